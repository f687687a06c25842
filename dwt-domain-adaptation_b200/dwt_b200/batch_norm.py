"""Drop-in for the reference's ``utils/batch_norm.py``: batch norm whose running statistics are
tensors owned by the caller (shared by the three domain branches in the ResNet script), on the
group-size-1 member of the whitening kernel family.

Reference: /root/reference/utils/batch_norm.py:14-305.
"""
from __future__ import annotations

import torch
import torch.nn.init as init
from torch.nn.modules.module import Module
from torch.nn.parameter import Parameter

from . import functional as F


class _BatchNorm(Module):
    _version = 2

    def __init__(self, num_features, running_m, running_v, eps=1e-5, momentum=0.1, affine=True,
                 track_running_stats=True):
        super().__init__()
        self.num_features = num_features
        self.eps = eps
        self.momentum = momentum
        self.affine = affine
        self.running_m = running_m
        self.running_v = running_v
        self.track_running_stats = track_running_stats
        if self.affine:
            self.weight = Parameter(torch.Tensor(num_features))
            self.bias = Parameter(torch.Tensor(num_features))
        else:
            self.register_parameter('weight', None)
            self.register_parameter('bias', None)
        if self.track_running_stats:
            self.register_buffer('running_mean', self.running_m)
            self.register_buffer('running_var', self.running_v)
            self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))
        else:
            self.register_parameter('running_mean', None)
            self.register_parameter('running_var', None)
            self.register_parameter('num_batches_tracked', None)
        self.reset_parameters()

    def reset_running_stats(self):
        # the statistics belong to the caller: only the step counter is reset (batch_norm.py:42-44)
        if self.track_running_stats:
            self.num_batches_tracked.zero_()

    def reset_parameters(self):
        self.reset_running_stats()
        if self.affine:
            init.uniform_(self.weight)
            init.zeros_(self.bias)

    def _check_input_dim(self, input):
        raise NotImplementedError

    def forward(self, input):
        self._check_input_dim(input)
        factor = 0.0
        if self.training and self.track_running_stats:
            self.num_batches_tracked += 1
            if self.momentum is None:      # cumulative moving average
                factor = 1.0 / self.num_batches_tracked.item()
            else:
                factor = self.momentum
        batch_stats = self.training or not self.track_running_stats
        if batch_stats and input.numel() // input.shape[1] <= 1:
            raise ValueError('Expected more than 1 value per channel when training, got input size {}'.format(
                tuple(input.shape)))
        update = self.training and self.track_running_stats
        return F.norm(input, self.weight, self.bias, kind="bn", group_size=1, n_domains=1,
                      training_stats=batch_stats, eps=self.eps, momentum=factor, update_running=update,
                      running=[(self.running_mean, self.running_var)] if self.track_running_stats else [(None, None)])

    def extra_repr(self):
        return '{num_features}, eps={eps}, momentum={momentum}, affine={affine}, ' \
               'track_running_stats={track_running_stats}'.format(**self.__dict__)

    def _load_from_state_dict(self, state_dict, prefix, metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        version = metadata.get('version', None)
        if (version is None or version < 2) and self.track_running_stats:
            key = prefix + 'num_batches_tracked'
            if key not in state_dict:
                state_dict[key] = torch.tensor(0, dtype=torch.long)
        super()._load_from_state_dict(state_dict, prefix, metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)


class BatchNorm1d(_BatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 2 and input.dim() != 3:
            raise ValueError('expected 2D or 3D input (got {}D input)'.format(input.dim()))


class BatchNorm2d(_BatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError('expected 4D input (got {}D input)'.format(input.dim()))


class BatchNorm3d(_BatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 5:
            raise ValueError('expected 5D input (got {}D input)'.format(input.dim()))
