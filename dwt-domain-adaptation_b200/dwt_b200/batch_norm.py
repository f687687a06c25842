"""Drop-in for the reference's ``utils/batch_norm.py``: batch norm whose running statistics are
tensors owned by the caller (shared by the three domain branches in the ResNet script), on the
group-size-1 member of the whitening kernel family.

Reference: /root/reference/utils/batch_norm.py:14-305.
"""
from __future__ import annotations

import torch
from torch import nn

from . import functional as F

_STAT_NAMES = ("running_mean", "running_var", "num_batches_tracked")


class _BatchNorm(nn.Module):
    _version = 2                       # batch_norm.py:15: version 2 added the step counter to the state dict
    _ranks: tuple = ()                 # accepted input ranks, set by the 1d / 2d / 3d subclasses
    _rank_text = ""

    def __init__(self, num_features, running_m, running_v, eps=1e-5, momentum=0.1, affine=True,
                 track_running_stats=True):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.affine, self.track_running_stats = affine, track_running_stats
        self.running_m, self.running_v = running_m, running_v
        for name in ("weight", "bias"):                               # learnable scale/shift or explicit Nones
            self.register_parameter(name, nn.Parameter(torch.empty(num_features)) if affine else None)
        if track_running_stats:                                       # batch_norm.py:32-35: the caller's tensors
            for name, tensor in zip(_STAT_NAMES, (running_m, running_v, torch.zeros((), dtype=torch.long))):
                self.register_buffer(name, tensor)
        else:
            for name in _STAT_NAMES:
                self.register_parameter(name, None)
        self.reset_parameters()

    def reset_running_stats(self):
        # the statistics belong to the caller: only the step counter is reset (batch_norm.py:42-44)
        if self.track_running_stats:
            self.num_batches_tracked.zero_()

    def reset_parameters(self):
        self.reset_running_stats()
        if self.affine:                                               # batch_norm.py:46-50
            nn.init.uniform_(self.weight)
            nn.init.zeros_(self.bias)

    def _check_input_dim(self, input):
        if not self._ranks:
            raise NotImplementedError
        if input.dim() not in self._ranks:
            raise ValueError("expected {} input (got {}D input)".format(self._rank_text, input.dim()))

    def forward(self, input):
        self._check_input_dim(input)
        tracking, factor = self.track_running_stats, 0.0
        if self.training and tracking:                                # batch_norm.py:57-64
            self.num_batches_tracked += 1
            # momentum None = cumulative moving average over the batches seen so far
            factor = self.momentum if self.momentum is not None else 1.0 / self.num_batches_tracked.item()
        batch_stats = self.training or not tracking
        if batch_stats and input.numel() // input.shape[1] <= 1:      # F.batch_norm's own guard
            raise ValueError("Expected more than 1 value per channel when training, got input size {}".format(
                tuple(input.shape)))
        return F.norm(input, self.weight, self.bias, kind="bn", group_size=1, n_domains=1,
                      training_stats=batch_stats, eps=self.eps, momentum=factor,
                      update_running=self.training and tracking,
                      running=[(self.running_mean, self.running_var)] if tracking else [(None, None)])

    def extra_repr(self):
        return (f"{self.num_features}, eps={self.eps}, momentum={self.momentum}, affine={self.affine}, "
                f"track_running_stats={self.track_running_stats}")

    def _load_from_state_dict(self, state_dict, prefix, metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        # checkpoints older than version 2 have no step counter: start it at zero instead of failing the load
        counter = prefix + _STAT_NAMES[2]
        if self.track_running_stats and (metadata.get("version") or 0) < 2 and counter not in state_dict:
            state_dict[counter] = torch.zeros((), dtype=torch.long)
        super()._load_from_state_dict(state_dict, prefix, metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)


class BatchNorm1d(_BatchNorm):         # batch_norm.py:157-160
    _ranks, _rank_text = (2, 3), "2D or 3D"


class BatchNorm2d(_BatchNorm):         # batch_norm.py:229-232
    _ranks, _rank_text = (4,), "4D"


class BatchNorm3d(_BatchNorm):         # batch_norm.py:302-305
    _ranks, _rank_text = (5,), "5D"
