"""Drop-in for the reference's ``utils/whitening.py`` (same class names, constructor and
forward signatures, attributes, buffers and error texts), computing on hand-written sm_100a
kernels through libdwt_b200.so.

Reference: /root/reference/utils/whitening.py:5-71.
Differences that are deliberate:
  * buffers default to the input's device lazily instead of "CUDA if available"
    (whitening.py:23-24) -- they are created on CPU like any nn.Module buffer and move with
    ``.to(device)``; externally owned buffers are registered as-is, never copied (aliasing
    across the three domain modules survives);
  * CUDA only: a CPU tensor raises (there is no CPU fallback by design).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as F


class _Whitening(nn.Module):
    def __init__(self, num_features, group_size, running_m=None, running_var=None, momentum=0.1,
                 track_running_stats=True, eps=1e-3, alpha=1):
        super().__init__()
        self.num_features = num_features
        self.momentum = momentum
        self.track_running_stats = track_running_stats
        self.eps = eps
        self.alpha = alpha
        self.group_size = min(self.num_features, group_size)
        self.num_groups = self.num_features // self.group_size
        self.running_m = running_m
        self.running_var = running_var
        if self.track_running_stats and self.running_m is not None:
            self.register_buffer("running_mean", self.running_m)
            self.register_buffer("running_variance", self.running_var)
        else:
            self.register_buffer("running_mean", torch.zeros(1, self.num_features, 1, 1))
            # an all-ones matrix per group, not the identity (whitening.py:24)
            self.register_buffer("running_variance",
                                 torch.ones(self.num_groups, self.group_size, self.group_size))

    def _check_input_dim(self, input):
        raise NotImplementedError

    def _check_group_size(self):
        raise NotImplementedError

    def forward(self, x):
        self._check_input_dim(x)
        self._check_group_size()
        inference = (not self.training) and self.track_running_stats
        # the reference updates the buffers in train mode even under no_grad and even when they
        # were default-constructed (whitening.py:57-59)
        update = self.training and self.track_running_stats
        return F.norm(x, None, None, kind="whiten", group_size=self.group_size, n_domains=1,
                      training_stats=not inference, eps=self.eps, momentum=self.momentum,
                      update_running=update, running=[(self.running_mean, self.running_variance)])


class WTransform2d(_Whitening):
    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError('expected 4D input (got {}D input)'.format(input.dim()))

    def _check_group_size(self):
        if self.num_features % self.group_size != 0:
            raise ValueError('expected number of channels divisible by group_size (got {} group_size\
				for {} number of features'.format(self.group_size, self.num_features))
