"""CPU port of the reference layers on stock ATen ops (TEST / BASELINE INFRASTRUCTURE).

These modules restate the reference's operator sequence with plain PyTorch so the
reference's CPU cost can be timed on the GPU box (where /root/reference does not
exist) and so host-side logic (harness model, data-parallel plumbing) can be
exercised on CPU under gloo.  They keep the reference's constructor signatures
and buffer names; autograd derives the backward exactly as it does for the
reference.  Unlike the reference they follow ``x.device`` instead of "CUDA if
available" (utils/whitening.py:23-24,48), so they run on CPU tensors in a process
that can see a GPU.

Only tests, ``__graft_entry__.smoke()`` and ``bench.py``'s reference / cpu_baseline
legs may import this.  Never the product package.

Restated lines: utils/whitening.py:7-61, utils/consensus_loss.py:6-24,
utils/batch_norm.py:17-89,157-160,229-232,302-305.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class WTransform2d(nn.Module):
    def __init__(self, num_features, group_size, running_m=None, running_var=None, momentum=0.1,
                 track_running_stats=True, eps=1e-3, alpha=1):
        super().__init__()
        self.num_features, self.momentum, self.eps, self.alpha = num_features, momentum, eps, alpha
        self.track_running_stats = track_running_stats
        self.group_size = min(num_features, group_size)
        self.num_groups = num_features // self.group_size
        self.running_m, self.running_var = running_m, running_var
        if track_running_stats and running_m is not None:
            self.register_buffer("running_mean", running_m)
            self.register_buffer("running_variance", running_var)
        else:
            self.register_buffer("running_mean", torch.zeros(1, num_features, 1, 1))
            self.register_buffer("running_variance",
                                 torch.ones(self.num_groups, self.group_size, self.group_size))

    def forward(self, x):
        if x.dim() != 4:
            raise ValueError("expected 4D input (got {}D input)".format(x.dim()))
        if self.num_features % self.group_size != 0:
            raise ValueError("expected number of channels divisible by group_size")
        g, gs = self.num_groups, self.group_size
        inference = (not self.training) and self.track_running_stats
        mu = x.mean(0).view(self.num_features, -1).mean(-1).view(1, -1, 1, 1)
        if inference:
            mu = self.running_mean
        xc = x - mu
        t = xc.transpose(0, 1).contiguous().view(g, gs, -1)
        cov = torch.bmm(t, t.transpose(1, 2)) / t.shape[-1]
        eye = torch.eye(gs, dtype=x.dtype, device=x.device).expand(g, gs, gs)
        src = self.running_variance if inference else cov
        shrunk = (1 - self.eps) * src + self.eps * eye
        w = torch.inverse(torch.linalg.cholesky(shrunk)).contiguous().view(self.num_features, gs, 1, 1)
        y = F.conv2d(xc, w, groups=g)
        if self.training and self.track_running_stats:
            with torch.no_grad():
                self.running_mean.mul_(1 - self.momentum).add_(mu.detach(), alpha=self.momentum)
                self.running_variance.mul_(1 - self.momentum).add_(cov.detach(), alpha=self.momentum)
        return y


class MinEntropyConsensusLoss(nn.Module):
    def __init__(self, num_classes, device):
        super().__init__()
        self.num_classes, self.device = num_classes, device

    def forward(self, x, y):
        s = -0.5 * (F.log_softmax(x, dim=1) + F.log_softmax(y, dim=1))
        return s.min(dim=1)[0].mean()


class _BatchNorm(nn.Module):
    _version = 2
    _dims = ()

    def __init__(self, num_features, running_m, running_v, eps=1e-5, momentum=0.1, affine=True,
                 track_running_stats=True):
        super().__init__()
        self.num_features, self.eps, self.momentum, self.affine = num_features, eps, momentum, affine
        self.running_m, self.running_v = running_m, running_v
        self.track_running_stats = track_running_stats
        if affine:
            self.weight = nn.Parameter(torch.empty(num_features).uniform_())
            self.bias = nn.Parameter(torch.zeros(num_features))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        if track_running_stats:
            self.register_buffer("running_mean", running_m)
            self.register_buffer("running_var", running_v)
            self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        else:
            self.register_parameter("running_mean", None)
            self.register_parameter("running_var", None)
            self.register_parameter("num_batches_tracked", None)

    def forward(self, x):
        if x.dim() not in self._dims:
            raise ValueError("expected {} input (got {}D input)".format(
                " or ".join("%dD" % d for d in self._dims), x.dim()))
        factor = 0.0
        if self.training and self.track_running_stats:
            self.num_batches_tracked += 1
            factor = 1.0 / self.num_batches_tracked.item() if self.momentum is None else self.momentum
        return F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias,
                            self.training or not self.track_running_stats, factor, self.eps)


class BatchNorm1d(_BatchNorm):
    _dims = (2, 3)


class BatchNorm2d(_BatchNorm):
    _dims = (4,)


class BatchNorm3d(_BatchNorm):
    _dims = (5,)
