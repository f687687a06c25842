"""Fused domain-triple norm site (extension beyond the reference's module API).

The reference runs every norm site as
    split/3 -> bns(x_s) | bnt(x_t) | bnt_aug(x_t') -> cat -> cat -> * gamma + beta -> relu
(resnet50_dwt_mec_officehome.py:220-222,335-337): three module calls, two full-tensor
concatenations, an affine pass and a ReLU pass.  ``DomainTripleNorm`` does the whole site in two
tensor-sized kernel launches (statistics, apply; a small finalize launch between them) over the
un-split tensor, the three domains batched on grid.z, gamma/beta/ReLU (and the Bottleneck's residual
tail) folded into the apply pass, and the running-statistic EMA applied source -> target ->
target-aug in order so aliased buffers end exactly as after three sequential module calls
(SURVEY.md H5).  Backward is likewise two tensor-sized launches and also yields dgamma/dbeta and,
for the residual tail, the gradient of the identity branch.

It owns no state: it borrows the running buffers of the three domain modules at call time.

``replicated=True`` is the statistics-collection pass (SURVEY.md §8f-3;
resnet50_dwt_mec_officehome.py:380-389): the reference feeds ``cat((data, data, data))`` through
the network in train mode under ``no_grad`` so that every domain branch folds the target batch
into its running buffers.  The three thirds are identical, so are their statistics and their
outputs: here x is the single copy [N, C, H, W], the statistics are computed once, the output is
written once, and a buffer shared by k branches receives the k-fold EMA in one update,
r <- (1-m)^k r + (1 - (1-m)^k) s -- a third of the traffic at every site and a third of the
convolution work between them.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as F


class DomainTripleNorm(nn.Module):
    def __init__(self, kind, num_features, group_size=4, n_domains=3):
        super().__init__()
        if kind not in ("whiten", "bn"):
            raise ValueError("kind must be 'whiten' or 'bn'")
        self.kind, self.num_features, self.n_domains = kind, num_features, n_domains
        self.group_size = min(num_features, group_size) if kind == "whiten" else 1
        # group sizes 1, 2, 4: gamma / beta / ReLU / residual are folded into the apply kernels.  Larger groups (the
        # tensor-core kernels; e.g. ResNet(..., group_size=64), resnet50_dwt_mec_officehome.py:266): the three domains
        # still share ONE statistics + ONE apply launch with the ordered EMA, the shared affine / ReLU / residual
        # follow as plain tensor ops (two extra elementwise passes; autograd differentiates them).
        self.kernel_epilogue = self.group_size in (1, 2, 4)

    def forward(self, x, domain_modules, gamma, beta, relu=False, residual=None, replicated=False, count_batches=True):
        """x: [n_domains*N, C, H, W]; domain_modules: the per-domain WTransform2d / BatchNorm2d
        modules (training mode), whose buffers receive the EMA updates; gamma/beta: [C,1,1];
        residual (needs relu=True): out = relu(norm(x)*gamma + beta + residual), the Bottleneck tail
        (resnet50_dwt_mec_officehome.py:239-240) folded into the apply pass.
        count_batches=False: the caller has already done the `num_batches_tracked += 1` of the three BatchNorm
        modules (batch_norm.py:58) -- a model bumps all its counters with ONE multi-tensor launch per step instead of
        126 one-element kernels."""
        mods = list(domain_modules)
        if len(mods) != self.n_domains:
            raise ValueError(f"expected {self.n_domains} domain modules")
        if x.dim() != 4:
            raise ValueError('expected 4D input (got {}D input)'.format(x.dim()))
        if replicated:
            return self._forward_replicated(x, mods, gamma, beta, relu, residual, count_batches)
        m0 = mods[0]
        if self.kind == "whiten":
            running = [(m.running_mean, m.running_variance) for m in mods]
            eps, momentum = m0.eps, m0.momentum
        else:
            if count_batches:
                counters = [m.num_batches_tracked for m in mods if m.training and m.track_running_stats]
                if counters:
                    torch._foreach_add_(counters, 1)
            running = [(m.running_mean, m.running_var) for m in mods]
            eps = m0.eps
            momentum = m0.momentum if m0.momentum is not None else 1.0 / m0.num_batches_tracked.item()
        update = m0.training and m0.track_running_stats
        if not self.kernel_epilogue:
            y = F.norm(x, None, None, kind=self.kind, group_size=self.group_size, n_domains=self.n_domains,
                       training_stats=True, eps=eps, momentum=momentum, update_running=update, running=running)
            return self._tensor_epilogue(y, gamma, beta, relu, residual)
        return F.norm(x, gamma, beta, kind=self.kind, group_size=self.group_size, n_domains=self.n_domains,
                      training_stats=True, eps=eps, momentum=momentum, update_running=update,
                      running=running, relu=relu, residual=residual)

    @staticmethod
    def _tensor_epilogue(y, gamma, beta, relu, residual):
        if residual is not None and not relu:
            raise ValueError("a fused residual needs relu=True")
        if gamma is not None:
            y = y * gamma + beta
        if residual is not None:
            y = y + residual
        return torch.relu(y) if relu else y

    def _forward_replicated(self, x, mods, gamma, beta, relu, residual, count_batches=True):
        """One copy of the batch stands for all n_domains branches (see the module docstring)."""
        second = "running_variance" if self.kind == "whiten" else "running_var"
        keep = {}                                  # distinct buffer pair -> product of (1 - factor) over its branches
        for m in mods:
            if not (m.training and m.track_running_stats):
                continue
            if self.kind == "bn":
                if count_batches:
                    m.num_batches_tracked += 1
                f = m.momentum if m.momentum is not None else 1.0 / m.num_batches_tracked.item()
            else:
                f = m.momentum
            rm, rv = m.running_mean, getattr(m, second)
            key = (rm.data_ptr(), rv.data_ptr())
            prod, _ = keep.get(key, (1.0, None))
            keep[key] = (prod * (1.0 - f), (rm, rv))
        means = {k[0] for k in keep}
        seconds = {k[1] for k in keep}
        if len(means) != len(keep) or len(seconds) != len(keep):
            raise ValueError("replicated statistics need each running_mean paired with one second-moment buffer")
        m0 = mods[0]
        common = dict(kind=self.kind, group_size=self.group_size, n_domains=1, training_stats=True, eps=m0.eps)
        if self.kernel_epilogue:
            common.update(relu=relu, residual=residual)
            g_arg, b_arg = gamma, beta
        else:
            g_arg = b_arg = None
        if not keep:
            out = F.norm(x, g_arg, b_arg, momentum=0.0, update_running=False,
                         running=[(m0.running_mean, getattr(m0, second))], **common)
        else:
            out = None
            for prod, pair in keep.values():       # one launch per distinct buffer set (one, in a loaded model)
                y = F.norm(x, g_arg, b_arg, momentum=1.0 - prod, update_running=True, running=[pair], **common)
                out = y if out is None else out
        return out if self.kernel_epilogue else self._tensor_epilogue(out, gamma, beta, relu, residual)
