"""autograd Functions over the C ABI: one forward call + one hand-derived backward call each.

Saved for backward: the layer input x plus the tiny per-group statistics (mean, W) -- never a
centred copy, a transposed copy or the covariance graph the reference's autograd keeps
(utils/whitening.py:44-55).
"""
from __future__ import annotations

import torch

from . import _native as nv


def _dense(x: torch.Tensor, group_size: int):
    """[N, C, *spatial] -> (dense tensor, N, C, HW, channels_last?).

    A 4-D tensor that is already dense in torch.channels_last order is used as it is (NHWC kernels) when
    the geometry has a channels-last build; anything else is made NCHW-contiguous."""
    n, c = x.shape[0], x.shape[1]
    hw = 1
    for s in x.shape[2:]:
        hw *= s
    nhwc = (x.dim() == 4 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last)
            and nv.channels_last_supported(c, group_size))
    if not nhwc and not x.is_contiguous():
        x = x.contiguous()
    return x, n, c, hw, nhwc


class _NormFunction(torch.autograd.Function):
    """Shared by whitening (kind='whiten') and domain batch norm (kind='bn').

    x is [n_domains*N, C, *]; `running` is a list of n_domains (mean, second-moment) buffer pairs
    (entries may alias); gamma/beta are [C]-sized or None; relu fuses max(.,0) behind the affine.
    """

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, kind, group_size, n_domains, mode, eps, momentum, update_running,
                running, relu):
        lib = nv.lib()
        gs = group_size if kind == "whiten" else 1
        x, n_all, c, hw, nhwc = _dense(x, gs)
        layout = nv.LAYOUT_NHWC if nhwc else 0
        if n_all % n_domains != 0:
            raise ValueError(f"batch of {n_all} does not split into {n_domains} domains")
        n = n_all // n_domains
        dev = nv.require_cuda(x, gamma, beta, residual, *[t for pair in running for t in pair])
        epi = nv.EPI_NONE
        if residual is not None:
            if gamma is None or not relu or residual.shape != x.shape:
                raise ValueError("a fused residual needs gamma/beta, relu=True and a tensor shaped like x")
            residual = residual.contiguous(memory_format=torch.channels_last) if nhwc else residual.contiguous()
        if gamma is not None:
            epi = nv.EPI_AFFINE | (nv.EPI_RELU if relu else 0) | (nv.EPI_RESIDUAL if residual is not None else 0)
            gamma_c, beta_c = gamma.detach().reshape(-1).contiguous(), beta.detach().reshape(-1).contiguous()
        else:
            gamma_c = beta_c = None
        y = torch.empty_like(x)                      # keeps x's memory format
        save_mean = torch.empty(n_domains, c, dtype=torch.float32, device=dev)
        save_w = torch.empty(n_domains, c // gs, gs, gs, dtype=torch.float32, device=dev)
        ws = nv.workspace(dev, n, c, hw, gs, n_domains)
        need_running = (mode == nv.MODE_EVAL) or update_running
        rm = nv.ptr_array([p[0] for p in running]) if need_running else None
        rv = nv.ptr_array([p[1] for p in running]) if need_running else None
        with torch.cuda.device(dev):
            if kind == "whiten":
                rc = lib.dwt_whiten_fwd(nv.ptr(x), nv.ptr(y), n, c, hw, gs, n_domains, mode | layout, eps, momentum,
                                        int(update_running), rm, rv, nv.ptr(gamma_c), nv.ptr(beta_c), nv.ptr(residual), epi,
                                        nv.ptr(save_mean), nv.ptr(save_w), nv.ptr(ws), ws.numel(), nv.stream_ptr(dev))
            else:
                rc = lib.dwt_bn_fwd(nv.ptr(x), nv.ptr(y), n, c, hw, n_domains, mode | layout, eps, momentum,
                                    int(update_running), rm, rv, nv.ptr(gamma_c), nv.ptr(beta_c), nv.ptr(residual), epi,
                                    nv.ptr(save_mean), nv.ptr(save_w), nv.ptr(ws), ws.numel(), nv.stream_ptr(dev))
        nv.check(rc)
        if residual is not None:
            # backward of relu(z + residual): dz = dout * (out > 0) is also the residual's gradient; the norm's own
            # backward then runs on dz with the plain affine epilogue -- same bytes as masking inside the kernels
            ctx.save_for_backward(x, save_mean, save_w, gamma_c, beta_c, y)
            epi = nv.EPI_AFFINE
        else:
            ctx.save_for_backward(x, save_mean, save_w, gamma_c, beta_c)
        ctx.has_residual = residual is not None
        ctx.cfg = (kind, gs, n_domains, mode | layout, eps, epi, n, c, hw, None if gamma is None else gamma.shape)
        return y

    @staticmethod
    def backward(ctx, dout):
        lib = nv.lib()
        if ctx.has_residual:
            x, save_mean, save_w, gamma_c, beta_c, out = ctx.saved_tensors
            dout = torch.ops.aten.threshold_backward(dout, out, 0)
        else:
            x, save_mean, save_w, gamma_c, beta_c = ctx.saved_tensors
        kind, gs, n_domains, mode, eps, epi, n, c, hw, gshape = ctx.cfg
        dout = dout.contiguous(memory_format=torch.channels_last) if (mode & nv.LAYOUT_NHWC) else dout.contiguous()
        dev = nv.require_cuda(dout)
        dx = torch.empty_like(x)
        want_affine = gamma_c is not None and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        d_res = dout if (ctx.has_residual and ctx.needs_input_grad[3]) else None
        dgamma = torch.empty(c, dtype=torch.float32, device=dev) if want_affine else None
        dbeta = torch.empty(c, dtype=torch.float32, device=dev) if want_affine else None
        ws = nv.workspace(dev, n, c, hw, gs, n_domains)
        with torch.cuda.device(dev):
            if kind == "whiten":
                rc = lib.dwt_whiten_bwd(nv.ptr(x), nv.ptr(dout), nv.ptr(dx), n, c, hw, gs, n_domains, mode, eps,
                                        nv.ptr(save_mean), nv.ptr(save_w), nv.ptr(gamma_c), nv.ptr(beta_c), epi,
                                        nv.ptr(dgamma), nv.ptr(dbeta), nv.ptr(ws), ws.numel(), nv.stream_ptr(dev))
            else:
                rc = lib.dwt_bn_bwd(nv.ptr(x), nv.ptr(dout), nv.ptr(dx), n, c, hw, n_domains, mode,
                                    nv.ptr(save_mean), nv.ptr(save_w), nv.ptr(gamma_c), nv.ptr(beta_c), epi,
                                    nv.ptr(dgamma), nv.ptr(dbeta), nv.ptr(ws), ws.numel(), nv.stream_ptr(dev))
        nv.check(rc)
        if want_affine:
            dgamma, dbeta = dgamma.view(gshape), dbeta.view(gshape)
        return (dx, dgamma, dbeta, d_res) + (None,) * 9


def norm(x, gamma, beta, *, kind, group_size, n_domains, training_stats, eps, momentum, update_running,
         running, relu=False, residual=None):
    mode = nv.MODE_TRAIN if training_stats else nv.MODE_EVAL
    return _NormFunction.apply(x, gamma, beta, residual, kind, group_size, n_domains, mode, float(eps),
                               float(momentum), bool(update_running), running, bool(relu))


class _MecFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        lib = nv.lib()
        if x.dim() != 2 or x.shape != y.shape:
            raise ValueError(f"expected two [N, K] logit tensors, got {tuple(x.shape)} and {tuple(y.shape)}")
        x, y = x.contiguous(), y.contiguous()
        dev = nv.require_cuda(x, y)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        gx, gy = torch.empty_like(x), torch.empty_like(y)
        with torch.cuda.device(dev):
            rc = lib.dwt_mec_fwd_bwd(nv.ptr(x), nv.ptr(y), x.shape[0], x.shape[1], nv.ptr(loss), nv.ptr(gx),
                                     nv.ptr(gy), nv.stream_ptr(dev))
        nv.check(rc)
        ctx.save_for_backward(gx, gy)
        return loss

    @staticmethod
    def backward(ctx, g):
        gx, gy = ctx.saved_tensors
        return g * gx, g * gy


def mec_loss(x, y):
    return _MecFunction.apply(x, y)


class _HeadLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, lam):
        lib = nv.lib()
        if logits.dim() != 2 or logits.shape[0] % 3 != 0 or labels.dim() != 1 or labels.shape[0] * 3 != logits.shape[0]:
            raise ValueError(f"expected logits [3B, K] and labels [B], got {tuple(logits.shape)} and {tuple(labels.shape)}")
        logits = logits.contiguous()
        dev = nv.require_cuda(logits)
        if not labels.is_cuda or labels.dtype != torch.int64:
            raise nv.NativeError("labels must be an int64 CUDA tensor")
        labels = labels.contiguous()
        losses = torch.empty(3, dtype=torch.float32, device=dev)
        grad = torch.empty_like(logits)
        with torch.cuda.device(dev):
            rc = lib.dwt_head_loss_fwd_bwd(nv.ptr(logits), nv.ptr(labels), labels.shape[0], logits.shape[1], float(lam),
                                           nv.ptr(losses), nv.ptr(grad), nv.stream_ptr(dev))
        nv.check(rc)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(losses)
        return losses[0], losses

    @staticmethod
    def backward(ctx, g, _unused):
        (grad,) = ctx.saved_tensors
        return g * grad, None, None


def head_loss(logits, labels, lambda_mec):
    """-> (total loss (differentiable), tensor [total, classification, lambda*MEC])."""
    return _HeadLossFunction.apply(logits, labels, lambda_mec)
