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


def _check_param(name, t, numel):
    """The C ABI takes raw pointers: a strided or mis-sized statistics / affine tensor would be read or written out
    of bounds on the device, where the reference raises a shape error.  Validate before taking data_ptr()."""
    if t is None:
        return
    if t.numel() != numel:
        raise ValueError(f"{name} has {t.numel()} elements, expected {numel}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous (it is written / read through a raw pointer)")


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
        for d, (rm_t, rv_t) in enumerate(running):
            _check_param(f"running mean of domain {d}", rm_t, c)
            _check_param(f"running second moment of domain {d}", rv_t, c * gs)
        _check_param("gamma / weight", gamma, c)
        _check_param("beta / bias", beta, c)
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
        # residual tail on the channels-last kernels: the apply pass leaves one byte per float4 with the four
        # (out > 0) bits, which is all the backward needs of the output
        mask = torch.empty(x.numel() // 4, dtype=torch.uint8, device=dev) if (residual is not None and nhwc) else None
        save_mean = torch.empty(n_domains, c, dtype=torch.float32, device=dev)
        save_w = torch.empty(n_domains, c // gs, gs, gs, dtype=torch.float32, device=dev)
        ws = nv.workspace(dev, n, c, hw, gs, n_domains)
        need_running = (mode == nv.MODE_EVAL) or update_running
        rm = nv.ptr_array([p[0] for p in running]) if need_running else None
        rv = nv.ptr_array([p[1] for p in running]) if need_running else None
        with torch.cuda.device(dev):
            if kind == "whiten":
                rc = lib.dwt_whiten_fwd(nv.ptr(x), nv.ptr(y), n, c, hw, gs, n_domains, mode | layout, eps, momentum,
                                        int(update_running), rm, rv, nv.ptr(gamma_c), nv.ptr(beta_c), nv.ptr(residual),
                                        nv.ptr(mask), epi, nv.ptr(save_mean), nv.ptr(save_w), nv.ptr(ws), ws.numel(),
                                        nv.stream_ptr(dev))
            else:
                rc = lib.dwt_bn_fwd(nv.ptr(x), nv.ptr(y), n, c, hw, n_domains, mode | layout, eps, momentum,
                                    int(update_running), rm, rv, nv.ptr(gamma_c), nv.ptr(beta_c), nv.ptr(residual),
                                    nv.ptr(mask), epi, nv.ptr(save_mean), nv.ptr(save_w), nv.ptr(ws), ws.numel(),
                                    nv.stream_ptr(dev))
        nv.check(rc)
        nv.poll_status(dev)
        if update_running and mode == nv.MODE_TRAIN:
            # the kernels wrote the running buffers in place behind autograd's back: bump their version counters so a
            # graph that saved one of them notices (the reference's in-place EMA, whitening.py:58-59, does the same)
            seen = set()
            for pair in running:
                for buf in pair:
                    if buf is not None and id(buf) not in seen:
                        seen.add(id(buf))
                        torch.autograd.graph.increment_version(buf)
        # backward of relu(z + residual): dz = dout * (out > 0) is also the residual's gradient
        ctx.residual_mode = None
        if mask is not None:
            # channels-last: both backward kernels mask dout with the saved bits, bwd_apply also writes dz
            ctx.save_for_backward(x, save_mean, save_w, gamma_c, beta_c, mask)
            ctx.residual_mode = "mask"
        elif residual is not None:
            # NCHW: dz is formed by one ATen pass from the saved output; the kernels then run the plain affine epilogue
            ctx.save_for_backward(x, save_mean, save_w, gamma_c, beta_c, y)
            epi = nv.EPI_AFFINE
            ctx.residual_mode = "aten"
        else:
            ctx.save_for_backward(x, save_mean, save_w, gamma_c, beta_c)
        ctx.cfg = (kind, gs, n_domains, mode | layout, eps, epi, n, c, hw, None if gamma is None else gamma.shape)
        return y

    @staticmethod
    def backward(ctx, dout):
        lib = nv.lib()
        mask = None
        if ctx.residual_mode == "mask":
            x, save_mean, save_w, gamma_c, beta_c, mask = ctx.saved_tensors
        elif ctx.residual_mode == "aten":
            x, save_mean, save_w, gamma_c, beta_c, out = ctx.saved_tensors
            extra = ctx.__dict__.pop("_dwt_extra_grad", None)
            if extra is not None:
                dout = dout + extra
            dout = torch.ops.aten.threshold_backward(dout, out, 0)
        else:
            x, save_mean, save_w, gamma_c, beta_c = ctx.saved_tensors
        kind, gs, n_domains, mode, eps, epi, n, c, hw, gshape = ctx.cfg
        # second addend of the incoming gradient, left here by fork_for_sum's backward (see there): the channels-last
        # kernels add it where they read dout; any other path adds it now
        dout2 = ctx.__dict__.pop("_dwt_extra_grad", None)
        if dout2 is not None and not ((mode & nv.LAYOUT_NHWC) and dout2.shape == dout.shape and dout2.dtype == torch.float32
                                      and dout2.is_contiguous(memory_format=torch.channels_last)):
            dout, dout2 = dout + dout2, None
        dout = dout.contiguous(memory_format=torch.channels_last) if (mode & nv.LAYOUT_NHWC) else dout.contiguous()
        dev = nv.require_cuda(dout)
        dx = torch.empty_like(x)
        want_affine = gamma_c is not None and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        d_res = None
        if ctx.needs_input_grad[3]:
            if ctx.residual_mode == "mask":
                d_res = torch.empty_like(x)
            elif ctx.residual_mode == "aten":
                d_res = dout
        dgamma = torch.empty(c, dtype=torch.float32, device=dev) if want_affine else None
        dbeta = torch.empty(c, dtype=torch.float32, device=dev) if want_affine else None
        ws = nv.workspace(dev, n, c, hw, gs, n_domains)
        with torch.cuda.device(dev):
            if kind == "whiten":
                rc = lib.dwt_whiten_bwd(nv.ptr(x), nv.ptr(dout), nv.ptr(dout2), nv.ptr(dx), n, c, hw, gs, n_domains, mode, eps,
                                        nv.ptr(save_mean), nv.ptr(save_w), nv.ptr(gamma_c), nv.ptr(beta_c), nv.ptr(mask),
                                        nv.ptr(d_res) if mask is not None else None, epi,
                                        nv.ptr(dgamma), nv.ptr(dbeta), nv.ptr(ws), ws.numel(), nv.stream_ptr(dev))
            else:
                rc = lib.dwt_bn_bwd(nv.ptr(x), nv.ptr(dout), nv.ptr(dout2), nv.ptr(dx), n, c, hw, n_domains, mode,
                                    nv.ptr(save_mean), nv.ptr(save_w), nv.ptr(gamma_c), nv.ptr(beta_c), nv.ptr(mask),
                                    nv.ptr(d_res) if mask is not None else None, epi,
                                    nv.ptr(dgamma), nv.ptr(dbeta), nv.ptr(ws), ws.numel(), nv.stream_ptr(dev))
        nv.check(rc)
        if want_affine:
            dgamma, dbeta = dgamma.view(gshape), dbeta.view(gshape)
        return (dx, dgamma, dbeta, d_res) + (None,) * 9


class _ForkForSum(torch.autograd.Function):
    """a, b = fork(y): two aliases of y whose gradients are NOT summed by autograd.  y must be the output of a
    _NormFunction node (the producer): backward hands the first gradient on as y's gradient and parks the second on the
    producer's node, whose backward passes it to the kernels as the second addend (dwt_whiten_bwd's dout2).  If y gets
    other gradients as well, autograd adds them to the first one as usual -- the parked addend is independent of that."""

    @staticmethod
    def forward(ctx, y):
        ctx.producer = y.grad_fn
        return y.view_as(y), y.view_as(y)

    @staticmethod
    def backward(ctx, ga, gb):
        producer, ctx.producer = ctx.producer, None
        if ga is None or gb is None:
            return ga if gb is None else gb
        if producer is None or "_dwt_extra_grad" in producer.__dict__:
            return ga + gb
        producer.__dict__["_dwt_extra_grad"] = gb
        return ga


def fork_for_sum(y):
    """Use the output of a fused site twice -- e.g. as the next Bottleneck's input AND its identity branch
    (resnet50_dwt_mec_officehome.py:217-240) -- without autograd's elementwise addition of the two gradients: returns
    (a, b), two aliases of y.  Falls back to (y, y) when y was not produced by one of this package's norm sites or no
    gradient is being recorded; results are identical either way."""
    fn = getattr(y, "grad_fn", None)
    if not torch.is_grad_enabled() or fn is None or not isinstance(fn, _NormFunction._backward_cls):
        return y, y
    return _ForkForSum.apply(y)


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
                                           nv.ptr(losses), nv.ptr(grad), nv.status_ptr(dev), nv.stream_ptr(dev))
        nv.check(rc)
        nv.poll_status(dev)
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
