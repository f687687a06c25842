"""CPU oracle for the DWT hot path (TEST INFRASTRUCTURE, never the product path).

A numpy restatement of the three reference operators, written so that every
function can run in float64 (to judge both the CUDA kernels and the fp32
reference against something tighter than either) or float32.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import this module.  The product package (``dwt-domain-adaptation_b200/``)
must never import it: it has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the reference itself, imported in the
build container by ``tests/golden/make_golden.py`` and committed as
``tests/golden/*.npz``; ``tests/test_oracle_vs_golden.py`` checks every fixture.

Reference lines restated (paths relative to /root/reference):
  * whitening forward ....... utils/whitening.py:37-61
  * whitening buffers init .. utils/whitening.py:19-24
  * whitening backward ...... autograd through utils/whitening.py:41-55, in the
                              closed form of SURVEY.md §8(a)
  * MEC loss ................ utils/consensus_loss.py:11-24
  * domain batch norm ....... utils/batch_norm.py:54-69 (F.batch_norm semantics)
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------- #
# whitening
# --------------------------------------------------------------------------- #
def _grouped(xc: np.ndarray, gs: int) -> np.ndarray:
    """[N,C,H,W] -> [G,gs,M] with M = N*H*W (utils/whitening.py:46)."""
    n, c = xc.shape[:2]
    return np.ascontiguousarray(np.moveaxis(xc.reshape(n, c, -1), 1, 0)).reshape(c // gs, gs, -1)


def whiten_stats(x: np.ndarray, gs: int):
    """Per-channel mean and per-group biased covariance (utils/whitening.py:41-47)."""
    n, c = x.shape[:2]
    mean = x.reshape(n, c, -1).mean(axis=(0, 2))
    t = _grouped(x - mean.reshape(1, c, 1, 1), gs)
    cov = t @ np.swapaxes(t, 1, 2) / t.shape[-1]
    return mean, cov


def whiten_matrix(cov: np.ndarray, eps: float) -> np.ndarray:
    """W = inverse(cholesky((1-eps)*cov + eps*I)), lower-triangular (utils/whitening.py:48,53)."""
    gs = cov.shape[-1]
    s = (1.0 - eps) * cov + eps * np.eye(gs, dtype=cov.dtype)
    return np.linalg.inv(np.linalg.cholesky(s))


def whiten_apply(x: np.ndarray, mean: np.ndarray, w: np.ndarray) -> np.ndarray:
    """y = W (x - mean), group by group (the grouped 1x1 conv of utils/whitening.py:55)."""
    n, c = x.shape[:2]
    g, gs, _ = w.shape
    xc = (x - mean.reshape(1, c, 1, 1)).reshape(n, g, gs, -1)
    y = np.einsum("gij,ngjm->ngim", w, xc)
    return y.reshape(x.shape)


def whiten_forward(x, gs, eps=1e-3, momentum=0.1, running_mean=None, running_cov=None,
                   training=True, track_running_stats=True):
    """Full forward of utils/whitening.py:37-61.

    Returns (y, mean_used, w, new_running_mean, new_running_cov, batch_cov).
    ``running_mean`` is [C] here (the reference stores it as [1,C,1,1]);
    ``running_cov`` is [G,gs,gs].  The EMA keeps the UN-shrunk covariance (:59).
    """
    c = x.shape[1]
    gs = min(c, gs)
    bmean, bcov = whiten_stats(x, gs)
    use_running = (not training) and track_running_stats
    mean = running_mean if use_running else bmean
    if use_running:
        # eval: the batch covariance is computed and discarded (:47 vs :50-51)
        w = whiten_matrix(running_cov, eps)
    else:
        w = whiten_matrix(bcov, eps)
    y = whiten_apply(x, mean, w)
    new_rm, new_rc = running_mean, running_cov
    if training and track_running_stats and running_mean is not None:
        new_rm = momentum * bmean + (1.0 - momentum) * running_mean
        new_rc = momentum * bcov + (1.0 - momentum) * running_cov
    return y, mean, w, new_rm, new_rc, bcov


def whiten_backward(x, dy, mean, w, eps=1e-3):
    """Closed-form d(loss)/dx for the TRAIN-mode forward (mean and W from the batch).

    With xc = x - mean, M = N*H*W, per group:
        dW = dy xc^T                      Q = -dW W^T
        P  = tril(Q) with halved diagonal S = sym(W^T P W)
        dx = W^T (dy - mean_M dy) + (2 (1-eps) / M) S xc
    (autograd through utils/whitening.py:41-55).
    """
    n, c = x.shape[:2]
    g, gs, _ = w.shape
    xc = _grouped(x - mean.reshape(1, c, 1, 1), gs)          # [G,gs,M]
    gy = _grouped(dy, gs)
    m = xc.shape[-1]
    dw = gy @ np.swapaxes(xc, 1, 2)
    q = -dw @ np.swapaxes(w, 1, 2)
    p = np.tril(q)
    idx = np.arange(gs)
    p[:, idx, idx] *= 0.5
    s = np.swapaxes(w, 1, 2) @ p @ w
    s = 0.5 * (s + np.swapaxes(s, 1, 2))
    gbar = gy.mean(axis=2, keepdims=True)
    dxg = np.swapaxes(w, 1, 2) @ (gy - gbar) + (2.0 * (1.0 - eps) / m) * (s @ xc)
    hw = x.shape[2:]
    return np.ascontiguousarray(np.moveaxis(dxg.reshape(c, n, *hw), 0, 1))


def whiten_backward_eval(dy, w):
    """Eval-mode backward: mean and W are constants, dx = W^T dy."""
    n, c = dy.shape[:2]
    g, gs, _ = w.shape
    gy = dy.reshape(n, g, gs, -1)
    return np.einsum("gji,ngjm->ngim", w, gy).reshape(dy.shape)


def scale_shift_relu(y, gamma, beta, relu):
    """out = y*gamma + beta (resnet50_dwt_mec_officehome.py:59-63,221-222), optional ReLU."""
    c = y.shape[1]
    out = y * gamma.reshape(1, c, 1, 1) + beta.reshape(1, c, 1, 1)
    return np.maximum(out, 0) if relu else out


# --------------------------------------------------------------------------- #
# Min-Entropy-Consensus loss
# --------------------------------------------------------------------------- #
def _log_softmax(a):
    z = a - a.max(axis=1, keepdims=True)
    return z - np.log(np.exp(z).sum(axis=1, keepdims=True))


def mec_loss(x, y):
    """loss = mean_n min_k -(lsm(x)+lsm(y))[n,k]/2 (utils/consensus_loss.py:13-22).

    Returns (loss, dloss/dx, dloss/dy, argmin_k).  Ties resolve to the first
    minimum like torch.min (the gradient then flows to that single class).
    """
    lx, ly = _log_softmax(x), _log_softmax(y)
    s = -0.5 * (lx + ly)
    k = s.argmin(axis=1)
    n = x.shape[0]
    loss = s[np.arange(n), k].mean()
    onehot = np.zeros_like(x)
    onehot[np.arange(n), k] = 1.0
    gx = (np.exp(lx) - onehot) / (2.0 * n)
    gy = (np.exp(ly) - onehot) / (2.0 * n)
    return loss, gx, gy, k


# --------------------------------------------------------------------------- #
# domain batch norm (externally owned running stats)
# --------------------------------------------------------------------------- #
def bn_forward(x, running_mean, running_var, weight=None, bias=None, training=True,
               factor=0.1, eps=1e-5):
    """F.batch_norm as called at utils/batch_norm.py:66-69.

    x is [N,C,*].  Batch variance is biased for normalisation and unbiased into
    running_var.  Returns (y, mean_used, invstd, new_running_mean, new_running_var).
    """
    n, c = x.shape[:2]
    xr = x.reshape(n, c, -1)
    m = n * xr.shape[2]
    shape = (1, c) + (1,) * (x.ndim - 2)
    if training:
        mean = xr.mean(axis=(0, 2))
        var = xr.var(axis=(0, 2))
        new_rm, new_rv = running_mean, running_var
        if running_mean is not None:
            new_rm = (1 - factor) * running_mean + factor * mean
            new_rv = (1 - factor) * running_var + factor * var * (m / max(m - 1, 1))
    else:
        mean, var = running_mean, running_var
        new_rm, new_rv = running_mean, running_var
    invstd = 1.0 / np.sqrt(var + eps)
    y = (x - mean.reshape(shape)) * invstd.reshape(shape)
    if weight is not None:
        y = y * weight.reshape(shape)
    if bias is not None:
        y = y + bias.reshape(shape)
    return y, mean, invstd, new_rm, new_rv


def bn_backward(x, dy, mean, invstd, weight=None, training=True):
    """Returns (dx, dweight, dbias) for bn_forward."""
    n, c = x.shape[:2]
    shape = (1, c) + (1,) * (x.ndim - 2)
    axes = (0,) + tuple(range(2, x.ndim))
    xhat = (x - mean.reshape(shape)) * invstd.reshape(shape)
    dbias = dy.sum(axis=axes)
    dweight = (dy * xhat).sum(axis=axes)
    w = np.ones(c, x.dtype) if weight is None else weight
    scale = (w * invstd).reshape(shape)
    if training:
        m = x.size // c
        dx = scale * (dy - (dbias / m).reshape(shape) - xhat * (dweight / m).reshape(shape))
    else:
        dx = scale * dy
    return dx, dweight, dbias
