"""CPU emulation of the split-TF32 apply GEMM (csrc/norm_tc_apply.cu) against fp64: how far from the fp32 reference
does y = W (x - mu) land as the covariance gets ill-conditioned and the data moves away from zero?

  A (matrix)  hi = RN_tf32(W), lo = RN_tf32(W - hi)
  B (tile)    hi = trunc_tf32(x)  (what the tensor core reads from the raw fp32 words), lo = RN_tf32(x - hi)
  y = A_hi B_hi + A_lo B_hi + A_hi B_lo + A_lo B_lo  - (W mu)        products exact, fp32 accumulation emulated two ways:
        'f64acc'  one rounding at the end (best case),  'f32seq' sequential fp32 adds over the 4 x 64 terms (worst case)
Compared with: plain fp32 (x - mu first, then the product with sequential fp32 adds) = what the reference computes.
Prints norm-wise relative errors vs fp64.  Pure numpy; no GPU.      python tools/tf32_split_accuracy.py
"""
import numpy as np


def rn_tf32(a):
    u = a.astype(np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def trunc_tf32(a):
    return (a.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def seq32(terms):
    """sum over axis 0 with sequential fp32 adds"""
    acc = np.zeros(terms.shape[1:], np.float32)
    for t in terms:
        acc = (acc + t.astype(np.float32)).astype(np.float32)
    return acc


def run(cond, offset, gs=64, npx=4096, seed=0):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((gs, gs)))
    ev = np.logspace(0, -np.log10(cond), gs)                      # covariance eigenvalues 1 .. 1/cond
    x64 = (q * np.sqrt(ev)) @ rng.standard_normal((gs, npx)) + offset
    mu = x64.mean(1, keepdims=True)
    cov = np.cov(x64, bias=True)
    W64 = np.linalg.inv(np.linalg.cholesky((1 - 1e-3) * cov + 1e-3 * np.eye(gs)))
    y64 = W64 @ (x64 - mu)
    W, x, m = W64.astype(np.float32), x64.astype(np.float32), mu.astype(np.float32)
    a_hi = rn_tf32(W); a_lo = rn_tf32(W - a_hi)
    b_hi = trunc_tf32(x); b_lo = rn_tf32(x - b_hi)
    terms = [a[:, :, None].astype(np.float64) * b[None, :, :].astype(np.float64)         # [i, k, px], exact products
             for a in (a_hi, a_lo) for b in (b_hi, b_lo)]
    rc = (W.astype(np.float64) @ m.astype(np.float64)).astype(np.float32)                 # row constant (fp32 FMA chain)
    y_best = (sum(t.sum(1) for t in terms)).astype(np.float32) - rc
    y_worst = seq32(np.concatenate([np.moveaxis(t, 1, 0) for t in terms], 0)) - rc
    xc = (x - m).astype(np.float32)
    y_fp32 = seq32(np.moveaxis(W[:, :, None] * xc[None, :, :], 1, 0))
    single = (rn_tf32(W).astype(np.float64) @ rn_tf32(xc).astype(np.float64))            # one-pass TF32 for contrast
    rel = lambda y: float(np.linalg.norm(y - y64) / np.linalg.norm(y64))                  # noqa: E731
    return rel(y_best), rel(y_worst), rel(y_fp32), rel(single)


if __name__ == "__main__":
    print(f"{'cond(cov)':>10} {'offset/sigma':>12} | {'split f64acc':>12} {'split f32seq':>12} {'plain fp32':>12} {'1-pass tf32':>12}")
    for cond in (1e1, 1e3, 1e5):
        for offset in (0.0, 2.0, 50.0):
            r = run(cond, offset)
            print(f"{cond:10.0e} {offset:12.1f} | " + " ".join(f"{v:12.2e}" for v in r))
