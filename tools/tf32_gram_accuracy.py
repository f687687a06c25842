"""CPU emulation of the single-pass TF32 Gram contraction (csrc/norm_tc.cu): covariance error vs sample count.

  G = sum_m RN_tf32(x_m - K) RN_tf32(x_m - K)^T   (products exact, fp32 accumulation emulated in fp64: the kernel
  keeps per-CTA partials of <= a few thousand samples each and reduces them in fixed order)
compared with the fp64 covariance; also the same with TRUNCATED operands (what feeding the raw tile would do) and the
resulting error of y = W (x - mu).  Pure numpy.      python tools/tf32_gram_accuracy.py
"""
import numpy as np


def rn(a):
    u = a.astype(np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)


def tr(a):
    return (a.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)


def run(M, gs=64, cond=1e2, offset=2.0, seed=0):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((gs, gs)))
    x = (q * np.sqrt(np.logspace(0, -np.log10(cond), gs))) @ rng.standard_normal((gs, M)) + offset
    K = x[:, M // 2 - 16:M // 2 + 16].mean(1, keepdims=True) if M >= 32 else x.mean(1, keepdims=True)   # pilot shift
    xs = (x.astype(np.float32) - K.astype(np.float32)).astype(np.float32).astype(np.float64)
    s1 = xs.sum(1, keepdims=True) / M

    def cov_from(op):
        o = op(xs)
        return o @ o.T / M - s1 @ s1.T

    exact = np.cov(x, bias=True)
    w = lambda c: np.linalg.inv(np.linalg.cholesky((1 - 1e-3) * c + 1e-3 * np.eye(gs)))       # noqa: E731
    y64 = w(exact) @ (x - x.mean(1, keepdims=True))
    out = []
    for op in (rn, tr):
        c = cov_from(op)
        out.append(np.linalg.norm(c - exact) / np.linalg.norm(exact))
        out.append(np.linalg.norm(w(c) @ (x - x.mean(1, keepdims=True)) - y64) / np.linalg.norm(y64))
    return out


if __name__ == "__main__":
    print(f"{'M':>8} | {'cov err RN':>11} {'y err RN':>11} | {'cov err trunc':>13} {'y err trunc':>11}")
    for M in (144, 1024, 4096, 65536, 802816 // 8):
        r = run(M)
        print(f"{M:8d} | {r[0]:11.2e} {r[1]:11.2e} | {r[2]:13.2e} {r[3]:11.2e}")
