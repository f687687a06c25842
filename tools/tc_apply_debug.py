"""Scratch: identity-whitening probe of the tensor-core apply kernel (eval mode, cov = I, mean = 0, eps = 0
=> y must equal x bit for bit up to the tf32 split).  Prints where y's values come from if not."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dwt-domain-adaptation_b200"), ROOT]
import torch, dwt_b200

dev = torch.device("cuda", 0)
for (N, C, H, gs) in [(8, 64, 32, 64), (8, 128, 32, 32), (6, 64, 28, 64)]:
    torch.manual_seed(0)
    x = torch.randn(N, C, H, H, device=dev)
    m = dwt_b200.WTransform2d(C, gs, running_m=torch.zeros(1, C, 1, 1, device=dev),
                              running_var=torch.eye(gs, device=dev).repeat(C // gs, 1, 1), eps=0.0).eval()
    # distinct scales per channel so a channel permutation is visible
    m.running_variance.mul_(1.0)
    y = m(x)
    torch.cuda.synchronize()
    d = (y - x).abs().max().item()
    print(f"N={N} C={C} HW={H*H} gs={gs}: max|y-x| = {d:.3e}   |y|max={y.abs().max().item():.3f} zeros={(y==0).float().mean().item():.3f}")
    if d > 1e-4:
        xf, yf = x[0].reshape(C, -1), y[0].reshape(C, -1)
        for c in (0, 1, 17):
            for p in (0, 1, 5, 33, 70):
                v = yf[c, p]
                hit = (xf - v).abs() < 1e-6
                idx = hit.nonzero()[:3].tolist()
                print(f"   y[0,{c},{p}] = {v.item():+.5f}  x there = {xf[c,p].item():+.5f}  matches x at {idx}")
    # backward identity: dx = W^T dy = dy
    x.requires_grad_(True)
    y = m(x)
    dy = torch.randn_like(y)
    (dx,) = torch.autograd.grad(y, x, dy)
    print(f"   bwd: max|dx-dy| = {(dx-dy).abs().max().item():.3e}")

# offset probe: non-zero running mean, data far from zero (the centring is applied after the product)
for off in (2.0, 50.0):
    torch.manual_seed(1)
    N, C, H, gs = 8, 64, 32, 64
    x = torch.randn(N, C, H, H, device=dev) + off
    mu = torch.full((1, C, 1, 1), off, device=dev) + 0.1 * torch.randn(1, C, 1, 1, device=dev)
    m = dwt_b200.WTransform2d(C, gs, running_m=mu.clone(), running_var=torch.eye(gs, device=dev).repeat(C // gs, 1, 1), eps=0.0).eval()
    y = m(x)
    print(f"offset probe off={off}: max|y-(x-mu)| = {(y - (x - mu)).abs().max().item():.3e}")
