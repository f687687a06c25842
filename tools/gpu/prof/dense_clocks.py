"""Development: phase clocks of the dense finalize kernels (norm_dense.cu built with -DDWT_PROF_DENSE into
tools/gpu/prof/libdwt_b200_prof.so by _variant.py; see tools/gpu/run_p.sh).  Prints one line per kernel launch (CTA 0)."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "..", "dwt-domain-adaptation_b200"))
sys.path.insert(0, HERE)
import _variant
os.environ["DWT_B200_LIB"] = _variant.build("prof", "DWT_PROF_DENSE", "norm_dense.cu")      # build it HERE (nvcc), run it on the GPU box
import torch
import dwt_b200
dev = torch.device("cuda:0")
torch.manual_seed(0)
N = 64
x = (torch.randn(N, 256, 56, 56, device=dev) + 2.0).requires_grad_(True)
dy = torch.randn(N, 256, 56, 56, device=dev)
m = dwt_b200.WTransform2d(256, 64).to(dev).train()
for _ in range(3):
    y = m(x)
    torch.autograd.grad(y, x, dy)
torch.cuda.synchronize()
