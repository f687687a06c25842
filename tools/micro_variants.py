"""Per-kernel times of the config-2 whitening layer (CUDA-event brackets of the library), repeated per argument.
The DWT_TC_DBG switches this script drove during the round-1 experiments (profiles/tc_apply_experiments_r01.md)
were removed from the kernels again; the arguments now only label repeated runs."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dwt-domain-adaptation_b200"))
import torch
import dwt_b200
from dwt_b200 import _native
dev = torch.device("cuda:0")
N = 256
torch.manual_seed(0)
x = (torch.randn(N, 256, 56, 56, device=dev) + 2.0).requires_grad_(True)
dy = torch.randn(N, 256, 56, 56, device=dev)
m = dwt_b200.WTransform2d(256, 64).to(dev).train()
def step():
    y = m(x)
    torch.autograd.grad(y, x, dy)
for dbg in [int(a) for a in sys.argv[1:]] or [0]:
    os.environ["DWT_TC_DBG"] = str(dbg)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    _native.profile_begin()
    for _ in range(10):
        step()
    prof = _native.by_family(_native.profile_end())
    print("dbg", dbg, {k: round(1e3 * v["ms"] / v["launches"], 1) for k, v in prof.items()})
