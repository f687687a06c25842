"""Is fwd+bwd of WTransform2d graph-capturable at a given shape?  python tools/graph_probe.py N C H gs"""
import os, sys, traceback
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dwt-domain-adaptation_b200"))
import torch
import dwt_b200
dev = torch.device("cuda:0")
N, C, H, gs = [int(a) for a in sys.argv[1:5]]
torch.manual_seed(0)
x = (torch.randn(N, C, H, H, device=dev) + 2.0).requires_grad_(True)
dy = torch.randn(N, C, H, H, device=dev)
m = dwt_b200.WTransform2d(C, gs).to(dev).train()
def step():
    y = m(x)
    return y, torch.autograd.grad(y, x, dy)[0]
for _ in range(2):
    step()
torch.cuda.synchronize()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        y, dx = step()
    g.replay(); torch.cuda.synchronize()
    print(sys.argv[1:], "capture OK", float(y.abs().max()), float(dx.abs().max()))
except Exception as e:
    print(sys.argv[1:], "capture FAILED:", str(e).splitlines()[0])
