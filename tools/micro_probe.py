"""Where does the config-2 microbench step go?  Eager vs profiled vs CUDA-graph, and host-side cost per call."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dwt-domain-adaptation_b200"))
import torch
import dwt_b200
from dwt_b200 import _native

dev = torch.device("cuda:0")
N, C, H, gs = 256, 256, 56, 64
torch.manual_seed(0)
x = (torch.randn(N, C, H, H, device=dev) + 2.0).requires_grad_(True)
dy = torch.randn(N, C, H, H, device=dev)
m = dwt_b200.WTransform2d(C, gs).to(dev).train()

def step():
    y = m(x)
    torch.autograd.grad(y, x, dy)

def timed(fn, k=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(k):
        fn()
    e1.record(); host = (time.perf_counter() - t0) / k * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, host

for _ in range(5):
    step()
print("eager   gpu ms/step %.3f  host ms/step %.3f" % timed(step))
_native.profile_begin()
r = timed(step)
prof = _native.by_family(_native.profile_end())
print("profiled gpu ms/step %.3f  host ms/step %.3f" % r)
print("  kernel sum ms/step %.3f" % (sum(v["ms"] for v in prof.values()) / 10))
with torch.no_grad():
    print("fwd only gpu %.3f host %.3f" % timed(lambda: m(x)))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
g.replay()
print("graph   gpu ms/step %.3f  host ms/step %.3f" % timed(g.replay))
