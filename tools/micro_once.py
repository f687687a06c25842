"""Two fwd+bwd passes of the config-2 whitening layer (for ncu captures): python tools/micro_once.py [N]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dwt-domain-adaptation_b200"))
import torch
import dwt_b200
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = (torch.randn(N, 256, 56, 56, device=dev) + 2.0).requires_grad_(True)
dy = torch.randn(N, 256, 56, 56, device=dev)
m = dwt_b200.WTransform2d(256, 64).to(dev).train()
for _ in range(2):
    y = m(x)
    torch.autograd.grad(y, x, dy)
torch.cuda.synchronize()
