"""Development: clock64 timeline of 48 consecutive tiles of one CTA of tc_gram_kernel (norm_tc.cu built with
-DDWT_PROF_GRAM into tools/gpu/prof/libdwt_b200_gramprof.so).  Columns: cycles since the TMA issue of the first recorded
tile -- producer: TMA issued; transform warp (quarter 0 of the tile's set): tile landed, A slot free, arrived on `ready`;
MMA issuer: `ready` seen, MMAs + commits issued."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _variant
os.environ["DWT_B200_LIB"] = _variant.build("gramprof", "DWT_PROF_GRAM", "norm_tc.cu")
sys.path.insert(0, os.path.join(HERE, "..", "..", "..", "dwt-domain-adaptation_b200"))
import torch
import dwt_b200
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(256, 256, 56, 56, device=dev) + 2.0
m = dwt_b200.WTransform2d(256, 64).to(dev).train()
with torch.no_grad():
    for _ in range(2):
        m(x)
torch.cuda.synchronize()
