"""Time the paired augmentation kernel at Office-Home geometry: python tools/augment_time.py [B]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dwt-domain-adaptation_b200"))
import numpy as np, torch
from dwt_b200 import PairedAugment, draw_params, _native
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
img = torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8, device=dev)
p = {k: v.to(dev) for k, v in draw_params(B, 256, 224, rng).items()}
pa = PairedAugment(224)
for cl in (False, True):
    fmt = torch.channels_last if cl else torch.contiguous_format
    batch = torch.empty(2 * B, 3, 224, 224, device=dev).contiguous(memory_format=fmt)
    for _ in range(5):
        pa(img, out_plain=batch[:B], out_aug=batch[B:], channels_last=cl, **p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        pa(img, out_plain=batch[:B], out_aug=batch[B:], channels_last=cl, **p)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    byts = 2 * B * 3 * 224 * 224 * 4 + B * 256 * 256 * 3
    _native.profile_begin()
    for _ in range(20):
        pa(img, out_plain=batch[:B], out_aug=batch[B:], channels_last=cl, **p)
    prof = _native.by_family(_native.profile_end())["augment_pair"]
    kus = 1e3 * prof["ms"] / prof["launches"]
    print(f"   kernel alone (CUDA events around the launch): {kus:.1f} us, {byts / kus / 1e3:.0f} GB/s")
    print(f"B={B} {'NHWC' if cl else 'NCHW'}: {us:.1f} us per launch (incl. host call), {byts / us / 1e3:.0f} GB/s of {byts / 1e6:.1f} MB, "
          f"{2 * B / us * 1e6:.0f} views/s")
