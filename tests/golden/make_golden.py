"""Generate golden fixtures by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py          # needs /root/reference, writes tests/golden/*.npz

The reference ships no tests or golden vectors, so these outputs of
the reference itself are the parity pin for the oracle and for the CUDA path.
Everything is fp32 on CPU (torch build recorded in each file).  Inputs are
seeded; sizes are kept tiny so the fixtures stay a few hundred KB.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "utils"))
warnings.filterwarnings("ignore")

import whitening as ref_whitening            # noqa: E402  (reference utils/whitening.py)
import consensus_loss as ref_mec             # noqa: E402
import batch_norm as ref_bn                  # noqa: E402

assert ref_whitening.__file__.startswith(REF), ref_whitening.__file__

# (name, N, C, H, W, group_size)
WHITEN_CASES = [
    ("w_c8_g4", 6, 8, 5, 5, 4),          # HW not a multiple of 4
    ("w_c64_g4", 3, 64, 8, 8, 4),        # ResNet layer1 shape family
    ("w_c16_g1", 5, 16, 4, 6, 1),        # gs=1: per-channel standardisation
    ("w_c16_g2", 3, 16, 7, 7, 2),
    ("w_c32_g8", 4, 32, 6, 6, 8),
    ("w_c32_g16", 4, 32, 6, 6, 16),
    ("w_c64_g32", 4, 64, 6, 6, 32),
    ("w_c64_g64", 4, 64, 6, 6, 64),     # microbench group size
    ("w_c128_g64", 3, 128, 6, 6, 64),
    ("w_c4_g8clamp", 6, 4, 5, 5, 8),     # group_size > C clamps to C (whitening.py:14)
    ("w_c48_g4_lenet", 2, 48, 14, 14, 4),  # usps_mnist.py conv2 site
]


def correlated(gen, n, c, h, w):
    """Correlated, non-zero-mean channels: N(0,1) mixed by a fixed CxC matrix, offset 2."""
    z = torch.randn(n, c, h, w, generator=gen)
    mix = torch.randn(c, c, generator=gen) / c ** 0.5 + torch.eye(c)
    return torch.einsum("dc,nchw->ndhw", mix, z) + 2.0


def whiten_case(name, n, c, h, w, gs, seed):
    gen = torch.Generator().manual_seed(seed)
    x1, x2 = correlated(gen, n, c, h, w), correlated(gen, n, c, h, w)
    dy = torch.randn(n, c, h, w, generator=gen)
    gse = min(c, gs)
    rm0 = 0.1 * torch.randn(1, c, 1, 1, generator=gen)
    a = torch.randn(c // gse, gse, gse, generator=gen)
    rv0 = a @ a.transpose(1, 2) / gse + 0.5 * torch.eye(gse)
    out = dict(x1=x1, x2=x2, dy=dy, rm0=rm0, rv0=rv0, gs=np.int64(gs))

    # externally owned buffers, two training steps, then eval
    m = ref_whitening.WTransform2d(c, gs, running_m=rm0.clone(), running_var=rv0.clone())
    m.train()
    xa = x1.clone().requires_grad_(True)
    y1 = m(xa)
    (dx1,) = torch.autograd.grad(y1, xa, dy)
    out.update(y1=y1.detach(), dx1=dx1, rm1=m.running_mean.clone(), rv1=m.running_variance.clone())
    with torch.no_grad():                       # no-grad train-mode forward (stats collection, :382-389)
        y2 = m(x2)
    out.update(y2=y2, rm2=m.running_mean.clone(), rv2=m.running_variance.clone())
    m.eval()
    xe = x1.clone().requires_grad_(True)
    ye = m(xe)
    (dxe,) = torch.autograd.grad(ye, xe, dy)
    out.update(y_eval=ye.detach(), dx_eval=dxe, rm_eval=m.running_mean.clone())

    # default-constructed buffers (zeros / all-ones), one training step
    d = ref_whitening.WTransform2d(c, gs)
    d.train()
    with torch.no_grad():
        d(x1)
    out.update(rm_default1=d.running_mean.clone(), rv_default1=d.running_variance.clone())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), torch=torch.__version__,
                        **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()})


def mec_cases():
    gen = torch.Generator().manual_seed(7)
    cases = {}
    for name, n, k, scale in [("plain", 6, 65, 1.0), ("big", 5, 65, 40.0), ("k10", 9, 10, 3.0), ("n1", 1, 7, 1.0)]:
        x = (scale * torch.randn(n, k, generator=gen)).requires_grad_(True)
        y = (scale * torch.randn(n, k, generator=gen)).requires_grad_(True)
        cases[name] = (x, y)
    # exact ties: identical rows / two equal maxima
    xt = torch.zeros(4, 5)
    xt[1, 2] = xt[1, 4] = 3.0
    xt[2] = torch.tensor([1.0, 2.0, 2.0, 0.0, -1.0])
    cases["ties"] = (xt.clone().requires_grad_(True), xt.clone().requires_grad_(True))
    out = {}
    for name, (x, y) in cases.items():
        crit = ref_mec.MinEntropyConsensusLoss(num_classes=x.shape[1], device="cpu")
        loss = crit(x, y)
        gx, gy = torch.autograd.grad(loss, (x, y))
        out.update({f"{name}_x": x.detach(), f"{name}_y": y.detach(), f"{name}_loss": loss.detach(),
                    f"{name}_gx": gx, f"{name}_gy": gy})
    np.savez_compressed(os.path.join(HERE, "mec.npz"), torch=torch.__version__,
                        **{k: v.numpy() for k, v in out.items()})


def bn_cases():
    gen = torch.Generator().manual_seed(11)
    out = {}
    specs = [("bn2d_affine", ref_bn.BatchNorm2d, (4, 6, 5, 5), True, 0.1),
             ("bn2d_plain", ref_bn.BatchNorm2d, (6, 16, 7, 7), False, 0.1),
             ("bn2d_hw4", ref_bn.BatchNorm2d, (3, 8, 4, 8), False, 0.1),
             ("bn2d_cma", ref_bn.BatchNorm2d, (4, 6, 3, 3), False, None),
             ("bn1d_2", ref_bn.BatchNorm1d, (8, 10), True, 0.1),
             ("bn1d_3", ref_bn.BatchNorm1d, (4, 6, 9), False, 0.1),
             ("bn3d", ref_bn.BatchNorm3d, (3, 4, 2, 3, 5), True, 0.3)]
    for name, cls, shape, affine, mom in specs:
        c = shape[1]
        x1 = 1.5 * torch.randn(*shape, generator=gen) + 0.7
        x2 = 0.5 * torch.randn(*shape, generator=gen) - 1.0
        dy = torch.randn(*shape, generator=gen)
        rm0, rv0 = 0.1 * torch.randn(c, generator=gen), 0.5 + torch.rand(c, generator=gen)
        m = cls(c, rm0.clone(), rv0.clone(), affine=affine, momentum=mom)
        if affine:
            with torch.no_grad():
                m.weight.copy_(0.5 + torch.rand(c, generator=gen))
                m.bias.copy_(0.1 * torch.randn(c, generator=gen))
            out[f"{name}_weight"], out[f"{name}_bias"] = m.weight.detach().clone(), m.bias.detach().clone()
        m.train()
        xa = x1.clone().requires_grad_(True)
        y1 = m(xa)
        y1.backward(dy)
        out.update({f"{name}_x1": x1, f"{name}_x2": x2, f"{name}_dy": dy, f"{name}_rm0": rm0, f"{name}_rv0": rv0,
                    f"{name}_y1": y1.detach(), f"{name}_dx1": xa.grad.clone(),
                    f"{name}_rm1": m.running_mean.clone(), f"{name}_rv1": m.running_var.clone()})
        if affine:
            out[f"{name}_dweight"], out[f"{name}_dbias"] = m.weight.grad.clone(), m.bias.grad.clone()
        with torch.no_grad():
            m(x2)
        out.update({f"{name}_rm2": m.running_mean.clone(), f"{name}_rv2": m.running_var.clone(),
                    f"{name}_nbt2": m.num_batches_tracked.clone()})
        m.eval()
        xe = x1.clone().requires_grad_(True)
        ye = m(xe)
        ye.backward(dy)
        out.update({f"{name}_y_eval": ye.detach(), f"{name}_dx_eval": xe.grad.clone()})
    np.savez_compressed(os.path.join(HERE, "bn.npz"), torch=torch.__version__,
                        **{k: v.numpy() for k, v in out.items()})


def resnet_case(size=96, name="resnet_tiny"):
    """Full reference model (resnet50_dwt_mec_officehome.py ResNet) on a small synthetic batch.
    size=96: tiny spatial sites; size=224: the real site shapes of BASELINE configs[2]
    (112^2 / 56^2 / 28^2 / 14^2 / 7^2) at 4 images per domain."""
    from harness.synth import synth_batch, synth_state_dict
    cwd = os.getcwd()
    os.chdir(REF)
    sys.path.insert(0, REF)
    import resnet50_dwt_mec_officehome as script
    os.chdir(cwd)
    sd = synth_state_dict(seed=1)
    x, labels = synth_batch(seed=2, per_domain=4, size=size)
    model = script.ResNet(script.Bottleneck, [3, 4, 6, 3], {k: v.clone() for k, v in sd.items()})
    model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=False)
    model.train()
    logits = model(x)
    s, t, a = torch.split(logits, logits.shape[0] // 3, dim=0)
    cls = F.nll_loss(F.log_softmax(s, dim=1), labels)
    mec = 0.1 * ref_mec.MinEntropyConsensusLoss(num_classes=65, device="cpu")(t, a)
    (cls + mec).backward()
    params = dict(model.named_parameters())
    bufs = model.state_dict()
    pick = ["conv1.weight", "gamma1", "beta1", "layer1.0.gamma2", "layer1.2.conv3.weight", "layer2.0.downsample_gamma",
            "layer3.5.beta3", "fc_out.bias"]
    out = dict(logits=logits.detach(), cls_loss=cls.detach(), mec_loss=mec.detach())
    for k in pick:
        out["grad/" + k] = params[k].grad
    out["gradnorms"] = torch.stack([p.grad.norm() for p in params.values()])
    out["gradnames"] = np.array(list(params.keys()))
    for k in ["bns1.wh.running_mean", "bnt1_aug.wh.running_variance", "layer1.1.bnt2.wh.running_variance",
              "layer2.0.bns1.running_var", "layer4.2.bnt3_aug.running_mean"]:
        out["buf/" + k] = bufs[k].clone()
    model.eval()
    with torch.no_grad():
        out["logits_eval"] = model(x)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), torch=torch.__version__,
                        **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()})


if __name__ == "__main__":
    torch.set_num_threads(8)
    for i, case in enumerate(WHITEN_CASES):
        whiten_case(*case, seed=100 + i)
    mec_cases()
    bn_cases()
    resnet_case()
    resnet_case(size=224, name="resnet_224")
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
