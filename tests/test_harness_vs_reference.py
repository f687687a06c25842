"""The measurement harness (harness/resnet50_dwt.py) against the UNMODIFIED reference model
(/root/reference/resnet50_dwt_mec_officehome.py ResNet) on CPU, both built on stock ATen layers:
same synthetic checkpoint, same batch -> same logits, loss, gradients, buffers, eval logits.
Guards the topology restatement (incl. the Bottleneck tail that the fused site folds in)."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "resnet50_dwt_mec_officehome.py")),
                                reason="reference not mounted")


def test_harness_equals_reference_model(monkeypatch):
    import importlib
    import oracle.torch_port as port
    from harness.resnet50_dwt import build_resnet50_dwt
    from harness.synth import synth_batch, synth_state_dict
    monkeypatch.syspath_prepend(os.path.join(REF, "utils"))
    monkeypatch.syspath_prepend(REF)
    monkeypatch.chdir(REF)
    ref_script = importlib.import_module("resnet50_dwt_mec_officehome")
    ref_mec = importlib.import_module("consensus_loss")
    torch.set_num_threads(8)
    sd = synth_state_dict(1)
    x, y = synth_batch(2, 2, size=64)
    clone = lambda: {k: v.clone() for k, v in sd.items()}
    ref = ref_script.ResNet(ref_script.Bottleneck, [3, 4, 6, 3], clone())
    ref.load_state_dict(clone(), strict=False)
    mine = build_resnet50_dwt(clone(), port)
    assert set(ref.state_dict()) == set(mine.state_dict())

    def step(m, mec):
        m.train()
        out = m(x)
        s, t, a = out.split(out.shape[0] // 3)
        loss = F.nll_loss(F.log_softmax(s, 1), y) + 0.1 * mec(t, a)
        loss.backward()
        return out.detach(), loss.item()

    o1, l1 = step(ref, ref_mec.MinEntropyConsensusLoss(65, "cpu"))
    o2, l2 = step(mine, port.MinEntropyConsensusLoss(65, "cpu"))
    assert torch.allclose(o1, o2, atol=1e-5) and abs(l1 - l2) < 1e-5
    g1, g2 = dict(ref.named_parameters()), dict(mine.named_parameters())
    for k in g1:
        assert (g1[k].grad - g2[k].grad).norm() <= 1e-4 * g1[k].grad.norm() + 1e-9, k
    b1, b2 = ref.state_dict(), mine.state_dict()
    for k in b1:
        assert torch.allclose(b1[k].float(), b2[k].float(), atol=1e-5), k
    ref.eval(); mine.eval()
    with torch.no_grad():
        assert torch.allclose(ref(x), mine(x), atol=1e-4)


@pytest.mark.skipif(not os.path.isdir("/root/reference/utils"), reason="reference not mounted (GPU box)")
def test_module_surface_is_the_references_own():
    """Build-container check against the UNMODIFIED reference modules: attributes, buffer identity, state-dict keys
    and initial values (same RNG stream), repr, checkpoint-version shim and every error text, string for string."""
    import importlib.util
    import warnings
    warnings.filterwarnings("ignore")
    from dwt_b200 import batch_norm as B, whitening as W

    def load_ref(name):
        spec = importlib.util.spec_from_file_location("ref_" + name, f"/root/reference/utils/{name}.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    def err(fn, *a):
        try:
            fn(*a)
        except ValueError as e:
            return str(e)
        return None

    RW, RB = load_ref("whitening"), load_ref("batch_norm")
    for args in [dict(num_features=8, group_size=4), dict(num_features=8, group_size=16), dict(num_features=48, group_size=32),
                 dict(num_features=8, group_size=4, running_m=torch.zeros(1, 8, 1, 1), running_var=torch.ones(2, 4, 4)),
                 dict(num_features=8, group_size=4, running_m=torch.zeros(1, 8, 1, 1), running_var=torch.ones(2, 4, 4),
                      track_running_stats=False, momentum=0.3, eps=1e-2, alpha=2)]:
        a, r = W.WTransform2d(**args), RW.WTransform2d(**args)
        for k in ["num_features", "momentum", "track_running_stats", "eps", "alpha", "group_size", "num_groups"]:
            assert getattr(a, k) == getattr(r, k), (k, args)
        assert a.running_m is args.get("running_m") and a.running_var is args.get("running_var")
        assert list(a.state_dict()) == list(r.state_dict())
        assert all(torch.equal(a.state_dict()[k], r.state_dict()[k].cpu()) for k in a.state_dict())
        for bad in [torch.zeros(2, 8, 3), torch.zeros(2, args["num_features"], 3, 3)]:
            check = lambda m: (m._check_input_dim(bad), m._check_group_size())      # noqa: E731
            assert err(check, a) == err(check, r)
    for cls in ("BatchNorm1d", "BatchNorm2d", "BatchNorm3d"):
        for kw in [dict(affine=True), dict(affine=False), dict(affine=True, track_running_stats=False),
                   dict(momentum=None, eps=1e-3)]:
            torch.manual_seed(0)
            a = getattr(B, cls)(6, torch.zeros(6), torch.ones(6), **kw)
            torch.manual_seed(0)
            r = getattr(RB, cls)(6, torch.zeros(6), torch.ones(6), **kw)
            assert repr(a) == repr(r) and a._version == r._version
            assert list(a.state_dict()) == list(r.state_dict())
            assert all(torch.equal(a.state_dict()[k], r.state_dict()[k]) for k in a.state_dict())
            assert [n for n, _ in a.named_parameters()] == [n for n, _ in r.named_parameters()]
            assert [n for n, _ in a.named_buffers()] == [n for n, _ in r.named_buffers()]
            for nd in range(2, 7):
                x = torch.zeros(*([2, 6] + [3] * (nd - 2)))
                assert err(a._check_input_dim, x) == err(r._check_input_dim, x), (cls, nd)
            if kw.get("track_running_stats", True):
                v1 = {k: v for k, v in a.state_dict().items() if "num_batches" not in k}
                a.load_state_dict(v1)                             # pre-version-2 checkpoint: no counter


def test_space_to_depth_stem_equals_the_reference_stem(monkeypatch):
    """The harness option stem_s2d (7x7/2 convolution evaluated as the 4x4/1 convolution of the 2x2 space-to-depth image)
    against the UNMODIFIED reference model's conv1: same output and the same gradient of the [64, 3, 7, 7] weight (fp32
    summation order is the only difference), in both memory formats; and the whole model's logits on top of it.  (The
    GPU golden test at 224x224 runs the same option on the CUDA layers.)"""
    import importlib
    import oracle.torch_port as port
    from harness.resnet50_dwt import build_resnet50_dwt
    from harness.synth import synth_batch, synth_state_dict
    monkeypatch.syspath_prepend(os.path.join(REF, "utils"))
    monkeypatch.syspath_prepend(REF)
    monkeypatch.chdir(REF)
    ref_script = importlib.import_module("resnet50_dwt_mec_officehome")
    torch.set_num_threads(8)
    sd = synth_state_dict(1)
    x, _ = synth_batch(3, 2, size=64)
    clone = lambda: {k: v.clone() for k, v in sd.items()}
    ref = ref_script.ResNet(ref_script.Bottleneck, [3, 4, 6, 3], clone())
    ref.load_state_dict(clone(), strict=False)
    ref.train()
    probe = torch.randn(6, 64, 32, 32)
    r_stem = ref.conv1(x)
    (gr,) = torch.autograd.grad((r_stem * probe).sum(), ref.conv1.weight)
    r_logits = ref(x)
    for cl in (False, True):
        mine = build_resnet50_dwt(clone(), port, channels_last=cl, stem_s2d=True).train()
        xi = x.contiguous(memory_format=torch.channels_last) if cl else x
        assert mine.conv1.weight.shape == ref.conv1.weight.shape == (64, 3, 7, 7)
        m_stem = mine._stem(xi)
        assert m_stem.shape == r_stem.shape and torch.allclose(r_stem, m_stem, atol=1e-5), (r_stem - m_stem).abs().max()
        (gm,) = torch.autograd.grad((m_stem * probe).sum(), mine.conv1.weight)
        assert (gr - gm).norm() <= 1e-5 * gr.norm(), ((gr - gm).norm() / gr.norm()).item()
        # 2 x 2 pixels per image reach layer4 at this size: the batch statistics amplify the rounding differences
        assert (r_logits - mine(xi)).abs().max() <= 1e-3 * r_logits.abs().max()
