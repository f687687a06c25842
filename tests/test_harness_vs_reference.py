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
