"""BASELINE.json configs[0]: usps_mnist.py USPS->MNIST, group_size=4, 1 epoch -- plumbing.

CPU side (build container, needs /root/reference): the UNMODIFIED reference script is driven end to end
on synthetic dataset files, its `from whitening import WTransform2d` resolved by sys.path order to a
stand-in directory (tests/support/port_utils) -- the same zero-edit mechanism that puts the CUDA layers
under it on a B200.  The LeNet restatement used on the GPU box is checked against
the script's own LeNet.  GPU side: one epoch of the restated LeNet on the CUDA layers.
"""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
SUPPORT = os.path.join(ROOT, "tests", "support")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "usps_mnist.py")), reason="reference not mounted")


@needs_ref
@pytest.mark.timeout(900)
def test_reference_script_runs_unmodified_on_synthetic_files(tmp_path):
    from harness.synth_digits import write_digit_files
    write_digit_files(str(tmp_path / "data"), seed=0)
    work = tmp_path / "work"
    work.mkdir()
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="4",
               PYTHONPATH=os.pathsep.join([os.path.join(SUPPORT, "stubs"), os.path.join(SUPPORT, "port_utils"), ROOT]))
    cmd = [sys.executable, os.path.join(REF, "usps_mnist.py"), "--group_size", "4", "--source", "usps", "--target",
           "mnist", "--epochs", "1", "--num_workers", "0", "--log_interval", "4", "--seed", "1"]
    res = subprocess.run(cmd, cwd=str(work), env=env, capture_output=True, text=True, timeout=800)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "Train Epoch: 0" in res.stdout and "Test set: Classification loss" in res.stdout, res.stdout[-1000:]


@needs_ref
def test_lenet_restatement_matches_reference_lenet(monkeypatch):
    import importlib
    import oracle.torch_port as port
    from harness.lenet_dwt import LeNetDWT
    monkeypatch.syspath_prepend(os.path.join(SUPPORT, "port_utils"))
    monkeypatch.syspath_prepend(os.path.join(SUPPORT, "stubs"))
    monkeypatch.syspath_prepend(REF)
    monkeypatch.chdir(REF)
    script = importlib.import_module("usps_mnist")
    torch.manual_seed(0)
    ref = script.LeNet(group_size=4)
    mine = LeNetDWT(port, group_size=4)
    missing, unexpected = mine.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(16, 1, 28, 28)
    ref.train(); mine.train()
    assert torch.allclose(ref(x), mine(x), atol=1e-5)
    ref.eval(); mine.eval()
    assert torch.allclose(ref(x), mine(x), atol=1e-5)


@pytest.mark.gpu
def test_config1_lenet_epoch_on_cuda_layers():
    import dwt_b200
    import oracle.torch_port as port
    from dwt_b200 import _native
    from harness.lenet_dwt import LeNetDWT, train_epoch
    from harness.synth_digits import digit_batches
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    model = LeNetDWT(dwt_b200, group_size=4).to(dev)
    twin = LeNetDWT(port, group_size=4).to(dev)           # same topology on stock ATen ops, same weights
    twin.load_state_dict(model.state_dict())
    batches = digit_batches(seed=3, steps=16)
    log = train_epoch(model, torch.optim.SGD(model.parameters(), lr=1e-2, momentum=0.5), batches, dev)
    log_twin = train_epoch(twin, torch.optim.SGD(twin.parameters(), lr=1e-2, momentum=0.5), batches, dev)   # (Adam's sign-like steps amplify 1e-7 differences)
    assert _native.status(dev) == 0
    assert all(abs(a[0] - b[0]) < 2e-2 * max(1.0, abs(b[0])) for a, b in zip(log, log_twin)), (log[-3:], log_twin[-3:])
    assert log[-1][0] < log[0][0]                          # the plumbing run learns
    model.eval(); twin.eval()
    x = batches[0][2].to(dev)
    with torch.no_grad():
        assert (model(x) - twin(x)).abs().max() < 5e-2 * twin(x).abs().max()
