"""World-size-2 data-parallel step on CPU (gloo): the N>1 path of bench.py without a GPU.

The product layers have no CPU path, so the harness model is built on the CPU port of the
reference layers (oracle/torch_port.py) -- what is under test is the host-side plumbing of
bench.py (FlatGradAllReduce + train_step): one process per rank, rank-local whitening statistics
(never exchanged), gradients averaged through the flat buffer in three segments that are all-reduced
as backward completes them (post-accumulate hooks, async collectives joined in reduce()), SURVEY.md §8e.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PICK = ["conv1.weight", "gamma1", "layer1.0.beta2", "layer2.1.gamma3", "layer4.2.conv3.weight", "fc_out.bias"]


def _step(rank_seed, ddp, gather="accumulate", segments=3):
    sys.path[:0] = [ROOT]
    import bench
    import oracle.torch_port as port
    from harness.synth import synth_batch
    torch.manual_seed(0)
    torch.set_num_threads(2)
    model = bench.build_model(port, torch.device("cpu"), "modules")
    sync = bench.FlatGradAllReduce(model, 2, segments=segments, gather=gather) if ddp else None   # the data-parallel plumbing of bench.py
    opt = torch.optim.SGD(model.parameters(), lr=0.0)               # lr 0: inspect the synchronised gradients
    mec = port.MinEntropyConsensusLoss(bench.NUM_CLASSES, "cpu")
    images, labels = synth_batch(seed=rank_seed, per_domain=1, size=64)
    bench.train_step(model, mec, opt, images, labels, sync)
    params = dict(model.named_parameters())
    return {k: params[k].grad.clone() for k in PICK}, model.state_dict()["bns1.wh.running_mean"].clone()


def _worker(rank, world, port_no, out_dir, gather, segments):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port_no}", rank=rank, world_size=world)
    grads, buf = _step(100 + rank, ddp=True, gather=gather, segments=segments)
    torch.save({"grads": grads, "buf": buf}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("gather,segments", [("accumulate", 3), ("copy", 1), ("copy", 3)])   # overlapped segments / one copy + one collective / both
def test_two_rank_step_averages_gradients_and_keeps_statistics_local(tmp_path, gather, segments):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_no = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port_no, str(tmp_path), gather, segments), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt") for r in range(2))
    g0, b0 = _step(100, ddp=False)
    g1, b1 = _step(101, ddp=False)
    for k in PICK:
        assert torch.allclose(r0["grads"][k], r1["grads"][k], rtol=0, atol=0), k          # identical after all-reduce
        mean = 0.5 * (g0[k] + g1[k])
        assert torch.allclose(r0["grads"][k], mean, rtol=1e-4, atol=1e-6 * mean.abs().max().item()), k
    # whitening statistics stay rank-local: each rank's buffer equals its own single-process run
    assert torch.allclose(r0["buf"], b0, atol=1e-6) and torch.allclose(r1["buf"], b1, atol=1e-6)
    assert not torch.allclose(r0["buf"], r1["buf"], atol=1e-6)
