#!/usr/bin/env python
"""bench.py -- ResNet-50-DWT training-step throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload resnet|microbench] [--per-domain B] [--site-mode fused|modules]

One "step" = one full training step of the ResNet-50-DWT harness model (harness/resnet50_dwt.py,
the reference topology with this repo's CUDA layers dropped in) on one synthetic Office-Home-shaped
batch: B source + B target + B target-aug images of 3x224x224 (B = 64 per domain -> 192 images per
GPU), forward, NLL(source) + 0.1 * MEC(target, target-aug), backward, SGD(momentum)
update.  N > 1: one process per GPU (torchrun), plain data parallel, gradients averaged by ONE NCCL
all-reduce of a flat 94.6 MB buffer per step (rank-local whitening statistics, nothing else exchanged).

Prints ONE JSON line on rank 0 (see the task contract): `value` = images/s with inputs resident in
HBM; `e2e` = images/s through the same public call with the step's images copied from pinned host
memory and the loss read back inside the timed region; `roofline` = the dominant hand-written
kernel family against the measured HBM peak; `cpu_baseline` = the CPU port of the reference
layers (oracle/torch_port.py, same harness model) timed on this box's host cores.

--impl reference: the reference's own CPU path (no GPU): the same step with the reference's
operator sequence on stock ATen CPU ops (oracle/torch_port.py), all host threads, a bounded sample.
--workload microbench: BASELINE.json configs[1] (WTransform2d N=256 C=256 56x56 gs=64 fwd+bwd).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PLUGIN_DIR = os.path.join(ROOT, "dwt-domain-adaptation_b200")
for _p in (PLUGIN_DIR, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch                                     # noqa: E402
import torch.nn.functional as F                  # noqa: E402

LAMBDA_MEC = 0.1                                 # resnet50_dwt_mec_officehome.py:508
NUM_CLASSES = 65


def host_cores() -> int:
    """Usable host cores: min(affinity mask, cgroup CPU quota).  The GPU boxes expose 128 logical
    CPUs but cap the container at 16 by cgroup quota; oversubscribing the quota is ~10x slower."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        if os.environ.get("DWT_NO_CLOCK_SAMPLER"):          # diagnosis only: is the sampler perturbing the run?
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            # nvidia-smi's start-up (NVML init, driver queries) briefly stalls launches on the GPU it opens: let
            # it deliver its first sample before the timed region starts
            t_end = time.time() + 5.0
            while not self.rows and time.time() < t_end:
                time.sleep(0.02)
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------- model
def build_model(layers, device, site_mode, seed=1, channels_last=False):
    from harness.resnet50_dwt import build_resnet50_dwt
    from harness.synth import synth_state_dict
    sd = {k: v.to(device) for k, v in synth_state_dict(seed=seed).items()}
    model = build_resnet50_dwt(sd, layers, site_mode=site_mode, channels_last=channels_last).to(device)
    return model.train()


def make_optimizer(model):
    """SGD with the reference's two parameter groups (resnet50_dwt_mec_officehome.py:578-590)."""
    head = [p for n, p in model.named_parameters() if n.startswith("fc_out")]
    body = [p for n, p in model.named_parameters() if not n.startswith("fc_out")]
    return torch.optim.SGD([{"params": body, "lr": 1e-3}, {"params": head, "lr": 1e-2}], momentum=0.9,
                           weight_decay=5e-4)


class FlatGradAllReduce:
    """Plain data parallelism: every parameter's .grad is a view into ONE flat fp32 buffer
    (94.6 MB for ResNet-50-DWT); after backward a single all-reduce averages it over the ranks (NCCL over
    NVLink on the B200 box, gloo in the CPU test).  One collective per step, no hooks, no bucketing logic --
    so the whole step, collective included, can be captured into a CUDA graph.  Whitening / BN statistics are
    never exchanged: every rank normalises its own minibatch."""

    def __init__(self, model, world):
        import torch.distributed as dist
        self.world, self.dist = world, dist
        params = [p for p in model.parameters() if p.requires_grad]
        total = sum(p.numel() for p in params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=params[0].device)
        off = 0
        for p in params:
            p.grad = self.flat[off:off + p.numel()].as_strided(p.shape, p.stride())   # same memory format as p
            off += p.numel()
        if world > 1:
            for p in params:                         # identical by construction (seeded); make it explicit
                dist.broadcast(p.data, 0)

    def zero(self):
        self.flat.zero_()

    def reduce(self):
        if self.world > 1:
            if self.flat.is_cuda:
                self.dist.all_reduce(self.flat, op=self.dist.ReduceOp.AVG)
            else:
                self.dist.all_reduce(self.flat, op=self.dist.ReduceOp.SUM)
                self.flat.div_(self.world)


def train_step(model, mec, opt, images, labels, sync=None, head=None):
    """head: dwt_b200.HeadLoss -- the same NLL + lambda*MEC as one fused launch (GPU arm)."""
    if sync is None:
        opt.zero_grad(set_to_none=True)
    else:
        sync.zero()
    logits = model(images)
    if head is not None:
        loss = head(logits, labels)
    else:
        src, tgt, aug = torch.split(logits, logits.shape[0] // 3, dim=0)
        loss = F.nll_loss(F.log_softmax(src, dim=1), labels) + LAMBDA_MEC * mec(tgt, aug)
    loss.backward()
    if sync is not None:
        sync.reduce()
    opt.step()
    return loss


# ----------------------------------------------------------------------------------------- arms
def run_reference(args):
    """The reference's CPU path: harness model + CPU port of the reference layers, all host threads."""
    import oracle.torch_port as port
    from harness.synth import synth_batch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    torch.set_num_threads(cores)
    per_domain = args.cpu_per_domain
    dev = torch.device("cpu")
    model = build_model(port, dev, "modules")
    opt = make_optimizer(model)
    mec = port.MinEntropyConsensusLoss(NUM_CLASSES, dev)
    images, labels = synth_batch(seed=2, per_domain=per_domain)
    for _ in range(args.warmup):
        train_step(model, mec, opt, images, labels)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        train_step(model, mec, opt, images, labels)
    dt = (time.perf_counter() - t0) / args.steps
    val = 3 * per_domain / dt
    sample = f"{args.steps} steps of {3 * per_domain} images (3x{per_domain}), same model/step as the GPU arm"
    print(json.dumps({
        "impl": "reference", "metric": "ResNet-50-DWT images/sec fwd+bwd", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, per_domain=per_domain, site_mode="modules"),
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, per_domain, site_mode):
    return {"workload": "ResNet-50-DWT synthetic Office-Home 224x224, train step = fwd + NLL + 0.1*MEC + bwd + SGD",
            "per_domain_batch": per_domain, "images_per_gpu": 3 * per_domain, "global_images": 3 * per_domain * args.gpus,
            "group_size": 4, "site_mode": site_mode, "memory_format": getattr(args, "memory_format", "nchw"),
            "parallelism": f"dp{args.gpus}", "grad_sync": "one flat NCCL all-reduce (AVG) per step",
            "l2": "no explicit flush: per-step working set (activations) is tens of GB >> 126 MB L2",
            "launch": "CUDA-graph replay of the whole step" if getattr(args, "cuda_graph", False) and getattr(args, "impl", "ours") == "ours" else "eager"}


def cpu_baseline(args):
    """Bounded CPU sample on rank 0: a few steps of the CPU port at a small per-domain batch."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--cpu-per-domain", str(args.cpu_per_domain)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="0", WORLD_SIZE="1")
    for k in ("LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600).stdout.strip().splitlines()
        return json.loads(out[-1])["cpu_baseline"]
    except Exception as e:                                   # the GPU number must not die with the CPU leg
        return {"value": None, "unit": "images/s", "cores": host_cores(), "kind": "port", "sample": f"failed: {e}"}


def timed_loop(step_fn, steps, device, distributed):
    import torch.distributed as dist
    if distributed:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_fn()
    e1.record()
    torch.cuda.synchronize(device)
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if distributed:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def finish(distributed, device):
    """Leave a multi-rank run without tearing NCCL down: destroying a process group whose collectives were
    captured into a live CUDA graph can dead-lock, and a benchmark that hangs after printing its line is worse
    than one that skips destructors."""
    sys.stdout.flush()
    sys.stderr.flush()
    if distributed:
        torch.cuda.synchronize(device)
        os._exit(0)


def run_ours(args):
    import torch.distributed as dist
    import dwt_b200
    from dwt_b200 import _native
    from harness.synth import synth_batch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if distributed:
        # NCCL prints its version banner on STDOUT at communicator creation; keep stdout for the JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize(device)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.backends.cudnn.benchmark = True
    _native.lib()                                           # fail loudly if the extension is missing

    if args.workload == "microbench":
        return run_microbench(args, device, rank)

    nhwc = args.memory_format == "nhwc"
    model = build_model(dwt_b200, device, args.site_mode, channels_last=nhwc)
    net = model
    sync = FlatGradAllReduce(model, world) if distributed else None    # one GPU: plain autograd .grad tensors
    opt = make_optimizer(model)
    mec = dwt_b200.MinEntropyConsensusLoss(NUM_CLASSES, device)
    head = dwt_b200.HeadLoss(NUM_CLASSES, LAMBDA_MEC)
    B = args.per_domain
    host_images, host_labels = synth_batch(seed=100 + rank, per_domain=B)
    if nhwc:
        host_images = host_images.contiguous(memory_format=torch.channels_last)
    host_images, host_labels = host_images.pin_memory(), host_labels.pin_memory()
    images, labels = host_images.to(device), host_labels.to(device)
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    def step_resident():
        train_step(net, mec, opt, images, labels, sync, head)

    def step_e2e():
        im = host_images.to(device, non_blocking=True)
        lb = host_labels.to(device, non_blocking=True)
        loss = train_step(net, mec, opt, im, lb, sync, head)
        loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream(device).synchronize()     # the user reads the loss every step

    for _ in range(args.warmup):
        step_resident()
    # Pass 1 (eager): every kernel launch of the library is bracketed by CUDA events -> per-kernel roofline.
    n0 = _native.launch_count()
    with ClockSampler(local) as clocks_eager:
        _native.profile_begin()
        ms_eager = timed_loop(step_resident, args.steps, device, distributed)
        prof_sites = _native.profile_end()
    prof = _native.by_family(prof_sites)
    launches = (_native.launch_count() - n0) // args.steps

    # Pass 2: the same step captured once into a CUDA graph and replayed (the step is ~1400 launches, the
    # host launch path is as long as the GPU work).  This is the reported `value`; identical kernels and work.
    graph, static_loss = None, None
    if args.cuda_graph:
        try:
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(3):
                    step_resident()
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = train_step(net, mec, opt, images, labels, sync, head)
            torch.cuda.synchronize(device)
        except Exception as e:                      # capture is an optimisation, never a requirement
            graph = None
            sys.stderr.write(f"[bench] CUDA-graph capture unavailable ({type(e).__name__}: {e}); timing eagerly\n")
            torch.cuda.synchronize(device)

    if graph is not None:
        def step_timed():
            graph.replay()

        # e2e: the batch of step i+1 crosses PCIe (pinned host -> staging buffer, side stream) while the graph of
        # step i runs; each step still performs exactly one H2D of a full batch and one D2H of the loss.
        copy_stream = torch.cuda.Stream(device)
        stage_images, stage_labels = torch.empty_like(images), torch.empty_like(labels)
        with torch.cuda.stream(copy_stream):
            stage_images.copy_(host_images, non_blocking=True)
            stage_labels.copy_(host_labels, non_blocking=True)

        def step_e2e():
            cur = torch.cuda.current_stream(device)
            cur.wait_stream(copy_stream)                     # this step's batch has landed
            images.copy_(stage_images, non_blocking=True)    # device-to-device into the graph's static input
            labels.copy_(stage_labels, non_blocking=True)
            copy_stream.wait_stream(cur)                     # staging buffer is free again
            with torch.cuda.stream(copy_stream):
                stage_images.copy_(host_images, non_blocking=True)
                stage_labels.copy_(host_labels, non_blocking=True)
            graph.replay()
            loss_host.copy_(static_loss.detach(), non_blocking=True)
            cur.synchronize()                                # the user reads the loss every step
    else:
        step_timed = step_resident
    for _ in range(2):
        step_timed()
    with ClockSampler(local) as clocks:
        ms = timed_loop(step_timed, args.steps, device, distributed)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed_loop(step_e2e, args.steps, device, distributed)

    # e2e from RESIZED UINT8 images (what the reference's loader workers hold before RandomCrop): B source + B target
    # images of 256x256x3 cross PCIe (4.6x fewer bytes than the three float views) and one dwt_augment_pair launch per
    # domain writes crop / flip / cv2-exact affine / normalise straight into the graph's static input.
    e2e_u8 = None
    if graph is not None and not distributed:
        try:
            import numpy as np
            from dwt_b200 import PairedAugment, draw_params
            aug = PairedAugment(crop=224)
            rng = np.random.default_rng(1234)
            host_u8 = torch.randint(0, 256, (2 * B, 256, 256, 3), dtype=torch.uint8).pin_memory()
            host_par = [{k: v.pin_memory() for k, v in draw_params(B, 256, 224, rng).items()} for _ in range(2)]
            dev_u8 = torch.empty(host_u8.shape, dtype=torch.uint8, device=device)
            dev_par = [{k: torch.empty(v.shape, dtype=v.dtype, device=device) for k, v in hp.items()} for hp in host_par]
            fmt_cl = args.memory_format == "nhwc"

            def upload():
                dev_u8.copy_(host_u8, non_blocking=True)
                stage_labels.copy_(host_labels, non_blocking=True)
                for dp, hp in zip(dev_par, host_par):
                    for k in dp:
                        dp[k].copy_(hp[k], non_blocking=True)

            with torch.cuda.stream(copy_stream):
                upload()

            def step_u8():
                cur = torch.cuda.current_stream(device)
                cur.wait_stream(copy_stream)
                aug(dev_u8[:B], crop_plain=dev_par[0]["crop_plain"], out_plain=images[:B], want_aug=False,
                    channels_last=fmt_cl)                                                   # source view
                aug(dev_u8[B:], out_plain=images[B:2 * B], out_aug=images[2 * B:], channels_last=fmt_cl, **dev_par[1])
                labels.copy_(stage_labels, non_blocking=True)
                copy_stream.wait_stream(cur)
                with torch.cuda.stream(copy_stream):
                    upload()
                graph.replay()
                loss_host.copy_(static_loss.detach(), non_blocking=True)
                cur.synchronize()

            for _ in range(2):
                step_u8()
            ms_u8 = timed_loop(step_u8, args.steps, device, False)
            e2e_u8 = {"value": 3 * B * args.steps / (ms_u8 / 1e3), "unit": "images/s", "ms_per_step": ms_u8 / args.steps,
                      "h2d_bytes_per_step": host_u8.numel() + host_labels.numel() * 8 + sum(v.numel() * v.element_size() for hp in host_par for v in hp.values()),
                      "d2h_bytes_per_step": 4,
                      "input": "2B resized uint8 images 256x256x3; crop/flip/affine/normalise on the GPU (dwt_augment_pair)"}
        except Exception as e:                      # an extra measurement, never a requirement
            sys.stderr.write(f"[bench] uint8-input e2e skipped ({type(e).__name__}: {e})\n")

    per_gpu = 3 * B
    value = per_gpu * world * args.steps / (ms / 1e3)
    e2e = per_gpu * world * args.steps / (ms_e2e / 1e3)
    if rank != 0:
        finish(distributed, device)
        return
    peak, peak_src = measured_peaks()
    if args.sites_out:
        rows = [dict(zip(("kernel", "C", "HW", "GS", "D", "N"), k.split("|")), launches=v["launches"],
                     us_per_launch=1e3 * v["ms"] / v["launches"], gbs=v["bytes"] / max(v["ms"], 1e-9) / 1e6,
                     frac_of_peak=v["bytes"] / max(v["ms"], 1e-9) / 1e6 / peak, ms_per_step=v["ms"] / args.steps)
                for k, v in sorted(prof_sites.items(), key=lambda kv: -kv[1]["ms"])]
        os.makedirs(os.path.dirname(os.path.abspath(args.sites_out)), exist_ok=True)
        json.dump(rows, open(args.sites_out, "w"), indent=1)
    fams = {k: dict(v, gbs=(v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else None),
                    us_per_launch=(1e3 * v["ms"] / v["launches"] if v["launches"] else None),
                    share_of_step=v["ms"] / ms_eager) for k, v in prof.items()}
    dom = max((k for k in fams if fams[k]["bytes"] > 0), key=lambda k: fams[k]["ms"], default=None)
    roof = None
    traffic_tab = {}
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "traffic_r01.json")))
    except Exception:
        pass
    if dom is not None:
        f = fams[dom]
        ratio = traffic_tab.get("traffic_over_algorithmic", {}).get(dom)
        roof = {"bound": "hbm", "kernel": dom, "achieved": f["gbs"], "peak": peak, "unit": "GB/s",
                "frac": f["gbs"] / peak,
                # DRAM bytes per launch: this run's algorithmic bytes per launch x the traffic/algorithmic ratio of the
                # committed `ncu --set full` capture of the same kernel (never measured under the profiler here)
                "traffic": (ratio * f["bytes"] / f["launches"]) if ratio else None,
                "traffic_source": traffic_tab.get("source") if ratio else None, "peak_source": peak_src,
                "launches_timed": f["launches"], "avg_launch_us": f["us_per_launch"],
                "algorithmic_bytes_per_launch": f["bytes"] / f["launches"],
                "all_norm_kernels_share_of_step": sum(v["ms"] for v in fams.values()) / ms_eager,
                "timed_in": "eager pass of this run (launches bracketed by CUDA events; graph replay hides them)"}
    out = {
        "metric": "ResNet-50-DWT images/sec fwd+bwd", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, per_domain=B, site_mode=args.site_mode),
        "e2e": {"value": e2e, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": host_images.numel() * 4 + host_labels.numel() * 8, "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "clocks": clocks.summary(), "roofline": roof, "kernels": fams,
        "e2e_uint8_input": e2e_u8,
        "cuda_graph": graph is not None, "eager_ms_per_step": ms_eager / args.steps,
        "conv_math": "cuDNN, TF32 allowed (torch default) -- convolutions are not part of the hot path",
    }
    if args.cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(out))
    finish(distributed, device)


def run_microbench(args, device, rank):
    """BASELINE.json configs[1]: WTransform2d N=256 C=256 H=W=56 group_size=64, fwd+bwd."""
    import dwt_b200
    from dwt_b200 import _native
    N, C, H, gs = args.micro_n, 256, 56, args.micro_gs
    torch.manual_seed(0)
    mix = torch.randn(C, C, device=device) / C ** 0.5 + torch.eye(C, device=device)
    # .contiguous(): einsum returns a permuted view, and a strided input would add a layout copy (and its backward) of
    # the 822 MB tensor to every step -- 0.5 ms that is not the layer's
    x = (torch.einsum("dc,nchw->ndhw", mix, torch.randn(N, C, H, H, device=device)) + 2.0).contiguous().requires_grad_(True)
    dy = torch.randn(N, C, H, H, device=device)
    m = dwt_b200.WTransform2d(C, gs).to(device).train()

    def step():
        y = m(x)
        torch.autograd.grad(y, x, dy)

    for _ in range(args.warmup):
        step()
    graph = None
    if args.cuda_graph:                              # the four launches as one graph: no host time between them
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream(device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        graph.replay()
    with ClockSampler(device.index) as clocks:
        _native.profile_begin()
        eager_ms = timed_loop(step, args.steps, device, False)
        prof = _native.by_family(_native.profile_end())
        ms = timed_loop(graph.replay, args.steps, device, False) if graph is not None else eager_ms
    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    elems = N * C * H * H
    total_bytes = 32.0 * elems                       # 12 B/elem forward + 20 B/elem backward
    fams = {k: dict(v, gbs=v["bytes"] / (v["ms"] * 1e-3) / 1e9, us_per_launch=1e3 * v["ms"] / v["launches"])
            for k, v in prof.items() if v["ms"] > 0}
    gbs = total_bytes * args.steps / (ms * 1e-3) / 1e9
    ratio, ratio_src = None, None
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "traffic_r01.json")))["tensor_core_path"]
        if gs >= 8:
            ratio, ratio_src = tab["traffic_over_algorithmic"]["fwd+bwd (all launches)"], tab["source"]
    except Exception:
        pass
    print(json.dumps({
        "metric": "WTransform2d fwd+bwd microbench", "value": args.steps / (ms * 1e-3), "unit": "iterations/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "eager_ms_per_step": eager_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"WTransform2d N={N} C={C} H=W={H} group_size={gs} fwd+bwd",
                   "l2": f"tensor of {elems * 4 / 1e6:.0f} MB > 126 MB L2",
                   "launch": "cuda-graph replay" if graph is not None else "eager"},
        "roofline": {"bound": "hbm", "kernel": "fwd+bwd (4 launches)", "achieved": gbs, "peak": peak, "unit": "GB/s",
                     "frac": gbs / peak, "traffic": ratio * total_bytes if ratio else None, "traffic_source": ratio_src,
                     "algorithmic_bytes_per_step": total_bytes, "peak_source": peak_src},
        "kernels": fams, "clocks": clocks.summary(),
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["resnet", "microbench"], default="resnet")
    ap.add_argument("--per-domain", type=int, default=64)
    ap.add_argument("--site-mode", choices=["fused", "modules"], default="fused")
    ap.add_argument("--memory-format", choices=["nchw", "nhwc"], default="nhwc",
                    help="activation layout of the GPU arm (nhwc = torch.channels_last end to end)")
    ap.add_argument("--cpu-per-domain", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false",
                    help="time the eager step instead of a CUDA-graph replay of it")
    ap.add_argument("--sites-out", default="", help="write the per-site kernel table (JSON) here")
    ap.add_argument("--micro-n", type=int, default=256)
    ap.add_argument("--micro-gs", type=int, default=64)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
