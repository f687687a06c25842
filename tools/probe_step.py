"""Scratch: host-core probe + torch.profiler kernel table of one training step (analysis only,
never a reported number)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dwt-domain-adaptation_b200"), ROOT]
import torch

def host():
    print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        if os.path.exists(p): print(p, open(p).read().strip())
    x = torch.randn(8, 64, 112, 112); w = torch.randn(64, 64, 3, 3)
    for nt in (4, 8, 16, 32, 64, 128):
        torch.set_num_threads(nt)
        torch.nn.functional.conv2d(x, w, padding=1)
        t = time.perf_counter()
        for _ in range(3): torch.nn.functional.conv2d(x, w, padding=1)
        print("threads", nt, "conv ms", (time.perf_counter() - t) / 3 * 1e3, flush=True)

def gpu(site_mode):
    import bench, dwt_b200
    from harness.synth import synth_batch
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    model = bench.build_model(dwt_b200, dev, site_mode, channels_last=True)
    opt = bench.make_optimizer(model); mec = dwt_b200.MinEntropyConsensusLoss(65, dev)
    im, lb = synth_batch(3, 64); im, lb = im.to(dev).contiguous(memory_format=torch.channels_last), lb.to(dev)
    for _ in range(3): bench.train_step(model, mec, opt, im, lb)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2): bench.train_step(model, mec, opt, im, lb)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))

if __name__ == "__main__":
    if "host" in sys.argv: host()
    if "gpu" in sys.argv: gpu(sys.argv[sys.argv.index("gpu") + 1])
