#!/bin/bash
# round-2 GPU call H (2 GPUs): lane-parallel finalize + prefetch: parity, step time; data-parallel gather modes; microbench
mkdir -p gpurun_out/h
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/h/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/h/pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --sites-out gpurun_out/h/sites.json > gpurun_out/h/bench_1gpu.json 2> gpurun_out/h/bench_1gpu.err; echo "1gpu rc=$?"
for mode in copy accumulate; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --grad-gather $mode > gpurun_out/h/bench_2gpu_$mode.json 2> gpurun_out/h/bench_2gpu_$mode.err; echo "2gpu $mode rc=$?"
done
timeout 600 python bench.py --workload microbench --no-cpu-baseline > gpurun_out/h/micro.json 2> gpurun_out/h/micro.err; echo "micro rc=$?"
python - <<'PY'
import json
for n in ("bench_1gpu","bench_2gpu_copy","bench_2gpu_accumulate"):
    try:
        d=json.loads(open(f"gpurun_out/h/{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "eager", round(d["eager_ms_per_step"],3), "launches", d["gpu_launches"])
        if n=="bench_1gpu":
            for k,v in sorted(d["kernels"].items()): print("   %-18s %6.0f GB/s  %8.1f us/launch  share %.3f"%(k, v["gbs"] or 0, v["us_per_launch"], v["share_of_step"]))
    except Exception as e:
        print(n, "failed", e)
try:
    mb=json.loads(open("gpurun_out/h/micro.json").read().strip().splitlines()[-1])
    print("micro", mb["ms_per_step"], mb["eager_ms_per_step"], mb["roofline"]["frac"], "vs ref gpu", mb.get("vs_reference_gpu"))
    for k,v in sorted(mb.get("kernels",{}).items()): print("   ",k, round(v["us_per_launch"],1), v.get("frac"))
except Exception as e: print("micro parse failed", e)
PY
