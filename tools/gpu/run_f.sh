#!/bin/bash
# round-2 GPU call F (2 GPUs): data-parallel step with the segmented, overlapped gradient all-reduce vs one collective
mkdir -p gpurun_out/f
timeout 300 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -p no:cacheprovider -k "maxpool" > gpurun_out/f/pytest_pool.log 2>&1; echo "pytest pool rc=$?"; tail -2 gpurun_out/f/pytest_pool.log
for seg in 3 1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --grad-segments $seg > gpurun_out/f/bench_2gpu_seg$seg.json 2> gpurun_out/f/bench_2gpu_seg$seg.err; echo "2gpu seg$seg rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/f/bench_1gpu.json 2> gpurun_out/f/bench_1gpu.err; echo "1gpu rc=$?"
python - <<'PY'
import json
for n in ("bench_2gpu_seg3","bench_2gpu_seg1","bench_1gpu"):
    try:
        d=json.loads(open(f"gpurun_out/f/{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "n_gpus", d["n_gpus"])
        if n=="bench_1gpu":
            for k,v in sorted(d["kernels"].items()): print("   %-18s %6.0f GB/s  %8.1f us/launch  share %.3f"%(k, v["gbs"] or 0, v["us_per_launch"], v["share_of_step"]))
    except Exception as e:
        print(n, "failed", e)
        try: print(open(f"gpurun_out/f/{n}.err").read()[-1500:])
        except Exception: pass
PY
