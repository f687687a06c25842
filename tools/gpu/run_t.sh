#!/bin/bash
# round-2 GPU call T (N GPUs, N = $1): the data-parallel step of the final build, default bench flags
N=${1:-2}
mkdir -p gpurun_out/t
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/t/bench_${N}gpu.json 2> gpurun_out/t/bench_${N}gpu.err; echo "bench ${N}gpu rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/t/bench_1gpu_same_box_${N}.json 2> gpurun_out/t/bench_1gpu_same_box_${N}.err; echo "bench 1gpu rc=$?"
python - <<PY
import json
for f in ("gpurun_out/t/bench_${N}gpu.json", "gpurun_out/t/bench_1gpu_same_box_${N}.json"):
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n_gpus", r["n_gpus"], "ms/step", r["ms_per_step"], "img/s", r["value"], "e2e", r["e2e"]["value"])
    except Exception as e: print(f, "parse failed", e)
PY
tail -3 gpurun_out/t/bench_${N}gpu.err
