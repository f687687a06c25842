#!/bin/bash
# round-2 GPU call L (N GPUs): the driver's scaling command lines on the final build
N=${1:-4}
mkdir -p gpurun_out/l
for n in $N; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/l/bench_${n}gpu.json 2> gpurun_out/l/bench_${n}gpu.err; echo "${n}gpu rc=$?"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --impl reference --gpus $n --steps 2 --warmup 1 > gpurun_out/l/ref_${n}gpu.json 2> gpurun_out/l/ref_${n}gpu.err; echo "ref ${n}gpu rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/l/bench_1gpu.json 2> gpurun_out/l/bench_1gpu.err; echo "1gpu rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/l/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n", d.get("n_gpus"), "ms/step", round(d["ms_per_step"],3), "value", round(d["value"],1), d.get("impl",""))
    except Exception as e:
        print(f, "failed", e)
PY
