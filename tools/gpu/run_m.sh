#!/bin/bash
# round-2 GPU call M (4 GPUs): gradient gather / segment modes of the data-parallel step
mkdir -p gpurun_out/m
i=0
for mode in "copy 1" "copy 3" "accumulate 3" "copy 2"; do
  set -- $mode; i=$((i+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2953$i bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --grad-gather $1 --grad-segments $2 > gpurun_out/m/bench_4gpu_$1_$2.json 2> gpurun_out/m/bench_4gpu_$1_$2.err; echo "4gpu $1 $2 rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/m/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n", d.get("n_gpus"), "ms/step", round(d["ms_per_step"],3), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1))
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
