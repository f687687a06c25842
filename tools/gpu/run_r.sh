#!/bin/bash
# round-2 GPU call R (1 GPU): space-to-depth stem (harness option) -- parity at 224 x 224 and the step time with / without
mkdir -p gpurun_out/r
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -p no:cacheprovider -k "resnet_224" > gpurun_out/r/pytest_224.log 2>&1; echo "pytest 224 rc=$?"
grep -E "passed|failed|error" gpurun_out/r/pytest_224.log | tail -3
for st in s2d direct s2d; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --stem $st > gpurun_out/r/bench_$st.json 2> gpurun_out/r/bench_$st.err; echo "bench $st rc=$?"
python - <<PY
import json
try:
    r=json.loads(open("gpurun_out/r/bench_$st.json").read().strip().splitlines()[-1])
    print("bench $st ms/step", r["ms_per_step"], "img/s", r["value"], "e2e", r["e2e"]["value"])
except Exception as e: print("parse failed", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r/launches_s2d.csv python tools/one_step.py --stem s2d > gpurun_out/r/ncu_step.log 2>&1; echo "ncu rc=$?"
