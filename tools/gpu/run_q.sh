#!/bin/bash
# round-2 GPU call Q (1 GPU): packed-add tc_gram transform, fused blocked factor+inverse in fwd_factor, tightened bwd_coef; full GPU suite
mkdir -p gpurun_out/q
timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_parity.py -m gpu -q -s -p no:cacheprovider  > gpurun_out/q/pytest_tc.log 2>&1; echo "pytest tc rc=$?"
grep -E "passed|failed|error" gpurun_out/q/pytest_tc.log | tail -3
for q in 1 2; do
timeout 600 python bench.py --workload microbench --no-cpu-baseline > gpurun_out/q/micro_$q.json 2> gpurun_out/q/micro_$q.err; echo "micro $q rc=$?"
python - <<PY
import json
try:
    mb=json.loads(open("gpurun_out/q/micro_$q.json").read().strip().splitlines()[-1])
    print("micro $q", mb["ms_per_step"], mb["roofline"]["frac"], {k: round(v["us_per_launch"],1) for k,v in sorted(mb.get("kernels",{}).items())})
except Exception as e: print("micro parse failed", e)
PY
done
timeout 300 python tools/gpu/prof/dense_clocks.py > gpurun_out/q/dense_clocks.log 2>&1; echo "clocks rc=$?"; tail -8 gpurun_out/q/dense_clocks.log
echo skip-ncu
