#!/bin/bash
# round-2 GPU call Y (1 GPU): tcgen05 instructions issued by one elected lane with warp-uniform operands (all three TC kernels)
mkdir -p gpurun_out/y
timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "tensor_core or run_to_run or non_pd or whitening_vs or microbench_shape or fused_triple or graph_capturable" > gpurun_out/y/pytest_tc.log 2>&1; echo "pytest tc rc=$?"
grep -E "passed|failed|error" gpurun_out/y/pytest_tc.log | tail -3
for q in 1 2; do
timeout 600 python bench.py --workload microbench --no-cpu-baseline > gpurun_out/y/micro_$q.json 2> gpurun_out/y/micro_$q.err; echo "micro $q rc=$?"
python - <<PY
import json
try:
    mb=json.loads(open("gpurun_out/y/micro_$q.json").read().strip().splitlines()[-1])
    print("micro $q", mb["ms_per_step"], mb["roofline"]["frac"], {k: round(v["us_per_launch"],1) for k,v in sorted(mb.get("kernels",{}).items())})
except Exception as e: print("micro parse failed", e)
PY
done
timeout 300 python tools/gpu/prof/gram_timeline.py > gpurun_out/y/gram_timeline.log 2>&1; echo "timeline rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gram|tc_apply" -s 3 -c 3 -o gpurun_out/y/prof_tc python tools/micro_once.py 256 > gpurun_out/y/ncu_gram.log 2>&1; echo "ncu rc=$?"
