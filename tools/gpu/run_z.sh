#!/bin/bash
# round-2 GPU call Z (1 GPU): tile order of the tc_apply kernels -- contiguous range per CTA vs interleaved (DWT_TC_INTERLEAVE=1)
mkdir -p gpurun_out/z
DWT_TC_INTERLEAVE=1 timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "tensor_core or run_to_run or non_pd or whitening_vs or microbench_shape or fused_triple or graph_capturable" > gpurun_out/z/pytest_tc.log 2>&1; echo "pytest tc (interleave) rc=$?"
grep -E "passed|failed|error" gpurun_out/z/pytest_tc.log | tail -3
for v in 0 1 0 1; do
DWT_TC_INTERLEAVE=$v timeout 600 python bench.py --workload microbench --no-cpu-baseline > gpurun_out/z/micro_$v.json 2> gpurun_out/z/micro_$v.err; echo "micro interleave=$v rc=$?"
python - <<PY
import json
try:
    mb=json.loads(open("gpurun_out/z/micro_$v.json").read().strip().splitlines()[-1])
    print("interleave $v", mb["ms_per_step"], mb["roofline"]["frac"], {k: round(v["us_per_launch"],1) for k,v in sorted(mb.get("kernels",{}).items())})
except Exception as e: print("micro parse failed", e)
PY
done
