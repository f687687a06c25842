#!/bin/bash
# round-2 GPU call U (1 GPU): mbarrier try_wait with / without the suspend-time hint (A/B on one box), TC tests
mkdir -p gpurun_out/u
timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "tensor_core or run_to_run or non_pd or whitening_vs or microbench_shape or fused_triple or graph_capturable" > gpurun_out/u/pytest_tc.log 2>&1; echo "pytest tc rc=$?"
grep -E "passed|failed|error" gpurun_out/u/pytest_tc.log | tail -3
for v in hint nohint hint nohint; do
if [ $v = nohint ]; then export DWT_B200_LIB=$PWD/tools/gpu/prof/libdwt_b200_nohint.so; else unset DWT_B200_LIB; fi
timeout 600 python bench.py --workload microbench --no-cpu-baseline > gpurun_out/u/micro_$v.json 2> gpurun_out/u/micro_$v.err; echo "micro $v rc=$?"
python - <<PY
import json
try:
    mb=json.loads(open("gpurun_out/u/micro_$v.json").read().strip().splitlines()[-1])
    print("micro $v", mb["ms_per_step"], mb["roofline"]["frac"], {k: round(v["us_per_launch"],1) for k,v in sorted(mb.get("kernels",{}).items())})
except Exception as e: print("micro parse failed", e)
PY
done
