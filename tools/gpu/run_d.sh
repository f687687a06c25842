#!/bin/bash
# round-2 GPU call D: reworked tc_gram (ones-row sums, branch-free split) parity + timing + ncu full captures
mkdir -p gpurun_out/d
timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_parity.py -m gpu -q -s -p no:cacheprovider -k "tensor_core or non_pd or whitening_vs or microbench_shape or fused_triple" > gpurun_out/d/pytest_tc.log 2>&1; echo "pytest tc rc=$?"
grep -E "passed|failed|error" gpurun_out/d/pytest_tc.log | tail -3
timeout 600 python bench.py --workload microbench --no-cpu-baseline > gpurun_out/d/micro.json 2> gpurun_out/d/micro.err; echo "micro rc=$?"
python - <<'PY'
import json
try:
    mb=json.loads(open("gpurun_out/d/micro.json").read().strip().splitlines()[-1])
    print("micro", mb["ms_per_step"], mb["eager_ms_per_step"], mb["roofline"]["frac"], "vs ref gpu", mb.get("vs_reference_gpu"))
    for k,v in sorted(mb.get("kernels",{}).items()): print("   ",k, round(v["us_per_launch"],1), v.get("frac"))
except Exception as e: print("micro parse failed", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_gram|tc_apply|tc_contract|fwd_factor|bwd_coef|partial_reduce' -s 8 -c 8 -o gpurun_out/d/prof_tc python tools/micro_once.py 256 > gpurun_out/d/ncu_tc.log 2>&1; echo "ncu tc rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cl_(stats|apply|bwd_reduce|bwd_apply)_kernel|vec_reduce|finalize' -s 16 -c 8 -o gpurun_out/d/prof_cl python tools/site_probe.py 0 cl > gpurun_out/d/ncu_cl.log 2>&1; echo "ncu cl rc=$?"
ls -la gpurun_out/d
