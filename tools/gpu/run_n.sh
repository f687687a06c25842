#!/bin/bash
# round-2 GPU call N (1 GPU): tc_gram with per-parity barrier rings; TC tests, microbench x2, ncu of tc_gram
mkdir -p gpurun_out/n
timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_parity.py -m gpu -q -s -p no:cacheprovider -k "tensor_core or run_to_run or non_pd or whitening_vs or microbench_shape or fused_triple or graph_capturable" > gpurun_out/n/pytest_tc.log 2>&1; echo "pytest tc rc=$?"
grep -E "passed|failed|error" gpurun_out/n/pytest_tc.log | tail -3
for i in 1 2; do
timeout 600 python bench.py --workload microbench --no-cpu-baseline > gpurun_out/n/micro$i.json 2> gpurun_out/n/micro$i.err; echo "micro rc=$?"
done
python - <<'PY'
import json
for i in (1,2):
  try:
    mb=json.loads(open(f"gpurun_out/n/micro{i}.json").read().strip().splitlines()[-1])
    print("micro", mb["ms_per_step"], mb["eager_ms_per_step"], mb["roofline"]["frac"], "vs ref gpu", mb.get("vs_reference_gpu"))
    for k,v in sorted(mb.get("kernels",{}).items()): print("   ",k, round(v["us_per_launch"],1), v.get("frac"))
  except Exception as e: print("micro parse failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_gram' -s 1 -c 1 -o gpurun_out/n/prof_gram python tools/micro_once.py 256 > gpurun_out/n/ncu_gram.log 2>&1; echo "ncu rc=$?"
