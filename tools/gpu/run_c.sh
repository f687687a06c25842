#!/bin/bash
# round-2 GPU call C: split-precision Gram kernel (tc_gram) parity + timing, fused sites for gs >= 8, stem-pad experiment
mkdir -p gpurun_out/c
timeout 600 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -p no:cacheprovider -k "tensor_core or non_pd" > gpurun_out/c/pytest_tc.log 2>&1; echo "pytest tc rc=$?"
grep -E "passed|failed|error" gpurun_out/c/pytest_tc.log | tail -3
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c/pytest_all.log 2>&1; echo "pytest all rc=$?"
tail -15 gpurun_out/c/pytest_all.log
timeout 600 python bench.py --workload microbench --no-cpu-baseline > gpurun_out/c/micro.json 2> gpurun_out/c/micro.err; echo "micro rc=$?"
for pad in 0 4 8; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --stem-pad $pad > gpurun_out/c/bench_pad$pad.json 2> gpurun_out/c/bench_pad$pad.err; echo "bench pad$pad rc=$?"
done
python - <<'PY'
import json
try:
    mb=json.loads(open("gpurun_out/c/micro.json").read().strip().splitlines()[-1])
    print("micro", mb["ms_per_step"], mb["eager_ms_per_step"], mb["roofline"]["frac"], "vs ref gpu", mb.get("vs_reference_gpu"))
    for k,v in sorted(mb.get("kernels",{}).items()): print("   ",k, round(v["us_per_launch"],1), v.get("frac"))
except Exception as e: print("micro parse failed", e)
for pad in (0,4,8):
    try:
        d=json.loads(open(f"gpurun_out/c/bench_pad{pad}.json").read().strip().splitlines()[-1])
        print("pad",pad,"ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], "status", d.get("status_word"))
    except Exception as e: print("pad",pad,"failed", e)
PY
