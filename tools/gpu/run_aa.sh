#!/bin/bash
# round-2 GPU call AA (1 GPU): 2x2-patch max-pool kernels for the stem geometry (backward, then forward)
mkdir -p gpurun_out/aa
timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "maxpool or resnet_224 or resnet_tiny or fork" > gpurun_out/aa/pytest_pool.log 2>&1; echo "pytest pool rc=$?"
grep -E "passed|failed|error" gpurun_out/aa/pytest_pool.log | tail -3
timeout 300 python __graft_entry__.py --smoke > gpurun_out/aa/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/aa/bench.json 2> gpurun_out/aa/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/aa/bench.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "worst", d["roofline"]["norm_path"].get("worst_family"), d["roofline"]["norm_path"].get("worst_frac"))
    for k,v in sorted(d["kernels"].items()):
        if "pool" in k: print("   %-18s %6.0f GB/s  %8.1f us/launch"%(k, v["gbs"] or 0, v["us_per_launch"]))
except Exception as e: print("parse failed", e)
PY
