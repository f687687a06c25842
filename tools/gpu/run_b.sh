#!/bin/bash
# round-2 GPU call B: new parity tests, the full default bench line, ncu launch lists (our step, the reference on GPU)
mkdir -p gpurun_out/b
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -p no:cacheprovider > gpurun_out/b/pytest_r2.log 2>&1; echo "pytest r2 rc=$?"
grep -E "passed|failed|error" gpurun_out/b/pytest_r2.log | tail -3
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/b/bench_full.json 2> gpurun_out/b/bench_full.err ) 2> gpurun_out/b/bench_full.time; echo "bench rc=$?"; tail -3 gpurun_out/b/bench_full.time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/b/launches_ours.csv python tools/one_step.py > gpurun_out/b/ncu_ours.log 2>&1; echo "ncu ours rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/b/launches_ref_micro.csv python tools/ref_once.py micro > gpurun_out/b/ncu_ref_micro.log 2>&1; echo "ncu ref micro rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/b/launches_ref_resnet.csv python tools/ref_once.py resnet > gpurun_out/b/ncu_ref_resnet.log 2>&1; echo "ncu ref resnet rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/b/bench_full.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
    print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("norm_path"))
    print("ref_on_gpu", d.get("reference_on_gpu"))
    mb=d.get("microbench",{})
    print("micro", {k:mb.get(k) for k in ("ms_per_step","eager_ms_per_step","vs_reference_gpu","error")}, mb.get("roofline",{}).get("frac"))
    for k,v in sorted(mb.get("kernels",{}).items()): print("   ",k, round(v["us_per_launch"],1), v.get("frac"))
    print("micro ref", mb.get("reference_on_gpu"), mb.get("cpu_baseline"))
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("parse failed", e)
PY
