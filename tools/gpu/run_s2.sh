#!/bin/bash
# round-2 GPU call S (1 GPU): final build of round 2 (s2d stem, quarter-per-warp tc_gram, blocked dense sweep): parity suite, smoke, default bench line, reference arm, launch list
mkdir -p gpurun_out/s2
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/s2/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/s2/pytest_all.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/s2/smoke.log 2>&1; echo "smoke rc=$?"
( time timeout 900 python bench.py --steps 20 --warmup 5 --sites-out gpurun_out/s2/sites.json > gpurun_out/s2/bench_full.json 2> gpurun_out/s2/bench_full.err ) 2> gpurun_out/s2/bench_full.time; echo "bench rc=$?"; tail -3 gpurun_out/s2/bench_full.time
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s2/bench_reference_arm.json 2> gpurun_out/s2/bench_reference_arm.err; echo "ref arm rc=$?"
echo skip-launch-list
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/s2/bench_full.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "eager", d["eager_ms_per_step"], "status", d.get("status_word"))
    print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("norm_path"))
    print("ref_on_gpu", {k:d["reference_on_gpu"].get(k) for k in ("value","ms_per_step","speedup_value","speedup_e2e","error")})
    mb=d.get("microbench",{})
    print("micro", {k:mb.get(k) for k in ("ms_per_step","eager_ms_per_step","vs_reference_gpu","error")}, mb.get("roofline",{}).get("frac"))
    for k,v in sorted(mb.get("kernels",{}).items()): print("   ",k, round(v["us_per_launch"],1), v.get("frac"))
    print("micro cpu", mb.get("cpu_baseline")); print("cpu", d.get("cpu_baseline")); print("u8", d.get("e2e_uint8_input"))
    for k,v in sorted(d["kernels"].items()): print("   %-18s %6.0f GB/s  %8.1f us/launch  share %.3f"%(k, v["gbs"] or 0, v["us_per_launch"], v["share_of_step"]))
except Exception as e:
    print("parse failed", e)
PY
