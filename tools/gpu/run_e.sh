#!/bin/bash
# round-2 GPU call E: tc_gram with homogeneous hi/lo warps + 8-lane finalize kernels: parity, timings, short ncu
mkdir -p gpurun_out/e
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/e/pytest_all.log 2>&1; echo "pytest all rc=$?"
tail -6 gpurun_out/e/pytest_all.log
for i in 1 2; do
timeout 600 python bench.py --workload microbench --no-cpu-baseline > gpurun_out/e/micro$i.json 2> gpurun_out/e/micro$i.err; echo "micro rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --sites-out gpurun_out/e/sites.json > gpurun_out/e/bench.json 2> gpurun_out/e/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for i in (1,2):
  try:
    mb=json.loads(open(f"gpurun_out/e/micro{i}.json").read().strip().splitlines()[-1])
    print("micro", mb["ms_per_step"], mb["eager_ms_per_step"], mb["roofline"]["frac"], "vs ref gpu", mb.get("vs_reference_gpu"))
    for k,v in sorted(mb.get("kernels",{}).items()): print("   ",k, round(v["us_per_launch"],1), v.get("frac"))
  except Exception as e: print("micro parse failed", e)
try:
    d=json.loads(open("gpurun_out/e/bench.json").read().strip().splitlines()[-1])
    print("ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "eager", round(d["eager_ms_per_step"],3), "launches", d["gpu_launches"], "status", d.get("status_word"))
    print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("norm_path"))
    for k,v in sorted(d["kernels"].items()): print("   %-18s %6.0f GB/s  %8.1f us/launch  share %.3f"%(k, v["gbs"] or 0, v["us_per_launch"], v["share_of_step"]))
except Exception as e: print("bench parse failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_gram' -s 1 -c 1 -o gpurun_out/e/prof_gram python tools/micro_once.py 256 > gpurun_out/e/ncu_gram.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'finalize|vec_reduce' -s 16 -c 8 --csv --log-file gpurun_out/e/fin_times.csv python tools/site_probe.py 0 cl > /dev/null 2>&1; echo "ncu fin rc=$?"; grep -v "^==" gpurun_out/e/fin_times.csv | cut -d, -f5,15 | tail -8
