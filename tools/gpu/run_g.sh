#!/bin/bash
# round-2 GPU call G (1 GPU): row-CTA max-pool kernels + warp-per-column finalize (vec_reduce merged): parity, step time, launch list
mkdir -p gpurun_out/g
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/g/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/g/pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --sites-out gpurun_out/g/sites.json > gpurun_out/g/bench_1gpu.json 2> gpurun_out/g/bench_1gpu.err; echo "1gpu rc=$?"
python - <<'PY'
import json
for n in ("bench_1gpu",):
    try:
        d=json.loads(open(f"gpurun_out/g/{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "eager", round(d["eager_ms_per_step"],3))
        for k,v in sorted(d["kernels"].items()): print("   %-18s %6.0f GB/s  %8.1f us/launch  share %.3f"%(k, v["gbs"] or 0, v["us_per_launch"], v["share_of_step"]))
    except Exception as e:
        print(n, "failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/g/launches_ours.csv python tools/one_step.py > gpurun_out/g/ncu_ours.log 2>&1; echo "ncu ours rc=$?"
timeout 300 python __graft_entry__.py --smoke > gpurun_out/g/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/g/smoke.log
