#!/bin/bash
# round-2 GPU call AC (1 GPU): gradient sum of the two uses of a block input inside the site's backward kernels (fork_for_sum)
mkdir -p gpurun_out/ac
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -s -k "fork or resnet_224 or resnet_tiny" > gpurun_out/ac/pytest_fork.log 2>&1; echo "pytest fork rc=$?"
grep -E "passed|failed|error|fork vs add" gpurun_out/ac/pytest_fork.log | tail -5
for r in fork nofork fork nofork; do
flag=""; [ $r = nofork ] && flag="--no-grad-fork"
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline $flag > gpurun_out/ac/bench_$r.json 2> gpurun_out/ac/bench_$r.err; echo "bench $r rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ac/bench_$r.json").read().strip().splitlines()[-1])
    print("$r ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "norm path", d["roofline"]["norm_path"]["frac"], {k: (round(v["gbs"] or 0), round(v["us_per_launch"],1)) for k,v in d["kernels"].items() if "bwd" in k})
except Exception as e: print("parse failed", e); print(open("gpurun_out/ac/bench_$r.err").read()[-1500:])
PY
done
