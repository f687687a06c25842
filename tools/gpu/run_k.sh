#!/bin/bash
# round-2 GPU call K (1 GPU): programmatic dependent launch on the channels-last chains, A/B on one box
mkdir -p gpurun_out/k
DWT_PDL=1 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/k/pytest_pdl.log 2>&1; echo "pytest pdl rc=$?"; tail -3 gpurun_out/k/pytest_pdl.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "channels_last or fused or resnet or residual" > gpurun_out/k/pytest_nopdl.log 2>&1; echo "pytest nopdl rc=$?"; tail -2 gpurun_out/k/pytest_nopdl.log
for rep in 1 2; do for pdl in 0 1; do
  DWT_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/k/bench_pdl${pdl}_$rep.json 2> gpurun_out/k/bench_pdl${pdl}_$rep.err; echo "bench pdl$pdl rc=$?"
done; done
python - <<'PY'
import json
for rep in (1,2):
  for pdl in (0,1):
    try:
        d=json.loads(open(f"gpurun_out/k/bench_pdl{pdl}_{rep}.json").read().strip().splitlines()[-1])
        print("pdl",pdl,"rep",rep,"ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "eager", round(d["eager_ms_per_step"],3), "status", d.get("status_word"))
    except Exception as e:
        print("pdl",pdl,"failed", e)
PY
