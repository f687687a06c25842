#!/bin/bash
# round-2 GPU call X (1 GPU): tc_gram ring depth sweep (8 KB stages, two CTAs per SM): 5 / 7 / 9 / 11 / 13 stages on one box
mkdir -p gpurun_out/x
for n in 13 5 7 9 11 13; do
if [ $n = 13 ]; then unset DWT_B200_LIB; else export DWT_B200_LIB=$PWD/tools/gpu/prof/libdwt_b200_stages$n.so; fi
timeout 600 python bench.py --workload microbench --no-cpu-baseline > gpurun_out/x/micro_$n.json 2> gpurun_out/x/micro_$n.err; echo "micro stages=$n rc=$?"
python - <<PY
import json
try:
    mb=json.loads(open("gpurun_out/x/micro_$n.json").read().strip().splitlines()[-1])
    print("stages $n", mb["ms_per_step"], {k: round(v["us_per_launch"],1) for k,v in sorted(mb.get("kernels",{}).items()) if k in ("tc_stats","tc_bwd_reduce")})
except Exception as e: print("micro parse failed", e)
PY
done
