#!/bin/bash
# round-2 GPU call J (1 GPU): specialised max-pool kernels, 2-column finalize blocks, stem convolution in NCHW (experiment)
mkdir -p gpurun_out/j
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/j/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/j/pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --sites-out gpurun_out/j/sites.json > gpurun_out/j/bench.json 2> gpurun_out/j/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --stem-nchw > gpurun_out/j/bench_stem_nchw.json 2> gpurun_out/j/bench_stem_nchw.err; echo "bench stem rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'finalize|maxpool' -c 400 --csv --log-file gpurun_out/j/fin_times.csv python tools/one_step.py > /dev/null 2>&1; echo "ncu fin rc=$?"
python - <<'PY'
import json, csv, collections
for n in ("bench","bench_stem_nchw"):
    try:
        d=json.loads(open(f"gpurun_out/j/{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "eager", round(d["eager_ms_per_step"],3))
        if n=="bench":
            for k,v in sorted(d["kernels"].items()): print("   %-18s %6.0f GB/s  %8.1f us/launch  share %.3f"%(k, v["gbs"] or 0, v["us_per_launch"], v["share_of_step"]))
    except Exception as e:
        print(n, "failed", e)
try:
    lines=[l for l in open("gpurun_out/j/fin_times.csv") if not l.startswith("==")]
    acc=collections.defaultdict(list)
    for r in csv.DictReader(lines): acc[r["Kernel Name"][:70]].append(float(r["Metric Value"].replace(",","")))
    for k,v in acc.items(): print("%-72s n=%3d avg %.1f us"%(k,len(v),sum(v)/len(v)/1e3))
except Exception as e: print("fin parse failed", e)
PY
