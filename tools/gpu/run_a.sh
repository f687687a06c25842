#!/bin/bash
# round-2 GPU call A: parity suite on the reworked channels-last kernels + step time with / without the domain-sequential sweep
mkdir -p gpurun_out/a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a/pytest.log
tail -5 gpurun_out/a/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sites-out gpurun_out/a/sites_seq.json > gpurun_out/a/bench_seq.json 2> gpurun_out/a/bench_seq.err; echo "bench rc=$?"
DWT_CL_SEQ_MB=100000000 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sites-out gpurun_out/a/sites_par.json > gpurun_out/a/bench_par.json 2> gpurun_out/a/bench_par.err; echo "bench par rc=$?"
DWT_CL_SEQ_MB=40 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sites-out gpurun_out/a/sites_seq40.json > gpurun_out/a/bench_seq40.json 2> gpurun_out/a/bench_seq40.err; echo "bench seq40 rc=$?"
python - <<'PY'
import json
for n in ("seq","par","seq40"):
    try:
        d=json.loads(open(f"gpurun_out/a/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "eager", round(d["eager_ms_per_step"],3), "launches", d["gpu_launches"])
        for k,v in sorted(d["kernels"].items()):
            print("   %-18s %6.0f GB/s  %8.1f us/launch  share %.3f"%(k, v["gbs"] or 0, v["us_per_launch"], v["share_of_step"]))
    except Exception as e:
        print(n, "failed", e)
PY
