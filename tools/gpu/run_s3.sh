#!/bin/bash
# round-2 GPU call S3 (2 GPUs): final build (ABI v5, fork_for_sum, 2x2-patch max-pool backward): parity suite, smoke, default
# bench line with extras, reference arm, 2-GPU data-parallel step
mkdir -p gpurun_out/s3
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/s3/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/s3/pytest_all.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/s3/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 --sites-out gpurun_out/s3/sites.json > gpurun_out/s3/bench_full.json 2> gpurun_out/s3/bench_full.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s3/bench_reference_arm.json 2> gpurun_out/s3/bench_reference_arm.err; echo "ref arm rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s3/bench_2gpu.json 2> gpurun_out/s3/bench_2gpu.err; echo "bench 2gpu rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/s3/launches_ours.csv python tools/one_step.py > gpurun_out/s3/ncu_ours.log 2>&1; echo "ncu ours rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/s3/bench_full.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "eager", d["eager_ms_per_step"], "status", d.get("status_word"))
    print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("norm_path"))
    print("ref_on_gpu", {k:d["reference_on_gpu"].get(k) for k in ("value","ms_per_step","speedup_value","speedup_e2e","error")})
    mb=d.get("microbench",{})
    print("micro", {k:mb.get(k) for k in ("ms_per_step","vs_reference_gpu","error")}, mb.get("roofline",{}).get("frac"))
    for k,v in sorted(d["kernels"].items()): print("   %-18s %6.0f GB/s  %8.1f us/launch  share %.3f"%(k, v["gbs"] or 0, v["us_per_launch"], v["share_of_step"]))
    r=json.loads(open("gpurun_out/s3/bench_2gpu.json").read().strip().splitlines()[-1])
    print("2gpu ms/step", r["ms_per_step"], "img/s", r["value"], "e2e", r["e2e"]["value"])
except Exception as e:
    print("parse failed", e)
PY
