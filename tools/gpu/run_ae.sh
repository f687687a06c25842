#!/bin/bash
# round-2 GPU call AE (1 GPU): ncu --set full of the final tensor-core path (config 2: N=256 C=256 56x56 gs=64), one fwd + bwd
mkdir -p gpurun_out/ae
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_gram|tc_apply|tc_contract|fwd_factor|bwd_coef|partial_reduce' -s 8 -c 8 -o gpurun_out/ae/prof_tc_path python tools/micro_once.py 256 > gpurun_out/ae/ncu.log 2>&1; echo "ncu rc=$?"
