#!/bin/bash
# round-2 GPU call AD (1 GPU): ncu --set full of the final channels-last kernels at the C=256 56x56 tail site (residual + dout2)
mkdir -p gpurun_out/ad
timeout 300 python tools/site_probe.py 0 cl tail > gpurun_out/ad/site_tail.log 2>&1; echo "probe rc=$?"; cat gpurun_out/ad/site_tail.log | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cl_(stats|apply|bwd_reduce|bwd_apply|fwd_finalize|bwd_finalize)_kernel' -s 12 -c 6 -o gpurun_out/ad/prof_cl_tail python tools/site_probe.py 0 cl tail > gpurun_out/ad/ncu.log 2>&1; echo "ncu rc=$?"
