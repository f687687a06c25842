#!/bin/bash
# round-2 GPU call V (1 GPU): tile timeline of tc_gram
mkdir -p gpurun_out/v
timeout 300 python tools/gpu/prof/gram_timeline.py > gpurun_out/v/gram_timeline.log 2>&1; echo "timeline rc=$?"; tail -5 gpurun_out/v/gram_timeline.log
