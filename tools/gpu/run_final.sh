#!/bin/bash
# round-2 last GPU call: the committed tree -- full parity suite + smoke (the bench line of this tree: calls S3 / AA)
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/final/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -2 gpurun_out/final/pytest_all.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/final/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final/smoke.log
