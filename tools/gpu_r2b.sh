#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_post_gpu.py -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_pytest.log
for B in 1024 2048 4096 16384 65536; do timeout 200 python tools/tc_forward.py $B >> gpurun_out/r2b_tc.log 2>&1; done
tail -5 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_tc.log
