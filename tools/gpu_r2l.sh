#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py -x -q -k "tc_kernel or any_hidden or batches_vs_oracle or full_size" > gpurun_out/r2l_pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_pytest_tc.log
tail -5 gpurun_out/r2l_pytest_tc.log
timeout 200 python tools/tc_marks.py 4096 > gpurun_out/r2l_marks.log 2>&1
timeout 300 python tools/tc_time.py 256 1024 4096 8192 65536 >> gpurun_out/r2l_marks.log 2>&1
cat gpurun_out/r2l_marks.log
