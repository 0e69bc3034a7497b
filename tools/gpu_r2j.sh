#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2j_time.log
for cfg in "128 1" "128 0" "256 1" "256 0"; do set -- $cfg
  MLB_TC_N=$1 MLB_TC_MC=$2 timeout 300 python tools/tc_time.py 256 1024 4096 8192 65536 >> gpurun_out/r2j_time.log 2>&1
done
cat gpurun_out/r2j_time.log
timeout 900 python -m pytest tests/test_forward_gpu.py -x -q -k "tc_kernel or any_hidden or batches_vs_oracle or full_size" > gpurun_out/r2j_pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_pytest_tc.log
tail -5 gpurun_out/r2j_pytest_tc.log
