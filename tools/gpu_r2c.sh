#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_post_gpu.py -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest.log
timeout 300 python tools/tc_diag.py 1536 2048 2560 4096 8192 12288 16384 > gpurun_out/r2c_diag.log 2>&1
tail -5 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_diag.log
