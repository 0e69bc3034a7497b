#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_forward_gpu.py -x -q -k "wide2" > gpurun_out/r2r_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_pytest.log
tail -15 gpurun_out/r2r_pytest.log
timeout 120 python tools/fwd_marks.py 16 wide2 > gpurun_out/r2r_marks.log 2>&1; cat gpurun_out/r2r_marks.log
timeout 120 python tools/fwd_marks.py 16 wide >> gpurun_out/r2r_marks.log 2>&1; tail -3 gpurun_out/r2r_marks.log
