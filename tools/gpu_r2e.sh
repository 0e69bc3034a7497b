#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py -x -q -k "tc_kernel or any_hidden or batches_vs_oracle or full_size" > gpurun_out/r2e_pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_pytest_tc.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?" >> gpurun_out/r2e_bench.err
tail -15 gpurun_out/r2e_pytest_tc.log; tail -15 gpurun_out/r2e_pytest.log; tail -3 gpurun_out/r2e_bench.err; head -c 1500 gpurun_out/r2e_bench.json
