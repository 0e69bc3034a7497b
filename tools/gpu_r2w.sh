#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2w_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2w_pytest.log
tail -4 gpurun_out/r2w_pytest.log
