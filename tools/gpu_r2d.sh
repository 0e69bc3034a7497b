#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in 0 1 2; do echo "== MLB_TC_MODE=$m" >> gpurun_out/r2d_diag.log; MLB_TC_MODE=$m timeout 300 python tools/tc_diag.py 4096 8192 16384 >> gpurun_out/r2d_diag.log 2>&1; done
echo "== mode 2, 33 / 36 clusters" >> gpurun_out/r2d_diag.log
MLB_TC_MODE=2 MLB_TC_CLUSTERS=33 timeout 300 python tools/tc_diag.py 8192 >> gpurun_out/r2d_diag.log 2>&1
MLB_TC_MODE=2 MLB_TC_CLUSTERS=36 timeout 300 python tools/tc_diag.py 8192 >> gpurun_out/r2d_diag.log 2>&1
cat gpurun_out/r2d_diag.log
