#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2n_marks.log
MLB_TC_CLUSTERS=16 timeout 200 python tools/tc_marks.py 2048 >> gpurun_out/r2n_marks.log 2>&1
MLB_TC_CLUSTERS=8 timeout 200 python tools/tc_marks.py 1024 >> gpurun_out/r2n_marks.log 2>&1
timeout 200 python tools/tc_marks.py 4096 >> gpurun_out/r2n_marks.log 2>&1
cat gpurun_out/r2n_marks.log
