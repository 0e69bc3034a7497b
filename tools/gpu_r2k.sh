#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2k_marks.log
for cfg in "256 0" "128 0" "128 1"; do set -- $cfg
  MLB_TC_N=$1 MLB_TC_MC=$2 timeout 200 python tools/tc_marks.py 4096 >> gpurun_out/r2k_marks.log 2>&1
done
MLB_TC_N=128 MLB_TC_MC=0 timeout 200 python tools/tc_time.py 1024 4096 8192 >> gpurun_out/r2k_marks.log 2>&1
cat gpurun_out/r2k_marks.log
