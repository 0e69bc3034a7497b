#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:loco_forward_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2_tc_v2 python tools/prof_tc.py 4096 > gpurun_out/r2m_ncu.log 2>&1
ncu -i gpurun_out/r2_tc_v2.ncu-rep --page raw --csv > gpurun_out/r2_tc_v2_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_tc_v2.ncu-rep --page source --csv > gpurun_out/r2_tc_v2_source.csv 2>/dev/null
tail -3 gpurun_out/r2m_ncu.log
