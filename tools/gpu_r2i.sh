#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MLB_TC_N=256 MLB_TC_MC=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:loco_forward_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2_tc_v1 python tools/prof_tc.py 4096 > gpurun_out/r2i_ncu.log 2>&1
ncu -i gpurun_out/r2_tc_v1.ncu-rep --page raw --csv > gpurun_out/r2_tc_v1_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_tc_v1.ncu-rep --page source --csv > gpurun_out/r2_tc_v1_source.csv 2>/dev/null
ncu -i gpurun_out/r2_tc_v1.ncu-rep --page details > gpurun_out/r2_tc_v1_details.txt 2>/dev/null
tail -5 gpurun_out/r2i_ncu.log; ls -la gpurun_out/r2_tc_v1*
