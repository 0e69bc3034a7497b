#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,launch__cluster_max_active,launch__occupancy_cluster_pct,launch__waves_per_multiprocessor --clock-control none -k regex:"loco_forward_wide" -c 12 --csv --log-file gpurun_out/r2t_wide.csv python tools/prof_tc.py 16 > gpurun_out/r2t.log 2>&1
grep -c . gpurun_out/r2t_wide.csv; cut -d, -f5,12- gpurun_out/r2t_wide.csv | tail -30
