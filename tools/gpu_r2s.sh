#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/lat_time.py > gpurun_out/r2s_lat.log 2>&1; cat gpurun_out/r2s_lat.log
