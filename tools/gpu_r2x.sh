#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/gather_overhead.py > gpurun_out/r2x.log 2>&1; tail -3 gpurun_out/r2x.log
