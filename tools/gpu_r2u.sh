#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
FLUSH=1 timeout 120 python tools/fwd_marks.py 16 wide2 > gpurun_out/r2u_marks.log 2>&1; cat gpurun_out/r2u_marks.log
