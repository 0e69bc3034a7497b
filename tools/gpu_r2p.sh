#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py -x -q -k "tc_kernel or any_hidden or batches_vs_oracle or full_size" > gpurun_out/r2p_pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_pytest_tc.log
tail -3 gpurun_out/r2p_pytest_tc.log
timeout 300 python tools/tc_time.py 256 4096 8192 65536 > gpurun_out/r2p_time.log 2>&1; cat gpurun_out/r2p_time.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:loco_forward_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2_tc_v4 python tools/prof_tc.py 4096 > gpurun_out/r2p_ncu.log 2>&1
ncu -i gpurun_out/r2_tc_v4.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
for a,b,c in zip(h,u,v):
    if a in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct'): print(a,c,b)
"
