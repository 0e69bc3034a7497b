#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2o_pytest.log
tail -4 gpurun_out/r2o_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?" >> gpurun_out/r2o_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2o_bench_ref.json 2>> gpurun_out/r2o_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:loco_forward_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2_tc_v3 python tools/prof_tc.py 4096 > gpurun_out/r2o_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:loco_train_kernel -s 2 -c 1 -f -o gpurun_out/r2_train_v2 python tools/prof_train.py >> gpurun_out/r2o_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:loco_forward_wide_kernel -s 2 -c 1 -f -o gpurun_out/r2_wide_v1 python tools/prof_tc.py 16 >> gpurun_out/r2o_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r2o_ncu.log; tail -2 gpurun_out/r2o_bench.err; head -c 600 gpurun_out/r2o_bench.json
