#!/bin/bash
# round-2 GPU call A (1 GPU): full gpu suite, tcgen05 probes (each in its own process), bench line
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
export MLB_EXPERIMENTAL=1
timeout 120 python -m pytest tests/test_probe_tc_gpu.py -q -x -k "tf32x3_probe" > gpurun_out/r2a_probe1.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_probe1.log
timeout 120 python tools/probe_tc.py 1024 >> gpurun_out/r2a_probe1.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_probe1.log
timeout 120 python -m pytest tests/test_probe_tc_gpu.py -q -x -k "layer_probe" > gpurun_out/r2a_probe2.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_probe2.log
timeout 120 python tools/probe_tc.py layer 4096 1024 1024 >> gpurun_out/r2a_probe2.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_probe2.log
timeout 180 python -m pytest tests/test_probe_tc_gpu.py -q -k "tc_forward" > gpurun_out/r2a_probe3.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_probe3.log
unset MLB_EXPERIMENTAL
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_pytest.log; tail -3 gpurun_out/r2a_probe1.log gpurun_out/r2a_probe2.log gpurun_out/r2a_probe3.log; tail -c 600 gpurun_out/r2a_bench.json
