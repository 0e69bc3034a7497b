#!/bin/bash
# 2 GPUs: the sharded gather tests + a 2-rank bench line (fused and nccl)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest.log
tail -30 gpurun_out/r2g_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g_bench2.json 2> gpurun_out/r2g_bench2.err; echo "rc=$?" >> gpurun_out/r2g_bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --gather nccl > gpurun_out/r2g_bench2_nccl.json 2> gpurun_out/r2g_bench2_nccl.err; echo "rc=$?" >> gpurun_out/r2g_bench2_nccl.err
tail -3 gpurun_out/r2g_bench2.err; cat gpurun_out/r2g_bench2.json | tail -c 2500; echo; cat gpurun_out/r2g_bench2_nccl.json | head -c 600
