#!/bin/bash
# 8 GPUs: 2- and 4-rank gather tests, then the scaling bench at 8 / 4 (fused) and 8 (nccl)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -q > gpurun_out/r2h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_pytest.log
tail -8 gpurun_out/r2h_pytest.log
for N in 8 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2h_bench$N.json 2> gpurun_out/r2h_bench$N.err; echo "rc=$?" >> gpurun_out/r2h_bench$N.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 --gather nccl > gpurun_out/r2h_bench8_nccl.json 2> gpurun_out/r2h_bench8_nccl.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2h_bench8_b.json 2> gpurun_out/r2h_bench8_b.err
for f in gpurun_out/r2h_bench8.json gpurun_out/r2h_bench4.json gpurun_out/r2h_bench2.json gpurun_out/r2h_bench8_nccl.json gpurun_out/r2h_bench8_b.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d['n_gpus'], 'ms/step %.4f'%d['ms_per_step'], d['ms_per_step_stats'], 'value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d.get('gather_check'), {k:(v if k!='gather_check' else v) for k,v in d.get('config4_131072_per_gpu',{}).items()})
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
