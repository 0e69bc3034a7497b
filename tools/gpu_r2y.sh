#!/bin/bash
# 8 GPUs, short: 2- and 4-rank gather tests + one 8-rank and one 4-rank bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -q > gpurun_out/r2y_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_pytest.log
tail -3 gpurun_out/r2y_pytest.log
for N in 8 4; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2y_bench$N.json 2> gpurun_out/r2y_bench$N.err
python - gpurun_out/r2y_bench$N.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['n_gpus'], 'ms/step %.4f'%d['ms_per_step'], d['ms_per_step_stats'], 'value %.4e'%d['value'], 'e2e %.4e'%d['e2e']['value'], d.get('gather_check'), d['config4_131072_per_gpu']['ms_per_step'], d['config4_131072_per_gpu']['value'])
PY
done
