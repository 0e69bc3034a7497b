#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
MLB_GRAD_RELL2=1e-5 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_pytest.log
tail -6 gpurun_out/r2q_pytest.log
python -c "
import __graft_entry__ as g
g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; tail -3 gpurun_out/r2q_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?" >> gpurun_out/r2q_bench.err
tail -2 gpurun_out/r2q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_selection'])
print({k:(v['ms'],v['kernel'][:24]) for k,v in d['extras']['forward_ms_by_batch'].items()})
PY
