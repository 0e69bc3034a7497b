#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_pytest.log
tail -4 gpurun_out/r2v_pytest.log
FLUSH=1 timeout 120 python tools/fwd_marks.py 16 wide2 > gpurun_out/r2v_marks.log 2>&1; grep "head\|staged" gpurun_out/r2v_marks.log
timeout 120 python tools/fwd_marks.py 16 wide2 >> gpurun_out/r2v_marks.log 2>&1; grep "head\|staged" gpurun_out/r2v_marks.log | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?" >> gpurun_out/r2v_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print({k:(round(v['ms'],4),v['kernel'][:26]) for k,v in d['extras']['forward_ms_by_batch'].items()})
print(d['extras']['small_batch_roofline'])
PY
timeout 300 ncu --set full --clock-control none -k regex:loco_forward_wide2_kernel -s 2 -c 1 -f -o gpurun_out/r2_wide2_v1 python tools/prof_tc.py 16 > gpurun_out/r2v_ncu.log 2>&1; tail -2 gpurun_out/r2v_ncu.log
