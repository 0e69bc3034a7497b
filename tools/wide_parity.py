"""Parity margin of the fused forward at hidden widths outside the reference default, against the live-reference
fixtures tests/golden/ref_wide_*.npz (oracle/gen_golden.py `wide`): prints err / tolerance of the raw outputs per fixture
and kernel, plus the same against the fp64 oracle on a larger synthetic batch.  GPU box only.

    python tools/wide_parity.py [out.json [substring of the fixture names to run]]
"""
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monoloco_b200 import synthetic, engine  # noqa: E402
from oracle import loco_oracle as O  # noqa: E402  (checker only)


def main():
    rows = []
    only = sys.argv[2] if len(sys.argv) > 2 else ''
    for path in sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'ref_wide_*.npz')), reverse=True):
        if only not in os.path.basename(path):
            continue
        f = np.load(path)
        isz, osz, L, st, seed = [int(v) for v in f['cfg'][:5]]
        kind = str(f['kind'])
        sd = synthetic.make_state_dict(kind, isz, osz, L, st, seed)
        eng = engine.LocoEngine(sd)
        kernels = [None] + (['tile'] if L <= 1024 else [])   # the cluster kernel is 1024-only
        for k in kernels:
            out = eng.forward(torch.from_numpy(f['x']).cuda(), kernel=k)
            ok, worst = O.close(out['raw'].cpu().numpy(), f['out'])
            rows.append(dict(fixture=os.path.basename(path), L=L, stages=st, kernel=eng.last_kernel(), rows=int(f['x'].shape[0]),
                             against='live reference', ok=bool(ok), err_over_tol=float(worst)))
            print(rows[-1], flush=True)
        x = synthetic.make_inputs(700, isz, seed=5)
        sd64 = {k: np.asarray(v, dtype=np.float64) for k, v in sd.items()}
        ref = O.model_forward(sd64, x.astype(np.float64))
        out = eng.forward(torch.from_numpy(x).cuda())
        ok, worst = O.close(out['raw'].cpu().numpy(), ref)
        rows.append(dict(fixture=os.path.basename(path), L=L, stages=st, kernel=eng.last_kernel(), rows=700,
                         against='fp64 oracle', ok=bool(ok), err_over_tol=float(worst)))
        print(rows[-1], flush=True)
        eng.close()
    if len(sys.argv) > 1:
        with open(sys.argv[1], 'w') as fh:
            json.dump(rows, fh, indent=1)


if __name__ == '__main__':
    main()
