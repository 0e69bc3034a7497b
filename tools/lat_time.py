"""Latency of the small-batch kernels (CUDA events, L2 warm and flushed): python tools/lat_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monoloco_b200 import synthetic, engine, _lib as L_
eng = engine.LocoEngine(synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for B in (1, 8, 16):
    kps = torch.from_numpy(synthetic.make_keypoints(B, seed=1)).cuda()
    for kernel in ('wide2', 'wide'):
        res = {}
        for cold in (False, True):
            ts = []
            for _ in range(30):
                if cold:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); eng.forward(kps, kk=synthetic.KITTI_K, kind=L_.IN_KPS, kernel=kernel); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            res[cold] = np.median(ts[5:])
        print('B=%2d %-6s warm %.1f us   cold (L2 flushed) %.1f us' % (B, kernel, res[False], res[True]), flush=True)
