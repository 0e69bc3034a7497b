"""A/B timings of the throughput kernel (residual stash: Tensor Memory vs L2 scratch) at B=4096 / 65536."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monoloco_b200 import synthetic, engine, _lib as L_
sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
eng = engine.LocoEngine(sd)
for B in (4096, 65536):
    x = torch.from_numpy(synthetic.make_keypoints(B, seed=1)).cuda()
    for tmem in (True, False):
        kw = dict(kk=synthetic.KITTI_K, kind=L_.IN_KPS, res_tmem=tmem, kernel='tile')
        for _ in range(3):
            eng.forward(x, **kw)
        torch.cuda.synchronize()
        n = 30 if B <= 4096 else 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            eng.forward(x, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print("B=%6d res_tmem=%d  %.4f ms  %.1f TFLOP/s" % (B, tmem, ms, B * 16865280 / ms / 1e9))
