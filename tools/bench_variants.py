"""Time the fused forward for several (batch, rows_per_group) combinations on cuda:0 (device-timed, L2 warm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monoloco_b200 import synthetic, engine, _lib as L_

sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
eng = engine.LocoEngine(sd)
combos = [(4096, 0), (4096, 14), (4096, 16), (256, 0), (1, 0), (32, 0), (65536, 0), (131072, 0)]
if len(sys.argv) > 1:
    combos = [tuple(int(v) for v in a.split(':')) for a in sys.argv[1:]]
for B, tm in combos:
    x = torch.from_numpy(synthetic.make_keypoints(B, seed=1)).cuda()
    for _ in range(3):
        eng.forward(x, kk=synthetic.KITTI_K, kind=L_.IN_KPS, rows_per_group=tm)
    torch.cuda.synchronize()
    n = 20 if B <= 4096 else 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.forward(x, kk=synthetic.KITTI_K, kind=L_.IN_KPS, rows_per_group=tm)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("B=%7d tm=%2d  %.4f ms  %.3f Mdet/s  %.1f TFLOP/s" % (B, tm, ms, B / ms / 1e3, B * 16865280 / ms / 1e9))
