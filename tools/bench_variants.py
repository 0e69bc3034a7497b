"""Time the fused forward for several (batch, rows_per_group) combinations on cuda:0 (device-timed, L2 warm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monoloco_b200 import synthetic, engine, _lib as L_

sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
eng = engine.LocoEngine(sd)
combos = [(1, 0), (16, 0), (64, 0), (256, 0), (288, 0), (512, 0), (1024, 0), (2048, 0), (4096, 0), (65536, 0)]
if len(sys.argv) > 1:
    combos = [tuple(int(v) for v in a.split(':')) for a in sys.argv[1:]]
def run(B, tm, kernel):
    x = torch.from_numpy(synthetic.make_keypoints(B, seed=1)).cuda()
    for _ in range(3):
        eng.forward(x, kk=synthetic.KITTI_K, kind=L_.IN_KPS, rows_per_group=tm, kernel=kernel)
    torch.cuda.synchronize()
    n = 20 if B <= 4096 else 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.forward(x, kk=synthetic.KITTI_K, kind=L_.IN_KPS, rows_per_group=tm, kernel=kernel)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for B, tm in combos:
    ms = run(B, tm, None)
    line = "B=%7d tm=%2d  auto %.4f ms  %.3f Mdet/s  %.1f TFLOP/s" % (B, tm, ms, B / ms / 1e3, B * 16865280 / ms / 1e9)
    if B <= 4096 and tm == 0:
        line += "   | tile %.4f ms | cluster %.4f ms" % (run(B, 0, 'tile'), run(B, 0, 'cluster'))
    print(line)
