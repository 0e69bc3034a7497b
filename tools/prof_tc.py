"""One forward of the default kernel at batch B on cuda:0 (ncu target): python tools/prof_tc.py [B] [kernel]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monoloco_b200 import synthetic, engine, _lib as L_
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kernel = sys.argv[2] if len(sys.argv) > 2 else None
eng = engine.LocoEngine(synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0))
kps = torch.from_numpy(synthetic.make_keypoints(B, seed=0)).cuda()
for _ in range(3):
    out = eng.forward(kps, kk=synthetic.KITTI_K, kind=L_.IN_KPS, kernel=kernel)
torch.cuda.synchronize()
print(eng.last_kernel(), float(out['raw'].abs().sum()))
