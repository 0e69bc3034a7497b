import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monoloco_b200 import synthetic, engine, _lib as L_
sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
eng = engine.LocoEngine(sd)
x = torch.from_numpy(synthetic.make_keypoints(16, seed=1)).cuda()
for _ in range(6):
    eng.forward(x, kk=synthetic.KITTI_K, kind=L_.IN_KPS, kernel='cluster')
torch.cuda.synchronize()
