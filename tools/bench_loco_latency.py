"""Wall-clock latency of the reference-facing call `Loco.forward(keypoints, kk)` per image (m detections), as
predict.py uses it: python lists in, dict of CPU tensors out."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monoloco_b200 import synthetic
from monoloco_b200.network import Loco
from monoloco_b200.network.architectures import LocoModel

sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 1)
m = LocoModel(34, 9, 1024)
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
net = Loco(model=m, mode='mono', device=torch.device('cuda'))
for n in (1, 16, 64):
    kps = synthetic.make_keypoints(n, seed=1).tolist()
    for _ in range(5):
        net.forward(kps, synthetic.KITTI_K)
    t0 = time.perf_counter()
    reps = 200
    for _ in range(reps):
        net.forward(kps, synthetic.KITTI_K)
    dt = (time.perf_counter() - t0) / reps
    print("Loco.forward m=%3d: %.3f ms per image" % (n, dt * 1e3))
