"""Kernel time of the three inference kernels at one image's worth of detections (CUDA events, 200 launches each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monoloco_b200 import synthetic, engine, _lib as L_

eng = engine.LocoEngine(synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0))
for B in [int(v) for v in (sys.argv[1:] or ['1', '8', '16', '32'])]:
    x = torch.from_numpy(synthetic.make_keypoints(B, seed=1)).cuda()
    res = []
    for kern in ('wide', 'cluster', 'tile'):
        if kern == 'wide' and B > 64:
            res.append('wide n/a')
            continue
        kw = dict(kk=synthetic.KITTI_K, kind=L_.IN_KPS, kernel=kern)
        for _ in range(20):
            eng.forward(x, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            eng.forward(x, **kw)
        e1.record()
        torch.cuda.synchronize()
        res.append('%s %.1f us' % (kern, e0.elapsed_time(e1) * 5.0))
    print('B=%3d: ' % B + ' | '.join(res))
