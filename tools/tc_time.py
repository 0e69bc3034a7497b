"""Time the tensor-core kernel (forced) at several batch sizes: python tools/tc_time.py [B ...]   (env: MLB_TC_N, MLB_TC_MC)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monoloco_b200 import synthetic, engine, _lib as L_
from oracle import loco_oracle as O

sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
eng = engine.LocoEngine(sd)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for B in [int(v) for v in (sys.argv[1:] or ['128', '256', '1024', '2048', '4096', '8192', '65536'])]:
    kps = torch.from_numpy(synthetic.make_keypoints(B, seed=1)).cuda()
    out = eng.forward(kps, kk=synthetic.KITTI_K, kind=L_.IN_KPS, kernel='tc')
    torch.cuda.synchronize()
    idx = np.random.RandomState(0).choice(B, min(B, 128), replace=False)
    x = O.preprocess_monoloco(kps.cpu().numpy()[idx], synthetic.KITTI_K)
    ok, worst = O.close(out['raw'].cpu().numpy()[idx], O.loco_model_forward(sd, x))
    ts = []
    for _ in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.forward(kps, kk=synthetic.KITTI_K, kind=L_.IN_KPS, kernel='tc'); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    print('N=%s MC=%s B=%6d  median %.4f ms  min %.4f  -> %.2f M det/s   parity ok=%s worst/tol=%.3f'
          % (os.environ.get('MLB_TC_N', 'auto'), os.environ.get('MLB_TC_MC', 'auto'), B, ts[len(ts) // 2], ts[0],
             B / ts[len(ts) // 2] / 1e3, ok, worst), flush=True)
