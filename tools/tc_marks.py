"""Timeline of CTA (0,0) inside the tensor-core forward kernel, first tile (mlb_debug_fwd_marks): per layer, us.
    python tools/tc_marks.py [B]     (env: MLB_TC_N, MLB_TC_MC)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from monoloco_b200 import synthetic, _lib as L_
from monoloco_b200.engine import LocoEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = LocoEngine(synthetic.make_state_dict('loco', 34, 9, 1024, 3, 7))
kps = torch.from_numpy(synthetic.make_keypoints(B, seed=1)).cuda()
kw = dict(kk=synthetic.KITTI_K, kind=L_.IN_KPS, kernel='tc')
lib = L_.lib()
buf = torch.zeros(256, dtype=torch.int64, device='cuda')
for _ in range(3):
    eng.forward(kps, **kw)
torch.cuda.synchronize()
L_.check(lib.mlb_debug_fwd_marks(C.c_void_p(buf.data_ptr())), 'marks')
eng.forward(kps, **kw)
torch.cuda.synchronize()
L_.check(lib.mlb_debug_fwd_marks(C.c_void_p(0)), 'marks')
m = buf.cpu().numpy().astype(np.int64)
print('N=%s MC=%s B=%d resident clusters %d' % (os.environ.get('MLB_TC_N', 'auto'), os.environ.get('MLB_TC_MC', 'auto'), B,
                                                 lib.mlb_tc_resident_clusters(eng._h)))
t0 = m[0]
for g in range(9):
    s, first, issued, acc, epi, bar, prod, acc1 = m[8 * g:8 * g + 8]
    if s == 0:
        break
    print('layer %d @%7.1f: first stage +%5.1f | MMAs issued +%5.1f | accumulators done +%5.1f | epilogue end +%5.1f | '
          'cluster barrier +%5.1f | producer done +%5.1f' % (g, (s - t0) / 1e3, (first - s) / 1e3, (issued - s) / 1e3, (acc - s) / 1e3,
                                                             (epi - s) / 1e3, (bar - s) / 1e3, (prod - s) / 1e3))
