"""Timeline of CTA 0 inside the tile forward kernel (mlb_debug_fwd_marks): per op GEMM / epilogue / sync / act-write us."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from monoloco_b200 import synthetic, _lib as L_
from monoloco_b200.engine import LocoEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
KERNEL = sys.argv[2] if len(sys.argv) > 2 else 'tile'
eng = LocoEngine(synthetic.make_state_dict('loco', 34, 9, 1024, 3, 7))
kps = torch.from_numpy(synthetic.make_keypoints(B, seed=1)).cuda()
kw = dict(kk=synthetic.KITTI_K, kind=L_.IN_KPS, kernel=KERNEL)
lib = L_.lib()
buf = torch.zeros(256, dtype=torch.int64, device='cuda')
for _ in range(3):
    eng.forward(kps, **kw)
torch.cuda.synchronize()
L_.check(lib.mlb_debug_fwd_marks(C.c_void_p(buf.data_ptr())), 'marks')
if os.environ.get('FLUSH'):   # cold: rewrite a buffer larger than L2 right before the marked launch
    torch.empty(256 << 20, dtype=torch.uint8, device='cuda').zero_()
eng.forward(kps, **kw)
torch.cuda.synchronize()
L_.check(lib.mlb_debug_fwd_marks(C.c_void_p(0)), 'marks')
m = buf.cpu().numpy().astype(np.int64)
t0 = m[0]
n_ops = eng.packed.desc['n_ops']
print('input staged %.1f us' % ((m[1] - t0) / 1e3))
if KERNEL == 'wide2':
    print('wide2 CTA 0: input staged +%.1f us' % ((m[1] - t0) / 1e3))
    prev = m[1]
    for i in range(9):
        a, b, c, d = m[2 + 4 * i:6 + 4 * i]
        if a == 0:
            break
        f1, f2 = m[64 + 2 * i], m[65 + 2 * i]
        print('gemm %d: layer start +%.1f | input pairs landed %.1f | math %.1f | DSMEM stores %.1f | cluster barrier %.1f | finalise + publish %.1f  (us)'
              % (i, (a - prev) / 1e3, (b - a) / 1e3, (f1 - b) / 1e3, (f2 - f1) / 1e3, (c - f2) / 1e3, (d - c) / 1e3))
        prev = d
    h0, h1, h2 = m[2 + 36], m[3 + 36], m[4 + 36]
    print('head partials published +%.1f us, collected +%.1f us, stored +%.1f us, total %.1f us' % ((h0 - prev) / 1e3, (h1 - h0) / 1e3, (h2 - h1) / 1e3, (h2 - t0) / 1e3))
    sys.exit(0)
if KERNEL == 'wide':
  for base in (0, 128):
    m = buf.cpu().numpy().astype(np.int64)[base:]
    if m[0] == 0:
        continue
    print('CTA %d (start %+.1f us vs CTA 0)' % (64 if base else 0, (m[0] - t0) / 1e3))
    prev = m[1]
    for i in range(9):
        a, b, c, d = m[2 + 4 * i:6 + 4 * i]
        if c == 0:
            c = d = b
        print('gemm %d: wait weights %.1f  partial sums %.1f  epilogue+barrier %.1f  exchange %.1f  (us)' % (i, (a - prev) / 1e3, (b - a) / 1e3, (c - b) / 1e3, (d - c) / 1e3))
        prev = d
    if m[39]:
        print('heads +%.1f us, stored +%.1f us, total %.1f us' % ((m[38] - prev) / 1e3, (m[39] - m[38]) / 1e3, (m[39] - t0) / 1e3))
  sys.exit(0)
prev = m[1]
for i in range(n_ops):
    g, e, s, w = m[2 + 4 * i:6 + 4 * i]
    if g == 0:
        continue
    print('op %2d: gemm %.1f  epilogue %.1f  sync %.1f  act write %.1f   (us)' % (i, (g - prev) / 1e3, (e - g) / 1e3, (s - e) / 1e3, (w - s) / 1e3))
    prev = w
print('heads done +%.1f us, rows stored +%.1f us, total %.1f us' % ((m[2 + 4 * n_ops] - prev) / 1e3, (m[3 + 4 * n_ops] - m[2 + 4 * n_ops]) / 1e3, (m[3 + 4 * n_ops] - t0) / 1e3))
