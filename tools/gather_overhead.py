"""Cost of the fused all-gather protocol without any peer: world size 1 (self gather) vs plain forward, batch 4096."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
os.environ.setdefault('MASTER_ADDR', '127.0.0.1'); os.environ.setdefault('MASTER_PORT', '29533')
dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda', 0))
from monoloco_b200 import synthetic, engine, distributed as D, _lib as L_
eng = engine.LocoEngine(synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0))
B = 4096
kps = torch.from_numpy(synthetic.make_keypoints(B, seed=0)).cuda()
sh = D.ShardedLoco(eng, B, mode='fused')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return np.median(ts)
print('plain forward        %.4f ms' % timed(lambda: eng.forward(kps, kk=synthetic.KITTI_K, kind=L_.IN_KPS)))
print('fused gather, world 1 %.4f ms' % timed(lambda: sh.forward(kps, synthetic.KITTI_K)))
sh.close(); dist.destroy_process_group()
