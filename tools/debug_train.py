import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import torch_port as T
from monoloco_b200 import synthetic
from monoloco_b200.train import train_step
from monoloco_b200.network.architectures import LocoModel

def run(L, st, B, p, tm):
    os.environ['MLB_TRAIN_ROWS_PER_GROUP'] = str(tm)
    sd = synthetic.make_state_dict('loco', 34, 9, L, st, 7)
    m = LocoModel(34, 9, L, p_dropout=p, num_stage=st)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()}); m.cuda().train()
    x = synthetic.make_inputs(B, 34, seed=3); y = synthetic.make_labels(B, seed=4)
    n_bn = 2 * st + 2
    masks = (np.random.RandomState(5).uniform(size=(n_bn, B, L)) >= p).astype(np.uint8)
    tasks = ('d', 'x', 'y', 'h', 'w', 'l', 'ori')
    loss, vals, out = train_step(m, torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), tasks,
                                 drop_mask=torch.from_numpy(masks).cuda() if p > 0 else None)
    tsd = T.to_torch(sd, requires_grad=True)
    ro = T.model_forward(tsd, torch.from_numpy(x), training=True, p_dropout=p, masks=[torch.from_numpy(mm) for mm in masks] if p > 0 else None)
    rl, _ = T.multi_task_loss(ro, torch.from_numpy(y), tasks); rl.backward()
    print("L=%d st=%d B=%d p=%.1f tm=%d loss %.6f ref %.6f out err %.2e" % (L, st, B, p, tm, float(loss), float(rl), float((out.cpu()-ro.detach()).abs().max())))
    for n, prm in m.named_parameters():
        g, r = prm.grad.cpu().numpy().astype(np.float64), tsd[n].grad.numpy().astype(np.float64)
        rel = np.linalg.norm(g - r) / max(np.linalg.norm(r), 1e-30)
        flag = '  <<<<' if rel > 2e-5 and np.abs(r).max() > 1e-6 else ''
        print("   %-36s relL2 %.2e  max|ref| %.2e maxerr %.2e%s" % (n, rel, np.abs(r).max(), np.abs(g - r).max(), flag))

for cfg in [(1024, 3, 4096, 0.2, 0), (1024, 3, 4096, 0.0, 0), (256, 2, 300, 0.0, 14), (256, 2, 300, 0.0, 16), (256, 2, 300, 0.0, 10)]:
    run(*cfg)
