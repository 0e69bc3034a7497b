import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from monoloco_b200 import synthetic, engine, _lib as L_
from oracle import loco_oracle as O
sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
eng = engine.LocoEngine(sd)
xn = synthetic.make_inputs(900, 34, seed=9)
x = torch.from_numpy(xn).cuda()
ref = O.loco_model_forward(sd, xn)
for tm in (0, 8, 14, 16):
    a = eng.forward(x, kernel='tile', rows_per_group=tm)['raw']
    b = eng.forward(x, res_tmem=True, kernel='tile', rows_per_group=tm)['raw']
    print('tm', tm, 'equal', torch.equal(a, b), 'maxdiff', float((a - b).abs().max()), 'scratch ok', O.close(a.cpu().numpy(), ref), 'tmem ok', O.close(b.cpu().numpy(), ref))
    bad = (a != b).any(1).nonzero().flatten().cpu().numpy()
    print('   bad rows', bad[:20], len(bad))
