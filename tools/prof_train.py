"""A few fused train steps at batch 4096 on cuda:0 (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monoloco_b200 import synthetic
from monoloco_b200.network.architectures import LocoModel
from monoloco_b200.train import train_step
sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
m = LocoModel(34, 9, 1024, p_dropout=0.2, num_stage=3)
m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
m.cuda().train()
x = torch.from_numpy(synthetic.make_inputs(4096, 34, seed=3)).cuda()
y = torch.from_numpy(synthetic.make_labels(4096, seed=4)).cuda()
for _ in range(3):
    loss, _, _ = train_step(m, x, y, ('d', 'x', 'y', 'h', 'w', 'l', 'ori'))
torch.cuda.synchronize()
print(float(loss))
