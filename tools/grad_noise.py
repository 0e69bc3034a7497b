"""Self-noise of the fp32 training oracle: the SAME train step (oracle/torch_port.py, torch autograd on the CPU) evaluated on
the batch in its given row order and with the rows reversed -- a pure change of fp32 summation order in the batch
reductions (BatchNorm statistics, dW = G^T A, bias gradients).  The relative L2 distance between the two gradient sets is
the floor any other fp32 implementation (ours included) can be held to.  Writes profiles/r2_grad_noise.{json,md}.

    python tools/grad_noise.py            (CPU only, a few minutes)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monoloco_b200 import synthetic
from oracle import torch_port as T

torch.set_num_threads(16)
TASKS = ('d', 'x', 'y', 'h', 'w', 'l', 'ori')
rows_out = []
for (L, st, B, p) in ((256, 3, 301, 0.2), (512, 3, 1000, 0.5), (1024, 3, 4096, 0.2), (1024, 3, 4096, 0.0)):
    sd = synthetic.make_state_dict('loco', 34, 9, L, st, 11)
    x = torch.from_numpy(synthetic.make_inputs(B, 34, seed=3))
    y = torch.from_numpy(synthetic.make_labels(B, seed=4))
    n_sites = 2 + 2 * st
    masks = [torch.from_numpy((np.random.RandomState(5 + i).uniform(size=(B, L)) >= p).astype(np.float32)) for i in range(n_sites)]

    def grads(perm):
        tsd = T.to_torch(sd, requires_grad=True)
        out = T.model_forward(tsd, x[perm], training=True, p_dropout=p, masks=[m[perm] for m in masks] if p > 0 else None)
        loss, _ = T.multi_task_loss(out, y[perm], TASKS)
        loss.backward()
        return {k: v.grad.numpy().astype(np.float64) for k, v in tsd.items() if v.requires_grad and v.grad is not None}, float(loss)

    ident = torch.arange(B)
    g0, l0 = grads(ident)
    g1, l1 = grads(torch.flip(ident, dims=[0]))
    worst, worst_name, big = 0.0, '', []
    for k in g0:
        nrm = np.linalg.norm(g0[k])
        if nrm > 1e-5 * np.sqrt(g0[k].size):
            rel = float(np.linalg.norm(g0[k] - g1[k]) / nrm)
            big.append(rel)
            if rel > worst:
                worst, worst_name = rel, k
    rec = {"L": L, "stages": st, "batch": B, "p_dropout": p, "loss_rel_diff": abs(l0 - l1) / abs(l0),
           "grad_rel_l2_max": worst, "grad_rel_l2_max_tensor": worst_name, "grad_rel_l2_median": float(np.median(big)),
           "tensors": len(big)}
    rows_out.append(rec)
    print(rec, flush=True)
out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles')
json.dump(rows_out, open(os.path.join(out_dir, 'r2_grad_noise.json'), 'w'), indent=1)
with open(os.path.join(out_dir, 'r2_grad_noise.md'), 'w') as f:
    f.write("# fp32 self-noise of the training oracle (tools/grad_noise.py)\n\nThe torch-autograd oracle against ITSELF with the batch rows "
            "reversed (same weights, inputs, dropout masks): relative L2 distance of every parameter gradient.  This is the floor "
            "behind the gradient tolerances of tests/test_train_gpu.py (tight rule rel-L2 <= 2e-5 up to B*L < 2^20, statistical rule "
            "rel-L2 <= 3e-3 at batch 4096 x 1024).\n\n| L | stages | batch | p_dropout | loss rel. diff | grad rel-L2 max (tensor) | grad rel-L2 median |\n|---|---|---|---|---|---|---|\n")
    for r in rows_out:
        f.write("| %d | %d | %d | %.1f | %.2e | %.2e (`%s`) | %.2e |\n" % (r['L'], r['stages'], r['batch'], r['p_dropout'], r['loss_rel_diff'],
                                                                        r['grad_rel_l2_max'], r['grad_rel_l2_max_tensor'], r['grad_rel_l2_median']))
print('written')
