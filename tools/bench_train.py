"""Time the fused train step (single cooperative launch) vs torch-eager CUDA autograd of the same step on cuda:0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monoloco_b200 import synthetic
from monoloco_b200.train import train_step, CompositeLoss, MultiTaskLoss
from monoloco_b200.network.architectures import LocoModel
from oracle import torch_port as T

tasks = ('d', 'x', 'y', 'h', 'w', 'l', 'ori')


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for B in [int(v) for v in (sys.argv[1:] or ['4096', '512'])]:
    sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 7)
    PD = float(os.environ.get('MLB_BENCH_PDROP', '0.2'))
    m = LocoModel(34, 9, 1024, p_dropout=PD, num_stage=3)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    m.cuda().train()
    x = torch.from_numpy(synthetic.make_inputs(B, 34, seed=3)).cuda()
    y = torch.from_numpy(synthetic.make_labels(B, seed=4)).cuda()
    mt = MultiTaskLoss(*CompositeLoss(tasks)(), (1,) * len(tasks), tasks)
    t_fused = timeit(lambda: train_step(m, x, y, tasks))
    from monoloco_b200.train.fused import phase_times
    print('   phases (ms):', ' '.join('%s%d:%.3f' % (n, b, ms) for n, b, ms in phase_times(m)))
    if os.environ.get('MLB_SUBPHASES'):
        from monoloco_b200.train.fused import subphase_times
        for n, b, pts in subphase_times(m):
            if n in ('FWD', 'BWD', 'FWD_FINAL'):
                # CTA first | last: stats loaded, rows finished, act written, tile ready, GEMM done, epilogue done, barrier left
                print('   %-9s %d  ' % (n, b) + ' | '.join(' '.join('%.3f' % pts[c][k] for k in (4, 5, 6, 0, 1, 2, 3)) for c in (0, 2)))

    def dropin():
        m.zero_grad(set_to_none=True)
        loss, _ = mt(m(x), y, phase='train')
        loss.backward()
    t_dropin = timeit(dropin)
    # torch-eager CUDA (cuBLAS SGEMM, TF32 off) -- the only GPU implementation the reference has
    tsd = {k: (torch.as_tensor(v).cuda().requires_grad_(True) if ('running' not in k and torch.as_tensor(v).is_floating_point())
               else torch.as_tensor(v).cuda()) for k, v in sd.items()}

    def eager():
        for v in tsd.values():
            v.grad = None
        out = T.model_forward(tsd, x, training=True, p_dropout=PD)
        loss, _ = T.multi_task_loss(out, y, tasks)
        loss.backward()
    t_eager = timeit(eager)
    fl = 3 * 16865280 * B
    print("train step B=%5d: fused 1-launch %.3f ms (%.1f TFLOP/s) | drop-in autograd (2 launches) %.3f ms | torch-eager CUDA %.3f ms | x%.2f"
          % (B, t_fused, fl / t_fused / 1e9, t_dropin, t_eager, t_eager / t_fused))
