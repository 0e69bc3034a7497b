"""GPU x2 (skipped on a single-GPU box): detections sharded over two ranks, outputs all-gathered either by NCCL or
by the kernel's fused peer stores over NVLink; both must equal the oracle on the full batch."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from monoloco_b200 import synthetic, engine, distributed as D
    sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
    eng = engine.LocoEngine(sd, device=torch.device('cuda', rank))
    kps = synthetic.make_keypoints(n_total, seed=5)
    res = {}
    for mode in ('nccl', 'fused'):
        sh = D.ShardedLoco(eng, n_total, mode=mode)
        local = torch.from_numpy(kps[sh.start:sh.stop]).cuda()
        rows = sh.forward(local, synthetic.KITTI_K)
        torch.cuda.synchronize()
        res[mode] = rows.cpu().numpy().copy()
        dist.barrier()
        sh.close()
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_total', [4096 * 2, 1001, 40])
def test_two_gpu_gather(n_total):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import loco_oracle as O
    from monoloco_b200 import synthetic, _lib as L_
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=150) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
    kps = synthetic.make_keypoints(n_total, seed=5)
    ref_raw = O.loco_model_forward(sd, O.preprocess_monoloco(kps, synthetic.KITTI_K))
    ref = O.extract_outputs(ref_raw)
    for rank in (0, 1):
        for mode in ('nccl', 'fused'):
            rows = got[rank][mode]
            assert rows.shape == (n_total, L_.GATHER_LD)
            ok, worst = O.close(rows[:, :9], ref_raw)
            assert ok, (rank, mode, worst)
            ok, worst = O.close(rows[:, L_.GATHER_DEC:L_.GATHER_DEC + 4], ref['xyzd'], col_scale=False)
            assert ok, (rank, mode, worst)
    assert np.array_equal(got[0]['fused'], got[1]['fused']) and np.array_equal(got[0]['nccl'], got[0]['fused'])
