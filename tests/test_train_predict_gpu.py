"""GPU: the reference's integration scenario (tests/test_train_mono.py:15-39, test_train_stereo.py:15-30) without the
openpifpaf front-end: train on the reference's own sample joints (331 / 406 instances, fixture X/Y), save the
state_dict the way Trainer does (trainer.py:242), load it into `Loco`, predict on the pifpaf fixture, post-process."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _train(mode, epochs, lr, tmp_path):
    from monoloco_b200.network.architectures import LocoModel
    from monoloco_b200.train import CompositeLoss, MultiTaskLoss
    f = np.load(os.path.join(GOLDEN, 'kat_%s_train.npz' % mode))
    v = np.load(os.path.join(GOLDEN, 'kat_%s_val.npz' % mode))
    tasks = ('d', 'x', 'y', 'h', 'w', 'l', 'ori') + (('aux',) if mode == 'stereo' else ())
    isz, osz = (34, 9) if mode == 'mono' else (68, 10)
    torch.manual_seed(1)
    model = LocoModel(isz, osz, linear_size=1024, p_dropout=0.2, num_stage=3).cuda()  # run.py defaults: hidden 1024, 3 stages
    mt = MultiTaskLoss(*CompositeLoss(tasks)(), (1,) * len(tasks), tasks)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=30, gamma=0.98)
    x, y = torch.from_numpy(f['X']).cuda(), torch.from_numpy(f['Y']).cuda()
    xv, yv = torch.from_numpy(v['X']).cuda(), torch.from_numpy(v['Y']).cuda()
    hist = []
    for _ in range(epochs):
        model.train()  # trainer.py:147-161: one mini-batch per epoch at the default bs=512 (331 / 406 rows)
        opt.zero_grad()
        loss, _ = mt(model(x), y, phase='train')
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 3)
        opt.step()
        sched.step()
        model.eval()
        with torch.no_grad():
            lv, vals = mt(model(xv), yv, phase='val')
        hist.append((float(loss), float(lv), float(vals[0])))
    path = str(tmp_path / ('%s-test.pkl' % ('monoloco_pp' if mode == 'mono' else 'monstereo')))
    torch.save(model.state_dict(), path)
    return path, hist


def test_train_then_predict_mono(tmp_path):
    from monoloco_b200.network import Loco, preprocess_pifpaf, load_calibration
    path, hist = _train('mono', epochs=10, lr=0.001, tmp_path=tmp_path)
    assert all(np.isfinite(h).all() for h in hist)
    # the train loss goes down; the eval-mode loss is NOT asserted: with 10 running-stat updates (momentum 0.1) the real
    # reference's val loss swings between 27 and 5955 in the same 10 epochs (checked with /root/reference on CPU)
    assert hist[-1][0] < hist[0][0]
    with open(os.path.join(GOLDEN, 'pifpaf_002282.json')) as fh:
        boxes, keypoints = preprocess_pifpaf(json.load(fh), im_size=(1238, 374))
    kk = load_calibration('kitti', (1238, 374))
    for n_dropout in (0, 10):
        net = Loco(model=path, mode='mono', device=torch.device('cuda'), n_dropout=n_dropout)
        dic = net.forward(keypoints, kk)
        out = Loco.post_process(dic, boxes, keypoints, kk)
        assert len(out['xyz_pred']) == 16 and len(out['angles']) == 16 and len(out['stds_epi']) == 16
        assert np.isfinite(np.array(out['dds_pred'])).all() and np.isfinite(np.array(out['confs'])).all()
        if n_dropout:
            assert (np.array(out['stds_epi']) > 0).all()


def test_train_then_predict_stereo(tmp_path):
    from monoloco_b200.network import Loco
    path, hist = _train('stereo', epochs=20, lr=0.002, tmp_path=tmp_path)
    assert hist[-1][0] < hist[0][0]
    v = np.load(os.path.join(GOLDEN, 'kat_stereo_val.npz'))
    left, right = v['kps'][:10, :, :17].tolist(), v['kps'][:7, :, 17:].tolist()
    net = Loco(model=path, mode='stereo', device=torch.device('cuda'))
    dic = net.forward(left, v['K'][0].tolist(), right)
    assert dic['xyzd'].shape[0] >= 10 and dic['aux'].shape[1] == 1
    assert ((dic['aux'] >= 0) & (dic['aux'] <= 1)).all() and np.isfinite(dic['d'].numpy()).all()


def test_device_loader_feeds_fused_training_loop(tmp_path):
    """N4 + A12 + N1 together: device-resident dataset -> one-launch train step -> fused clip + Adam (3 launches / batch),
    every row seen exactly once per epoch, no per-batch host->device copy."""
    from monoloco_b200 import synthetic
    from monoloco_b200.network.architectures import LocoModel
    from monoloco_b200.train import DeviceLoader, KeypointsDataset, FusedClipAdam, train_step
    path = str(tmp_path / 'joints.json')
    synthetic.make_joints_json(path, n_train=300, n_val=40, seed=3)
    ds = KeypointsDataset(path, 'train')
    loader = DeviceLoader(ds, batch_size=128, shuffle=True, device='cuda')
    torch.manual_seed(0)
    model = LocoModel(34, 9, linear_size=256, p_dropout=0.2, num_stage=2).cuda().train()
    params = list(model.parameters())
    opt = FusedClipAdam(params, lr=1e-3, max_norm=3, clip_params=params)
    tasks = ('d', 'x', 'y', 'h', 'w', 'l', 'ori')
    for epoch in range(2):
        seen = []
        for inputs, labels, names, kps in loader:
            assert inputs.is_cuda and labels.is_cuda and kps.is_cuda and names is None
            assert inputs.shape[1] == 34 and labels.shape[1] == 10 and kps.shape[1:] == (3, 17)
            seen += [int(v) for v in inputs[:, 0].tolist()]
            loss, vals, out = train_step(model, inputs, labels, tasks)
            opt.step()
            assert torch.isfinite(loss) and out.shape == (inputs.shape[0], 9)
        assert sorted(seen) == list(range(300)) and len(loader) == 3
