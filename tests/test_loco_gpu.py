"""GPU: the reference-facing Python API (Loco, LocoModel, MonolocoModel, process helpers) against the
live-reference fixtures -- these read like the reference's own usage (predict.py:166-171,231-245;
generate_kitti.py:41-48,104-120)."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _load_module(kind, isz, osz, L, st, seed):
    from monoloco_b200 import synthetic
    from monoloco_b200.network.architectures import LocoModel, MonolocoModel
    sd = synthetic.make_state_dict(kind, isz, osz, L, st, seed)
    m = LocoModel(isz, osz, L, num_stage=st) if kind == 'loco' else MonolocoModel(isz, osz, L, num_stage=st)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    return m, sd


def _cmp(O, dic, f, prefix, keys):
    for k in keys:
        ok, worst = O.close(dic[k].numpy(), f[prefix + k], col_scale=(k != 'xyzd'))
        assert ok, (k, worst)
    assert O.angle_close(dic['yaw'][0].numpy(), f[prefix + 'yaw_pred'])[0]
    assert O.angle_close(dic['yaw'][1].numpy(), f[prefix + 'yaw_orig'], rtol=3e-5)[0]


def test_loco_forward_mono_from_checkpoint_path(tmp_path):
    """Loco(model=<path to state_dict pickle>) as predict.py / generate_kitti.py construct it."""
    from oracle import loco_oracle as O
    from monoloco_b200.network import Loco, preprocess_pifpaf
    f = np.load(os.path.join(GOLDEN, 'ref_loco_mono_pifpaf.npz'))
    m, sd = _load_module('loco', 34, 9, 1024, 3, 1)
    path = str(tmp_path / 'monoloco_pp-test.pkl')
    torch.save(m.state_dict(), path)  # trainer.py:242
    net = Loco(model=path, mode='mono', device=torch.device('cuda'))
    with open(os.path.join(GOLDEN, 'pifpaf_002282.json')) as fh:
        boxes, keypoints = preprocess_pifpaf(json.load(fh), im_size=(1238, 374))
    kk = f['K'].tolist()
    dic = net.forward(keypoints, kk)
    assert all(not v.is_cuda for k, v in dic.items() if isinstance(v, torch.Tensor))
    _cmp(O, dic, f, 'out_', ('xyzd', 'bi', 'd', 'h', 'w', 'l', 'ori'))
    assert dic['epi'] == [0.] * 16
    assert net.forward([], kk) is None  # net.py:88-89
    ok, worst = O.close(dic['xyz_c'].numpy(), f['xyz_from_distance'], col_scale=False)
    assert ok, worst

    # post_process against the reference's output on the same inputs
    with open(os.path.join(GOLDEN, 'ref_api.json')) as fh:
        api = json.load(fh)
    for gt, key in ((api['dic_gt'], 'post'), (None, 'post_nogt')):
        post = Loco.post_process(dic, boxes, keypoints, kk, dic_gt=gt)
        ref = api[key]
        assert sorted(post.keys()) == sorted(ref.keys())
        assert post['gt'] == ref['gt'] and post['uv_centers'] == ref['uv_centers'] and post['uv_heads'] == ref['uv_heads']
        assert post['uv_shoulders'] == ref['uv_shoulders'] and post['boxes'] == ref['boxes']
        for k in ('confs', 'dds_pred', 'stds_ale', 'xyz_pred', 'angles', 'angles_egocentric'):
            assert np.allclose(np.array(post[k]), np.array(ref[k]), rtol=3e-5, atol=2e-4), k
        if gt:
            assert post['dds_real'] == ref['dds_real'] and post['boxes_gt'] == ref['boxes_gt']
            assert np.allclose(np.array(post['xyz_real']), np.array(ref['xyz_real']), rtol=1e-5)


def test_loco_forward_stereo():
    from oracle import loco_oracle as O
    from monoloco_b200.network import Loco
    f = np.load(os.path.join(GOLDEN, 'ref_loco_stereo.npz'))
    m, sd = _load_module('loco', 68, 10, 1024, 3, 2)
    net = Loco(model=m, mode='stereo', device=torch.device('cuda'))
    dic = net.forward(f['left'].tolist(), f['K'].tolist(), f['right'].tolist())
    _cmp(O, dic, f, 'out_', ('xyzd', 'bi', 'd', 'aux', 'ori', 'h', 'w', 'l'))
    assert dic['epi'] == [0.] * 12
    dic1 = net.forward(f['left'].tolist(), f['K'].tolist(), None)  # no right poses: net.py:115-116
    _cmp(O, dic1, f, 'noright_', ('xyzd', 'bi', 'd', 'aux'))


def test_module_forward_eval_matches_reference_module():
    """nn.Module.forward in eval mode (Trainer.evaluate, trainer.py:199-231): CPU tensor in -> CPU tensor out."""
    from oracle import loco_oracle as O
    for name in ('loco_mono_l1024', 'loco_stereo_l128', 'monoloco_l256_o2', 'monoloco_l1024_o9'):
        f = np.load(os.path.join(GOLDEN, 'ref_fwd_%s.npz' % name))
        isz, osz, L, st, seed = [int(v) for v in f['cfg'][:5]]
        m, _ = _load_module(str(f['kind']), isz, osz, L, st, seed)
        m.eval().cuda()
        with torch.no_grad():
            out = m(torch.from_numpy(f['x']))
        assert not out.is_cuda and out.shape == f['out'].shape
        ok, worst = O.close(out.numpy(), f['out'])
        assert ok, (name, worst)
        out2 = m(torch.from_numpy(f['x']).cuda())
        assert out2.is_cuda and torch.equal(out2.cpu(), out)


def test_module_reload_after_weight_change():
    """load_state_dict / in-place parameter updates invalidate the packed device copy."""
    from oracle import loco_oracle as O
    from monoloco_b200 import synthetic
    m, sd = _load_module('loco', 34, 9, 256, 2, 4)
    m.eval().cuda()
    x = synthetic.make_inputs(33, 34, seed=2)
    with torch.no_grad():
        a = m(torch.from_numpy(x)).numpy()
        assert O.close(a, O.model_forward(sd, x))[0]
        sd2 = synthetic.make_state_dict('loco', 34, 9, 256, 2, 5)
        m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd2.items()})
        b = m(torch.from_numpy(x)).numpy()
        assert O.close(b, O.model_forward(sd2, x))[0]
        m.w_fin.bias.add_(1.0)
        c = m(torch.from_numpy(x)).numpy()
        assert np.allclose(c[:, :8], b[:, :8] + 1.0, atol=1e-5)


def test_epistemic_uncertainty_statistics():
    """MC dropout (net.py:135-161): n_dropout passes + 100 Laplace samples each; compared in distribution with the
    analytic mixture std computed by the oracle from explicit-mask passes."""
    from oracle import loco_oracle as O
    from monoloco_b200 import synthetic, _lib as L_
    from monoloco_b200.network import Loco
    m, sd = _load_module('loco', 34, 9, 1024, 3, 1)
    n_drop = 20
    net = Loco(model=m, mode='mono', device=torch.device('cuda'), n_dropout=n_drop)
    kps = synthetic.make_keypoints(40, seed=3)
    dic = net.forward(kps.tolist(), synthetic.KITTI_K)
    epi = dic['epi'].numpy()
    assert epi.shape == (40,) and np.isfinite(epi).all() and (epi > 0).all()
    # dropout must be off again afterwards and outputs deterministic
    assert net.model.dropout.training is False
    # analytic expectation: Var = mean_n(2 b_n^2) + var_n(mu_n), with the engine's own stochastic passes
    eng = net.model.engine()
    x = torch.from_numpy(O.preprocess_monoloco(kps, synthetic.KITTI_K)).cuda()
    out = eng.forward(x.repeat(n_drop, 1), dropout=True, drop_seed=1)  # all passes as independent rows of one launch
    mus = out['raw'][:, 2].cpu().numpy().reshape(n_drop, -1)
    bis = np.abs(out['dec'][:, 4].cpu().numpy().reshape(n_drop, -1))
    assert np.abs(mus - mus[0]).max() > 1e-4  # every replica drew its own dropout mask
    expect = np.sqrt((2 * bis ** 2).mean(0) + mus.var(0))
    # 20 x 100 draws per row: the std estimator of a Laplace mixture has a 1-sigma error of sqrt(5 / (4 * 2000)) = 2.5 %;
    # 40 rows -> allow 4 sigma.  The sampler itself is checked to 1 % in test_laplace_std_kernel_large_sample.
    assert np.allclose(epi, expect, rtol=0.10), np.abs(epi / expect - 1).max()


def test_laplace_std_kernel_large_sample():
    """The device sampler alone (mlb_laplace_std: inverse-CDF Laplace(mu, |b|) draws, counter RNG) at 200 000 draws per row:
    std = sqrt(2) |b| within 1 % (3 sigma of the estimator is 0.75 %), independent of mu and of the sign of b; and the
    mixture over passes follows Var = mean_n(2 b_n^2) + var_n(mu_n) (net.py:150-158 concatenates the passes).  The
    reference's own sampler is pinned to the same Laplace(mu, b) by tests/test_oracle_golden.py (KS test on its draws)."""
    import ctypes as C
    from monoloco_b200 import _lib as L_
    lib = L_.lib()
    rows = 64
    rng = np.random.RandomState(0)
    mu = rng.uniform(2, 60, rows).astype(np.float32)
    b = (rng.uniform(0.05, 6, rows) * rng.choice([-1, 1], rows)).astype(np.float32)
    d_bi = torch.from_numpy(np.stack([mu, b], 1)[None].copy()).cuda()   # [1 pass, rows, 2]
    std = torch.empty(rows, dtype=torch.float32, device='cuda')
    L_.check(lib.mlb_laplace_std(d_bi.data_ptr(), 1, rows, 200000, 7, std.data_ptr(), None), 'laplace_std')
    got = std.cpu().numpy()
    assert np.allclose(got, np.sqrt(2) * np.abs(b), rtol=1e-2), np.abs(got / (np.sqrt(2) * np.abs(b)) - 1).max()
    n_pass = 8
    mus = rng.uniform(5, 40, (n_pass, rows)).astype(np.float32)
    bs = rng.uniform(0.1, 3, (n_pass, rows)).astype(np.float32)
    d_bi = torch.from_numpy(np.stack([mus, bs], 2).copy()).cuda()
    L_.check(lib.mlb_laplace_std(d_bi.data_ptr(), n_pass, rows, 50000, 9, std.data_ptr(), None), 'laplace_std')
    expect = np.sqrt((2 * bs.astype(np.float64) ** 2).mean(0) + mus.astype(np.float64).var(0))
    assert np.allclose(std.cpu().numpy(), expect, rtol=1e-2)


def test_process_helpers():
    from oracle import loco_oracle as O
    from monoloco_b200.network.process import (preprocess_monoloco, preprocess_monstereo, extract_outputs,
                                               unnormalize_bi, cluster_outputs)
    f = np.load(os.path.join(GOLDEN, 'ref_loco_stereo.npz'))
    x, clusters = preprocess_monstereo(f['left'], f['right'], f['K'])
    assert clusters == [9] * 12 and np.abs(x.cpu().numpy() - f['pairs_x']).max() < 6e-6
    g = np.load(os.path.join(GOLDEN, 'ref_fwd_loco_stereo_l1024.npz'))
    dic = extract_outputs(torch.from_numpy(g['out']))
    for k in ('xyzd', 'bi', 'd', 'aux'):
        assert O.close(dic[k].numpy(), g['dec_' + k], col_scale=(k != 'xyzd'))[0], k
    cols = extract_outputs(torch.from_numpy(g['out']), tasks=('d', 'ori', 'aux'))
    assert cols[0].shape[1] == 2 and cols[1].shape[1] == 2 and cols[2].shape[1] == 1
    with pytest.raises(AssertionError):
        unnormalize_bi(torch.zeros(3, 3))
    with pytest.raises(AssertionError):
        cluster_outputs(torch.zeros(7, 10), 3)
    kat = np.load(os.path.join(GOLDEN, 'kat_mono_val.npz'))
    k = kat['K'][0]
    rows = np.where((kat['K'].reshape(-1, 9) == k.reshape(-1)).all(1))[0]
    xk = preprocess_monoloco(kat['kps'][rows].tolist(), k.tolist())
    assert np.abs(xk.cpu().numpy() - kat['X'][rows]).max() < 4e-6
