"""CPU: the numpy oracle (oracle/loco_oracle.py) against the golden fixtures produced by the real
reference (oracle/gen_golden.py) and against the reference's own known-answer fixtures."""
import glob
import os

import numpy as np
import pytest

from oracle import loco_oracle as O
from monoloco_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _sd(f):
    isz, osz, L, st, seed = [int(v) for v in f['cfg'][:5]]
    sd = synthetic.make_state_dict(str(f['kind']) if 'kind' in f.files else 'loco', isz, osz, L, st, seed)
    chk = float(sum(float(np.asarray(v, dtype=np.float64).sum()) for k, v in sorted(sd.items())))
    assert abs(chk - float(f['checksum'])) <= 1e-9 * max(1.0, abs(chk)), "synthetic weights drifted"
    return sd


@pytest.mark.parametrize('mode,phase', [('mono', 'train'), ('mono', 'val'), ('stereo', 'train'), ('stereo', 'val')])
def test_preprocess_kat(mode, phase):
    """Reference fixture KAT: stored X == preprocess(kps, K)."""
    f = np.load(os.path.join(GOLDEN, 'kat_%s_%s.npz' % (mode, phase)))
    kps, K, X = f['kps'], f['K'], f['X']
    worst = 0.0
    for k in np.unique(K.reshape(-1, 9), axis=0):
        rows = np.where((K.reshape(-1, 9) == k).all(1))[0]
        kk = k.reshape(3, 3)
        if mode == 'mono':
            x = O.preprocess_monoloco(kps[rows], kk)
        else:
            le = O.preprocess_monoloco(kps[rows][:, :, :17], kk)
            ri = O.preprocess_monoloco(kps[rows][:, :, 17:], kk)
            x = np.concatenate([le, le - ri], axis=1)
        worst = max(worst, float(np.abs(x - X[rows]).max()))
    assert worst < 4e-6, worst


def test_pixel_to_camera_linearity():
    """reference tests/test_utils.py:18-25."""
    kk = np.array(synthetic.KITTI_K, dtype=np.float32)
    uv = np.array([[100., 50.], [700., 300.]], dtype=np.float32)
    a = O.pixel_to_camera(uv, kk, 1) * np.float32(7.5)
    b = O.pixel_to_camera(uv, kk, 7.5)
    assert np.allclose(a, b, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN, 'ref_fwd_*.npz')) + glob.glob(os.path.join(GOLDEN, 'ref_wide_*.npz'))))
def test_forward_vs_reference(path):
    """The oracle's network against the real nn.Module: the default widths (ref_fwd_*) and the widths outside the default
    that `--hidden_size` allows (ref_wide_*: 2048, 1500, 300, 200)."""
    f = np.load(path)
    sd = _sd(f)
    out = O.model_forward(sd, f['x'])
    ok, worst = O.close(out, f['out'], rtol=1e-5, atol=1e-6)
    assert ok, (path, worst)
    if 'dec_xyzd' in f.files:
        dec = O.extract_outputs(f['out']) if str(f['kind']) == 'loco' else O.extract_outputs_mono(f['out'])
        for k in ('xyzd', 'bi', 'd', 'h', 'w', 'l', 'ori'):
            ok, worst = O.close(dec[k], f['dec_' + k], rtol=2e-6, atol=1e-6)
            assert ok, (path, k, worst)
        assert O.angle_close(dec['yaw'][0], f['dec_yaw_pred'], rtol=2e-6)[0]
        assert O.angle_close(dec['yaw'][1], f['dec_yaw_orig'], rtol=2e-6)[0]
        if 'dec_aux' in f.files:
            assert O.close(dec['aux'], f['dec_aux'], rtol=2e-6)[0]


def test_loco_forward_mono_pifpaf():
    f = np.load(os.path.join(GOLDEN, 'ref_loco_mono_pifpaf.npz'))
    sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 1)
    dic = O.loco_forward(sd, f['keypoints'], f['K'], mode='mono')
    for k in ('xyzd', 'bi', 'd', 'h', 'w', 'l', 'ori'):
        ok, worst = O.close(dic[k], f['out_' + k])
        assert ok, (k, worst)
    assert O.angle_close(dic['yaw'][0], f['out_yaw_pred'])[0]
    assert O.angle_close(dic['yaw'][1], f['out_yaw_orig'])[0]
    assert len(dic['epi']) == 16
    uvc = O.get_keypoints(f['keypoints'], 'center')
    xyc = O.pixel_to_camera(uvc, f['K'], 1)
    assert O.close(xyc, f['xy_centers'])[0]
    assert O.close(O.xyz_from_distance(dic['d'], xyc), f['xyz_from_distance'])[0]
    assert O.loco_forward(sd, [], f['K']) is None  # net.py:88-89


def test_loco_forward_stereo():
    f = np.load(os.path.join(GOLDEN, 'ref_loco_stereo.npz'))
    sd = synthetic.make_state_dict('loco', 68, 10, 1024, 3, 2)
    x, clusters = O.preprocess_monstereo(f['left'], f['right'], f['K'])
    assert clusters == [9] * 12
    assert O.close(x, f['pairs_x'], rtol=1e-6, atol=4e-6)[0]
    raw = O.model_forward(sd, f['pairs_x'])
    assert O.close(raw, f['pairs_raw'])[0]
    fin, mask = O.filter_outputs(O.cluster_outputs(f['pairs_raw'], 9))
    assert (mask == f['filter_mask']).all() and np.array_equal(fin, f['filter_out'])
    dic = O.loco_forward(sd, f['left'], f['K'], f['right'], mode='stereo')
    for k in ('xyzd', 'bi', 'd', 'aux', 'ori'):
        ok, worst = O.close(dic[k], f['out_' + k])
        assert ok, (k, worst)
    dic1 = O.loco_forward(sd, f['left'], f['K'], None, mode='stereo')
    assert O.close(dic1['xyzd'], f['noright_xyzd'])[0]


@pytest.mark.parametrize('mode', ['mono', 'stereo'])
@pytest.mark.parametrize('auto', [False, True])
def test_loss_values(mode, auto):
    f = np.load(os.path.join(GOLDEN, 'ref_train_%s_%s.npz' % (mode, 'auto' if auto else 'mtl')))
    tasks = ('d', 'x', 'y', 'h', 'w', 'l', 'ori') + (('aux',) if mode == 'stereo' else ())
    ls = f['log_sigmas'] if auto else None
    loss, vals = O.multi_task_loss(f['out'], f['y'], tasks, log_sigmas=ls)
    assert abs(float(loss) - float(f['loss'])) <= 2e-6 * abs(float(f['loss']))
    assert np.allclose(np.array(vals, dtype=np.float64), f['vals'], rtol=3e-6)


def test_laplace_population_std():
    f = np.load(os.path.join(GOLDEN, 'ref_laplace_sampling.npz'))
    est = f['samples'].std(0, ddof=1)
    pop = O.laplace_std(f['mu_bi'][:, 1])
    assert np.allclose(est, pop, rtol=0.35)  # 100 samples: statistical agreement only
    assert np.allclose(O.unnormalize_bi(np.array([[10.0, -1.0], [25.0, 0.2]], dtype=np.float32)), f['bi'], rtol=1e-6)


@pytest.mark.parametrize('mode', ['mono', 'stereo'])
@pytest.mark.parametrize('auto', [False, True])
def test_torch_port_train_step(mode, auto):
    """oracle/torch_port.py (autograd) vs the live reference: loss, per-task values, d(out), every parameter grad,
    BN running-stat update."""
    import torch
    from oracle import torch_port as T
    f = np.load(os.path.join(GOLDEN, 'ref_train_%s_%s.npz' % (mode, 'auto' if auto else 'mtl')))
    isz, osz, L, st, seed, B = [int(v) for v in f['cfg']]
    tasks = ('d', 'x', 'y', 'h', 'w', 'l', 'ori') + (('aux',) if mode == 'stereo' else ())
    sd = T.to_torch(synthetic.make_state_dict('loco', isz, osz, L, st, seed), requires_grad=True)
    ls = torch.tensor(f['log_sigmas'], requires_grad=True) if auto else None
    out = T.model_forward(sd, torch.from_numpy(f['x']), training=True, p_dropout=0.0)
    loss, vals = T.multi_task_loss(out, torch.from_numpy(f['y']), tasks, log_sigmas=ls)
    loss.backward()
    assert abs(float(loss) - float(f['loss'])) <= 2e-6 * abs(float(f['loss']))
    assert np.allclose(out.detach().numpy(), f['out'], rtol=1e-5, atol=1e-5)
    for k in f.files:
        if k.startswith('grad.') and k != 'grad.log_sigmas':
            g = sd[k[5:]].grad.numpy()
            assert np.allclose(g, f[k], rtol=1e-4, atol=1e-6 * max(1.0, float(np.abs(f[k]).max()))), k
        if k.startswith('buf.') and 'num_batches' not in k:
            assert np.allclose(sd[k[4:]].detach().numpy(), f[k], rtol=1e-5, atol=1e-6), k
    if auto:
        assert np.allclose(ls.grad.numpy(), f['grad.log_sigmas'], rtol=1e-5)


def test_reference_laplace_sampler_is_laplace_mu_b():
    """The distribution our device sampler must reproduce (process.py:101-122): the live reference's 100 draws per (mu, b)
    pass a Kolmogorov-Smirnov test against Laplace(mu, scale=b) -- and fail it against the neighbouring hypotheses
    (scale b / sqrt(2), i.e. "b is the std", and scale 2 b) when the three columns are pooled after standardisation."""
    f = np.load(os.path.join(GOLDEN, 'ref_laplace_sampling.npz'))
    mu, b, xs = f['mu_bi'][:, 0].astype(np.float64), f['mu_bi'][:, 1].astype(np.float64), f['samples'].astype(np.float64)

    def ks(z):   # KS distance of standardised draws z = (x - mu) / scale against Laplace(0, 1)
        z = np.sort(z)
        cdf = np.where(z < 0, 0.5 * np.exp(z), 1 - 0.5 * np.exp(-z))
        n = len(z)
        return max(np.max(np.arange(1, n + 1) / n - cdf), np.max(cdf - np.arange(0, n) / n))

    crit = 1.63 / np.sqrt(xs.shape[0])   # alpha = 0.01
    for j in range(3):
        assert ks((xs[:, j] - mu[j]) / b[j]) < crit, j
    pooled = lambda s: np.concatenate([(xs[:, j] - mu[j]) / (s * b[j]) for j in range(3)])  # noqa: E731
    crit3 = 1.63 / np.sqrt(3 * xs.shape[0])
    assert ks(pooled(1.0)) < crit3
    assert ks(pooled(1 / np.sqrt(2))) > crit3 and ks(pooled(2.0)) > crit3
