"""GPU parity tests of the fused forward (through the C ABI) against the numpy oracle and the
live-reference golden fixtures.  Rule: |a-b| <= 1e-5 * max(|b|, column scale) + 1e-6;
yaw compared modulo 2*pi."""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _mods():
    from oracle import loco_oracle as O
    from monoloco_b200 import synthetic, engine, _lib
    return O, synthetic, engine, _lib


def _check_dec(O, dec, ref, stereo, raw_ref=None):
    """dec: [B,8] numpy from the kernel; ref: oracle dict; raw_ref: the reference raw outputs [B,out] (conditions the
    angle tolerances: atan2 of two raw outputs that are both near zero is ill-conditioned for ANY fp32 implementation)."""
    # x, y, z are d * (products of sin/cos): their error scale is |d| (cos(theta) ~ 0 amplifies the relative error
    # of x), so the whole xyzd block is compared against one scale = max |d| (SURVEY.md §0.5 "per-column scale s")
    ok, worst = O.close(dec[:, 0:4], ref['xyzd'], col_scale=False)
    assert ok, ('xyzd', worst)
    ok, worst = O.close(dec[:, 4:5], ref['bi'])
    assert ok, ('bi', worst)
    rad = lin = rad2 = lin2 = None
    if raw_ref is not None and raw_ref.shape[1] >= 9:
        # yaw_pred = atan2(o7, o8): the parity rule grants each of them 1e-5 * column scale + 1e-6
        rad = np.hypot(raw_ref[:, 7:8], raw_ref[:, 8:9])
        lin = 1e-5 * float(np.abs(raw_ref[:, 7:9]).max()) + 1e-6
        # yaw_orig adds atan2(x, z): x, z carry 1e-5 * max|d|
        xyzd = np.asarray(ref['xyzd'])
        rad2 = np.minimum(rad / lin, np.hypot(xyzd[:, 0:1], xyzd[:, 2:3]) / (1e-5 * float(np.abs(xyzd[:, 3]).max()) + 1e-6))
        lin2 = 2.0  # radius already divided by the tolerances: two such terms add up
    ok, worst = O.angle_close(dec[:, 5:6], ref['yaw'][0], radius=rad, lin_tol=lin)
    assert ok, ('yaw_pred', worst)
    ok, worst = O.angle_close(dec[:, 6:7], ref['yaw'][1], rtol=3e-5, radius=rad2, lin_tol=lin2)  # atan2(x,z) amplifies the 1e-5 of x,z
    assert ok, ('yaw_orig', worst)
    if stereo:
        ok, worst = O.close(dec[:, 7:8], ref['aux'])
        assert ok, ('aux', worst)


@pytest.fixture(scope='module')
def mono1024():
    O, synthetic, engine, _ = _mods()
    sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
    return sd, engine.LocoEngine(sd)


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN, 'ref_fwd_*.npz'))))
def test_forward_golden(path):
    """Same weights + inputs as the real reference nn.Module (fixtures from oracle/gen_golden.py)."""
    O, synthetic, engine, L_ = _mods()
    f = np.load(path)
    isz, osz, L, st, seed = [int(v) for v in f['cfg'][:5]]
    sd = synthetic.make_state_dict(str(f['kind']), isz, osz, L, st, seed)
    eng = engine.LocoEngine(sd)
    out = eng.forward(torch.from_numpy(f['x']).cuda())
    raw = out['raw'].cpu().numpy()
    ok, worst = O.close(raw, f['out'])
    assert ok, (path, worst)
    if 'dec_xyzd' in f.files:
        dec = out['dec'].cpu().numpy()
        ref = {'xyzd': f['dec_xyzd'], 'bi': f['dec_bi'], 'yaw': (f['dec_yaw_pred'], f['dec_yaw_orig'])}
        if 'dec_aux' in f.files:
            ref['aux'] = f['dec_aux']
        _check_dec(O, dec, ref, 'dec_aux' in f.files, raw_ref=f['out'])
    eng.close()


@pytest.mark.parametrize('B', [1, 5, 27, 28, 29, 32, 33, 148, 256, 1000, 4096])
def test_forward_batches_vs_oracle(mono1024, B):
    O, synthetic, engine, L_ = _mods()
    sd, eng = mono1024
    x = synthetic.make_inputs(B, 34, seed=B)
    ref = O.loco_model_forward(sd, x)
    out = eng.forward(torch.from_numpy(x).cuda())
    ok, worst = O.close(out['raw'].cpu().numpy(), ref)
    assert ok, worst
    _check_dec(O, out['dec'].cpu().numpy(), O.extract_outputs(ref), False, raw_ref=ref)


@pytest.mark.parametrize('tm', [8, 10, 12, 14, 16])
def test_forward_tile_shapes(mono1024, tm):
    """Every rows-per-group instantiation gives the same answer (ragged last tile included)."""
    O, synthetic, engine, L_ = _mods()
    sd, eng = mono1024
    B = 613
    x = synthetic.make_inputs(B, 34, seed=7)
    ref = O.loco_model_forward(sd, x)
    out = eng.forward(torch.from_numpy(x).cuda(), rows_per_group=tm)
    ok, worst = O.close(out['raw'].cpu().numpy(), ref)
    assert ok, (tm, worst)


def test_forward_residual_in_tensor_memory(mono1024):
    O, synthetic, engine, L_ = _mods()
    sd, eng = mono1024
    x = torch.from_numpy(synthetic.make_inputs(900, 34, seed=9)).cuda()
    a = eng.forward(x, res_tmem=False, kernel='tile')['raw']  # L2-resident global scratch
    b = eng.forward(x, res_tmem=True, kernel='tile')['raw']   # Tensor Memory (the default)
    assert torch.equal(a, b)  # same kernel, same summation order: the stash location must not change a bit


def test_empty_batch(mono1024):
    sd, eng = mono1024
    out = eng.forward(torch.empty((0, 34), dtype=torch.float32, device='cuda'))
    assert out['raw'].shape == (0, 9)


def test_preprocess_kat_mono():
    """Reference fixture KAT through the stand-alone kernel and through the fused prologue."""
    O, synthetic, engine, L_ = _mods()
    f = np.load(os.path.join(GOLDEN, 'kat_mono_val.npz'))
    sd = synthetic.make_state_dict('loco', 34, 9, 128, 1, 5)
    eng = engine.LocoEngine(sd)
    for k in np.unique(f['K'].reshape(-1, 9), axis=0):
        rows = np.where((f['K'].reshape(-1, 9) == k).all(1))[0]
        kps = torch.from_numpy(f['kps'][rows]).cuda()
        x1 = engine.preprocess_device(kps, k.reshape(3, 3)).cpu().numpy()
        assert np.abs(x1 - f['X'][rows]).max() < 4e-6
        out = eng.forward(kps, kk=k.reshape(3, 3), kind=L_.IN_KPS, want_x=True, want_xyzc=True)
        x2 = out['x'].cpu().numpy()
        assert np.abs(x2 - f['X'][rows]).max() < 4e-6
        ref = O.loco_model_forward(sd, f['X'][rows])
        ok, worst = O.close(out['raw'].cpu().numpy(), ref)
        assert ok, worst
        # zero-centred legacy variant (net.py:96)
        xz = engine.preprocess_device(kps, k.reshape(3, 3), zero_center=True).cpu().numpy()
        assert np.abs(xz - O.preprocess_monoloco(f['kps'][rows], k.reshape(3, 3), zero_center=True)).max() < 4e-6
        # xyz_from_distance on the bbox-centre ray (net.py:195,213)
        uvc = O.get_keypoints(f['kps'][rows], 'center')
        xyc = O.pixel_to_camera(uvc, k.reshape(3, 3), 1)
        xyz = O.xyz_from_distance(ref[:, 2:3], xyc)
        ok, worst = O.close(out['xyzc'].cpu().numpy()[:, :3], xyz)
        assert ok, worst
    eng.close()


def test_stereo_pairs_and_filter():
    """all-vs-all pair rows built in the kernel prologue + arg-max filter vs the live-reference fixture."""
    O, synthetic, engine, L_ = _mods()
    f = np.load(os.path.join(GOLDEN, 'ref_loco_stereo.npz'))
    sd = synthetic.make_state_dict('loco', 68, 10, 1024, 3, 2)
    eng = engine.LocoEngine(sd)
    left, right = torch.from_numpy(f['left']).cuda(), torch.from_numpy(f['right']).cuda()
    out = eng.forward(left, x_right=right, kk=f['K'], kind=L_.IN_KPS_STEREO, want_x=True)
    assert np.abs(out['x'].cpu().numpy() - f['pairs_x']).max() < 6e-6
    ok, worst = O.close(out['raw'].cpu().numpy(), f['pairs_raw'])
    assert ok, worst
    sel_raw, sel_dec, sel_idx = eng.stereo_filter(out['raw'], out['dec'], 12, 9)
    assert np.array_equal(np.where(f['filter_mask'].reshape(-1))[0], sel_idx.cpu().numpy())
    ok, worst = O.close(sel_raw.cpu().numpy(), f['filter_out'])
    assert ok, worst
    ref = {'xyzd': f['out_xyzd'], 'bi': f['out_bi'], 'yaw': (f['out_yaw_pred'], f['out_yaw_orig']), 'aux': f['out_aux']}
    _check_dec(O, sel_dec.cpu().numpy(), ref, True)
    # ties are all kept, in row-major order (process.py:324-326)
    raw = out['raw'].clone()
    raw[:, 9] = 0.25
    s_raw, _, s_idx = eng.stereo_filter(raw, out['dec'], 12, 9)
    assert s_idx.cpu().numpy().tolist() == list(range(108))
    eng.close()


def test_stereo_kat_diagonal():
    O, synthetic, engine, L_ = _mods()
    f = np.load(os.path.join(GOLDEN, 'kat_stereo_val.npz'))
    sd = synthetic.make_state_dict('loco', 68, 10, 128, 1, 6)
    eng = engine.LocoEngine(sd)
    for k in np.unique(f['K'].reshape(-1, 9), axis=0):
        rows = np.where((f['K'].reshape(-1, 9) == k).all(1))[0]
        le = torch.from_numpy(np.ascontiguousarray(f['kps'][rows][:, :, :17])).cuda()
        ri = torch.from_numpy(np.ascontiguousarray(f['kps'][rows][:, :, 17:])).cuda()
        n = len(rows)
        out = eng.forward(le, x_right=ri, kk=k.reshape(3, 3), kind=L_.IN_KPS_STEREO, want_x=True, want_dec=False)
        x = out['x'].cpu().numpy().reshape(n, n, 68)[np.arange(n), np.arange(n)]
        assert np.abs(x - f['X'][rows]).max() < 6e-6
    eng.close()


def test_mc_dropout_with_explicit_masks(mono1024):
    """MC-dropout pass (net.py:141: only the two top-level dropout sites) with torch-style keep masks."""
    O, synthetic, engine, L_ = _mods()
    sd, eng = mono1024
    B = 300
    rng = np.random.RandomState(3)
    masks = (rng.uniform(size=(2, B, 1024)) >= 0.2).astype(np.uint8)
    x = synthetic.make_inputs(B, 34, seed=31)
    ref = O.loco_model_forward(sd, x, drop_masks=(masks[0], masks[1]), p_dropout=0.2)
    out = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_mask=torch.from_numpy(masks).cuda())
    ok, worst = O.close(out['raw'].cpu().numpy(), ref)
    assert ok, worst
    # in-kernel RNG: keep rate ~ 0.8 changes the output, deterministic per seed
    a = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_seed=5)['raw']
    b = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_seed=5)['raw']
    c = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_seed=6)['raw']
    assert torch.equal(a, b) and not torch.equal(a, c)


def test_forward_host_buffers(mono1024):
    O, synthetic, engine, L_ = _mods()
    sd, eng = mono1024
    kps = synthetic.make_keypoints(777, seed=4)
    kk = synthetic.KITTI_K
    out = eng.forward_host(torch.from_numpy(kps).pin_memory(), kk=kk, kind=L_.IN_KPS, want_xyzc=True)
    ref = O.loco_forward(sd, kps, kk, mode='mono')
    _check_dec(O, out['dec'].numpy(), ref, False, raw_ref=O.loco_model_forward(sd, O.preprocess_monoloco(kps, kk)))


def test_full_size_properties(mono1024):
    """BASELINE size (65536): row independence -- replicated inputs give bit-identical outputs wherever the row
    lands in a tile / wave; and a spot-check of 512 rows against the oracle."""
    O, synthetic, engine, L_ = _mods()
    sd, eng = mono1024
    base = synthetic.make_inputs(4099, 34, seed=77)
    reps = 16
    x = torch.from_numpy(np.tile(base, (reps, 1))[:65536]).cuda()
    raw = eng.forward(x)['raw']
    first = raw[:4099]
    for r in range(1, 15):
        assert torch.equal(raw[r * 4099:(r + 1) * 4099], first)
    idx = np.random.RandomState(1).choice(4099, 512, replace=False)
    ok, worst = O.close(first.cpu().numpy()[idx], O.loco_model_forward(sd, base[idx]))
    assert ok, worst


# ------------------------------------------------------------------------------------------------ small-batch kernel
@pytest.mark.parametrize('B', [1, 5, 16, 17, 100, 256, 300, 1000])
def test_cluster_kernel_batches_vs_oracle(mono1024, B):
    """8-CTA-cluster kernel (forward_small.cu): column-split layers + DSMEM all-gather, forced on for every size
    (B > 288 exercises the multi-tile loop of a cluster)."""
    O, synthetic, engine, L_ = _mods()
    sd, eng = mono1024
    kps = synthetic.make_keypoints(B, seed=40 + B)
    out = eng.forward(torch.from_numpy(kps).cuda(), kk=synthetic.KITTI_K, kind=L_.IN_KPS, want_x=True, want_xyzc=True,
                      kernel='cluster')
    x = O.preprocess_monoloco(kps, synthetic.KITTI_K)
    assert np.abs(out['x'].cpu().numpy() - x).max() < 6e-6
    ref = O.loco_model_forward(sd, x)
    ok, worst = O.close(out['raw'].cpu().numpy(), ref)
    assert ok, worst
    _check_dec(O, out['dec'].cpu().numpy(), O.extract_outputs(ref), False, raw_ref=ref)
    tile = eng.forward(torch.from_numpy(kps).cuda(), kk=synthetic.KITTI_K, kind=L_.IN_KPS, want_xyzc=True, kernel='tile')
    assert torch.allclose(out['raw'], tile['raw'], rtol=2e-5, atol=2e-5) and torch.allclose(out['xyzc'], tile['xyzc'], rtol=1e-5, atol=1e-5)


def test_cluster_kernel_stereo_dropout_and_legacy_models():
    O, synthetic, engine, L_ = _mods()
    f = np.load(os.path.join(GOLDEN, 'ref_loco_stereo.npz'))
    sd = synthetic.make_state_dict('loco', 68, 10, 1024, 3, 2)
    eng = engine.LocoEngine(sd)
    left, right = torch.from_numpy(f['left']).cuda(), torch.from_numpy(f['right']).cuda()
    out = eng.forward(left, x_right=right, kk=f['K'], kind=L_.IN_KPS_STEREO, want_x=True, kernel='cluster')
    assert np.abs(out['x'].cpu().numpy() - f['pairs_x']).max() < 6e-6
    ok, worst = O.close(out['raw'].cpu().numpy(), f['pairs_raw'])
    assert ok, worst
    eng.close()
    # MC-dropout masks and the in-kernel RNG are the same function of (seed, site, row, col) in both kernels
    sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
    eng = engine.LocoEngine(sd)
    B = 77
    masks = (np.random.RandomState(3).uniform(size=(2, B, 1024)) >= 0.2).astype(np.uint8)
    x = synthetic.make_inputs(B, 34, seed=31)
    ref = O.loco_model_forward(sd, x, drop_masks=(masks[0], masks[1]), p_dropout=0.2)
    out = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_mask=torch.from_numpy(masks).cuda(), kernel='cluster')
    ok, worst = O.close(out['raw'].cpu().numpy(), ref)
    assert ok, worst
    a = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_seed=9, kernel='cluster')['raw']
    b = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_seed=9, kernel='tile')['raw']
    assert torch.allclose(a, b, rtol=2e-5, atol=2e-5)
    eng.close()
    # MonolocoModel (heads only at the end) through the cluster kernel
    g = np.load(os.path.join(GOLDEN, 'ref_fwd_monoloco_l1024_o9.npz'))
    eng = engine.LocoEngine(synthetic.make_state_dict('monoloco', 34, 9, 1024, 3, 3))
    out = eng.forward(torch.from_numpy(g['x']).cuda(), kernel='cluster')
    ok, worst = O.close(out['raw'].cpu().numpy(), g['out'])
    assert ok, worst
    eng.close()


# ------------------------------------------------------------------------------------------------ whole-grid latency kernel
@pytest.mark.parametrize('B', [1, 2, 7, 16, 17, 31, 32, 33, 64, 75])
def test_wide_kernel_batches_vs_oracle(mono1024, B):
    """Whole-grid kernel (forward_wide.cu): every layer split by output columns over L/8 CTAs, grid barrier + TMA
    exchange per layer; both row-slot instantiations (<= 16, <= 32), repeated launches (monotonic barrier counter)."""
    O, synthetic, engine, L_ = _mods()
    sd, eng = mono1024
    kps = synthetic.make_keypoints(B, seed=140 + B)
    x = O.preprocess_monoloco(kps, synthetic.KITTI_K)
    ref = O.loco_model_forward(sd, x)
    for rep in range(3):
        out = eng.forward(torch.from_numpy(kps).cuda(), kk=synthetic.KITTI_K, kind=L_.IN_KPS, want_x=True, want_xyzc=True,
                          kernel='wide')
        assert np.abs(out['x'].cpu().numpy() - x).max() < 6e-6
        ok, worst = O.close(out['raw'].cpu().numpy(), ref)
        assert ok, (rep, worst)
        _check_dec(O, out['dec'].cpu().numpy(), O.extract_outputs(ref), False, raw_ref=ref)
    tile = eng.forward(torch.from_numpy(kps).cuda(), kk=synthetic.KITTI_K, kind=L_.IN_KPS, want_xyzc=True, kernel='tile')
    assert torch.allclose(out['raw'], tile['raw'], rtol=2e-5, atol=2e-5) and torch.allclose(out['xyzc'], tile['xyzc'], rtol=1e-5, atol=1e-5)
    if 16 < B <= 64:   # selected by default for 17 .. 64 rows (one launch per 32-row tile); <= 16 rows: the wide2 kernel
        auto = eng.forward(torch.from_numpy(kps).cuda(), kk=synthetic.KITTI_K, kind=L_.IN_KPS)
        assert torch.equal(auto['raw'], out['raw'])


def test_wide_kernel_stereo_dropout_and_legacy_models():
    O, synthetic, engine, L_ = _mods()
    f = np.load(os.path.join(GOLDEN, 'ref_loco_stereo.npz'))
    sd = synthetic.make_state_dict('loco', 68, 10, 1024, 3, 2)
    eng = engine.LocoEngine(sd)
    left, right = torch.from_numpy(f['left'][:5]).cuda(), torch.from_numpy(f['right']).cuda()   # 5 x 9 = 45 pairs: two tiles
    out = eng.forward(left, x_right=right, kk=f['K'], kind=L_.IN_KPS_STEREO, want_x=True, kernel='wide')
    n = out['raw'].shape[0]
    assert n == 5 * f['right'].shape[0]
    assert np.abs(out['x'].cpu().numpy() - f['pairs_x'][:n]).max() < 6e-6
    ok, worst = O.close(out['raw'].cpu().numpy(), f['pairs_raw'][:n])
    assert ok, worst
    eng.close()
    # MC-dropout: explicit masks vs the oracle, in-kernel RNG vs the tile kernel
    sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
    eng = engine.LocoEngine(sd)
    B = 43
    masks = (np.random.RandomState(3).uniform(size=(2, B, 1024)) >= 0.2).astype(np.uint8)
    x = synthetic.make_inputs(B, 34, seed=31)
    ref = O.loco_model_forward(sd, x, drop_masks=(masks[0], masks[1]), p_dropout=0.2)
    out = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_mask=torch.from_numpy(masks).cuda(), kernel='wide')
    ok, worst = O.close(out['raw'].cpu().numpy(), ref)
    assert ok, worst
    a = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_seed=9, kernel='wide')['raw']
    b = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_seed=9, kernel='tile')['raw']
    assert torch.allclose(a, b, rtol=2e-5, atol=2e-5)
    eng.close()
    # legacy MonolocoModel widths: L = 1024 (128 CTAs) and L = 256 (32 CTAs)
    for name, args in (('ref_fwd_monoloco_l1024_o9.npz', ('monoloco', 34, 9, 1024, 3, 3)),):
        g = np.load(os.path.join(GOLDEN, name))
        eng = engine.LocoEngine(synthetic.make_state_dict(*args))
        out = eng.forward(torch.from_numpy(g['x'][:32]).cuda(), kernel='wide')
        ok, worst = O.close(out['raw'].cpu().numpy(), g['out'][:32])
        assert ok, worst
        eng.close()
    sd = synthetic.make_state_dict('monoloco', 34, 2, 256, 3, 5)
    eng = engine.LocoEngine(sd)
    x = synthetic.make_inputs(9, 34, seed=2)
    out = eng.forward(torch.from_numpy(x).cuda(), kernel='wide')
    ok, worst = O.close(out['raw'].cpu().numpy(), O.monoloco_model_forward(sd, x))
    assert ok, worst
    eng.close()


# ------------------------------------------------------------------------------------------------ second-generation latency kernel
@pytest.mark.parametrize('B', [1, 2, 7, 15, 16])
def test_wide2_kernel_batches_vs_oracle(mono1024, B):
    """Latency kernel, second generation (forward_wide2.cu): 32 clusters x 4 CTAs, K x N split, partial sums through
    distributed shared memory, (value, epoch) pair exchange; repeated launches walk the three rotating buffers and the
    monotonic epochs."""
    O, synthetic, engine, L_ = _mods()
    sd, eng = mono1024
    kps = synthetic.make_keypoints(B, seed=340 + B)
    x = O.preprocess_monoloco(kps, synthetic.KITTI_K)
    ref = O.loco_model_forward(sd, x)
    for rep in range(5):
        out = eng.forward(torch.from_numpy(kps).cuda(), kk=synthetic.KITTI_K, kind=L_.IN_KPS, want_x=True, want_xyzc=True,
                          kernel='wide2')
        assert np.abs(out['x'].cpu().numpy() - x).max() < 6e-6
        ok, worst = O.close(out['raw'].cpu().numpy(), ref)
        assert ok, (rep, worst)
        _check_dec(O, out['dec'].cpu().numpy(), O.extract_outputs(ref), False, raw_ref=ref)
    tile = eng.forward(torch.from_numpy(kps).cuda(), kk=synthetic.KITTI_K, kind=L_.IN_KPS, want_xyzc=True, kernel='tile')
    assert torch.allclose(out['raw'], tile['raw'], rtol=2e-5, atol=2e-5) and torch.allclose(out['xyzc'], tile['xyzc'], rtol=1e-5, atol=1e-5)
    auto = eng.forward(torch.from_numpy(kps).cuda(), kk=synthetic.KITTI_K, kind=L_.IN_KPS)   # the default pick up to 16 rows
    assert torch.equal(auto['raw'], out['raw']) and eng.last_kernel()[0] == 4


def test_wide2_kernel_stereo_dropout_and_legacy_models():
    O, synthetic, engine, L_ = _mods()
    f = np.load(os.path.join(GOLDEN, 'ref_loco_stereo.npz'))
    sd = synthetic.make_state_dict('loco', 68, 10, 1024, 3, 2)
    eng = engine.LocoEngine(sd)
    left, right = torch.from_numpy(f['left'][:1]).cuda(), torch.from_numpy(f['right']).cuda()   # 1 x 9 pairs
    out = eng.forward(left, x_right=right, kk=f['K'], kind=L_.IN_KPS_STEREO, want_x=True, kernel='wide2')
    n = out['raw'].shape[0]
    assert n == f['right'].shape[0] and np.abs(out['x'].cpu().numpy() - f['pairs_x'][:n]).max() < 6e-6
    ok, worst = O.close(out['raw'].cpu().numpy(), f['pairs_raw'][:n])
    assert ok, worst
    eng.close()
    sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
    eng = engine.LocoEngine(sd)
    B = 13
    masks = (np.random.RandomState(3).uniform(size=(2, B, 1024)) >= 0.2).astype(np.uint8)
    x = synthetic.make_inputs(B, 34, seed=31)
    ref = O.loco_model_forward(sd, x, drop_masks=(masks[0], masks[1]), p_dropout=0.2)
    out = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_mask=torch.from_numpy(masks).cuda(), kernel='wide2')
    ok, worst = O.close(out['raw'].cpu().numpy(), ref)
    assert ok, worst
    a = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_seed=9, kernel='wide2')['raw']
    b = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_seed=9, kernel='tile')['raw']
    assert torch.allclose(a, b, rtol=2e-5, atol=2e-5)
    eng.close()
    g = np.load(os.path.join(GOLDEN, 'ref_fwd_monoloco_l1024_o9.npz'))
    eng = engine.LocoEngine(synthetic.make_state_dict('monoloco', 34, 9, 1024, 3, 3))
    out = eng.forward(torch.from_numpy(g['x'][:16]).cuda(), kernel='wide2')
    ok, worst = O.close(out['raw'].cpu().numpy(), g['out'][:16])
    assert ok, worst
    eng.close()
    sd = synthetic.make_state_dict('monoloco', 34, 2, 256, 3, 5)   # legacy width: 8 clusters, K slices of 64
    eng = engine.LocoEngine(sd)
    x = synthetic.make_inputs(9, 34, seed=2)
    out = eng.forward(torch.from_numpy(x).cuda(), kernel='wide2')
    ok, worst = O.close(out['raw'].cpu().numpy(), O.monoloco_model_forward(sd, x))
    assert ok, worst
    eng.close()


# ------------------------------------------------------------------------------------------------ tensor-core kernel
@pytest.mark.parametrize('B', [1, 100, 128, 129, 1000, 4224, 4500])
def test_tc_kernel_batches_vs_oracle(mono1024, B):
    """Tensor-core kernel (forward_tc.cu: 3xTF32 on tcgen05, persistent clusters over 128-row tiles), forced on for every
    size: ragged last tile, exactly one wave, more tiles than co-resident clusters; raw keypoints in, decoded rows out."""
    O, synthetic, engine, L_ = _mods()
    sd, eng = mono1024
    kps = synthetic.make_keypoints(B, seed=240 + B)
    x = O.preprocess_monoloco(kps, synthetic.KITTI_K)
    ref = O.loco_model_forward(sd, x)
    for rep in range(2):   # second launch: the cluster workspace slots are reused
        out = eng.forward(torch.from_numpy(kps).cuda(), kk=synthetic.KITTI_K, kind=L_.IN_KPS, want_x=True, want_xyzc=True,
                          kernel='tc')
        assert np.abs(out['x'].cpu().numpy() - x).max() < 6e-6
        ok, worst = O.close(out['raw'].cpu().numpy(), ref)
        assert ok, (rep, worst)
        _check_dec(O, out['dec'].cpu().numpy(), O.extract_outputs(ref), False, raw_ref=ref)
    tile = eng.forward(torch.from_numpy(kps).cuda(), kk=synthetic.KITTI_K, kind=L_.IN_KPS, want_xyzc=True, kernel='tile')
    assert torch.allclose(out['raw'], tile['raw'], rtol=2e-5, atol=2e-5) and torch.allclose(out['xyzc'], tile['xyzc'], rtol=1e-5, atol=1e-5)
    if B >= 300:   # the default pick for batches beyond one wave of FFMA clusters
        auto = eng.forward(torch.from_numpy(kps).cuda(), kk=synthetic.KITTI_K, kind=L_.IN_KPS)
        assert torch.equal(auto['raw'], out['raw'])


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN, 'ref_fwd_*.npz'))))
def test_tc_kernel_golden(path):
    """Every live-reference fixture whose width the tensor-core kernel covers (L % 256 == 0), forced through it."""
    O, synthetic, engine, L_ = _mods()
    f = np.load(path)
    isz, osz, L, st, seed = [int(v) for v in f['cfg'][:5]]
    if L % 256:
        pytest.skip("L = %d runs on the FFMA kernels" % L)
    eng = engine.LocoEngine(synthetic.make_state_dict(str(f['kind']), isz, osz, L, st, seed))
    out = eng.forward(torch.from_numpy(f['x']).cuda(), kernel='tc')
    ok, worst = O.close(out['raw'].cpu().numpy(), f['out'])
    assert ok, (path, worst)
    if 'dec_xyzd' in f.files:
        ref = {'xyzd': f['dec_xyzd'], 'bi': f['dec_bi'], 'yaw': (f['dec_yaw_pred'], f['dec_yaw_orig'])}
        if 'dec_aux' in f.files:
            ref['aux'] = f['dec_aux']
        _check_dec(O, out['dec'].cpu().numpy(), ref, 'dec_aux' in f.files, raw_ref=f['out'])
    eng.close()


def test_tc_kernel_stereo_dropout_and_legacy_models():
    O, synthetic, engine, L_ = _mods()
    f = np.load(os.path.join(GOLDEN, 'ref_loco_stereo.npz'))
    sd = synthetic.make_state_dict('loco', 68, 10, 1024, 3, 2)
    eng = engine.LocoEngine(sd)
    left, right = torch.from_numpy(f['left']).cuda(), torch.from_numpy(f['right']).cuda()
    out = eng.forward(left, x_right=right, kk=f['K'], kind=L_.IN_KPS_STEREO, want_x=True, kernel='tc')
    assert np.abs(out['x'].cpu().numpy() - f['pairs_x']).max() < 6e-6
    ok, worst = O.close(out['raw'].cpu().numpy(), f['pairs_raw'])
    assert ok, worst
    eng.close()
    # MC-dropout: explicit masks vs the oracle; the in-kernel RNG is the same function of (seed, site, row, col) everywhere
    sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)
    eng = engine.LocoEngine(sd)
    B = 333
    masks = (np.random.RandomState(3).uniform(size=(2, B, 1024)) >= 0.2).astype(np.uint8)
    x = synthetic.make_inputs(B, 34, seed=31)
    ref = O.loco_model_forward(sd, x, drop_masks=(masks[0], masks[1]), p_dropout=0.2)
    out = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_mask=torch.from_numpy(masks).cuda(), kernel='tc')
    ok, worst = O.close(out['raw'].cpu().numpy(), ref)
    assert ok, worst
    a = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_seed=9, kernel='tc')['raw']
    b = eng.forward(torch.from_numpy(x).cuda(), dropout=True, drop_seed=9, kernel='tile')['raw']
    assert torch.allclose(a, b, rtol=2e-5, atol=2e-5)
    # zero-centred legacy pre-process (net.py:96)
    kps = synthetic.make_keypoints(200, seed=8)
    xz = O.preprocess_monoloco(kps, synthetic.KITTI_K, zero_center=True)
    out = eng.forward(torch.from_numpy(kps).cuda(), kk=synthetic.KITTI_K, kind=L_.IN_KPS, zero_center=True, want_x=True, kernel='tc')
    assert np.abs(out['x'].cpu().numpy() - xz).max() < 6e-6
    ok, worst = O.close(out['raw'].cpu().numpy(), O.loco_model_forward(sd, xz))
    assert ok, worst
    eng.close()
    # legacy MonolocoModel: L = 256 is a one-CTA "cluster", the head is the only output layer
    sd = synthetic.make_state_dict('monoloco', 34, 2, 256, 3, 5)
    eng = engine.LocoEngine(sd)
    x = synthetic.make_inputs(300, 34, seed=2)
    out = eng.forward(torch.from_numpy(x).cuda(), kernel='tc')
    ok, worst = O.close(out['raw'].cpu().numpy(), O.monoloco_model_forward(sd, x))
    assert ok, worst
    eng.close()


def test_tc_kernel_many_waves_properties(mono1024):
    """40 000 rows = 313 tiles over ~33 persistent clusters: row independence (replicated inputs give bit-identical rows
    wherever they land) plus an oracle spot check."""
    O, synthetic, engine, L_ = _mods()
    sd, eng = mono1024
    base = synthetic.make_inputs(2003, 34, seed=78)
    x = torch.from_numpy(np.tile(base, (20, 1))[:40000]).cuda()
    raw = eng.forward(x, kernel='tc')['raw']
    first = raw[:2003]
    for r in range(1, 19):
        assert torch.equal(raw[r * 2003:(r + 1) * 2003], first)
    idx = np.random.RandomState(1).choice(2003, 256, replace=False)
    ok, worst = O.close(first.cpu().numpy()[idx], O.loco_model_forward(sd, base[idx]))
    assert ok, worst


@pytest.mark.parametrize('L,kind', [(2048, 'loco'), (300, 'loco'), (1500, 'loco'), (200, 'monoloco')])
def test_any_hidden_size(L, kind):
    """`--hidden_size` is free in the reference (run.py:101,122; hyp_tuning.py:52 uses 2048): widths beyond 1024 run on the
    tensor-core kernel (multiples of 256 up to 2048), every other width is zero-padded by the packer."""
    O, synthetic, engine, L_ = _mods()
    sd = synthetic.make_state_dict(kind, 34, 9, L, 2, 4)
    eng = engine.LocoEngine(sd)
    for B in (7, 70, 600):
        x = synthetic.make_inputs(B, 34, seed=B)
        out = eng.forward(torch.from_numpy(x).cuda())
        ok, worst = O.close(out['raw'].cpu().numpy(), O.model_forward(sd, x))
        assert ok, (L, B, worst)
    eng.close()
