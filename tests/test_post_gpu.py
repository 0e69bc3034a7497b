"""GPU: the steps right after the network on the device (BASELINE configs[2]):
`mlb_post_process` over a batch of images against the live reference (tests/golden/ref_api.json) and against the host
re-statement image by image, `mlb_kitti_rows` against the byte-identical host writer, and the monstereo arg-max filter at
its stated size (64 x 64 = 4096 pairs) against the oracle."""
import json
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

EXACT = ('gt', 'boxes', 'uv_kps', 'uv_centers', 'uv_shoulders', 'uv_heads', 'dds_real', 'boxes_gt', 'dds_pred', 'stds_ale',
         'stds_epi', 'angles', 'angles_egocentric', 'aux')
CLOSE = ('confs', 'xyz_pred', 'xyz_real')


def _same(post, ref, what, rtol=2e-5):
    assert sorted(post.keys()) == sorted(ref.keys()), what
    for k in EXACT:
        if k in ref:
            assert post[k] == ref[k], (what, k)
    for k in CLOSE:
        if k in ref:
            assert len(post[k]) == len(ref[k]), (what, k)
            if len(ref[k]):
                assert np.allclose(np.array(post[k], dtype=np.float64), np.array(ref[k], dtype=np.float64), rtol=rtol,
                                   atol=2e-5), (what, k)


def _synthetic_image(rng, m, g, with_gt=True):
    """m detections (boxes around the key points, confidences with a deliberate tie) and g ground truths of which some
    overlap detections, some compete for the same detection and some match nothing."""
    from monoloco_b200 import synthetic
    kps = synthetic.make_keypoints(m, seed=int(rng.randint(1 << 30)))
    boxes = []
    for j in range(m):
        u, v = kps[j, 0], kps[j, 1]
        boxes.append([float(u.min()), float(v.min()), float(u.max()), float(v.max()), float(rng.uniform(0.2, 1.0))])
    if m > 3:
        boxes[2][4] = boxes[1][4]  # equal confidences: stable order
    dic_in = {'d': torch.from_numpy(rng.uniform(3, 60, (m, 1)).astype(np.float32)),
              'bi': torch.from_numpy(rng.uniform(0.1, 5, (m, 1)).astype(np.float32)),
              'epi': [0.] * m if m % 2 else torch.from_numpy(rng.uniform(0, 1, m).astype(np.float32)),
              'yaw': (torch.from_numpy(rng.uniform(-3, 3, (m, 1)).astype(np.float32)),
                      torch.from_numpy(rng.uniform(-3, 3, (m, 1)).astype(np.float32)))}
    if m % 3 == 0:
        dic_in['aux'] = torch.from_numpy(rng.uniform(0, 1, (m, 1)).astype(np.float32))
    dic_gt = None
    if with_gt and g > 0:
        gtb, ys = [], []
        for k in range(g):
            if k < m and rng.uniform() < 0.7:
                b = boxes[int(rng.randint(m))]
                jit = rng.uniform(-0.15, 0.15, 4) * (b[2] - b[0] + b[3] - b[1]) / 2
                gtb.append([b[0] + jit[0], b[1] + jit[1], b[2] + jit[2], b[3] + jit[3]])
            else:
                x0, y0 = rng.uniform(0, 1200), rng.uniform(0, 300)
                gtb.append([x0, y0, x0 + rng.uniform(10, 80), y0 + rng.uniform(20, 150)])
            ys.append([0., 0., 0., float(rng.uniform(3, 60))])
        dic_gt = {'boxes': gtb, 'ys': ys}
    return dic_in, boxes, kps.tolist(), dic_gt


def test_post_process_batch_vs_live_reference():
    """One real image (pifpaf fixture) through the device kernel == the reference's own post_process output."""
    from monoloco_b200 import synthetic
    from monoloco_b200.network import Loco, preprocess_pifpaf
    from monoloco_b200.network.post import post_process_batch
    from monoloco_b200.network.architectures import LocoModel
    f = np.load(os.path.join(GOLDEN, 'ref_loco_mono_pifpaf.npz'))
    sd = synthetic.make_state_dict('loco', 34, 9, 1024, 3, 1)
    m = LocoModel(34, 9, 1024, num_stage=3)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    net = Loco(model=m, mode='mono', device=torch.device('cuda'))
    with open(os.path.join(GOLDEN, 'pifpaf_002282.json')) as fh:
        boxes, keypoints = preprocess_pifpaf(json.load(fh), im_size=(1238, 374))
    kk = f['K'].tolist()
    dic = net.forward(keypoints, kk)
    with open(os.path.join(GOLDEN, 'ref_api.json')) as fh:
        api = json.load(fh)
    posts = post_process_batch([(dic, boxes, keypoints, kk, api['dic_gt']), (dic, boxes, keypoints, kk, None),
                                (None, [], [], kk, None)])
    assert len(posts[2]) == 0
    for post, key in ((posts[0], 'post'), (posts[1], 'post_nogt')):
        ref = api[key]
        assert sorted(post.keys()) == sorted(ref.keys())
        for k in ('gt', 'uv_centers', 'uv_heads', 'uv_shoulders', 'boxes'):
            assert post[k] == ref[k], k
        for k in ('confs', 'dds_pred', 'stds_ale', 'xyz_pred', 'angles', 'angles_egocentric'):
            assert np.allclose(np.array(post[k]), np.array(ref[k]), rtol=3e-5, atol=2e-4), k
    assert posts[0]['dds_real'] == api['post']['dds_real'] and posts[0]['boxes_gt'] == api['post']['boxes_gt']
    assert np.allclose(np.array(posts[0]['xyz_real']), np.array(api['post']['xyz_real']), rtol=1e-5)


@pytest.mark.parametrize('reorder', [True, False])
def test_post_process_batch_vs_host_many_images(reorder):
    """200 synthetic images (0..40 detections, 0..12 ground truths, ties, competing matches): indices, order and every
    list entry equal the per-image host implementation (itself pinned to the live reference in test_loco_gpu.py)."""
    from monoloco_b200 import synthetic
    from monoloco_b200.network import Loco
    from monoloco_b200.network.post import post_process_batch
    rng = np.random.RandomState(11)
    kk = synthetic.KITTI_K
    items = []
    for i in range(200):
        m = int(rng.choice([1, 2, 3, 5, 8, 13, 21, 40]))
        g = int(rng.choice([0, 1, 2, 4, 7, 12]))
        dic_in, boxes, kps, dic_gt = _synthetic_image(rng, m, g)
        kki = [[kk[0][0] * (1 + 0.01 * (i % 3)), 0., kk[0][2]], [0., kk[1][1], kk[1][2]], [0., 0., 1.]]
        items.append((dic_in, boxes, kps, kki, dic_gt))
    got = post_process_batch(items, iou_min=0.3, reorder=reorder)
    n_matched = 0
    for i, it in enumerate(items):
        dic_in = dict(it[0])
        dic_in['xyz_c'] = None
        ref = Loco.post_process(dic_in, it[1], it[2], it[3], dic_gt=it[4], iou_min=0.3, reorder=reorder)
        _same(got[i], ref, i)
        n_matched += sum(ref['gt'])
    assert n_matched > 100  # the cases do exercise the matcher


def test_kitti_rows_device_byte_identical(tmp_path):
    """`save_txts` numbers from the device == the host writer (itself byte-identical to the live reference)."""
    from monoloco_b200 import synthetic, engine, _lib as L_
    from monoloco_b200.network.post import kitti_rows_device
    from monoloco_b200.utils.kitti import kitti_rows
    for net, (isz, osz) in (('monoloco_pp', (34, 9)), ('monstereo', (68, 10))):
        sd = synthetic.make_state_dict('loco', isz, osz, 256, 2, 3)
        eng = engine.LocoEngine(sd)
        n = 37
        x = torch.from_numpy(synthetic.make_inputs(n, isz, seed=4)).cuda()
        out = eng.forward(x, kind=L_.IN_X)
        raw, dec = out['raw'], out['dec']
        rng = np.random.RandomState(5)
        boxes = [[float(v) for v in np.round(rng.uniform(0, 1200, 4), 2)] + [float(rng.uniform(0.1, 1.0))] for _ in range(n)]
        epi = torch.from_numpy(rng.uniform(0, 1, n).astype(np.float32))
        rows = kitti_rows_device(boxes, raw, dec, epi=epi, net=net)
        r, d = raw.cpu(), dec.cpu()
        outs = [d[:, 0:4], d[:, 4:5], epi, (d[:, 5:6], d[:, 6:7]), r[:, 4:5], r[:, 5:6], r[:, 6:7]]
        _, table = kitti_rows(boxes, outs, None, net=net, cat=[0.] * n)
        fmt = lambda t: ''.join(('%f ' * 15 + '\n') % tuple(row) for row in t.tolist())  # noqa: E731
        assert fmt(rows) == fmt(table)
        assert np.array_equal(rows, table)
        eng.close()


def test_stereo_64x64_pairs_filter_vs_oracle():
    """BASELINE configs[2] at its stated size: 64 left x 64 right = 4096 pair rows, arg-max filter, decode and
    xyz_from_distance against the numpy oracle."""
    from oracle import loco_oracle as O
    from monoloco_b200 import synthetic, engine, _lib as L_
    sd = synthetic.make_state_dict('loco', 68, 10, 1024, 3, 2)
    eng = engine.LocoEngine(sd)
    le, ri = synthetic.make_keypoints(64, seed=3, right=True)
    out = eng.forward(torch.from_numpy(le).cuda(), x_right=torch.from_numpy(ri).cuda(), kk=synthetic.KITTI_K,
                      kind=L_.IN_KPS_STEREO, want_xyzc=True)
    pairs, _ = O.preprocess_monstereo(le, ri, synthetic.KITTI_K)
    assert pairs.shape == (4096, 68)
    ref_raw = O.loco_model_forward(sd, pairs)
    ok, worst = O.close(out['raw'].cpu().numpy(), ref_raw)
    assert ok, worst
    # filter on the ENGINE's logits (the arg-max is exact index work on identical inputs) ...
    raw_np = out['raw'].cpu().numpy()
    _, mask = O.filter_outputs(raw_np.reshape(64, 64, 10))
    sel_raw, sel_dec, sel_idx, sel_xyzc = eng.stereo_filter(out['raw'], out['dec'], 64, 64, xyzc=out['xyzc'])
    keep = np.where(mask.reshape(-1))[0]
    assert np.array_equal(sel_idx.cpu().numpy(), keep)
    assert np.array_equal(sel_raw.cpu().numpy(), raw_np[keep])
    # ... and the winners agree with the oracle's own winners wherever the top two logits are not within the tolerance
    _, ref_mask = O.filter_outputs(ref_raw.reshape(64, 64, 10))
    srt = np.sort(ref_raw.reshape(64, 64, 10)[:, :, 9], axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > 1e-4
    assert clear.sum() >= 56
    assert np.array_equal(mask[clear], ref_mask[clear])
    # decoded fields + xyz_from_distance of the kept rows
    ref = O.extract_outputs(ref_raw[keep])
    dec = sel_dec.cpu().numpy()
    ok, worst = O.close(dec[:, 0:4], ref['xyzd'], col_scale=False)
    assert ok, worst
    ok, worst = O.close(dec[:, 4:5], ref['bi'])
    assert ok, worst
    ok, worst = O.close(dec[:, 7:8], ref['aux'])
    assert ok, worst
    cen = O.get_keypoints(le[keep // 64], 'center')
    xyc = O.pixel_to_camera(cen, synthetic.KITTI_K, 1.0)
    xyz = O.xyz_from_distance(ref_raw[keep][:, 2:3], xyc)
    ok, worst = O.close(sel_xyzc.cpu().numpy()[:, :3], xyz)
    assert ok, worst
    # host-returning variant used by Loco.forward: same rows with a single synchronisation
    h_raw, h_dec, h_xyzc = eng.stereo_filter_host(out['raw'], out['dec'], out['xyzc'], 64, 64)
    assert np.array_equal(h_raw.numpy(), sel_raw.cpu().numpy()) and np.array_equal(h_dec.numpy(), dec)
    assert np.array_equal(h_xyzc.numpy(), sel_xyzc.cpu().numpy())
    eng.close()


def test_stereo_filter_ties_nan_and_wide_right():
    """ties keep every row in row-major order; a NaN logit empties that left pose's selection (torch.max propagates
    NaN, process.py:321-326); more right poses than a warp has lanes."""
    from monoloco_b200 import synthetic, engine
    eng = engine.LocoEngine(synthetic.make_state_dict('loco', 68, 10, 128, 1, 2))
    rng = np.random.RandomState(0)
    n_left, n_right = 37, 75
    raw = rng.standard_normal((n_left * n_right, 10)).astype(np.float32)
    raw3 = raw.reshape(n_left, n_right, 10)
    raw3[3, :, 9] = 0.5                      # whole row tied
    raw3[5, [1, 40, 74], 9] = 9.0            # three-way tie across warp passes
    raw3[7, 13, 9] = np.nan                  # NaN -> nothing kept for left pose 7
    dec = rng.standard_normal((n_left * n_right, 8)).astype(np.float32)
    exp = []
    for l in range(n_left):
        v = raw3[l, :, 9]
        if np.isnan(v).any():
            continue
        exp += [l * n_right + r for r in np.where(v >= v.max())[0]]
    traw, tdec = torch.from_numpy(raw).cuda(), torch.from_numpy(dec).cuda()
    sel_raw, sel_dec, sel_idx = eng.stereo_filter(traw, tdec, n_left, n_right)
    assert sel_idx.cpu().numpy().tolist() == exp
    assert np.array_equal(sel_raw.cpu().numpy(), raw[exp], equal_nan=True)
    assert np.array_equal(sel_dec.cpu().numpy(), dec[exp])
    h_raw, h_dec, _ = eng.stereo_filter_host(traw, tdec, torch.zeros((n_left * n_right, 4), device='cuda'), n_left, n_right)
    assert np.array_equal(h_raw.numpy(), raw[exp], equal_nan=True)   # more kept rows than the first batch: second fetch
    eng.close()
