"""CPU: host-side API surface against fixtures produced by the real reference (tests/golden/ref_api.json):
checkpoint ABI, preprocess_pifpaf, load_calibration, IoU matching; plus the C-ABI library exports."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def api():
    with open(os.path.join(GOLDEN, 'ref_api.json')) as f:
        return json.load(f)


def test_checkpoint_abi(api):
    """state_dict keys and shapes are the reference's (SURVEY.md §8b 'Checkpoint ABI')."""
    from monoloco_b200.network.architectures import LocoModel, MonolocoModel
    models = {'loco_34_9_1024': LocoModel(34, 9, 1024), 'loco_68_10_1024': LocoModel(68, 10, 1024),
              'loco_34_9_256_s2': LocoModel(34, 9, 256, num_stage=2), 'monoloco_34_2_256': MonolocoModel(34, 2, 256),
              'monoloco_34_9_1024': MonolocoModel(34, 9, 1024)}
    for name, m in models.items():
        got = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert list(got.keys()) == list(api['abi'][name].keys()), name
        assert got == api['abi'][name], name


def test_state_dict_roundtrip_through_packer():
    """A reference-format checkpoint loads into the mirror module and packs to the same blob as the raw dict."""
    import torch
    from monoloco_b200 import synthetic, packing
    from monoloco_b200.network.architectures import LocoModel
    sd = synthetic.make_state_dict('loco', 34, 9, 256, 2, 4)
    m = LocoModel(34, 9, 256, num_stage=2)
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    a = packing.pack_state_dict(sd)
    b = packing.pack_state_dict(m.state_dict())
    assert np.array_equal(a.blob, b.blob) and a.ops == b.ops and a.desc == b.desc


def test_packing_layout():
    from monoloco_b200 import synthetic, packing, _lib as L_
    sd = synthetic.make_state_dict('loco', 68, 10, 256, 3, 9)
    pm = packing.pack_state_dict(sd)
    assert pm.desc['input_size'] == 68 and pm.desc['output_size'] == 10 and pm.desc['decode_kind'] == L_.DECODE_LOCO
    gemms = [o for o in pm.ops if o['type'] == L_.OP_GEMM]
    heads = [o for o in pm.ops if o['type'] == L_.OP_HEAD]
    assert len(gemms) == 1 + 2 * 3 + 2 and len(heads) == 2
    assert gemms[0]['Kpad'] == 72 and gemms[0]['K'] == 68
    o = gemms[1]
    wt = pm.blob[o['w_off']:o['w_off'] + 256 * 256].reshape(256, 256)
    assert np.array_equal(wt, sd['linear_stages.0.w1.weight'].T)
    s = pm.blob[o['scale_off']:o['scale_off'] + 256]
    ref = sd['linear_stages.0.batch_norm1.weight'] / np.sqrt(sd['linear_stages.0.batch_norm1.running_var'] + 1e-5)
    assert np.allclose(s, ref, rtol=1e-6)
    assert all(o['w_off'] % 32 == 0 for o in pm.ops)
    assert packing.flops_per_detection(synthetic.make_state_dict('loco', 34, 9, 1024, 3, 0)) == 16865280
    assert packing.flops_per_detection(synthetic.make_state_dict('monoloco', 34, 9, 1024, 3, 0)) == 12670976


def test_preprocess_pifpaf(api):
    from monoloco_b200.network import preprocess_pifpaf
    # the reference's own test input (tests/002282.png.pifpaf.json), committed as a data fixture by gen_golden.py
    src = os.path.join(GOLDEN, 'pifpaf_002282.json')
    with open(src) as f:
        ann = json.load(f)
    boxes, kps = preprocess_pifpaf(json.loads(json.dumps(ann)), im_size=(1238, 374))
    assert np.allclose(np.array(boxes), np.array(api['boxes']), rtol=1e-12) and kps == api['keypoints']
    boxes2, _ = preprocess_pifpaf(json.loads(json.dumps(ann)), im_size=None, enlarge_boxes=False, min_conf=0.3)
    assert np.allclose(np.array(boxes2), np.array(api['boxes_noenlarge_conf03']), rtol=1e-12)


def test_load_calibration(api):
    from monoloco_b200.network import load_calibration
    assert np.allclose(load_calibration('kitti', (1238, 374)), api['calib']['kitti_1238x374'])
    assert np.allclose(load_calibration('nuscenes', (800, 450)), api['calib']['nuscenes_800x450'])
    assert np.allclose(load_calibration('custom', (1920, 1080), focal_length=5.7), api['calib']['custom_1920x1080'])


def test_iou_matching(api):
    from monoloco_b200.utils import get_iou_matches, reorder_matches, get_iou_matrix
    boxes, gt = api['boxes'], api['dic_gt']['boxes']
    matches = get_iou_matches(boxes, gt, iou_min=0.3)
    assert sorted(m[0] for m in matches) == [2, 3, 4, 5, 6]
    ordered = reorder_matches(matches, boxes, mode='left_right')
    assert [b for b in api['post']['boxes_gt']] == [gt[j] for _, j in ordered]
    assert get_iou_matrix(boxes, gt).shape == (16, 6)  # reference tests/test_utils.py:11-15
    assert get_iou_matches([], gt) == []


def test_pixel_to_camera_linearity():
    """reference tests/test_utils.py:18-25 on the mirror helper."""
    import torch
    from monoloco_b200.utils import pixel_to_camera
    kk = [[718.3351, 0., 600.3891], [0., 718.3351, 181.5122], [0., 0., 1.]]
    uv = torch.tensor([[100., 50.], [700., 300.]])
    assert torch.allclose(pixel_to_camera(uv, kk, 1) * 7.5, pixel_to_camera(uv, kk, 7.5), rtol=1e-6)


def test_c_abi_exports():
    """The shared library loads and exports every function include/monoloco_b200.h declares (no compute calls)."""
    from monoloco_b200 import build, _lib
    path = build.build()
    lib = ctypes.CDLL(path)
    with open(os.path.join(ROOT, 'include', 'monoloco_b200.h')) as f:
        header = f.read()
    declared = set(re.findall(r'\b(mlb_[a-z0-9_]+)\s*\(', header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTS)
    lib.mlb_abi_version.restype = ctypes.c_int
    assert lib.mlb_abi_version() == _lib.MLB_ABI_VERSION


def test_product_path_has_no_cpu_fallback():
    """Without a GPU every product entry point raises instead of silently computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from monoloco_b200 import synthetic, engine
    from monoloco_b200.network import Loco
    from monoloco_b200.network.architectures import LocoModel
    with pytest.raises(RuntimeError):
        engine.LocoEngine(synthetic.make_state_dict('loco', 34, 9, 128, 1, 0))
    with pytest.raises(RuntimeError):
        Loco(model=LocoModel(34, 9, 128), mode='mono')
    with pytest.raises(RuntimeError):
        LocoModel(34, 9, 128).eval()(torch.zeros(2, 34))


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under monoloco_b200/ may import it."""
    pkg = os.path.join(ROOT, 'monoloco_b200')
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith('.py'):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), os.path.join(dp, fn)


def test_activity_helpers(api):
    """Loco.social_distance / raising_hand building blocks vs outputs of the reference's activity.py (fixture)."""
    from monoloco_b200 import activity as M
    with open(os.path.join(GOLDEN, 'ref_activity.json')) as f:
        ref = json.load(f)
    assert [M.is_raising_hand(k) for k in api['keypoints']] == ref['raising']
    n = len(ref['centers'])
    prob = [M.social_interactions(i, ref['centers'], ref['angles'], ref['dds'], stds=ref['stds']) for i in range(n)]
    det = [M.social_interactions(i, ref['centers'], ref['angles'], ref['dds'], stds=ref['stds'], n_samples=1) for i in range(n)]
    assert prob == ref['prob'] and det == ref['det']


def test_model_copy_and_pickle_drop_the_engine_handle():
    """copy.deepcopy(model) / torch.save(model) after an eval forward (best-model snapshots, EMA) must not trip
    over the ctypes engine handle; the copy rebuilds its own engine lazily."""
    import copy
    import ctypes
    import io
    import torch
    from monoloco_b200.network.architectures import LocoModel
    m = LocoModel(34, 9, 128, num_stage=1)
    object.__setattr__(m, '_engine', ctypes.c_void_p(1234))   # what engine() stores after the first eval forward
    object.__setattr__(m, '_engine_key', ('k',))
    c = copy.deepcopy(m)
    assert c._engine is None and c._engine_key is None and '_engine' not in c.__dict__
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), c.state_dict().values()))
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    r = torch.load(buf, weights_only=False)
    assert r._engine is None and sorted(r.state_dict()) == sorted(m.state_dict())
    assert m._engine.value == 1234   # the original keeps its handle


def _run_packed_program(pm, x):
    """Host interpreter of the packed layer program (include/monoloco_b200.h `mlb_op`), fp64: what every fused forward
    kernel executes, minus dropout.  Used only to check that the packer's layout / BN folding / width padding keep the
    reference's arithmetic; the kernels themselves are checked on the GPU."""
    from monoloco_b200 import _lib as L_
    L = pm.desc['linear_size']
    blob = pm.blob.astype(np.float64)
    cur, res = None, None
    out = np.zeros((x.shape[0], pm.desc['output_size']))
    for o in pm.ops:
        if o['type'] == L_.OP_GEMM:
            src = x if o['flags'] & L_.F_IN_XIN else cur
            wt = blob[o['w_off']:o['w_off'] + o['Kpad'] * L].reshape(o['Kpad'], L)
            y = src @ wt[:src.shape[1]]
            y = y * blob[o['scale_off']:o['scale_off'] + L] + blob[o['shift_off']:o['shift_off'] + L]
            if o['flags'] & L_.F_RELU:
                y = np.maximum(y, 0)
            if o['flags'] & L_.F_ADD_RES:
                y = y + res
            if o['flags'] & L_.F_SAVE_RES:
                res = y
            cur = y
        else:
            w = blob[o['w_off']:o['w_off'] + o['N'] * L].reshape(o['N'], L)
            out[:, o['out_col']:o['out_col'] + o['N']] = cur @ w.T + blob[o['shift_off']:o['shift_off'] + o['N']]
    return out, cur


@pytest.mark.parametrize('kind,width,n_out,stages', [('loco', 256, 9, 3), ('loco', 300, 10, 2), ('loco', 1100, 9, 1),
                                                     ('monoloco', 200, 9, 3), ('monoloco', 128, 2, 0)])
def test_packed_program_matches_oracle_any_width(kind, width, n_out, stages):
    """The packed program reproduces the reference network (oracle restatement of architectures.py:8-46, 111-133) for
    hidden widths that need padding too: padded units carry exact zeros through every layer."""
    from monoloco_b200 import synthetic, packing
    from oracle import loco_oracle as O
    n_in = 68 if kind == 'loco' else 34
    sd = synthetic.make_state_dict(kind, n_in, n_out, width, stages, seed=width)
    pm = packing.pack_state_dict(sd)
    Lp = packing.padded_width(width)
    assert pm.desc['linear_size'] == Lp and Lp % 128 == 0 and Lp >= width
    x = np.random.RandomState(1).randn(37, n_in).astype(np.float32)
    got, hidden = _run_packed_program(pm, x.astype(np.float64))
    ref = O.model_forward({k: np.asarray(v, dtype=np.float64) for k, v in sd.items()}, x.astype(np.float64))
    assert got.shape == ref.shape
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-5), np.abs(got - ref).max()
    assert not hidden[:, width:].any()          # padded hidden units are exactly zero


def test_packed_program_matches_live_reference_wide():
    """Same interpreter against the live-reference fixtures of the padded widths (tests/golden/ref_wide_*.npz)."""
    import glob
    from monoloco_b200 import synthetic, packing
    from oracle import loco_oracle as O
    paths = sorted(glob.glob(os.path.join(GOLDEN, 'ref_wide_*.npz')))
    assert len(paths) >= 5
    for path in paths:
        f = np.load(path)
        isz, osz, L, st, seed = [int(v) for v in f['cfg'][:5]]
        pm = packing.pack_state_dict(synthetic.make_state_dict(str(f['kind']), isz, osz, L, st, seed))
        got, _ = _run_packed_program(pm, f['x'].astype(np.float64))
        ok, worst = O.close(got, f['out'])
        assert ok, (path, worst)
