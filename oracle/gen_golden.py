"""
Generate the golden fixtures under tests/golden/ by running the REAL reference (imported from
/root/reference, read-only) in the build container.  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden.py

The reference is Python and cannot travel to the GPU box, so its outputs are committed as small
.npz fixtures; weights are NOT stored -- they are regenerated from `monoloco_b200.synthetic`
seeds (numpy RandomState, machine independent) and a checksum is stored instead.

matplotlib is absent here; the reference imports pyplot at module import (net.py:19 ->
activity.py:10, losses.py:12), so it is stubbed with MagicMock before import.
"""
import json
import math
import os
import sys
from unittest.mock import MagicMock

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
OUT = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)

for _m in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.patches', 'matplotlib.cm']:
    sys.modules[_m] = MagicMock()
sys.path.insert(0, REF)

import torch  # noqa: E402
from monoloco.network import Loco  # noqa: E402
from monoloco.network.architectures import LocoModel, MonolocoModel  # noqa: E402
from monoloco.network.process import (preprocess_monoloco, preprocess_monstereo, extract_outputs,  # noqa: E402
                                      extract_outputs_mono, cluster_outputs, filter_outputs,
                                      preprocess_pifpaf, laplace_sampling, unnormalize_bi)
from monoloco.utils import xyz_from_distance, pixel_to_camera, get_keypoints  # noqa: E402
from monoloco.train.losses import CompositeLoss, MultiTaskLoss, AutoTuneMultiTaskLoss  # noqa: E402

from monoloco_b200 import synthetic  # noqa: E402

torch.set_num_threads(8)


def sd_to_torch(sd):
    return {k: torch.from_numpy(np.array(v)) for k, v in sd.items()}


def sd_checksum(sd):
    return float(sum(float(np.asarray(v, dtype=np.float64).sum()) for k, v in sorted(sd.items())))


def build(kind, input_size, output_size, linear_size, num_stage, seed, p_dropout=0.2):
    sd = synthetic.make_state_dict(kind, input_size, output_size, linear_size, num_stage, seed)
    if kind == 'loco':
        m = LocoModel(input_size, output_size, linear_size, p_dropout, num_stage, device='cpu')
    else:
        m = MonolocoModel(input_size, output_size, linear_size, p_dropout, num_stage)
    m.load_state_dict(sd_to_torch(sd))
    m.eval()
    return m, sd


def dic_to_np(dic, prefix=''):
    out = {}
    for k, v in dic.items():
        if k == 'yaw':
            out[prefix + 'yaw_pred'] = v[0].numpy()
            out[prefix + 'yaw_orig'] = v[1].numpy()
        elif k == 'epi':
            out[prefix + 'epi'] = np.asarray(v, dtype=np.float32)
        else:
            out[prefix + k] = v.numpy()
    return out


def kat_preprocess():
    """Known-answer fixtures of the reference's own tests: stored X == preprocess(kps, K)."""
    for mode in ('mono', 'stereo'):
        d = json.load(open(os.path.join(REF, 'tests', 'sample_joints-kitti-%s.json' % mode)))
        for phase in ('train', 'val'):
            kps = np.asarray(d[phase]['kps'], dtype=np.float32)[:, 0]  # (N,3,17|34)
            xs = np.asarray(d[phase]['X'], dtype=np.float32)
            ys = np.asarray(d[phase]['Y'], dtype=np.float32)
            ks = []
            for k in d[phase]['K']:
                if k not in ks:
                    ks.append(k)
            k_rows = np.zeros((kps.shape[0], 3, 3), dtype=np.float32)
            worst = 0.0
            for i in range(kps.shape[0]):
                best = None
                for k in ks:
                    kt = torch.tensor(k)
                    if mode == 'mono':
                        x = preprocess_monoloco(torch.from_numpy(kps[i:i + 1]), kt)[0].numpy()
                    else:
                        le = preprocess_monoloco(torch.from_numpy(kps[i:i + 1, :, :17]), kt)[0]
                        ri = preprocess_monoloco(torch.from_numpy(kps[i:i + 1, :, 17:]), kt)[0]
                        x = torch.cat((le, le - ri)).numpy()  # prep/preprocess_kitti.py:244-247
                    err = float(np.abs(x - xs[i]).max())
                    if best is None or err < best[0]:
                        best = (err, k)
                k_rows[i] = np.asarray(best[1], dtype=np.float32)
                worst = max(worst, best[0])
            print('kat', mode, phase, kps.shape, 'max |X - preprocess(kps,K)| =', worst)
            assert worst < 2e-6
            np.savez_compressed(os.path.join(OUT, 'kat_%s_%s.npz' % (mode, phase)), kps=kps, K=k_rows, X=xs, Y=ys)


def ref_forward():
    cfgs = [
        # name, kind, in, out, L, stages, seed, n_rows
        ('loco_mono_l128', 'loco', 34, 9, 128, 3, 11, 77),
        ('loco_stereo_l128', 'loco', 68, 10, 128, 3, 12, 77),
        ('loco_mono_l256_s2', 'loco', 34, 9, 256, 2, 13, 41),
        ('monoloco_l256_o2', 'monoloco', 34, 2, 256, 3, 14, 50),
        ('monoloco_l128_o9', 'monoloco', 34, 9, 128, 3, 15, 50),
        ('loco_mono_l1024', 'loco', 34, 9, 1024, 3, 1, 169),
        ('loco_stereo_l1024', 'loco', 68, 10, 1024, 3, 2, 96),
        ('monoloco_l1024_o9', 'monoloco', 34, 9, 1024, 3, 3, 96),
    ]
    kat = np.load(os.path.join(OUT, 'kat_mono_val.npz'))
    for name, kind, isz, osz, L, st, seed, n in cfgs:
        model, sd = build(kind, isz, osz, L, st, seed)
        if name == 'loco_mono_l1024':
            x = kat['X'][:n]  # BASELINE config 1: the reference fixture's val rows
        else:
            x = synthetic.make_inputs(n, isz, seed=100 + seed)
        with torch.no_grad():
            out = model(torch.from_numpy(x))
        save = dict(x=x, out=out.numpy(), cfg=np.array([isz, osz, L, st, seed]), kind=kind,
                    checksum=sd_checksum(sd))
        if kind == 'loco' or osz == 9:
            dec = extract_outputs(out) if kind == 'loco' else extract_outputs_mono(out)
            save.update(dic_to_np(dec, 'dec_'))
        np.savez_compressed(os.path.join(OUT, 'ref_fwd_%s.npz' % name), **save)
        print('fwd', name, out.shape, float(out.abs().mean()))


def ref_forward_wide():
    """`--hidden_size` values outside the reference's default (run.py:101,122; hyp_tuning.py:52 searches 2048): the real
    nn.Module at the widest width the kernels cover and at widths the packer has to pad.  Written as ref_wide_*.npz so
    that the ref_fwd_* parametrisations stay as they were."""
    cfgs = [
        # name, kind, in, out, L, stages, seed, n_rows
        ('loco_mono_l2048', 'loco', 34, 9, 2048, 3, 21, 150),
        ('loco_stereo_l2048', 'loco', 68, 10, 2048, 3, 22, 64),
        ('loco_mono_l1500_s1', 'loco', 34, 9, 1500, 1, 23, 90),
        ('loco_stereo_l300_s2', 'loco', 68, 10, 300, 2, 24, 70),
        ('monoloco_l200_o9', 'monoloco', 34, 9, 200, 3, 25, 50),
    ]
    for name, kind, isz, osz, L, st, seed, n in cfgs:
        model, sd = build(kind, isz, osz, L, st, seed)
        x = synthetic.make_inputs(n, isz, seed=100 + seed)
        with torch.no_grad():
            out = model(torch.from_numpy(x))
        save = dict(x=x, out=out.numpy(), cfg=np.array([isz, osz, L, st, seed]), kind=kind, checksum=sd_checksum(sd))
        dec = extract_outputs(out) if kind == 'loco' else extract_outputs_mono(out)
        save.update(dic_to_np(dec, 'dec_'))
        np.savez_compressed(os.path.join(OUT, 'ref_wide_%s.npz' % name), **save)
        print('wide', name, out.shape, float(out.abs().mean()))


def ref_loco_forward():
    """Full Loco.forward (pre + model + post) on the reference's pifpaf fixture (mono) and on the
    stereo fixture's left/right keypoints (stereo, all-vs-all + filter)."""
    ann = json.load(open(os.path.join(REF, 'tests', '002282.png.pifpaf.json')))
    boxes, keypoints = preprocess_pifpaf(ann, im_size=(1238, 374))
    kk = synthetic.KITTI_K
    model, sd = build('loco', 34, 9, 1024, 3, 1)
    net = Loco(model=model, mode='mono', device=torch.device('cpu'))
    dic = net.forward(keypoints, kk)
    save = dict(keypoints=np.asarray(keypoints, dtype=np.float32), boxes=np.asarray(boxes, dtype=np.float32),
                K=np.asarray(kk, dtype=np.float32), checksum=sd_checksum(sd))
    save.update(dic_to_np(dic, 'out_'))
    # post_process pieces used by Loco.post_process (net.py:192-215)
    uv_centers = get_keypoints(keypoints, mode='center')
    xy_centers = pixel_to_camera(uv_centers, kk, 1)
    xyz = xyz_from_distance(dic['d'], xy_centers)
    save['xy_centers'] = xy_centers.numpy()
    save['xyz_from_distance'] = xyz.numpy()
    np.savez_compressed(os.path.join(OUT, 'ref_loco_mono_pifpaf.npz'), **save)
    print('loco mono', {k: v.shape for k, v in save.items() if hasattr(v, 'shape')})

    kat = np.load(os.path.join(OUT, 'kat_stereo_val.npz'))
    left = kat['kps'][:12, :, :17]
    right = kat['kps'][:9, :, 17:]
    model, sd = build('loco', 68, 10, 1024, 3, 2)
    net = Loco(model=model, mode='stereo', device=torch.device('cpu'))
    dic = net.forward(left.tolist(), kk, right.tolist())
    save = dict(left=left, right=right, K=np.asarray(kk, dtype=np.float32), checksum=sd_checksum(sd))
    save.update(dic_to_np(dic, 'out_'))
    inputs, _ = preprocess_monstereo(torch.from_numpy(left), torch.from_numpy(right), torch.tensor(kk))
    with torch.no_grad():
        raw = model(inputs)
    clustered = cluster_outputs(raw, right.shape[0])
    fin, mask = filter_outputs(clustered)
    save['pairs_x'] = inputs.numpy()
    save['pairs_raw'] = raw.numpy()
    save['filter_mask'] = mask.numpy()
    save['filter_out'] = fin.numpy()
    # no right poses: net.py:115-116
    dic1 = net.forward(left.tolist(), kk, None)
    save.update(dic_to_np(dic1, 'noright_'))
    np.savez_compressed(os.path.join(OUT, 'ref_loco_stereo.npz'), **save)
    print('loco stereo', save['out_xyzd'].shape, save['filter_mask'].sum())


def ref_losses():
    """MultiTaskLoss / AutoTuneMultiTaskLoss values and the full train-step gradients
    (train-mode BN batch statistics, dropout p=0 so no RNG is involved)."""
    for mode, isz, osz, seed in (('mono', 34, 9, 21), ('stereo', 68, 10, 22)):
        tasks = ('d', 'x', 'y', 'h', 'w', 'l', 'ori') + (('aux',) if mode == 'stereo' else ())
        lambdas = (1,) * len(tasks)
        L, st, B = 128, 2, 48
        sd = synthetic.make_state_dict('loco', isz, osz, L, st, seed)
        model = LocoModel(isz, osz, L, 0.0, st, device='cpu')
        model.load_state_dict(sd_to_torch(sd))
        model.train()
        x = synthetic.make_inputs(B, isz, seed=200 + seed)
        y = synthetic.make_labels(B, stereo=(mode == 'stereo'), seed=300 + seed)
        for auto in (False, True):
            losses_tr, losses_val = CompositeLoss(tasks)()
            if auto:
                mt = AutoTuneMultiTaskLoss(losses_tr, losses_val, lambdas, tasks)
                with torch.no_grad():
                    mt.log_sigmas.copy_(torch.linspace(-0.3, 0.4, len(tasks)))
            else:
                mt = MultiTaskLoss(losses_tr, losses_val, lambdas, tasks)
            model.load_state_dict(sd_to_torch(sd))
            model.zero_grad()
            out = model(torch.from_numpy(x))
            out.retain_grad()
            loss, vals = mt(out, torch.from_numpy(y), phase='train')
            loss.backward()
            save = dict(x=x, y=y, out=out.detach().numpy(), loss=float(loss), vals=np.array([float(v) for v in vals]),
                        dout=out.grad.numpy(), cfg=np.array([isz, osz, L, st, seed, B]), checksum=sd_checksum(sd))
            for k, p in model.named_parameters():
                save['grad.' + k] = p.grad.numpy()
            for k, b in model.named_buffers():
                save['buf.' + k] = b.detach().numpy()  # running stats after one train-mode forward
            if auto:
                save['grad.log_sigmas'] = mt.log_sigmas.grad.numpy()
                save['log_sigmas'] = mt.log_sigmas.detach().numpy()
            with torch.no_grad():
                _, vals_val = mt(out.detach(), torch.from_numpy(y), phase='val')
            save['vals_val'] = np.array([float(v) for v in vals_val])
            np.savez_compressed(os.path.join(OUT, 'ref_train_%s_%s.npz' % (mode, 'auto' if auto else 'mtl')), **save)
            print('train', mode, auto, float(loss))


def ref_epistemic():
    """laplace_sampling (process.py:101-122) population check + unnormalize_bi."""
    mu = torch.tensor([[10.0, 0.5], [25.0, 2.0], [40.0, 4.0]])
    xx = laplace_sampling(mu, 100)
    np.savez_compressed(os.path.join(OUT, 'ref_laplace_sampling.npz'), mu_bi=mu.numpy(), samples=xx.numpy(),
                        bi=unnormalize_bi(torch.tensor([[10.0, -1.0], [25.0, 0.2]])).numpy())


def ref_api():
    """Host-side API fixtures: checkpoint ABI (state_dict keys/shapes), preprocess_pifpaf, load_calibration,
    Loco.post_process with a synthetic ground truth."""
    from monoloco.network.process import load_calibration
    abi = {}
    for name, m in (('loco_34_9_1024', LocoModel(34, 9, 1024, device='cpu')),
                    ('loco_68_10_1024', LocoModel(68, 10, 1024, device='cpu')),
                    ('loco_34_9_256_s2', LocoModel(34, 9, 256, num_stage=2, device='cpu')),
                    ('monoloco_34_2_256', MonolocoModel(34, 2, 256)),
                    ('monoloco_34_9_1024', MonolocoModel(34, 9, 1024))):
        abi[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
    ann = json.load(open(os.path.join(REF, 'tests', '002282.png.pifpaf.json')))
    with open(os.path.join(OUT, 'pifpaf_002282.json'), 'w') as f:  # data fixture: 16 pifpaf detections
        json.dump(ann, f)
    boxes, keypoints = preprocess_pifpaf(ann, im_size=(1238, 374))
    boxes2, _ = preprocess_pifpaf(json.load(open(os.path.join(REF, 'tests', '002282.png.pifpaf.json'))),
                                  im_size=None, enlarge_boxes=False, min_conf=0.3)
    calib = {'kitti_1238x374': load_calibration('kitti', (1238, 374)),
             'nuscenes_800x450': load_calibration('nuscenes', (800, 450)),
             'custom_1920x1080': load_calibration('custom', (1920, 1080), focal_length=5.7)}
    kk = synthetic.KITTI_K
    model, sd = build('loco', 34, 9, 1024, 3, 1)
    net = Loco(model=model, mode='mono', device=torch.device('cpu'))
    dic = net.forward(keypoints, kk)
    # synthetic ground truth: shifted copies of 5 detections (matched) + 1 far-away box (unmatched)
    boxes_gt = [[b[0] + 3, b[1] - 2, b[2] + 3, b[3] - 2] for b in boxes[2:7]] + [[5., 5., 20., 40.]]
    ys = [[0, 0, 0, 10.0 + 3 * i, 0, 0, 0, 0, 0, 0] for i in range(len(boxes_gt))]
    dic_gt = {'boxes': boxes_gt, 'ys': ys}
    post = Loco.post_process(dic, boxes, keypoints, kk, dic_gt=dic_gt)
    post_nogt = Loco.post_process(dic, boxes, keypoints, kk, dic_gt=None)
    out = {'abi': abi, 'boxes': boxes, 'keypoints': keypoints, 'boxes_noenlarge_conf03': boxes2, 'calib': calib,
           'dic_gt': dic_gt, 'post': {k: v for k, v in post.items()}, 'post_nogt': {k: v for k, v in post_nogt.items()}}
    with open(os.path.join(OUT, 'ref_api.json'), 'w') as f:
        json.dump(out, f)
    print('api', len(boxes), len(boxes2), sorted(post.keys()))


def ref_dataset_order():
    """Batches of the reference's `DataLoader(KeypointsDataset(...), batch_size=bs, shuffle=True)` (trainer.py:102-103)
    under torch.manual_seed: row ids per batch for two epochs of train and one of val, plus cluster sizes."""
    import tempfile
    from torch.utils.data import DataLoader
    from monoloco.train.datasets import KeypointsDataset
    path = os.path.join(tempfile.mkdtemp(), 'joints.json')
    synthetic.make_joints_json(path)
    out = {'seed': 11, 'bs': 32, 'joints_seed': 5}
    torch.manual_seed(out['seed'])
    loaders = {ph: DataLoader(KeypointsDataset(path, phase=ph), batch_size=out['bs'], shuffle=True) for ph in ('train', 'val')}
    order = []
    for epoch in range(2):
        for ph in ('train', 'val'):
            ids = []
            for inputs, labels, names, kps in loaders[ph]:
                ids.append([int(v) for v in inputs[:, 0].tolist()])
                assert [int(nm[:6]) for nm in names] == ids[-1] and kps.shape[1:] == (3, 17)
            order.append({'epoch': epoch, 'phase': ph, 'batches': ids})
    out['order'] = order
    ds = KeypointsDataset(path, phase='val')
    out['clusters'] = {k: int(ds.get_cluster_annotations(k)[2]) for k in ('10', '20', '30', '>30')}
    out['x_sum'] = float(ds.inputs_all.double().sum())
    out['version'] = ds.get_version()
    with open(os.path.join(OUT, 'ref_dataset_order.json'), 'w') as f:
        json.dump(out, f)
    print('ref_dataset_order.json', len(order), 'passes')


def ref_kitti_txt():
    """Text of the reference's `save_txts` (eval/generate_kitti.py:202-253) for every `net` branch on synthetic detections.
    monoloco.eval asserts a data/logs directory at import time (eval_kitti.py:48), so the import runs from a scratch cwd."""
    import tempfile
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, 'data', 'logs'))
    os.chdir(tmp)
    try:
        from monoloco.eval.generate_kitti import save_txts
    finally:
        os.chdir(cwd)
    out = {}
    for i, (net, n) in enumerate((('monoloco_pp', 7), ('monstereo', 5), ('monoloco', 6), ('geometric', 6), ('baseline', 4),
                                  ('monoloco_pp', 0), ('monoloco_pp', 33))):
        boxes, outs, params, cat = synthetic.make_kitti_case(net, n, seed=i)
        path = os.path.join(tmp, 'out_%d.txt' % i)
        save_txts(path, boxes, outs, params, net=net, cat=cat)
        out['%d:%s:%d' % (i, net, n)] = open(path).read()
    with open(os.path.join(OUT, 'ref_kitti_txt.json'), 'w') as f:
        json.dump(out, f)
    print('ref_kitti_txt.json', len(out), 'files')


def ref_activity():
    """Outputs of the reference's activity helpers (monoloco/activity.py:17-67 social_interactions, probabilistic and
    deterministic; :70-117 is_raising_hand) on a seeded crowd of 12 people standing 0.6-3 m apart and on the pifpaf
    fixture's poses.  The seed is the first one whose flags contain both outcomes."""
    from monoloco.activity import social_interactions, is_raising_hand
    with open(os.path.join(OUT, 'pifpaf_002282.json')) as f:
        _, keypoints = preprocess_pifpaf(json.load(f), im_size=(1238, 374))
    raising = [is_raising_hand(k) for k in keypoints]
    for seed in range(200):
        rng = np.random.RandomState(seed)
        n = 12
        centers = np.stack([rng.uniform(-1.0, 2.5, n), rng.uniform(4.5, 9.5, n)], 1).tolist()
        angles = rng.uniform(-math.pi, math.pi, n).tolist()
        dds = [math.hypot(c[0], c[1]) for c in centers]
        stds = rng.uniform(0.2, 1.0, n).tolist()
        prob = [bool(social_interactions(i, centers, angles, dds, stds=stds)) for i in range(n)]
        det = [bool(social_interactions(i, centers, angles, dds, stds=stds, n_samples=1)) for i in range(n)]
        if 2 <= sum(prob) <= n - 2 and 1 <= sum(det) and prob != det:
            break
    out = {'seed': seed, 'centers': centers, 'angles': angles, 'dds': dds, 'stds': stds, 'prob': prob, 'det': det,
           'raising': raising}
    with open(os.path.join(OUT, 'ref_activity.json'), 'w') as f:
        json.dump(out, f)
    print('ref_activity.json seed', seed, 'prob', sum(prob), 'det', sum(det))


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    steps = {'kat': kat_preprocess, 'forward': ref_forward, 'loco': ref_loco_forward, 'losses': ref_losses,
             'epistemic': ref_epistemic, 'api': ref_api, 'dataset': ref_dataset_order, 'kitti': ref_kitti_txt,
             'activity': ref_activity, 'wide': ref_forward_wide}
    for name in (sys.argv[1:] or list(steps)):   # python oracle/gen_golden.py [step ...]
        steps[name]()
    print('done')
