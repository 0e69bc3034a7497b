"""
CPU oracle for the monoloco per-detection hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy (float32) restatement of the reference algorithm.  It is NOT part of the
product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it, and only as the checker / the timed CPU baseline.  `monoloco_b200/` never
imports anything from `oracle/`.

Parity pinning: the restatement is checked (tests/test_oracle_golden.py) against
  * the reference's own fixtures (`tests/golden/kat_*.npz` extracted from
    /root/reference/tests/sample_joints-kitti-{mono,stereo}.json: stored X == preprocess(kps, K)),
  * outputs of the *real* reference modules (imported from /root/reference in the build container
    by `oracle/gen_golden.py`, committed under tests/golden/ref_*.npz): LocoModel / MonolocoModel
    forward, preprocess_monoloco / preprocess_monstereo, extract_outputs(_mono), cluster/filter,
    xyz_from_distance, LaplacianLoss / MultiTaskLoss values and gradients.
The reference holds no golden vectors for the network forward itself; those are
pinned by the live-reference fixtures above.

Every function cites the reference file:line it follows (paths relative to /root/reference).
All state dicts are {key: np.ndarray} with the reference's state_dict key names.
"""
import math

import numpy as np

F32 = np.float32
BN_EPS = F32(1e-5)  # torch.nn.BatchNorm1d default eps, monoloco/network/architectures.py:25,84


# ------------------------------------------------------------------------------------------------
# utils/camera.py
# ------------------------------------------------------------------------------------------------
def pixel_to_camera(uv, kk, z_met):
    """monoloco/utils/camera.py:10-29.  uv: (m,2) or (m,x,2) or (m,2,x); kk: (3,3)."""
    uv = np.asarray(uv, dtype=F32)
    kk = np.asarray(kk, dtype=F32)
    if uv.shape[-1] != 2:
        uv = np.transpose(uv, (0, 2, 1))  # camera.py:21
        assert uv.shape[-1] == 2, "Tensor size not recognized"
    ones = np.ones(uv.shape[:-1] + (1,), dtype=F32)
    uv_padded = np.concatenate([uv, ones], axis=-1)  # camera.py:23
    kk_1 = np.linalg.inv(kk).astype(F32)  # camera.py:25 (torch.inverse, fp32)
    xyz = np.matmul(uv_padded, kk_1.T).astype(F32)  # camera.py:26
    return (xyz * F32(z_met)).astype(F32)  # camera.py:27


def get_keypoints(keypoints, mode):
    """monoloco/utils/camera.py:69-107.  keypoints (m,3,17) or (3,17) -> (m,2)."""
    kps = np.asarray(keypoints, dtype=F32)
    if kps.ndim == 2:
        kps = kps[None]
    assert kps.ndim == 3 and kps.shape[1] == 3, "tensor dimensions not recognized"
    assert mode in ['center', 'bottom', 'head', 'shoulder', 'hip', 'ankle']
    kps_in = kps[:, 0:2, :]
    if mode == 'center':
        kmax = kps_in.max(2)
        kmin = kps_in.min(2)
        return ((kmax - kmin) / F32(2) + kmin).astype(F32)  # camera.py:86
    if mode == 'bottom':
        kmax = kps_in.max(2)
        kmin = kps_in.min(2)
        x = (kmax[:, 0:1] - kmin[:, 0:1]) / F32(2) + kmin[:, 0:1]
        return np.concatenate([x, kmax[:, 1:2]], -1).astype(F32)
    sl = {'head': slice(0, 5), 'shoulder': slice(5, 7), 'hip': slice(11, 13), 'ankle': slice(15, 17)}[mode]
    return kps_in[:, :, sl].mean(2).astype(F32)


def xyz_from_distance(distances, xy_centers):
    """monoloco/utils/camera.py:161-177."""
    d = np.asarray(distances, dtype=F32)
    if d.ndim == 0:
        d = d[None]
    if d.ndim == 1:
        d = d[:, None]
    c = np.asarray(xy_centers, dtype=F32)
    if c.ndim == 1:
        c = c[None]
    assert c.shape[-1] == 3 and d.shape[-1] == 1, "Size of tensor not recognized"
    return (c * d / np.sqrt(F32(1) + c[:, 0:1] ** 2 + c[:, 1:2] ** 2)).astype(F32)


def back_correct_angles(yaws, xyz):
    """monoloco/utils/camera.py:202-208 (single wrap by +-2pi, applied sequentially)."""
    corrections = np.arctan2(xyz[:, 0], xyz[:, 2]).astype(F32)
    yaws = (yaws + corrections.reshape(-1, 1)).astype(F32)
    yaws = np.where(yaws > F32(math.pi), (yaws - F32(2 * math.pi)).astype(F32), yaws)
    yaws = np.where(yaws < F32(-math.pi), (yaws + F32(2 * math.pi)).astype(F32), yaws)
    return yaws.astype(F32)


def to_cartesian_xy(rtp):
    """monoloco/utils/camera.py:226-237, modes 'x' and 'y' on (m,3) = (theta, psi, r)."""
    t, p, r = rtp[:, 0], rtp[:, 1], rtp[:, 2]
    x = (r * np.sin(p) * np.cos(t)).astype(F32)
    y = (r * np.cos(p)).astype(F32)
    return x.reshape(-1, 1), y.reshape(-1, 1)


# ------------------------------------------------------------------------------------------------
# network/process.py
# ------------------------------------------------------------------------------------------------
def preprocess_monoloco(keypoints, kk, zero_center=False):
    """monoloco/network/process.py:47-67.  (m,3,17) + K -> (m,34) interleaved x0,y0,...,x16,y16."""
    kps = np.asarray(keypoints, dtype=F32)
    uv_center = get_keypoints(kps, mode='center')
    xy1_center = pixel_to_camera(uv_center, kk, 10)
    xy1_all = pixel_to_camera(kps[:, 0:2, :], kk, 10)  # (m,17,3)
    if zero_center:
        kps_norm = xy1_all - xy1_center[:, None, :]
    else:
        kps_norm = xy1_all
    return np.ascontiguousarray(kps_norm[:, :, 0:2]).reshape(kps_norm.shape[0], -1).astype(F32)


def preprocess_monstereo(keypoints, keypoints_r, kk):
    """monoloco/network/process.py:25-44.  all-vs-all rows (l*R + r) = cat(l, l - r) -> (L*R, 68)."""
    inputs_l = preprocess_monoloco(keypoints, kk)
    inputs_r = preprocess_monoloco(keypoints_r, kk)
    n_l, n_r = inputs_l.shape[0], inputs_r.shape[0]
    left = np.repeat(inputs_l, n_r, axis=0)
    right = np.tile(inputs_r, (n_l, 1))
    inputs = np.concatenate([left, (left - right).astype(F32)], axis=1).astype(F32)
    clusters = [n_r] * n_l
    return inputs, clusters


def unnormalize_bi(loc):
    """monoloco/network/process.py:125-133: bi = exp(s) * d."""
    assert loc.shape[1] == 2, "size of the output tensor should be (m, 2)"
    return (np.exp(loc[:, 1:2]) * loc[:, 0:1]).astype(F32)


def _sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x))).astype(F32)


def extract_outputs(outputs, tasks=()):
    """monoloco/network/process.py:231-278."""
    outputs = np.asarray(outputs, dtype=F32)
    dic_out = {'x': outputs[:, 0:1], 'y': outputs[:, 1:2], 'd': outputs[:, 2:4], 'h': outputs[:, 4:5],
               'w': outputs[:, 5:6], 'l': outputs[:, 6:7], 'ori': outputs[:, 7:9]}
    if outputs.shape[1] == 10:
        dic_out['aux'] = outputs[:, 9:10]
    if len(tasks) >= 1:
        assert isinstance(tasks, tuple), "tasks need to be a tuple"
        return [dic_out[task] for task in tasks]
    dic_out['bi'] = unnormalize_bi(dic_out['d'])
    with np.errstate(invalid='ignore'):
        x, y = to_cartesian_xy(outputs[:, 0:3])
        d = dic_out['d'][:, 0:1]
        z = np.sqrt(d ** 2 - x ** 2 - y ** 2).astype(F32)  # NaN when inconsistent, as the reference
    dic_out['xyzd'] = np.concatenate([x, y, z, d], axis=1).astype(F32)
    dic_out.pop('x')
    dic_out.pop('y')
    dic_out['d'] = d
    yaw_pred = np.arctan2(dic_out['ori'][:, 0:1], dic_out['ori'][:, 1:2]).astype(F32)
    yaw_orig = back_correct_angles(yaw_pred.copy(), dic_out['xyzd'][:, 0:3])
    dic_out['yaw'] = (yaw_pred, yaw_orig)
    if outputs.shape[1] == 10:
        dic_out['aux'] = _sigmoid(dic_out['aux'])
    return dic_out


def extract_outputs_mono(outputs, tasks=None):
    """monoloco/network/process.py:330-360 (legacy monoloco_p)."""
    outputs = np.asarray(outputs, dtype=F32)
    dic_out = {'xyz': outputs[:, 0:3], 'zb': outputs[:, 2:4], 'h': outputs[:, 4:5], 'w': outputs[:, 5:6],
               'l': outputs[:, 6:7], 'ori': outputs[:, 7:9]}
    if tasks is not None:
        assert isinstance(tasks, tuple), "tasks need to be a tuple"
        return [dic_out[task] for task in tasks]
    bi = unnormalize_bi(dic_out['zb'])
    dd = np.sqrt((dic_out['xyz'].astype(F32) ** 2).sum(1)).astype(F32).reshape(-1, 1)
    dic_out['xyzd'] = np.concatenate([dic_out['xyz'], dd], axis=1).astype(F32)
    dic_out['d'], dic_out['bi'] = dd, bi
    yaw_pred = np.arctan2(dic_out['ori'][:, 0:1], dic_out['ori'][:, 1:2]).astype(F32)
    yaw_orig = back_correct_angles(yaw_pred.copy(), dic_out['xyzd'][:, 0:3])
    dic_out['yaw'] = (yaw_pred, yaw_orig)
    return dic_out


def extract_labels(labels, tasks=None):
    """monoloco/network/process.py:293-304."""
    dic = {'x': labels[:, 0:1], 'y': labels[:, 1:2], 'z': labels[:, 2:3], 'd': labels[:, 3:4],
           'h': labels[:, 4:5], 'w': labels[:, 5:6], 'l': labels[:, 6:7], 'ori': labels[:, 7:9],
           'aux': labels[:, 10:11]}
    if tasks is not None:
        assert isinstance(tasks, tuple), "tasks need to be a tuple"
        return [dic[t] for t in tasks]
    return dic


def cluster_outputs(outputs, clusters):
    """monoloco/network/process.py:307-316."""
    if clusters == 0:
        clusters = max(1, round(outputs.shape[0] / 2))
    assert outputs.shape[0] % clusters == 0, "Unexpected number of inputs"
    return outputs.reshape(-1, clusters, outputs.shape[1])


def filter_outputs(outputs):
    """monoloco/network/process.py:319-327: keep, per left pose, every row whose aux logit >= max (ties kept)."""
    val = outputs[:, :, -1]
    best = val.max(axis=1, keepdims=True)
    mask = val >= best
    return outputs[mask], mask


# ------------------------------------------------------------------------------------------------
# network/architectures.py  (eval mode: BN running stats, Dropout = identity unless a mask is given)
# ------------------------------------------------------------------------------------------------
def _linear(x, sd, name):
    """nn.Linear: y = x W^T + b (weight [out,in])."""
    return (x @ sd[name + '.weight'].T.astype(F32) + sd[name + '.bias']).astype(F32)


def _bn_eval(x, sd, name):
    """nn.BatchNorm1d eval: (x - running_mean) / sqrt(running_var + eps) * weight + bias."""
    inv = (F32(1) / np.sqrt(sd[name + '.running_var'].astype(F32) + BN_EPS)).astype(F32)
    return ((x - sd[name + '.running_mean']) * inv * sd[name + '.weight'] + sd[name + '.bias']).astype(F32)


def _relu(x):
    return np.maximum(x, F32(0))


def _drop(x, mask, p):
    """nn.Dropout in training mode with an explicit keep-mask (1 = keep): x * mask / (1 - p)."""
    if mask is None:
        return x
    return (x * mask.astype(F32) * F32(1.0 / (1.0 - p))).astype(F32)


def _stage(x, sd, prefix):
    """MyLinearSimple / MyLinear forward, architectures.py:88-102 / 162-176 (eval)."""
    y = _relu(_bn_eval(_linear(x, sd, prefix + '.w1'), sd, prefix + '.batch_norm1'))
    y = _relu(_bn_eval(_linear(y, sd, prefix + '.w2'), sd, prefix + '.batch_norm2'))
    return (x + y).astype(F32)


def num_stages(sd):
    n = 0
    while 'linear_stages.%d.w1.weight' % n in sd:
        n += 1
    return n


def loco_model_forward(sd, x, drop_masks=None, p_dropout=0.2):
    """LocoModel.forward, architectures.py:48-71.  drop_masks = (mask_after_bn1, mask_after_bn3) keep
    masks for the MC-dropout mode of net.py:141 (only the top-level self.dropout is re-enabled)."""
    x = np.asarray(x, dtype=F32)
    m1, m3 = drop_masks if drop_masks is not None else (None, None)
    y = _relu(_bn_eval(_linear(x, sd, 'w1'), sd, 'batch_norm1'))
    y = _drop(y, m1, p_dropout)
    for i in range(num_stages(sd)):
        y = _stage(y, sd, 'linear_stages.%d' % i)
    y = _linear(y, sd, 'w2')  # architectures.py:59
    aux = _linear(y, sd, 'w_aux')  # :60
    y = _relu(_bn_eval(_linear(y, sd, 'w3'), sd, 'batch_norm3'))  # :63-65
    y = _drop(y, m3, p_dropout)
    y = _linear(y, sd, 'w_fin')  # :67
    return np.concatenate([y, aux], axis=1).astype(F32)  # :70


def monoloco_model_forward(sd, x, drop_masks=None, p_dropout=0.2):
    """MonolocoModel.forward, architectures.py:135-145."""
    x = np.asarray(x, dtype=F32)
    m1 = drop_masks[0] if drop_masks is not None else None
    y = _relu(_bn_eval(_linear(x, sd, 'w1'), sd, 'batch_norm1'))
    y = _drop(y, m1, p_dropout)
    for i in range(num_stages(sd)):
        y = _stage(y, sd, 'linear_stages.%d' % i)
    return _linear(y, sd, 'w2')


def model_forward(sd, x, **kw):
    """Dispatch on the checkpoint ABI: LocoModel has w_fin/w_aux/w3."""
    if 'w_fin.weight' in sd:
        return loco_model_forward(sd, x, **kw)
    return monoloco_model_forward(sd, x, **kw)


# ------------------------------------------------------------------------------------------------
# network/net.py :: Loco.forward (net='monoloco_pp' / 'monstereo', the only reachable branches: SURVEY §8b)
# ------------------------------------------------------------------------------------------------
def loco_forward(sd, keypoints, kk, keypoints_r=None, mode='mono'):
    """monoloco/network/net.py:83-133 without MC dropout; returns the dic_out of numpy arrays."""
    if len(keypoints) == 0:
        return None
    kps = np.asarray(keypoints, dtype=F32)
    kk = np.asarray(kk, dtype=F32)
    if mode == 'mono':
        inputs = preprocess_monoloco(kps, kk)
        outputs = model_forward(sd, inputs)
        dic_out = extract_outputs(outputs)
    else:
        if keypoints_r is not None and len(keypoints_r):
            kps_r = np.asarray(keypoints_r, dtype=F32)
        else:
            kps_r = kps[0:1].copy()  # net.py:116
        inputs, _ = preprocess_monstereo(kps, kps_r, kk)
        outputs = model_forward(sd, inputs)
        outputs = cluster_outputs(outputs, kps_r.shape[0])
        outputs_fin, _ = filter_outputs(outputs)
        dic_out = extract_outputs(outputs_fin)
    dic_out['epi'] = [0.] * outputs.shape[0]  # net.py:130 (NB: stereo -> number of LEFT poses, outputs is 3-D)
    return dic_out


def laplace_std(bi):
    """Analytic std of Laplace(mu, b) = sqrt(2) b: the population value `laplace_sampling`
    (process.py:101-122) + `.std(0)` (net.py:159) estimates with 100 seeded samples per pass."""
    return (np.abs(bi) * F32(math.sqrt(2.0))).astype(F32)


# ------------------------------------------------------------------------------------------------
# train/losses.py (values only; gradients are pinned with torch autograd in oracle/torch_port.py)
# ------------------------------------------------------------------------------------------------
def laplacian_loss(mu_si, xx):
    """LaplacianLoss.forward, losses.py:112-142 (size_average=True, reduce=True, evaluate=False)."""
    mu, si = mu_si[:, 0:1], mu_si[:, 1:2]
    norm = F32(1) - mu / xx
    values = np.abs(norm) * np.exp(-si) + F32(0.01) + si + F32(2)
    return F32(values.astype(F32).mean(dtype=F32))


def l1_loss(a, b):
    return F32(np.abs(a - b).astype(F32).mean(dtype=F32))


def bce_with_logits(z, t):
    """nn.BCEWithLogitsLoss (mean): max(z,0) - z t + log(1 + exp(-|z|))."""
    v = np.maximum(z, F32(0)) - z * t + np.log1p(np.exp(-np.abs(z)))
    return F32(v.astype(F32).mean(dtype=F32))


def multi_task_loss(outputs, labels, tasks, lambdas=None, log_sigmas=None):
    """MultiTaskLoss.forward (losses.py:59-73) / AutoTuneMultiTaskLoss.forward (:28-43), phase='train'.
    Returns (loss, [per-task weighted values])."""
    outputs = np.asarray(outputs, dtype=F32)
    labels = np.asarray(labels, dtype=F32)
    lambdas = lambdas if lambdas is not None else (1,) * len(tasks)
    out = extract_outputs(outputs, tasks=tuple(tasks))
    gt = extract_labels(labels, tasks=tuple(tasks))
    vals = []
    for i, (t, o, g) in enumerate(zip(tasks, out, gt)):
        if t == 'd':
            v = laplacian_loss(o, g)
        elif t == 'aux':
            v = bce_with_logits(o, g)
        else:
            v = l1_loss(o, g)
        v = F32(lambdas[i]) * v
        if log_sigmas is not None:
            v = v / (F32(2.0) * np.exp(F32(log_sigmas[i])) ** 2)
        vals.append(F32(v))
    loss = F32(sum(vals))
    if log_sigmas is not None:
        loss = F32(loss + F32(np.sum(np.asarray(log_sigmas, dtype=F32))))
    return loss, vals


# ------------------------------------------------------------------------------------------------
# comparison rule: |a-b| <= rtol * max(|b|, scale) + tiny atol, NaN == NaN
# ------------------------------------------------------------------------------------------------
def close(a, b, rtol=1e-5, atol=1e-6, col_scale=True):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    nan_ok = np.isnan(a) == np.isnan(b)
    fin = ~np.isnan(b)
    if b.ndim >= 2 and col_scale:
        scale = np.nanmax(np.abs(np.where(fin, b, 0.0)), axis=0, keepdims=True)
    else:
        scale = np.nanmax(np.abs(np.where(fin, b, 0.0))) if b.size else 0.0
    tol = rtol * np.maximum(np.abs(np.where(fin, b, 0.0)), scale) + atol
    err = np.abs(np.where(fin & ~np.isnan(a), a - b, 0.0))
    ok = bool(nan_ok.all() and (err <= tol).all())
    worst = float((err / np.maximum(tol, 1e-300)).max()) if err.size else 0.0
    return ok, worst


def angle_close(a, b, rtol=1e-5, atol=1e-6, radius=None, lin_tol=None):
    """yaw comparison modulo 2pi (wrap at +-pi, camera.py:205-206).

    An angle atan2(s, c) is only as well determined as its arguments: an error e on (s, c) moves it by up to e / hypot(s, c).
    With `radius` (per-row hypot of the two arguments, reference values) and `lin_tol` (the tolerance the parity rule grants
    those arguments), the tolerance of a row is max(rtol * pi + atol, lin_tol / radius) -- the angle subtended by the allowed
    argument error.  Without them the rule is the flat rtol * pi + atol."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    diff = np.abs(np.where(np.isnan(a) | np.isnan(b), 0.0, a - b))
    diff = np.minimum(diff, np.abs(diff - 2 * math.pi))
    tol = np.full(diff.shape, rtol * math.pi + atol)
    if radius is not None:
        r = np.maximum(np.asarray(radius, dtype=np.float64).reshape(diff.shape), 1e-30)
        tol = np.maximum(tol, float(lin_tol) / r)
    ok = bool((np.isnan(a) == np.isnan(b)).all() and (diff <= tol).all())
    return ok, float((diff / tol).max()) if diff.size else 0.0
