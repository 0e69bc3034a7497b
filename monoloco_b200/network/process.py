"""Reference-facing helpers of monoloco/network/process.py, backed by the CUDA kernels.

preprocess_monoloco / preprocess_monstereo / extract_outputs / unnormalize_bi / filter_outputs keep the
reference's names, argument meaning and error behaviour (asserts) but run on the GPU through
libmonoloco_b200.so; the host-side list munging (pifpaf json -> lists, calibration yaml) is plain Python.
"""
import json
import os

import numpy as np
import torch

from ..engine import preprocess_device

Sx, Sy = 7.2, 5.4  # nuScenes sensor size in mm (process.py:21-22)

# camera intrinsics of the reference (monoloco/network/intrinsics.yaml:1-21)
INTRINSICS = {
    'kitti': {'intrinsics': [[718.3351, 0., 600.3891], [0., 718.3351, 181.5122], [0., 0., 1.]], 'im_size': [1238, 374]},
    'wv': {'intrinsics': [[1070.9498, 0., 987.4846], [0., 1070.726, 605.5297], [0., 0., 1.]], 'im_size': [1920, 1200]},
    'nuscenes': {'intrinsics': [[1070.9498, 0., 987.4846], [0., 1070.726, 605.5297], [0., 0., 1.]],
                 'im_size': [1600, 900]},
}


def _cuda(t):
    if isinstance(t, (list, np.ndarray)):
        t = torch.tensor(t, dtype=torch.float32)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("monoloco_b200: no CUDA device -- the hot path has no CPU fallback")
        t = t.cuda()
    return t.float()


def preprocess_monoloco(keypoints, kk, zero_center=False):
    """process.py:47-67: (m,3,17) pixel keypoints + K -> (m,34) metres at z=10, on the GPU (mlb_preprocess)."""
    kps = _cuda(keypoints)
    assert kps.dim() == 3 and kps.shape[1] == 3 and kps.shape[2] == 17, "tensor dimensions not recognized"
    return preprocess_device(kps, kk, zero_center=zero_center)


def preprocess_monstereo(keypoints, keypoints_r, kk):
    """process.py:25-44: all-vs-all rows cat(l, l - r) -> ((L*R, 68), clusters).  (The fused forward builds these
    rows inside the kernel and never materialises them; this stand-alone form exists for dataset preparation.)"""
    inputs_l = preprocess_monoloco(keypoints, kk)
    inputs_r = preprocess_monoloco(keypoints_r, kk)
    n_l, n_r = inputs_l.shape[0], inputs_r.shape[0]
    left = inputs_l.repeat_interleave(n_r, dim=0)
    right = inputs_r.repeat(n_l, 1)
    return torch.cat((left, left - right), dim=1), [n_r] * n_l


def unnormalize_bi(loc):
    """process.py:125-133."""
    assert loc.size()[1] == 2, "size of the output tensor should be (m, 2)"
    return torch.exp(loc[:, 1:2]) * loc[:, 0:1]


def extract_outputs(outputs, tasks=()):
    """process.py:231-278.  With `tasks` returns the raw column views used by the losses; without, the decoded
    dictionary of CPU tensors.  The decode itself (spherical -> xyz, bi, yaw) is what the fused kernel's epilogue
    computes; for a raw [m,9|10] tensor handed in from outside it is re-done here with the same op order."""
    dic_out = {'x': outputs[:, 0:1], 'y': outputs[:, 1:2], 'd': outputs[:, 2:4], 'h': outputs[:, 4:5],
               'w': outputs[:, 5:6], 'l': outputs[:, 6:7], 'ori': outputs[:, 7:9]}
    if outputs.shape[1] == 10:
        dic_out['aux'] = outputs[:, 9:10]
    if len(tasks) >= 1:
        assert isinstance(tasks, tuple), "tasks need to be a tuple"
        return [dic_out[task] for task in tasks]
    from ..engine import decode_device
    return decode_device(outputs)


def extract_labels_aux(labels, tasks=None):
    """process.py:281-290."""
    dic = {'aux': labels[:, 0:1]}
    if tasks is not None:
        assert isinstance(tasks, tuple), "tasks need to be a tuple"
        return [dic[t] for t in tasks]
    return {k: v.detach().cpu() for k, v in dic.items()}


def extract_labels(labels, tasks=None):
    """process.py:293-304."""
    dic = {'x': labels[:, 0:1], 'y': labels[:, 1:2], 'z': labels[:, 2:3], 'd': labels[:, 3:4], 'h': labels[:, 4:5],
           'w': labels[:, 5:6], 'l': labels[:, 6:7], 'ori': labels[:, 7:9], 'aux': labels[:, 10:11]}
    if tasks is not None:
        assert isinstance(tasks, tuple), "tasks need to be a tuple"
        return [dic[t] for t in tasks]
    return {k: v.detach().cpu() for k, v in dic.items()}


def cluster_outputs(outputs, clusters):
    """process.py:307-316."""
    if clusters == 0:
        clusters = max(1, round(outputs.shape[0] / 2))
    assert outputs.shape[0] % clusters == 0, "Unexpected number of inputs"
    return outputs.view(-1, clusters, outputs.shape[1])


def load_calibration(calibration, im_size, focal_length=5.7):
    """process.py:70-86."""
    if calibration == 'custom':
        return [[im_size[0] * focal_length / Sx, 0., im_size[0] / 2],
                [0., im_size[1] * focal_length / Sy, im_size[1] / 2],
                [0., 0., 1.]]
    cfg = INTRINSICS[calibration]
    kk = [list(r) for r in cfg['intrinsics']]
    scale = [size / orig for size, orig in zip(im_size, cfg['im_size'])]
    kk[0] = [el * scale[0] for el in kk[0]]
    kk[1] = [el * scale[1] for el in kk[1]]
    return kk


def factory_for_gt(path_gt, name=None):
    """process.py:89-98."""
    assert os.path.exists(path_gt), "Ground-truth file not found"
    with open(path_gt, 'r') as f:
        dic_names = json.load(f)
    return dic_names[name], dic_names[name]['K']


def prepare_pif_kps(kps_in):
    """process.py:208-216: flat list of 51 -> [xs, ys, confs]."""
    assert len(kps_in) % 3 == 0, "keypoints expected as a multiple of 3"
    return [kps_in[0:][::3], kps_in[1:][::3], kps_in[2:][::3]]


def preprocess_pifpaf(annotations, im_size=None, enlarge_boxes=True, min_conf=0.):
    """process.py:155-205: pifpaf annotations -> (boxes [x1,y1,x2,y2,conf], keypoints [3][17])."""
    boxes, keypoints = [], []
    enlarge = 1 if enlarge_boxes else 2
    for dic in annotations:
        kps = prepare_pif_kps(dic['keypoints'])
        box = list(dic['bbox'])
        if 'score' in dic:
            conf = dic['score']
            delta_h = box[3] / (10 * enlarge)
            delta_w = box[2] / (5 * enlarge)
            box[2] += box[0]
            box[3] += box[1]
        else:
            conf = float(np.mean(np.array(kps[2])))
            delta_h = (box[3] - box[1]) / (7 * enlarge)
            delta_w = (box[2] - box[0]) / (3.5 * enlarge)
            assert delta_h > -5 and delta_w > -5, "Bounding box <=0"
        box[0] -= delta_w
        box[1] -= delta_h
        box[2] += delta_w
        box[3] += delta_h
        if im_size is not None:
            box[0] = max(0, box[0])
            box[1] = max(0, box[1])
            box[2] = min(box[2], im_size[0])
            box[3] = min(box[3], im_size[1])
        if conf >= min_conf:
            box.append(conf)
            boxes.append(box)
            keypoints.append(kps)
    return boxes, keypoints
