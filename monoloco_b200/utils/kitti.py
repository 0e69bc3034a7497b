"""
KITTI result files: the txt wire format `eval/generate_kitti.py:202-253 save_txts` writes after
`Loco.forward`, one line per detection:

    <Pedestrian|Cyclist> -1 -1 alpha x1 y1 x2 y2 h w l x y z ry conf bi epi        (15 numbers, "%f ")

Computed column-wise (float64 numpy, the same operation order as the reference's per-instance Python arithmetic, so the
files are byte-identical) and formatted with one template per row instead of 18 writes per detection.
"""
import numpy as np

from .camera import xyz_from_distance

_CONF_SCALE = {'monoloco_pp': 0.035, 'monstereo': 0.033}   # "(approximately) same recall at evaluation", generate_kitti.py:231
_ROW = '%s -1 -1 ' + '%f ' * 15 + '\n'


def _col(v, n):
    """[n] float64 column from a tensor [n,1] / [n], a list of floats or a list of 0-d tensors."""
    if hasattr(v, 'detach'):
        v = v.detach().cpu().numpy()
    return np.asarray([float(e) for e in v] if isinstance(v, (list, tuple)) else v, dtype=np.float64).reshape(n)


def kitti_rows(all_inputs, all_outputs, all_params, net='monoloco', cat=None):
    """(category names, [n,15] float64 table) of one image -- the numbers `save_txts` prints."""
    assert net in ('monoloco', 'monstereo', 'geometric', 'baseline', 'monoloco_pp')
    uv_boxes = all_inputs
    n = len(uv_boxes)
    tt = np.zeros(3)
    zzs_geom = None
    if net in ('monstereo', 'monoloco_pp'):
        xyzd, bis, epis, yaws, hs, ws, ls = all_outputs[:]
        xyz = xyzd[:, 0:3]
    elif net in ('monoloco', 'geometric'):
        dds, bis, epis, zzs_geom, xy_centers = all_outputs[:]
        xyz = xyz_from_distance(dds, xy_centers)
    else:
        _, tt_in = all_params[:]
        tt = np.asarray([float(t) for t in tt_in], dtype=np.float64)
        xyz, bis, epis, zzs_geom, xy_centers = all_outputs[:]
    assert n == len(list(xyz)), "Number of inputs different from number of outputs"
    table = np.zeros((n, 15), dtype=np.float64)
    if n == 0:
        return [], table
    if hasattr(xyz, 'detach'):
        xyz = xyz.detach().cpu().numpy()
    cam = np.asarray([[float(c) for c in row[:3]] for row in xyz], dtype=np.float64).reshape(n, 3) - tt
    if net == 'geometric':
        cam[:, 2] = _col(zzs_geom, n)
    boxes = np.asarray([[float(b) for b in box] for box in uv_boxes], dtype=np.float64)   # x1 y1 x2 y2 conf
    bi, epi = _col(bis, n), _col(epis, n)
    if net in _CONF_SCALE:
        table[:, 0], table[:, 11] = _col(yaws[0], n), _col(yaws[1], n)   # alpha, ry
        table[:, 5], table[:, 6], table[:, 7] = _col(hs, n), _col(ws, n), _col(ls, n)
        conf_scale = _CONF_SCALE[net]
    else:
        table[:, 0] = table[:, 11] = -10.
        conf_scale = 0.05
    table[:, 1:5] = boxes[:, :4]
    table[:, 8:11] = cam
    # conf_scale * box_conf / (bi / |xyz|), generate_kitti.py:236
    table[:, 12] = conf_scale * boxes[:, -1] / (bi / np.sqrt(cam[:, 0] ** 2 + cam[:, 1] ** 2 + cam[:, 2] ** 2))
    table[:, 13], table[:, 14] = bi, epi
    names = ['Pedestrian' if float(c) < 0.1 else 'Cyclist' for c in cat]
    return names, table


def save_txts(path_txt, all_inputs, all_outputs, all_params, net='monoloco', cat=None):
    """Drop-in for `monoloco.eval.generate_kitti.save_txts` (same arguments, byte-identical file)."""
    names, table = kitti_rows(all_inputs, all_outputs, all_params, net=net, cat=cat)
    with open(path_txt, 'w+') as ff:
        ff.write(''.join(_ROW % ((name,) + tuple(row)) for name, row in zip(names, table.tolist())))
