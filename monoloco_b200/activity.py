"""Per-image activity heuristics behind `Loco.social_distance` / `Loco.raising_hand` (net.py:250-271).

Host-side geometry on a handful of people per image (outside the hot path); restated with numpy
from the behaviour of monoloco/activity.py:17-67 (F-formation test with Laplace-sampled distances), :70-117 (raised
hand from elbow angles) and :120-165 (o-space check)."""
import math

import numpy as np
import torch

_KP = dict(nose=0, l_ear=3, r_ear=4, l_sho=5, r_sho=6, l_elb=7, r_elb=8, l_hand=9, r_hand=10)


def _arm_state(kp, side):
    """(hand above shoulder, elbow angle in the reference's 90/pi units, hand tucked next to the head)."""
    xs, ys = np.asarray(kp[0], dtype=float), np.asarray(kp[1], dtype=float)
    sho, elb, hand = _KP[side + '_sho'], _KP[side + '_elb'], _KP[side + '_hand']
    fore = np.array([xs[hand] - xs[elb], ys[hand] - ys[elb]])
    upper = np.array([xs[sho] - xs[elb], ys[sho] - ys[elb]])
    cosang = float(np.dot(fore / np.linalg.norm(fore), upper / np.linalg.norm(upper)))
    angle = (90.0 / np.pi) * np.arccos(cosang)
    head_top = ys[_KP['nose']] - (xs[_KP['l_ear']] - xs[_KP['r_ear']])
    inward = xs[hand] <= xs[sho] if side == 'l' else xs[hand] >= xs[sho]
    return ys[hand] < ys[sho], angle, bool(inward and ys[hand] >= head_top)


def is_raising_hand(kp):
    """'left' | 'right' | 'both' | None for one pose [3][17] (activity.py:70-117)."""
    risen = {}
    for side in ('l', 'r'):
        up, angle, tucked = _arm_state(kp, side)
        risen[side] = bool(up and angle >= 30 and not tucked)
    if risen['l'] and risen['r']:
        return 'both'
    return 'left' if risen['l'] else ('right' if risen['r'] else None)


def check_f_formations(idx, idx_t, centers, angles, radii, social_distance=False):
    """Two people form an F-formation if, for some radius, their o-space candidates (one step along each gaze) are
    closer to each other than the people are to the o-space centre and nobody else stands inside (activity.py:120-165)."""
    pts = np.asarray([[float(c[0]), float(c[1])] for c in centers])
    others = np.delete(pts, [idx, idx_t], axis=0)
    for radius in radii:
        mu = [pts[i] + radius * np.array([math.cos(angles[i]), -math.sin(angles[i])]) for i in (idx, idx_t)]
        centre = (mu[0] + mu[1]) / 2
        gap = np.linalg.norm(mu[0] - mu[1]) * (0.5 if social_distance else 1.0)
        nearest_other = np.min(np.linalg.norm(others - centre, axis=1)) if len(others) else 100.0
        if gap <= min(np.linalg.norm(pts[idx] - centre), np.linalg.norm(pts[idx_t] - centre)) and nearest_other > radius:
            return True
    return False


def social_interactions(idx, centers, angles, dds, stds=None, social_distance=False, n_samples=100,
                        threshold_prob=0.25, threshold_dist=2, radii=(0.3, 0.5)):
    """True if person `idx` is (probably) interacting with a neighbour closer than threshold_dist (activity.py:17-67).
    With n_samples >= 2 the radial distances of both people are re-drawn from Laplace(d, std) (seeded like
    process.py:103) and the F-formation must hold in >= threshold_prob of the draws."""
    pts = np.asarray([[float(c[0]), float(c[1])] for c in centers])
    dist = np.linalg.norm(pts - pts[idx], axis=1)
    near = [j for j in np.argsort(dist)[1:] if dist[j] <= threshold_dist]
    if n_samples < 2:
        return any(check_f_formations(idx, j, centers, angles, radii, social_distance) for j in near)
    torch.manual_seed(1)
    mu = torch.tensor(dds, dtype=torch.float32)
    draws = torch.distributions.Laplace(mu, torch.abs(torch.tensor(stds, dtype=torch.float32))).sample((n_samples,)).numpy()
    for j in near:
        hits = 0
        for s in range(n_samples):
            moved = pts.copy()
            for el in (idx, j):
                # the reference does this step on float32 tensors (dds[el] - sample, then += on the python list entry)
                delta = np.float32(np.float32(dds[el]) - draws[s, el])
                theta = math.atan2(moved[el][1], moved[el][0])
                moved[el][0] = np.float32(np.float32(delta * np.float32(math.cos(theta))) + np.float32(moved[el][0]))
                moved[el][1] = np.float32(np.float32(delta * np.float32(math.sin(theta))) + np.float32(moved[el][1]))
            hits += check_f_formations(idx, j, moved, angles, radii, social_distance)
        if hits / n_samples >= threshold_prob:
            return True
    return False
