"""Box matching helpers used by Loco.post_process (host-side list logic, a handful of boxes per image).
Semantics follow monoloco/utils/iou.py:6-29, 32-41, 44-64, 87-101."""
import numpy as np


def calculate_iou(box1, box2):
    """IoU of two (x1, y1, x2, y2[, conf]) boxes (utils/iou.py:6-29)."""
    iw = max(min(box1[2], box2[2]) - max(box1[0], box2[0]), 0)
    ih = max(min(box1[3], box2[3]) - max(box1[1], box2[1]), 0)
    inter = iw * ih
    union = (box1[2] - box1[0]) * (box1[3] - box1[1]) + (box2[2] - box2[0]) * (box2[3] - box2[1]) - inter
    return inter / union


def get_iou_matrix(boxes, boxes_gt):
    """[len(boxes), len(boxes_gt)] IoU matrix (utils/iou.py:32-41)."""
    mat = np.zeros((len(boxes), len(boxes_gt)))
    for i, b in enumerate(boxes):
        for j, g in enumerate(boxes_gt):
            mat[i, j] = calculate_iou(b, g)
    return mat


def get_iou_matches(boxes, boxes_gt, iou_min=0.3):
    """Greedy matching in decreasing confidence order; each ground truth used once (utils/iou.py:44-64)."""
    if not boxes or not boxes_gt:
        return []
    # ties: the reference's default np.argsort is not a stable sort (its order of equal confidences depends on numpy's
    # SIMD dispatch); here equal confidences are ordered by index, on the host and in the device kernel alike
    order = list(np.argsort([b[4] for b in boxes], kind='stable'))[::-1]
    matches, used = [], []
    for idx in order:
        ious = [calculate_iou(boxes[idx], g) for g in boxes_gt]
        j = int(np.argmax(ious))
        if ious[j] >= iou_min and j not in used:
            matches.append((int(idx), j))
            used.append(j)
    return matches


def reorder_matches(matches, boxes, mode='left_right'):
    """Sort matches by the detections' left edge (utils/iou.py:87-101)."""
    assert mode == 'left_right'
    ordered = np.argsort([b[0] for b in boxes], kind='stable')
    left = [int(i) for i, _ in matches]
    return [matches[left.index(i)] for i in ordered if i in left]
