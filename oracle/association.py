"""Oracle restatement of the association costs -- TEST INFRASTRUCTURE ONLY.

Follows (relative to /root/reference/boxmot):
  * trackers/association/iou.py:134-150       iou_batch (no epsilon in the denominator)
  * trackers/association/matching.py:46-80    iou_distance = 1 - IoU (empty -> float32 zeros of shape (T,D))
  * trackers/association/matching.py:85-107   embedding_distance = max(0, cdist(track, det, 'cosine'))
  * trackers/association/matching.py:139-147  fuse_score
  * trackers/association/matching.py:28-43    linear_assignment(cost, thresh) -> lap.lapjv(extend_cost, cost_limit)
dtype rules that matter for bit parity: track boxes are float64, detection boxes float32; numpy
computes the detection AREA in float32 and everything else after promotion to float64.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.distance import cdist

from .lap import lapjv


def iou_batch(b1: np.ndarray, b2: np.ndarray) -> np.ndarray:
    b2 = np.expand_dims(b2, 0)
    b1 = np.expand_dims(b1, 1)
    xx1 = np.maximum(b1[..., 0], b2[..., 0])
    yy1 = np.maximum(b1[..., 1], b2[..., 1])
    xx2 = np.minimum(b1[..., 2], b2[..., 2])
    yy2 = np.minimum(b1[..., 3], b2[..., 3])
    w = np.maximum(0.0, xx2 - xx1)
    h = np.maximum(0.0, yy2 - yy1)
    wh = w * h
    return wh / (
        (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
        + (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
        - wh
    )


def iou_cost(a_xyxy, b_xyxy) -> np.ndarray:
    """a_xyxy / b_xyxy: lists (or arrays) of xyxy boxes, each keeping its own dtype."""
    out = np.zeros((len(a_xyxy), len(b_xyxy)), dtype=np.float32)
    if out.size == 0:
        return out
    return 1 - iou_batch(np.asarray(a_xyxy), np.asarray(b_xyxy))


def embedding_cost(track_feats, det_feats) -> np.ndarray:
    out = np.zeros((len(track_feats), len(det_feats)), dtype=np.float32)
    if out.size == 0:
        return out
    d = np.asarray(det_feats, dtype=np.float32)
    t = np.asarray(track_feats, dtype=np.float32)
    return np.maximum(0.0, cdist(t, d, "cosine"))


def fuse_score(cost: np.ndarray, det_confs) -> np.ndarray:
    if cost.size == 0:
        return cost
    iou_sim = 1 - cost
    confs = np.array(list(det_confs))
    confs = np.expand_dims(confs, axis=0).repeat(cost.shape[0], axis=0)
    return 1 - iou_sim * confs


def linear_assignment(cost: np.ndarray, thresh: float):
    """Returns (matches list of (row, col) in ascending row order, unmatched rows, unmatched cols)."""
    if cost.size == 0:
        return [], list(range(cost.shape[0])), list(range(cost.shape[1]))
    _, x, y = lapjv(cost, extend_cost=True, cost_limit=thresh)
    matches = [(int(i), int(j)) for i, j in enumerate(x) if j >= 0]
    return matches, [int(i) for i in np.where(x < 0)[0]], [int(j) for j in np.where(y < 0)[0]]
