"""Oracle restatement of StrongSORT's per-frame update -- TEST INFRASTRUCTURE ONLY.

Follows (relative to /root/reference/boxmot):
  * trackers/bbox/strongsort/strongsort.py:69-123          StrongSort._update_impl (conf >= min_conf filter, camera
                                                           update whenever tracks exist, predict, update, emit)
  * trackers/bbox/strongsort/sort/tracker.py:62-169        Tracker.predict/update/_match/_initiate_track (appearance
                                                           stage over confirmed tracks, IoU stage over tentative +
                                                           just-missed tracks, births, gallery refit)
  * trackers/bbox/strongsort/sort/track.py:66-196          Track (n_init confirmation, EMA appearance, NSA Kalman
                                                           update with the detection confidence, camera_update,
                                                           mark_missed)
  * trackers/bbox/strongsort/sort/linear_assignment.py:12-70   min_cost_matching (clip to max_distance + 1e-5,
                                                           scipy linear_sum_assignment, unmatched ordering)
  * trackers/bbox/strongsort/sort/linear_assignment.py:73-110  matching_cascade (one level; unmatched tracks come
                                                           back through a Python set)
  * trackers/bbox/strongsort/sort/linear_assignment.py:145-198 gate_cost_matrix (squared Mahalanobis gate at
                                                           chi2inv95[4], mc_lambda blend, INFTY 1e5)
  * trackers/bbox/strongsort/sort/linear_assignment.py:201-221,266-283,304-344  NN cosine metric + sample gallery
  * trackers/bbox/strongsort/sort/iou_matching.py:9-87     tlwh IoU cost (INFTY rows for time_since_update > 1)
  * motion/kalman_filters/base.py:253-268,286-355,523-551  single-track predict, project/update, gating_distance
  * motion/kalman_filters/xyah.py:22-68                    XYAH std tables
Third-party arithmetic: `scipy.optimize.linear_sum_assignment` (scipy 1.18.1 here and on the GPU box; a modified
Jonker-Volgenant shortest-augmenting-path solver, Crouse 2016).  The oracle calls scipy itself, exactly like the
reference; the DEVICE solver restates it tie-for-tie (boxmot_b200/csrc/lsa_sap.cuh) because min_cost_matching's
clipped entries tie by construction and scipy's choice among them orders the unmatched detections and so the ids.
Two orderings come from CPython's `set` and are kept by using `set` the same way (tracker.py:156,
linear_assignment.py:108).

Track state lives in flat records, the gallery is a list per id like the reference's; the camera warp is an input
(its estimation is out of scope) but camera_update is applied whenever tracks exist, identity included.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
from scipy.optimize import linear_sum_assignment

INFTY_COST = 1e5
CHI2_4 = 9.4877
W_POS = 1.0 / 20
W_VEL = 1.0 / 160

_F = np.eye(8)
for _i in range(4):
    _F[_i, 4 + _i] = 1.0
_H = np.eye(4, 8)

TENTATIVE, CONFIRMED, DELETED = 1, 2, 3


def _clamp(mean):
    mean[2] = max(float(mean[2]), 1e-4)
    mean[3] = max(float(mean[3]), 1e-4)
    return mean


def kf_initiate(z):
    z = np.asarray(z, dtype=float).copy()
    mean = np.r_[z, np.zeros_like(z)]
    std = [2 * W_POS * z[3], 2 * W_POS * z[3], 1e-2, 2 * W_POS * z[3],
           10 * W_VEL * z[3], 10 * W_VEL * z[3], 1e-5, 10 * W_VEL * z[3]]
    return _clamp(mean), np.diag(np.square(std))


def kf_predict(mean, cov):
    std_pos = [W_POS * mean[3], W_POS * mean[3], 1e-2, W_POS * mean[3]]
    std_vel = [W_VEL * mean[3], W_VEL * mean[3], 1e-5, W_VEL * mean[3]]
    motion_cov = np.diag(np.square(np.r_[std_pos, std_vel]))
    mean = np.dot(mean, _F.T)
    cov = np.linalg.multi_dot((_F, cov, _F.T)) + motion_cov
    return _clamp(mean), cov


def kf_project(mean, cov, confidence=0.0):
    std = [W_POS * mean[3], W_POS * mean[3], 1e-1, W_POS * mean[3]]
    std = [(1 - confidence) * x for x in std]
    return np.dot(_H, mean), np.linalg.multi_dot((_H, cov, _H.T)) + np.diag(np.square(std))


def kf_update(mean, cov, z, confidence):
    pm, pc = kf_project(mean, cov, confidence)
    chol, lower = scipy.linalg.cho_factor(pc, lower=True, check_finite=False)
    gain = scipy.linalg.cho_solve((chol, lower), np.dot(cov, _H.T).T, check_finite=False).T
    new_mean = mean + np.dot(z - pm, gain.T)
    new_cov = cov - np.linalg.multi_dot((gain, pc, gain.T))
    return _clamp(new_mean), new_cov


def gating_distance(mean, cov, measurements):
    pm, pc = kf_project(mean, cov)
    d = measurements - pm
    chol = np.linalg.cholesky(pc)
    z = scipy.linalg.solve_triangular(chol, d.T, lower=True, check_finite=False, overwrite_b=True)
    return np.sum(z * z, axis=0)


def tlwh_to_xyah(tlwh):
    r = tlwh.copy()
    r[:2] += r[2:] / 2
    r[2] /= r[3]
    return r


def tlwh_iou(box, cand):
    tl0, br0 = box[:2], box[:2] + box[2:]
    ctl, cbr = cand[:, :2], cand[:, :2] + cand[:, 2:]
    tl = np.c_[np.maximum(tl0[0], ctl[:, 0])[:, None], np.maximum(tl0[1], ctl[:, 1])[:, None]]
    br = np.c_[np.minimum(br0[0], cbr[:, 0])[:, None], np.minimum(br0[1], cbr[:, 1])[:, None]]
    wh = np.maximum(0.0, br - tl)
    inter = wh.prod(axis=1)
    return inter / (box[2:].prod() + cand[:, 2:].prod(axis=1) - inter)


class _Det:
    __slots__ = ("tlwh", "conf", "cls", "ind", "feat")

    def __init__(self, tlwh, conf, cls, ind, feat):
        self.tlwh, self.conf, self.cls, self.ind, self.feat = tlwh, conf, cls, ind, feat


class _Track:
    def __init__(self, det, tid, n_init, max_age, alpha):
        self.id, self.conf, self.cls, self.ind = tid, det.conf, det.cls, det.ind
        self.hits, self.age, self.tsu, self.state = 1, 1, 0, TENTATIVE
        self.n_init, self.max_age, self.alpha = n_init, max_age, alpha
        det.feat /= np.linalg.norm(det.feat)
        self.feat = det.feat
        self.mean, self.cov = kf_initiate(tlwh_to_xyah(det.tlwh))

    def tlwh(self):
        r = self.mean[:4].copy()
        r[2] *= r[3]
        r[:2] -= r[2:] / 2
        return r

    def tlbr(self):
        r = self.tlwh()
        r[2:] = r[:2] + r[2:]
        return r

    def camera_update(self, warp):
        a, b = warp
        m = np.array([a, b, [0, 0, 1]]).tolist()
        x1, y1, x2, y2 = self.tlbr()
        x1_, y1_, _ = m @ np.array([x1, y1, 1]).T
        x2_, y2_, _ = m @ np.array([x2, y2, 1]).T
        w, h = x2_ - x1_, y2_ - y1_
        self.mean[:4] = [x1_ + w / 2, y1_ + h / 2, w / h, h]

    def predict(self):
        self.mean, self.cov = kf_predict(self.mean, self.cov)
        self.age += 1
        self.tsu += 1

    def update(self, det):
        self.conf, self.cls, self.ind = det.conf, det.cls, det.ind
        self.mean, self.cov = kf_update(self.mean, self.cov, tlwh_to_xyah(det.tlwh), det.conf)
        f = det.feat / np.linalg.norm(det.feat)
        s = self.alpha * self.feat + (1 - self.alpha) * f
        s /= np.linalg.norm(s)
        self.feat = s
        self.hits += 1
        self.tsu = 0
        if self.state == TENTATIVE and self.hits >= self.n_init:
            self.state = CONFIRMED

    def mark_missed(self):
        if self.state == TENTATIVE or self.tsu > self.max_age:
            self.state = DELETED


def min_cost_matching(cost, max_distance, track_idx, det_idx):
    """linear_assignment.py:12-70 with the cost matrix already built by the caller."""
    if len(det_idx) == 0 or len(track_idx) == 0:
        return [], list(track_idx), list(det_idx)
    cost[cost > max_distance] = max_distance + 1e-5
    rows, cols = linear_sum_assignment(cost)
    matches, un_t, un_d = [], [], []
    for c, d in enumerate(det_idx):
        if c not in cols:
            un_d.append(d)
    for r, t in enumerate(track_idx):
        if r not in rows:
            un_t.append(t)
    for r, c in zip(rows, cols):
        if cost[r, c] > max_distance:
            un_t.append(track_idx[r])
            un_d.append(det_idx[c])
        else:
            matches.append((track_idx[r], det_idx[c]))
    return matches, un_t, un_d


class StrongSortOracle:
    def __init__(self, reid_model=None, min_conf=0.1, max_cos_dist=0.2, max_iou_dist=0.7, n_init=3, nn_budget=100,
                 mc_lambda=0.98, ema_alpha=0.9, max_age=30):
        self.model = reid_model
        self.min_conf, self.max_cos_dist, self.max_iou_dist = min_conf, max_cos_dist, max_iou_dist
        self.n_init, self.budget, self.mc_lambda, self.alpha, self.max_age = n_init, nn_budget, mc_lambda, ema_alpha, max_age
        self.tracks: list[_Track] = []
        self.samples: dict[int, list] = {}
        self._next_id = 1
        self.frame_count = 0
        self.trace = None

    # -- appearance stage cost: NN cosine over each track's gallery, then the motion gate ---------------------
    def _appearance_cost(self, dets, t_idx, d_idx):
        feats = np.array([dets[i].feat for i in d_idx])
        cost = np.zeros((len(t_idx), len(d_idx)))
        b = np.asarray(feats) / np.linalg.norm(feats, axis=1, keepdims=True)
        for r, ti in enumerate(t_idx):
            a = np.asarray(self.samples[self.tracks[ti].id])
            a = a / np.linalg.norm(a, axis=1, keepdims=True)
            cost[r, :] = (1.0 - np.dot(a, b.T)).min(axis=0)
        meas = np.asarray([tlwh_to_xyah(dets[i].tlwh) for i in d_idx])
        for r, ti in enumerate(t_idx):
            t = self.tracks[ti]
            g = gating_distance(t.mean, t.cov, meas)
            cost[r, g > CHI2_4] = INFTY_COST
            cost[r] = self.mc_lambda * cost[r] + (1 - self.mc_lambda) * g
        return cost

    def _iou_cost(self, dets, t_idx, d_idx):
        cost = np.zeros((len(t_idx), len(d_idx)))
        cand = np.asarray([dets[i].tlwh for i in d_idx])
        for r, ti in enumerate(t_idx):
            t = self.tracks[ti]
            if t.tsu > 1:
                cost[r, :] = INFTY_COST
                continue
            cost[r, :] = 1.0 - tlwh_iou(t.tlwh(), cand)
        return cost

    def _match(self, dets):
        confirmed = [i for i, t in enumerate(self.tracks) if t.state == CONFIRMED]
        unconfirmed = [i for i, t in enumerate(self.tracks) if t.state != CONFIRMED]
        d_all = list(range(len(dets)))
        if len(d_all) == 0 or len(confirmed) == 0:
            m_a, un_d = [], d_all
        else:
            m_a, _, un_d = min_cost_matching(self._appearance_cost(dets, confirmed, d_all), self.max_cos_dist,
                                             confirmed, d_all)
        un_t_a = list(set(confirmed) - set(k for k, _ in m_a))          # CPython set order, as the reference
        cand = unconfirmed + [k for k in un_t_a if self.tracks[k].tsu == 1]
        un_t_a = [k for k in un_t_a if self.tracks[k].tsu != 1]
        if len(un_d) == 0 or len(cand) == 0:
            m_b, un_t_b = [], cand
        else:
            m_b, un_t_b, un_d = min_cost_matching(self._iou_cost(dets, cand, un_d), self.max_iou_dist, cand, un_d)
        return m_a + m_b, list(set(un_t_a + un_t_b)), un_d

    def update(self, dets, img=None, embs=None, warp=None):
        self.frame_count += 1
        dets = np.asarray(dets)
        if dets.size == 0:
            dets = np.empty((0, 7), dtype=np.float32)
        else:
            assert dets.ndim == 2 and dets.shape[1] == 6, "Unsupported 'dets' 2nd dimension length, valid length is 6"
            dets = np.hstack([dets, np.arange(len(dets), dtype=np.int32).reshape(-1, 1)])
        keep = dets[:, 4] >= self.min_conf
        dets = dets[keep]
        xyxy = dets[:, 0:4]
        if len(self.tracks) >= 1:
            w = np.eye(2, 3) if warp is None else np.asarray(warp)
            for t in self.tracks:
                t.camera_update(w)
        if embs is not None:
            feats = np.asarray(embs)[keep]
        else:
            feats = self.model.get_features(xyxy, img)
        tlwh = np.copy(xyxy)
        tlwh[:, 2] = xyxy[:, 2] - xyxy[:, 0]
        tlwh[:, 3] = xyxy[:, 3] - xyxy[:, 1]
        D = [_Det(tlwh[i], dets[i, 4], dets[i, 5], dets[i, 6], feats[i]) for i in range(len(dets))]
        for t in self.tracks:
            t.predict()
        matches, un_t, un_d = self._match(D)
        if self.trace is not None:
            self.trace.append(dict(matches=list(matches), unmatched_tracks=list(un_t), unmatched_dets=list(un_d)))
        for ti, di in matches:
            self.tracks[ti].update(D[di])
        for ti in un_t:
            self.tracks[ti].mark_missed()
        for di in un_d:
            self.tracks.append(_Track(D[di], self._next_id, self.n_init, self.max_age, self.alpha))
            self._next_id += 1
        self.tracks = [t for t in self.tracks if t.state != DELETED]
        active = [t.id for t in self.tracks if t.state == CONFIRMED]
        for t in self.tracks:
            if t.state != CONFIRMED:
                continue
            self.samples.setdefault(t.id, []).append(t.feat)
            if self.budget is not None:
                self.samples[t.id] = self.samples[t.id][-self.budget:]
        self.samples = {k: self.samples[k] for k in active}
        rows = []
        for t in self.tracks:
            if t.state != CONFIRMED or t.tsu >= 1:
                continue
            rows.append(np.concatenate((t.tlbr(), [t.id], [t.conf], [t.cls], [t.ind])).reshape(1, -1))
        return np.concatenate(rows) if rows else np.empty((0, 8), dtype=np.float32)

    def state_snapshot(self):
        ids = np.asarray([t.id for t in self.tracks], dtype=np.int64)
        means = np.asarray([t.mean for t in self.tracks], dtype=np.float64).reshape(-1, 8)
        covs = np.asarray([t.cov for t in self.tracks], dtype=np.float64).reshape(-1, 8, 8)
        return ids, means, covs
