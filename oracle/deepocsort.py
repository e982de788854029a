"""Oracle restatement of DeepOCSORT's per-frame update -- TEST INFRASTRUCTURE ONLY.

Follows (relative to /root/reference/boxmot):
  * trackers/bbox/deepocsort/deepocsort.py:13-49    k_previous_obs, convert_x_to_bbox, speed_direction
  * trackers/bbox/deepocsort/deepocsort.py:51-233   KalmanBoxTracker (7-state XYSR filter, observation-centric
                                                    bookkeeping: last_observation, observations by age, velocity)
  * trackers/bbox/deepocsort/deepocsort.py:302-492  DeepOcSort._update_impl (embedding trust, first association,
                                                    OCR second round on last observations, births, emit / cull)
  * motion/kalman_filters/xysr.py:379-476           freeze / unfreeze (virtual-trajectory replay) and update
  * motion/kalman_filters/base.py:366-459           predict_state, update_state (symmetrised S, Joseph form)
  * trackers/association/association.py:8-152       speed_direction_batch, compute_aw_max_metric, associate
  * trackers/common/geometry.py:103-124             xyxy2xysr
The filter is restated as a plain record (x, P, the frozen copy, the last measurement and the gap length) instead of
the reference's deepcopy of `__dict__` plus a 50-entry history deque: within `max_age` < `max_obs` frames the only
entries the replay ever reads are the last observed measurement and the number of missed frames.
CMC is not applied (cmc_off=True); ids start at 1 per tracker instance.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

from .association import iou_batch
from .lap import lapjv

_F = np.eye(7)
_F[0, 4] = _F[1, 5] = _F[2, 6] = 1.0
_H = np.zeros((4, 7))
_H[:, :4] = np.eye(4)
_I7 = np.eye(7)


def xyxy2xysr(b):
    b = np.asarray(b, dtype=float)[0:4]
    w = b[2] - b[0]
    h = b[3] - b[1]
    return np.array([b[0] + w / 2.0, b[1] + h / 2.0, w * h, w / (h + 1e-6)]).reshape((4, 1))


def x_to_bbox(x):
    w = np.sqrt(x[2] * x[3])
    h = x[2] / w
    return np.array([x[0] - w / 2.0, x[1] - h / 2.0, x[0] + w / 2.0, x[1] + h / 2.0]).reshape((1, 4))


def speed_direction(b1, b2):
    cx1, cy1 = (b1[0] + b1[2]) / 2.0, (b1[1] + b1[3]) / 2.0
    cx2, cy2 = (b2[0] + b2[2]) / 2.0, (b2[1] + b2[3]) / 2.0
    speed = np.array([cy2 - cy1, cx2 - cx1])
    norm = np.sqrt((cy2 - cy1) ** 2 + (cx2 - cx1) ** 2) + 1e-6
    return speed / norm


SAFE_CHOLESKY_EVENTS = [0]   # diagnostics for the tests: how often the jitter path was taken


def _safe_cho_factor(matrix):
    """BaseKalmanFilter._safe_cho_factor (base.py:462-500): plain Cholesky; on failure a ridge of
    max|diag| * 10**e, e = -12 .. 3, is added until it succeeds; last resort: clip the eigenvalues."""
    try:
        return scipy.linalg.cho_factor(matrix, lower=True, check_finite=False)
    except scipy.linalg.LinAlgError:
        pass
    SAFE_CHOLESKY_EVENTS[0] += 1
    n = matrix.shape[0]
    diag = np.diagonal(matrix)
    scale = float(np.max(np.abs(diag))) if diag.size else 1.0
    if not np.isfinite(scale) or scale <= 0.0:
        scale = 1.0
    eye = np.eye(n)
    for exponent in range(-12, 4):
        jitter = scale * (10.0 ** exponent)
        try:
            return scipy.linalg.cho_factor(matrix + jitter * eye, lower=True, check_finite=False)
        except scipy.linalg.LinAlgError:
            continue
    symmetric = 0.5 * (matrix + matrix.T)
    eigvals, eigvecs = np.linalg.eigh(symmetric)
    floor = max(scale * 1e-6, 1e-12)
    eigvals = np.clip(eigvals, floor, None)
    repaired = (eigvecs * eigvals) @ eigvecs.T
    repaired = 0.5 * (repaired + repaired.T)
    return scipy.linalg.cho_factor(repaired, lower=True, check_finite=False)


class XYSRFilter:
    """KalmanFilterXYSR(dim_x=7, dim_z=4) as configured by KalmanBoxTracker (deepocsort.py:82-114)."""

    def __init__(self, z0, q_xy=0.01, q_s=0.0001):
        self.x = np.zeros((7, 1))
        self.P = np.eye(7)
        self.Q = np.eye(7)
        self.R = np.eye(4)
        self.R[2:, 2:] *= 10.0
        self.P[4:, 4:] *= 1000.0
        self.P *= 10.0
        self.Q[4:6, 4:6] *= q_xy
        self.Q[-1, -1] *= q_s
        self.x[:4] = z0
        self.observed = False
        self.saved = None      # (x, P, z_last) frozen when the track stops being observed
        self.z_last = None     # last observed (prepared) measurement
        self.gap = 0           # consecutive missed frames since the freeze

    def _enforce(self):
        self.x[2, 0] = max(float(self.x[2, 0]), 1e-6)
        self.x[3, 0] = max(float(self.x[3, 0]), 1e-6)
        self.P = 0.5 * (self.P + self.P.T)

    def predict(self):
        self.x = np.dot(_F, self.x)
        self.P = 1.0 * np.dot(np.dot(_F, self.P), _F.T) + self.Q
        self._enforce()

    def _update_state(self, z):
        pm = np.dot(_H, self.x)
        pc = np.dot(np.dot(_H, self.P), _H.T) + self.R
        pc = 0.5 * (pc + pc.T)
        chol = _safe_cho_factor(pc)
        K = scipy.linalg.cho_solve(chol, np.dot(self.P, _H.T).T, check_finite=False).T
        y = z - pm
        self.x = self.x + np.dot(K, y)
        i_kh = _I7 - np.dot(K, _H)
        self.P = np.linalg.multi_dot((i_kh, self.P, i_kh.T)) + np.linalg.multi_dot((K, self.R, K.T))
        self.P = 0.5 * (self.P + self.P.T)

    def apply_affine(self, m, t):
        """KalmanFilterXYSR.apply_affine_correction (xysr.py:311-366), AABB: position and velocity through the 2x2 part,
        their covariance blocks likewise; the frozen copy too while the track is unobserved.  The pre-gap measurement
        the un-freeze replay starts from is the LIVE history entry, which the reference does not warp."""
        def one(x, P):
            x[:2] = m @ x[:2] + t
            x[4:6] = m @ x[4:6]
            P[:2, :2] = m @ P[:2, :2] @ m.T
            P[4:6, 4:6] = m @ P[4:6, 4:6] @ m.T

        one(self.x, self.P)
        if not self.observed and self.saved is not None:
            one(self.saved[0], self.saved[1])
        self._enforce()

    @staticmethod
    def _prepare(z):
        m = np.asarray(z, dtype=float).reshape((4, 1)).copy()
        m[2, 0] = max(float(m[2, 0]), 1e-6)
        m[3, 0] = max(float(m[3, 0]), 1e-6)
        return m

    def _observe(self, m):
        self._update_state(m)
        self._enforce()
        self.z_last = m.copy()

    def _unfreeze(self, m_new):
        if self.saved is None:
            return
        self.x, self.P, z_a = self.saved[0].copy(), self.saved[1].copy(), self.saved[2]
        self.saved = None
        x1, y1, s1, r1 = z_a[:4, 0]
        w1, h1 = np.sqrt(s1 * r1), np.sqrt(s1 / r1)
        x2, y2, s2, r2 = m_new[:4, 0]
        w2, h2 = np.sqrt(s2 * r2), np.sqrt(s2 / r2)
        gap = self.gap + 1
        dx, dy = (x2 - x1) / gap, (y2 - y1) / gap
        dw, dh = (w2 - w1) / gap, (h2 - h1) / gap
        for i in range(gap):
            x = x1 + (i + 1) * dx
            y = y1 + (i + 1) * dy
            w = w1 + (i + 1) * dw
            h = h1 + (i + 1) * dh
            box = np.array([x, y, w * h, w / float(h)], dtype=float).reshape((4, 1))
            self._observe(self._prepare(box))
            if i != gap - 1:
                self.predict()

    def update(self, z):
        if z is None:
            if self.observed:
                self.saved = (self.x.copy(), self.P.copy(), self.z_last.copy())
                self.gap = 0
            self.gap += 1
            self.observed = False
            return
        m = self._prepare(z)
        if not self.observed:
            self._unfreeze(m)
        self.observed = True
        self.gap = 0
        self._observe(m)


class _Track:
    def __init__(self, det, tid, delta_t, emb, q_xy, q_s):
        self.conf, self.cls, self.det_ind = det[4], det[5], det[6]
        self.kf = XYSRFilter(xyxy2xysr(det[0:5]), q_xy, q_s)
        self.time_since_update = 0
        self.id = tid
        self.hits = 0
        self.hit_streak = 0
        self.age = 0
        self.last_observation = np.array([-1, -1, -1, -1, -1])
        self.observations = {}
        self.velocity = None
        self.delta_t = delta_t
        self.emb = emb

    def update(self, det):
        if det is not None:
            bbox = det[0:5]
            self.conf, self.cls, self.det_ind = det[4], det[5], det[6]
            if self.last_observation.sum() >= 0:
                prev = None
                for dt in range(self.delta_t, 0, -1):
                    if self.age - dt in self.observations:
                        prev = self.observations[self.age - dt]
                        break
                if prev is None:
                    prev = self.last_observation
                self.velocity = speed_direction(prev, bbox)
            self.last_observation = bbox
            self.observations[self.age] = bbox
            self.time_since_update = 0
            self.hits += 1
            self.hit_streak += 1
            self.kf.update(xyxy2xysr(bbox))
        else:
            self.kf.update(None)

    def apply_affine_correction(self, affine):
        """KalmanBoxTracker.apply_affine_correction (deepocsort.py:189-206), statement for statement: the last
        observation and the observations inside the velocity window are views of the same detection rows, so the
        newest one is warped by both loops -- as in the reference."""
        affine = np.asarray(affine, dtype=float)
        m = affine[:, :2]
        t = affine[:, 2].reshape(2, 1)
        if self.last_observation.sum() > 0:
            ps = self.last_observation[:4].reshape(2, 2).T
            ps = m @ ps + t
            self.last_observation[:4] = ps.T.reshape(-1)
        for dt in range(self.delta_t, -1, -1):
            if self.age - dt in self.observations:
                ps = self.observations[self.age - dt][:4].reshape(2, 2).T
                ps = m @ ps + t
                self.observations[self.age - dt][:4] = ps.T.reshape(-1)
        self.kf.apply_affine(m, t)

    def update_emb(self, emb, alpha=0.9):
        self.emb = alpha * self.emb + (1 - alpha) * emb
        self.emb /= np.linalg.norm(self.emb)

    def predict(self):
        if (self.kf.x[6] + self.kf.x[2]) <= 0:
            self.kf.x[6] *= 0.0
        self.kf.predict()
        self.age += 1
        if self.time_since_update > 0:
            self.hit_streak = 0
        self.time_since_update += 1
        return x_to_bbox(self.kf.x)

    def k_previous_obs(self, k):
        if len(self.observations) == 0:
            return [-1, -1, -1, -1, -1]
        for i in range(k):
            dt = k - i
            if self.age - dt in self.observations:
                return self.observations[self.age - dt]
        return self.observations[max(self.observations.keys())]


def speed_direction_batch(dets, tracks):
    tracks = tracks[..., np.newaxis]
    cx1, cy1 = (dets[:, 0] + dets[:, 2]) / 2.0, (dets[:, 1] + dets[:, 3]) / 2.0
    cx2, cy2 = (tracks[:, 0] + tracks[:, 2]) / 2.0, (tracks[:, 1] + tracks[:, 3]) / 2.0
    dx = cx1 - cx2
    dy = cy1 - cy2
    norm = np.sqrt(dx ** 2 + dy ** 2) + 1e-6
    return dy / norm, dx / norm


def linear_assignment(cost):
    _, x, y = lapjv(cost, extend_cost=True)
    return np.array([[y[i], i] for i in x if i >= 0])


def aw_max_metric(emb_cost, w_emb0, bottom=0.5):
    w_emb = np.full_like(emb_cost, w_emb0)
    for i in range(emb_cost.shape[0]):
        inds = np.argsort(-emb_cost[i])
        if len(inds) < 2:
            continue
        if emb_cost[i, inds[0]] == 0:
            rw = 0
        else:
            rw = 1 - max((emb_cost[i, inds[1]] / emb_cost[i, inds[0]]) - bottom, 0) / (1 - bottom)
        w_emb[i] *= rw
    for j in range(emb_cost.shape[1]):
        inds = np.argsort(-emb_cost[:, j])
        if len(inds) < 2:
            continue
        if emb_cost[inds[0], j] == 0:
            cw = 0
        else:
            cw = 1 - max((emb_cost[inds[1], j] / emb_cost[inds[0], j]) - bottom, 0) / (1 - bottom)
        w_emb[:, j] *= cw
    return w_emb * emb_cost


def associate(dets, trks, iou_threshold, velocities, prev_obs, vdc_weight, emb_cost, w_emb, aw_off, aw_param):
    if len(trks) == 0:
        return np.empty((0, 2), dtype=int), np.arange(len(dets)), np.empty((0, 5), dtype=int)
    Y, X = speed_direction_batch(dets, prev_obs)
    iy, ix = velocities[:, 0], velocities[:, 1]
    iy = np.repeat(iy[:, np.newaxis], Y.shape[1], axis=1)
    ix = np.repeat(ix[:, np.newaxis], X.shape[1], axis=1)
    cosang = np.clip(ix * X + iy * Y, a_min=-1, a_max=1)
    ang = (np.pi / 2.0 - np.abs(np.arccos(cosang))) / np.pi
    valid = np.ones(prev_obs.shape[0])
    valid[np.where(prev_obs[:, 4] < 0)] = 0
    iou = iou_batch(dets, trks)
    scores = np.repeat(dets[:, -1][:, np.newaxis], trks.shape[0], axis=1)
    valid = np.repeat(valid[:, np.newaxis], X.shape[1], axis=1)
    ang_cost = ((valid * ang) * vdc_weight).T * scores
    if min(iou.shape):
        a = (iou > iou_threshold).astype(np.int32)
        if a.sum(1).max() == 1 and a.sum(0).max() == 1:
            matched = np.stack(np.where(a), axis=1)
        else:
            if emb_cost is None:
                emb_cost = 0
            else:
                emb_cost[iou <= 0] = 0
                if not aw_off:
                    emb_cost = aw_max_metric(emb_cost, w_emb, bottom=aw_param)
                else:
                    emb_cost *= w_emb
            matched = linear_assignment(-(iou + ang_cost + emb_cost))
            if matched.size == 0:
                matched = np.empty(shape=(0, 2))
    else:
        matched = np.empty(shape=(0, 2))
    un_d = [d for d in range(len(dets)) if d not in matched[:, 0]]
    un_t = [t for t in range(len(trks)) if t not in matched[:, 1]]
    matches = []
    for m in matched:
        if iou[m[0], m[1]] < iou_threshold:
            un_d.append(m[0])
            un_t.append(m[1])
        else:
            matches.append(m.reshape(1, 2))
    matches = np.concatenate(matches, axis=0) if matches else np.empty((0, 2), dtype=int)
    return matches, np.array(un_d), np.array(un_t)


class DeepOcSortOracle:
    def __init__(self, reid_model=None, delta_t=3, inertia=0.2, w_association_emb=0.5, alpha_fixed_emb=0.95,
                 aw_param=0.5, embedding_off=False, aw_off=False, Q_xy_scaling=0.01, Q_s_scaling=0.0001,
                 det_thresh=0.3, max_age=30, min_hits=3, iou_threshold=0.3):
        self.delta_t, self.inertia, self.w_emb, self.af = delta_t, inertia, w_association_emb, alpha_fixed_emb
        self.aw_param, self.embedding_off, self.aw_off = aw_param, embedding_off, aw_off
        self.q_xy, self.q_s = Q_xy_scaling, Q_s_scaling
        self.det_thresh, self.max_age, self.min_hits, self.iou_threshold = det_thresh, max_age, min_hits, iou_threshold
        self.model = reid_model
        self.tracks = []
        self.frame_count = 0
        self._next = 1

    def update(self, dets, img=None, embs=None, warp=None):
        """`warp`: the 2x3 matrix the reference's CMC estimator would return for this frame (deepocsort.py:345-348);
        estimation itself is outside the path."""
        dets = np.asarray(dets)
        if dets.size == 0:
            dets = np.empty((0, 6), dtype=np.float32)
        assert dets.ndim == 2 and dets.shape[1] == 6
        if embs is not None:
            assert len(dets) == len(embs), "Missmatch between detections and embeddings sizes"
        self.frame_count += 1
        scores = dets[:, 4]
        dets = np.hstack([dets, np.arange(len(dets)).reshape(-1, 1)])
        keep = scores > self.det_thresh
        dets = dets[keep]
        if self.embedding_off or dets.shape[0] == 0:
            dets_embs = np.ones((dets.shape[0], 1))
        elif embs is not None:
            dets_embs = np.asarray(embs)[keep]
        else:
            dets_embs = self.model.get_features(dets[:, 0:4], img)
        if warp is not None:
            for trk in self.tracks:
                trk.apply_affine_correction(warp)
        trust = (dets[:, 4] - self.det_thresh) / (1 - self.det_thresh)
        dets_alpha = self.af + (1 - self.af) * (1 - trust)
        trks = np.zeros((len(self.tracks), 5))
        trk_embs, to_del = [], []
        for t, trk in enumerate(trks):
            pos = self.tracks[t].predict()[0]
            trk[:] = [pos[0], pos[1], pos[2], pos[3], 0]
            if np.any(np.isnan(pos)):
                to_del.append(t)
            else:
                trk_embs.append(self.tracks[t].emb)
        trks = np.ma.compress_rows(np.ma.masked_invalid(trks))
        trk_embs = np.vstack(trk_embs) if len(trk_embs) > 0 else np.array(trk_embs)
        for t in reversed(to_del):
            self.tracks.pop(t)
        velocities = np.array([t.velocity if t.velocity is not None else np.array((0, 0)) for t in self.tracks])
        last_boxes = np.array([t.last_observation for t in self.tracks])
        k_obs = np.array([t.k_previous_obs(self.delta_t) for t in self.tracks])
        if self.embedding_off or dets.shape[0] == 0 or trk_embs.shape[0] == 0:
            emb_cost = None
        else:
            emb_cost = dets_embs @ trk_embs.T
        matched, un_d, un_t = associate(dets[:, 0:5], trks, self.iou_threshold, velocities, k_obs, self.inertia,
                                        emb_cost, self.w_emb, self.aw_off, self.aw_param)
        for m in matched:
            self.tracks[m[1]].update(dets[m[0], :])
            self.tracks[m[1]].update_emb(dets_embs[m[0]], alpha=dets_alpha[m[0]])
        if un_d.shape[0] > 0 and un_t.shape[0] > 0:
            left_dets = dets[un_d]
            left_trks = last_boxes[un_t]
            iou_left = np.array(iou_batch(left_dets, left_trks))
            if iou_left.max() > self.iou_threshold:
                rem = linear_assignment(-iou_left)
                rm_d, rm_t = [], []
                for m in rem:
                    di, ti = un_d[m[0]], un_t[m[1]]
                    if iou_left[m[0], m[1]] < self.iou_threshold:
                        continue
                    self.tracks[ti].update(dets[di, :])
                    self.tracks[ti].update_emb(dets_embs[di], alpha=dets_alpha[di])
                    rm_d.append(di)
                    rm_t.append(ti)
                un_d = np.setdiff1d(un_d, np.array(rm_d))
                un_t = np.setdiff1d(un_t, np.array(rm_t))
        for m in un_t:
            self.tracks[m].update(None)
        for i in un_d:
            self.tracks.append(_Track(dets[i], self._next, self.delta_t, dets_embs[i], self.q_xy, self.q_s))
            self._next += 1
        ret = []
        i = len(self.tracks)
        for trk in reversed(self.tracks):
            if trk.last_observation.sum() < 0:
                d = x_to_bbox(trk.kf.x)[0]
            else:
                d = trk.last_observation[:4]
            if trk.time_since_update < 1 and (trk.hit_streak >= self.min_hits or self.frame_count <= self.min_hits):
                ret.append(np.concatenate((d, [trk.id], [trk.conf], [trk.cls], [trk.det_ind])).reshape(1, -1))
            i -= 1
            if trk.time_since_update > self.max_age:
                self.tracks.pop(i)
        if len(ret) > 0:
            return np.concatenate(ret).astype(np.float32)
        return np.empty((0, 8), dtype=np.float32)

    def state_snapshot(self):
        return {t.id: (t.kf.x[:, 0].copy(), t.kf.P.copy()) for t in self.tracks}
