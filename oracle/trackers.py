"""Oracle restatement of the ByteTrack and BoT-SORT per-frame update -- TEST INFRASTRUCTURE ONLY.

Follows (relative to /root/reference/boxmot):
  * trackers/bbox/bytetrack/bytetrack.py:17-195   STrack (XYAH measurement, f32 detection geometry)
  * trackers/bbox/bytetrack/bytetrack.py:259-447  ByteTrack._update_impl, joint/sub/remove_duplicate
  * trackers/bbox/botsort/botsort.py:177-500      BotSort._update_impl and its five stages
  * trackers/bbox/botsort/botsort_track.py:16-115,232-282  STrack (EMA feature, class vote, predict, update)
  * trackers/bbox/botsort/botsort_utils.py:10-82  joint / sub / remove_duplicate
  * trackers/bbox/{bytetrack,botsort}/basetrack.py  id counter and state constants
  * trackers/common/detection_layout.py:48-52     det_ind column (hstack promotes f32 dets to f64)
Differences by design: the id counter is per tracker instance (the reference keeps it process-global, SURVEY
N4; goldens are generated with one fresh counter per stream), CMC is not applied, OBB and
per_class are out of scope.  Track state is held in plain records; list order is the reference's list order
because it defines output row order and duplicate removal.
"""
from __future__ import annotations

from collections import deque

import numpy as np

from . import kalman
from .association import embedding_cost, fuse_score, iou_cost, linear_assignment

NEW, TRACKED, LOST, REMOVED = 0, 1, 2, 3


def _xyxy_to_xywh32(b):
    y = np.copy(b)
    y[0] = (b[0] + b[2]) / 2
    y[1] = (b[1] + b[3]) / 2
    y[2] = b[2] - b[0]
    y[3] = b[3] - b[1]
    return y


def _xywh_to_xyxy(x):
    y = np.copy(x)
    y[0] = x[0] - x[2] / 2
    y[1] = x[1] - x[3] / 2
    y[2] = x[0] + x[2] / 2
    y[3] = x[1] + x[3] / 2
    return y


class _Rec:
    """One detection-or-track record (the reference uses one STrack object for both roles)."""

    __slots__ = ("xywh", "meas", "conf", "cls", "det_ind", "mean", "cov", "state", "activated", "id",
                 "frame_id", "start_frame", "tracklet_len", "curr_feat", "smooth_feat", "cls_hist", "kind")

    def __init__(self, det_row, kind: str):
        det = np.asarray(det_row, dtype=np.float32)
        self.kind = kind
        self.xywh = _xyxy_to_xywh32(det[0:4])
        if kind == "xyah":
            tlwh = np.copy(self.xywh)
            tlwh[0] = self.xywh[0] - self.xywh[2] / 2.0
            tlwh[1] = self.xywh[1] - self.xywh[3] / 2.0
            xyah = np.copy(tlwh)
            xyah[0] = tlwh[0] + (tlwh[2] / 2)
            xyah[1] = tlwh[1] + (tlwh[3] / 2)
            xyah[2] = tlwh[2] / tlwh[3]
            self.meas = xyah
        else:
            self.meas = self.xywh
        self.conf = det[4]
        self.cls = det[5]
        self.det_ind = det[6]
        self.mean = None
        self.cov = None
        self.state = NEW
        self.activated = False
        self.id = -1
        self.frame_id = 0
        self.start_frame = 0
        self.tracklet_len = 0
        self.curr_feat = None
        self.smooth_feat = None
        self.cls_hist = None

    # geometry ---------------------------------------------------------------------------------
    def xyxy(self):
        if self.mean is None:
            return _xywh_to_xyxy(self.xywh.copy())
        ret = self.mean[:4].copy()
        if self.kind == "xyah":
            ret[2] *= ret[3]
        return _xywh_to_xyxy(ret)

    # BoT-SORT appearance / class bookkeeping ------------------------------------------------------
    def vote_cls(self, cls, conf):
        best = 0
        seen = False
        for c in self.cls_hist:
            if cls == c[0]:
                c[1] += conf
                seen = True
            if c[1] > best:
                best = c[1]
                self.cls = c[0]
        if not seen:
            self.cls_hist.append([cls, conf])
            self.cls = cls

    def absorb_feature(self, feat, alpha=0.9):
        feat /= np.linalg.norm(feat)
        self.curr_feat = feat
        if self.smooth_feat is None:
            self.smooth_feat = feat
        else:
            self.smooth_feat = alpha * self.smooth_feat + (1 - alpha) * feat
        self.smooth_feat /= np.linalg.norm(self.smooth_feat)


def _joint(a, b):
    seen = {}
    out = []
    for t in a:
        seen[t.id] = 1
        out.append(t)
    for t in b:
        if not seen.get(t.id, 0):
            seen[t.id] = 1
            out.append(t)
    return out


def _sub(a, b):
    keep = {}
    for t in a:
        keep[t.id] = t
    for t in b:
        if t.id in keep:
            del keep[t.id]
    return list(keep.values())


def _drop_duplicates(a, b):
    pdist = iou_cost([t.xyxy() for t in a], [t.xyxy() for t in b])
    pa, pb = np.where(pdist < 0.15)
    dupa, dupb = set(), set()
    for p, q in zip(pa, pb):
        tp = a[p].frame_id - a[p].start_frame
        tq = b[q].frame_id - b[q].start_frame
        if tp > tq:
            dupb.add(int(q))
        else:
            dupa.add(int(p))
    return [t for i, t in enumerate(a) if i not in dupa], [t for i, t in enumerate(b) if i not in dupb]


class _TrackerBase:
    kind = "xyah"
    vel_zero = slice(7, 8)

    def __init__(self):
        self.frame_count = 0
        self.active = []
        self.lost = []
        self._next = 0
        self.trace = None  # optional: list of per-frame dicts for debugging / finer-grained parity

    def _new_id(self):
        self._next += 1
        return self._next

    def _with_ind(self, dets):
        dets = np.asarray(dets)
        if dets.size == 0:
            return np.empty((0, 7), dtype=np.float32)
        assert dets.ndim == 2 and dets.shape[1] == 6, "Unsupported 'dets' 2nd dimension length, valid length is 6"
        inds = np.arange(len(dets), dtype=np.int32).reshape(-1, 1)
        return np.hstack([dets, inds])

    def _predict(self, pool):
        if not pool:
            return
        mean = np.asarray([t.mean.copy() for t in pool])
        cov = np.asarray([t.cov for t in pool])
        for i, t in enumerate(pool):
            if t.state != TRACKED:
                mean[i][self.vel_zero] = 0
        mean, cov = kalman.multi_predict(self.kind, mean, cov)
        for t, m, c in zip(pool, mean, cov):
            t.mean, t.cov = m, c

    def _kf_update(self, t, det):
        t.mean, t.cov = kalman.update(self.kind, t.mean, t.cov, det.meas)

    def _activate(self, t, frame):
        t.id = self._new_id()
        t.mean, t.cov = kalman.initiate(self.kind, t.meas)
        t.tracklet_len = 0
        t.state = TRACKED
        if frame == 1:
            t.activated = True
        t.frame_id = frame
        t.start_frame = frame

    def _rows(self):
        rows = []
        for t in self.active:
            if t.activated:
                rows.append([*t.xyxy(), t.id, t.conf, t.cls, t.det_ind])
        return np.asarray(rows, dtype=np.float32) if rows else np.empty((0, 8), dtype=np.float32)

    def state_snapshot(self):
        """(id -> (mean, cov, state)) for every track still referenced by the tracker."""
        return {t.id: (t.mean.copy(), t.cov.copy(), t.state) for t in self.active + self.lost}


class ByteTrackOracle(_TrackerBase):
    kind = "xyah"
    vel_zero = slice(7, 8)

    def __init__(self, min_conf=0.1, track_thresh=0.45, match_thresh=0.8, track_buffer=25, frame_rate=30):
        super().__init__()
        self.min_conf = min_conf
        self.track_thresh = track_thresh
        self.match_thresh = match_thresh
        self.det_thresh = track_thresh
        self.max_time_lost = int(frame_rate / 30.0 * track_buffer)
        self.removed = []

    def _match_update(self, t, det, frame, activated, refind):
        if t.state == TRACKED:
            t.frame_id = frame
            t.tracklet_len += 1
            self._kf_update(t, det)
            t.state = TRACKED
            t.activated = True
            t.conf, t.cls, t.det_ind = det.conf, det.cls, det.det_ind
            activated.append(t)
        else:
            self._kf_update(t, det)
            t.tracklet_len = 0
            t.state = TRACKED
            t.activated = True
            t.frame_id = frame
            t.conf, t.cls, t.det_ind = det.conf, det.cls, det.det_ind
            refind.append(t)

    def update(self, dets, img=None, embs=None):
        dets = self._with_ind(dets)
        self.frame_count += 1
        frame = self.frame_count
        activated, refind, lost_now, removed_now = [], [], [], []
        confs = dets[:, 4] if dets.size else np.empty((0,))
        first = confs > self.track_thresh
        second = np.logical_and(confs > self.min_conf, confs < self.track_thresh)
        dets_second = dets[second]
        dets_first = dets[first]
        detections = [_Rec(d, self.kind) for d in dets_first]

        unconfirmed = [t for t in self.active if not t.activated]
        tracked = [t for t in self.active if t.activated]

        pool = _joint(tracked, self.lost)
        self._predict(pool)
        d1 = iou_cost([t.xyxy() for t in pool], [d.xyxy() for d in detections])
        d1 = fuse_score(d1, [d.conf for d in detections])
        m1, u_track, u_det = linear_assignment(d1, self.match_thresh)
        for it, idet in m1:
            self._match_update(pool[it], detections[idet], frame, activated, refind)

        detections_second = [_Rec(d, self.kind) for d in dets_second]
        r_tracked = [pool[i] for i in u_track if pool[i].state == TRACKED]
        d2 = iou_cost([t.xyxy() for t in r_tracked], [d.xyxy() for d in detections_second])
        m2, u_track2, _ = linear_assignment(d2, 0.5)
        for it, idet in m2:
            self._match_update(r_tracked[it], detections_second[idet], frame, activated, refind)
        for it in u_track2:
            t = r_tracked[it]
            if t.state != LOST:
                t.state = LOST
                lost_now.append(t)

        rest = [detections[i] for i in u_det]
        d3 = iou_cost([t.xyxy() for t in unconfirmed], [d.xyxy() for d in rest])
        d3 = fuse_score(d3, [d.conf for d in rest])
        m3, u_unc, u_det3 = linear_assignment(d3, 0.7)
        for it, idet in m3:
            t, det = unconfirmed[it], rest[idet]
            t.frame_id = frame
            t.tracklet_len += 1
            self._kf_update(t, det)
            t.state = TRACKED
            t.activated = True
            t.conf, t.cls, t.det_ind = det.conf, det.cls, det.det_ind
            activated.append(t)
        for it in u_unc:
            unconfirmed[it].state = REMOVED
            removed_now.append(unconfirmed[it])

        for inew in u_det3:
            det = rest[inew]
            if det.conf < self.det_thresh:
                continue
            self._activate(det, frame)
            activated.append(det)

        for t in self.lost:
            if frame - t.frame_id > self.max_time_lost:
                t.state = REMOVED
                removed_now.append(t)

        self.active = [t for t in self.active if t.state == TRACKED]
        self.active = _joint(self.active, activated)
        self.active = _joint(self.active, refind)
        self.lost = _sub(self.lost, self.active)
        self.lost.extend(lost_now)
        self.lost = _sub(self.lost, self.removed)
        self.removed.extend(removed_now)
        self.active, self.lost = _drop_duplicates(self.active, self.lost)
        if self.trace is not None:
            self.trace.append({"m1": m1, "m2": m2, "m3": m3, "n_pool": len(pool), "n_first": len(detections),
                               "n_second": len(detections_second), "n_unc": len(unconfirmed)})
        return self._rows()


class BotSortOracle(_TrackerBase):
    kind = "xywh"
    vel_zero = slice(6, 8)

    def __init__(self, track_high_thresh=0.5, track_low_thresh=0.1, new_track_thresh=0.6, track_buffer=30,
                 match_thresh=0.8, proximity_thresh=0.5, appearance_thresh=0.25, frame_rate=30,
                 fuse_first_associate=False, with_reid=True, second_match_thresh=0.5,
                 unconfirmed_match_thresh=0.7, unconfirmed_emb_scale=2.0, removed_stracks_buffer=100,
                 reid_model=None):
        super().__init__()
        self.track_high_thresh = track_high_thresh
        self.track_low_thresh = track_low_thresh
        self.new_track_thresh = new_track_thresh
        self.match_thresh = match_thresh
        self.max_time_lost = int(frame_rate / 30.0 * track_buffer)
        self.proximity_thresh = proximity_thresh
        self.appearance_thresh = appearance_thresh
        self.second_match_thresh = second_match_thresh
        self.unconfirmed_match_thresh = unconfirmed_match_thresh
        self.unconfirmed_emb_scale = unconfirmed_emb_scale
        self.with_reid = with_reid
        self.fuse_first_associate = fuse_first_associate
        self.removed = deque(maxlen=removed_stracks_buffer)
        self.model = reid_model if with_reid else None

    def _make_det(self, row, feat=None):
        d = _Rec(row, self.kind)
        d.cls_hist = []
        d.vote_cls(d.cls, d.conf)
        if feat is not None:
            d.absorb_feature(feat)
        return d

    def _match_update(self, t, det, frame, activated, refind):
        if t.state == TRACKED:
            t.frame_id = frame
            t.tracklet_len += 1
            self._kf_update(t, det)
            if det.curr_feat is not None:
                t.absorb_feature(det.curr_feat)
            t.state = TRACKED
            t.activated = True
            t.conf, t.cls, t.det_ind = det.conf, det.cls, det.det_ind
            t.vote_cls(det.cls, det.conf)
            activated.append(t)
        else:
            self._kf_update(t, det)
            if det.curr_feat is not None:
                t.absorb_feature(det.curr_feat)
            t.tracklet_len = 0
            t.state = TRACKED
            t.activated = True
            t.frame_id = frame
            t.conf, t.cls, t.det_ind = det.conf, det.cls, det.det_ind
            t.vote_cls(det.cls, det.conf)
            refind.append(t)

    @staticmethod
    def _gmc(tracks, H):
        """STrack.multi_gmc (botsort_track.py:117-132) for a supplied 2x3 warp."""
        if not tracks:
            return
        H = np.asarray(H, dtype=float)
        R8 = np.kron(np.eye(4), H[:2, :2])
        t = H[:2, 2]
        for st in tracks:
            mean = R8.dot(st.mean)
            mean[:2] += t
            st.mean = mean
            st.cov = R8.dot(st.cov).dot(R8.T)

    def update(self, dets, img=None, embs=None, warp=None):
        if embs is not None:
            assert len(dets) == len(embs), "Missmatch between detections and embeddings sizes"
        dets = self._with_ind(dets)
        self.frame_count += 1
        frame = self.frame_count
        activated, refind, lost_now, removed_now = [], [], [], []

        confs = dets[:, 4] if dets.size else np.empty((0,))
        second_mask = np.logical_and(confs > self.track_low_thresh, confs < self.track_high_thresh)
        first_mask = confs > self.track_high_thresh
        dets_second = dets[second_mask]
        dets_first = dets[first_mask]
        embs_first = np.asarray(embs)[first_mask] if embs is not None else None

        if self.with_reid and embs is None:
            feats = self.model.get_features(dets_first[:, 0:4], img)
        else:
            feats = embs_first if embs_first is not None else []
        if len(dets_first) > 0:
            if self.with_reid:
                detections = [self._make_det(d, np.array(f)) for d, f in zip(dets_first, feats)]
            else:
                detections = [self._make_det(d) for d in dets_first]
        else:
            detections = []

        unconfirmed = [t for t in self.active if not t.activated]
        tracked = [t for t in self.active if t.activated]
        pool = _joint(tracked, self.lost)

        # first association ---------------------------------------------------------------------
        self._predict(pool)
        if warp is not None:  # camera-motion compensation with a supplied warp (botsort.py:301)
            self._gmc(pool, warp)
            self._gmc(unconfirmed, warp)
        iou_d = iou_cost([t.xyxy() for t in pool], [d.xyxy() for d in detections])
        far = iou_d > self.proximity_thresh
        if self.fuse_first_associate:
            iou_d = fuse_score(iou_d, [d.conf for d in detections])
        if self.with_reid:
            emb_d = embedding_cost([t.smooth_feat for t in pool], [d.curr_feat for d in detections])
            emb_d[emb_d > self.appearance_thresh] = 1.0
            emb_d[far] = 1.0
            d1 = np.minimum(iou_d, emb_d)
        else:
            d1 = iou_d
        m1, u_track, u_det = linear_assignment(d1, self.match_thresh)
        for it, idet in m1:
            self._match_update(pool[it], detections[idet], frame, activated, refind)

        # second association ---------------------------------------------------------------------
        detections_second = [self._make_det(d) for d in dets_second]
        r_tracked = [pool[i] for i in u_track if pool[i].state == TRACKED]
        d2 = iou_cost([t.xyxy() for t in r_tracked], [d.xyxy() for d in detections_second])
        m2, u_track2, _ = linear_assignment(d2, self.second_match_thresh)
        for it, idet in m2:
            self._match_update(r_tracked[it], detections_second[idet], frame, activated, refind)
        for it in u_track2:
            t = r_tracked[it]
            if t.state != LOST:
                t.state = LOST
                lost_now.append(t)

        # unconfirmed ----------------------------------------------------------------------------
        rest = [detections[i] for i in u_det]
        iou_u = iou_cost([t.xyxy() for t in unconfirmed], [d.xyxy() for d in rest])
        far_u = iou_u > self.proximity_thresh
        iou_u = fuse_score(iou_u, [d.conf for d in rest])
        if self.with_reid:
            emb_u = embedding_cost([t.smooth_feat for t in unconfirmed], [d.curr_feat for d in rest])
            emb_u = emb_u / self.unconfirmed_emb_scale
            emb_u[emb_u > self.appearance_thresh] = 1.0
            emb_u[far_u] = 1.0
            d3 = np.minimum(iou_u, emb_u)
        else:
            d3 = iou_u
        m3, u_unc, u_det3 = linear_assignment(d3, self.unconfirmed_match_thresh)
        for it, idet in m3:
            t, det = unconfirmed[it], rest[idet]
            # the reference calls STrack.update (the Tracked branch) unconditionally here
            t.frame_id = frame
            t.tracklet_len += 1
            self._kf_update(t, det)
            if det.curr_feat is not None:
                t.absorb_feature(det.curr_feat)
            t.state = TRACKED
            t.activated = True
            t.conf, t.cls, t.det_ind = det.conf, det.cls, det.det_ind
            t.vote_cls(det.cls, det.conf)
            activated.append(t)
        for it in u_unc:
            unconfirmed[it].state = REMOVED
            removed_now.append(unconfirmed[it])

        # births ---------------------------------------------------------------------------------
        for inew in u_det3:
            det = rest[inew]
            if det.conf < self.new_track_thresh:
                continue
            self._activate(det, frame)
            activated.append(det)

        for t in self.lost:
            if frame - t.frame_id > self.max_time_lost:
                t.state = REMOVED
                removed_now.append(t)

        self.active = [t for t in self.active if t.state == TRACKED]
        self.active = _joint(self.active, activated)
        self.active = _joint(self.active, refind)
        self.lost = _sub(self.lost, self.active)
        self.lost.extend(lost_now)
        self.lost = _sub(self.lost, self.removed)
        self.removed.extend(removed_now)
        self.active, self.lost = _drop_duplicates(self.active, self.lost)
        if self.trace is not None:
            self.trace.append({"m1": m1, "m2": m2, "m3": m3, "n_pool": len(pool), "n_first": len(detections),
                               "n_second": len(detections_second), "n_unc": len(unconfirmed)})
        return self._rows()
