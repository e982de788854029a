"""Oracle restatement of OC-SORT -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/boxmot/trackers/bbox/ocsort/ocsort.py:331-547 (OcSort.__init__ / _update_impl) and its
KalmanBoxTracker :60-318 for axis-aligned boxes: the same XYSR filter (motion/kalman_filters/xysr.py), the same
`associate()` (trackers/association/association.py:61-152, called without an appearance cost), the same
observation-centric second round and emit / cull loop as DeepOCSORT (trackers/bbox/deepocsort/deepocsort.py:302-492)
with its embedding and camera-motion terms removed -- so the restatement is the DeepOCSORT oracle with
`embedding_off=True`.  Pinned row for row against goldens dumped from the unmodified OcSort class
(tests/golden/make_ocsort_golden.py, tests/test_oracle_ocsort.py).

Differences of the reference class that do not reach the output for AABB input with `use_byte=False`: ids count from 0
and are emitted as id + 1 (DeepOCSORT counts from 1 and emits the id); `min_conf` only feeds the BYTE second association;
`max_obs` sizes a history deque nothing on this path reads.  `use_byte=True` (a ByteTrack-style extra round on
low-confidence detections, ocsort.py:455-482) and OBB input are outside the restatement."""
from __future__ import annotations

from oracle.deepocsort import DeepOcSortOracle


class OcSortOracle(DeepOcSortOracle):
    def __init__(self, min_conf=0.1, delta_t=3, inertia=0.2, use_byte=False, Q_xy_scaling=0.01, Q_s_scaling=0.0001,
                 det_thresh=0.3, max_age=30, min_hits=3, iou_threshold=0.3):
        if use_byte:
            raise NotImplementedError("use_byte=True is outside the restated path")
        self.min_conf = min_conf
        super().__init__(reid_model=None, delta_t=delta_t, inertia=inertia, embedding_off=True, aw_off=True,
                         Q_xy_scaling=Q_xy_scaling, Q_s_scaling=Q_s_scaling, det_thresh=det_thresh, max_age=max_age,
                         min_hits=min_hits, iou_threshold=iou_threshold)

    def update(self, dets, img=None, embs=None, **kw):
        return super().update(dets, img, None, **kw)
