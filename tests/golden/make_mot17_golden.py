"""A realistic, non-synthetic parity stream: the public FRCNN detections of the reference's own
`assets/MOT17-mini` sequences, run through the UNMODIFIED reference trackers.

    python tests/golden/make_mot17_golden.py      # needs /root/reference; writes the files below (committed)

* `mot17_mini_dets.npz`  -- the INPUT fixture: per sequence the first frames of `det/det.txt` in the cache row format
  `(frame_id, x1, y1, x2, y2, conf, cls)` float32 (conversion of `engine/eval/cache.py:407-420`), and per detection the
  ground-truth identity it overlaps (IoU >= 0.5 with a visible pedestrian box of `gt/gt.txt`, else -1) from which the
  tests derive appearance vectors -- so embeddings are consistent along a person's track like real ReID features.
* `<tracker>_mot17_<seq>.npz` -- the reference's `(M, 8)` rows per frame (+ Kalman snapshots), as in make_golden.py.
"""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1]))

import refharness  # noqa: E402

ASSETS = refharness.REFERENCE_ROOT / "assets" / "MOT17-mini" / "train"
SEQS = {"04": ("MOT17-04-FRCNN", 220), "02": ("MOT17-02-FRCNN", 300)}


def _iou(a, b):
    x1, y1 = np.maximum(a[:, None, 0], b[None, :, 0]), np.maximum(a[:, None, 1], b[None, :, 1])
    x2, y2 = np.minimum(a[:, None, 2], b[None, :, 2]), np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    ua = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    ub = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (ua[:, None] + ub[None, :] - inter)


def build_inputs():
    out = {}
    for key, (seq, n_frames) in SEQS.items():
        raw = np.loadtxt(ASSETS / seq / "det" / "det.txt", delimiter=",")
        raw = raw[raw[:, 0] <= n_frames]
        raw = raw[np.argsort(raw[:, 0], kind="stable")]
        x, y, w, h = raw[:, 2], raw[:, 3], raw[:, 4], raw[:, 5]
        dets = np.column_stack([raw[:, 0], x, y, x + w, y + h, raw[:, 6], np.zeros(len(raw))]).astype(np.float32)
        gt = np.loadtxt(ASSETS / seq / "gt" / "gt.txt", delimiter=",")
        gt = gt[(gt[:, 6] == 1) & (gt[:, 7] == 1)]
        gid = np.full(len(dets), -1, np.int32)
        for f in np.unique(dets[:, 0]).astype(int):
            di = np.flatnonzero(dets[:, 0] == f)
            g = gt[gt[:, 0] == f]
            if not len(g):
                continue
            gb = np.column_stack([g[:, 2], g[:, 3], g[:, 2] + g[:, 4], g[:, 3] + g[:, 5]])
            iou = _iou(dets[di, 1:5].astype(np.float64), gb)
            best = iou.argmax(1)
            ok = iou[np.arange(len(di)), best] >= 0.5
            gid[di[ok]] = g[best[ok], 1].astype(np.int32)
        out[f"dets_{key}"] = dets
        out[f"gid_{key}"] = gid
        out[f"frames_{key}"] = np.int32(n_frames)
    np.savez_compressed(HERE / "mot17_mini_dets.npz", **out)
    return out


def main():
    build_inputs()
    refharness.install_reference()
    spec = importlib.util.spec_from_file_location("b200_tests_common", HERE.parent / "common.py")
    common = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(common)
    spec = importlib.util.spec_from_file_location("b200_make_golden", HERE / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)   # run(), the reference classes, the StrongSORT warp shim

    from boxmot.trackers.bbox.botsort.botsort import BotSort
    from boxmot.trackers.bbox.bytetrack import basetrack as bt_base
    from boxmot.trackers.bbox.bytetrack.bytetrack import ByteTrack
    from boxmot.trackers.bbox.deepocsort.deepocsort import DeepOcSort
    from boxmot.trackers.bbox.strongsort.strongsort import StrongSort

    img = np.zeros((1080, 1920, 3), np.uint8)
    for name, (kind, kwargs, make_frames, make_embs) in common.CASES.items():
        if "_mot17_" not in name:
            continue
        frames = make_frames()
        embs = make_embs(frames) if make_embs else None
        if kind == "bytetrack":
            bt_base.BaseTrack._count = 0
            trk = ByteTrack(**kwargs)
        elif kind == "botsort":
            trk = BotSort(reid_model=None, use_cmc=False, **kwargs)
        elif kind == "deepocsort":
            trk = DeepOcSort(reid_model=None, cmc_off=True, **kwargs)
        else:
            trk = StrongSort(reid_model=None, **kwargs)
            trk.cmc = mg._FrameWarps([np.eye(2, 3)] * len(frames))
        res = mg.run(trk, frames, img, embs)
        np.savez_compressed(HERE / f"{name}.npz", **res)
        ids = np.unique(res["rows"][:, 4]) if len(res["rows"]) else []
        print(name, "frames", len(frames), "rows", len(res["rows"]), "ids", len(ids), (HERE / f"{name}.npz").stat().st_size)


if __name__ == "__main__":
    main()
