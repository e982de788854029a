"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) in this container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Inputs are regenerated from seeds by oracle/streams.py at test time; only the reference OUTPUTS are stored
(per-frame `(M,8)` rows, plus Kalman mean/cov snapshots keyed by track id on a few frames).  The reference is
imported through tests/golden/refharness.py (stubs for gdown/ftfy/yacs; `lap` -> oracle/lap.py because lapx is
not installed -- SURVEY section 8(c)); CMC is disabled; the ByteTrack id counter is reset per
stream.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1]))

import refharness  # noqa: E402

refharness.install_reference()

from boxmot.trackers.bbox.botsort.botsort import BotSort  # noqa: E402
from boxmot.trackers.bbox.bytetrack import basetrack as bt_base  # noqa: E402
from boxmot.trackers.bbox.bytetrack.bytetrack import ByteTrack  # noqa: E402

from oracle.streams import bench_stream, stress_embeddings, stress_stream  # noqa: E402

BYTETRACK_YAML = dict(min_conf=0.1, track_thresh=0.6, track_buffer=30, match_thresh=0.9, frame_rate=30)
BOTSORT_YAML = dict(
    track_high_thresh=0.6296854875023994, track_low_thresh=0.1014392537025336,
    new_track_thresh=0.6246494191492591, track_buffer=40, match_thresh=0.7722224024589055,
    proximity_thresh=0.6084297894561342, appearance_thresh=0.6188818853936099,
    unconfirmed_emb_scale=2.5445206391993294, second_match_thresh=0.28795081514328974,
    unconfirmed_match_thresh=0.41148010638233784, removed_stracks_buffer=329, fuse_first_associate=True,
    frame_rate=30, with_reid=True)
SNAP_FRAMES = (1, 2, 10, 50, 150, 299)


def _snapshot_docs(tracker):
    ids, means, covs = [], [], []
    for t in tracker.active_tracks:
        ids.append(t.id)
        means.append(np.r_[np.asarray(t.kf.x, dtype=np.float64).reshape(-1), 0.0])
        c = np.zeros((8, 8))
        c[:7, :7] = t.kf.P
        covs.append(c)
    return (np.asarray(ids, dtype=np.int64), np.asarray(means).reshape(-1, 8), np.asarray(covs).reshape(-1, 8, 8))


def _snapshot(tracker):
    ids, means, covs = [], [], []
    for t in list(tracker.active_tracks) + list(tracker.lost_stracks):
        ids.append(t.id)
        means.append(np.asarray(t.mean, dtype=np.float64))
        covs.append(np.asarray(t.covariance, dtype=np.float64))
    return (np.asarray(ids, dtype=np.int64), np.asarray(means).reshape(-1, 8), np.asarray(covs).reshape(-1, 8, 8))


class _GivenWarps:
    """Stands in for the reference's CMC estimator (motion/cmc/*, OpenCV): returns the supplied warp."""

    def __init__(self, warps):
        self.warps, self.i = warps, 0

    def apply(self, img, dets):
        w = self.warps[self.i]
        self.i += 1
        return w


def _snapshot_ss(tracker):
    tr = tracker.tracker.tracks
    return (np.asarray([t.id for t in tr], dtype=np.int64),
            np.asarray([t.mean for t in tr], dtype=np.float64).reshape(-1, 8),
            np.asarray([t.covariance for t in tr], dtype=np.float64).reshape(-1, 8, 8))


class _FrameWarps:
    """StrongSORT only asks its CMC for a warp when tracks exist; this stub answers by frame index."""

    def __init__(self, warps):
        self.warps, self.f = warps, 0

    def apply(self, img, dets):
        return self.warps[self.f]


def run(tracker, frames, img, embs=None):
    rows, offsets, snaps = [], [0], {}
    for f, dets in enumerate(frames):
        if isinstance(getattr(tracker, "cmc", None), _FrameWarps):
            tracker.cmc.f = f
        e = None if embs is None else embs[f].copy()
        if e is not None and len(dets) == 0:
            e = np.empty((0, e.shape[1] if e.ndim == 2 else 512), np.float32)
        out = np.asarray(tracker.update(dets.copy(), img, e) if e is not None else tracker.update(dets.copy(), img))
        out = out.reshape(-1, 8) if out.size else np.empty((0, 8), np.float32)
        rows.append(out.astype(np.float32))
        offsets.append(offsets[-1] + len(out))
        if (f + 1) in SNAP_FRAMES:
            kind = tracker.__class__.__name__
            ids, m, c = (_snapshot_docs(tracker) if kind == "DeepOcSort" else
                         _snapshot_ss(tracker) if kind == "StrongSort" else _snapshot(tracker))
            snaps[f"snap{f + 1}_ids"] = ids
            snaps[f"snap{f + 1}_mean"] = m
            snaps[f"snap{f + 1}_cov"] = c
    return dict(rows=np.concatenate(rows, 0), offsets=np.asarray(offsets, np.int64), **snaps)


def make_reid_golden():
    """Crops and embeddings from the reference classes (BaseModelBackend.get_crops / get_features over the
    reference OSNet) with seeded weights from oracle.reid.make_osnet_state (no pretrained files exist here)."""
    import hashlib

    import torch
    from boxmot.reid.backbones.osnet import osnet_x0_25
    from boxmot.reid.backends.base_backend import BaseModelBackend
    from boxmot.reid.core.preprocessing import get_preprocess_fn

    from oracle import reid as orid

    class RefBackend(BaseModelBackend):
        def __init__(self, model):
            self.device = torch.device("cpu")
            self.half = False
            self.input_shape = (256, 128)
            self.nhwc = False
            self.preprocess_fn = get_preprocess_fn(None)
            self.mean_array = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
            self.std_array = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
            self.model = model

        def forward(self, x):
            return self.model(x)

        def load_model(self, w):
            pass

    sd = orid.make_osnet_state("osnet_x0_25", seed=7)
    model = osnet_x0_25(num_classes=1041, pretrained=False)
    model.load_state_dict(sd, strict=True)
    model.eval()
    img = np.random.default_rng(123).integers(0, 255, size=(720, 1280, 3), dtype=np.uint8)
    r2 = np.random.default_rng(5)
    boxes = []
    for _ in range(12):
        cx, cy = r2.uniform(0, 1280), r2.uniform(0, 720)
        w, h = r2.uniform(8, 160), r2.uniform(16, 320)
        boxes.append([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2])
    boxes += [[-50, -50, -10, -10], [10.5, 20.5, 11.4, 300.2], [0, 0, 1280, 720], [600, 300, 728, 556]]
    boxes = np.array(boxes, np.float32)
    backend = RefBackend(model)
    crops = backend.get_crops(boxes, img).numpy()
    feats = backend.get_features(boxes, img)
    np.savez_compressed(HERE / "reid_osnet_x0_25.npz", boxes=boxes,
                        crops_sha256=hashlib.sha256(np.ascontiguousarray(crops).tobytes()).hexdigest(),
                        crops_sample=crops[[0, 13]], crops_sample_index=np.array([0, 13]),
                        features=feats.astype(np.float32), weight_seed=7, image_seed=123)


def make_strongsort_goldens():
    """StrongSORT: GITHUB_ACTIONS unset, the ECC estimator replaced by a given warp (identity unless
    the case supplies one) while camera_update itself still runs."""
    from boxmot.trackers.bbox.strongsort.strongsort import StrongSort
    import importlib.util

    spec = importlib.util.spec_from_file_location("b200_tests_common", HERE.parent / "common.py")  # the reference
    common = importlib.util.module_from_spec(spec)                                               # has a `tests` too
    spec.loader.exec_module(common)
    CASES, WARPS = common.CASES, common.WARPS

    img = np.zeros((360, 640, 3), np.uint8)
    for name, (kind, kwargs, make_frames, make_embs) in CASES.items():
        if kind != "strongsort":
            continue
        frames = make_frames()
        embs = make_embs(frames)
        trk = StrongSort(reid_model=None, **kwargs)
        trk.cmc = _FrameWarps(WARPS[name]() if name in WARPS else [np.eye(2, 3)] * len(frames))
        np.savez_compressed(HERE / f"{name}.npz", **run(trk, frames, img, embs))


def main():
    img = np.zeros((360, 640, 3), np.uint8)
    if len(sys.argv) > 1 and sys.argv[1] == "strongsort":
        make_strongsort_goldens()
        return

    bt_base.BaseTrack._count = 0
    _, frames = bench_stream(64, 300)
    np.savez_compressed(HERE / "bytetrack_bench64.npz", **run(ByteTrack(**BYTETRACK_YAML), frames, img))

    bt_base.BaseTrack._count = 0
    frames = stress_stream(96, 300)
    np.savez_compressed(HERE / "bytetrack_stress96.npz", **run(ByteTrack(**BYTETRACK_YAML), frames, img))

    bt_base.BaseTrack._count = 0
    frames = stress_stream(48, 200, seed=19, n_classes=3, empty_every=37)
    np.savez_compressed(HERE / "bytetrack_stress48_gaps.npz", **run(ByteTrack(**BYTETRACK_YAML), frames, img))

    frames = stress_stream(96, 300)
    embs = stress_embeddings(frames, 96)
    np.savez_compressed(HERE / "botsort_stress96.npz",
                        **run(BotSort(reid_model=None, use_cmc=False, **BOTSORT_YAML), frames, img, embs))

    frames = stress_stream(48, 200, seed=19, n_classes=3, empty_every=37)
    embs = stress_embeddings(frames, 48, seed=5)
    np.savez_compressed(HERE / "botsort_stress48_gaps.npz",
                        **run(BotSort(reid_model=None, use_cmc=False, **BOTSORT_YAML), frames, img, embs))

    # constructor defaults instead of YAML defaults, no appearance
    frames = stress_stream(64, 200, seed=23)
    np.savez_compressed(HERE / "botsort_noreid_stress64.npz",
                        **run(BotSort(reid_model=None, use_cmc=False, with_reid=False), frames, img))

    from oracle.streams import warp_sequence
    frames = stress_stream(64, 150, seed=29)
    embs = stress_embeddings(frames, 64, seed=31)
    trk = BotSort(reid_model=None, use_cmc=False, **BOTSORT_YAML)
    trk.cmc = _GivenWarps(warp_sequence(150))
    np.savez_compressed(HERE / "botsort_warp_stress64.npz", **run(trk, frames, img, embs))

    _, frames = bench_stream(256, 40)
    embs = stress_embeddings(frames, 256, seed=3)
    np.savez_compressed(HERE / "botsort_bench256.npz",
                        **run(BotSort(reid_model=None, use_cmc=False, **BOTSORT_YAML), frames, img, embs))
    from boxmot.trackers.bbox.deepocsort.deepocsort import DeepOcSort
    from oracle.streams import unit_embeddings
    for name, frames in (("deepocsort_stress96", stress_stream(96, 300)),
                         ("deepocsort_stress48_gaps", stress_stream(48, 200, seed=19, n_classes=3, empty_every=37)),
                         ("deepocsort_bench128", bench_stream(128, 60, hw=(360, 640))[1])):
        embs = unit_embeddings(frames, 96, seed=5)
        np.savez_compressed(HERE / f"{name}.npz", **run(DeepOcSort(reid_model=None, cmc_off=True), frames, img, embs))
    make_reid_golden()
    make_strongsort_goldens()
    for p in sorted(HERE.glob("*.npz")):
        print(p.name, p.stat().st_size)


if __name__ == "__main__":
    main()
