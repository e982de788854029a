"""Randomised soak of the device tracker source (host simulation) against the oracle: many seeded stress streams with
varied crowding, drop-out, class counts, empty frames and thresholds, all four trackers.  A disagreement is a parity bug
in the device core or the oracle.   python tests/tools/soak_hostsim.py [n_cases] [first_seed]"""
from __future__ import annotations

import os
import sys
import time

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):   # tiny matrices: BLAS threads only contend
    os.environ.setdefault(_v, "1")
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.deepocsort import DeepOcSortOracle  # noqa: E402
from oracle.streams import stress_embeddings, stress_stream, unit_embeddings, warp_sequence  # noqa: E402
from oracle.strongsort import StrongSortOracle  # noqa: E402
from oracle.trackers import BotSortOracle, ByteTrackOracle  # noqa: E402
from tests.common import BOTSORT_YAML, BYTETRACK_YAML, DEEPOCSORT_YAML, STRONGSORT_YAML, assert_rows_match  # noqa: E402
from tests.hostsim import (HostSimDeepOcSort, HostSimStrongSort, HostSimTracker, botsort_cfg, bytetrack_cfg,  # noqa: E402
                           deepocsort_cfg, strongsort_cfg)


def case_with_warps(seed, build=True):
    """case(seed) plus, for every second seed of the trackers that take one, a supplied camera-motion warp per frame.
    build=False: inputs only (no host simulation / oracle objects), for the GPU soak."""
    kind, kw, frames, embs, sim, orc = case(seed, build)
    warps = None
    if kind != "bytetrack" and (seed // 4) % 2 == 1:
        warps = warp_sequence(len(frames), seed=seed + 77)
    return kind, kw, frames, embs, sim, orc, warps


def case(seed, build=True):
    rng = np.random.default_rng(seed)
    kind = ["bytetrack", "botsort", "deepocsort", "strongsort"][seed % 4]
    heavy = os.environ.get("SOAK_HEAVY") == "1"   # crowded, long streams (slow): rare lifecycle paths
    n_obj = int(rng.integers(100, 200)) if heavy else int(rng.integers(8, 90))
    n_frames = int(rng.integers(150, 260)) if heavy else int(rng.integers(40, 140))
    kw_stream = dict(seed=int(rng.integers(1, 10**6)), n_classes=int(rng.integers(1, 4)),
                     empty_every=int(rng.choice([0, 0, 17, 29])), dropout=float(rng.choice([0.05, 0.2, 0.4])))
    frames = stress_stream(n_obj, n_frames, **kw_stream)
    dim = 64
    real = (seed // 8) % 3 == 0   # every third block of seeds: a window of the MOT17-mini public detections instead
    if real:
        from tests.common import mot17_embeddings, mot17_stream

        seq = "04" if seed % 2 else "02"
        full = mot17_stream(seq)
        n_frames = min(n_frames, len(full) - 1)
        start = int(rng.integers(0, len(full) - n_frames))
        all_embs = mot17_embeddings(seq, full, dim=dim, seed=seed + 9, unit=(kind == "deepocsort"))
        frames, real_embs = full[start:start + n_frames], all_embs[start:start + n_frames]
    if kind == "bytetrack":
        kw = dict(BYTETRACK_YAML, track_thresh=float(rng.uniform(0.3, 0.7)), match_thresh=float(rng.uniform(0.6, 0.95)),
                  track_buffer=int(rng.integers(5, 40)), frame_rate=int(rng.choice([25, 30])),
                  min_conf=float(rng.uniform(0.05, 0.3)))
        if not build:
            return kind, kw, frames, None, None, None
        return kind, kw, frames, None, HostSimTracker(bytetrack_cfg(**kw)), ByteTrackOracle(**kw)
    if kind == "botsort":
        kw = dict(BOTSORT_YAML, track_high_thresh=float(rng.uniform(0.4, 0.7)), new_track_thresh=float(rng.uniform(0.4, 0.75)),
                  match_thresh=float(rng.uniform(0.6, 0.9)), appearance_thresh=float(rng.uniform(0.2, 0.7)),
                  proximity_thresh=float(rng.uniform(0.4, 0.7)), fuse_first_associate=bool(rng.integers(0, 2)),
                  track_buffer=int(rng.integers(5, 50)), removed_stracks_buffer=int(rng.choice([3, 20, 329])),
                  track_low_thresh=float(rng.uniform(0.05, 0.3)), second_match_thresh=float(rng.uniform(0.2, 0.6)),
                  unconfirmed_match_thresh=float(rng.uniform(0.3, 0.8)), unconfirmed_emb_scale=float(rng.uniform(1.0, 3.0)),
                  with_reid=bool(rng.integers(0, 4) > 0), frame_rate=int(rng.choice([25, 30])))
        embs = real_embs if real else stress_embeddings(frames, n_obj, dim=dim, seed=seed + 5)
        if not build:
            return kind, kw, frames, embs, None, None
        return kind, kw, frames, embs, HostSimTracker(botsort_cfg(feat_dim=dim, **kw)), BotSortOracle(**kw)
    if kind == "deepocsort":
        kw = dict(DEEPOCSORT_YAML, det_thresh=float(rng.uniform(0.2, 0.6)), w_association_emb=float(rng.uniform(0.2, 0.9)),
                  inertia=float(rng.uniform(0.05, 0.4)), delta_t=int(rng.integers(1, 5)), max_age=int(rng.integers(5, 35)),
                  min_hits=int(rng.integers(1, 4)), aw_off=bool(rng.integers(0, 2)), iou_threshold=float(rng.uniform(0.2, 0.4)),
                  embedding_off=bool(rng.integers(0, 5) == 0), alpha_fixed_emb=float(rng.uniform(0.8, 0.98)),
                  aw_param=float(rng.uniform(0.3, 0.7)), Q_xy_scaling=float(rng.choice([0.01, 0.05])),
                  Q_s_scaling=float(rng.choice([0.0001, 0.001])))
        embs = real_embs if real else unit_embeddings(frames, n_obj, dim=dim, seed=seed + 5)
        if not build:
            int(rng.integers(0, 3))
            return kind, kw, frames, embs, None, None
        sim = HostSimDeepOcSort(deepocsort_cfg(feat_dim=dim, **kw))
        mode = int(rng.integers(0, 3))
        sim.set_jv_wide(int(os.environ.get("SOAK_JV_MODE", mode)))   # SOAK_JV_MODE=3: the default variant with every shortcut
        return kind, kw, frames, embs, sim, DeepOcSortOracle(**kw)
    kw = dict(STRONGSORT_YAML, min_conf=float(rng.uniform(0.2, 0.6)), max_cos_dist=float(rng.uniform(0.2, 0.5)),
              n_init=int(rng.integers(1, 4)), nn_budget=int(rng.choice([5, 30, 100])), max_age=int(rng.integers(5, 35)),
              ema_alpha=float(rng.choice([0.8, 0.9])), mc_lambda=float(rng.choice([0.9, 0.98])),
              max_iou_dist=float(rng.uniform(0.5, 0.9)))
    embs = real_embs if real else stress_embeddings(frames, n_obj, dim=dim, seed=seed + 5)
    if not build:
        return kind, kw, frames, embs, None, None
    return kind, kw, frames, embs, HostSimStrongSort(strongsort_cfg(cap_tracks=512, cap_dets=256, feat_dim=dim, **kw)), StrongSortOracle(**kw)


def same_up_to_an_id_permutation(seed, make_device=None) -> bool:
    """True when the two outputs of a diverged case agree in every row except for a consistent renaming of track ids
    (the StrongSORT birth-order limit of DESIGN 3.1c), False for any other difference.  `make_device(kind, kw)` swaps
    the host simulation for another implementation of the same interface (the GPU soak passes the device tracker)."""
    kind, kw, frames, embs, sim, orc, warps = case_with_warps(seed)
    if make_device is not None:
        sim = make_device(kind, kw)
    fwd, bwd = {}, {}
    for f, d in enumerate(frames):
        e = None if embs is None else embs[f]
        x = {} if warps is None else {"warp": warps[f]}
        got = np.asarray(sim.update(d, None, e, **x), np.float32).reshape(-1, 8)
        want = orc.update(d, None) if embs is None else orc.update(d, None, e.copy(), **x)
        want = np.asarray(want, np.float32).reshape(-1, 8)
        if got.shape != want.shape:
            return False
        a = got[np.argsort(got[:, 7], kind="stable")]
        b = want[np.argsort(want[:, 7], kind="stable")]
        if not np.array_equal(a[:, 5:], b[:, 5:]) or not np.allclose(a[:, :4], b[:, :4], rtol=1e-4, atol=1e-3):
            return False
        for i, j in zip(a[:, 4].tolist(), b[:, 4].tolist()):
            if fwd.setdefault(i, j) != j or bwd.setdefault(j, i) != i:
                return False
    return True


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    t0 = time.time()
    for seed in range(first, first + n):
        kind, kw, frames, embs, sim, orc, warps = case_with_warps(seed)
        try:
            for f, d in enumerate(frames):
                e = None if embs is None else embs[f]
                x = {} if warps is None else {"warp": warps[f]}
                got = sim.update(d, None, e, **x)
                want = orc.update(d, None) if embs is None else orc.update(d, None, e.copy(), **x)
                assert_rows_match(got, want, f, box_rtol=1e-4)
        except AssertionError as ex:
            bad += 1
            tag = "ids permuted, everything else equal" if same_up_to_an_id_permutation(seed) else "REAL DIFFERENCE"
            print(f"seed {seed} {kind} warps={warps is not None} DIVERGED ({tag}): {str(ex).splitlines()[0]}  kw={kw}")
        if (seed - first + 1) % 50 == 0:
            print(f"  ... {seed - first + 1} cases, {bad} diverged, {time.time() - t0:.0f} s", flush=True)
    print(f"{n} cases, {bad} diverged, {time.time() - t0:.0f} s")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
