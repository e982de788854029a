"""Seeded synthetic detection streams -- TEST / BENCH INPUT GENERATORS (no tracker arithmetic here).

`bench_stream` restates the reference's own throughput-sweep generator
(/root/reference/tests/performance/benchmark_fps.py:60-94: `_make_random_dets` + `_jitter_dets`, seed 42+n,
fixed random uint8 image) -- SURVEY section 8(d).  `stress_stream` is the harder parity stream the survey
asks for: constant-velocity motion with wall bounce, per-frame dropout, per-frame confidences spanning every
threshold, crowding, births and deaths; it exercises lost / re-activated / unconfirmed / removed tracks.
"""
from __future__ import annotations

import numpy as np


from boxmot_b200.synthetic import bench_image, bench_stream  # noqa: E402,F401


def stress_stream(n_objects: int, n_frames: int, hw=(360, 640), seed: int = 7, dropout: float = 0.2,
                  n_classes: int = 1, empty_every: int = 0):
    """Moving, crowded, flickering objects.  Returns [dets_f32 (n_t,6)] * n_frames (n_t varies, may be 0)."""
    h, w = hw
    rng = np.random.default_rng(seed)
    cx = rng.uniform(30, w - 30, size=n_objects)
    cy = rng.uniform(30, h - 30, size=n_objects)
    bw = rng.uniform(20, 60, size=n_objects)
    bh = rng.uniform(40, 110, size=n_objects)
    vx = rng.normal(0.0, 3.0, size=n_objects)
    vy = rng.normal(0.0, 2.0, size=n_objects)
    cls = rng.integers(0, n_classes, size=n_objects).astype(np.float64)
    # cohorts: object k lives during [birth, death)
    birth = rng.integers(0, max(1, n_frames // 2), size=n_objects)
    birth[: n_objects // 2] = 0
    death = birth + rng.integers(n_frames // 4, n_frames, size=n_objects)
    frames = []
    for f in range(n_frames):
        cx += vx
        cy += vy
        hit = (cx < 20) | (cx > w - 20)
        vx[hit] *= -1
        hit = (cy < 20) | (cy > h - 20)
        vy[hit] *= -1
        alive = (birth <= f) & (f < death) & (rng.random(n_objects) >= dropout)
        if empty_every and f > 0 and f % empty_every == 0:
            alive[:] = False
        idx = np.nonzero(alive)[0]
        jx = rng.normal(0.0, 1.5, size=idx.size)
        jy = rng.normal(0.0, 1.5, size=idx.size)
        x1 = np.clip(cx[idx] + jx - bw[idx] / 2, 0, w - 1)
        y1 = np.clip(cy[idx] + jy - bh[idx] / 2, 0, h - 1)
        x2 = np.clip(cx[idx] + jx + bw[idx] / 2, 1, w)
        y2 = np.clip(cy[idx] + jy + bh[idx] / 2, 1, h)
        conf = rng.uniform(0.05, 0.95, size=idx.size)
        dets = np.stack([x1, y1, x2, y2, conf, cls[idx]], axis=1).astype(np.float32) if idx.size else \
            np.zeros((0, 6), dtype=np.float32)
        frames.append(dets)
    return frames


def warp_sequence(n_frames: int, seed: int = 17):
    """Small seeded affine camera motions (2x3, float64), one per frame."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_frames):
        a = rng.normal(0.0, 0.003, size=(2, 2))
        t = rng.normal(0.0, 2.0, size=2)
        out.append(np.hstack([np.eye(2) + a, t[:, None]]))
    return out


def unit_embeddings(frames, n_objects_hint: int, dim: int = 512, seed: int = 11):
    """`stress_embeddings` rows L2-normalised in float32, as a ReID backend's get_features returns them."""
    out = []
    for e in stress_embeddings(frames, n_objects_hint, dim=dim, seed=seed):
        e = np.asarray(e, np.float32)
        out.append(e / np.linalg.norm(e, axis=1, keepdims=True) if len(e) else e)
    return out


def stress_embeddings(frames, n_objects_hint: int, dim: int = 512, seed: int = 11, noise: float = 0.35):
    """Per-detection appearance vectors for `stress_stream` frames (same seed => same vectors).

    Identity is recovered from box size (w,h are constant per object up to clipping), hashed to a prototype;
    the prototype plus per-frame noise gives embeddings whose cosine distances straddle the BoT-SORT gates.
    Rows are NOT normalised (the tracker normalises them, as the reference does).
    """
    rng = np.random.default_rng(seed)
    protos = {}
    out = []
    for dets in frames:
        e = np.zeros((len(dets), dim), dtype=np.float32)
        for i, d in enumerate(dets):
            key = (int(round(float(d[2] - d[0]) * 8)), int(round(float(d[3] - d[1]) * 8)))
            if key not in protos:
                protos[key] = np.abs(rng.normal(size=dim)).astype(np.float32)
            e[i] = np.maximum(protos[key] + noise * rng.normal(size=dim).astype(np.float32), 0.0)
        out.append(e)
    return out
