"""Standalone hot-path kernels through the C ABI vs the oracle: linear assignment (bit-exact indices), batched
Kalman predict/update (float64, 1e-9), IoU and cosine cost matrices; at config sizes and with the edge cases
the domain has (empty, ragged, fully gated, duplicate boxes)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import kalman as okf
from oracle.association import embedding_cost, iou_batch
from oracle.lap import lapjv


def _lib():
    from boxmot_b200 import _lib as L

    return L.require_device()


def gpu_lap(cost, thresh):
    lib = _lib()
    cost = np.ascontiguousarray(cost, np.float64)
    T, D = cost.shape
    x = np.empty(max(T, 1), np.int32)
    y = np.empty(max(D, 1), np.int32)
    assert lib.boxmot_b200_lap_solve(cost.ctypes.data, T, D, thresh, x.ctypes.data, y.ctypes.data) == 1
    return x[:T], y[:D]


@pytest.mark.parametrize("T,D,density", [(1, 1, 1.0), (64, 64, 0.1), (300, 256, 0.03), (256, 300, 0.05),
                                         (17, 240, 0.5), (2000, 512, 0.01), (512, 2000, 0.01), (100, 100, 1.0)])
def test_lap_matches_lapjv(T, D, density):
    rng = np.random.default_rng(T * 7919 + D)
    cost = rng.random((T, D))
    cost[rng.random((T, D)) > density] = 1.0
    for thresh in (0.3, 0.8, 0.95):
        x, y = gpu_lap(cost, thresh)
        _, xo, yo = lapjv(cost, extend_cost=True, cost_limit=thresh)
        assert np.array_equal(x, xo) and np.array_equal(y, yo)


def test_lap_edge_cases():
    x, y = gpu_lap(np.ones((5, 7)), 0.8)  # everything gated
    assert (x == -1).all() and (y == -1).all()
    x, y = gpu_lap(np.zeros((0, 4)), 0.8)
    assert x.size == 0 and (y == -1).all()
    c = np.full((3, 3), np.nan)
    c[0, 1] = 0.1
    x, y = gpu_lap(c, 0.8)  # NaN entries are never candidates
    assert list(x) == [1, -1, -1]


def test_lap_tracking_like_structure():
    """IoU-shaped costs from a real stream: near-diagonal with crowding."""
    from oracle.streams import stress_stream

    frames = stress_stream(200, 12, seed=3, dropout=0.1)
    for a, b in zip(frames[:-1], frames[1:]):
        cost = 1 - iou_batch(a[:, :4].astype(np.float64), b[:, :4])
        x, y = gpu_lap(cost, 0.9)
        _, xo, yo = lapjv(cost, extend_cost=True, cost_limit=0.9)
        assert np.array_equal(x, xo) and np.array_equal(y, yo)


@pytest.mark.parametrize("kind,name", [(0, "xyah"), (1, "xywh")])
def test_kalman_batch(kind, name):
    lib = _lib()
    rng = np.random.default_rng(kind)
    n = 2000
    box = np.stack([rng.uniform(0, 1900, n), rng.uniform(0, 1000, n), rng.uniform(20, 120, n),
                    rng.uniform(40, 250, n)], 1).astype(np.float32)
    meas = box.copy()
    if name == "xyah":
        meas[:, 2] = box[:, 2] / box[:, 3]
    mean = np.zeros((n, 8))
    cov = np.zeros((n, 8, 8))
    assert lib.boxmot_b200_kalman_initiate(kind, meas.ctypes.data, mean.ctypes.data, cov.ctypes.data, n) == 1
    om, oc = zip(*[okf.initiate(name, m) for m in meas])
    om, oc = np.array(om), np.array(oc)
    assert np.array_equal(mean, om) and np.array_equal(cov, oc)  # initiate is exactly reproducible
    tracked = (rng.random(n) < 0.7).astype(np.int32)
    for step in range(5):
        mean[:, 4:] += rng.normal(0, 0.5, (n, 4)) * (step == 0)
        om = mean.copy()
        oc = cov.copy()
        assert lib.boxmot_b200_kalman_predict(kind, mean.ctypes.data, cov.ctypes.data, tracked.ctypes.data, n) == 1
        for i in range(n):
            if not tracked[i]:
                om[i, 7 if name == "xyah" else slice(6, 8)] = 0
        om, oc = okf.multi_predict(name, om, oc)
        assert np.array_equal(mean, om), "predict mean must be bit-exact"
        assert np.array_equal(cov, oc), "predict covariance must be bit-exact (no FMA contraction)"
        z = (meas + rng.normal(0, 1.0, meas.shape) * np.array([1, 1, 0.01 if name == "xyah" else 1, 1])).astype(np.float32)
        assert lib.boxmot_b200_kalman_update(kind, mean.ctypes.data, cov.ctypes.data, z.ctypes.data, n) == 1
        for i in range(n):
            om[i], oc[i] = okf.update(name, om[i], oc[i], z[i])
        np.testing.assert_allclose(mean, om, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(cov, oc, rtol=1e-9, atol=1e-10)
        mean, cov = om.copy(), oc.copy()


def test_iou_and_cosine_cost():
    lib = _lib()
    rng = np.random.default_rng(0)
    T, D, F = 300, 256, 512
    t = np.sort(rng.uniform(0, 500, (T, 2, 2)), axis=1).transpose(0, 1, 2).reshape(T, 4)[:, [0, 1, 2, 3]]
    t = np.stack([t[:, 0], t[:, 1], t[:, 2], t[:, 3]], 1)
    t = np.stack([np.minimum(t[:, 0], t[:, 2]), np.minimum(t[:, 1], t[:, 3]), np.maximum(t[:, 0], t[:, 2]) + 1,
                  np.maximum(t[:, 1], t[:, 3]) + 1], 1)
    d = (t[rng.integers(0, T, D)] + rng.normal(0, 5, (D, 4))).astype(np.float32)
    out = np.empty((T, D))
    assert lib.boxmot_b200_iou_cost(t.ctypes.data, T, d.ctypes.data, D, out.ctypes.data) == 1
    want = 1 - iou_batch(t, d)
    assert np.array_equal(out, want), "IoU cost is elementwise IEEE arithmetic: must be bit-exact"
    a = np.abs(rng.normal(size=(T, F))).astype(np.float32)
    b = np.abs(rng.normal(size=(D, F))).astype(np.float32)
    assert lib.boxmot_b200_cosine_cost(a.ctypes.data, T, b.ctypes.data, D, F, out.ctypes.data) == 1
    np.testing.assert_allclose(out, embedding_cost(a, b), rtol=0, atol=1e-13)


@pytest.mark.parametrize("wide", [3, 7, 11, 19, 35, 2, 1, 0])
@pytest.mark.parametrize("seed,r,c", [(0, 1, 1), (1, 7, 30), (2, 33, 32), (3, 64, 64), (4, 65, 200), (5, 200, 65),
                                      (6, 130, 131), (7, 300, 300), (8, 97, 512), (9, 512, 97), (10, 40, 700)])
def test_dense_jv_reproduces_lapjv_ties(seed, r, c, wide):
    """GPU dense Jonker-Volgenant vs the oracle's lapjv on tie-heavy matrices: identical x / y, i.e. identical
    tie-breaking, which DeepOCSORT's birth order (ids) depends on.  Every augmentation variant: 3 = column-owned with
    every exact shortcut (the default), 7 / 11 / 19 / 35 = mode 3 with ONE shortcut (no-op band columns / parallel
    _find_dense tail / hit list / CTA-wide row reduction: `3 | bit << 2`), 2 = CTA-wide with owned columns in registers,
    1 = CTA-wide over list positions, 0 = one warp."""
    lib = _lib()
    assert lib.boxmot_b200_jv_dense_mode(wide) == 1
    try:
        rng = np.random.default_rng(seed)
        x = np.empty(r, np.int32)
        y = np.empty(c, np.int32)
        k = int(rng.integers(0, r * c // 2 + 1))
        tie = np.zeros((r, c))
        tie.flat[rng.choice(r * c, size=k, replace=False)] = -np.round(rng.random(k), 2)
        sparse = np.zeros((r, c))   # mostly zeros: long runs of equal distances (the config-3 regime)
        k2 = max(1, r * c // 50)
        sparse.flat[rng.choice(r * c, size=k2, replace=False)] = -np.round(rng.random(k2), 1)
        # continuous costs as well (unique optimum), and all-zero
        for cost in (tie, sparse, -rng.random((r, c)), np.zeros((r, c))):
            assert lib.boxmot_b200_jv_dense(cost.ctypes.data, r, c, x.ctypes.data, y.ctypes.data) == 1
            _, xo, yo = lapjv(cost, extend_cost=True)
            assert np.array_equal(x, xo) and np.array_equal(y, yo)
    finally:
        lib.boxmot_b200_jv_dense_mode(3)
