"""Control-flow parity of the CUDA tracker source (compiled for the host, tests/hostsim.py) against the
goldens dumped from the unmodified reference.  ids / det_ind / conf / cls must be bit-exact on every frame;
boxes and Kalman state within 1e-4 relative (BASELINE.json north_star); in practice they agree to ~1e-12."""
import numpy as np
import pytest

from oracle.lap import lapjv
from tests.common import CASES, WARPS, assert_rows_match, load_golden
from tests.hostsim import (HostSimDeepOcSort, HostSimStrongSort, HostSimTracker, botsort_cfg, bytetrack_cfg,
                           deepocsort_cfg, strongsort_cfg)


def _make(kind, kwargs, feat_dim=None):
    if kind == "bytetrack":
        return HostSimTracker(bytetrack_cfg(**kwargs))
    if kind == "deepocsort":
        return HostSimDeepOcSort(deepocsort_cfg(**kwargs))
    if kind == "strongsort":
        return HostSimStrongSort(strongsort_cfg(cap_tracks=512, cap_dets=256, feat_dim=feat_dim, **kwargs))
    return HostSimTracker(botsort_cfg(**kwargs))


@pytest.mark.parametrize("name", sorted(CASES))
def test_hostsim_matches_reference_golden(name):
    kind, kwargs, make_frames, make_embs = CASES[name]
    frames = make_frames()
    embs = make_embs(frames) if make_embs else None
    want, snaps = load_golden(name)
    trk = _make(kind, kwargs, None if embs is None else next(e.shape[1] for e in embs if e.ndim == 2 and len(e)))
    warps = WARPS[name]() if name in WARPS else None
    worst = 0.0
    for f, dets in enumerate(frames):
        extra = {} if warps is None else {"warp": warps[f]}
        got = trk.update(dets, None, None if embs is None else embs[f], **extra)
        assert_rows_match(got, want[f], f, box_rtol=1e-4)
        if (f + 1) in snaps:
            ids, mean, cov = snaps[f + 1]
            st = trk.state_snapshot()
            assert sorted(st) == sorted(ids.tolist()), f"live track ids differ at frame {f + 1}"
            for i, m, c in zip(ids, mean, cov):
                k = len(st[int(i)][0])  # 8-state STrack filters or the 7-state XYSR filter
                np.testing.assert_allclose(st[int(i)][0], m[:k], rtol=1e-4, atol=1e-7)
                np.testing.assert_allclose(st[int(i)][1], c[:k, :k], rtol=1e-4, atol=1e-6)
                worst = max(worst, float(np.max(np.abs(st[int(i)][0] - m[:k]) / (np.abs(m[:k]) + 1e-6))))
    assert worst < 1e-6


@pytest.mark.parametrize("seed", range(30))
def test_hostsim_lap_equals_lapjv(seed):
    rng = np.random.default_rng(seed)
    T, D = rng.integers(1, 120, 2)
    cost = rng.random((T, D))
    cost[rng.random((T, D)) < rng.uniform(0.3, 0.95)] = 1.0
    thresh = float(rng.uniform(0.2, 0.95))
    trk = HostSimTracker(bytetrack_cfg(cap_tracks=128, cap_dets=128))
    x, y = trk.lap(cost, thresh)
    _, xo, yo = lapjv(cost, extend_cost=True, cost_limit=thresh)
    assert np.array_equal(x, xo) and np.array_equal(y, yo)


def test_hostsim_empty_and_capacity():
    trk = HostSimTracker(bytetrack_cfg(cap_tracks=8, cap_dets=4))
    assert trk.update(np.zeros((0, 6), np.float32)).shape == (0, 8)
    with pytest.raises(RuntimeError):
        trk.update(np.zeros((5, 6), np.float32))


def _tie_heavy(rng, r, c):
    """DeepOCSORT-shaped costs: mostly exact zeros, a few negative entries, some exact duplicates."""
    cost = np.zeros((r, c))
    k = rng.integers(0, r * c // 2 + 1)
    cost.flat[rng.choice(r * c, size=k, replace=False)] = -np.round(rng.random(k), 2)  # rounded -> value ties
    return cost


@pytest.mark.parametrize("seed", range(40))
def test_hostsim_dense_jv_reproduces_lapjv_ties(seed):
    """The dense JV restatement must pick the SAME optimum as the oracle's lapjv among ties (ids depend on it)."""
    rng = np.random.default_rng(seed)
    r, c = rng.integers(1, 40, 2)
    cost = _tie_heavy(rng, r, c)
    sim = HostSimDeepOcSort(deepocsort_cfg(cap_tracks=64, cap_dets=64, feat_dim=0))
    x, y = sim.jv(cost)
    _, xo, yo = lapjv(cost, extend_cost=True)
    assert np.array_equal(x, xo) and np.array_equal(y, yo)


@pytest.mark.parametrize("mode", [1, 2, 3, 7, 11, 19, 35])
@pytest.mark.parametrize("seed", range(24))
def test_hostsim_dense_jv_wide_augmentation_reproduces_lapjv_ties(seed, mode):
    """The CTA-wide augmentation (jv_augment_wide: relax every open position, then replay the band-minimum hits in
    position order; zero-padding rows never loaded) against the oracle's lapjv, at sizes where it is active (n >= 64)."""
    rng = np.random.default_rng(1000 + seed)
    r, c = rng.integers(1, 180, 2)
    if max(r, c) < 64:
        c = 64 + seed
    sim = HostSimDeepOcSort(deepocsort_cfg(cap_tracks=256, cap_dets=256, feat_dim=0))
    sim.set_jv_wide(mode)
    sparse = np.zeros((r, c))
    k2 = max(1, r * c // 50)
    sparse.flat[rng.choice(r * c, size=k2, replace=False)] = -np.round(rng.random(k2), 1)
    for cost in (_tie_heavy(rng, r, c), -rng.random((r, c)), -np.round(rng.random((r, c)), 1), sparse, np.zeros((r, c))):
        x, y = sim.jv(cost)
        _, xo, yo = lapjv(cost, extend_cost=True)
        assert np.array_equal(x, xo) and np.array_equal(y, yo)


@pytest.mark.parametrize("mode", [1, 2, 3, 7, 11, 19, 35])
@pytest.mark.parametrize("name", sorted(n for n in CASES if CASES[n][0] == "deepocsort"))
def test_hostsim_deepocsort_golden_with_wide_augmentation(name, mode):
    kind, kwargs, make_frames, make_embs = CASES[name]
    frames = make_frames()
    embs = make_embs(frames) if make_embs else None
    want, _ = load_golden(name)
    trk = _make(kind, kwargs)
    trk.set_jv_wide(mode)
    warps = WARPS[name]() if name in WARPS else None
    for f, dets in enumerate(frames):
        extra = {} if warps is None else {"warp": warps[f]}
        assert_rows_match(trk.update(dets, None, None if embs is None else embs[f], **extra), want[f], f, box_rtol=1e-4)


@pytest.mark.parametrize("seed", range(40))
def test_hostsim_lsa_reproduces_scipy_ties(seed):
    """lsa_sap.cuh must return scipy.optimize.linear_sum_assignment's own choice among tied optima: StrongSORT's
    clipped cost matrices tie by construction and the choice orders the births (ids)."""
    from scipy.optimize import linear_sum_assignment

    rng = np.random.default_rng(seed)
    sim = HostSimStrongSort(strongsort_cfg(cap_tracks=64, cap_dets=64, feat_dim=4, nn_budget=2))
    for trial in range(25):
        r, c = rng.integers(1, 48, 2)
        mode = (seed + trial) % 4
        if mode == 0:
            cost = rng.integers(0, 3, (r, c)).astype(float)
        elif mode == 1:
            cost = rng.random((r, c))
            cost[cost > 0.5] = 0.5 + 1e-5
        elif mode == 2:
            cost = np.zeros((r, c))
        else:
            cost = rng.random((r, c))
            cost[cost > 0.3] = 0.7 + 1e-5
            cost[rng.integers(0, r)] = 0.7 + 1e-5
        ri, ci = sim.lsa(cost)
        ro, co = linear_sum_assignment(cost)
        assert np.array_equal(ri, ro) and np.array_equal(ci, co)


@pytest.mark.parametrize("seed", range(10))
def test_hostsim_set_order_equals_cpython(seed):
    """pyset.cuh against the interpreter's own `list(set(a) - set(b))` (linear_assignment.py:108)."""
    import random

    rnd = random.Random(seed)
    sim = HostSimStrongSort(strongsort_cfg(cap_tracks=1024, cap_dets=8, feat_dim=4, nn_budget=2))
    for _ in range(600):
        T = rnd.choice([5, 20, 60, 200, 1000])
        n = rnd.randint(0, T)
        a = sorted(rnd.sample(range(T), n))
        m = rnd.randint(0, n) if rnd.random() < 0.7 else rnd.randint(0, max(0, n // 6))
        b = set(rnd.sample(a, m))
        assert sim.set_difference(a, [k in b for k in a]) == list(set(a) - set(k for k in a if k in b))
    assert sim.set_difference([3, 17], [False, False]) == [17, 3]
