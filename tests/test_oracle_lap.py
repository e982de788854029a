"""Pin the lapjv restatement (oracle/lapjv.c + oracle/lap.py): brute force for tiny problems and an
independent exact solver (scipy) on the same extended matrix.  lapx itself is not available -> the oracle
header says 'parity unpinned' at that boundary; these checks are what stands in for it."""
import itertools

import numpy as np
import pytest

from oracle.lap import extend_cost_matrix, lapjv, lapjv_via_scipy


def _brute(cost, limit):
    r, c = cost.shape
    best, best_x = None, None
    # choose for each row a column or -1, columns distinct
    def rec(i, used, x, tot):
        nonlocal best, best_x
        if i == r:
            t = tot + limit / 2.0 * ((r - sum(v >= 0 for v in x)) + (c - len(used)))
            if best is None or t < best - 1e-15:
                best, best_x = t, list(x)
            return
        rec(i + 1, used, x + [-1], tot)
        for j in range(c):
            if j not in used:
                rec(i + 1, used | {j}, x + [j], tot + cost[i, j])
    rec(0, frozenset(), [], 0.0)
    return best_x


@pytest.mark.parametrize("seed", range(40))
def test_bruteforce_small(seed):
    rng = np.random.default_rng(seed)
    r, c = rng.integers(1, 6, 2)
    cost = rng.random((r, c))
    limit = float(rng.uniform(0.2, 1.0))
    _, x, y = lapjv(cost, extend_cost=True, cost_limit=limit)
    assert list(x) == _brute(cost, limit)
    for i, j in enumerate(x):
        if j >= 0:
            assert y[j] == i
    assert sorted(np.where(y < 0)[0]) == sorted(set(range(c)) - set(int(v) for v in x if v >= 0))


@pytest.mark.parametrize("seed", range(10))
def test_against_scipy_gated(seed):
    rng = np.random.default_rng(100 + seed)
    r, c = rng.integers(50, 300, 2)
    cost = rng.random((r, c))
    cost[rng.random((r, c)) < 0.85] = 1.0  # gated entries tie exactly, as in the trackers
    _, x, y = lapjv(cost, extend_cost=True, cost_limit=0.8)
    xs, ys = lapjv_via_scipy(cost, 0.8)
    assert np.array_equal(x, xs) and np.array_equal(y, ys)


def test_extension_layout():
    ext = extend_cost_matrix(np.arange(6, dtype=float).reshape(2, 3), 0.5)
    assert ext.shape == (5, 5)
    assert np.all(ext[:2, 3:] == 0.25) and np.all(ext[2:, :3] == 0.25) and np.all(ext[2:, 3:] == 0.0)
    assert np.array_equal(ext[:2, :3], np.arange(6, dtype=float).reshape(2, 3))


def test_rectangular_without_limit_is_optimal():
    rng = np.random.default_rng(5)
    for _ in range(50):
        r, c = rng.integers(1, 25, 2)
        cost = rng.random((r, c))
        _, x, _ = lapjv(cost, extend_cost=True)
        xs, _ = lapjv_via_scipy(cost)
        a = cost[np.nonzero(x >= 0)[0], x[x >= 0]].sum()
        b = cost[np.nonzero(xs >= 0)[0], xs[xs >= 0]].sum()
        assert abs(a - b) < 1e-12
        assert (x >= 0).sum() == min(r, c)


def test_square_requires_flag():
    with pytest.raises(ValueError):
        lapjv(np.zeros((2, 3)))
