"""Oracle restatement of ``lap.lapjv`` (lapx 0.9.4 contract) -- TEST INFRASTRUCTURE ONLY.

Reference call sites: boxmot/trackers/association/matching.py:28-43 (``extend_cost=True, cost_limit=thresh``)
and boxmot/trackers/association/association.py:20-24 (``extend_cost=True``).  lapx is third-party and absent
from /root/reference (pinned lapx 0.9.4, uv.lock:2522-2523): the wrapper below restates its documented
contract; the dense solver is ``oracle/lapjv.c``.  PARITY UNPINNED against a real lapx binary; pinned against
scipy's exact solver and brute force in tests/test_oracle_lap.py.
"""
from __future__ import annotations

import ctypes

import numpy as np

from .build import build_oracle

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build_oracle()))
        _LIB.oracle_lapjv_square.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _LIB.oracle_lapjv_square.restype = ctypes.c_int
    return _LIB


def extend_cost_matrix(cost: np.ndarray, cost_limit: float = np.inf) -> np.ndarray:
    """The square matrix lapjv actually solves (lapx `lapjv`, extend_cost / cost_limit branches)."""
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    r, c = cost.shape
    if cost_limit < np.inf:
        n = r + c
        ext = np.empty((n, n), dtype=np.float64)
        ext[:] = cost_limit / 2.0
        ext[r:, c:] = 0.0
        ext[:r, :c] = cost
        return ext
    if r != c:
        n = max(r, c)
        ext = np.zeros((n, n), dtype=np.float64)
        ext[:r, :c] = cost
        return ext
    return cost


def lapjv(cost, extend_cost: bool = False, cost_limit: float = np.inf, return_cost: bool = True):
    cost = np.asarray(cost)
    if cost.ndim != 2:
        raise ValueError("2-dimensional array expected")
    r, c = cost.shape
    if r != c and not extend_cost:
        raise ValueError("Square cost array expected. If cost is intentionally non-square, pass extend_cost=True.")
    costd = np.ascontiguousarray(cost, dtype=np.float64)
    ext = extend_cost_matrix(costd, cost_limit)
    n = ext.shape[0]
    x = np.empty(n, dtype=np.int32)
    y = np.empty(n, dtype=np.int32)
    if n > 0:
        _lib().oracle_lapjv_square(n, ext.ctypes.data, x.ctypes.data, y.ctypes.data)
    x = x.astype(np.int64)
    y = y.astype(np.int64)
    if n != r or n != c:
        x[x >= c] = -1
        y[y >= r] = -1
        x = x[:r]
        y = y[:c]
    if return_cost:
        rows = np.nonzero(x >= 0)[0]
        opt = float(costd[rows, x[rows]].sum()) if rows.size else 0.0
        return opt, x, y
    return x, y


def lapjv_via_scipy(cost, cost_limit: float = np.inf):
    """Independent exact solver on the same extended matrix (cross-check only)."""
    from scipy.optimize import linear_sum_assignment

    cost = np.asarray(cost, dtype=np.float64)
    r, c = cost.shape
    ext = extend_cost_matrix(cost, cost_limit)
    ri, ci = linear_sum_assignment(ext)
    x = np.full(ext.shape[0], -1, dtype=np.int64)
    y = np.full(ext.shape[0], -1, dtype=np.int64)
    x[ri] = ci
    y[ci] = ri
    x[x >= c] = -1
    y[y >= r] = -1
    return x[:r], y[:c]
