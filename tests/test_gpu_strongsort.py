"""StrongSORT pieces on the GPU beyond the golden table in test_gpu_trackers.py: the scipy-faithful assignment
solver, the tiled gallery-distance kernel against its plain restatement, multi-stream equality, on-device ReID."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.common import CASES, assert_rows_match, load_golden


def _lsa(cost):
    import ctypes

    from boxmot_b200 import _lib

    lib = _lib.require_device()
    cost = np.ascontiguousarray(cost, np.float64)
    r, c = cost.shape
    ri = np.empty(max(min(r, c), 1), np.int32)
    ci = np.empty(max(min(r, c), 1), np.int32)
    n = ctypes.c_int(0)
    assert lib.boxmot_b200_lsa_solve(cost.ctypes.data, r, c, ri.ctypes.data, ci.ctypes.data, ctypes.byref(n)) == 1, \
        _lib.last_error(lib)
    return ri[: n.value], ci[: n.value]


@pytest.mark.parametrize("seed", range(12))
def test_gpu_lsa_reproduces_scipy_ties(seed):
    from scipy.optimize import linear_sum_assignment

    rng = np.random.default_rng(seed)
    for trial in range(6):
        r, c = rng.integers(1, 200, 2)
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
        ri, ci = _lsa(cost)
        ro, co = linear_sum_assignment(cost)
        assert np.array_equal(ri, ro) and np.array_equal(ci, co)


def test_gpu_lsa_large_and_value():
    """600 x 300 generic costs: same optimum value as scipy and, costs being distinct, the same assignment."""
    from scipy.optimize import linear_sum_assignment

    rng = np.random.default_rng(5)
    cost = rng.random((600, 300))
    ri, ci = _lsa(cost)
    ro, co = linear_sum_assignment(cost)
    assert np.array_equal(ri, ro) and np.array_equal(ci, co)


def test_gpu_strongsort_matches_oracle_live_wide_features():
    """1792-d rows (MobileNetV2 width, SURVEY config C4) through the tiled gallery kernel, against the oracle."""
    import boxmot_b200 as bb
    from oracle.streams import stress_embeddings, stress_stream
    from oracle.strongsort import StrongSortOracle

    frames = stress_stream(40, 60, seed=77, n_classes=2)
    embs = stress_embeddings(frames, 40, dim=1792, seed=78)
    kw = dict(min_conf=0.3, max_cos_dist=0.3, n_init=2, nn_budget=7, max_age=12)
    orc = StrongSortOracle(**kw)
    gpu = bb.StrongSort(cap_tracks=256, cap_dets=128, feat_dim=1792, **kw)
    for f, (d, e) in enumerate(zip(frames, embs)):
        assert_rows_match(gpu.update(d, None, e), orc.update(d, None, e.copy()), f)


def test_gpu_strongsort_multistream_equals_independent():
    import boxmot_b200 as bb

    names = ["strongsort_stress96", "strongsort_warp_stress64", "strongsort_bench128"]
    streams = [CASES[n][2]() for n in names]
    embs = [CASES[n][3](fr) for n, fr in zip(names, streams)]
    golds = [load_golden(n)[0] for n in names]
    from tests.common import WARPS

    warps = WARPS["strongsort_warp_stress64"]()
    ms = bb.MultiStreamTracker("strongsort", n_streams=3, cap_tracks=512, cap_dets=256, feat_dim=512)
    for f in range(30):
        ms.set_warp(1, warps[f])
        outs = ms.update([s[f] for s in streams], None, [e[f] for e in embs])
        for k in range(3):
            assert_rows_match(outs[k], golds[k][f], f)


def test_gpu_strongsort_device_reid_matches_oracle(tmp_path):
    """Frames in, ids out: crops -> OSNet -> gallery distances -> assignment, all on the device, against the oracle
    fed with the oracle ReID features."""
    import boxmot_b200 as bb
    from boxmot_b200.reid import B200ReID
    from boxmot_b200.synthetic import bench_stream, make_osnet_state
    from boxmot_b200.weights import export_blob
    from oracle import reid as orid
    from oracle.strongsort import StrongSortOracle

    img, frames = bench_stream(24, 12, hw=(360, 640))
    sd = make_osnet_state("osnet_x0_25", seed=7)
    reid = B200ReID(export_blob(sd, tmp_path / "m.b200reid"))
    kw = dict(min_conf=0.3, max_cos_dist=0.4, n_init=2)
    gpu = bb.StrongSort(reid_model=reid, cap_tracks=128, cap_dets=64, **kw)
    orc = StrongSortOracle(reid_model=orid.OracleReID(sd), **kw)
    for f, d in enumerate(frames):
        assert_rows_match(gpu.update(d, img), orc.update(d, img), f)


def test_unbounded_gallery_nn_budget_none_matches_oracle():
    """nn_budget=None (the reference's unbounded sample list, linear_assignment.py:307-331): the device keeps a ring of
    UNBOUNDED_GALLERY_CAP samples per track, identical to the reference until a track outgrows it."""
    import boxmot_b200 as bb
    from oracle.streams import stress_embeddings, stress_stream
    from oracle.strongsort import StrongSortOracle

    frames = stress_stream(40, 150, seed=61, dropout=0.1)
    embs = stress_embeddings(frames, 40, dim=64, seed=62)
    kw = dict(min_conf=0.3, max_cos_dist=0.4, n_init=2, nn_budget=None)
    with pytest.warns(UserWarning, match="nn_budget=None"):
        gpu = bb.StrongSort(cap_tracks=128, cap_dets=64, feat_dim=64, **kw)
    orc = StrongSortOracle(**kw)
    longest = 0
    for f, (d, e) in enumerate(zip(frames, embs)):
        assert_rows_match(gpu.update(d, None, e), orc.update(d, None, e.copy()), f)
        longest = max([longest] + [len(v) for v in orc.samples.values()])
    assert longest > 100, "the stream must fill galleries beyond the default budget of 100"
