"""The numpy oracle (oracle/trackers.py) against golden vectors dumped from the unmodified reference
(tests/golden/make_golden.py).  Same machine-independent integer outputs (ids, det_ind) must be bit-exact;
on the machine that generated the goldens the float columns are bit-exact too, elsewhere BLAS may differ in
the last ulp, so boxes are compared at 1e-6 relative."""
import numpy as np
import pytest

from oracle.deepocsort import DeepOcSortOracle
from oracle.strongsort import StrongSortOracle
from oracle.trackers import BotSortOracle, ByteTrackOracle
from tests.common import CASES, WARPS, assert_rows_match, load_golden


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
    kind, kwargs, make_frames, make_embs = CASES[name]
    frames = make_frames()
    embs = make_embs(frames) if make_embs else None
    want, snaps = load_golden(name)
    trk = {"bytetrack": ByteTrackOracle, "botsort": BotSortOracle, "deepocsort": DeepOcSortOracle,
           "strongsort": StrongSortOracle}[kind](**kwargs)
    warps = WARPS[name]() if name in WARPS else None
    img = np.zeros((360, 640, 3), np.uint8)
    for f, dets in enumerate(frames):
        extra = {} if warps is None else {"warp": warps[f]}
        got = trk.update(dets.copy(), img, None if embs is None else embs[f].copy(), **extra)
        assert_rows_match(got, want[f], f, box_rtol=1e-6)
        if (f + 1) in snaps:
            ids, mean, cov = snaps[f + 1]
            st = trk.state_snapshot()
            if kind == "strongsort":
                st = {int(i): (m, c) for i, m, c in zip(*st)}
            assert sorted(st) == sorted(ids.tolist())
            for i, m, c in zip(ids, mean, cov):
                k = len(st[int(i)][0])  # 8 (STrack filters) or 7 (XYSR)
                np.testing.assert_allclose(st[int(i)][0], m[:k], rtol=1e-9, atol=1e-12)
                np.testing.assert_allclose(st[int(i)][1], c[:k, :k], rtol=1e-9, atol=1e-12)
