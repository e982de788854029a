"""Row a15 of SURVEY section 8: applying a SUPPLIED camera-motion warp to the Kalman state (STrack.multi_gmc).
Golden = the unmodified reference BoT-SORT with its CMC estimator replaced by the supplied warps."""
import numpy as np
import pytest

from oracle.streams import stress_embeddings, stress_stream, warp_sequence
from oracle.trackers import BotSortOracle
from tests.common import BOTSORT_YAML, assert_rows_match, load_golden
from tests.hostsim import HostSimTracker, botsort_cfg


def _inputs():
    frames = stress_stream(64, 150, seed=29)
    return frames, stress_embeddings(frames, 64, seed=31), warp_sequence(150)


def test_oracle_warp_matches_reference_golden():
    frames, embs, warps = _inputs()
    want, snaps = load_golden("botsort_warp_stress64")
    trk = BotSortOracle(**BOTSORT_YAML)
    for f, d in enumerate(frames):
        assert_rows_match(trk.update(d, None, embs[f].copy(), warp=warps[f]), want[f], f, box_rtol=1e-6)


def test_hostsim_warp_matches_reference_golden():
    frames, embs, warps = _inputs()
    want, snaps = load_golden("botsort_warp_stress64")
    trk = HostSimTracker(botsort_cfg(**BOTSORT_YAML))
    for f, d in enumerate(frames):
        assert_rows_match(trk.update(d, None, embs[f], warp=warps[f]), want[f], f)
        if (f + 1) in snaps:
            ids, mean, cov = snaps[f + 1]
            st = trk.state_snapshot()
            assert sorted(st) == sorted(ids.tolist())
            for i, m, c in zip(ids, mean, cov):
                np.testing.assert_allclose(st[int(i)][0], m, rtol=1e-4, atol=1e-7)
                np.testing.assert_allclose(st[int(i)][1], c, rtol=1e-4, atol=1e-7)


@pytest.mark.gpu
def test_gpu_warp_matches_reference_golden():
    import boxmot_b200 as bb

    frames, embs, warps = _inputs()
    want, snaps = load_golden("botsort_warp_stress64")
    trk = bb.BotSort(cap_tracks=512, cap_dets=256, **BOTSORT_YAML)
    for f, d in enumerate(frames):
        assert_rows_match(trk.update(d, None, embs[f], warp=warps[f]), want[f], f)
        if (f + 1) in snaps:
            ids, mean, cov = snaps[f + 1]
            st = trk.snapshot()
            assert sorted(st) == sorted(ids.tolist())
            for i, m, c in zip(ids, mean, cov):
                np.testing.assert_allclose(st[int(i)][0], m, rtol=1e-4, atol=1e-7)
    with pytest.raises(bb.B200Error):
        bb.ByteTrack(cap_tracks=64, cap_dets=16)._engine.set_warp(0, np.eye(2, 3))
