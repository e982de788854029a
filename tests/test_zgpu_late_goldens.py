"""GPU golden parity for the cases added after the round-1 GPU budget was spent (tests/common.py::LATE_CASES): the
MOT17-mini streams through all four trackers and DeepOCSORT with a supplied camera-motion warp.  Same check as
tests/test_gpu_trackers.py::test_gpu_tracker_matches_reference_golden; this file sorts last on purpose."""
import pytest

pytestmark = pytest.mark.gpu

from tests.common import LATE_CASES
from tests.test_gpu_trackers import run_golden_case


@pytest.mark.parametrize("name", sorted(LATE_CASES))
def test_gpu_tracker_matches_reference_golden_late_cases(name):
    run_golden_case(name)


@pytest.mark.parametrize("name", sorted(__import__("tests.common", fromlist=["OCSORT_CASES"]).OCSORT_CASES))
def test_ocsort_matches_reference_golden_on_the_device(name):
    """OC-SORT on the DeepOCSORT device core (appearance term off) against goldens from the unmodified OcSort class."""
    import numpy as np

    import boxmot_b200 as bb
    from tests.common import OCSORT_CASES, assert_rows_match, load_golden

    _, kwargs, make_frames, _ = OCSORT_CASES[name]
    want, snaps = load_golden(name)
    trk = bb.OcSort(cap_tracks=512, cap_dets=256, **kwargs)
    img = np.zeros((64, 64, 3), np.uint8)
    for f, dets in enumerate(make_frames()):
        assert_rows_match(trk.update(dets, img), want[f], f, box_rtol=1e-4)
        if (f + 1) in snaps:
            ids, mean, cov = snaps[f + 1]
            st = trk.snapshot()
            assert sorted(st) == sorted(ids.tolist())
            for i, m, c in zip(ids, mean, cov):
                np.testing.assert_allclose(st[int(i)][0], m, rtol=1e-4, atol=1e-7)
                np.testing.assert_allclose(st[int(i)][1], c, rtol=1e-4, atol=1e-6)
    with pytest.raises(NotImplementedError):
        bb.OcSort(use_byte=True)
