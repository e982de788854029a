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
