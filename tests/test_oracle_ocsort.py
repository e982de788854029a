"""OC-SORT: the oracle restatement (oracle/ocsort.py) and the host simulation of the device core against
goldens dumped from the unmodified reference OcSort class (tests/golden/make_ocsort_golden.py)."""
import numpy as np
import pytest

from oracle.ocsort import OcSortOracle
from tests.common import OCSORT_CASES, assert_rows_match, load_golden


@pytest.mark.parametrize("name", sorted(OCSORT_CASES))
def test_ocsort_oracle_matches_reference_golden(name):
    _, kwargs, make_frames, _ = OCSORT_CASES[name]
    frames = make_frames()
    want, snaps = load_golden(name)
    trk = OcSortOracle(**kwargs)
    img = np.zeros((360, 640, 3), np.uint8)
    for f, dets in enumerate(frames):
        assert_rows_match(trk.update(dets.copy(), img), want[f], f, box_rtol=1e-6)
        if (f + 1) in snaps:
            ids, mean, cov = snaps[f + 1]
            st = trk.state_snapshot()
            assert sorted(st) == sorted(ids.tolist())
            for i, m, c in zip(ids, mean, cov):
                np.testing.assert_allclose(st[int(i)][0], m[:7], rtol=1e-9, atol=1e-12)
                np.testing.assert_allclose(st[int(i)][1], c[:7, :7], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", sorted(OCSORT_CASES))
def test_ocsort_host_simulation_of_the_device_core_matches_golden(name):
    from tests.hostsim import HostSimDeepOcSort, deepocsort_cfg

    _, kwargs, make_frames, _ = OCSORT_CASES[name]
    kw = {k: v for k, v in kwargs.items() if k not in ("min_conf", "use_byte")}
    sim = HostSimDeepOcSort(deepocsort_cfg(embedding_off=True, aw_off=True, **kw))
    want, _ = load_golden(name)
    for f, dets in enumerate(make_frames()):
        assert_rows_match(sim.update(dets, None, None), want[f], f, box_rtol=1e-4)


def test_ocsort_contract():
    from boxmot_b200.trackers import TRACKER_DEFAULTS, resolve_tracker_args

    kind, cls, args = resolve_tracker_args("ocsort")
    assert cls.__name__ == "OcSort" and args["det_thresh"] == 0.6 and args["inertia"] == 0.1 and args["use_byte"] is False
    assert TRACKER_DEFAULTS["ocsort"]["min_conf"] == 0.1
    with pytest.raises(NotImplementedError):
        OcSortOracle(use_byte=True)
