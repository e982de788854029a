"""BASELINE config 3 shape (DeepOCSORT, 512 dets/frame drawn from 2048 objects in 4 cohorts -> ~2000 live
tracks): ids against the oracle on the first frames, and DeepOCSORT with on-device OSNet_x1_0 ReID."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


from boxmot_b200.synthetic import cohort_stream  # noqa: E402  (the BASELINE config 3 generator, shared with bench.py)


def test_config3_shape_ids_match_oracle_and_timing():
    import boxmot_b200 as bb
    from oracle.deepocsort import DeepOcSortOracle
    from tests.common import assert_rows_match

    dets, embs = cohort_stream(frames=9)
    gpu = bb.DeepOcSort(cap_tracks=2600, cap_dets=512)
    orc = DeepOcSortOracle()
    times = []
    for f, (d, e) in enumerate(zip(dets, embs)):
        t0 = time.perf_counter()
        got = gpu.update(d, None, e)
        times.append(time.perf_counter() - t0)
        assert_rows_match(got, orc.update(d, None, e.copy()), f)
    print("DeepOCSORT 512 dets/frame, live tracks:", len(gpu.snapshot()), "ms/frame:", [round(t * 1e3, 1) for t in times])


def test_deepocsort_with_device_reid_osnet_x1_0(tmp_path):
    import boxmot_b200 as bb
    from boxmot_b200.reid import B200ReID
    from boxmot_b200.synthetic import bench_stream, make_osnet_state
    from boxmot_b200.weights import export_blob
    from oracle import reid as orid
    from oracle.deepocsort import DeepOcSortOracle
    from tests.common import assert_rows_match

    sd = make_osnet_state("osnet_x1_0", seed=8)
    reid = B200ReID(export_blob(sd, tmp_path / "osnet_x1_0.b200reid"))
    img, frames = bench_stream(24, 8, hw=(360, 640))
    gpu = bb.DeepOcSort(reid_model=reid, cap_tracks=128, cap_dets=64)
    orc = DeepOcSortOracle(reid_model=orid.OracleReID(sd))
    for f, d in enumerate(frames):
        assert_rows_match(gpu.update(d, img), orc.update(d, img), f)
