"""BASELINE config 3 shape (DeepOCSORT, 512 dets/frame drawn from 2048 objects in 4 cohorts -> ~2000 live
tracks): ids against the oracle on the first frames, and DeepOCSORT with on-device OSNet_x1_0 ReID."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def cohort_stream(n_objects=2048, cohorts=4, frames=10, hw=(1080, 1920), seed=3, dim=512):
    rng = np.random.default_rng(seed)
    h, w = hw
    cx, cy = rng.uniform(40, w - 40, n_objects), rng.uniform(60, h - 60, n_objects)
    bw, bh = rng.uniform(20, 50, n_objects), rng.uniform(40, 100, n_objects)
    conf = rng.uniform(0.4, 0.95, n_objects)
    proto = np.abs(rng.normal(size=(n_objects, dim))).astype(np.float32)
    dets, embs = [], []
    for f in range(frames):
        idx = np.arange(f % cohorts, n_objects, cohorts)
        jx, jy = rng.normal(0, 2.0, idx.size), rng.normal(0, 2.0, idx.size)
        d = np.stack([cx[idx] + jx - bw[idx] / 2, cy[idx] + jy - bh[idx] / 2, cx[idx] + jx + bw[idx] / 2,
                      cy[idx] + jy + bh[idx] / 2, conf[idx], np.zeros(idx.size)], 1).astype(np.float32)
        e = np.maximum(proto[idx] + 0.3 * rng.normal(size=(idx.size, dim)).astype(np.float32), 0)
        dets.append(d)
        embs.append((e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32))
    return dets, embs


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
