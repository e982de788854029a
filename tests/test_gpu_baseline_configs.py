"""End-to-end parity at the BASELINE.json configuration shapes, on the device, ids bit-exact against the oracle tracker:
  C2  BoT-SORT + on-device OSNet_x0_25, 1280x720, a 256-object stress stream (births, losses, re-activations, the
      low-confidence second round), 100 frames;
  C3  DeepOCSORT + on-device OSNet_x1_0, 1920x1080, 512 detections per frame out of 2048 objects in 4 cohorts
      (~1900 live tracks), 40 frames;
  C4  StrongSORT + on-device MobileNetV2_x1_4 (1792-d), 8 x 1080p streams in one handle, 20 frames.
The oracle trackers run on the host; for C3 / C4 their ReID convolutions run through PyTorch on the GPU
(tests/common.py::TorchDeviceOracleReID -- test infrastructure, float32, TF32 off), for C2 on the host like everywhere else."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.common import BOTSORT_YAML, DEEPOCSORT_YAML, STRONGSORT_YAML, TorchDeviceOracleReID, assert_rows_match


def _blob(tmp_path, arch, seed):
    from boxmot_b200.reid import B200ReID
    from boxmot_b200.synthetic import make_mobilenetv2_state, make_osnet_state
    from boxmot_b200.weights import export_blob

    sd = make_mobilenetv2_state(1.4, seed=seed) if arch.startswith("mobilenet") else make_osnet_state(arch, seed=seed)
    return sd, B200ReID(export_blob(sd, tmp_path / f"{arch}.b200reid"))


def test_config2_botsort_osnet_x0_25_stress_256_objects_100_frames(tmp_path):
    import boxmot_b200 as bb
    from oracle import reid as orid
    from oracle.streams import stress_stream
    from oracle.trackers import BotSortOracle

    sd, reid = _blob(tmp_path, "osnet_x0_25", 21)
    img = np.random.default_rng(5).integers(0, 255, size=(720, 1280, 3), dtype=np.uint8)
    frames = stress_stream(256, 100, hw=(720, 1280), seed=23)
    gpu = bb.BotSort(reid_model=reid, cap_tracks=1024, cap_dets=256, **BOTSORT_YAML)
    orc = BotSortOracle(reid_model=orid.OracleReID(sd), **BOTSORT_YAML)
    ids, n_low = set(), 0
    for f, d in enumerate(frames):
        got = gpu.update(d, img)
        assert_rows_match(got, orc.update(d, img), f)
        ids.update(np.asarray(got)[:, 4].astype(int).tolist())
        n_low += int(((d[:, 4] > BOTSORT_YAML["track_low_thresh"]) & (d[:, 4] < BOTSORT_YAML["track_high_thresh"])).sum())
    assert len(ids) > 150 and n_low > 1000, "the stream must exercise births and the low-confidence round"


def test_config3_deepocsort_osnet_x1_0_2000_tracks_40_frames(tmp_path):
    import boxmot_b200 as bb
    from boxmot_b200.synthetic import cohort_stream
    from oracle.deepocsort import DeepOcSortOracle

    sd, reid = _blob(tmp_path, "osnet_x1_0", 22)
    img = np.random.default_rng(6).integers(0, 255, size=(1080, 1920, 3), dtype=np.uint8)
    dets, _ = cohort_stream(frames=40, conf_lo=0.55)   # all 2048 objects above det_thresh = 0.5
    gpu = bb.DeepOcSort(reid_model=reid, cap_tracks=2600, cap_dets=512, **DEEPOCSORT_YAML)
    orc = DeepOcSortOracle(reid_model=TorchDeviceOracleReID(sd), **DEEPOCSORT_YAML)
    live = 0
    for f, d in enumerate(dets):
        assert_rows_match(gpu.update(d, img), orc.update(d, img), f)
        live = max(live, len(gpu.snapshot()))
    # 2048 objects, every one above det_thresh; overlapping neighbours share or lose a track now and then
    assert live >= 1800, f"config 3 asks for ~2000 live tracks, saw {live}"


def test_config4_strongsort_mobilenetv2_8_streams_1080p(tmp_path):
    import boxmot_b200 as bb
    from boxmot_b200.synthetic import bench_stream
    from oracle.strongsort import StrongSortOracle

    sd, reid = _blob(tmp_path, "mobilenetv2_x1_4", 23)
    S, F, n = 8, 20, 64
    rng = np.random.default_rng(7)
    imgs = [rng.integers(0, 255, size=(1080, 1920, 3), dtype=np.uint8) for _ in range(S)]
    streams = [bench_stream(n, F, hw=(1080, 1920), stream=s)[1] for s in range(S)]
    gpu = bb.MultiStreamTracker("strongsort", n_streams=S, cap_tracks=256, cap_dets=n, feat_dim=1792, reid_blob=str(reid.blob_path),
                                **STRONGSORT_YAML)
    oreid = TorchDeviceOracleReID(sd)
    orcs = [StrongSortOracle(reid_model=oreid, **STRONGSORT_YAML) for _ in range(S)]
    for f in range(F):
        got = gpu.update([streams[s][f] for s in range(S)], imgs)
        for s in range(S):
            assert_rows_match(got[s], orcs[s].update(streams[s][f], imgs[s]), f)
