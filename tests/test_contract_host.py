"""Host-side contract of the Python class seam that needs no GPU: the input assertions of
basetracker.py:356-372 as the reference's own tests exercise them (tests/unit/test_trackers.py:561-598, 639-650),
`TrackResults` (track_results.py:12-31), `create_tracker` name / scope errors."""
import numpy as np
import pytest

from boxmot_b200.trackers import TRACKER_DEFAULTS, TrackResults, _check_inputs, create_tracker


def test_input_assertions_match_the_reference_messages():
    img = np.zeros((640, 640, 3), np.uint8)
    ok = np.array([[10, 10, 20, 20, 0.7, 0]], np.float32)
    _check_inputs(ok, img, None)
    _check_inputs(ok, None, np.zeros((1, 512), np.float32))
    with pytest.raises(AssertionError, match="Missmatch between detections and embeddings sizes"):
        _check_inputs(ok, img, np.random.rand(2, 512))                       # test_emb_trackers_requires_embeddings
    with pytest.raises(AssertionError, match="2nd dimension length"):
        _check_inputs(np.random.rand(2, 5), img, None)                       # test_invalid_det_array_shape
    with pytest.raises(AssertionError, match="valid format is np.ndarray"):
        _check_inputs([[1, 2, 3, 4, 0.5, 0]], img, None)
    with pytest.raises(AssertionError, match="valid number of dimensions is two"):
        _check_inputs(np.zeros((6,), np.float32), img, None)
    with pytest.raises(AssertionError, match="img_numpy"):
        _check_inputs(ok, "frame.jpg", None)


def test_track_results_views():
    rows = np.array([[1, 2, 3, 4, 7, 0.5, 2, 9], [5, 6, 7, 8, 8, 0.25, 0, 3]], np.float64)
    tr = TrackResults(rows)
    assert tr.dtype == np.float32 and tr.shape == (2, 8)
    assert tr.id.tolist() == [7, 8] and tr.cls.tolist() == [2, 0] and tr.det_ind.tolist() == [9, 3]
    assert np.array_equal(tr.xyxy, rows[:, :4].astype(np.float32)) and tr.conf.tolist() == [0.5, 0.25]
    assert TrackResults(rows[0]).shape == (1, 8)                 # a single 1-d row becomes (1, 8)
    assert TrackResults(np.array([])).shape == (0, 0)            # DeepOCSORT's empty return
    assert TrackResults(np.empty((0, 8))).shape == (0, 8)


def test_create_tracker_rejects_unknown_and_out_of_scope_names():
    # the reference's own test matches this text (tests/unit/test_trackers.py:639-650)
    with pytest.raises(ValueError, match="Unknown tracker type: 'nonexistent_tracker'"):
        create_tracker(tracker_type="nonexistent_tracker", tracker_config=None, reid_weights="x.pt", device="cpu",
                       half=False, per_class=False)
    with pytest.raises(ValueError, match="Unknown tracker type"):
        create_tracker("hybridsort")
    assert set(TRACKER_DEFAULTS) == {"bytetrack", "botsort", "deepocsort", "strongsort", "ocsort"}
    # configs/trackers/ocsort.yaml defaults
    assert TRACKER_DEFAULTS["ocsort"]["use_byte"] is False and TRACKER_DEFAULTS["ocsort"]["delta_t"] == 3
    # the YAML defaults the reference's create_tracker would read
    assert TRACKER_DEFAULTS["botsort"]["track_high_thresh"] == 0.6296854875023994
    assert TRACKER_DEFAULTS["botsort"]["removed_stracks_buffer"] == 329
    assert TRACKER_DEFAULTS["bytetrack"]["track_thresh"] == 0.6 and TRACKER_DEFAULTS["strongsort"]["min_conf"] == 0.6


def test_create_tracker_has_the_reference_signature_and_argument_resolution(tmp_path):
    """Positional callers of tracker_zoo.create_tracker keep working (tracker_type, tracker_config, reid_weights, device,
    half, per_class, evolve_param_dict, reid_preprocess, reid_model, tracker_backend), and the constructor arguments
    are assembled like tracker_zoo.py:103-147: evolve_param_dict replaces the YAML defaults wholesale, a YAML path in
    the reference's format is flattened (nested conditional parameters included), keys the reference's constructors
    swallow in **kwargs are dropped, CMC is forced off."""
    import inspect

    from boxmot_b200.trackers import BotSort, DeepOcSort, resolve_tracker_args

    names = list(inspect.signature(create_tracker).parameters)
    assert names[:10] == ["tracker_type", "tracker_config", "reid_weights", "device", "half", "per_class",
                          "evolve_param_dict", "reid_preprocess", "reid_model", "tracker_backend"]
    kind, cls, args = resolve_tracker_args("BotSort")
    assert (kind, cls) == ("botsort", BotSort) and args["use_cmc"] is False and "cmc_method" not in args
    assert args["proximity_thresh"] == 0.6084297894561342 and args["track_buffer"] == 40
    # a YAML file in the reference's layout: typed entries with `default`, conditional children nested under a parent
    y = tmp_path / "deepocsort.yaml"
    y.write_text("""
det_thresh:
  type: uniform
  default: 0.41
  range: [0.3, 0.6]
iou_thresh:
  type: uniform
  default: 0.25
asso_func:
  type: choice
  default: iou
embedding_off:
  type: choice
  default: false
  conditional:
    false:
      w_association_emb:
        type: uniform
        default: 0.66
cmc_off:
  type: choice
  default: false
""")
    kind, cls, args = resolve_tracker_args("deepocsort", str(y))
    assert cls is DeepOcSort and args["det_thresh"] == 0.41 and args["w_association_emb"] == 0.66
    assert args["cmc_off"] is True and "iou_thresh" not in args and args["asso_func"] == "iou"
    # evolve_param_dict: used INSTEAD of the YAML; everything else falls back to the constructor defaults
    _, _, args = resolve_tracker_args("botsort", None, {"track_high_thresh": 0.7, "not_a_parameter": 1, "cmc_method": "ecc"})
    assert args == {"track_high_thresh": 0.7, "use_cmc": False}
    _, _, args = resolve_tracker_args("bytetrack", None, None, {"track_thresh": 0.5, "cap_dets": 64})
    assert args["track_thresh"] == 0.5 and args["cap_dets"] == 64 and args["match_thresh"] == 0.9


def test_tracker_defaults_are_the_reference_yaml_defaults():
    """`create_tracker` must start from what the reference's `create_tracker` reads out of configs/trackers/*.yaml
    (dumped by tests/golden/make_yaml_defaults.py), not from the constructor defaults."""
    import json
    from pathlib import Path

    ref = json.loads((Path(__file__).parent / "golden" / "tracker_yaml_defaults.json").read_text())
    # documented deviations: camera-motion estimation is outside the path (CMC off), and two YAML keys the reference's
    # constructors swallow in **kwargs without effect
    deviations = {("botsort", "use_cmc"), ("botsort", "cmc_method"), ("deepocsort", "cmc_off"),
                  ("deepocsort", "iou_thresh"), ("deepocsort", "asso_func")}
    for kind, params in ref.items():
        mine = TRACKER_DEFAULTS[kind]
        for name, value in params.items():
            if (kind, name) in deviations:
                continue
            assert name in mine, f"{kind}.{name} missing"
            assert mine[name] == value, f"{kind}.{name}: {mine[name]} != YAML {value}"
    assert TRACKER_DEFAULTS["deepocsort"]["cmc_off"] is True and TRACKER_DEFAULTS["deepocsort"]["iou_threshold"] == 0.3


def test_reid_box_normalisation_follows_base_backend():
    """`B200ReID._boxes` == `BaseModelBackend._boxes_to_xyxy` (base_backend.py:125-146) on the AABB inputs of the
    reference's tests/unit/test_base_backend.py:36-43; OBB layouts (5 / 7 / 9 columns) are refused, not mis-read."""
    from boxmot_b200.reid import B200ReID

    xyxy = B200ReID._boxes(np.array([[10, 20, 30, 40, 0.9, 0]], dtype=np.float32))
    assert xyxy.shape == (1, 4) and xyxy.dtype == np.float32 and xyxy.flags["C_CONTIGUOUS"]
    np.testing.assert_array_equal(xyxy[0], np.array([10, 20, 30, 40], dtype=np.float32))
    assert B200ReID._boxes(np.array([1.0, 2.0, 3.0, 4.0])).shape == (1, 4)            # 1-d row
    assert B200ReID._boxes(np.empty((0, 6))).shape == (0, 4) and B200ReID._boxes([]).shape == (0, 4)
    assert B200ReID._boxes(np.zeros((3, 8))).shape == (3, 4)                           # tracker output rows
    with pytest.raises(ValueError, match="at least 4 coordinates"):
        B200ReID._boxes(np.zeros((2, 3)))
    for cols in (5, 7, 9):
        with pytest.raises(NotImplementedError, match="OBB"):
            B200ReID._boxes(np.zeros((1, cols)))


def test_constructor_signatures_match_the_reference():
    """Same parameter names and default values as the reference constructors (tests/golden/make_ctor_defaults.py),
    apart from the documented camera-motion switches (estimation is outside this path, so the safe value is the default)."""
    import inspect
    import json
    from pathlib import Path

    import boxmot_b200 as bb
    from boxmot_b200.trackers import _SingleStreamTracker

    ref = json.loads((Path(__file__).parent / "golden" / "tracker_ctor_defaults.json").read_text())
    deviations = {("BotSort", "use_cmc"): False, ("DeepOcSort", "cmc_off"): True}
    base = {n: p.default for n, p in inspect.signature(_SingleStreamTracker.__init__).parameters.items()}
    for name, value in ref["BaseTracker"].items():
        assert name in base and base[name] == value, f"BaseTracker.{name}"
    for cls_name in ("ByteTrack", "BotSort", "DeepOcSort", "StrongSort"):
        mine = {n: p.default for n, p in inspect.signature(getattr(bb, cls_name).__init__).parameters.items()}
        for name, value in ref[cls_name].items():
            assert name in mine, f"{cls_name}.{name} missing"
            want = deviations.get((cls_name, name), value)
            assert mine[name] == want, f"{cls_name}.{name}: {mine[name]} != {want}"


def test_track_results_exports_equal_the_reference(tmp_path):
    """`xywh`, `summary`, `to_json`, `to_csv`, `save_csv`, `save_mot` against strings produced by the reference class
    (tests/golden/make_trackresults_golden.py)."""
    import json
    from pathlib import Path

    from tests.golden.make_trackresults_golden import ROWS

    g = json.loads((Path(__file__).parent / "golden" / "trackresults_exports.json").read_text())
    tr = TrackResults(np.asarray(ROWS, dtype=np.float64))
    assert tr.summary() == g["summary"] and tr.to_json() == g["json"] and tr.to_json(indent=1) == g["json_indent"]
    assert tr.to_csv() == g["csv"] and tr.to_csv(frame_id=17) == g["csv_frame"]
    assert tr.xywh.tolist() == g["xywh"] and tr.is_obb is g["is_obb"] and tr.masks is None
    p = tmp_path / "a" / "t.csv"
    tr.save_csv(p, frame_id=3)
    tr.save_csv(p, frame_id=4)
    assert p.read_text() == g["save_csv"]
    m = tmp_path / "b" / "t.txt"
    tr.save_mot(m, frame_id=9)
    assert m.read_text() == g["save_mot"]
    assert TrackResults(np.empty((0, 8))).xywh.shape == (0, 4) and tr[:2].masks is None


def test_create_tracker_call_chain_with_a_stub_engine(monkeypatch):
    """The whole construction path (create_tracker -> tracker class -> engine arguments; TrackerRuntime.create) with the
    CUDA engine replaced by a recorder: positional reference-style calls, on-device ReID detection, warm-up only for
    foreign backends."""
    import boxmot_b200.trackers as T
    from boxmot_b200.runtime import TimingStats, TrackerRuntime

    made = []

    class Engine:
        def __init__(self, kind, n_streams, cap_tracks, cap_dets, feat_dim, reid_blob=None, **params):
            self.kind, self.params, self.feat_dim = kind, params, feat_dim
            self.with_reid, self.has_reid_model = kind != "bytetrack", reid_blob is not None
            made.append((kind, cap_tracks, cap_dets, feat_dim, reid_blob, params))

        def set_cmc(self, method):
            self.cmc = method

    monkeypatch.setattr(T, "MultiStreamTracker", Engine)

    class OnDevice:
        blob_path, feature_dim = "/tmp/m.b200reid", 1792

        def warmup(self):
            raise AssertionError("an on-device backend has nothing to warm up")

    warmed = []

    class Foreign:
        def get_features(self, xyxys, img):
            return np.zeros((len(xyxys), 512), np.float32)

        def warmup(self):
            warmed.append(1)

    rt = TrackerRuntime.create("botsort", reid_model=OnDevice(), timing_stats=TimingStats(), cap_tracks=128, cap_dets=64)
    kind, ct, cd, fd, blob, params = made[-1]
    assert (kind, ct, cd, fd, blob) == ("botsort", 128, 64, 1792, "/tmp/m.b200reid") and rt.tracker.provides_reid
    assert params["track_high_thresh"] == 0.6296854875023994 and "use_cmc" not in params
    for name in ("bytetrack", "deepocsort", "strongsort"):   # the reference tests call positionally like this
        t = T.create_tracker(name, None, None, "cpu", False, False)
        assert made[-1][0] == name and t.per_class is False
    assert made[-2][5]["det_thresh"] == 0.5 and made[-2][5]["w_association_emb"] == 0.75      # deepocsort YAML values
    T.create_tracker("deepocsort", evolve_param_dict={"det_thresh": 0.4, "embedding_off": True}, reid_weights="never_loaded.pt")
    assert made[-1][5]["det_thresh"] == 0.4 and made[-1][5]["w_association_emb"] == 0.5     # constructor default now
    t = T.create_tracker("strongsort", reid_model=Foreign())
    assert warmed == [1] and made[-1][4] is None
    assert t._engine.cmc == "ecc"   # strongsort.py:67: the reference always estimates with ECC -> on the device here
    t = T.create_tracker("botsort", evolve_param_dict={"use_cmc": True, "cmc_method": "ecc", "with_reid": False})
    assert t._engine.cmc == "ecc"
    with pytest.raises(NotImplementedError, match="only 'ecc'"):
        T.BotSort(use_cmc=True, cmc_method="sof", with_reid=False)
    with pytest.raises(NotImplementedError, match="per_class"):
        T.create_tracker("bytetrack", per_class=True)
