"""Camera-motion estimation on the device: the ECC estimator against the installed OpenCV (the third-party
arithmetic the reference calls) and, inside StrongSORT / BoT-SORT, against the oracle trackers fed with the warps the
reference's ECC class semantics produce from cv2.  Sorts last on purpose (first run on hardware at the end of round 2)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

cv2 = pytest.importorskip("cv2")

from boxmot_b200.synthetic import camera_pan_sequence  # noqa: E402

CRIT = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 100, 1e-5)


class Cv2Ecc:
    """boxmot/motion/cmc/ecc.py:46-108 with its defaults, on the installed OpenCV (test-side reference)."""

    def __init__(self, scale=0.15):
        self.scale, self.prev = scale, None

    def apply(self, img):
        w = np.eye(2, 3, dtype=np.float32)
        cur = cv2.resize(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), (0, 0), fx=self.scale, fy=self.scale,
                         interpolation=cv2.INTER_LINEAR)
        if self.prev is None:
            self.prev = cur
            return w
        try:
            _, w = cv2.findTransformECC(self.prev, cur, w, cv2.MOTION_TRANSLATION, CRIT, None, 1)
        except cv2.error as e:
            assert e.code == cv2.Error.StsNoConv
            self.prev = cur
            return w
        w = w.copy()
        w[0, 2] /= self.scale
        w[1, 2] /= self.scale
        self.prev = cur
        return w


def _device_ecc(prev, cur, scale=0.15):
    from boxmot_b200 import _lib

    lib = _lib.require_device()
    rows, cols = prev.shape[:2]
    h, w = int(np.rint(rows * scale)), int(np.rint(cols * scale))
    warp = np.zeros((2, 3), np.float32)
    reg = np.zeros((h, w), np.uint8)
    st = ctypes.c_int(-1)
    prev, cur = np.ascontiguousarray(prev), np.ascontiguousarray(cur)
    ok = lib.boxmot_b200_cmc_ecc(prev.ctypes.data, cur.ctypes.data, rows, cols, scale, 1e-5, 100, warp.ctypes.data,
                                 ctypes.byref(st), reg.ctypes.data)
    assert ok == 1, _lib.last_error(lib)
    return st.value, warp, reg


@pytest.mark.parametrize("hw,seed", [((360, 640), 5), ((720, 1280), 6), ((1080, 1920), 7), ((475, 801), 8)])
def test_device_ecc_matches_opencv(hw, seed):
    """Registration image bit-exact; warp of ECC().apply within 1e-4 (relative to max(1, |t|)) of OpenCV's -- observed
    ~1e-6; the iteration follows the same 1/32-pixel quantised path."""
    frames, _, _, _ = camera_pan_sequence(5, hw=hw, seed=seed)
    worst = 0.0
    for a, b in zip(frames[:-1], frames[1:]):
        ref = Cv2Ecc()
        ref.apply(a)
        want = ref.apply(b)
        st, got, reg = _device_ecc(a, b)
        assert np.array_equal(reg, ref.prev), "BaseCMC.preprocess must be bit-exact"
        assert st == 0 and np.any(want[:, 2] != 0)
        assert np.array_equal(got[:, :2], np.eye(2, dtype=np.float32))
        err = np.abs(got[:, 2] - want[:, 2]) / np.maximum(1.0, np.abs(want[:, 2]))
        worst = max(worst, float(err.max()))
    print("device ECC vs cv2, worst relative warp error:", worst)
    assert worst < 1e-4


@pytest.mark.parametrize("kind", ["constant", "inverted"])
def test_device_ecc_reports_non_convergence_as_identity(kind):
    """Both StsNoConv exits of cv::findTransformECC: a constant frame (zero variance -> NaN correlation) and the inverted
    frame (anti-correlated -> lambda_d <= 0).  ECC.apply then returns the identity (ecc.py:69-79); so does the device."""
    a = camera_pan_sequence(1, hw=(360, 640), seed=5)[0][0]
    b = np.full_like(a, 128) if kind == "constant" else 255 - a
    ref = Cv2Ecc()
    ref.apply(a)
    assert np.array_equal(ref.apply(b), np.eye(2, 3, dtype=np.float32))   # cv2 raised StsNoConv
    st, got, _ = _device_ecc(a, b)
    assert st == 1 and np.array_equal(got, np.eye(2, 3, dtype=np.float32))


def test_strongsort_with_device_ecc_matches_oracle_with_opencv_ecc():
    """StrongSort(cmc="ecc"): estimator gated on live tracks like strongsort.py:83-86, warp consumed by camera_update in
    the same frame.  Oracle: StrongSortOracle with the warp cv2 gives for the same frames."""
    import boxmot_b200 as bb
    from oracle.strongsort import StrongSortOracle
    from tests.common import assert_rows_match

    frames, dets, _, embs = camera_pan_sequence(24, dim=512, seed=11)
    kw = dict(min_conf=0.3, max_cos_dist=0.4, n_init=2)
    gpu = bb.StrongSort(cmc="ecc", cap_tracks=128, cap_dets=64, **kw)
    orc, ref = StrongSortOracle(**kw), Cv2Ecc()
    moved = 0
    for f, (im, d, e) in enumerate(zip(frames, dets, embs)):
        if f in (7, 8):   # two empty frames in the middle: tracks survive, the estimator keeps running
            d, e = d[:0], e[:0]
        warp = ref.apply(im) if len(orc.tracks) >= 1 else None
        moved += warp is not None and bool(np.any(warp[:, 2] != 0))
        want = orc.update(d, im, e, warp=warp)
        assert_rows_match(gpu.update(d, im, e), want, f)
    assert moved >= 15
    gpu.reset()   # ECC.prev_img is dropped with the tracks
    orc, ref = StrongSortOracle(**kw), Cv2Ecc()
    for f in range(4):
        warp = ref.apply(frames[f]) if len(orc.tracks) >= 1 else None
        assert_rows_match(gpu.update(dets[f], frames[f], embs[f]), orc.update(dets[f], frames[f], embs[f], warp=warp), f)


def test_botsort_with_device_ecc_matches_oracle_with_opencv_ecc():
    """BotSort(use_cmc=True, cmc_method="ecc"): the estimator runs on every frame (botsort.py:142), multi_gmc applies it."""
    import boxmot_b200 as bb
    from oracle.trackers import BotSortOracle
    from tests.common import assert_rows_match

    frames, dets, _, _ = camera_pan_sequence(24, seed=12)
    kw = dict(with_reid=False, track_high_thresh=0.6, new_track_thresh=0.65)
    gpu = bb.BotSort(use_cmc=True, cmc_method="ecc", cap_tracks=128, cap_dets=64, **kw)
    orc, ref = BotSortOracle(**kw), Cv2Ecc()
    for f, (im, d) in enumerate(zip(frames, dets)):
        warp = ref.apply(im)
        assert_rows_match(gpu.update(d, im), orc.update(d, im, warp=warp), f)
    with pytest.raises(Exception, match="frame"):
        gpu.update(dets[0], None)
    with pytest.raises(NotImplementedError):
        bb.BotSort(use_cmc=True, cmc_method="sof", with_reid=False)


def test_reference_abi_botsort_accepts_cmc_method_ecc():
    """boxmot_botsort_create with cmc_method="ecc" (the field of botsort/c_api.hpp:17-33 the previous rounds rejected)."""
    from boxmot_b200 import _lib

    lib = _lib.require_device()
    cfg = _lib.BoxMOTBotSortConfig()
    cfg.track_high_thresh, cfg.track_low_thresh, cfg.new_track_thresh = 0.6, 0.1, 0.65
    cfg.track_buffer, cfg.match_thresh, cfg.proximity_thresh, cfg.appearance_thresh = 30, 0.8, 0.5, 0.25
    cfg.cmc_method, cfg.frame_rate, cfg.fuse_first_associate, cfg.with_reid, cfg.max_obs = b"ecc", 30, 0, 0, 50
    h = lib.boxmot_botsort_create(ctypes.byref(cfg))
    assert h, _lib.last_error(lib)
    lib.boxmot_botsort_destroy(h)
    cfg.cmc_method = b"sof"
    assert not lib.boxmot_botsort_create(ctypes.byref(cfg))
