"""Camera-motion estimation, CPU side: oracle/cmc.py pinned on the installed OpenCV and on goldens dumped from
the unmodified reference ECC class (tests/golden/make_cmc_golden.py); the device source (boxmot_b200/csrc/cmc_ecc.cuh)
compiled for the host against both."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from boxmot_b200.synthetic import camera_pan_sequence  # noqa: E402
from oracle import cmc  # noqa: E402
from tests import hostsim as hs  # noqa: E402
from tests.common import GOLDEN as GOLDEN_DIR  # noqa: E402

CRIT = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 100, 1e-5)


def _cv2_ecc(prev, cur):
    try:
        _, w = cv2.findTransformECC(prev, cur, np.eye(2, 3, dtype=np.float32), cv2.MOTION_TRANSLATION, CRIT, None, 1)
        return 0, w[0, 2], w[1, 2]
    except cv2.error as e:
        assert e.code == cv2.Error.StsNoConv
        return 1, np.float32(0), np.float32(0)


@pytest.mark.parametrize("hw,scale", [((720, 1280), 0.15), ((1080, 1920), 0.15), ((719, 1277), 0.15), ((360, 640), 0.1), ((377, 700), 0.15), ((597, 802), 0.15),
                                      ((333, 517), 0.15)])
def test_preprocess_is_bit_exact_against_opencv(hw, scale):
    """cvtColor(BGR2GRAY) + resize(fx, fy, INTER_LINEAR) on uint8: oracle and host-compiled device source == cv2.  377 and 597
    rows are sizes where rint(rows * 0.15) rounds up and the last output row samples past the last input row: OpenCV clips the
    row indices there but keeps both weights (found by tests/tools/soak_cmc.py)."""
    img = np.random.default_rng(hw[0]).integers(0, 256, (*hw, 3), dtype=np.uint8)
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    want = cv2.resize(gray, (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(cmc.bgr2gray_u8(img), gray)
    assert np.array_equal(cmc.preprocess(img, scale), want)
    assert np.array_equal(hs.cmc_prepare(img, scale), want)


@pytest.mark.parametrize("t", [(0.0, 0.0), (0.3, -0.7), (1.51, 2.49), (-3.26, 0.015), (5.999, -6.001), (0.015625, 0.046875)])
def test_warp_affine_restatement_is_bit_exact(t):
    """warpAffine(INTER_LINEAR | WARP_INVERSE_MAP) of a float image by a translation: 1/32-pixel fixed point, zero border;
    INTER_NEAREST for the validity mask."""
    a = (np.random.default_rng(3).random((108, 192)) * 255).astype(np.float32)
    m = np.array([[1, 0, t[0]], [0, 1, t[1]]], np.float32)
    want = cv2.warpAffine(a, m, (192, 108), flags=cv2.INTER_LINEAR + cv2.WARP_INVERSE_MAP)
    assert np.array_equal(cmc.warp_translate_f32(a, np.float32(t[0]), np.float32(t[1])), want)
    mask = cv2.warpAffine(np.ones((108, 192), np.uint8), m, (192, 108), flags=cv2.INTER_NEAREST + cv2.WARP_INVERSE_MAP)
    assert np.array_equal(cmc.warp_mask(108, 192, np.float32(t[0]), np.float32(t[1])), mask.astype(bool))


def test_ecc_restatements_match_opencv_on_a_panning_camera():
    """findTransformECC (translation) on consecutive frames of the synthetic pan: oracle and host-compiled device source
    against cv2, 2e-6 registration pixels (observed 5e-7); the estimates recover the true camera motion."""
    frames, _, offs, _ = camera_pan_sequence(10)
    regs = [cmc.preprocess(f) for f in frames]
    for f in range(1, len(regs)):
        st, tx, ty = _cv2_ecc(regs[f - 1], regs[f])
        assert st == 0
        _, ox, oy = cmc.find_transform_ecc_translation(regs[f - 1], regs[f])
        hst, hx, hy = hs.ecc_translation(regs[f - 1], regs[f])
        assert hst == 0
        assert max(abs(ox - tx), abs(oy - ty), abs(hx - tx), abs(hy - ty)) < 2e-6
        true = -(offs[f] - offs[f - 1]) * 0.15
        assert abs(tx - true[0]) < 0.08 and abs(ty - true[1]) < 0.08


@pytest.mark.parametrize("kind", ["constant", "inverted", "flipped"])
def test_ecc_non_convergence_is_reported_like_opencv(kind):
    """cv2 raises StsNoConv on a constant image (NaN correlation) and on anti-correlated images (lambda_d <= 0): the oracle
    raises NoConvergence, the device source returns status 1."""
    a = cmc.preprocess(camera_pan_sequence(1, hw=(360, 640), seed=5)[0][0])
    b = {"constant": np.full_like(a, 128), "inverted": 255 - a, "flipped": a[::-1, ::-1].copy()}[kind]
    assert _cv2_ecc(a, b)[0] == 1
    with pytest.raises(cmc.NoConvergence):
        cmc.find_transform_ecc_translation(a, b)
    assert hs.ecc_translation(a, b)[0] == 1


def test_ecc_oracle_reproduces_the_reference_class():
    """EccOracle.apply == boxmot.motion.cmc.ecc.ECC.apply (golden from the unmodified reference): the synthetic pan frame by
    frame (first frame identity, translation divided by the scale in float32), and the reference's MOT17-mini frames
    through their registration images."""
    g = np.load(GOLDEN_DIR / "cmc_ecc.npz")
    frames, _, offs, _ = camera_pan_sequence(12)
    assert np.array_equal(offs, g["pan_offsets"])
    orc = cmc.EccOracle()
    for f, im in enumerate(frames):
        w = orc.apply(im)
        assert w.dtype == np.float32 and w.shape == (2, 3)
        np.testing.assert_allclose(w, g["pan_warps"][f], rtol=0, atol=2e-5)
    regs, warps = g["mot17_reg"], g["mot17_warps"]
    for k in range(1, len(regs)):
        if not np.any(warps[k][:, 2]):   # first frame of a sequence: identity
            continue
        _, tx, ty = cmc.find_transform_ecc_translation(regs[k - 1], regs[k])
        st, hx, hy = hs.ecc_translation(regs[k - 1], regs[k])
        want = warps[k][:, 2]
        got = np.array([tx, ty], np.float32) / np.float32(0.15)
        hgot = np.array([hx, hy], np.float32) / np.float32(0.15)
        assert st == 0 and np.abs(got - want).max() < 2e-6 and np.abs(hgot - want).max() < 2e-6
