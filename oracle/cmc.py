"""Oracle restatement of the reference's ECC camera-motion estimator -- TEST INFRASTRUCTURE ONLY.

Reference: boxmot/motion/cmc/ecc.py:14-108 (`ECC.apply`: preprocess -> cv2.findTransformECC(prev, curr, eye(2,3),
MOTION_TRANSLATION, (EPS|COUNT, 100, 1e-5), None, 1) -> translation divided by the scale) and
boxmot/motion/cmc/base_cmc.py:29-60 (`preprocess`: cv2.cvtColor BGR2GRAY, cv2.resize(fx=fy=scale, INTER_LINEAR)).
Callers: StrongSORT always (strongsort.py:67,83-86, only on frames that start with at least one track), BoT-SORT when
`cmc_method == "ecc"` (botsort.py:116-117,142; the constructor default).

The arithmetic lives in OpenCV (third party, installed here and on the GPU box: cv2 4.13): this file restates
`cv::findTransformECC` (modules/video/src/ecc.cpp) for MOTION_TRANSLATION without an input mask and with
gaussFiltSize = 1, `cv::warpAffine` (INTER_LINEAR | WARP_INVERSE_MAP on float images: source coordinates in 10-bit fixed
point rounded to 1/32 pixel, constant-zero border; INTER_NEAREST for the mask), `cv::cvtColor(BGR2GRAY)` on uint8 and
the uint8 `cv::resize`.  PINNED against the installed cv2 in tests/test_oracle_cmc.py (gray / resize bit for bit, warp
to 1e-6 on seeded image pairs and on MOT17-mini frame pairs).
"""
from __future__ import annotations

import numpy as np

AB_BITS = 10
AB_SCALE = 1 << AB_BITS
INTER_BITS = 5
INTER_TAB = 1 << INTER_BITS


def bgr2gray_u8(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, COLOR_BGR2GRAY) on uint8: 15-bit fixed-point weights, round to nearest."""
    b = img[..., 0].astype(np.int64)
    g = img[..., 1].astype(np.int64)
    r = img[..., 2].astype(np.int64)
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def _coeffs(dst_n: int, src_n: int, scale: float, clamp: bool = True):
    """Source index and 11-bit weights of cv2.resize(INTER_LINEAR) on uint8 for an explicit scale = 1 / fx.  OpenCV resets the
    weights at the border only along x (`clamp`); along y the weights stay and the row indices are clipped."""
    idx = np.zeros(dst_n, np.int64)
    a0 = np.zeros(dst_n, np.int64)
    a1 = np.zeros(dst_n, np.int64)
    for d in range(dst_n):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp and s < 0:
            s, f = 0, np.float32(0)
        if clamp and s >= src_n - 1:
            s, f = src_n - 1, np.float32(0)
        idx[d] = s
        a0[d] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
        a1[d] = int(np.rint(np.float32(f * np.float32(2048))))
    return idx, a0, a1


def scaled_size(rows: int, cols: int, scale: float):
    """dsize of cv2.resize(src, (0, 0), fx=scale, fy=scale): saturate_cast<int>(n * f) rounds half to even."""
    return int(np.rint(rows * scale)), int(np.rint(cols * scale))


def resize_gray_u8(src: np.ndarray, scale: float) -> np.ndarray:
    """cv2.resize(gray_u8, (0, 0), fx=scale, fy=scale, interpolation=INTER_LINEAR)."""
    sh, sw = src.shape
    dh, dw = scaled_size(sh, sw, scale)
    inv = 1.0 / scale
    xi, xa0, xa1 = _coeffs(dw, sw, inv)
    yi, ya0, ya1 = _coeffs(dh, sh, inv, clamp=False)
    s = src.astype(np.int64)
    x1 = np.minimum(xi + 1, sw - 1)
    hor = s[:, xi] * xa0[None, :] + s[:, x1] * xa1[None, :]
    y0 = np.clip(yi, 0, sh - 1)
    y1 = np.clip(yi + 1, 0, sh - 1)
    out = (((ya0[:, None] * (hor[y0] >> 4)) >> 16) + ((ya1[:, None] * (hor[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def preprocess(img: np.ndarray, scale: float = 0.15) -> np.ndarray:
    """BaseCMC.preprocess (base_cmc.py:29-60) with grayscale=True and a float scale."""
    return resize_gray_u8(bgr2gray_u8(img), scale)


def _saturate_int(x: float) -> int:
    return int(np.rint(x))   # cvRound: round half to even


def _shift_params(t: np.float32, n: int, nearest: bool):
    """Integer source offset and 1/32 fraction of warpAffine for the matrix row [1 0 t] (or [0 1 t])."""
    delta = AB_SCALE // 2 if nearest else AB_SCALE // INTER_TAB // 2
    x0 = _saturate_int(float(t) * AB_SCALE) + delta
    k = np.arange(n, dtype=np.int64)
    adelta = np.rint(k.astype(np.float64) * AB_SCALE).astype(np.int64)   # m00 = 1
    if nearest:
        return (x0 + adelta) >> AB_BITS, None
    X = (x0 + adelta) >> (AB_BITS - INTER_BITS)
    return X >> INTER_BITS, X & (INTER_TAB - 1)


def warp_translate_f32(src: np.ndarray, tx: np.float32, ty: np.float32) -> np.ndarray:
    """cv2.warpAffine(src_f32, [[1,0,tx],[0,1,ty]], size, INTER_LINEAR | WARP_INVERSE_MAP), constant-zero border."""
    h, w = src.shape
    sx, ax = _shift_params(tx, w, False)
    sy, ay = _shift_params(ty, h, False)

    def at(yy, xx):
        ok = (yy >= 0) & (yy < h)
        okx = (xx >= 0) & (xx < w)
        v = src[np.clip(yy, 0, h - 1)[:, None], np.clip(xx, 0, w - 1)[None, :]]
        return np.where(ok[:, None] & okx[None, :], v, np.float32(0))

    fx = (ax.astype(np.float32) / np.float32(INTER_TAB))
    fy = (ay.astype(np.float32) / np.float32(INTER_TAB))
    w00 = ((np.float32(1) - fy)[:, None] * (np.float32(1) - fx)[None, :]).astype(np.float32)
    w01 = ((np.float32(1) - fy)[:, None] * fx[None, :]).astype(np.float32)
    w10 = (fy[:, None] * (np.float32(1) - fx)[None, :]).astype(np.float32)
    w11 = (fy[:, None] * fx[None, :]).astype(np.float32)
    out = at(sy, sx) * w00 + at(sy, sx + 1) * w01 + at(sy + 1, sx) * w10 + at(sy + 1, sx + 1) * w11
    return out.astype(np.float32)


def warp_mask(h: int, w: int, tx: np.float32, ty: np.float32) -> np.ndarray:
    """cv2.warpAffine(ones_u8, M, size, INTER_NEAREST | WARP_INVERSE_MAP): 1 where the source pixel exists."""
    sx, _ = _shift_params(tx, w, True)
    sy, _ = _shift_params(ty, h, True)
    return ((sy >= 0) & (sy < h))[:, None] & ((sx >= 0) & (sx < w))[None, :]


def gradients(img: np.ndarray):
    """filter2D with [-0.5, 0, 0.5] (and its transpose), BORDER_REFLECT_101."""
    p = np.pad(img, 1, mode="reflect")
    gx = (p[1:-1, 2:] * np.float32(0.5) - p[1:-1, :-2] * np.float32(0.5)).astype(np.float32)
    gy = (p[2:, 1:-1] * np.float32(0.5) - p[:-2, 1:-1] * np.float32(0.5)).astype(np.float32)
    return gx, gy


class NoConvergence(Exception):
    """cv2.error StsNoConv (NaN correlation or lambda_d <= 0): ECC.apply returns the identity (ecc.py:69-79)."""


def find_transform_ecc_translation(template_u8: np.ndarray, image_u8: np.ndarray, eps: float = 1e-5, max_iter: int = 100):
    """cv::findTransformECC(template, image, eye(2,3), MOTION_TRANSLATION, (COUNT|EPS, max_iter, eps), noArray(), 1).

    Returns (rho, tx, ty) with tx, ty float32 (warp[0,2], warp[1,2])."""
    T = template_u8.astype(np.float32)
    I = image_u8.astype(np.float32)
    h, w = T.shape
    gx, gy = gradients(I)
    tx = np.float32(0)
    ty = np.float32(0)
    rho, last_rho = -1.0, -eps
    it = 1
    while it <= max_iter and abs(rho - last_rho) >= eps:
        Iw = warp_translate_f32(I, tx, ty)
        gxw = warp_translate_f32(gx, tx, ty)
        gyw = warp_translate_f32(gy, tx, ty)
        m = warp_mask(h, w, tx, ty)
        n = int(m.sum())
        if n:   # cv::meanStdDev: float64 sum and sum of squares, std = sqrt(max(sq / n - mean^2, 0))
            iw64, t64 = Iw[m].astype(np.float64), T[m].astype(np.float64)
            i_mean, t_mean = float(iw64.sum() / n), float(t64.sum() / n)
            i_std = float(np.sqrt(max((iw64 * iw64).sum() / n - i_mean * i_mean, 0.0)))
            t_std = float(np.sqrt(max((t64 * t64).sum() / n - t_mean * t_mean, 0.0)))
        else:
            i_mean = t_mean = i_std = t_std = 0.0
        # cv::subtract(Mat32f, Scalar, dst, mask) computes in float32; pixels outside the mask keep their value
        Iz = np.where(m, Iw - np.float32(i_mean), Iw).astype(np.float32)
        Tz = np.where(m, T - np.float32(t_mean), np.float32(0)).astype(np.float32)
        t_norm = np.sqrt(n * t_std * t_std)
        i_norm = np.sqrt(n * i_std * i_std)
        d = lambda a, b: float(np.dot(a.astype(np.float64).ravel(), b.astype(np.float64).ravel()))   # Mat::dot
        H = np.array([[d(gxw, gxw), d(gxw, gyw)], [d(gxw, gyw), d(gyw, gyw)]], np.float32)
        Hd = H.astype(np.float64)
        det = Hd[0, 0] * Hd[1, 1] - Hd[0, 1] * Hd[1, 0]
        if det == 0.0:
            Hinv = np.zeros((2, 2), np.float32)
        else:
            Hinv = (np.array([[Hd[1, 1], -Hd[0, 1]], [-Hd[1, 0], Hd[0, 0]]]) * (1.0 / det)).astype(np.float32)
        corr = d(Tz, Iz)
        last_rho = rho
        rho = corr / (i_norm * t_norm) if (i_norm * t_norm) != 0.0 else float("nan")
        if np.isnan(rho):
            raise NoConvergence("NaN encountered.")
        ip = np.array([d(gxw, Iz), d(gyw, Iz)], np.float32)
        tp = np.array([d(gxw, Tz), d(gyw, Tz)], np.float32)
        iph = (Hinv @ ip).astype(np.float32)
        lam_n = i_norm * i_norm - float(np.dot(ip.astype(np.float64), iph.astype(np.float64)))
        lam_d = corr - float(np.dot(tp.astype(np.float64), iph.astype(np.float64)))
        if lam_d <= 0.0:
            raise NoConvergence("The algorithm stopped before its convergence.")
        lam = lam_n / lam_d
        err = (np.float32(lam) * Tz - Iz).astype(np.float32)
        ep = np.array([d(gxw, err), d(gyw, err)], np.float32)
        dp = (Hinv @ ep).astype(np.float32)
        tx = np.float32(tx + dp[0])
        ty = np.float32(ty + dp[1])
        it += 1
    return rho, tx, ty


class EccOracle:
    """ECC.apply (ecc.py:46-108) for the default arguments: translation, eps 1e-5, 100 iterations, scale 0.15."""

    def __init__(self, eps: float = 1e-5, max_iter: int = 100, scale: float = 0.15):
        self.eps, self.max_iter, self.scale = float(eps), int(max_iter), float(scale)
        self.prev = None

    def apply(self, img: np.ndarray, dets=None) -> np.ndarray:
        warp = np.eye(2, 3, dtype=np.float32)
        cur = preprocess(img, self.scale)
        if self.prev is None:
            self.prev = cur
            return warp
        try:
            _, tx, ty = find_transform_ecc_translation(self.prev, cur, self.eps, self.max_iter)
        except NoConvergence:
            self.prev = cur
            return warp
        warp[0, 2] = tx
        warp[1, 2] = ty
        if self.scale < 1.0:
            warp[0, 2] /= self.scale
            warp[1, 2] /= self.scale
        self.prev = cur
        return warp
