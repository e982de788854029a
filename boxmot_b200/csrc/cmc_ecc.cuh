// cmc_ecc.cuh -- camera-motion estimation on the device: the reference's ECC estimator.
//
// Replaces boxmot/motion/cmc/ecc.py:46-108 (`ECC.apply` with its defaults: MOTION_TRANSLATION, eps 1e-5, 100 iterations,
// scale 0.15, grayscale) and boxmot/motion/cmc/base_cmc.py:29-60 (`preprocess`: cv2.cvtColor BGR2GRAY + cv2.resize
// fx = fy = scale, INTER_LINEAR).  Callers in the reference: StrongSORT on every frame that starts with at least one
// track (strongsort.py:67,83-86), BoT-SORT with cmc_method "ecc" (botsort.py:116-117,142).
//
// The arithmetic is OpenCV's (third party): cv::findTransformECC for a translation without mask and blur, cv::warpAffine
// (INTER_LINEAR | WARP_INVERSE_MAP on float images: source coordinates in 10-bit fixed point rounded to 1/32 pixel --
// for a translation every pixel shares one integer offset and one pair of 1/32 fractions --, zero border; nearest for
// the validity mask), the uint8 gray conversion (15-bit weights) and the uint8 bilinear resize (11-bit weights).
// oracle/cmc.py restates the same steps and is pinned on the installed cv2; this file follows that restatement step by
// step so that the iteration takes the same path (same 1/32-pixel quantisation, float32 image arithmetic, float64
// sums).  Same source compiles for the host simulation (tests/_hostsim) like the tracker cores.
//
// Data: the registration image is rows*scale x cols*scale (108 x 192 at 720p, 162 x 288 at 1080p): uint8, 20-47 KB per
// stream, L1 resident.  One CTA per stream; an iteration is three passes over the pixels (moments -> projections ->
// error projection) with a block reduction of float64 partial sums after each; the gradient of the input image is
// recomputed from the uint8 pixels at every tap instead of being stored (exact: 0.5 * u8 differences).
#pragma once
#include <stdint.h>

#include "tracker_core.cuh"

namespace bmb {

#if BMB_DEVICE
#define BMB_D2I_RN(x) __double2int_rn(x)
#define BMB_F2I_RN(x) __float2int_rn(x)
#else
#define BMB_D2I_RN(x) ((int)nearbyint(x))
#define BMB_F2I_RN(x) ((int)nearbyintf(x))
#endif

// source index and 11-bit weights of cv2.resize(INTER_LINEAR) on uint8 (oracle/cmc.py::_coeffs).  OpenCV resets the
// weights at the image border only along x (`clamp`); along y both taps keep their weights and the ROW INDICES are clipped
// (it matters when rint(rows * scale) rounds up and the last output row samples past the last input row).
BMB_FN void cmc_coeff(int d, int src_n, double inv_scale, bool clamp, int& idx, int& a0, int& a1) {
    float f = (float)((d + 0.5) * inv_scale - 0.5);
    int s = (int)floorf(f);
    f = f - (float)s;
    if (clamp && s < 0) { s = 0; f = 0.f; }
    if (clamp && s >= src_n - 1) { s = src_n - 1; f = 0.f; }
    idx = s;
    a0 = BMB_F2I_RN((1.0f - f) * 2048.0f);
    a1 = BMB_F2I_RN(f * 2048.0f);
}

BMB_FN int cmc_gray(const uint8_t* p) {   // BGR -> gray, cv2.cvtColor on uint8
    return ((int)p[0] * 3735 + (int)p[1] * 19235 + (int)p[2] * 9798 + (1 << 14)) >> 15;
}

// one pixel of preprocess(): gray of the four source pixels, horizontal then vertical fixed-point blend
BMB_FN uint8_t cmc_prepare_pixel(const uint8_t* img, int rows, int cols, double inv_scale, int dy, int dx) {
    int xi, xa0, xa1, yi, ya0, ya1;
    cmc_coeff(dx, cols, inv_scale, true, xi, xa0, xa1);
    cmc_coeff(dy, rows, inv_scale, false, yi, ya0, ya1);
    const int x1 = xi + 1 < cols ? xi + 1 : cols - 1;
    const int y0 = yi < 0 ? 0 : (yi > rows - 1 ? rows - 1 : yi);
    const int y1 = yi + 1 < 0 ? 0 : (yi + 1 > rows - 1 ? rows - 1 : yi + 1);
    const uint8_t* r0 = img + (size_t)y0 * cols * 3;
    const uint8_t* r1 = img + (size_t)y1 * cols * 3;
    const int hor0 = cmc_gray(r0 + 3 * xi) * xa0 + cmc_gray(r0 + 3 * x1) * xa1;
    const int hor1 = cmc_gray(r1 + 3 * xi) * xa0 + cmc_gray(r1 + 3 * x1) * xa1;
    int out = (((ya0 * (hor0 >> 4)) >> 16) + ((ya1 * (hor1 >> 4)) >> 16) + 2) >> 2;
    out = out < 0 ? 0 : (out > 255 ? 255 : out);
    return (uint8_t)out;
}

// sums of K float64 partials over the CTA; every thread leaves with the same totals (fixed summation order)
template <int K>
BMB_FN void ecc_block_sum(double (&v)[K], double* red) {
#if BMB_DEVICE
#pragma unroll
    for (int k = 0; k < K; ++k)
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    if (BMB_LANE == 0)
#pragma unroll
        for (int k = 0; k < K; ++k) red[BMB_WARP * K + k] = v[k];
    BMB_SYNC();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a = 0.0;
        for (int wq = 0; wq < BMB_NW; ++wq) a += red[wq * K + k];
        v[k] = a;
    }
    BMB_SYNC();
#else
    (void)v; (void)red;
#endif
}

struct EccTap {   // the warp of one iteration: integer offsets, bilinear weights, nearest offsets of the mask
    int ox, oy, mx, my;
    float w00, w01, w10, w11;
};

BMB_FN EccTap ecc_tap(float tx, float ty) {
    EccTap t;
    const int X0 = BMB_D2I_RN((double)tx * 1024.0), Y0 = BMB_D2I_RN((double)ty * 1024.0);   // saturate_cast<int>(m * AB_SCALE)
    const int qx = (X0 + 16) >> 5, qy = (Y0 + 16) >> 5;   // + AB_SCALE / INTER_TAB_SIZE / 2, >> (AB_BITS - INTER_BITS)
    t.ox = qx >> 5; t.oy = qy >> 5;
    const float fx = (float)(qx & 31) * (1.0f / 32.0f), fy = (float)(qy & 31) * (1.0f / 32.0f);
    t.w00 = (1.0f - fy) * (1.0f - fx);
    t.w01 = (1.0f - fy) * fx;
    t.w10 = fy * (1.0f - fx);
    t.w11 = fy * fx;
    t.mx = (X0 + 512) >> 10; t.my = (Y0 + 512) >> 10;   // INTER_NEAREST: + AB_SCALE / 2, >> AB_BITS
    return t;
}

// input image and its [-0.5 0 0.5] gradients (BORDER_REFLECT_101) at source pixel (yy, xx); zero outside the image
BMB_FN void ecc_src(const uint8_t* I, int h, int w, int yy, int xx, float& v, float& gx, float& gy) {
    if (yy < 0 || yy >= h || xx < 0 || xx >= w) { v = 0.f; gx = 0.f; gy = 0.f; return; }
    const uint8_t* row = I + (size_t)yy * w;
    v = (float)row[xx];
    const int xl = xx == 0 ? 1 : xx - 1, xr = xx == w - 1 ? w - 2 : xx + 1;
    const int yu = yy == 0 ? 1 : yy - 1, yd = yy == h - 1 ? h - 2 : yy + 1;
    gx = (float)row[xr] * 0.5f - (float)row[xl] * 0.5f;
    gy = (float)I[(size_t)yd * w + xx] * 0.5f - (float)I[(size_t)yu * w + xx] * 0.5f;
}

BMB_FN void ecc_sample(const uint8_t* I, int h, int w, const EccTap& t, int y, int x, float& iw, float& gxw, float& gyw) {
    float v00, v01, v10, v11, a00, a01, a10, a11, b00, b01, b10, b11;
    const int sy = y + t.oy, sx = x + t.ox;
    ecc_src(I, h, w, sy, sx, v00, a00, b00);
    ecc_src(I, h, w, sy, sx + 1, v01, a01, b01);
    ecc_src(I, h, w, sy + 1, sx, v10, a10, b10);
    ecc_src(I, h, w, sy + 1, sx + 1, v11, a11, b11);
    iw = v00 * t.w00 + v01 * t.w01 + v10 * t.w10 + v11 * t.w11;
    gxw = a00 * t.w00 + a01 * t.w01 + a10 * t.w10 + a11 * t.w11;
    gyw = b00 * t.w00 + b01 * t.w01 + b10 * t.w10 + b11 * t.w11;
}

// cv::findTransformECC(T, I, eye(2,3), MOTION_TRANSLATION, (COUNT | EPS, max_iter, eps), noArray(), 1).
// Returns 0 and the translation (float32, registration-image pixels) in txy, or 1 when OpenCV would raise StsNoConv
// (NaN correlation, lambda_d <= 0): the reference then keeps the identity (ecc.py:69-79).  `red`: [BMB_NW * 8] doubles.
BMB_FN int ecc_translation(const uint8_t* T, const uint8_t* I, int h, int w, double eps, int max_iter, double* red, float* txy) {
    float tx = 0.f, ty = 0.f;
    double rho = -1.0, last_rho = -eps;
    const int npx = h * w;
    for (int it = 1; it <= max_iter && fabs(rho - last_rho) >= eps; ++it) {
        const EccTap t = ecc_tap(tx, ty);
        // pass A: moments of the warped image and of the template over the valid pixels (cv::meanStdDev with the mask)
        double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        for (int p = BMB_TID; p < npx; p += BMB_NT) {
            const int y = p / w, x = p - y * w;
            const int my = y + t.my, mx = x + t.mx;
            if (my < 0 || my >= h || mx < 0 || mx >= w) continue;
            float iw, gxw, gyw;
            ecc_sample(I, h, w, t, y, x, iw, gxw, gyw);
            const double di = (double)iw, dt = (double)T[p];
            a[0] += 1.0; a[1] += di; a[2] += di * di; a[3] += dt; a[4] += dt * dt;
        }
        ecc_block_sum<5>(a, red);
        const double n = a[0];
        const double i_mean = n > 0.0 ? a[1] / n : 0.0, t_mean = n > 0.0 ? a[3] / n : 0.0;
        double i_var = n > 0.0 ? a[2] / n - i_mean * i_mean : 0.0, t_var = n > 0.0 ? a[4] / n - t_mean * t_mean : 0.0;
        const double i_std = sqrt(i_var > 0.0 ? i_var : 0.0), t_std = sqrt(t_var > 0.0 ? t_var : 0.0);
        const double i_norm = sqrt(n * i_std * i_std), t_norm = sqrt(n * t_std * t_std);
        const float i_mean_f = (float)i_mean, t_mean_f = (float)t_mean;   // subtract(Mat32f, Scalar): float32 arithmetic
        // pass B: Hessian, correlation, projections of the zero-mean images onto the Jacobian (= warped gradients)
        double b[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int p = BMB_TID; p < npx; p += BMB_NT) {
            const int y = p / w, x = p - y * w;
            const int my = y + t.my, mx = x + t.mx;
            const bool m = !(my < 0 || my >= h || mx < 0 || mx >= w);
            float iw, gxw, gyw;
            ecc_sample(I, h, w, t, y, x, iw, gxw, gyw);
            const float iz = m ? iw - i_mean_f : iw;               // outside the mask the warped image keeps its value
            const float tz = m ? (float)T[p] - t_mean_f : 0.f;     // ... and the zero-mean template is zero
            const double gx = (double)gxw, gy = (double)gyw, di = (double)iz, dt = (double)tz;
            b[0] += gx * gx; b[1] += gx * gy; b[2] += gy * gy; b[3] += dt * di;
            b[4] += gx * di; b[5] += gy * di; b[6] += gx * dt; b[7] += gy * dt;
        }
        ecc_block_sum<8>(b, red);
        const float H00 = (float)b[0], H01 = (float)b[1], H11 = (float)b[2];
        const double det = (double)H00 * (double)H11 - (double)H01 * (double)H01;
        float Hi00 = 0.f, Hi01 = 0.f, Hi11 = 0.f;   // Mat::inv of a singular matrix is the zero matrix
        if (det != 0.0) {
            const double dinv = 1.0 / det;
            Hi00 = (float)((double)H11 * dinv);
            Hi11 = (float)((double)H00 * dinv);
            Hi01 = (float)(-(double)H01 * dinv);
        }
        const double corr = b[3];
        last_rho = rho;
        rho = corr / (i_norm * t_norm);
        if (rho != rho) return 1;   // "NaN encountered."
        const float ip0 = (float)b[4], ip1 = (float)b[5], tp0 = (float)b[6], tp1 = (float)b[7];
        const float iph0 = Hi00 * ip0 + Hi01 * ip1, iph1 = Hi01 * ip0 + Hi11 * ip1;
        const double lam_n = i_norm * i_norm - ((double)ip0 * (double)iph0 + (double)ip1 * (double)iph1);
        const double lam_d = corr - ((double)tp0 * (double)iph0 + (double)tp1 * (double)iph1);
        if (lam_d <= 0.0) return 1;   // "The algorithm stopped before its convergence."
        const float lam = (float)(lam_n / lam_d);
        // pass C: projection of the error image lambda * Tz - Iz
        double c[2] = {0.0, 0.0};
        for (int p = BMB_TID; p < npx; p += BMB_NT) {
            const int y = p / w, x = p - y * w;
            const int my = y + t.my, mx = x + t.mx;
            const bool m = !(my < 0 || my >= h || mx < 0 || mx >= w);
            float iw, gxw, gyw;
            ecc_sample(I, h, w, t, y, x, iw, gxw, gyw);
            const float iz = m ? iw - i_mean_f : iw;
            const float tz = m ? (float)T[p] - t_mean_f : 0.f;
            const float e = lam * tz - iz;
            c[0] += (double)gxw * (double)e; c[1] += (double)gyw * (double)e;
        }
        ecc_block_sum<2>(c, red);
        const float ep0 = (float)c[0], ep1 = (float)c[1];
        tx = tx + (Hi00 * ep0 + Hi01 * ep1);
        ty = ty + (Hi01 * ep0 + Hi11 * ep1);
    }
    txy[0] = tx;
    txy[1] = ty;
    return 0;
}

}  // namespace bmb
