// reid_model.cu -- ReID appearance-embedding path on the GPU: detection crops -> OSNet -> L2-normalised rows.
//
// Replaces (relative to /root/reference/boxmot):
//   reid/backends/base_backend.py:148-207   get_crops (cv2.resize INTER_LINEAR, BGR2RGB, /255, mean/std) and
//                                           get_features (forward + row-wise L2 normalisation)
//   reid/backbones/osnet.py:27-260,380-405  ConvLayer / Conv1x1 / Conv1x1Linear / LightConv3x3 / ChannelGate /
//                                           OSBlock / OSNet.forward (eval mode, BatchNorm folded offline)
//   native/cpp/trackers/base/src/reid_onnx.cpp:51-383  (per-crop batch-1 ORT forward of the native path)
//
// Data layout: activations are NHWC float32 in HBM, one chunk of crops at a time (256 by default: measured on B200,
// one full-width chunk beats two L2-resident half chunks -- 415 vs 351 frames/s at 208 crops -- because the small
// tile kernels are latency-bound and want full waves); weights are a BN-folded float32 blob
// (boxmot_b200/weights.py) uploaded once.  Round-1 kernels are float32 CUDA-core kernels with shared-memory
// tiling; every 1x1 convolution goes through one GEMM-shaped kernel (k_pointwise) whose prologue can build
// the gated branch sum on the fly and whose epilogue fuses bias / residual / ReLU.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "engine.h"
#include "pointwise_tc.cuh"

namespace bmb {

#define RCUDA_OK(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
    } while (0)

constexpr int IN_H = 256, IN_W = 128;
constexpr uint32_t BLOB_MAGIC = 0x45523242u;

__device__ __forceinline__ int chunk_count(const int* d_n, int off, int cap) {
    int n = *d_n - off;
    n = n < 0 ? 0 : n;
    return n > cap ? cap : n;
}

// ---------------------------------------------------------------------------------------------------
// K1: crop + OpenCV-exact bilinear resize + BGR->RGB + /255 + mean/std  ->  (N,256,128,3) float32
// One CTA per crop.  cv2.resize(INTER_LINEAR) on uint8 is integer arithmetic: 11-bit coefficients derived from a
// float32 phase, horizontal pass in int32, vertical (((b0*(S0>>4))>>16)+((b1*(S1>>4))>>16)+2)>>2.  x phases are
// clamped at the borders, y rows are clipped at fetch (pinned against cv2 in tests/test_oracle_reid.py).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void linear_coeff(int d, int src_n, double scale, bool clamp, int& idx, int& a0, int& a1) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f = f - (float)s;
    if (clamp) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src_n - 1) { s = src_n - 1; f = 0.f; }
    }
    idx = s;
    a0 = (int)rintf((1.0f - f) * 2048.0f);
    a1 = (int)rintf(f * 2048.0f);
}

__global__ void __launch_bounds__(256) k_crop_resize_norm(const uint8_t* __restrict__ images, size_t image_stride,
                                                          int rows, int cols, const CropDesc* __restrict__ crops,
                                                          const int* __restrict__ d_n, int off, int cap,
                                                          float* __restrict__ blob, int pad_mode) {
    const int n = blockIdx.x;
    if (n >= chunk_count(d_n, off, cap)) return;
    const CropDesc cd = crops[off + n];
    __shared__ int xi[IN_W], xa0[IN_W], xa1[IN_W];
    __shared__ int yi[IN_H], ya0[IN_H], ya1[IN_H];
    // box.round().astype(int): round half to even
    const int x1 = (int)rintf(cd.x1), y1 = (int)rintf(cd.y1), x2 = (int)rintf(cd.x2), y2 = (int)rintf(cd.y2);
    const int cx1 = max(0, x1), cy1 = max(0, y1), cx2 = min(cols, x2), cy2 = min(rows, y2);
    const bool valid = cx2 > cx1 && cy2 > cy1;
    const int sw = cx2 - cx1, sh = cy2 - cy1;
    // resize_pad (preprocessing.py:21-45): scale = min(W / w, H / h); new = int(size * scale); centred, ImageNet-mean border
    int nw = IN_W, nh = IN_H, pl = 0, pt = 0;
    if (valid && pad_mode) {
        const double sc = fmin((double)IN_W / (double)sw, (double)IN_H / (double)sh);
        nw = max(1, (int)((double)sw * sc));
        nh = max(1, (int)((double)sh * sc));
        pl = (IN_W - nw) / 2;
        pt = (IN_H - nh) / 2;
    }
    if (valid) {
        const double sx = 1.0 / ((double)nw / (double)sw), sy = 1.0 / ((double)nh / (double)sh);
        for (int d = threadIdx.x; d < nw; d += blockDim.x) linear_coeff(d, sw, sx, true, xi[d], xa0[d], xa1[d]);
        for (int d = threadIdx.x; d < nh; d += blockDim.x) linear_coeff(d, sh, sy, false, yi[d], ya0[d], ya1[d]);
    }
    __syncthreads();
    const uint8_t* img = images + (size_t)cd.image * image_stride;
    float* out = blob + (size_t)n * IN_H * IN_W * 3;
    const float mean[3] = {0.485f, 0.456f, 0.406f};
    const float stdv[3] = {0.229f, 0.224f, 0.225f};
    for (int p = threadIdx.x; p < IN_H * IN_W; p += blockDim.x) {
        const int py = p / IN_W, px = p - py * IN_W;
        const int dy = py - pt, dx = px - pl;
        int v[3] = {0, 0, 0};
        if (valid && (dx < 0 || dx >= nw || dy < 0 || dy >= nh)) {
            v[0] = 104; v[1] = 116; v[2] = 124;      // IMAGENET_MEAN_BGR (the crop is BGR until the channel flip below)
        } else if (valid) {
            const int sx0 = xi[dx], sx1 = min(sx0 + 1, sw - 1);
            const int r0 = min(max(yi[dy], 0), sh - 1), r1 = min(max(yi[dy] + 1, 0), sh - 1);
            const uint8_t* p0 = img + ((size_t)(cy1 + r0) * cols + cx1) * 3;
            const uint8_t* p1 = img + ((size_t)(cy1 + r1) * cols + cx1) * 3;
            const int a0 = xa0[dx], a1 = xa1[dx], b0 = ya0[dy], b1 = ya1[dy];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int h0 = (int)p0[sx0 * 3 + c] * a0 + (int)p0[sx1 * 3 + c] * a1;
                const int h1 = (int)p1[sx0 * 3 + c] * a0 + (int)p1[sx1 * 3 + c] * a1;
                int r = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                v[c] = min(max(r, 0), 255);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {  // output channel c is RGB: source channel 2-c
            const float f = __fdiv_rn((float)v[2 - c], 255.0f);
            out[(size_t)p * 3 + c] = __fdiv_rn(f - mean[c], stdv[c]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K2: stem 7x7 stride-2 conv (3 -> C0) + folded BN + ReLU : (N,256,128,3) -> (N,128,64,C0)
// CTA = 8 output rows x 64 columns of one crop; the 21 x 134 x 3 input window and the weight chunk live in
// shared memory; a thread owns two output pixels (x, x+32) x 16 output channels.
// ---------------------------------------------------------------------------------------------------
constexpr int ST_R = 8, ST_IR = 2 * ST_R + 5, ST_IC = IN_W + 6;
__global__ void __launch_bounds__(256) k_stem(const float* __restrict__ blob, const float* __restrict__ w,
                                              const float* __restrict__ bias, int C0, const int* __restrict__ d_n,
                                              int off, int cap, float* __restrict__ out) {
    const int n = blockIdx.y;
    if (n >= chunk_count(d_n, off, cap)) return;
    extern __shared__ __align__(16) float smem[];
    float* sin = smem;                         // [ST_IR][ST_IC][3]
    float* sw = smem + ((ST_IR * ST_IC * 3 + 3) & ~3);  // [147][16], 16-byte aligned for float4 reads
    const int oy0 = blockIdx.x * ST_R;
    const int iy0 = oy0 * 2 - 3;
    const float* src = blob + (size_t)n * IN_H * IN_W * 3;
    for (int e = threadIdx.x; e < ST_IR * ST_IC * 3; e += blockDim.x) {
        const int r = e / (ST_IC * 3), rem = e - r * (ST_IC * 3);
        const int cidx = rem / 3, ch = rem - cidx * 3;
        const int iy = iy0 + r, ix = cidx - 3;
        sin[e] = (iy >= 0 && iy < IN_H && ix >= 0 && ix < IN_W) ? src[((size_t)iy * IN_W + ix) * 3 + ch] : 0.f;
    }
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    for (int co0 = 0; co0 < C0; co0 += 16) {
        __syncthreads();
        for (int e = threadIdx.x; e < 147 * 16; e += blockDim.x) {
            const int k = e >> 4, c = e & 15;
            sw[e] = w[(size_t)k * C0 + co0 + c];
        }
        __syncthreads();
        // packed FP32x2 FMAs (FFMA2, sm_100): two output channels per instruction, IEEE per lane (= fmaf bit for bit)
        float2 acc0[8], acc1[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { acc0[c] = make_float2(0.f, 0.f); acc1[c] = make_float2(0.f, 0.f); }
        for (int kh = 0; kh < 7; ++kh) {
            const float* row = sin + (size_t)(ty * 2 + kh) * ST_IC * 3;
            for (int kw = 0; kw < 7; ++kw) {
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const float a0 = row[(tx * 2 + kw) * 3 + ci];
                    const float a1 = row[((tx + 32) * 2 + kw) * 3 + ci];
                    const float2 a0p = make_float2(a0, a0), a1p = make_float2(a1, a1);
                    const float4* wp = reinterpret_cast<const float4*>(sw + ((kh * 7 + kw) * 3 + ci) * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 wv = wp[q];
                        const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
                        acc0[q * 2 + 0] = __ffma2_rn(a0p, w01, acc0[q * 2 + 0]);
                        acc0[q * 2 + 1] = __ffma2_rn(a0p, w23, acc0[q * 2 + 1]);
                        acc1[q * 2 + 0] = __ffma2_rn(a1p, w01, acc1[q * 2 + 0]);
                        acc1[q * 2 + 1] = __ffma2_rn(a1p, w23, acc1[q * 2 + 1]);
                    }
                }
            }
        }
        float* o0 = out + (((size_t)n * 128 + oy0 + ty) * 64 + tx) * C0 + co0;
        float* o1 = o0 + (size_t)32 * C0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 b = *reinterpret_cast<const float4*>(bias + co0 + q * 4);
            float4 r0 = make_float4(fmaxf(acc0[q * 2].x + b.x, 0.f), fmaxf(acc0[q * 2].y + b.y, 0.f),
                                    fmaxf(acc0[q * 2 + 1].x + b.z, 0.f), fmaxf(acc0[q * 2 + 1].y + b.w, 0.f));
            float4 r1 = make_float4(fmaxf(acc1[q * 2].x + b.x, 0.f), fmaxf(acc1[q * 2].y + b.y, 0.f),
                                    fmaxf(acc1[q * 2 + 1].x + b.z, 0.f), fmaxf(acc1[q * 2 + 1].y + b.w, 0.f));
            reinterpret_cast<float4*>(o0)[q] = r0;
            reinterpret_cast<float4*>(o1)[q] = r1;
        }
    }
}

// K3: max pool 3x3 stride 2 pad 1 : (N,H,W,C) -> (N,H/2,W/2,C)
__global__ void k_maxpool3s2(const float* __restrict__ in, int H, int W, int C, const int* __restrict__ d_n, int off,
                             int cap, float* __restrict__ out) {
    const int n_crops = chunk_count(d_n, off, cap);
    const int OH = H / 2, OW = W / 2, C4 = C / 4;
    const size_t total = (size_t)n_crops * OH * OW * C4;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(e % C4);
        size_t r = e / C4;
        const int ox = (int)(r % OW); r /= OW;
        const int oy = (int)(r % OH);
        const int n = (int)(r / OH);
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy * 2 - 1 + ky;
            if (iy < 0 || iy >= H) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = ox * 2 - 1 + kx;
                if (ix < 0 || ix >= W) continue;
                const float4 v = *reinterpret_cast<const float4*>(in + (((size_t)n * H + iy) * W + ix) * C + c4 * 4);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        *reinterpret_cast<float4*>(out + (((size_t)n * OH + oy) * OW + ox) * C + c4 * 4) = m;
    }
}

// K4: average pool 2x2 stride 2
__global__ void k_avgpool2(const float* __restrict__ in, int H, int W, int C, const int* __restrict__ d_n, int off,
                           int cap, float* __restrict__ out) {
    const int n_crops = chunk_count(d_n, off, cap);
    const int OH = H / 2, OW = W / 2, C4 = C / 4;
    const size_t total = (size_t)n_crops * OH * OW * C4;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(e % C4);
        size_t r = e / C4;
        const int ox = (int)(r % OW); r /= OW;
        const int oy = (int)(r % OH);
        const int n = (int)(r / OH);
        const float* p = in + (((size_t)n * H + oy * 2) * W + ox * 2) * C + c4 * 4;
        const float4 a = *reinterpret_cast<const float4*>(p);
        const float4 b = *reinterpret_cast<const float4*>(p + C);
        const float4 c = *reinterpret_cast<const float4*>(p + (size_t)W * C);
        const float4 d = *reinterpret_cast<const float4*>(p + (size_t)W * C + C);
        float4 o;
        o.x = (a.x + b.x + c.x + d.x) * 0.25f; o.y = (a.y + b.y + c.y + d.y) * 0.25f;
        o.z = (a.z + b.z + c.z + d.z) * 0.25f; o.w = (a.w + b.w + c.w + d.w) * 0.25f;
        *reinterpret_cast<float4*>(out + (((size_t)n * OH + oy) * OW + ox) * C + c4 * 4) = o;
    }
}

// ---------------------------------------------------------------------------------------------------
// K5: pointwise (1x1) convolution as a GEMM:  out[M][N] = act( A[M][K] * W[K][N] + bias (+ residual) )
//   prologue PLAIN : A = in[M][K]
//   prologue GATED : A[m][k] = sum_b gates[crop(m)][b][k] * branch_b[m][k]      for k <  mid
//                            = x[m][k - mid]                                      for k >= mid  (downsample rows)
// CTA: 256 threads, BN output channels (16/32/64), BM = 2048/BN*4 rows; thread = 8 rows x 4 channels;
// K is streamed through shared memory in chunks of 16.
// ---------------------------------------------------------------------------------------------------
struct PwArgs {
    const float* in;          // PLAIN: [M][K];  GATED: x [M][K - mid] (may be null when K == mid)
    const float* branch[4];   // GATED: four [M][mid] tensors
    const float* gates;       // GATED: [crops][4][mid]
    const float* w;           // [K][N]
    const float* bias;        // [N]
    const float* residual;    // [M][N] or null
    float* out;               // [M][N]
    int K, N, mid, HW, relu;
    const float* w_tc;        // canonical hi/lo weight blocks for the tcgen05 path (null: CUDA-core kernel)
    int Kpad, Npad;
};

template <int BN, bool GATED>
__global__ void __launch_bounds__(256) k_pointwise(const PwArgs a, const int* __restrict__ d_n, int off, int cap) {
    constexpr int NTN = BN / 4, NTM = 256 / NTN, BM = NTM * 8, BK = 16, LDA = BM + 4;
    const int M = chunk_count(d_n, off, cap) * a.HW;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (m0 >= M) return;
    __shared__ __align__(16) float As[BK * LDA];
    __shared__ __align__(16) float Bs[BK * BN];
    const int tn = threadIdx.x % NTN, tm = threadIdx.x / NTN;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int K = a.K, N = a.N;
    const int KX = GATED ? K - a.mid : 0;
    for (int k0 = 0; k0 < K; k0 += BK) {
        // A chunk: BM rows x 16 k, loaded as float4 along k, stored k-major
        for (int e = threadIdx.x; e < BM * 4; e += 256) {
            const int r = e >> 2, kq = (e & 3) * 4;
            const int m = m0 + r, k = k0 + kq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < M && k < K) {
                if (!GATED) {
                    v = *reinterpret_cast<const float4*>(a.in + (size_t)m * K + k);
                } else if (k < a.mid) {
                    const float* g = a.gates + (size_t)(m / a.HW) * 4 * a.mid + k;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const float4 x = *reinterpret_cast<const float4*>(a.branch[b] + (size_t)m * a.mid + k);
                        const float4 gg = *reinterpret_cast<const float4*>(g + b * a.mid);
                        v.x = fmaf(x.x, gg.x, v.x); v.y = fmaf(x.y, gg.y, v.y);
                        v.z = fmaf(x.z, gg.z, v.z); v.w = fmaf(x.w, gg.w, v.w);
                    }
                } else {
                    v = *reinterpret_cast<const float4*>(a.in + (size_t)m * KX + (k - a.mid));
                }
            }
            As[(kq + 0) * LDA + r] = v.x; As[(kq + 1) * LDA + r] = v.y;
            As[(kq + 2) * LDA + r] = v.z; As[(kq + 3) * LDA + r] = v.w;
        }
        for (int e = threadIdx.x; e < BK * BN / 4; e += 256) {
            const int kk = e / (BN / 4), c = (e % (BN / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + kk < K && n0 + c < N) v = *reinterpret_cast<const float4*>(a.w + (size_t)(k0 + kk) * N + n0 + c);
            *reinterpret_cast<float4*>(Bs + kk * BN + c) = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(As + kk * LDA + tm * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(As + kk * LDA + tm * 8 + 4);
            const float4 b = *reinterpret_cast<const float4*>(Bs + kk * BN + tn * 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[i][0] = fmaf(av[i], b.x, acc[i][0]); acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
                acc[i][2] = fmaf(av[i], b.z, acc[i][2]); acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
            }
        }
        __syncthreads();
    }
    const int c = n0 + tn * 4;
    if (c >= N) return;
    const float4 bv = *reinterpret_cast<const float4*>(a.bias + c);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + tm * 8 + i;
        if (m >= M) break;
        float4 v = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
        if (a.residual) {
            const float4 r = *reinterpret_cast<const float4*>(a.residual + (size_t)m * N + c);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (a.relu == 2) { v.x = fminf(v.x, 6.f); v.y = fminf(v.y, 6.f); v.z = fminf(v.z, 6.f); v.w = fminf(v.w, 6.f); }  // ReLU6
        *reinterpret_cast<float4*>(a.out + (size_t)m * N + c) = v;
    }
}

// ---------------------------------------------------------------------------------------------------
// K6: LightConv3x3 = 1x1 (linear) -> depthwise 3x3 -> folded BN -> ReLU, fused per spatial tile.
// grid = (row tiles, branches of this level, crops).  Phase A computes T = pw(in) for the tile plus a one-pixel
// halo into shared memory (two pixels x 16 channels per thread per pass, weights broadcast from shared memory);
// phase B applies the depthwise taps from shared memory, writes the activation and, for the last layer of a
// branch, per-tile channel sums for the ChannelGate's global average pool (summed in fixed order later).
// ---------------------------------------------------------------------------------------------------
struct LightArgs {
    const float* in[4];
    float* out[4];
    const float* wpw[4];
    const float* wdw[4];
    const float* bias[4];
    float* sums[4];   // [crops][tiles][C] or null
    int H, W, C, R;   // R = tile rows
    const float* wtc[4];   // 1x1 weights as canonical K-major hi / lo blocks (tc::pack_weights) for the tcgen05 stage
};

__global__ void k_lightconv(const LightArgs a, const int* __restrict__ d_n, int off, int cap) {
    const int n = blockIdx.z;
    if (n >= chunk_count(d_n, off, cap)) return;
    const int br = blockIdx.y, tile = blockIdx.x;
    const int H = a.H, W = a.W, C = a.C, R = a.R;
    const int TW = W + 2, TR = R + 2;
    extern __shared__ __align__(16) float smem[];
    float* sT = smem;                       // [TR][TW][C]
    float* sW = sT + (size_t)TR * TW * C;   // [C][C]
    float* sD = sW + (size_t)C * C;         // [9][C]
    float* sP = sD + 9 * C;                 // partial sums [groups][C]
    const int y0 = tile * R;
    const float* in = a.in[br] + (size_t)n * H * W * C;
    for (int e = threadIdx.x; e < C * C; e += blockDim.x) sW[e] = a.wpw[br][e];
    for (int e = threadIdx.x; e < 9 * C; e += blockDim.x) sD[e] = a.wdw[br][e];
    __syncthreads();
    // ---- phase A: T = in * Wpw on the haloed tile ----
    const int n_px = TR * TW;
    const int n_pairs = (n_px + 1) / 2;
    const int n_cchunks = C / 8;  // 8 output channels per pass keeps 16 accumulators for two pixels
    for (int item = threadIdx.x; item < n_pairs * n_cchunks; item += blockDim.x) {
        const int pair = item / n_cchunks, cc = (item - pair * n_cchunks) * 8;
        const int p0 = pair * 2, p1 = p0 + 1;
        const int ty0 = p0 / TW, tx0 = p0 - ty0 * TW;
        const int ty1 = p1 / TW, tx1 = p1 - ty1 * TW;
        const int gy0 = y0 + ty0 - 1, gx0 = tx0 - 1, gy1 = y0 + ty1 - 1, gx1 = tx1 - 1;
        const bool v0 = gy0 >= 0 && gy0 < H && gx0 >= 0 && gx0 < W;
        const bool v1 = p1 < n_px && gy1 >= 0 && gy1 < H && gx1 >= 0 && gx1 < W;
        float acc0[8], acc1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
        if (v0 || v1) {
            const float* q0 = in + ((size_t)gy0 * W + gx0) * C;
            const float* q1 = in + ((size_t)gy1 * W + gx1) * C;
            for (int k = 0; k < C; k += 4) {
                const float4 x0 = v0 ? *reinterpret_cast<const float4*>(q0 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 x1 = v1 ? *reinterpret_cast<const float4*>(q1 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float xa[4] = {x0.x, x0.y, x0.z, x0.w};
                const float xb[4] = {x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const float4 w0 = *reinterpret_cast<const float4*>(sW + (size_t)(k + kk) * C + cc);
                    const float4 w1 = *reinterpret_cast<const float4*>(sW + (size_t)(k + kk) * C + cc + 4);
                    acc0[0] = fmaf(xa[kk], w0.x, acc0[0]); acc0[1] = fmaf(xa[kk], w0.y, acc0[1]);
                    acc0[2] = fmaf(xa[kk], w0.z, acc0[2]); acc0[3] = fmaf(xa[kk], w0.w, acc0[3]);
                    acc0[4] = fmaf(xa[kk], w1.x, acc0[4]); acc0[5] = fmaf(xa[kk], w1.y, acc0[5]);
                    acc0[6] = fmaf(xa[kk], w1.z, acc0[6]); acc0[7] = fmaf(xa[kk], w1.w, acc0[7]);
                    acc1[0] = fmaf(xb[kk], w0.x, acc1[0]); acc1[1] = fmaf(xb[kk], w0.y, acc1[1]);
                    acc1[2] = fmaf(xb[kk], w0.z, acc1[2]); acc1[3] = fmaf(xb[kk], w0.w, acc1[3]);
                    acc1[4] = fmaf(xb[kk], w1.x, acc1[4]); acc1[5] = fmaf(xb[kk], w1.y, acc1[5]);
                    acc1[6] = fmaf(xb[kk], w1.z, acc1[6]); acc1[7] = fmaf(xb[kk], w1.w, acc1[7]);
                }
            }
        }
        float4* t0 = reinterpret_cast<float4*>(sT + (size_t)p0 * C + cc);
        t0[0] = make_float4(acc0[0], acc0[1], acc0[2], acc0[3]);
        t0[1] = make_float4(acc0[4], acc0[5], acc0[6], acc0[7]);
        if (p1 < n_px) {
            float4* t1 = reinterpret_cast<float4*>(sT + (size_t)p1 * C + cc);
            t1[0] = make_float4(acc1[0], acc1[1], acc1[2], acc1[3]);
            t1[1] = make_float4(acc1[4], acc1[5], acc1[6], acc1[7]);
        }
    }
    __syncthreads();
    // ---- phase B: depthwise 3x3 + bias + ReLU ----
    const int C4 = C / 4;
    const int rows_here = min(R, H - y0);
    const int c4 = threadIdx.x % C4, grp = threadIdx.x / C4, n_grp = blockDim.x / C4;
    float4 wv[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wv[t] = *reinterpret_cast<const float4*>(sD + t * C + c4 * 4);
    const float4 bv = *reinterpret_cast<const float4*>(a.bias[br] + c4 * 4);
    float4 psum = make_float4(0.f, 0.f, 0.f, 0.f);
    float* outp = a.out[br] + (size_t)n * H * W * C;
    for (int p = grp; p < rows_here * W; p += n_grp) {
        const int y = p / W, x = p - y * W;
        float4 acc = bv;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4 t = *reinterpret_cast<const float4*>(sT + ((size_t)(y + ky) * TW + x + kx) * C + c4 * 4);
                const float4 w = wv[ky * 3 + kx];
                acc.x = fmaf(t.x, w.x, acc.x); acc.y = fmaf(t.y, w.y, acc.y);
                acc.z = fmaf(t.z, w.z, acc.z); acc.w = fmaf(t.w, w.w, acc.w);
            }
        acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
        *reinterpret_cast<float4*>(outp + ((size_t)(y0 + y) * W + x) * C + c4 * 4) = acc;
        psum.x += acc.x; psum.y += acc.y; psum.z += acc.z; psum.w += acc.w;
    }
    if (a.sums[br]) {
        *reinterpret_cast<float4*>(sP + (size_t)grp * C + c4 * 4) = psum;
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float s = 0.f;
            for (int g = 0; g < n_grp; ++g) s += sP[(size_t)g * C + c];
            a.sums[br][((size_t)n * gridDim.x + tile) * C + c] = s;
        }
    }
}

// K6 v2.  Same contract as k_lightconv, restructured around the two limits ncu showed for v1 (L1/TEX 79 %):
//   * phase A read each pixel's K-vector with per-thread float4 loads 64 B apart: every warp request touched 16
//     cache lines.  v2 stages the haloed input tile once with coalesced cp.async into a K-chunk-planar shared
//     layout sX[C/4][pixels] (zero fill outside the image), so phase A reads consecutive float4s (conflict-free),
//     one warp = 64 consecutive pixels x 8 output channels with the 1x1 weights broadcast;
//   * phase B loaded 9 float4 per output from shared memory.  v2 walks columns: a thread owns (x, 4 channels),
//     slides down its rows keeping a 3x3 window in registers and loads 3 float4 per output.
// Shared memory: sX + sT (2 x tile) + weights; pick_tile_rows2 keeps two CTAs per SM for the narrow models.
static inline size_t light2_smem_bytes(int n_px, int C, int threads, int ppl = 4, bool tc_stage = false) {
    const int n_pxp = ((n_px + 32 * ppl - 1) / (32 * ppl)) * (32 * ppl) + 2;
    return sizeof(float) * (2 * (size_t)n_pxp * C + (size_t)C * C * (tc_stage ? 2 : 1) + 9 * C +
                            (size_t)(threads / (C / 4)) * C);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
}

template <int C, int W, int R, int PPL = 4, int MINB = 2, bool TC = false, int NTHREADS = 256>
__global__ void __launch_bounds__(NTHREADS, MINB) k_lightconv2(const LightArgs a, const int* __restrict__ d_n, int off, int cap) {
    const int n = blockIdx.z;
    if (n >= chunk_count(d_n, off, cap)) return;
    const int br = blockIdx.y, tile = blockIdx.x;
    const int H = a.H;
    constexpr int TW = W + 2, TR = R + 2, C4 = C / 4;
    constexpr int n_px = TR * TW;
    constexpr int n_pxp = ((n_px + 32 * PPL - 1) / (32 * PPL)) * (32 * PPL) + 2;   // planes 8 banks apart; whole pixel groups in bounds
    constexpr int NT = NTHREADS;
    extern __shared__ __align__(16) float smem[];
    float4* sX = reinterpret_cast<float4*>(smem);            // [C4][n_pxp]
    float4* sT = sX + (size_t)C4 * n_pxp;                    // [C4][n_pxp]  (planar like sX: conflict-free both ways)
    float* sW = reinterpret_cast<float*>(sT + (size_t)C4 * n_pxp);   // [C][C]
    float* sD = sW + (size_t)C * C;                          // [9][C]
    float* sP = sD + 9 * C;                                  // [NT / C4][C]
    const int y0 = tile * R;
    const float* in = a.in[br] + (size_t)n * H * W * C;
    // ---- stage: weights + haloed input tile (compile-time shapes: the index arithmetic folds to mul/shift) ----
    for (int e = threadIdx.x; e < n_px * C4; e += NT) {
        const int p = e / C4, ch = e - p * C4;
        const int ty = p / TW, tx = p - ty * TW;
        const int gy = y0 + ty - 1, gx = tx - 1;
        float4* dst = sX + ch * n_pxp + p;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) cp_async16(dst, in + ((size_t)gy * W + gx) * C + ch * 4);
        else *dst = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int e = threadIdx.x; e < (n_pxp - n_px) * C4; e += NT) {
        const int q = e / C4, ch = e - q * C4;
        sX[ch * n_pxp + n_px + q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int e = threadIdx.x; e < C * C / 4; e += NT)
        reinterpret_cast<float4*>(sW)[e] = reinterpret_cast<const float4*>(a.wpw[br])[e];
    for (int e = threadIdx.x; e < 9 * C; e += NT) sD[e] = a.wdw[br][e];
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;");
    if constexpr (TC) {
        // ---- phase A on the tensor cores: T = X * Wpw as tcgen05.mma kind::tf32 with the 3-term hi/lo split
        // (float32-class accuracy).  The staged tile IS the canonical no-swizzle K-major operand: a core matrix is 8
        // consecutive pixels x 16 bytes of one K-chunk plane (contiguous 128 B), the next K chunk is one plane further
        // (LBO = n_pxp * 16 B), the next 8 pixels 128 B further (SBO).  X is split in place (hi) and into the T buffer
        // (lo); the accumulators of all 128-pixel tiles live in TMEM at once, one commit, then T overwrites lo.
        static_assert(!TC || (C % 16 == 0 && C <= 64), "tcgen05 stage: N must be a multiple of 16");
        __shared__ __align__(8) uint64_t bar_mma;
        __shared__ uint32_t tmem_base_s;
        constexpr int n_mt = (n_px + 127) / 128;
        constexpr uint32_t tmem_cols = n_mt * C <= 32 ? 32u : (n_mt * C <= 64 ? 64u : (n_mt * C <= 128 ? 128u : (n_mt * C <= 256 ? 256u : 512u)));
        static_assert(!TC || n_mt * 128 <= n_pxp, "tile rows must stay inside the padded planes");
        static_assert(!TC || n_mt * C <= 512, "accumulators must fit TMEM");
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_base_s)), "r"(tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        if (threadIdx.x == 0) {
            tc::mbar_init(&bar_mma, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        // packed weights (hi block then lo block, canonical [n][k]) over the plain copy in sW: 2 * C * C floats fit
        // because sW + sD + sP follow each other (C*C + 9C + 64C >= 2*C*C for C <= 64)
        __syncthreads();   // cp.async data + the plain sW copy are complete: sW may be overwritten
        float* sWtc = sW;
        for (int e = threadIdx.x; e < 2 * C * C / 4; e += NT)
            reinterpret_cast<float4*>(sWtc)[e] = reinterpret_cast<const float4*>(a.wtc[br])[e];
        // sD was staged before and sits behind sW: re-stage it after the packed weights
        float* sD2 = sWtc + 2 * C * C;
        for (int e = threadIdx.x; e < 9 * C; e += NT) sD2[e] = a.wdw[br][e];
        // split X -> hi (in place) / lo (T buffer), pads included
        for (int e = threadIdx.x; e < C4 * n_pxp; e += NT) {
            const float4 v = sX[e];
            float4 hi, lo;
            hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo.x = v.x - hi.x;
            hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo.y = v.y - hi.y;
            hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo.z = v.z - hi.z;
            hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo.w = v.w - hi.w;
            sX[e] = hi;
            sT[e] = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_base = tmem_base_s;
        if (threadIdx.x == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t a_hi = tc::smem_u32(sX), a_lo = tc::smem_u32(sT);
            const uint32_t b_hi = tc::smem_u32(sWtc), b_lo = tc::smem_u32(sWtc + C * C);
            const uint32_t lbo_a = (uint32_t)n_pxp * 16u, sbo_a = 128u, lbo_b = 128u, sbo_b = (uint32_t)C * 32u;
#pragma unroll 1
            for (int t = 0; t < n_mt; ++t) {
#pragma unroll
                for (int ks = 0; ks < C; ks += 8) {
                    const uint32_t ao = (uint32_t)t * 2048u + (uint32_t)(ks >> 2) * lbo_a, bo = (uint32_t)(ks >> 2) * 128u;
                    const uint64_t dah = tc::make_desc(a_hi + ao, lbo_a, sbo_a), dal = tc::make_desc(a_lo + ao, lbo_a, sbo_a);
                    const uint64_t dbh = tc::make_desc(b_hi + bo, lbo_b, sbo_b), dbl = tc::make_desc(b_lo + bo, lbo_b, sbo_b);
                    const uint32_t d = tmem_base + (uint32_t)(t * C);
                    tc::mma_tf32(d, dah, dbh, idesc, ks > 0 ? 1u : 0u);
                    tc::mma_tf32(d, dah, dbl, idesc, 1u);
                    tc::mma_tf32(d, dal, dbh, idesc, 1u);
                }
            }
            tc::mma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // TMEM -> T (planar): warp w reads lanes 32*(w%4).. of the tiles t = w/4, w/4 + 2, ...
        for (int t = warp >> 2; t < n_mt; t += 2) {
            const int p = t * 128 + (warp & 3) * 32 + lane;
#pragma unroll
            for (int c0 = 0; c0 < C; c0 += 8) {
                float v[8];
                tc::tmem_ld8(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t * C + c0), v);
                if (p < n_px) {
                    sT[(c0 / 4) * n_pxp + p] = make_float4(v[0], v[1], v[2], v[3]);
                    sT[(c0 / 4 + 1) * n_pxp + p] = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
        sD = sD2;
        sP = sD2 + 9 * C;
    } else {
    __syncthreads();
    // ---- phase A: T = X * Wpw, warp item = (128-pixel group, 8 output channels), 4 pixels per lane ----
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        constexpr int G = 32 * PPL;
        constexpr int n_pg = (n_px + G - 1) / G, n_cc = C >> 3;
        for (int item = warp; item < n_pg * n_cc; item += NT / 32) {
            const int pg = item % n_pg, cc = (item / n_pg) * 8;
            const int p0 = pg * G + lane;
            float acc[PPL][8];
#pragma unroll
            for (int q = 0; q < PPL; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
#pragma unroll 2
            for (int kc = 0; kc < C4; ++kc) {
                float xv[PPL][4];
#pragma unroll
                for (int q = 0; q < PPL; ++q) {
                    const float4 x = sX[kc * n_pxp + p0 + 32 * q];
                    xv[q][0] = x.x; xv[q][1] = x.y; xv[q][2] = x.z; xv[q][3] = x.w;
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const float4 w0 = *reinterpret_cast<const float4*>(sW + (kc * 4 + kk) * C + cc);
                    const float4 w1 = *reinterpret_cast<const float4*>(sW + (kc * 4 + kk) * C + cc + 4);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int q = 0; q < PPL; ++q)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[q][j] = fmaf(xv[q][kk], wv[j], acc[q][j]);
                }
            }
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                const int p = p0 + 32 * q;
                if (p < n_px) {
                    sT[(cc / 4) * n_pxp + p] = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
                    sT[(cc / 4 + 1) * n_pxp + p] = make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
                }
            }
        }
    }
    __syncthreads();
    }
    // ---- phase B: depthwise 3x3 + bias + ReLU, column walkers with a register window ----
    const int rows_here = min(R, H - y0);
    constexpr int walkers = W * C4;
    constexpr int n_grp = NT / C4;
    constexpr int act = n_grp * C4;    // threads that walk: a multiple of C4 so a thread keeps its channel group
    constexpr int n_split = (act / walkers) < 1 ? 1 : ((act / walkers) > R ? R : (act / walkers));
    constexpr int rows_per = (R + n_split - 1) / n_split;
    float* outp = a.out[br] + (size_t)n * H * W * C;
    float4 psum = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x < act) {
        for (int wk = threadIdx.x; wk < walkers * n_split; wk += act) {
            const int c4 = wk % C4, x = (wk / C4) % W, sp = wk / walkers;
            const int ya = sp * rows_per, yb = min(ya + rows_per, rows_here);
            float4 wv[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) wv[t] = *reinterpret_cast<const float4*>(sD + t * C + c4 * 4);
            const float4 bv = *reinterpret_cast<const float4*>(a.bias[br] + c4 * 4);
            const float4* tbase = sT + c4 * n_pxp + x;
            float4 win[3][3];   // rows (y, y+1, y+2) mod 3 live in fixed registers: the row loop is unrolled by 3
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                win[0][kx] = tbase[ya * TW + kx];
                win[1][kx] = tbase[(ya + 1) * TW + kx];
            }
            float* orow = outp + ((size_t)(y0 + ya) * W + x) * C + c4 * 4;
            for (int y = ya; y < yb; y += 3) {
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    if (y + u < yb) {
                        const int i0 = u % 3, i1 = (u + 1) % 3, i2 = (u + 2) % 3;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
                            win[i2][kx] = tbase[(y + u + 2) * TW + kx];
                        float4 acc = bv;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const float4 w0 = wv[kx], w1 = wv[3 + kx], w2 = wv[6 + kx];
                            const float4 t0 = win[i0][kx], t1 = win[i1][kx], t2 = win[i2][kx];
                            acc.x = fmaf(t0.x, w0.x, acc.x); acc.y = fmaf(t0.y, w0.y, acc.y);
                            acc.z = fmaf(t0.z, w0.z, acc.z); acc.w = fmaf(t0.w, w0.w, acc.w);
                            acc.x = fmaf(t1.x, w1.x, acc.x); acc.y = fmaf(t1.y, w1.y, acc.y);
                            acc.z = fmaf(t1.z, w1.z, acc.z); acc.w = fmaf(t1.w, w1.w, acc.w);
                            acc.x = fmaf(t2.x, w2.x, acc.x); acc.y = fmaf(t2.y, w2.y, acc.y);
                            acc.z = fmaf(t2.z, w2.z, acc.z); acc.w = fmaf(t2.w, w2.w, acc.w);
                        }
                        acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f);
                        acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
                        *reinterpret_cast<float4*>(orow + (size_t)u * W * C) = acc;
                        psum.x += acc.x; psum.y += acc.y; psum.z += acc.z; psum.w += acc.w;
                    }
                }
                orow += (size_t)3 * W * C;
            }
        }
    }
    if (a.sums[br]) {
        // every walking thread owns the fixed slot (threadIdx / C4, threadIdx % C4)
        if (threadIdx.x < act)
            *reinterpret_cast<float4*>(sP + (threadIdx.x / C4) * C + (threadIdx.x % C4) * 4) = psum;
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += NT) {
            float s = 0.f;
            for (int g = 0; g < n_grp; ++g) s += sP[g * C + c];
            a.sums[br][((size_t)n * gridDim.x + tile) * C + c] = s;
        }
    }
}

// K6 v3: a whole OSBlock branch (1-4 LightConv3x3 layers) per CTA.  The per-level kernel spends a third of its time
// waiting for its input tile and writes every intermediate level back to HBM; here a CTA stages the conv1 output once
// (tile rows + `depth` halo rows each side), runs  1x1 -> depthwise 3x3 + bias + ReLU  `depth` times between two
// shared-memory buffers (X -> T -> X ...), and only the last level goes to global memory (plus the per-tile channel
// sums for the ChannelGate).  Rows outside the image stay zero in both buffers, which is exactly the zero padding
// every layer of the reference applies; level l only computes the rows level `depth` still needs.
// grid = (row tiles, 4 branches (deepest first), crops); ~12 % redundant halo rows at R = 16, none for full-height tiles.
struct ChainArgs {
    const float* in;        // conv1 output [crops][H][W][C]
    float* out[4];          // final activation of each branch
    const float* wpw[10];   // per LightConv (index = branch*(branch+1)/2 + level-1)
    const float* wdw[10];
    const float* bias[10];
    float* sums[4];         // [crops][tiles][C]
    int H;
};

template <int C, int W, int R, int NT>
__host__ __device__ static constexpr int chain_n_pxp() {
    return (((R + 8) * (W + 2) + 64 + 7) / 8) * 8 + 2;   // room for a trailing 64-pixel group; planes 8 banks apart
}
template <int C, int W, int R, int NT>
static constexpr size_t chain_smem_bytes() {
    return sizeof(float) * ((size_t)2 * chain_n_pxp<C, W, R, NT>() * C + 4 * ((size_t)C * C + 9 * C + C));
}

template <int C, int W, int R, int NT>
__global__ void __launch_bounds__(NT) k_lightchain(const ChainArgs a, const int* __restrict__ d_n, int off, int cap) {
    const int n = blockIdx.z;
    if (n >= chunk_count(d_n, off, cap)) return;
    const int br = 3 - (int)blockIdx.y, depth = br + 1, tile = blockIdx.x;
    const int l0 = br * (br + 1) / 2;
    const int H = a.H;
    constexpr int TW = W + 2, C4 = C / 4, PPL = 2, G = 32 * PPL;   // buffers hold R + 8 padded rows
    constexpr int n_pxp = chain_n_pxp<C, W, R, NT>();
    extern __shared__ __align__(16) float smem[];
    float4* sX = reinterpret_cast<float4*>(smem);     // [C4][n_pxp]  level input (planar: K-chunk major)
    float4* sT = sX + (size_t)C4 * n_pxp;             // [C4][n_pxp]  1x1 result
    float* sW = reinterpret_cast<float*>(sT + (size_t)C4 * n_pxp);   // [4][C][C]
    float* sD = sW + 4 * C * C;                       // [4][9][C]
    float* sB = sD + 4 * 9 * C;                       // [4][C]
    float* sP = reinterpret_cast<float*>(sT);         // [NT / C4][C]  (aliases T: only used after the last level)
    const int y0 = tile * R;
    const int g0 = y0 - 4;                            // global row of local row 0
    const float* in = a.in + (size_t)n * H * W * C;
    // ---- stage: zero both buffers, weights of this branch, the input rows level 1 needs ----
    for (int e = threadIdx.x; e < 2 * C4 * n_pxp; e += NT) sX[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < depth; ++l) {
        for (int e = threadIdx.x; e < C * C / 4; e += NT)
            reinterpret_cast<float4*>(sW + l * C * C)[e] = reinterpret_cast<const float4*>(a.wpw[l0 + l])[e];
        for (int e = threadIdx.x; e < 9 * C; e += NT) sD[l * 9 * C + e] = a.wdw[l0 + l][e];
        for (int e = threadIdx.x; e < C; e += NT) sB[l * C + e] = a.bias[l0 + l][e];
    }
    __syncthreads();   // the zero fill must land before cp.async writes into the same buffer
    {
        const int ga = max(y0 - depth, 0), gb = min(y0 + R + depth, H);
        const int n_in = (gb - ga) * W * C4;
        for (int e = threadIdx.x; e < n_in; e += NT) {
            const int ch = e % C4, px = e / C4;
            const int gy = ga + px / W, gx = px % W;
            cp_async16(sX + ch * n_pxp + (gy - g0) * TW + gx + 1, in + ((size_t)gy * W + gx) * C + ch * 4);
        }
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 0;");
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int n_cc = C >> 3;
    constexpr int walkers = W * C4;
    constexpr int n_grp = NT / C4;
    constexpr int act = n_grp * C4;
    constexpr int n_split = (act / walkers) < 1 ? 1 : (act / walkers);
    float4 psum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int lv = 1; lv <= depth; ++lv) {
        const int ext = depth - lv;                                   // extra rows each side this level still feeds
        const float* w = sW + (lv - 1) * C * C;
        // ---- phase A: T = X * Wpw on image rows [ya-1, yb+1) of this level (whole padded rows) ----
        const int ya = max(y0 - ext, 0), yb = min(y0 + R + ext, H);   // rows phase B produces
        {
            const int ta = max(ya - 1, 0), tb = min(yb + 1, H);
            const int pa = (ta - g0) * TW, pb = (tb - g0) * TW;
            const int n_pg = (pb - pa + G - 1) / G;
            for (int item = warp; item < n_pg * n_cc; item += NT / 32) {
                const int pg = item % n_pg, cc = (item / n_pg) * 8;
                const int p0 = pa + pg * G + lane;
                float acc[PPL][8];
#pragma unroll
                for (int q = 0; q < PPL; ++q)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
#pragma unroll 2
                for (int kc = 0; kc < C4; ++kc) {
                    float xv[PPL][4];
#pragma unroll
                    for (int q = 0; q < PPL; ++q) {
                        const float4 x = sX[kc * n_pxp + p0 + 32 * q];
                        xv[q][0] = x.x; xv[q][1] = x.y; xv[q][2] = x.z; xv[q][3] = x.w;
                    }
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const float4 w0 = *reinterpret_cast<const float4*>(w + (kc * 4 + kk) * C + cc);
                        const float4 w1 = *reinterpret_cast<const float4*>(w + (kc * 4 + kk) * C + cc + 4);
                        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                        for (int q = 0; q < PPL; ++q)
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[q][j] = fmaf(xv[q][kk], wv[j], acc[q][j]);
                    }
                }
#pragma unroll
                for (int q = 0; q < PPL; ++q) {
                    const int p = p0 + 32 * q;
                    if (p < pb) {
                        sT[(cc / 4) * n_pxp + p] = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
                        sT[(cc / 4 + 1) * n_pxp + p] = make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
                    }
                }
            }
        }
        __syncthreads();
        // ---- phase B: depthwise 3x3 + bias + ReLU on rows [ya, yb): next level's X, or the branch output ----
        const bool last = lv == depth;
        const int rows_lv = yb - ya;
        const int rows_per = (rows_lv + n_split - 1) / n_split;
        float* outp = a.out[br] + (size_t)n * H * W * C;
        if (threadIdx.x < act) {
            for (int wk = threadIdx.x; wk < walkers * n_split; wk += act) {
                const int c4 = wk % C4, x = (wk / C4) % W, sp = wk / walkers;
                const int ra = ya + sp * rows_per, rb = min(ra + rows_per, yb);
                if (ra >= rb) continue;
                const float* dwp = sD + (lv - 1) * 9 * C + c4 * 4;
                float4 wv[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) wv[t] = *reinterpret_cast<const float4*>(dwp + t * C);
                const float4 bv = *reinterpret_cast<const float4*>(sB + (lv - 1) * C + c4 * 4);
                const float4* tbase = sT + c4 * n_pxp + x;              // column x-1 of the padded row
                float4* xdst = sX + c4 * n_pxp + x + 1;
                float4 win[3][3];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    win[0][kx] = tbase[(ra - 1 - g0) * TW + kx];
                    win[1][kx] = tbase[(ra - g0) * TW + kx];
                }
                for (int y = ra; y < rb; y += 3) {
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        if (y + u < rb) {
                            const int i0 = u % 3, i1 = (u + 1) % 3, i2 = (u + 2) % 3;
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) win[i2][kx] = tbase[(y + u + 1 - g0) * TW + kx];
                            float4 acc = bv;
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const float4 w0 = wv[kx], w1 = wv[3 + kx], w2 = wv[6 + kx];
                                const float4 t0 = win[i0][kx], t1 = win[i1][kx], t2 = win[i2][kx];
                                acc.x = fmaf(t0.x, w0.x, acc.x); acc.y = fmaf(t0.y, w0.y, acc.y);
                                acc.z = fmaf(t0.z, w0.z, acc.z); acc.w = fmaf(t0.w, w0.w, acc.w);
                                acc.x = fmaf(t1.x, w1.x, acc.x); acc.y = fmaf(t1.y, w1.y, acc.y);
                                acc.z = fmaf(t1.z, w1.z, acc.z); acc.w = fmaf(t1.w, w1.w, acc.w);
                                acc.x = fmaf(t2.x, w2.x, acc.x); acc.y = fmaf(t2.y, w2.y, acc.y);
                                acc.z = fmaf(t2.z, w2.z, acc.z); acc.w = fmaf(t2.w, w2.w, acc.w);
                            }
                            acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f);
                            acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
                            if (last) {
                                *reinterpret_cast<float4*>(outp + ((size_t)(y + u) * W + x) * C + c4 * 4) = acc;
                                psum.x += acc.x; psum.y += acc.y; psum.z += acc.z; psum.w += acc.w;
                            } else {
                                xdst[(y + u - g0) * TW] = acc;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    // per-tile channel sums of the branch output (fixed slot per thread, fixed combination order)
    if (threadIdx.x < act)
        *reinterpret_cast<float4*>(sP + (threadIdx.x / C4) * C + (threadIdx.x % C4) * 4) = psum;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += NT) {
        float s = 0.f;
        for (int g = 0; g < n_grp; ++g) s += sP[g * C + c];
        a.sums[br][((size_t)n * gridDim.x + tile) * C + c] = s;
    }
}

// K5 v2.  Same contract as k_pointwise.  ncu on v1: long-scoreboard bound (global loads serialised with the math:
// load chunk -> sync -> compute -> sync) and 16 scalar shared stores per thread per chunk for the k-major transpose.
// v2 streams A with cp.async straight into a K-chunk-planar layout As[4][BM] of float4 (rows interleaved over the
// threads so consecutive lanes read consecutive float4: conflict-free, no transpose) and double-buffers the chunks,
// so the copy of chunk c+1 overlaps the FMAs of chunk c.  The GATED prologue (gate-weighted branch sum) still goes
// through registers.
template <int BN, bool GATED, int NT = 256>
__global__ void __launch_bounds__(NT) k_pointwise2(const PwArgs a, const int* __restrict__ d_n, int off, int cap) {
    constexpr int NTN = BN / 4, NTM = NT / NTN, BM = NTM * 8, BK = 16, BMP = BM + 2;
    constexpr int STAGE_F4 = 4 * BMP + BK * BN / 4;   // float4 per stage: A planes + B rows
    const int M = chunk_count(d_n, off, cap) * a.HW;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (m0 >= M) return;
    extern __shared__ __align__(16) float4 pw_smem[];
    const int tn = threadIdx.x % NTN, tm = threadIdx.x / NTN;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int K = a.K, N = a.N;
    const int KX = GATED ? K - a.mid : 0;
    const int n_chunks = (K + BK - 1) / BK;
    auto load_chunk = [&](int c) {
        float4* As = pw_smem + (size_t)(c & 1) * STAGE_F4;
        float4* Bs = As + 4 * BMP;
        const int k0 = c * BK;
        for (int e = threadIdx.x; e < BM * 4; e += NT) {
            const int r = e >> 2, kq = e & 3;
            const int m = m0 + r, k = k0 + kq * 4;
            float4* dst = As + kq * BMP + r;
            if (m < M && k < K) {
                if (!GATED) {
                    cp_async16(dst, a.in + (size_t)m * K + k);
                } else if (k < a.mid) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float* g = a.gates + (size_t)(m / a.HW) * 4 * a.mid + k;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const float4 x = *reinterpret_cast<const float4*>(a.branch[b] + (size_t)m * a.mid + k);
                        const float4 gg = *reinterpret_cast<const float4*>(g + b * a.mid);
                        v.x = fmaf(x.x, gg.x, v.x); v.y = fmaf(x.y, gg.y, v.y);
                        v.z = fmaf(x.z, gg.z, v.z); v.w = fmaf(x.w, gg.w, v.w);
                    }
                    *dst = v;
                } else {
                    cp_async16(dst, a.in + (size_t)m * KX + (k - a.mid));
                }
            } else {
                *dst = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        for (int e = threadIdx.x; e < BK * BN / 4; e += NT) {
            const int kk = e / (BN / 4), c4 = e % (BN / 4);
            if (k0 + kk < K && n0 + c4 * 4 < N) cp_async16(Bs + e, a.w + (size_t)(k0 + kk) * N + n0 + c4 * 4);
            else Bs[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("cp.async.commit_group;");
    };
    load_chunk(0);
    for (int c = 0; c < n_chunks; ++c) {
        if (c + 1 < n_chunks) {
            load_chunk(c + 1);
            asm volatile("cp.async.wait_group 1;");
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
        __syncthreads();
        const float4* As = pw_smem + (size_t)(c & 1) * STAGE_F4;
        const float4* Bs = As + 4 * BMP;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            float xv[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 x = As[kc * BMP + tm + NTM * i];
                xv[i][0] = x.x; xv[i][1] = x.y; xv[i][2] = x.z; xv[i][3] = x.w;
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 b = Bs[(kc * 4 + kk) * (BN / 4) + tn];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc[i][0] = fmaf(xv[i][kk], b.x, acc[i][0]); acc[i][1] = fmaf(xv[i][kk], b.y, acc[i][1]);
                    acc[i][2] = fmaf(xv[i][kk], b.z, acc[i][2]); acc[i][3] = fmaf(xv[i][kk], b.w, acc[i][3]);
                }
            }
        }
        __syncthreads();
    }
    const int cn = n0 + tn * 4;
    if (cn >= N) return;
    const float4 bv = *reinterpret_cast<const float4*>(a.bias + cn);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + tm + NTM * i;
        if (m >= M) continue;
        float4 v = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
        if (a.residual) {
            const float4 r = *reinterpret_cast<const float4*>(a.residual + (size_t)m * N + cn);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (a.relu == 2) { v.x = fminf(v.x, 6.f); v.y = fminf(v.y, 6.f); v.z = fminf(v.z, 6.f); v.w = fminf(v.w, 6.f); }  // ReLU6
        *reinterpret_cast<float4*>(a.out + (size_t)m * N + cn) = v;
    }
}

// K6b (MobileNetV2): 3x3 stride-2 stem (3 -> C0) + folded BN + ReLU6 : (N,256,128,3) -> (N,128,64,C0)
__global__ void k_stem3(const float* __restrict__ blob, const float* __restrict__ w, const float* __restrict__ bias,
                        int C0, const int* __restrict__ d_n, int off, int cap, float* __restrict__ out) {
    const int n_crops = chunk_count(d_n, off, cap);
    const int C4 = C0 / 4;
    const size_t total = (size_t)n_crops * 128 * 64 * C4;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(e % C4);
        size_t r = e / C4;
        const int ox = (int)(r % 64); r /= 64;
        const int oy = (int)(r % 128);
        const int n = (int)(r / 128);
        float4 acc = *reinterpret_cast<const float4*>(bias + c4 * 4);
        const float* src = blob + (size_t)n * IN_H * IN_W * 3;
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy * 2 - 1 + ky;
            if (iy < 0 || iy >= IN_H) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = ox * 2 - 1 + kx;
                if (ix < 0 || ix >= IN_W) continue;
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const float a = src[((size_t)iy * IN_W + ix) * 3 + ci];
                    const float4 wv = *reinterpret_cast<const float4*>(w + (size_t)((ky * 3 + kx) * 3 + ci) * C0 + c4 * 4);
                    acc.x = fmaf(a, wv.x, acc.x); acc.y = fmaf(a, wv.y, acc.y);
                    acc.z = fmaf(a, wv.z, acc.z); acc.w = fmaf(a, wv.w, acc.w);
                }
            }
        }
        acc.x = fminf(fmaxf(acc.x, 0.f), 6.f); acc.y = fminf(fmaxf(acc.y, 0.f), 6.f);
        acc.z = fminf(fmaxf(acc.z, 0.f), 6.f); acc.w = fminf(fmaxf(acc.w, 0.f), 6.f);
        *reinterpret_cast<float4*>(out + (((size_t)n * 128 + oy) * 64 + ox) * C0 + c4 * 4) = acc;
    }
}

// K6c (MobileNetV2): depthwise 3x3, stride 1 or 2, pad 1, + folded BN + ReLU6, NHWC float4 channels
__global__ void k_dwconv3(const float* __restrict__ in, int H, int W, int C, int stride, const float* __restrict__ w9c,
                          const float* __restrict__ bias, const int* __restrict__ d_n, int off, int cap,
                          float* __restrict__ out) {
    const int n_crops = chunk_count(d_n, off, cap);
    const int OH = H / stride, OW = W / stride, C4 = C / 4;
    const size_t total = (size_t)n_crops * OH * OW * C4;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(e % C4);
        size_t r = e / C4;
        const int ox = (int)(r % OW); r /= OW;
        const int oy = (int)(r % OH);
        const int n = (int)(r / OH);
        float4 acc = *reinterpret_cast<const float4*>(bias + c4 * 4);
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy * stride - 1 + ky;
            if (iy < 0 || iy >= H) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = ox * stride - 1 + kx;
                if (ix < 0 || ix >= W) continue;
                const float4 v = *reinterpret_cast<const float4*>(in + (((size_t)n * H + iy) * W + ix) * C + c4 * 4);
                const float4 wv = *reinterpret_cast<const float4*>(w9c + (size_t)(ky * 3 + kx) * C + c4 * 4);
                acc.x = fmaf(v.x, wv.x, acc.x); acc.y = fmaf(v.y, wv.y, acc.y);
                acc.z = fmaf(v.z, wv.z, acc.z); acc.w = fmaf(v.w, wv.w, acc.w);
            }
        }
        acc.x = fminf(fmaxf(acc.x, 0.f), 6.f); acc.y = fminf(fmaxf(acc.y, 0.f), 6.f);
        acc.z = fminf(fmaxf(acc.z, 0.f), 6.f); acc.w = fminf(fmaxf(acc.w, 0.f), 6.f);
        *reinterpret_cast<float4*>(out + (((size_t)n * OH + oy) * OW + ox) * C + c4 * 4) = acc;
    }
}

// K7: ChannelGate (osnet.py:161-210): mean -> fc1 -> ReLU -> fc2 -> sigmoid, for the four branches of a block.
struct GateArgs {
    const float* sums[4];  // [crops][tiles][C]
    const float* w1; const float* b1; const float* w2; const float* b2;  // [C][hid], [hid], [hid][C], [C]
    float* gates;          // [crops][4][C]
    int C, hid, tiles, HW;
};
__global__ void k_gates(const GateArgs a, const int* __restrict__ d_n, int off, int cap) {
    const int n = blockIdx.x;
    if (n >= chunk_count(d_n, off, cap)) return;
    extern __shared__ __align__(16) float smem[];
    float* mean = smem;            // [4][C]
    float* hid = smem + 4 * a.C;   // [4][hid]
    const int C = a.C;
    for (int e = threadIdx.x; e < 4 * C; e += blockDim.x) {
        const int b = e / C, c = e - b * C;
        float s = 0.f;
        for (int t = 0; t < a.tiles; ++t) s += a.sums[b][((size_t)n * a.tiles + t) * C + c];
        mean[e] = s / (float)a.HW;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 4 * a.hid; e += blockDim.x) {
        const int b = e / a.hid, h = e - b * a.hid;
        float s = a.b1[h];
        for (int c = 0; c < C; ++c) s = fmaf(mean[b * C + c], a.w1[(size_t)c * a.hid + h], s);
        hid[e] = fmaxf(s, 0.f);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 4 * C; e += blockDim.x) {
        const int b = e / C, c = e - b * C;
        float s = a.b2[c];
        for (int h = 0; h < a.hid; ++h) s = fmaf(hid[b * a.hid + h], a.w2[(size_t)h * C + c], s);
        a.gates[(size_t)n * 4 * C + e] = 1.0f / (1.0f + expf(-s));
    }
}

// K8: head: global average pool over HW, fc (+ folded BatchNorm1d) + ReLU, row-wise L2 normalisation, scatter
// to the caller's row (base_backend.py:197-207).  One CTA per crop.
__global__ void k_head(const float* __restrict__ x, int HW, int C, const float* __restrict__ wfc,
                       const float* __restrict__ bfc, int FEAT, const CropDesc* __restrict__ crops,
                       const int* __restrict__ d_n, int off, int cap, float* __restrict__ out, int out_ld) {
    const int n = blockIdx.x;
    if (n >= chunk_count(d_n, off, cap)) return;
    extern __shared__ __align__(16) float smem[];
    float* pooled = smem;       // [C]
    float* red = smem + C;      // [32]
    const float* xp = x + (size_t)n * HW * C;
    // global average pool: consecutive threads read consecutive channels of one pixel (coalesced); the pixel
    // stripes of the thread groups are combined in a fixed order so the result is run-to-run deterministic
    float* part = red + 32;  // [groups][C]
    const int groups = blockDim.x / C > 0 ? blockDim.x / C : 1;
    if ((int)threadIdx.x < groups * C) {
        const int c = threadIdx.x % C, g = threadIdx.x / C;
        // four independent partial sums keep several loads in flight (fixed combination order: deterministic)
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int p = g;
        for (; p + 3 * groups < HW; p += 4 * groups) {
            s0 += xp[(size_t)p * C + c];
            s1 += xp[(size_t)(p + groups) * C + c];
            s2 += xp[(size_t)(p + 2 * groups) * C + c];
            s3 += xp[(size_t)(p + 3 * groups) * C + c];
        }
        for (; p < HW; p += groups) s0 += xp[(size_t)p * C + c];
        part[g * C + c] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        if (blockDim.x >= (unsigned)C) {
            for (int g = 0; g < groups; ++g) s += part[g * C + c];
        } else {
            for (int p = 0; p < HW; ++p) s += xp[(size_t)p * C + c];
        }
        pooled[c] = s / (float)HW;
    }
    __syncthreads();
    float* dst = out + (size_t)crops[off + n].out_row * out_ld;
    float sq = 0.f;
    for (int f = threadIdx.x; f < FEAT; f += blockDim.x) {
        float s;
        if (wfc) {
            // fc + folded BN1d + ReLU; eight independent chains hide the L2 latency of the weight rows
            float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            int c = 0;
            for (; c + 8 <= C; c += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = fmaf(pooled[c + u], wfc[(size_t)(c + u) * FEAT + f], t[u]);
            }
            for (; c < C; ++c) t[0] = fmaf(pooled[c], wfc[(size_t)c * FEAT + f], t[0]);
            s = bfc[f] + (((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7])));
            s = fmaxf(s, 0.f);
        } else {
            s = pooled[f];  // MobileNetV2: the pooled conv9 map is the embedding (mobilenetv2.py:186-193)
        }
        dst[f] = s;
        sq = fmaf(s, s, sq);
    }
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    const float nrm = sqrtf(tot);
    for (int f = threadIdx.x; f < FEAT; f += blockDim.x) dst[f] = dst[f] / nrm;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
enum { CLS_CROP = 0, CLS_STEM, CLS_MAXPOOL, CLS_POINTWISE, CLS_LIGHTCONV, CLS_GATES, CLS_AVGPOOL, CLS_HEAD };
struct LightW { size_t pw, dw, b; };
struct TcW { const float* w = nullptr; int Kpad = 0, Npad = 0; };  // packed weights of one pointwise layer
struct BlockW {
    int cin, cout, mid, hid, has_ds;
    size_t c1w, c1b;
    LightW light[10];
    size_t g1w, g1b, g2w, g2b;
    size_t cw, cb;
    TcW tc_c1, tc_c;
    TcW light_tc[10];   // LightConv 1x1 weights packed for the tcgen05 stage (mid % 16 == 0)
};

namespace tcx { struct Plan; }

struct MbBlock {
    int cin, cout, t, stride, cinp, midp, coutp;
    size_t we, be, wd, bd, wp, bp;
};

struct ReidModel {
    int arch = 1;                 // 1 OSNet, 2 MobileNetV2
    std::vector<MbBlock> mb;      // MobileNetV2 bottlenecks
    int mb_stem = 0, mb_stemp = 0, mb_last = 0;
    size_t mb_stem_w = 0, mb_stem_b = 0, mb_c9w = 0, mb_c9b = 0;
    int c[4] = {0, 0, 0, 0};
    int feat = 0;
    float* d_w = nullptr;
    size_t stem_w = 0, stem_b = 0;
    BlockW blocks[6];
    size_t trans_w[2] = {0, 0}, trans_b[2] = {0, 0};
    size_t c5w = 0, c5b = 0, fcw = 0, fcb = 0;
    TcW tc_trans[2], tc_c5;
    float* d_wtc = nullptr;    // all packed tensor-core weights
    bool use_tc = false;
    bool pw_small = true;   // 128-thread pointwise CTAs (measured 3 % faster than 256); BOXMOT_B200_PW_SMALL=0 for the 256-thread shape
    bool pw_v2 = true;      // BOXMOT_B200_PW_V1=1 selects the first-generation pointwise GEMM (A/B runs)
    bool light_small = false;  // BOXMOT_B200_LIGHT_SMALL=1: stage-2 LightConv as 8-row tiles, 128 threads, 4 CTAs / SM
    bool light_tc = false;     // BOXMOT_B200_LIGHT_TC=1: the stage-2 LightConv 1x1 stage on tcgen05 (tf32 x3)
    bool light_chain = true;   // BOXMOT_B200_LIGHT_CHAIN=0: per-level LightConv launches instead of whole-branch CTAs
    int chain_var = 2;         // BOXMOT_B200_CHAIN_VAR: 2 (default) stage 2 per level + stages 3-4 chained (measured best:
                               // 8-row stage-2 chain tiles recompute 25 % halo rows); 0 all chained; 1 stage-2 chain
                               // tiles of 16 rows with 512 threads (1 CTA / SM)
    bool light_v2 = true;   // BOXMOT_B200_LIGHT_V1=1 selects the first-generation LightConv kernel (A/B runs)
    // workspace for one chunk of crops
    int chunk = 256;
    float* blob = nullptr;
    float* bufA = nullptr;
    float* bufB = nullptr;
    float* x1 = nullptr;
    float* Y[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    float* sums[4] = {nullptr, nullptr, nullptr, nullptr};
    float* gates = nullptr;
    // per-kernel-class device timing (bench / profiles): events around every launch when enabled
    bool profile = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_cls;
    double prof_ms[REID_N_CLASSES] = {0};
    int prof_launches[REID_N_CLASSES] = {0};
    int preprocess = 0;        // 0 resize, 1 resize_pad
    tcx::Plan* tc = nullptr;   // tensor-core path (tcgen05 + TMA, reid_tc.cuh): the default for the widths it covers
    int debug_stop = -1;       // stop after this stage index and leave the tensor in debug_ptr
    const float* debug_ptr = nullptr;
    size_t debug_floats_per_crop = 0;
};

static const int kBranchOfLight[10] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3};
static const int kLevelOfLight[10] = {1, 1, 2, 1, 2, 3, 1, 2, 3, 4};
static const int kDepth[4] = {1, 2, 3, 4};

}  // namespace bmb
#include "reid_tc_host.cuh"
namespace bmb {
namespace tcx {
Plan* plan_build(ReidModel* m, const float* host_w);
bool plan_supported(const ReidModel* m);
}

ReidModel* reid_load(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ReID blob: ") + path);
    int32_t hdr[16];
    f.read(reinterpret_cast<char*>(hdr), sizeof(hdr));
    if (!f || (uint32_t)hdr[0] != BLOB_MAGIC || hdr[1] != 1 || (hdr[2] != 1 && hdr[2] != 2))
        throw std::runtime_error("not a version-1 .b200reid blob (export it with boxmot_b200.weights.export_blob)");
    ReidModel* m = new ReidModel();
    if (hdr[2] == 2) {
        // ---- MobileNetV2 (reid/backbones/mobilenetv2.py): stem, 17 inverted-residual blocks, conv9, GAP ----
        try {
            m->arch = 2;
            m->mb_stem = hdr[3];
            m->mb_stemp = (hdr[3] + 3) / 4 * 4;
            const int n_blocks = hdr[4];
            m->feat = hdr[7];
            const size_t n_floats = (size_t)hdr[8];
            if (n_blocks < 1 || n_blocks > 64 || m->feat < 1) throw std::runtime_error("bad MobileNetV2 blob header");
            std::vector<int32_t> table((size_t)n_blocks * 4);
            f.read(reinterpret_cast<char*>(table.data()), sizeof(int32_t) * table.size());
            std::vector<float> host(n_floats);
            f.read(reinterpret_cast<char*>(host.data()), sizeof(float) * n_floats);
            if (!f) throw std::runtime_error("truncated ReID blob");
            size_t o = 0;
            auto take = [&](size_t n) { size_t r = o; o += (n + 3) / 4 * 4; return r; };
            auto p4 = [](int n) { return (n + 3) / 4 * 4; };
            m->mb_stem_w = take((size_t)27 * m->mb_stemp);
            m->mb_stem_b = take(m->mb_stemp);
            int H = 128, Wd = 64;
            size_t max_x = (size_t)H * Wd * m->mb_stemp, max_e = 0, max_d = 0;
            for (int i = 0; i < n_blocks; ++i) {
                MbBlock b{};
                b.cin = table[i * 4]; b.cout = table[i * 4 + 1]; b.t = table[i * 4 + 2]; b.stride = table[i * 4 + 3];
                if (b.stride != 1 && b.stride != 2) throw std::runtime_error("bad MobileNetV2 stride");
                b.cinp = p4(b.cin); b.midp = p4(b.cin * b.t); b.coutp = p4(b.cout);
                b.we = take((size_t)b.cinp * b.midp); b.be = take(b.midp);
                b.wd = take((size_t)9 * b.midp); b.bd = take(b.midp);
                b.wp = take((size_t)b.midp * b.coutp); b.bp = take(b.coutp);
                max_e = std::max(max_e, (size_t)H * Wd * b.midp);
                H /= b.stride; Wd /= b.stride;
                max_d = std::max(max_d, (size_t)H * Wd * b.midp);
                max_x = std::max(max_x, (size_t)H * Wd * b.coutp);
                m->mb.push_back(b);
            }
            m->mb_last = m->mb.back().coutp;
            const int featp = p4(m->feat);
            m->mb_c9w = take((size_t)m->mb_last * featp);
            m->mb_c9b = take(featp);
            if (o != n_floats || featp != m->feat) throw std::runtime_error("ReID blob size does not match its header");
            max_x = std::max(max_x, (size_t)H * Wd * featp);
            RCUDA_OK(cudaMalloc(&m->d_w, sizeof(float) * n_floats));
            RCUDA_OK(cudaMemcpy(m->d_w, host.data(), sizeof(float) * n_floats, cudaMemcpyHostToDevice));
            m->chunk = 256;
            if (const char* ce = getenv("BOXMOT_B200_REID_CHUNK")) {
                const int v = atoi(ce);
                if (v >= 8 && v <= 1024) m->chunk = v;
            }
            const size_t CH = m->chunk;
            RCUDA_OK(cudaMalloc(&m->blob, sizeof(float) * CH * IN_H * IN_W * 3));
            RCUDA_OK(cudaMalloc(&m->bufA, sizeof(float) * CH * max_x));
            RCUDA_OK(cudaMalloc(&m->bufB, sizeof(float) * CH * max_x));
            RCUDA_OK(cudaMalloc(&m->x1, sizeof(float) * CH * max_e));
            RCUDA_OK(cudaMalloc(&m->Y[0][0], sizeof(float) * CH * max_d));
        } catch (...) {
            reid_free(m);
            throw;
        }
        return m;
    }
    try {
        for (int i = 0; i < 4; ++i) m->c[i] = hdr[3 + i];
        m->feat = hdr[7];
        const size_t n_floats = (size_t)hdr[8];
        if (m->c[0] % 16 != 0 || m->feat < 1) throw std::runtime_error("unsupported OSNet width (stem channels)");
        std::vector<float> host(n_floats);
        f.read(reinterpret_cast<char*>(host.data()), sizeof(float) * n_floats);
        if (!f) throw std::runtime_error("truncated ReID blob");
        size_t o = 0;
        auto take = [&](size_t n) { size_t r = o; o += (n + 3) / 4 * 4; return r; };  // 16-byte aligned tensors
        m->stem_w = take((size_t)147 * m->c[0]);
        m->stem_b = take(m->c[0]);
        for (int s = 0; s < 3; ++s) {
            for (int j = 0; j < 2; ++j) {
                BlockW& b = m->blocks[s * 2 + j];
                b.cin = j == 0 ? m->c[s] : m->c[s + 1];
                b.cout = m->c[s + 1];
                b.mid = b.cout / 4;
                b.hid = b.mid / 16;
                b.has_ds = b.cin != b.cout;
                if (b.mid % 8 != 0 || b.hid < 1) throw std::runtime_error("unsupported OSNet width (mid channels)");
                b.c1w = take((size_t)b.cin * b.mid);
                b.c1b = take(b.mid);
                for (int l = 0; l < 10; ++l) {
                    b.light[l].pw = take((size_t)b.mid * b.mid);
                    b.light[l].dw = take((size_t)9 * b.mid);
                    b.light[l].b = take(b.mid);
                }
                b.g1w = take((size_t)b.mid * b.hid);
                b.g1b = take(b.hid);
                b.g2w = take((size_t)b.hid * b.mid);
                b.g2b = take(b.mid);
                b.cw = take((size_t)(b.mid + (b.has_ds ? b.cin : 0)) * b.cout);
                b.cb = take(b.cout);
            }
            if (s < 2) {
                m->trans_w[s] = take((size_t)m->c[s + 1] * m->c[s + 1]);
                m->trans_b[s] = take(m->c[s + 1]);
            }
        }
        m->c5w = take((size_t)m->c[3] * m->c[3]);
        m->c5b = take(m->c[3]);
        m->fcw = take((size_t)m->c[3] * m->feat);
        m->fcb = take(m->feat);
        if (o != n_floats) throw std::runtime_error("ReID blob size does not match its header");
        RCUDA_OK(cudaMalloc(&m->d_w, sizeof(float) * n_floats));
        RCUDA_OK(cudaMemcpy(m->d_w, host.data(), sizeof(float) * n_floats, cudaMemcpyHostToDevice));
        {   // tensor-core copies of every 1x1 weight that fits the tcgen05 kernel's shared-memory budget
            const char* env = getenv("BOXMOT_B200_REID_TC");
            // Measured in round 1 (profiles/r1_tcgen05_pointwise.md): at OSNet_x0_25's K,N <= 128 every 1x1 layer is
            // bandwidth-bound and the float32 CUDA-core GEMM is 1.2-1.8x faster than this first (unpipelined)
            // tcgen05 kernel, so the tensor-core path is opt-in until it is pipelined / fused.
            m->use_tc = env && env[0] == '1';
            const char* lv = getenv("BOXMOT_B200_LIGHT_V1");
            m->light_v2 = !(lv && lv[0] == '1');
            if (const char* cv = getenv("BOXMOT_B200_LIGHT_CHAIN")) m->light_chain = !(cv[0] == '0');
            if (const char* cv = getenv("BOXMOT_B200_LIGHT_TC")) m->light_tc = cv[0] == '1';
            if (const char* cv = getenv("BOXMOT_B200_LIGHT_SMALL")) m->light_small = cv[0] == '1';
            if (const char* cv = getenv("BOXMOT_B200_CHAIN_VAR")) m->chain_var = atoi(cv);
            const char* pv = getenv("BOXMOT_B200_PW_V1");
            m->pw_v2 = !(pv && pv[0] == '1');
            if (const char* cv = getenv("BOXMOT_B200_PW_SMALL")) m->pw_small = !(cv[0] == '0');
            std::vector<float> packed;
            struct Todo { TcW* dst; size_t w; int K, N; size_t at; };
            std::vector<Todo> todo;
            auto add = [&](TcW* dst, size_t w, int K, int N) {
                const int Kpad = (K + 7) / 8 * 8, Npad = (N + 15) / 16 * 16;
                if (Npad > 256 || tc::smem_bytes(Kpad, Npad) > 200 * 1024) return;
                dst->Kpad = Kpad; dst->Npad = Npad;
                todo.push_back({dst, w, K, N, packed.size()});
                packed.resize(packed.size() + 2 * (size_t)Npad * Kpad);
            };
            for (int bi = 0; bi < 6; ++bi) {
                BlockW& b = m->blocks[bi];
                add(&b.tc_c1, b.c1w, b.cin, b.mid);
                add(&b.tc_c, b.cw, b.mid + (b.has_ds ? b.cin : 0), b.cout);
                if (b.mid % 16 == 0)
                    for (int l = 0; l < 10; ++l) add(&b.light_tc[l], b.light[l].pw, b.mid, b.mid);
            }
            for (int s = 0; s < 2; ++s) add(&m->tc_trans[s], m->trans_w[s], m->c[s + 1], m->c[s + 1]);
            add(&m->tc_c5, m->c5w, m->c[3], m->c[3]);
            for (auto& t : todo) tc::pack_weights(host.data() + t.w, t.K, t.N, t.dst->Kpad, t.dst->Npad, packed.data() + t.at);
            if (!packed.empty()) {
                RCUDA_OK(cudaMalloc(&m->d_wtc, sizeof(float) * packed.size()));
                RCUDA_OK(cudaMemcpy(m->d_wtc, packed.data(), sizeof(float) * packed.size(), cudaMemcpyHostToDevice));
                for (auto& t : todo) t.dst->w = m->d_wtc + t.at;
            }
        }
        // workspace
        if (const char* ce = getenv("BOXMOT_B200_REID_CHUNK")) {
            const int v = atoi(ce);
            if (v >= 8 && v <= 1024) m->chunk = v;
        }
        const size_t CH = m->chunk;
        const size_t big = (size_t)8192 * m->c[0] > (size_t)2048 * m->c[1] ? (size_t)8192 * m->c[0] : (size_t)2048 * m->c[1];
        const size_t mid_max = (size_t)2048 * (m->c[1] / 4);
        RCUDA_OK(cudaMalloc(&m->blob, sizeof(float) * CH * IN_H * IN_W * 3));
        RCUDA_OK(cudaMalloc(&m->bufA, sizeof(float) * CH * big));
        RCUDA_OK(cudaMalloc(&m->bufB, sizeof(float) * CH * big));
        RCUDA_OK(cudaMalloc(&m->x1, sizeof(float) * CH * mid_max));
        for (int b = 0; b < 4; ++b) {
            for (int k = 0; k < 2; ++k) RCUDA_OK(cudaMalloc(&m->Y[b][k], sizeof(float) * CH * mid_max));
            RCUDA_OK(cudaMalloc(&m->sums[b], sizeof(float) * CH * 64 * (m->c[3] / 4)));
        }
        RCUDA_OK(cudaMalloc(&m->gates, sizeof(float) * CH * 4 * (m->c[3] / 4)));
        {   // tensor-core path: the default wherever its kernel instances cover the widths (BOXMOT_B200_REID_FP32=1
            // keeps the float32 CUDA-core kernels of round 1, e.g. for A/B runs)
            const char* fe = getenv("BOXMOT_B200_REID_FP32");
            if (!(fe && fe[0] == '1') && tcx::plan_supported(m)) m->tc = tcx::plan_build(m, host.data());
        }
    } catch (...) {
        reid_free(m);
        throw;
    }
    return m;
}

void reid_free(ReidModel* m) {
    if (!m) return;
    cudaFree(m->d_w); cudaFree(m->d_wtc); cudaFree(m->blob); cudaFree(m->bufA); cudaFree(m->bufB); cudaFree(m->x1);
    for (int b = 0; b < 4; ++b) { cudaFree(m->Y[b][0]); cudaFree(m->Y[b][1]); cudaFree(m->sums[b]); }
    cudaFree(m->gates);
    tcx::plan_free(m->tc);
    delete m;
}

int reid_feature_dim(const ReidModel* m) { return m->feat; }
void reid_set_preprocess(ReidModel* m, int mode) { m->preprocess = mode ? 1 : 0; }
const float* reid_last_input_blob(const ReidModel* m) { return m->blob; }
void reid_set_profile(ReidModel* m, bool on) { m->profile = on; }
// Fold the events recorded since the last call into per-class totals (the stream must be idle).
void reid_profile_collect(ReidModel* m, double* ms, int* launches) {
    for (size_t i = 0; i < m->prof_cls.size(); ++i) {
        float t = 0.f;
        cudaEventElapsedTime(&t, m->prof_ev[2 * i], m->prof_ev[2 * i + 1]);
        m->prof_ms[m->prof_cls[i]] += t;
        m->prof_launches[m->prof_cls[i]] += 1;
        cudaEventDestroy(m->prof_ev[2 * i]);
        cudaEventDestroy(m->prof_ev[2 * i + 1]);
    }
    m->prof_ev.clear();
    m->prof_cls.clear();
    for (int c = 0; c < REID_N_CLASSES; ++c) {
        if (ms) ms[c] = m->prof_ms[c];
        if (launches) launches[c] = m->prof_launches[c];
        m->prof_ms[c] = 0;
        m->prof_launches[c] = 0;
    }
}
void reid_set_debug_stop(ReidModel* m, int stage) { m->debug_stop = stage; }
const float* reid_debug_tensor(const ReidModel* m, size_t* floats_per_crop) {
    if (floats_per_crop) *floats_per_crop = m->debug_floats_per_crop;
    return m->debug_ptr;
}

namespace {
struct Launcher {
    ReidModel* m;
    const int* d_n;
    int off, cap, upper;  // upper = host-side bound on crops in this chunk (grid sizing)
    cudaStream_t st;
    int launches = 0;

    void begin(int cls) {
        if (!m->profile) return;
        cudaEvent_t e0, e1;
        RCUDA_OK(cudaEventCreate(&e0));
        RCUDA_OK(cudaEventCreate(&e1));
        m->prof_ev.push_back(e0);
        m->prof_ev.push_back(e1);
        m->prof_cls.push_back(cls);
        RCUDA_OK(cudaEventRecord(e0, st));
    }
    void end() {
        if (!m->profile) return;
        RCUDA_OK(cudaEventRecord(m->prof_ev.back(), st));
    }

    void pointwise_tc(const PwArgs& a) {
        tc::Args t{};
        t.in = a.in;
        for (int b = 0; b < 4; ++b) t.branch[b] = a.branch[b];
        t.gates = a.gates; t.w_tc = a.w_tc; t.bias = a.bias; t.residual = a.residual; t.out = a.out;
        t.K = a.K; t.N = a.N; t.Kpad = a.Kpad; t.Npad = a.Npad; t.mid = a.mid; t.HW = a.HW; t.relu = a.relu;
        const size_t smem = tc::smem_bytes(a.Kpad, a.Npad);
        const int tmem_cols = a.Npad <= 32 ? 32 : (a.Npad <= 64 ? 64 : (a.Npad <= 128 ? 128 : 256));
        int per_sm = (int)((220 * 1024) / (smem + 1024));
        per_sm = per_sm < 1 ? 1 : (per_sm > 512 / tmem_cols ? 512 / tmem_cols : per_sm);
        per_sm = per_sm > 8 ? 8 : per_sm;
        const int tiles = (int)(((size_t)upper * a.HW) / tc::TILE_M);
        const int grid = tiles < 148 * per_sm ? tiles : 148 * per_sm;
        begin(CLS_POINTWISE);
        if (a.gates) {
            RCUDA_OK(cudaFuncSetAttribute(tc::k_pointwise_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            tc::k_pointwise_tc<true><<<grid, tc::THREADS, smem, st>>>(t, d_n, off, cap);
        } else {
            RCUDA_OK(cudaFuncSetAttribute(tc::k_pointwise_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            tc::k_pointwise_tc<false><<<grid, tc::THREADS, smem, st>>>(t, d_n, off, cap);
        }
        end();
        ++launches;
    }
    void pointwise(const PwArgs& a) {
        if (m->use_tc && a.w_tc && a.HW % tc::TILE_M == 0) { pointwise_tc(a); return; }
        const int N = a.N;
        const size_t Mmax = (size_t)upper * a.HW;
        if (N % 64 == 0) launch_pw<64>(a, Mmax);
        else if (N % 32 == 0 || N > 16) launch_pw<32>(a, Mmax);
        else launch_pw<16>(a, Mmax);
    }
    template <int BN>
    void launch_pw(const PwArgs& a, size_t Mmax) {
        constexpr int BM = (256 / (BN / 4)) * 8;
        dim3 grid((unsigned)((Mmax + BM - 1) / BM), (unsigned)((a.N + BN - 1) / BN));
        if (m->pw_v2 && m->pw_small) {   // 128-thread CTAs: half the rows per CTA, twice the CTAs (A/B switch)
            constexpr int BMs = (128 / (BN / 4)) * 8;
            constexpr size_t smem = 2 * sizeof(float4) * (4 * (BMs + 2) + 16 * BN / 4);
            dim3 g((unsigned)((Mmax + BMs - 1) / BMs), (unsigned)((a.N + BN - 1) / BN));
            begin(CLS_POINTWISE);
            if (a.gates) k_pointwise2<BN, true, 128><<<g, 128, smem, st>>>(a, d_n, off, cap);
            else k_pointwise2<BN, false, 128><<<g, 128, smem, st>>>(a, d_n, off, cap);
            end();
            ++launches;
            return;
        }
        if (m->pw_v2) {
            constexpr size_t smem = 2 * sizeof(float4) * (4 * (BM + 2) + 16 * BN / 4);
            begin(CLS_POINTWISE);
            if (a.gates) {
                if (smem > 48 * 1024)
                    RCUDA_OK(cudaFuncSetAttribute(k_pointwise2<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_pointwise2<BN, true><<<grid, 256, smem, st>>>(a, d_n, off, cap);
            } else {
                if (smem > 48 * 1024)
                    RCUDA_OK(cudaFuncSetAttribute(k_pointwise2<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_pointwise2<BN, false><<<grid, 256, smem, st>>>(a, d_n, off, cap);
            }
            end();
            ++launches;
            return;
        }
        begin(CLS_POINTWISE);
        if (a.gates) k_pointwise<BN, true><<<grid, 256, 0, st>>>(a, d_n, off, cap);
        else k_pointwise<BN, false><<<grid, 256, 0, st>>>(a, d_n, off, cap);
        end();
        ++launches;
    }
    template <int C, int W, int R, int PPL = 4, int MINB = 2, bool TC = false, int NT = 256>
    void launch_light2(const LightArgs& a, int n_branches) {
        const int tiles = (a.H + R - 1) / R;
        const size_t smem = light2_smem_bytes((R + 2) * (W + 2), C, NT, PPL, TC);
        if (smem > 48 * 1024)
            RCUDA_OK(cudaFuncSetAttribute(k_lightconv2<C, W, R, PPL, MINB, TC, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        begin(CLS_LIGHTCONV);
        k_lightconv2<C, W, R, PPL, MINB, TC, NT><<<dim3(tiles, n_branches, upper), NT, smem, st>>>(a, d_n, off, cap);
        end();
        ++launches;
    }
    template <int C, int W, int R, int NT>
    void launch_chain(const ChainArgs& a) {
        const int tiles = a.H / R;
        constexpr size_t smem = chain_smem_bytes<C, W, R, NT>();
        RCUDA_OK(cudaFuncSetAttribute(k_lightchain<C, W, R, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        begin(CLS_LIGHTCONV);
        k_lightchain<C, W, R, NT><<<dim3(tiles, 4, upper), NT, smem, st>>>(a, d_n, off, cap);
        end();
        ++launches;
    }
    // whole-branch LightConv chains for the osnet_x0_25 stage shapes; returns the tile rows used (0 = not covered)
    int light_chain(const ChainArgs& a, int C, int W) {
        if (C == 16 && W == 32) {
            if (m->chain_var == 2) return 0;   // stage 2 per level (8-row chain tiles recompute 25 % halo rows)
            if (m->chain_var == 1) { launch_chain<16, 32, 16, 512>(a); return 16; }
            launch_chain<16, 32, 8, 256>(a); return 8;
        }
        if (C == 24 && W == 16) { launch_chain<24, 16, 16, 256>(a); return 16; }
        if (C == 32 && W == 8) { launch_chain<32, 8, 16, 256>(a); return 16; }
        return 0;
    }
    // shape-specialised LightConv (the three OSBlock stages of osnet_x0_25 and osnet_x1_0); false = not covered
    bool light2(const LightArgs& a, int n_branches) {
#define BMB_LIGHT2(CC, WW, RR) \
        if (a.C == CC && a.W == WW && a.R == RR) { launch_light2<CC, WW, RR>(a, n_branches); return true; }
        if (m->light_small && a.C == 16 && a.W == 32 && a.R == 8) {   // four 128-thread CTAs per SM (A/B switch)
            launch_light2<16, 32, 8, 2, 4, false, 128>(a, n_branches);
            return true;
        }
        if (m->light_tc && a.C == 16 && a.W == 32 && a.R == 16 && a.wtc[0]) {   // tcgen05 1x1 stage (opt-in)
            launch_light2<16, 32, 16, 4, 2, true>(a, n_branches);
            return true;
        }
        BMB_LIGHT2(16, 32, 16) BMB_LIGHT2(24, 16, 16) BMB_LIGHT2(32, 8, 16)
        BMB_LIGHT2(64, 32, 8) BMB_LIGHT2(96, 16, 4) BMB_LIGHT2(128, 8, 8)
#undef BMB_LIGHT2
        return false;
    }
    void light(const LightArgs& a, int n_branches, int threads) {
        if (m->light_v2 && light2(a, n_branches)) return;
        const int tiles = (a.H + a.R - 1) / a.R;
        const int n_grp = threads / (a.C / 4);
        const size_t smem = sizeof(float) * ((size_t)(a.R + 2) * (a.W + 2) * a.C + (size_t)a.C * a.C + 9 * a.C +
                                             (size_t)n_grp * a.C);
        if (smem > 48 * 1024)
            RCUDA_OK(cudaFuncSetAttribute(k_lightconv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        begin(CLS_LIGHTCONV);
        k_lightconv<<<dim3(tiles, n_branches, upper), threads, smem, st>>>(a, d_n, off, cap);
        end();
        ++launches;
    }
};

// v2 keeps the staged input and the 1x1 result side by side: aim for two CTAs per SM, fall back to one
int pick_tile_rows2(int H, int W, int C) {
    for (size_t budget : {(size_t)100 * 1024, (size_t)220 * 1024})
        for (int R = H; R >= 1; --R) {
            if (H % R || R > 16) continue;
            if (light2_smem_bytes((R + 2) * (W + 2), C, 256) <= budget) return R;
        }
    return 1;
}

int pick_tile_rows(int H, int W, int C) {
    // largest divisor of H whose haloed tile (+ weights) stays under ~96 KB
    for (int R = H; R >= 1; --R) {
        if (H % R) continue;
        const size_t bytes = sizeof(float) * ((size_t)(R + 2) * (W + 2) * C + (size_t)C * C + 9 * C + 64 * C);
        if (bytes <= 96 * 1024 && R <= 16) return R;
    }
    return 1;
}
}  // namespace

}  // namespace bmb
#include "reid_tc_plan.cuh"
namespace bmb {

int reid_forward(ReidModel* m, const uint8_t* d_images, size_t image_stride, int rows, int cols,
                 const CropDesc* d_crops, const int* d_ncrops, int max_crops, float* d_out, int out_ld,
                 cudaStream_t st, int first_crop, int last_crop) {
    // [first_crop, last_crop) restricts the call to a slice of the crop list (two models on two streams split a frame)
    if (last_crop < 0 || last_crop > max_crops) last_crop = max_crops;
    int launches = 0;
    const float* W = m->d_w;
    m->debug_ptr = nullptr;
    if (m->arch == 2) {
        // MobileNetV2: stem -> [expand 1x1 + ReLU6 -> depthwise 3x3 + ReLU6 -> project 1x1 (+ residual)] x 17 -> conv9 -> GAP
        for (int off = first_crop; off < last_crop; off += m->chunk) {
            const int upper = (last_crop - off) < m->chunk ? (last_crop - off) : m->chunk;
            Launcher L{m, d_ncrops, off, upper, upper, st};
            L.begin(CLS_CROP);
            k_crop_resize_norm<<<upper, 256, 0, st>>>(d_images, image_stride, rows, cols, d_crops, d_ncrops, off, upper,
                                                      m->blob, m->preprocess);
            L.end();
            ++L.launches;
            float* X = m->bufA;
            float* Xo = m->bufB;
            L.begin(CLS_STEM);
            k_stem3<<<148 * 8, 256, 0, st>>>(m->blob, W + m->mb_stem_w, W + m->mb_stem_b, m->mb_stemp, d_ncrops, off,
                                              upper, X);
            L.end();
            ++L.launches;
            int H = 128, Wd = 64;
            for (const MbBlock& b : m->mb) {
                PwArgs e{};
                e.in = X; e.w = W + b.we; e.bias = W + b.be; e.out = m->x1;
                e.K = b.cinp; e.N = b.midp; e.HW = H * Wd; e.relu = 2;
                L.pointwise(e);
                L.begin(CLS_LIGHTCONV);
                k_dwconv3<<<148 * 8, 256, 0, st>>>(m->x1, H, Wd, b.midp, b.stride, W + b.wd, W + b.bd, d_ncrops, off,
                                                    upper, m->Y[0][0]);
                L.end();
                ++L.launches;
                H /= b.stride; Wd /= b.stride;
                PwArgs p{};
                p.in = m->Y[0][0]; p.w = W + b.wp; p.bias = W + b.bp; p.out = Xo;
                p.residual = (b.stride == 1 && b.cin == b.cout) ? X : nullptr;
                p.K = b.midp; p.N = b.coutp; p.HW = H * Wd; p.relu = 0;
                L.pointwise(p);
                float* t = X; X = Xo; Xo = t;
            }
            PwArgs c9{};
            c9.in = X; c9.w = W + m->mb_c9w; c9.bias = W + m->mb_c9b; c9.out = Xo;
            c9.K = m->mb_last; c9.N = m->feat; c9.HW = H * Wd; c9.relu = 2;
            L.pointwise(c9);
            L.begin(CLS_HEAD);
            k_head<<<upper, 256, sizeof(float) * (2 * m->feat + 32), st>>>(Xo, H * Wd, m->feat, nullptr, nullptr, m->feat,
                                                                           d_crops, d_ncrops, off, upper, d_out, out_ld);
            L.end();
            ++L.launches;
            launches += L.launches;
        }
        RCUDA_OK(cudaGetLastError());
        return launches;
    }
    for (int off = first_crop; off < last_crop; off += m->chunk) {
        const int upper = (last_crop - off) < m->chunk ? (last_crop - off) : m->chunk;
        Launcher L{m, d_ncrops, off, upper, upper, st};
        int stage_idx = 0;
        auto stop_here = [&](const float* ptr, size_t per_crop) {
            if (m->debug_stop == stage_idx) { m->debug_ptr = ptr; m->debug_floats_per_crop = per_crop; ++stage_idx; return true; }
            ++stage_idx;
            return false;
        };
        if (m->tc && m->debug_stop != 0 && m->debug_stop != 1) {
            // tensor-core path: crop staging, stem and max pool are one fused kernel (k_front_tc); diagnostic stops at the
            // blob / stem tensors (0, 1) run the float32 kernels below instead
            bool tc_stopped = false;
            const tcx::FrontInput fi{d_images, image_stride, rows, cols, d_crops, d_out, out_ld};
            L.launches += tcx::plan_run(m, fi, d_ncrops, off, upper, st, &tc_stopped, L);
            if (!tc_stopped && !(m->tc->head_fused && m->debug_stop < 0)) {
                const int C = m->c[3];
                L.begin(CLS_HEAD);
                k_head<<<upper, 256, sizeof(float) * (C + 32 + (256 / C > 0 ? 256 / C : 1) * C), st>>>(
                    m->tc->c5, 128, C, W + m->fcw, W + m->fcb, m->feat, d_crops, d_ncrops, off, upper, d_out, out_ld);
                L.end();
                ++L.launches;
            }
            launches += L.launches;
            continue;
        }
        L.begin(CLS_CROP);
        k_crop_resize_norm<<<upper, 256, 0, st>>>(d_images, image_stride, rows, cols, d_crops, d_ncrops, off, upper,
                                                  m->blob, m->preprocess);
        L.end();
        ++L.launches;
        if (stop_here(m->blob, (size_t)IN_H * IN_W * 3)) { launches += L.launches; continue; }
        {
            const size_t smem = sizeof(float) * ((size_t)((ST_IR * ST_IC * 3 + 3) & ~3) + 147 * 16);
            RCUDA_OK(cudaFuncSetAttribute(k_stem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            L.begin(CLS_STEM);
            k_stem<<<dim3(128 / ST_R, upper), 256, smem, st>>>(m->blob, W + m->stem_w, W + m->stem_b, m->c[0], d_ncrops,
                                                                off, upper, m->bufA);
            L.end();
            ++L.launches;
        }
        if (stop_here(m->bufA, (size_t)8192 * m->c[0])) { launches += L.launches; continue; }
        L.begin(CLS_MAXPOOL);
        k_maxpool3s2<<<148 * 8, 256, 0, st>>>(m->bufA, 128, 64, m->c[0], d_ncrops, off, upper, m->bufB);
        L.end();
        ++L.launches;
        if (stop_here(m->bufB, (size_t)2048 * m->c[0])) { launches += L.launches; continue; }
        float* X = m->bufB;
        float* Xo = m->bufA;
        int H = 64, Wd = 32;
        bool stopped = false;
        for (int s = 0; s < 3 && !stopped; ++s) {
            for (int j = 0; j < 2 && !stopped; ++j) {
                const BlockW& b = m->blocks[s * 2 + j];
                const int HW = H * Wd;
                PwArgs p{};
                p.in = X; p.w = W + b.c1w; p.bias = W + b.c1b; p.out = m->x1;
                p.K = b.cin; p.N = b.mid; p.HW = HW; p.relu = 1;
                p.w_tc = b.tc_c1.w; p.Kpad = b.tc_c1.Kpad; p.Npad = b.tc_c1.Npad;
                L.pointwise(p);
                int R = 0;
                if (m->light_chain && m->light_v2) {
                    ChainArgs ca{};
                    ca.in = m->x1; ca.H = H;
                    for (int br = 0; br < 4; ++br) { ca.out[br] = m->Y[br][kDepth[br] & 1]; ca.sums[br] = m->sums[br]; }
                    for (int l = 0; l < 10; ++l) {
                        ca.wpw[l] = W + b.light[l].pw; ca.wdw[l] = W + b.light[l].dw; ca.bias[l] = W + b.light[l].b;
                    }
                    R = L.light_chain(ca, b.mid, Wd);
                }
                const bool chained = R > 0;
                if (!chained) R = m->light_v2 ? pick_tile_rows2(H, Wd, b.mid) : pick_tile_rows(H, Wd, b.mid);
                if (!chained && m->light_small && m->light_v2 && b.mid == 16 && Wd == 32) R = 8;
                const int tiles = H / R;
                const int threads = (256 / (b.mid / 4)) * (b.mid / 4);  // a multiple of the channel groups
                for (int level = 1; level <= 4 && !chained; ++level) {
                    LightArgs la{};
                    la.H = H; la.W = Wd; la.C = b.mid; la.R = R;
                    int nb = 0;
                    for (int l = 0; l < 10; ++l) {
                        if (kLevelOfLight[l] != level) continue;
                        const int br = kBranchOfLight[l];
                        la.in[nb] = level == 1 ? m->x1 : m->Y[br][(level - 1) & 1];
                        la.out[nb] = m->Y[br][level & 1];
                        la.wpw[nb] = W + b.light[l].pw;
                        la.wdw[nb] = W + b.light[l].dw;
                        la.bias[nb] = W + b.light[l].b;
                        la.wtc[nb] = b.light_tc[l].w;
                        la.sums[nb] = kDepth[br] == level ? m->sums[br] : nullptr;
                        ++nb;
                    }
                    L.light(la, nb, threads);
                }
                GateArgs ga{};
                for (int br = 0; br < 4; ++br) ga.sums[br] = m->sums[br];
                ga.w1 = W + b.g1w; ga.b1 = W + b.g1b; ga.w2 = W + b.g2w; ga.b2 = W + b.g2b;
                ga.gates = m->gates; ga.C = b.mid; ga.hid = b.hid; ga.tiles = tiles; ga.HW = HW;
                L.begin(CLS_GATES);
                k_gates<<<upper, 128, sizeof(float) * (4 * b.mid + 4 * b.hid), st>>>(ga, d_ncrops, off, upper);
                L.end();
                ++L.launches;
                PwArgs c{};
                for (int br = 0; br < 4; ++br) c.branch[br] = m->Y[br][kDepth[br] & 1];
                c.gates = m->gates; c.mid = b.mid;
                c.in = b.has_ds ? X : nullptr;
                c.residual = b.has_ds ? nullptr : X;
                c.w = W + b.cw; c.bias = W + b.cb; c.out = Xo;
                c.K = b.mid + (b.has_ds ? b.cin : 0); c.N = b.cout; c.HW = HW; c.relu = 1;
                c.w_tc = b.tc_c.w; c.Kpad = b.tc_c.Kpad; c.Npad = b.tc_c.Npad;
                L.pointwise(c);
                float* t = X; X = Xo; Xo = t;
                if (stop_here(X, (size_t)HW * b.cout)) { stopped = true; break; }
            }
            if (stopped) break;
            if (s < 2) {
                const int C = m->c[s + 1];
                PwArgs p{};
                p.in = X; p.w = W + m->trans_w[s]; p.bias = W + m->trans_b[s]; p.out = Xo;
                p.K = C; p.N = C; p.HW = H * Wd; p.relu = 1;
                p.w_tc = m->tc_trans[s].w; p.Kpad = m->tc_trans[s].Kpad; p.Npad = m->tc_trans[s].Npad;
                L.pointwise(p);
                L.begin(CLS_AVGPOOL);
                k_avgpool2<<<148 * 4, 256, 0, st>>>(Xo, H, Wd, C, d_ncrops, off, upper, X);
                L.end();
                ++L.launches;
                H /= 2; Wd /= 2;
                if (stop_here(X, (size_t)H * Wd * C)) { stopped = true; break; }
            }
        }
        if (!stopped) {
            const int C = m->c[3];
            PwArgs p{};
            p.in = X; p.w = W + m->c5w; p.bias = W + m->c5b; p.out = Xo;
            p.K = C; p.N = C; p.HW = H * Wd; p.relu = 1;
            p.w_tc = m->tc_c5.w; p.Kpad = m->tc_c5.Kpad; p.Npad = m->tc_c5.Npad;
            L.pointwise(p);
            if (!stop_here(Xo, (size_t)H * Wd * C)) {
                L.begin(CLS_HEAD);
                k_head<<<upper, 256, sizeof(float) * (C + 32 + (256 / C > 0 ? 256 / C : 1) * C), st>>>(Xo, H * Wd, C, W + m->fcw, W + m->fcb, m->feat,
                                                                     d_crops, d_ncrops, off, upper, d_out, out_ld);
                L.end();
                ++L.launches;
            }
        }
        launches += L.launches;
    }
    RCUDA_OK(cudaGetLastError());
    return launches;
}


// Standalone 1x1-convolution GEMM on host arrays (parity tests / micro-benchmarks): out = act(A W + bias (+ res)).
void standalone_pointwise(const float* A, int M, int K, const float* W, int N, const float* bias, const float* residual,
                          int relu, int use_tc, float* out, float* elapsed_ms) {
    if (M <= 0 || K <= 0 || N <= 0 || K % 4 || N % 4) throw std::runtime_error("M,K,N > 0 and K,N multiples of 4 required");
    float *dA = nullptr, *dW = nullptr, *dB = nullptr, *dR = nullptr, *dO = nullptr, *dWtc = nullptr;
    int* dn = nullptr;
    auto cleanup = [&] { cudaFree(dA); cudaFree(dW); cudaFree(dB); cudaFree(dR); cudaFree(dO); cudaFree(dWtc); cudaFree(dn); };
    try {
        RCUDA_OK(cudaMalloc(&dA, sizeof(float) * (size_t)M * K));
        RCUDA_OK(cudaMalloc(&dW, sizeof(float) * (size_t)K * N));
        RCUDA_OK(cudaMalloc(&dB, sizeof(float) * N));
        RCUDA_OK(cudaMalloc(&dO, sizeof(float) * (size_t)M * N));
        RCUDA_OK(cudaMalloc(&dn, sizeof(int)));
        RCUDA_OK(cudaMemcpy(dA, A, sizeof(float) * (size_t)M * K, cudaMemcpyHostToDevice));
        RCUDA_OK(cudaMemcpy(dW, W, sizeof(float) * (size_t)K * N, cudaMemcpyHostToDevice));
        RCUDA_OK(cudaMemcpy(dB, bias, sizeof(float) * N, cudaMemcpyHostToDevice));
        if (residual) {
            RCUDA_OK(cudaMalloc(&dR, sizeof(float) * (size_t)M * N));
            RCUDA_OK(cudaMemcpy(dR, residual, sizeof(float) * (size_t)M * N, cudaMemcpyHostToDevice));
        }
        const int one = 1;
        RCUDA_OK(cudaMemcpy(dn, &one, sizeof(int), cudaMemcpyHostToDevice));
        ReidModel fake;
        fake.use_tc = use_tc != 0;
        PwArgs p{};
        p.in = dA; p.w = dW; p.bias = dB; p.residual = dR; p.out = dO; p.K = K; p.N = N; p.HW = M; p.relu = relu;
        if (use_tc) {
            if (M % tc::TILE_M) throw std::runtime_error("tensor-core path needs M % 128 == 0");
            const int Kpad = (K + 7) / 8 * 8, Npad = (N + 15) / 16 * 16;
            if (Npad > 256 || tc::smem_bytes(Kpad, Npad) > 200 * 1024) throw std::runtime_error("shape exceeds the tcgen05 kernel's shared-memory budget");
            std::vector<float> packed(2 * (size_t)Npad * Kpad);
            tc::pack_weights(W, K, N, Kpad, Npad, packed.data());
            RCUDA_OK(cudaMalloc(&dWtc, sizeof(float) * packed.size()));
            RCUDA_OK(cudaMemcpy(dWtc, packed.data(), sizeof(float) * packed.size(), cudaMemcpyHostToDevice));
            p.w_tc = dWtc; p.Kpad = Kpad; p.Npad = Npad;
        }
        cudaEvent_t e0, e1;
        RCUDA_OK(cudaEventCreate(&e0));
        RCUDA_OK(cudaEventCreate(&e1));
        Launcher L{&fake, dn, 0, 1, 1, nullptr};
        L.pointwise(p);  // warm-up
        RCUDA_OK(cudaDeviceSynchronize());
        RCUDA_OK(cudaEventRecord(e0));
        for (int r = 0; r < 10; ++r) L.pointwise(p);
        RCUDA_OK(cudaEventRecord(e1));
        RCUDA_OK(cudaDeviceSynchronize());
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (elapsed_ms) *elapsed_ms = ms / 10.f;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        RCUDA_OK(cudaMemcpy(out, dO, sizeof(float) * (size_t)M * N, cudaMemcpyDeviceToHost));
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}

}  // namespace bmb
