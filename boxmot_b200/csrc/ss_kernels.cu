// ss_kernels.cu -- the StrongSORT kernels around ss_core.cuh and their launch sequence.
//
//   k_ss_prepare     unit-norm copies of the detection appearance rows                (one warp per detection)
//   k_ss_appearance  nearest-neighbour cosine distance of every confirmed track's sample gallery to every
//                    detection (sort/linear_assignment.py:266-283,335-344): per track a (gallery x F) by (F x dets)
//                    product reduced with min over the gallery -- the GEMM-shaped part of StrongSORT, tiled over a
//                    wide grid (track, 64-detection tile, stream), float32 FMA like the reference's sgemm
//   k_ss_frame       ss_frame(): one CTA per stream, the assignment solver's state in shared memory
//   k_ss_features    appearance EMA / birth copies and the gallery append              (one warp per track)
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "engine.h"
#include "ss_core.cuh"
#include "tracker_layout.h"

namespace bmb {

#define CUDA_OK(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
    } while (0)

__global__ void __launch_bounds__(256) k_ss_prepare(const SsCfg cfg, SsStream* streams) {
    SsStream s = streams[blockIdx.y];
    const int D = min(*s.n_dets, cfg.cap_dets);
    const int d = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (d >= D) return;
    ss_prepare_row(cfg, s, d);
}

// ---- gallery x detections, min over the gallery ------------------------------------------------------------------
constexpr int APP_BM = 128, APP_BN = 64, APP_BK = 16;

__global__ void __launch_bounds__(256) k_ss_appearance(const SsCfg cfg, SsStream* streams) {
    const SsStream& s = streams[blockIdx.z];
    const int k = blockIdx.y;
    if (k >= s.scalars[SC_N_ACTIVE]) return;
    const int t = s.tracks[k];
    if (s.state[t] != SS_CONFIRMED) return;
    const int D = min(*s.n_dets, cfg.cap_dets);
    const int d0 = blockIdx.x * APP_BN;
    if (d0 >= D) return;
    const int F = cfg.feat_dim, n = s.gal_n[t];
    __shared__ __align__(16) float As[APP_BK][APP_BM + 4];
    __shared__ __align__(16) float Bs[APP_BK][APP_BN + 4];
    __shared__ float red[16][APP_BN];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const float* gal = s.gal + (size_t)t * cfg.budget * F;
    float best[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
    for (int m0 = 0; m0 < n; m0 += APP_BM) {
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < F; k0 += APP_BK) {
            // A tile: 128 gallery rows x 16, B tile: 64 detections x 16 (float4 along K, stored K-major)
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) {
                const int idx = tid + rep * 256, row = idx >> 2, kq = (idx & 3) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m0 + row < n && k0 + kq < F)
                    v = *reinterpret_cast<const float4*>(gal + (size_t)(m0 + row) * F + k0 + kq);
                As[kq + 0][row] = v.x; As[kq + 1][row] = v.y; As[kq + 2][row] = v.z; As[kq + 3][row] = v.w;
            }
            {
                const int row = tid >> 2, kq = (tid & 3) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (d0 + row < D && k0 + kq < F)
                    v = *reinterpret_cast<const float4*>(s.dfeatn + (size_t)(d0 + row) * F + k0 + kq);
                Bs[kq + 0][row] = v.x; Bs[kq + 1][row] = v.y; Bs[kq + 2][row] = v.z; Bs[kq + 3][row] = v.w;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < APP_BK; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (m0 + ty * 8 + i >= n) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float dist = 1.0f - acc[i][j];
                best[j] = dist < best[j] ? dist : best[j];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) red[ty][tx * 4 + j] = best[j];
    __syncthreads();
    if (tid < APP_BN && d0 + tid < D) {
        float m = red[0][tid];
#pragma unroll
        for (int r = 1; r < 16; ++r) m = red[r][tid] < m ? red[r][tid] : m;
        s.appc[(size_t)t * cfg.cap_dets + d0 + tid] = m;
    }
}

// shared-memory residency of the assignment solver's per-row / per-column state
__host__ __device__ inline size_t lsa_smem_bytes(int MX) {
    return (size_t)MX * (3 * sizeof(double) + 6 * sizeof(int)) + 16;
}

__global__ void __launch_bounds__(256) k_ss_frame(const SsCfg cfg, SsStream* streams, int lsa_in_smem) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    SsStream s = streams[blockIdx.x];
    if (lsa_in_smem) {
        const int MX = cfg.cap_tracks > cfg.cap_dets ? cfg.cap_tracks : cfg.cap_dets;
        double* pd = reinterpret_cast<double*>(dyn_smem);
        s.lsa_u = pd; pd += MX;
        s.lsa_v = pd; pd += MX;
        s.lsa_spc = pd; pd += MX;
        int* pi = reinterpret_cast<int*>(pd);
        s.lsa_path = pi; pi += MX;
        s.lsa_row4col = pi; pi += MX;
        s.lsa_col4row = pi; pi += MX;
        s.lsa_rem = pi; pi += MX;
        s.lsa_sr = pi; pi += MX;
        s.lsa_sc = pi;
    }
    ss_frame(cfg, s);
}

__global__ void __launch_bounds__(256) k_ss_features(const SsCfg cfg, SsStream* streams) {
    SsStream s = streams[blockIdx.y];
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= s.scalars[SC_N_ACTIVE]) return;
    ss_features_pos(cfg, s, k);
}

// crop list for on-device ReID: every detection with conf >= min_conf (strongsort.py:74-91), ordered
__global__ void k_ss_build_crops(const SsCfg cfg, SsStream* streams, int n_streams, CropDesc* crops, int* n_crops,
                                 int* hint) {
    if (threadIdx.x >= 32 || blockIdx.x != 0) return;
    const int lane = threadIdx.x;
    int n = 0;
    for (int sidx = 0; sidx < n_streams; ++sidx) {
        const SsStream& s = streams[sidx];
        const int D = min(*s.n_dets, cfg.cap_dets);
        for (int d0 = 0; d0 < D; d0 += 32) {
            const int d = d0 + lane;
            const bool p = d < D && (double)s.dets[d * 6 + 4] >= cfg.min_conf;
            const unsigned m = __ballot_sync(0xffffffffu, p);
            if (p) {
                const float* r = s.dets + d * 6;
                CropDesc c;
                c.x1 = r[0]; c.y1 = r[1]; c.x2 = r[2]; c.y2 = r[3];
                c.image = sidx;
                c.out_row = sidx * cfg.cap_dets + d;
                crops[n + __popc(m & ((1u << lane) - 1u))] = c;
            }
            n += __popc(m);
        }
    }
    if (lane == 0) { *n_crops = n; if (hint) *hint = n; }
}

void ss_build_crops(const SsCfg& cfg, SsStream* d_streams, int S, CropDesc* crops, int* n_crops, int* hint,
                    cudaStream_t stream) {
    k_ss_build_crops<<<1, 32, 0, stream>>>(cfg, d_streams, S, crops, n_crops, hint);
}

int ss_enqueue_frame(const SsCfg& cfg, SsStream* d_streams, int S, cudaStream_t stream) {
    const int CT = cfg.cap_tracks, CD = cfg.cap_dets;
    k_ss_prepare<<<dim3((CD + 7) / 8, S), 256, 0, stream>>>(cfg, d_streams);
    k_ss_appearance<<<dim3((CD + APP_BN - 1) / APP_BN, CT, S), 256, 0, stream>>>(cfg, d_streams);
    const int MX = CT > CD ? CT : CD;
    const size_t lb = lsa_smem_bytes(MX);
    const bool in_smem = lb <= 160 * 1024;
    if (in_smem && lb > 48 * 1024)
        CUDA_OK(cudaFuncSetAttribute(k_ss_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lb));
    k_ss_frame<<<S, 256, in_smem ? lb : 0, stream>>>(cfg, d_streams, in_smem ? 1 : 0);
    k_ss_features<<<dim3((CT + 7) / 8, S), 256, 0, stream>>>(cfg, d_streams);
    CUDA_OK(cudaGetLastError());
    return 4;
}

// ---- standalone entries for parity tests ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_lsa_only(SsStream* streams, int nr, int nc) {
    SsStream s = streams[blockIdx.x];
    lsa_solve(s, s.cost, nr, nc, nc);
}

// scipy.optimize.linear_sum_assignment on a host (R, C) float64 matrix; returns min(R, C) pairs, rows ascending
int standalone_lsa(const double* cost, int R, int C, int* row_ind, int* col_ind) {
    if (R < 0 || C < 0) throw std::runtime_error("negative shape");
    if (R == 0 || C == 0) return 0;
    SsCfg c{};
    c.cap_tracks = R < 8 ? 8 : R;
    c.cap_dets = C;
    c.feat_dim = 4;
    c.budget = 1;
    const size_t bytes = carve_ss(c, nullptr, nullptr, nullptr);
    uint8_t* mem = nullptr;
    SsStream hs, *ds = nullptr;
    CUDA_OK(cudaMalloc(&mem, bytes));
    CUDA_OK(cudaMemset(mem, 0, bytes));
    carve_ss(c, mem, &hs, nullptr);
    const bool tr = C < R;
    const int nr = tr ? C : R, nc = tr ? R : C;
    std::vector<double> m((size_t)R * C);
    for (int r = 0; r < R; ++r)
        for (int q = 0; q < C; ++q) m[tr ? (size_t)q * nc + r : (size_t)r * nc + q] = cost[(size_t)r * C + q];
    CUDA_OK(cudaMalloc(&ds, sizeof(SsStream)));
    CUDA_OK(cudaMemcpy(ds, &hs, sizeof(SsStream), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(hs.cost, m.data(), sizeof(double) * m.size(), cudaMemcpyHostToDevice));
    k_lsa_only<<<1, 256>>>(ds, nr, nc);
    std::vector<int> c4r(nr), r4c(nc), sc(SC_COUNT);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(c4r.data(), hs.lsa_col4row, sizeof(int) * nr, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(r4c.data(), hs.lsa_row4col, sizeof(int) * nc, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(sc.data(), hs.scalars, sizeof(int) * SC_COUNT, cudaMemcpyDeviceToHost);
    cudaFree(mem);
    cudaFree(ds);
    CUDA_OK(e);
    if (sc[SC_ERROR] == ERR_LSA_INFEASIBLE) throw std::runtime_error("cost matrix is infeasible");
    int n = 0;
    if (!tr) {
        for (int i = 0; i < nr; ++i) { row_ind[n] = i; col_ind[n] = c4r[i]; ++n; }
    } else {
        for (int r = 0; r < R; ++r)
            if (r4c[r] >= 0) { row_ind[n] = r; col_ind[n] = r4c[r]; ++n; }
    }
    return n;
}

}  // namespace bmb
