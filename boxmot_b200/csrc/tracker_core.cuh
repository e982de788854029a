// tracker_core.cuh -- the per-stream, per-frame track update for the STrack family (ByteTrack, BoT-SORT).
//
// One CTA owns one video stream and runs the whole frame: detection split, Kalman predict over the pool,
// IoU / appearance cost build, three linear-assignment rounds, Kalman update, EMA appearance update, the
// track state machine and output row assembly.  All track state is structure-of-arrays in HBM and never
// leaves the device; a frame costs one launch for every stream resident on the GPU.
//
// What it replaces in the reference (relative to /root/reference/boxmot):
//   trackers/bbox/bytetrack/bytetrack.py:259-447    ByteTrack._update_impl (+ joint/sub/remove_duplicate)
//   trackers/bbox/botsort/botsort.py:177-500        BotSort._update_impl
//   trackers/bbox/botsort/botsort_track.py:16-115,232-282   STrack feature EMA, class vote, update paths
//   motion/kalman_filters/base.py:234-355 + xyah.py / xywh.py  initiate / multi_predict / project / update
//   trackers/association/matching.py:28-147         iou_distance, embedding gates, fuse_score, lapjv
//   trackers/association/iou.py:134-150             iou_batch
//
// The same source compiles for the host (BMB_HOSTSIM, one "thread") so the control flow can be checked
// against the golden vectors in a container without a GPU; that build lives under tests/ only and is never
// linked into the product library.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__) && !defined(BMB_HOSTSIM)
#define BMB_FN __device__ __forceinline__
#define BMB_TID ((int)threadIdx.x)
#define BMB_NT ((int)blockDim.x)
#define BMB_SYNC() __syncthreads()
#define BMB_LANE ((int)(threadIdx.x & 31))
#define BMB_NL 32
#define BMB_WARP ((int)(threadIdx.x >> 5))
#define BMB_NW ((int)(blockDim.x >> 5))
#define BMB_SYNCWARP() __syncwarp()
#define BMB_BALLOT(p) __ballot_sync(0xffffffffu, (p))
#define BMB_POPC(x) __popc(x)
#define BMB_SHFL_DOWN_D(v, o) __shfl_down_sync(0xffffffffu, (v), (o))
#define BMB_SHFL_DOWN_I(v, o) __shfl_down_sync(0xffffffffu, (v), (o))
#define BMB_SHFL_D(v, l) __shfl_sync(0xffffffffu, (v), (l))
#define BMB_SHFL_I(v, l) __shfl_sync(0xffffffffu, (v), (l))
#define BMB_SHFL_XOR_F(v, o) __shfl_xor_sync(0xffffffffu, (v), (o))
#define BMB_DEVICE 1
#define BMB_CLOCK() clock64()
#define BMB_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define BMB_ATOMIC_MIN(p, v) atomicMin((p), (v))
#else
#define BMB_ATOMIC_MAX(p, v) do { if ((v) > *(p)) *(p) = (v); } while (0)
#define BMB_ATOMIC_MIN(p, v) do { if ((v) < *(p)) *(p) = (v); } while (0)
#define BMB_CLOCK() 0ll
#define BMB_FN static inline
#define BMB_TID 0
#define BMB_NT 1
#define BMB_SYNC() ((void)0)
#define BMB_LANE 0
#define BMB_NL 1
#define BMB_WARP 0
#define BMB_NW 1
#define BMB_SYNCWARP() ((void)0)
#define BMB_BALLOT(p) ((p) ? 1u : 0u)
#define BMB_POPC(x) ((int)((x) & 1u))
#define BMB_SHFL_DOWN_D(v, o) (v)
#define BMB_SHFL_DOWN_I(v, o) (v)
#define BMB_SHFL_D(v, l) (v)
#define BMB_SHFL_I(v, l) (v)
#define BMB_SHFL_XOR_F(v, o) (v)
#define BMB_DEVICE 0
#endif

namespace bmb {

enum { ST_NEW = 0, ST_TRACKED = 1, ST_LOST = 2, ST_REMOVED = 3 };
enum { KIND_XYAH = 0, KIND_XYWH = 1 };
enum { HIST_CAP = 8 };  // distinct classes remembered per track by the BoT-SORT class vote
enum {
    SC_N_ACTIVE = 0, SC_N_LOST, SC_FRAME, SC_NEXT_ID, SC_RING_HEAD, SC_RING_COUNT, SC_ERROR, SC_N_OUT,
    SC_LAP_STEPS, SC_N_EMA, SC_COUNT = 16
};
// mailbox slots (first MB_COUNT ints of free_l): counts produced by one thread, read by the whole CTA after a sync
enum {
    MB_N_FIRST = 0, MB_N_SECOND, MB_N_UNC, MB_N_POOL0, MB_N_POOL, MB_N_ACT, MB_N_REFIND, MB_N_EMA, MB_N_RT,
    MB_N_REST, MB_N_LOSTNOW, MB_N_REMNOW, MB_N_BIRTH, MB_M1, MB_M2, MB_M3, MB_Q1, MB_Q2, MB_COUNT = 32
};
enum { ERR_NONE = 0, ERR_TRACK_CAPACITY = 1, ERR_CLS_HIST = 2, ERR_DET_CAPACITY = 3, ERR_LSA_INFEASIBLE = 4 };

struct TrkCfg {
    int kind;             // KIND_XYAH (ByteTrack) / KIND_XYWH (BoT-SORT)
    int with_reid;        // appearance gates active
    int fuse_first;       // fuse_score in the first association round
    int proximity_mask;   // BoT-SORT proximity gate (rounds 1 and 3)
    int max_time_lost;
    int removed_cap;      // 0: unbounded removed set (ByteTrack); >0: deque(maxlen) of ids (BoT-SORT)
    int feat_dim;
    int cap_tracks;       // CT
    int cap_dets;         // CD
    int vote_cls;         // BoT-SORT class vote
    double high_thresh, low_thresh;
    float new_thresh_f32;  // np.float32 conf < python float compares in float32 (NEP 50)
    double match1, match2, match3;
    double proximity, appearance, unc_emb_scale;
};

struct TrkStream {
    // persistent track table (slot indexed)
    double* mean;  // [CT][8]
    double* cov;   // [CT][64]
    int* state;
    int* activated;
    int* id;
    int* frame_id;
    int* start_frame;
    int* tracklet_len;
    float* conf;
    float* cls;
    float* det_ind;
    float* smooth;     // [CT][F]
    float* hist_cls;   // [CT][HIST_CAP]
    float* hist_sum;   // [CT][HIST_CAP]
    int* hist_n;       // [CT]
    int* in_removed;   // [CT]   unbounded removed set as a slot flag
    int* removed_ring; // [removed_cap] ids
    int* active;       // [CT] ordered slot list
    int* lost;         // [CT] ordered slot list
    int* scalars;      // [SC_COUNT]
    long long* timers; // [16] accumulated clock64() ticks per phase (diagnostics)
    // per-frame inputs
    const float* dets;  // [CD][6]
    const int* n_dets;  // [1]
    const double* warp; // [8]: warp[0..5] = 2x3 camera-motion matrix (row major), warp[6] != 0 when one is pending
    // per-frame scratch
    float* dxywh;   // [CD][4]
    float* dmeas;   // [CD][4]
    float* dxyxy;   // [CD][4]
    float* dfeat;   // [CD][F]  detection appearance after the reference's two in-place normalisations
    double* embd;   // [CT][CD] max(0, cosine distance(smooth[slot], dfeat[det])) by the wide kernel
    double* cost;   // [CT][CD]
    double* txyxy;  // [CT][4]
    int* first;     // [CD]
    int* second;    // [CD]
    int* rest;      // [CD]
    int* pool;      // [CT]
    int* unconf;    // [CT]
    int* rtracked;  // [CT]
    int* act_l;     // [CT+CD]
    int* refind_l;  // [CT]
    int* lostnow_l; // [CT]
    int* remnow_l;  // [CT]
    int* tmp_a;     // [CT]
    int* tmp_b;     // [CT]
    int* mark;      // [CT]
    int* free_l;    // [MB_COUNT + CT]  mailbox + free-slot list
    int* ema_slot;  // [CD] matched (slot, detection) pairs whose appearance EMA is applied after the frame
    int* ema_det;   // [CD]
    // linear assignment scratch
    int* lap_x;       // [CT]
    int* lap_y;       // [CD]
    double* lap_u;    // [CT]
    double* lap_v;    // [CD]
    double* lap_spc;  // [CD]
    int* lap_path;    // [CD]
    int* lap_insc;    // [CD]
    int* lap_tl;      // [CD]
    int* lap_sc;      // [CD]
    int* csr_ptr;     // [CT+1]
    int* csr_col;     // [CT*CD]
    // output
    float* out;  // [CD][8]
};

// ---------------------------------------------------------------------------------------------------
// small geometry helpers (float32 detection geometry exactly as numpy evaluates it)
// ---------------------------------------------------------------------------------------------------
BMB_FN void det_geometry(const TrkCfg& c, const float* d, float* xywh, float* meas, float* xyxy) {
    // xyxy2xywh on a float32 row (trackers/common/geometry.py:10-24)
    float cx = (d[0] + d[2]) / 2.0f;
    float cy = (d[1] + d[3]) / 2.0f;
    float w = d[2] - d[0];
    float h = d[3] - d[1];
    xywh[0] = cx; xywh[1] = cy; xywh[2] = w; xywh[3] = h;
    if (c.kind == KIND_XYAH) {
        // xywh2tlwh then tlwh2xyah (bytetrack.py:37-40)
        float t0 = cx - w / 2.0f;
        float t1 = cy - h / 2.0f;
        meas[0] = t0 + (w / 2.0f);
        meas[1] = t1 + (h / 2.0f);
        meas[2] = w / h;
        meas[3] = h;
    } else {
        meas[0] = cx; meas[1] = cy; meas[2] = w; meas[3] = h;
    }
    // STrack.xyxy of a not-yet-activated detection: xywh2xyxy(self.xywh) in float32
    xyxy[0] = cx - w / 2.0f;
    xyxy[1] = cy - h / 2.0f;
    xyxy[2] = cx + w / 2.0f;
    xyxy[3] = cy + h / 2.0f;
}

BMB_FN void track_xyxy(const TrkCfg& c, const double* m, double* o) {
    double w = m[2], h = m[3];
    if (c.kind == KIND_XYAH) w = w * h;
    o[0] = m[0] - w / 2.0;
    o[1] = m[1] - h / 2.0;
    o[2] = m[0] + w / 2.0;
    o[3] = m[1] + h / 2.0;
}

// 1 - IoU between a float64 track box and a float32 detection box; numpy computes the detection area in
// float32 before promotion (association/iou.py:134-150).
BMB_FN double iou_dist_td(const double* a, const float* b) {
    double b0 = (double)b[0], b1 = (double)b[1], b2 = (double)b[2], b3 = (double)b[3];
    double xx1 = a[0] > b0 ? a[0] : b0;
    double yy1 = a[1] > b1 ? a[1] : b1;
    double xx2 = a[2] < b2 ? a[2] : b2;
    double yy2 = a[3] < b3 ? a[3] : b3;
    double w = xx2 - xx1; w = w > 0.0 ? w : 0.0;
    double h = yy2 - yy1; h = h > 0.0 ? h : 0.0;
    double wh = w * h;
    double area_a = (a[2] - a[0]) * (a[3] - a[1]);
    float area_b = (b[2] - b[0]) * (b[3] - b[1]);
    double o = wh / (area_a + (double)area_b - wh);
    return 1.0 - o;
}

BMB_FN double iou_dist_tt(const double* a, const double* b) {
    double xx1 = a[0] > b[0] ? a[0] : b[0];
    double yy1 = a[1] > b[1] ? a[1] : b[1];
    double xx2 = a[2] < b[2] ? a[2] : b[2];
    double yy2 = a[3] < b[3] ? a[3] : b[3];
    double w = xx2 - xx1; w = w > 0.0 ? w : 0.0;
    double h = yy2 - yy1; h = h > 0.0 ? h : 0.0;
    double wh = w * h;
    double o = wh / ((a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - wh);
    return 1.0 - o;
}

// ---------------------------------------------------------------------------------------------------
// Kalman filter (float64).  W_POS = 1/20, W_VEL = 1/160 as python evaluates them.
// ---------------------------------------------------------------------------------------------------
#define BMB_W_POS 0.05
#define BMB_W_VEL 0.00625

BMB_FN void kf_sizes(const TrkCfg& c, const double* v, double* s) {
    if (c.kind == KIND_XYAH) { s[0] = v[3]; s[1] = v[3]; s[2] = v[3]; s[3] = v[3]; }
    else { s[0] = v[2]; s[1] = v[3]; s[2] = v[2]; s[3] = v[3]; }
}

// base.py:234-244 + xyah.py:22-37 / xywh.py:22-36; measurement arrives as float32 and is widened first.
BMB_FN void kf_initiate(const TrkCfg& c, const float* meas32, double* mean, double* cov) {
    double m[4] = {(double)meas32[0], (double)meas32[1], (double)meas32[2], (double)meas32[3]};
    double s[4];
    kf_sizes(c, m, s);
    double sd[8];
    const double two_wp = 2 * BMB_W_POS, ten_wv = 10 * BMB_W_VEL;
    for (int i = 0; i < 4; ++i) { sd[i] = two_wp * s[i]; sd[4 + i] = ten_wv * s[i]; }
    if (c.kind == KIND_XYAH) { sd[2] = 1e-2; sd[6] = 1e-5; }
    for (int i = 0; i < 64; ++i) cov[i] = 0.0;
    for (int i = 0; i < 8; ++i) cov[i * 8 + i] = sd[i] * sd[i];
    for (int i = 0; i < 4; ++i) { mean[i] = m[i]; mean[4 + i] = 0.0; }
    if (mean[2] < 1e-4) mean[2] = 1e-4;
    if (mean[3] < 1e-4) mean[3] = 1e-4;
}

// base.py:311-327 (+ STrack.multi_predict velocity zeroing for non-Tracked tracks).  In place.
BMB_FN void kf_predict(const TrkCfg& c, int tracked, double* mean, double* cov) {
    if (!tracked) {
        if (c.kind == KIND_XYAH) mean[7] = 0.0;
        else { mean[6] = 0.0; mean[7] = 0.0; }
    }
    double s[4], q[8];
    kf_sizes(c, mean, s);   // process noise from the PRIOR mean
    for (int i = 0; i < 4; ++i) {
        double sp = BMB_W_POS * s[i], sv = BMB_W_VEL * s[i];
        q[i] = sp * sp; q[4 + i] = sv * sv;
    }
    if (c.kind == KIND_XYAH) { q[2] = 1e-2 * 1e-2; q[6] = 1e-5 * 1e-5; }
    for (int i = 0; i < 4; ++i) mean[i] = mean[i] + mean[4 + i];
    // left = F P ; P' = left F^T + Q, with the reference's summation order
    double L[64];
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) L[i * 8 + j] = (i < 4) ? (cov[i * 8 + j] + cov[(i + 4) * 8 + j]) : cov[i * 8 + j];
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            double v = (j < 4) ? (L[i * 8 + j] + L[i * 8 + j + 4]) : L[i * 8 + j];
            if (i == j) v = v + q[i];
            cov[i * 8 + j] = v;
        }
    if (mean[2] < 1e-4) mean[2] = 1e-4;
    if (mean[3] < 1e-4) mean[3] = 1e-4;
}

// base.py:286-309,329-355: S = HPH^T + R, Cholesky, K = P H^T S^-1, x += K y, P -= K S K^T (non-Joseph).
BMB_FN void kf_update(const TrkCfg& c, const float* meas32, double* mean, double* cov) {
    double s[4], S[16], Lc[16];
    kf_sizes(c, mean, s);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) S[i * 4 + j] = cov[i * 8 + j];
    for (int i = 0; i < 4; ++i) {
        double sd = BMB_W_POS * s[i];
        if (c.kind == KIND_XYAH && i == 2) sd = 1e-1;
        sd = 1.0 * sd;  // (1 - confidence) with confidence = 0
        S[i * 4 + i] = S[i * 4 + i] + sd * sd;
    }
    // lower Cholesky S = Lc Lc^T
    for (int i = 0; i < 16; ++i) Lc[i] = 0.0;
    for (int j = 0; j < 4; ++j) {
        double d = S[j * 4 + j];
        for (int k = 0; k < j; ++k) d -= Lc[j * 4 + k] * Lc[j * 4 + k];
        d = sqrt(d);
        Lc[j * 4 + j] = d;
        for (int i = j + 1; i < 4; ++i) {
            double v = S[i * 4 + j];
            for (int k = 0; k < j; ++k) v -= Lc[i * 4 + k] * Lc[j * 4 + k];
            Lc[i * 4 + j] = v / d;
        }
    }
    // K^T (4x8) = S^-1 (P H^T)^T ; column r of (P H^T)^T is row r of P restricted to the first 4 columns
    double KT[32];
    for (int r = 0; r < 8; ++r) {
        double y[4];
        for (int i = 0; i < 4; ++i) {
            double v = cov[r * 8 + i];
            for (int k = 0; k < i; ++k) v -= Lc[i * 4 + k] * y[k];
            y[i] = v / Lc[i * 4 + i];
        }
        for (int i = 3; i >= 0; --i) {
            double v = y[i];
            for (int k = i + 1; k < 4; ++k) v -= Lc[k * 4 + i] * KT[k * 8 + r];
            KT[i * 8 + r] = v / Lc[i * 4 + i];
        }
    }
    double inn[4];
    for (int i = 0; i < 4; ++i) inn[i] = (double)meas32[i] - mean[i];
    for (int r = 0; r < 8; ++r) {
        double acc = 0.0;
        for (int i = 0; i < 4; ++i) acc += inn[i] * KT[i * 8 + r];
        mean[r] = mean[r] + acc;
    }
    // M = S K^T (4x8); P -= K M
    double M[32];
    for (int i = 0; i < 4; ++i)
        for (int r = 0; r < 8; ++r) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += S[i * 4 + k] * KT[k * 8 + r];
            M[i * 8 + r] = acc;
        }
    for (int a = 0; a < 8; ++a)
        for (int b = 0; b < 8; ++b) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += KT[k * 8 + a] * M[k * 8 + b];
            cov[a * 8 + b] = cov[a * 8 + b] - acc;
        }
    if (mean[2] < 1e-4) mean[2] = 1e-4;
    if (mean[3] < 1e-4) mean[3] = 1e-4;
}

// STrack.multi_gmc (botsort_track.py:117-132): x <- kron(I4, R) x, x[:2] += t, P <- R8 P R8^T for a given 2x3 warp.
// The warp ESTIMATION (motion/cmc/*) is outside this path; applying a supplied warp is part of the Kalman kernel.
BMB_FN void kf_apply_warp(const double* H, double* mean, double* cov) {
    const double r00 = H[0], r01 = H[1], tx = H[2], r10 = H[3], r11 = H[4], ty = H[5];
    for (int b = 0; b < 4; ++b) {
        const double x = mean[2 * b], y = mean[2 * b + 1];
        mean[2 * b] = r00 * x + r01 * y;
        mean[2 * b + 1] = r10 * x + r11 * y;
    }
    mean[0] += tx;
    mean[1] += ty;
    double tmp[64];
    for (int b = 0; b < 4; ++b)       // tmp = R8 P
        for (int j = 0; j < 8; ++j) {
            const double p0 = cov[(2 * b) * 8 + j], p1 = cov[(2 * b + 1) * 8 + j];
            tmp[(2 * b) * 8 + j] = r00 * p0 + r01 * p1;
            tmp[(2 * b + 1) * 8 + j] = r10 * p0 + r11 * p1;
        }
    for (int i = 0; i < 8; ++i)       // P' = tmp R8^T
        for (int b = 0; b < 4; ++b) {
            const double q0 = tmp[i * 8 + 2 * b], q1 = tmp[i * 8 + 2 * b + 1];
            cov[i * 8 + 2 * b] = q0 * r00 + q1 * r01;
            cov[i * 8 + 2 * b + 1] = q0 * r10 + q1 * r11;
        }
}

// ---------------------------------------------------------------------------------------------------
// appearance: the float32 normalisations and EMA of botsort_track.py:58-67, one warp per vector
// ---------------------------------------------------------------------------------------------------
BMB_FN float warp_sum_f(float v) {
#if BMB_DEVICE
    for (int o = 16; o > 0; o >>= 1) v += BMB_SHFL_XOR_F(v, o);
#endif
    return v;
}

BMB_FN float vec_norm_f32(const float* x, int n) {
    float acc = 0.0f;
    for (int k = BMB_LANE; k < n; k += BMB_NL) acc += x[k] * x[k];
    acc = warp_sum_f(acc);
    return sqrtf(acc);
}

// dst = src / ||src|| twice (update_features on a fresh STrack: feat /= norm; smooth_feat is the same array
// and is normalised again in place).  Warp-cooperative.
BMB_FN void feat_prepare(const float* src, float* dst, int n) {
    float nr = vec_norm_f32(src, n);
    for (int k = BMB_LANE; k < n; k += BMB_NL) dst[k] = src[k] / nr;
    BMB_SYNCWARP();
    nr = vec_norm_f32(dst, n);
    BMB_SYNCWARP();
    for (int k = BMB_LANE; k < n; k += BMB_NL) dst[k] = dst[k] / nr;
    BMB_SYNCWARP();
}

// track.update_features(det.curr_feat): det feature normalised a third time in place, then the EMA.
BMB_FN void feat_ema(float* smooth, float* dfeat, int n) {
    float nr = vec_norm_f32(dfeat, n);
    BMB_SYNCWARP();
    for (int k = BMB_LANE; k < n; k += BMB_NL) {
        float f = dfeat[k] / nr;
        dfeat[k] = f;
        float a = 0.9f * smooth[k];
        float b = 0.1f * f;
        smooth[k] = a + b;
    }
    BMB_SYNCWARP();
    nr = vec_norm_f32(smooth, n);
    BMB_SYNCWARP();
    for (int k = BMB_LANE; k < n; k += BMB_NL) smooth[k] = smooth[k] / nr;
    BMB_SYNCWARP();
}

// botsort_track.py:69-82
BMB_FN void vote_cls(const TrkCfg& c, TrkStream& s, int slot, float cls, float conf) {
    float* hc = s.hist_cls + slot * HIST_CAP;
    float* hs = s.hist_sum + slot * HIST_CAP;
    int n = s.hist_n[slot];
    float best = 0.0f;
    int seen = 0;
    for (int k = 0; k < n; ++k) {
        if (cls == hc[k]) { hs[k] = hs[k] + conf; seen = 1; }
        if (hs[k] > best) { best = hs[k]; s.cls[slot] = hc[k]; }
    }
    if (!seen) {
        if (n < HIST_CAP) { hc[n] = cls; hs[n] = conf; s.hist_n[slot] = n + 1; }
        else s.scalars[SC_ERROR] = ERR_CLS_HIST;
        s.cls[slot] = cls;
    }
}

// max(0, cosine distance) of two float32 vectors evaluated in float64, as scipy's cdist(..., 'cosine') does
// after widening (matching.py:85-107): 1 - u.v / (|u| |v|), cosine clipped to [-1, 1].
BMB_FN double cosine_cost_f64(const float* a, const float* b, int n) {
    double dot = 0.0, na = 0.0, nb = 0.0;
    for (int k = 0; k < n; ++k) {
        double x = (double)a[k], y = (double)b[k];
        dot += x * y; na += x * x; nb += y * y;
    }
    double cs = dot / (sqrt(na) * sqrt(nb));
    if (fabs(cs) > 1.0) cs = cs > 0 ? 1.0 : -1.0;
    double d = 1.0 - cs;
    return d > 0.0 ? d : (d != d ? d : 0.0);
}

// ---------------------------------------------------------------------------------------------------
// Linear assignment with lapjv(extend_cost=True, cost_limit=thresh) semantics (matching.py:28-43).
//
// lapjv pads the T x D problem to (T+D)^2 with thresh/2 everywhere outside the real block and 0 in the
// dummy-dummy block: a real pair is worth matching iff it beats leaving both ends unmatched (thresh).  That
// is exactly: every row either takes a real column at c[i][j] or its private "stay unmatched" column at
// thresh.  Entries with c >= thresh can never be part of an optimum, so the graph is sparse (a track overlaps
// few detections).  We solve it exactly with shortest augmenting paths (row by row, dual prices u/v in
// float64) on CSR candidate lists; one warp runs the search, lanes relax candidates and reduce the frontier.
// Unique optimum  =>  same matches as lapjv (ties have measure zero on continuous costs).
// ---------------------------------------------------------------------------------------------------
// Ordered append by ONE warp: for every i in [0,n) with pred(i), in ascending i, out[base++] = val(i).
template <typename Pred, typename Val>
BMB_FN int warp_append(int n, int* out, int base, Pred pred, Val val) {
    for (int i0 = 0; i0 < n; i0 += BMB_NL) {
        const int i = i0 + BMB_LANE;
        const bool p = i < n && pred(i);
        const unsigned m = BMB_BALLOT(p);
        if (p) {
#if BMB_DEVICE
            out[base + __popc(m & ((1u << BMB_LANE) - 1u))] = val(i);
#else
            out[base] = val(i);
#endif
        }
        base += BMB_POPC(m);
    }
    return base;
}

// Candidate lists (CSR) plus a parallel start for the augmenting-path solver: every row takes its cheapest
// candidate as dual price u[i] (v = 0 keeps all reduced costs >= 0) and claims that column; a column claimed by
// several rows goes to the lowest row.  The claimed pairs are tight edges of a feasible dual, i.e. a valid partial
// matching for the exact search below, which then only has to place the rows that lost a claim.
template <typename S>
BMB_FN void lap_build_csr(S& s, int T, int D, int ld, double thresh) {
    for (int j = BMB_TID; j < D; j += BMB_NT) s.lap_tl[j] = 0x7fffffff;  // column claims (lap_tl is free here)
    for (int i = BMB_WARP; i < T; i += BMB_NW) {
        const double* ci = s.cost + (size_t)i * ld;
        int n = 0;
        double best = INFINITY;
        int bj = -1;
        for (int j0 = 0; j0 < D; j0 += BMB_NL) {
            const int j = j0 + BMB_LANE;
            const double v = j < D ? ci[j] : INFINITY;
            const bool cand = j < D && v < thresh;
            n += BMB_POPC(BMB_BALLOT(cand));
            if (cand && v < best) { best = v; bj = j; }
        }
#if BMB_DEVICE
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = BMB_SHFL_DOWN_D(best, o);
            const int oj = BMB_SHFL_DOWN_I(bj, o);
            if (oj >= 0 && (bj < 0 || ov < best || (ov == best && oj < bj))) { best = ov; bj = oj; }
        }
#endif
        if (BMB_LANE == 0) {
            s.lap_x[i] = n;  // temporarily the row count
            s.lap_u[i] = bj >= 0 ? best : 0.0;
            s.tmp_b[i] = bj;
        }
    }
    BMB_SYNC();
    if (BMB_TID == 0) {
        int acc = 0;
        for (int i = 0; i < T; ++i) { s.csr_ptr[i] = acc; acc += s.lap_x[i]; }
        s.csr_ptr[T] = acc;
    }
    for (int i = BMB_TID; i < T; i += BMB_NT) {
        const int bj = s.tmp_b[i];
        if (bj >= 0) {
#if BMB_DEVICE
            atomicMin(&s.lap_tl[bj], i);
#else
            if (i < s.lap_tl[bj]) s.lap_tl[bj] = i;
#endif
        }
    }
    BMB_SYNC();
    for (int i = BMB_WARP; i < T; i += BMB_NW) {
        const double* ci = s.cost + (size_t)i * ld;
        warp_append(D, s.csr_col, s.csr_ptr[i], [&](int j) { return ci[j] < thresh; }, [&](int j) { return j; });
    }
    BMB_SYNC();
}

// Called by the whole CTA; result in lap_x[0..T) (column or -1) and lap_y[0..D) (row or -1).
template <typename S>
BMB_FN void lap_solve(S& s, int T, int D, int ld, double thresh) {
    if (T == 0 || D == 0) {
        for (int i = BMB_TID; i < T; i += BMB_NT) s.lap_x[i] = -1;
        for (int j = BMB_TID; j < D; j += BMB_NT) s.lap_y[j] = -1;
        BMB_SYNC();
        return;
    }
    lap_build_csr(s, T, D, ld, thresh);
    for (int j = BMB_TID; j < D; j += BMB_NT) {
        s.lap_y[j] = -1; s.lap_v[j] = 0.0; s.lap_spc[j] = INFINITY; s.lap_insc[j] = 0;
    }
    BMB_SYNC();
    for (int i = BMB_TID; i < T; i += BMB_NT) {  // uncontested claims become the initial matching
        const int bj = s.tmp_b[i];
        if (bj >= 0 && s.lap_tl[bj] == i) { s.lap_x[i] = bj; s.lap_y[bj] = i; }
        else s.lap_x[i] = -1;
    }
    BMB_SYNC();
    if (BMB_WARP == 0) {
        const int lane = BMB_LANE;
        const double INF = INFINITY;
        int steps = 0;
        for (int cur = 0; cur < T; ++cur) {
            if (s.lap_x[cur] >= 0) continue;                        // placed by the parallel start
            if (s.csr_ptr[cur + 1] == s.csr_ptr[cur]) continue;  // no candidate: stays unmatched
            double minval = 0.0;
            int i = cur;
            int n_tl = 0, n_sc = 0;
            int sink = -1;
            double dmin = INF;   // cheapest "leave row r unmatched" terminal seen so far
            int drow = -1;
            while (true) {
                ++steps;
                double ui = s.lap_u[i];
                double cand = (minval + thresh) - ui;
                if (cand < dmin) { dmin = cand; drow = i; }
                // relax the candidates of row i
                const int e0 = s.csr_ptr[i], e1 = s.csr_ptr[i + 1];
                const double* ci = s.cost + (size_t)i * ld;
                for (int base = e0; base < e1; base += BMB_NL) {
                    int e = base + lane;
                    int fresh = 0;
                    int j = -1;
                    if (e < e1) {
                        j = s.csr_col[e];
                        if (!s.lap_insc[j]) {
                            double r = ((minval + ci[j]) - ui) - s.lap_v[j];
                            double old = s.lap_spc[j];
                            if (r < old) {
                                s.lap_spc[j] = r;
                                s.lap_path[j] = i;
                                fresh = (old == INF) ? 1 : 0;
                            }
                        }
                    }
                    unsigned m = BMB_BALLOT(fresh);
                    if (fresh) {
#if BMB_DEVICE
                        int pos = n_tl + __popc(m & ((1u << lane) - 1u));
#else
                        int pos = n_tl;
#endif
                        s.lap_tl[pos] = j;
                    }
                    n_tl += BMB_POPC(m);
                }
                BMB_SYNCWARP();
                // frontier minimum over touched, unscanned columns
                double best = INF;
                int bj = -1;
                for (int k = lane; k < n_tl; k += BMB_NL) {
                    int j = s.lap_tl[k];
                    if (!s.lap_insc[j]) {
                        double v = s.lap_spc[j];
                        if (v < best || (v == best && j < bj)) { best = v; bj = j; }
                    }
                }
#if BMB_DEVICE
                for (int o = 16; o > 0; o >>= 1) {
                    double ov = BMB_SHFL_DOWN_D(best, o);
                    int oj = BMB_SHFL_DOWN_I(bj, o);
                    if (oj >= 0 && (bj < 0 || ov < best || (ov == best && oj < bj))) { best = ov; bj = oj; }
                }
                best = BMB_SHFL_D(best, 0);
                bj = BMB_SHFL_I(bj, 0);
#endif
                if (bj < 0 || dmin <= best) {  // cheapest way out is to leave `drow` unmatched
                    minval = dmin;
                    sink = -1;
                    break;
                }
                minval = best;
                if (lane == 0) { s.lap_insc[bj] = 1; s.lap_sc[n_sc] = bj; }
                ++n_sc;
                BMB_SYNCWARP();
                if (s.lap_y[bj] < 0) { sink = bj; break; }
                i = s.lap_y[bj];
            }
            // dual update (lanes over the scanned columns)
            for (int k = lane; k < n_sc; k += BMB_NL) {
                int j = s.lap_sc[k];
                double dj = s.lap_spc[j];
                int r = s.lap_y[j];
                if (r >= 0) s.lap_u[r] = s.lap_u[r] + (minval - dj);
                s.lap_v[j] = s.lap_v[j] - (minval - dj);
            }
            if (lane == 0) s.lap_u[cur] = s.lap_u[cur] + minval;
            BMB_SYNCWARP();
            // augment (sequential walk back along the path)
            if (lane == 0) {
                int j;
                bool go = true;
                if (sink >= 0) {
                    j = sink;
                } else if (drow == cur) {
                    go = false;  // cur itself stays unmatched
                    j = -1;
                } else {
                    j = s.lap_x[drow];
                    s.lap_x[drow] = -1;
                }
                while (go) {
                    int r = s.lap_path[j];
                    s.lap_y[j] = r;
                    int prev = s.lap_x[r];
                    s.lap_x[r] = j;
                    j = prev;
                    if (r == cur) go = false;
                }
            }
            // reset the touched columns
            for (int k = lane; k < n_tl; k += BMB_NL) {
                int j = s.lap_tl[k];
                s.lap_spc[j] = INF;
                s.lap_insc[j] = 0;
            }
            BMB_SYNCWARP();
        }
        if (lane == 0) s.scalars[SC_LAP_STEPS] += steps;
    }
    BMB_SYNC();
}

// ---------------------------------------------------------------------------------------------------
// match application: Kalman update + per-slot bookkeeping (thread per match), ordered list appends (warp 0).
// The appearance EMA of matched tracks is only consumed by the NEXT frame's appearance cost, so the pairs are
// recorded here and applied by a wide kernel after this one (one warp per pair across the whole GPU).
// ---------------------------------------------------------------------------------------------------
// rows[i] = slot of LAP row i, cols[j] = detection index of LAP column j.  Counts live in the mailbox `mb`.
BMB_FN void apply_matches(const TrkCfg& c, TrkStream& s, int T, const int* rows, const int* cols, int frame,
                          int force_update_branch, int use_feat, int* mb) {
    for (int i = BMB_TID; i < T; i += BMB_NT) {
        const int j = s.lap_x[i];
        if (j < 0) { s.tmp_b[i] = 0; continue; }
        const int slot = rows[i];
        const int det = cols[j];
        kf_update(c, s.dmeas + det * 4, s.mean + slot * 8, s.cov + slot * 64);
        const float* d = s.dets + det * 6;
        const int tracked = force_update_branch || s.state[slot] == ST_TRACKED;
        if (tracked) {
            s.frame_id[slot] = frame;
            s.tracklet_len[slot] += 1;
        } else {
            s.tracklet_len[slot] = 0;
            s.frame_id[slot] = frame;
        }
        s.state[slot] = ST_TRACKED;
        s.activated[slot] = 1;
        s.conf[slot] = d[4];
        s.cls[slot] = d[5];
        s.det_ind[slot] = (float)det;
        if (c.vote_cls) vote_cls(c, s, slot, d[5], d[4]);
        s.tmp_b[i] = tracked ? 1 : 2;
    }
    BMB_SYNC();
    if (BMB_WARP == 0) {
        const int na = warp_append(T, s.act_l, mb[MB_N_ACT], [&](int i) { return s.tmp_b[i] == 1; },
                                   [&](int i) { return rows[i]; });
        const int nr = warp_append(T, s.refind_l, mb[MB_N_REFIND], [&](int i) { return s.tmp_b[i] == 2; },
                                   [&](int i) { return rows[i]; });
        int ne = mb[MB_N_EMA];
        if (use_feat) {
            warp_append(T, s.ema_slot, ne, [&](int i) { return s.tmp_b[i] != 0; }, [&](int i) { return rows[i]; });
            ne = warp_append(T, s.ema_det, ne, [&](int i) { return s.tmp_b[i] != 0; },
                             [&](int i) { return cols[s.lap_x[i]]; });
        }
        BMB_SYNCWARP();
        if (BMB_LANE == 0) { mb[MB_N_ACT] = na; mb[MB_N_REFIND] = nr; mb[MB_N_EMA] = ne; }
    }
    BMB_SYNC();
}

// removed-set membership of a live track (ByteTrack: unbounded list -> slot flag; BoT-SORT: deque of ids)
BMB_FN int in_removed_set(const TrkCfg& c, const TrkStream& s, int slot) {
    if (c.removed_cap == 0) return s.in_removed[slot];
    const int id = s.id[slot];
    const int n = s.scalars[SC_RING_COUNT];
    for (int k = 0; k < n; ++k)
        if (s.removed_ring[k] == id) return 1;
    return 0;
}

BMB_FN void push_removed(const TrkCfg& c, TrkStream& s, int slot) {
    if (c.removed_cap == 0) { s.in_removed[slot] = 1; return; }
    const int head = s.scalars[SC_RING_HEAD], n = s.scalars[SC_RING_COUNT];
    if (n < c.removed_cap) {
        s.removed_ring[(head + n) % c.removed_cap] = s.id[slot];
        s.scalars[SC_RING_COUNT] = n + 1;
    } else {
        s.removed_ring[head] = s.id[slot];
        s.scalars[SC_RING_HEAD] = (head + 1) % c.removed_cap;
    }
}

// cost rows, one warp per track row: the track box is loaded once, detections stream through the lanes
template <typename CostFn>
BMB_FN void build_cost(TrkStream& s, int T, int D, int ld, const int* rows, CostFn fn) {
    for (int i = BMB_WARP; i < T; i += BMB_NW) {
        const int t = rows[i];
        const double tb[4] = {s.txyxy[t * 4], s.txyxy[t * 4 + 1], s.txyxy[t * 4 + 2], s.txyxy[t * 4 + 3]};
        double* ci = s.cost + (size_t)i * ld;
        // four independent evaluations in flight per lane: the loads (L2 latency) and the float64 divide overlap
        for (int j0 = BMB_LANE; j0 < D; j0 += 4 * BMB_NL) {
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j0 + q * BMB_NL;
                v[q] = j < D ? fn(t, tb, j) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j0 + q * BMB_NL;
                if (j < D) ci[j] = v[q];
            }
        }
    }
    BMB_SYNC();
}

// ---------------------------------------------------------------------------------------------------
// the frame
// ---------------------------------------------------------------------------------------------------
// phase timers: thread 0 accumulates elapsed SM clocks since the previous boundary into timers[phase]
#define BMB_PHASE(idx)                                                        \
    do {                                                                      \
        if (BMB_TID == 0) {                                                   \
            long long _now = BMB_CLOCK();                                     \
            s.timers[idx] += _now - _t_prev;                                  \
            _t_prev = _now;                                                   \
        }                                                                     \
    } while (0)

BMB_FN void tracker_frame(const TrkCfg& c, TrkStream& s) {
    const int CT = c.cap_tracks, CD = c.cap_dets, F = c.feat_dim;
    long long _t_prev = BMB_CLOCK();
    int* mb = s.free_l;                  // mailbox: MB_COUNT ints written by one thread, read by all after a sync
    int* free_slots = s.free_l + MB_COUNT;
    int D = *s.n_dets;
    if (D > CD) {
        if (BMB_TID == 0) s.scalars[SC_ERROR] = ERR_DET_CAPACITY;
        D = CD;
    }
    const int frame = s.scalars[SC_FRAME] + 1;
    const int n_active0 = s.scalars[SC_N_ACTIVE], n_lost0 = s.scalars[SC_N_LOST];
    BMB_SYNC();

    // ---- S0/S1: detection geometry + confidence split; unconfirmed / pool lists ----
    for (int d = BMB_TID; d < D; d += BMB_NT)
        det_geometry(c, s.dets + d * 6, s.dxywh + d * 4, s.dmeas + d * 4, s.dxyxy + d * 4);
    for (int k = BMB_TID; k < CT; k += BMB_NT) s.mark[k] = 0;
    if (BMB_WARP == 0) {
        const int nf = warp_append(D, s.first, 0, [&](int d) { return (double)s.dets[d * 6 + 4] > c.high_thresh; },
                                   [&](int d) { return d; });
        const int ns = warp_append(D, s.second, 0, [&](int d) {
            const double cf = (double)s.dets[d * 6 + 4];
            return !(cf > c.high_thresh) && cf > c.low_thresh && cf < c.high_thresh; }, [&](int d) { return d; });
        const int nu = warp_append(n_active0, s.unconf, 0, [&](int k) { return !s.activated[s.active[k]]; },
                                   [&](int k) { return s.active[k]; });
        const int np = warp_append(n_active0, s.pool, 0, [&](int k) { return s.activated[s.active[k]] != 0; },
                                   [&](int k) { return s.active[k]; });
        if (BMB_LANE == 0) {
            mb[MB_N_FIRST] = nf; mb[MB_N_SECOND] = ns; mb[MB_N_UNC] = nu; mb[MB_N_POOL0] = np;
            mb[MB_N_ACT] = 0; mb[MB_N_REFIND] = 0; mb[MB_N_EMA] = 0;
        }
    }
    BMB_SYNC();
    const int n_first = mb[MB_N_FIRST], n_second = mb[MB_N_SECOND], n_unc = mb[MB_N_UNC];
    for (int k = BMB_TID; k < mb[MB_N_POOL0]; k += BMB_NT) s.mark[s.pool[k]] = 1;
    BMB_SYNC();
    if (BMB_WARP == 0) {  // joint(tracked, lost): a lost entry that is already in the pool is skipped
        const int np = warp_append(n_lost0, s.pool, mb[MB_N_POOL0], [&](int k) { return !s.mark[s.lost[k]]; },
                                   [&](int k) { return s.lost[k]; });
        if (BMB_LANE == 0) mb[MB_N_POOL] = np;
    }
    BMB_SYNC();
    const int n_pool = mb[MB_N_POOL];

    // ---- S2: Kalman predict over the pool (state is updated in place, as the reference does) ----
    const bool warp_pending = c.kind == KIND_XYWH && s.warp[6] != 0.0;  // BoT-SORT only (botsort.py:301)
    for (int k = BMB_TID; k < n_pool; k += BMB_NT) {
        const int t = s.pool[k];
        kf_predict(c, s.state[t] == ST_TRACKED, s.mean + t * 8, s.cov + t * 64);
        if (warp_pending) kf_apply_warp(s.warp, s.mean + t * 8, s.cov + t * 64);
        track_xyxy(c, s.mean + t * 8, s.txyxy + t * 4);
    }
    for (int k = BMB_TID; k < n_unc; k += BMB_NT) {
        const int t = s.unconf[k];
        if (warp_pending) kf_apply_warp(s.warp, s.mean + t * 8, s.cov + t * 64);
        track_xyxy(c, s.mean + t * 8, s.txyxy + t * 4);
    }
    BMB_SYNC();
    BMB_PHASE(0);  // split + pool + predict

    // ---- S3: first-association cost (botsort.py:306-317 / bytetrack.py:303-305) ----
    build_cost(s, n_pool, n_first, CD, s.pool, [&](int t, const double* tb, int j) {
        const int d = s.first[j];
        double iou_d = iou_dist_td(tb, s.dxyxy + d * 4);
        const int far = iou_d > c.proximity;
        if (c.fuse_first) {
            const double sim = 1.0 - iou_d;
            const double fs = sim * (double)s.dets[d * 6 + 4];
            iou_d = 1.0 - fs;
        }
        double v = iou_d;
        if (c.with_reid) {
            double em = s.embd[(size_t)t * CD + d];
            if (em > c.appearance) em = 1.0;
            if (c.proximity_mask && far) em = 1.0;
            v = (iou_d != iou_d || em != em) ? (double)NAN : (iou_d < em ? iou_d : em);
        }
        return v;
    });
    BMB_PHASE(1);  // cost build round 1
    lap_solve(s, n_pool, n_first, CD, c.match1);
    BMB_PHASE(2);  // assignment round 1
    apply_matches(c, s, n_pool, s.pool, s.first, frame, 0, c.with_reid, mb);
    BMB_PHASE(3);  // Kalman update + bookkeeping round 1

    // ---- S6: second association: remaining Tracked pool rows x low-confidence detections ----
    if (BMB_WARP == 0) {
        const int nrt = warp_append(n_pool, s.rtracked, 0,
                                    [&](int i) { return s.lap_x[i] < 0 && s.state[s.pool[i]] == ST_TRACKED; },
                                    [&](int i) { return s.pool[i]; });
        const int nrest = warp_append(n_first, s.rest, 0, [&](int j) { return s.lap_y[j] < 0; },
                                      [&](int j) { return s.first[j]; });
        if (BMB_LANE == 0) { mb[MB_N_RT] = nrt; mb[MB_N_REST] = nrest; }
    }
    BMB_SYNC();
    const int n_rt = mb[MB_N_RT], n_rest = mb[MB_N_REST];
    build_cost(s, n_rt, n_second, CD, s.rtracked,
               [&](int t, const double* tb, int j) { return iou_dist_td(tb, s.dxyxy + s.second[j] * 4); });
    lap_solve(s, n_rt, n_second, CD, c.match2);
    apply_matches(c, s, n_rt, s.rtracked, s.second, frame, 0, 0, mb);
    for (int i = BMB_TID; i < n_rt; i += BMB_NT) {
        int flag = 0;
        if (s.lap_x[i] < 0) {
            const int t = s.rtracked[i];
            if (s.state[t] != ST_LOST) { s.state[t] = ST_LOST; flag = 1; }
        }
        s.tmp_b[i] = flag;
    }
    BMB_SYNC();
    if (BMB_WARP == 0) {
        const int nl = warp_append(n_rt, s.lostnow_l, 0, [&](int i) { return s.tmp_b[i] != 0; },
                                   [&](int i) { return s.rtracked[i]; });
        if (BMB_LANE == 0) mb[MB_N_LOSTNOW] = nl;
    }
    BMB_SYNC();
    BMB_PHASE(4);  // second association

    // ---- S7: unconfirmed tracks x detections left over from round 1 ----
    build_cost(s, n_unc, n_rest, CD, s.unconf, [&](int t, const double* tb, int j) {
        const int d = s.rest[j];
        double iou_d = iou_dist_td(tb, s.dxyxy + d * 4);
        const int far = iou_d > c.proximity;
        {
            const double sim = 1.0 - iou_d;
            const double fs = sim * (double)s.dets[d * 6 + 4];
            iou_d = 1.0 - fs;
        }
        double v = iou_d;
        if (c.with_reid) {
            double em = s.embd[(size_t)t * CD + d] / c.unc_emb_scale;
            if (em > c.appearance) em = 1.0;
            if (c.proximity_mask && far) em = 1.0;
            v = (iou_d != iou_d || em != em) ? (double)NAN : (iou_d < em ? iou_d : em);
        }
        return v;
    });
    lap_solve(s, n_unc, n_rest, CD, c.match3);
    apply_matches(c, s, n_unc, s.unconf, s.rest, frame, 1, c.with_reid, mb);
    BMB_PHASE(5);  // unconfirmed round

    // ---- S8: removals of unmatched unconfirmed tracks, births (botsort.py:425-440 / bytetrack.py:362-373) ----
    for (int i = BMB_TID; i < n_unc; i += BMB_NT)
        if (s.lap_x[i] < 0) s.state[s.unconf[i]] = ST_REMOVED;
    for (int k = BMB_TID; k < CT; k += BMB_NT) s.mark[k] = 0;
    BMB_SYNC();
    for (int k = BMB_TID; k < n_active0; k += BMB_NT) s.mark[s.active[k]] = 1;
    for (int k = BMB_TID; k < n_lost0; k += BMB_NT) s.mark[s.lost[k]] = 1;
    BMB_SYNC();
    if (BMB_WARP == 0) {
        const int nrem = warp_append(n_unc, s.remnow_l, 0, [&](int i) { return s.lap_x[i] < 0; },
                                     [&](int i) { return s.unconf[i]; });
        // birth candidates: rest[j] unmatched in round 3 whose confidence passes the float32 gate
        int nb = warp_append(n_rest, s.tmp_a, 0,
                             [&](int j) { return s.lap_y[j] < 0 && !(s.dets[s.rest[j] * 6 + 4] < c.new_thresh_f32); },
                             [&](int j) { return s.rest[j]; });
        // that many free slots (anything not referenced by the active / lost lists)
        int nfree = 0;
        for (int k0 = 0; k0 < CT && nfree < nb; k0 += BMB_NL) {
            const int k = k0 + BMB_LANE;
            const bool p = k < CT && !s.mark[k];
            const unsigned m = BMB_BALLOT(p);
#if BMB_DEVICE
            const int pos = nfree + __popc(m & ((1u << BMB_LANE) - 1u));
#else
            const int pos = nfree;
#endif
            if (p && pos < nb) free_slots[pos] = k;
            nfree += BMB_POPC(m);
        }
        if (nfree < nb) {
            if (BMB_LANE == 0) s.scalars[SC_ERROR] = ERR_TRACK_CAPACITY;
            nb = nfree;
        }
        if (BMB_LANE == 0) { mb[MB_N_REMNOW] = nrem; mb[MB_N_BIRTH] = nb; }
    }
    BMB_SYNC();
    const int n_birth = mb[MB_N_BIRTH];
    {
        const int id0 = s.scalars[SC_NEXT_ID];
        for (int k = BMB_TID; k < n_birth; k += BMB_NT) {
            const int t = free_slots[k], d = s.tmp_a[k];
            const float cf = s.dets[d * 6 + 4];
            s.id[t] = id0 + 1 + k;
            s.tracklet_len[t] = 0;
            s.state[t] = ST_TRACKED;
            s.activated[t] = (frame == 1) ? 1 : 0;
            s.frame_id[t] = frame;
            s.start_frame[t] = frame;
            s.conf[t] = cf;
            s.cls[t] = s.dets[d * 6 + 5];
            s.det_ind[t] = (float)d;
            s.in_removed[t] = 0;
            s.hist_n[t] = 1;
            s.hist_cls[t * HIST_CAP] = s.dets[d * 6 + 5];
            s.hist_sum[t * HIST_CAP] = cf;
            kf_initiate(c, s.dmeas + d * 4, s.mean + t * 8, s.cov + t * 64);
        }
        if (c.with_reid) {
            for (int k = BMB_WARP; k < n_birth; k += BMB_NW) {
                const int t = free_slots[k], d = s.tmp_a[k];
                for (int q = BMB_LANE; q < F; q += BMB_NL) s.smooth[(size_t)t * F + q] = s.dfeat[(size_t)d * F + q];
            }
        }
    }
    // lost tracks that timed out (botsort.py:472-476): flags first, ordered append below
    for (int k = BMB_TID; k < n_lost0; k += BMB_NT) {
        const int t = s.lost[k];
        int flag = 0;
        if (frame - s.frame_id[t] > c.max_time_lost) { s.state[t] = ST_REMOVED; flag = 1; }
        s.tmp_b[k] = flag;
    }
    for (int k = BMB_TID; k < CT; k += BMB_NT) s.mark[k] = 0;
    BMB_SYNC();
    if (BMB_WARP == 0) {
        const int na = warp_append(n_birth, s.act_l, mb[MB_N_ACT], [&](int) { return true; },
                                   [&](int k) { return free_slots[k]; });
        const int nrem = warp_append(n_lost0, s.remnow_l, mb[MB_N_REMNOW], [&](int k) { return s.tmp_b[k] != 0; },
                                     [&](int k) { return s.lost[k]; });
        // active' part 1: still-Tracked members of the old active list
        const int m1 = warp_append(n_active0, s.pool, 0, [&](int k) { return s.state[s.active[k]] == ST_TRACKED; },
                                   [&](int k) { return s.active[k]; });
        if (BMB_LANE == 0) {
            mb[MB_N_ACT] = na; mb[MB_N_REMNOW] = nrem; mb[MB_M1] = m1;
            s.scalars[SC_NEXT_ID] += n_birth;
        }
    }
    BMB_SYNC();
    for (int k = BMB_TID; k < mb[MB_M1]; k += BMB_NT) s.mark[s.pool[k]] = 1;
    BMB_SYNC();
    if (BMB_WARP == 0) {  // joint with the activated list
        const int m2 = warp_append(mb[MB_N_ACT], s.pool, mb[MB_M1], [&](int k) { return !s.mark[s.act_l[k]]; },
                                   [&](int k) { return s.act_l[k]; });
        if (BMB_LANE == 0) mb[MB_M2] = m2;
    }
    BMB_SYNC();
    for (int k = mb[MB_M1] + BMB_TID; k < mb[MB_M2]; k += BMB_NT) s.mark[s.pool[k]] = 1;
    BMB_SYNC();
    if (BMB_WARP == 0) {  // joint with the refind list; lost' = (lost - active') ++ lost_now
        const int m3 = warp_append(mb[MB_N_REFIND], s.pool, mb[MB_M2], [&](int k) { return !s.mark[s.refind_l[k]]; },
                                   [&](int k) { return s.refind_l[k]; });
        if (BMB_LANE == 0) mb[MB_M3] = m3;
    }
    BMB_SYNC();
    for (int k = mb[MB_M2] + BMB_TID; k < mb[MB_M3]; k += BMB_NT) s.mark[s.pool[k]] = 1;
    BMB_SYNC();
    if (BMB_WARP == 0) {
        int q = warp_append(n_lost0, s.rtracked, 0, [&](int k) { return !s.mark[s.lost[k]]; },
                            [&](int k) { return s.lost[k]; });
        q = warp_append(mb[MB_N_LOSTNOW], s.rtracked, q, [&](int) { return true; },
                        [&](int k) { return s.lostnow_l[k]; });
        if (BMB_LANE == 0) mb[MB_Q1] = q;
    }
    BMB_SYNC();
    // ... minus everything in the removed set (evaluated BEFORE this frame's removals are pushed)
    for (int k = BMB_TID; k < mb[MB_Q1]; k += BMB_NT) s.tmp_b[k] = in_removed_set(c, s, s.rtracked[k]);
    BMB_SYNC();
    if (BMB_WARP == 0) {
        const int q2 = warp_append(mb[MB_Q1], s.tmp_a, 0, [&](int k) { return !s.tmp_b[k]; },
                                   [&](int k) { return s.rtracked[k]; });
        if (BMB_LANE == 0) {
            mb[MB_Q2] = q2;
            const int nrem = mb[MB_N_REMNOW];
            for (int k = 0; k < nrem; ++k) push_removed(c, s, s.remnow_l[k]);
        }
    }
    BMB_SYNC();
    BMB_PHASE(6);  // births + list algebra

    // ---- duplicate suppression (botsort_utils.py:53-82): IoU distance active' x lost' < 0.15 ----
    const int m = mb[MB_M3], q = mb[MB_Q2];
    for (int k = BMB_TID; k < m; k += BMB_NT) track_xyxy(c, s.mean + s.pool[k] * 8, s.txyxy + s.pool[k] * 4);
    for (int k = BMB_TID; k < q; k += BMB_NT) track_xyxy(c, s.mean + s.tmp_a[k] * 8, s.txyxy + s.tmp_a[k] * 4);
    BMB_SYNC();
    // duplicate flags go into `mark` (1 = member of active'): 2 = drop from lost', 3 = drop from active'
    for (int e = BMB_TID; e < m * q; e += BMB_NT) {
        const int pi = e / q, r = e - pi * q;
        const int ta = s.pool[pi], tb = s.tmp_a[r];
        const double dist = iou_dist_tt(s.txyxy + ta * 4, s.txyxy + tb * 4);
        if (dist < 0.15) {
            const int tp = s.frame_id[ta] - s.start_frame[ta];
            const int tq = s.frame_id[tb] - s.start_frame[tb];
            if (tp > tq) s.mark[tb] = 2; else s.mark[ta] = 3;
        }
    }
    BMB_SYNC();
    if (BMB_WARP == 0) {
        const int na = warp_append(m, s.active, 0, [&](int k) { return s.mark[s.pool[k]] != 3; },
                                   [&](int k) { return s.pool[k]; });
        const int nl = warp_append(q, s.lost, 0, [&](int k) { return s.mark[s.tmp_a[k]] != 2; },
                                   [&](int k) { return s.tmp_a[k]; });
        BMB_SYNCWARP();
        int n_out = warp_append(na, s.tmp_b, 0, [&](int k) { return s.activated[s.active[k]] != 0; },
                                [&](int k) { return s.active[k]; });
        if (n_out > CD) {
            if (BMB_LANE == 0) s.scalars[SC_ERROR] = ERR_DET_CAPACITY;
            n_out = CD;
        }
        if (BMB_LANE == 0) {
            s.scalars[SC_N_ACTIVE] = na;
            s.scalars[SC_N_LOST] = nl;
            s.scalars[SC_FRAME] = frame;
            s.scalars[SC_N_OUT] = n_out;
            s.scalars[SC_N_EMA] = mb[MB_N_EMA];
        }
    }
    BMB_SYNC();
    // output rows (botsort.py:494-500)
    for (int k = BMB_TID; k < s.scalars[SC_N_OUT]; k += BMB_NT) {
        const int t = s.tmp_b[k];
        double b[4];
        track_xyxy(c, s.mean + t * 8, b);
        float* o = s.out + k * 8;
        o[0] = (float)b[0]; o[1] = (float)b[1]; o[2] = (float)b[2]; o[3] = (float)b[3];
        o[4] = (float)s.id[t];
        o[5] = s.conf[t];
        o[6] = s.cls[t];
        o[7] = s.det_ind[t];
    }
    BMB_SYNC();
    BMB_PHASE(7);  // duplicate suppression + output rows
}

// Deferred appearance EMA of this frame's matches (botsort_track.py:58-67 via update()/re_activate()).
// Warp-cooperative; `pair` indexes the list recorded by tracker_frame.
BMB_FN void apply_feature_ema(const TrkCfg& c, TrkStream& s, int pair) {
    feat_ema(s.smooth + (size_t)s.ema_slot[pair] * c.feat_dim, s.dfeat + (size_t)s.ema_det[pair] * c.feat_dim, c.feat_dim);
}

}  // namespace bmb
