// docs_core.cuh -- DeepOCSORT per-stream, per-frame update (one CTA per stream), same conventions as
// tracker_core.cuh (the source also compiles for the host under BMB_HOSTSIM for GPU-less control-flow tests).
//
// Replaces (relative to /root/reference/boxmot):
//   trackers/bbox/deepocsort/deepocsort.py:51-233   KalmanBoxTracker (predict / update / update_emb, observation
//                                                   bookkeeping: last_observation, observations by age, velocity)
//   trackers/bbox/deepocsort/deepocsort.py:302-492  DeepOcSort._update_impl
//   motion/kalman_filters/xysr.py:379-476 + base.py:366-459  7-state XYSR filter: predict_state, update_state
//                                                   (symmetrised S, Joseph form), freeze / unfreeze replay
//   trackers/association/association.py:8-152       speed_direction_batch, compute_aw_max_metric, associate,
//                                                   linear_assignment (lapjv(extend_cost=True), no cost limit)
// The filter keeps (x, P), the copy frozen when the track stopped being observed, the last observed measurement
// and the number of missed frames: that is all the reference's deepcopy(__dict__) + history deque replay reads.
#pragma once
#include "tracker_core.cuh"
#include "jv_dense.cuh"

namespace bmb {

enum { DOCS_RING = 8 };  // observation ring: delta_t + 1 most recent (age, box) pairs are enough (delta_t <= 7)

struct DocsCfg {
    int cap_tracks, cap_dets, feat_dim, delta_t, max_age, min_hits, embedding_off, aw_off;
    float det_thresh_f32;  // `scores > det_thresh` compares a float32 array with a python float in float32
    double det_thresh, iou_threshold, inertia, w_emb, alpha_fixed, aw_param, q_xy, q_s;
};

struct DocsStream {
    // ---- persistent, slot indexed ----
    double* x;         // [CT][8]   (7 used)
    double* P;         // [CT][56]  (7x7 row major, stride 7)
    double* xs;        // frozen copies
    double* Ps;
    double* zlast;     // [CT][4] last observed measurement
    int* gap;
    int* has_saved;
    int* observed;
    int* age;
    int* hits;
    int* hit_streak;
    int* tsu;          // time_since_update
    int* id;
    double* conf;
    double* cls;
    double* det_ind;
    double* last_obs;  // [CT][5]
    double* obs_box;   // [CT][DOCS_RING][5]
    int* obs_age;      // [CT][DOCS_RING]
    int* obs_n;        // [CT] observations recorded so far (ring index = n % DOCS_RING)
    double* vel;       // [CT][2]
    int* has_vel;
    double* emb;       // [CT][F] float64 (the reference's EMA promotes to float64)
    int* tracks;       // [CT] ordered slot list (the reference's active_tracks list)
    int* scalars;      // [SC_COUNT]
    long long* timers;
    // ---- inputs ----
    const float* dets;   // [CD][6]
    const int* n_dets;
    const float* embs;   // [CD][F] float32 rows (unit norm, as get_features returns them), may be null
    // ---- scratch ----
    int* kdet;         // [CD] kept detection indices
    double* dbox;      // [CD][5] kept detection rows x1,y1,x2,y2,conf as float64
    double* dalpha;    // [CD]
    double* tbox;      // [CT][4] predicted boxes by list position
    double* kobs;      // [CT][5]
    double* iou;       // [CD][LD]
    double* embq;      // [CD][CT] appearance similarity by (raw detection, slot): filled by the wide k_docs_embcost
    double* embc;      // [CD][LD]
    double* cost;      // [max(CD,CT)][LD2] assignment cost (possibly transposed)
    double* top;       // [2*(CD+CT)] row / column top-2 values
    int* mrow;         // [CD] matched track position per kept detection (or -1)
    int* und;          // [CD] unmatched kept-detection positions
    int* unt;          // [CT] unmatched track positions
    int* tmp_a;        // [CT+CD]
    int* tmp_b;        // [max(CT,CD)]
    int* mark;         // [CT]
    int* free_l;       // [MB_COUNT + CT]
    int* lap_x; int* lap_y; double* lap_u; double* lap_v; double* lap_spc; int* lap_path; int* lap_insc;
    int* lap_tl; int* lap_sc; int* csr_ptr; int* csr_col;
    float* out;        // [CD][8]
    int jv_wide;       // CTA-wide augmentation in the dense JV solver (set per launch, not part of the carved state)
    const double* warp;  // [8]: supplied 2x3 camera-motion matrix (row major), warp[6] != 0 when one is pending; may be null
};

// one entry of the appearance similarity matrix (float32 detection row x float64 track EMA, float64 accumulate)
BMB_FN double docs_emb_dot(const DocsCfg& c, const DocsStream& s, int d_raw, int slot) {
    const float* de = s.embs + (size_t)d_raw * c.feat_dim;
    const double* te = s.emb + (size_t)slot * c.feat_dim;
    double acc = 0.0;
    for (int q = 0; q < c.feat_dim; ++q) acc += (double)de[q] * te[q];
    return acc;
}

BMB_FN double iou_ff(const double* a, const double* b) {
    double xx1 = a[0] > b[0] ? a[0] : b[0];
    double yy1 = a[1] > b[1] ? a[1] : b[1];
    double xx2 = a[2] < b[2] ? a[2] : b[2];
    double yy2 = a[3] < b[3] ? a[3] : b[3];
    double w = xx2 - xx1; w = w > 0.0 ? w : 0.0;
    double h = yy2 - yy1; h = h > 0.0 ? h : 0.0;
    double wh = w * h;
    return wh / ((a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - wh);
}

// ---- 7-state XYSR Kalman filter -----------------------------------------------------------------------------
BMB_FN void xysr_enforce(double* x, double* P) {
    if (x[2] < 1e-6) x[2] = 1e-6;
    if (x[3] < 1e-6) x[3] = 1e-6;
    for (int i = 0; i < 7; ++i)
        for (int j = i + 1; j < 7; ++j) {
            const double v = 0.5 * (P[i * 7 + j] + P[j * 7 + i]);
            P[i * 7 + j] = v;
            P[j * 7 + i] = v;
        }
}

BMB_FN void xysr_predict(const DocsCfg& c, double* x, double* P) {
    // x <- F x ; P <- F P F^T + Q   (F = I + e0 e4^T + e1 e5^T + e2 e6^T); exact two-term sums as in numpy
    for (int i = 0; i < 3; ++i) x[i] = x[i] + x[i + 4];
    double L[49];
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) L[i * 7 + j] = (i < 3) ? (P[i * 7 + j] + P[(i + 4) * 7 + j]) : P[i * 7 + j];
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) {
            double v = (j < 3) ? (L[i * 7 + j] + L[i * 7 + j + 4]) : L[i * 7 + j];
            if (i == j) v = v + ((i < 4) ? 1.0 : (i < 6 ? c.q_xy : c.q_s));
            P[i * 7 + j] = v;
        }
    xysr_enforce(x, P);
}

// lower Cholesky factor of the symmetric 4x4 S + ridge*I; false as soon as a pivot is not positive (LAPACK potrf's
// criterion: ajj <= 0 or NaN)
BMB_FN bool xysr_chol4(const double* S, double ridge, double* Lc) {
    for (int i = 0; i < 16; ++i) Lc[i] = 0.0;
    for (int j = 0; j < 4; ++j) {
        double d = S[j * 4 + j] + ridge;
        for (int k = 0; k < j; ++k) d -= Lc[j * 4 + k] * Lc[j * 4 + k];
        if (!(d > 0.0)) return false;
        d = sqrt(d);
        Lc[j * 4 + j] = d;
        for (int i = j + 1; i < 4; ++i) {
            double v = S[i * 4 + j];
            for (int k = 0; k < j; ++k) v -= Lc[i * 4 + k] * Lc[j * 4 + k];
            Lc[i * 4 + j] = v / d;
        }
    }
    return true;
}

// base.py:414-459 with R = diag(1,1,10,10), H = [I4 | 0]
BMB_FN void xysr_update_state(double* x, double* P, const double* z) {
    const double Rd[4] = {1.0, 1.0, 10.0, 10.0};
    double S[16], Lc[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) S[i * 4 + j] = P[i * 7 + j] + (i == j ? Rd[i] : 0.0);
    for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j) {
            const double v = 0.5 * (S[i * 4 + j] + S[j * 4 + i]);
            S[i * 4 + j] = v; S[j * 4 + i] = v;
        }
    // _safe_cho_factor (base.py:462-500): plain Cholesky; when a pivot is not positive, a ridge of
    // max|diag S| * 10^e, e = -12 .. 3, is added until the factorisation goes through.  Camera-motion correction
    // (apply_affine_correction transforms the position and velocity blocks of P but not their cross terms) makes this
    // the normal path of young tracks, not a rarity.  The eigenvalue-clipping last resort is only reachable with
    // non-finite input, where the reference raises as well.
    if (!xysr_chol4(S, 0.0, Lc)) {
        double scale = 0.0;
        for (int i = 0; i < 4; ++i) { const double a = fabs(S[i * 4 + i]); if (a > scale) scale = a; }
        if (!(scale > 0.0) || scale > 1.7976931348623157e308) scale = 1.0;
        const double p10[16] = {1e-12, 1e-11, 1e-10, 1e-09, 1e-08, 1e-07, 1e-06, 1e-05, 0.0001, 0.001, 0.01, 0.1,
                                1.0, 10.0, 100.0, 1000.0};
        for (int e = 0; e < 16; ++e)
            if (xysr_chol4(S, scale * p10[e], Lc)) break;
    }
    double K[28];  // [7][4]
    for (int r = 0; r < 7; ++r) {
        double y[4], kt[4];
        for (int i = 0; i < 4; ++i) {
            double v = P[r * 7 + i];
            for (int k = 0; k < i; ++k) v -= Lc[i * 4 + k] * y[k];
            y[i] = v / Lc[i * 4 + i];
        }
        for (int i = 3; i >= 0; --i) {
            double v = y[i];
            for (int k = i + 1; k < 4; ++k) v -= Lc[k * 4 + i] * kt[k];
            kt[i] = v / Lc[i * 4 + i];
        }
        for (int i = 0; i < 4; ++i) K[r * 4 + i] = kt[i];
    }
    double yv[4];
    for (int i = 0; i < 4; ++i) yv[i] = z[i] - x[i];
    for (int r = 0; r < 7; ++r) {
        double acc = 0.0;
        for (int i = 0; i < 4; ++i) acc += K[r * 4 + i] * yv[i];
        x[r] = x[r] + acc;
    }
    // Joseph form: P <- A (P A^T) + K R K^T with A = I - K H
    double A[49], PA[49];
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) A[i * 7 + j] = (i == j ? 1.0 : 0.0) - (j < 4 ? K[i * 4 + j] : 0.0);
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 7; ++k) acc += P[i * 7 + k] * A[j * 7 + k];
            PA[i * 7 + j] = acc;
        }
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 7; ++k) acc += A[i * 7 + k] * PA[k * 7 + j];
            double krk = 0.0;
            for (int k = 0; k < 4; ++k) krk += K[i * 4 + k] * (Rd[k] * K[j * 4 + k]);
            P[i * 7 + j] = acc + krk;
        }
    for (int i = 0; i < 7; ++i)
        for (int j = i + 1; j < 7; ++j) {
            const double v = 0.5 * (P[i * 7 + j] + P[j * 7 + i]);
            P[i * 7 + j] = v; P[j * 7 + i] = v;
        }
}

BMB_FN void xysr_prepare(const double* z, double* m) {
    m[0] = z[0]; m[1] = z[1];
    m[2] = z[2] < 1e-6 ? 1e-6 : z[2];
    m[3] = z[3] < 1e-6 ? 1e-6 : z[3];
}

BMB_FN void xysr_observe(double* x, double* P, double* zlast, const double* m) {
    xysr_update_state(x, P, m);
    xysr_enforce(x, P);
    for (int i = 0; i < 4; ++i) zlast[i] = m[i];
}

BMB_FN void xyxy_to_xysr(const double* b, double* z) {
    const double w = b[2] - b[0], h = b[3] - b[1];
    z[0] = b[0] + w / 2.0;
    z[1] = b[1] + h / 2.0;
    z[2] = w * h;
    z[3] = w / (h + 1e-6);
}

BMB_FN void xysr_to_box(const double* x, double* o) {
    const double w = sqrt(x[2] * x[3]);
    const double h = x[2] / w;
    o[0] = x[0] - w / 2.0; o[1] = x[1] - h / 2.0; o[2] = x[0] + w / 2.0; o[3] = x[1] + h / 2.0;
}

// KalmanFilterXYSR.update(z) for an observed box (xysr.py:440-476), including the unfreeze replay (:384-438)
BMB_FN void docs_kf_update(const DocsCfg& c, DocsStream& s, int t, const double* z) {
    double* x = s.x + t * 8;
    double* P = s.P + t * 56;
    double m[4];
    xysr_prepare(z, m);
    if (!s.observed[t] && s.has_saved[t]) {
        for (int i = 0; i < 7; ++i) x[i] = s.xs[t * 8 + i];
        for (int i = 0; i < 49; ++i) P[i] = s.Ps[t * 56 + i];
        s.has_saved[t] = 0;
        const double* za = s.zlast + t * 4;
        const double x1 = za[0], y1 = za[1], s1 = za[2], r1 = za[3];
        const double w1 = sqrt(s1 * r1), h1 = sqrt(s1 / r1);
        const double x2 = m[0], y2 = m[1], s2 = m[2], r2 = m[3];
        const double w2 = sqrt(s2 * r2), h2 = sqrt(s2 / r2);
        const int gap = s.gap[t] + 1;
        const double dx = (x2 - x1) / gap, dy = (y2 - y1) / gap, dw = (w2 - w1) / gap, dh = (h2 - h1) / gap;
        double zl[4];
        for (int i = 0; i < gap; ++i) {
            const double xx = x1 + (i + 1) * dx, yy = y1 + (i + 1) * dy;
            const double ww = w1 + (i + 1) * dw, hh = h1 + (i + 1) * dh;
            const double box[4] = {xx, yy, ww * hh, ww / hh};
            double mm[4];
            xysr_prepare(box, mm);
            xysr_observe(x, P, zl, mm);
            if (i != gap - 1) xysr_predict(c, x, P);
        }
    }
    s.observed[t] = 1;
    s.gap[t] = 0;
    xysr_observe(x, P, s.zlast + t * 4, m);
}

BMB_FN void docs_kf_miss(DocsStream& s, int t) {
    if (s.observed[t]) {
        for (int i = 0; i < 7; ++i) s.xs[t * 8 + i] = s.x[t * 8 + i];
        for (int i = 0; i < 49; ++i) s.Ps[t * 56 + i] = s.P[t * 56 + i];
        s.has_saved[t] = 1;
        s.gap[t] = 0;
    }
    s.gap[t] += 1;
    s.observed[t] = 0;
}

// KalmanBoxTracker.update(det) bookkeeping + filter update (deepocsort.py:138-177); det row = dbox[kd]
BMB_FN void docs_track_update(const DocsCfg& c, DocsStream& s, int t, int kd) {
    const double* bb = s.dbox + kd * 5;
    s.conf[t] = bb[4];
    s.cls[t] = (double)s.dets[s.kdet[kd] * 6 + 5];
    s.det_ind[t] = (double)s.kdet[kd];
    double* lo = s.last_obs + t * 5;
    if (lo[0] + lo[1] + lo[2] + lo[3] + lo[4] >= 0) {
        const double* prev = nullptr;
        for (int dt = c.delta_t; dt > 0 && !prev; --dt) {
            const int want = s.age[t] - dt;
            for (int k = 0; k < DOCS_RING; ++k)
                if (k < s.obs_n[t] && s.obs_age[t * DOCS_RING + k] == want) { prev = s.obs_box + (t * DOCS_RING + k) * 5; break; }
        }
        if (!prev) prev = lo;
        const double cx1 = (prev[0] + prev[2]) / 2.0, cy1 = (prev[1] + prev[3]) / 2.0;
        const double cx2 = (bb[0] + bb[2]) / 2.0, cy2 = (bb[1] + bb[3]) / 2.0;
        const double sy = cy2 - cy1, sx = cx2 - cx1;
        const double nrm = sqrt(sy * sy + sx * sx) + 1e-6;
        s.vel[t * 2] = sy / nrm;
        s.vel[t * 2 + 1] = sx / nrm;
        s.has_vel[t] = 1;
    }
    for (int i = 0; i < 5; ++i) lo[i] = bb[i];
    {   // observations[age] = bbox (an existing entry for the same age is overwritten)
        int slot = -1;
        const int n = s.obs_n[t];
        for (int k = 0; k < DOCS_RING && k < n; ++k)
            if (s.obs_age[t * DOCS_RING + k] == s.age[t]) slot = k;
        if (slot < 0) { slot = n % DOCS_RING; s.obs_n[t] = n + 1; }
        s.obs_age[t * DOCS_RING + slot] = s.age[t];
        for (int i = 0; i < 5; ++i) s.obs_box[(t * DOCS_RING + slot) * 5 + i] = bb[i];
    }
    s.tsu[t] = 0;
    s.hits[t] += 1;
    s.hit_streak[t] += 1;
    double z[4];
    xyxy_to_xysr(bb, z);
    docs_kf_update(c, s, t, z);
}

// update_emb (deepocsort.py:179-181) in float64; warp-cooperative over the feature dimension
BMB_FN void docs_emb_update(const DocsCfg& c, DocsStream& s, int t, int kd) {
    const int F = c.feat_dim;
    const double alpha = s.dalpha[kd];
    const float* e = s.embs + (size_t)s.kdet[kd] * F;
    double* te = s.emb + (size_t)t * F;
    double acc = 0.0;
    for (int k = BMB_LANE; k < F; k += BMB_NL) {
        const double v = alpha * te[k] + (1 - alpha) * (double)e[k];
        te[k] = v;
        acc += v * v;
    }
#if BMB_DEVICE
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
#endif
    const double nrm = sqrt(acc);
    BMB_SYNCWARP();
    for (int k = BMB_LANE; k < F; k += BMB_NL) te[k] = te[k] / nrm;
    BMB_SYNCWARP();
}

// KalmanBoxTracker.apply_affine_correction (deepocsort.py:189-206) + KalmanFilterXYSR.apply_affine_correction
// (xysr.py:311-366) for a supplied 2x3 warp W.  In the reference `last_observation` and `observations[age]` are views of
// the same detection row, so the newest observation is warped by the first statement AND again by the window loop when
// it lies inside the velocity window: both copies kept here follow that.  The pre-gap measurement used by the
// un-freeze replay comes from the live history, which the reference does not warp (zlast stays).
BMB_FN void docs_warp_box(const double* W, double* b) {
    const double x1 = W[0] * b[0] + W[1] * b[1] + W[2], y1 = W[3] * b[0] + W[4] * b[1] + W[5];
    const double x2 = W[0] * b[2] + W[1] * b[3] + W[2], y2 = W[3] * b[2] + W[4] * b[3] + W[5];
    b[0] = x1; b[1] = y1; b[2] = x2; b[3] = y2;
}

BMB_FN void docs_warp_state(const double* W, double* x, double* P) {
    const double m00 = W[0], m01 = W[1], m10 = W[3], m11 = W[4];
    const double px = m00 * x[0] + m01 * x[1] + W[2], py = m10 * x[0] + m11 * x[1] + W[5];
    x[0] = px; x[1] = py;
    const double vx = m00 * x[4] + m01 * x[5], vy = m10 * x[4] + m11 * x[5];
    x[4] = vx; x[5] = vy;
    for (int b = 0; b < 2; ++b) {   // P[o:o+2, o:o+2] <- (m @ P_blk) @ m^T, numpy's left-to-right order
        const int o = b * 4;
        const double a = P[o * 7 + o], bq = P[o * 7 + o + 1], cq = P[(o + 1) * 7 + o], dq = P[(o + 1) * 7 + o + 1];
        const double t00 = m00 * a + m01 * cq, t01 = m00 * bq + m01 * dq;
        const double t10 = m10 * a + m11 * cq, t11 = m10 * bq + m11 * dq;
        P[o * 7 + o] = t00 * m00 + t01 * m01;
        P[o * 7 + o + 1] = t00 * m10 + t01 * m11;
        P[(o + 1) * 7 + o] = t10 * m00 + t11 * m01;
        P[(o + 1) * 7 + o + 1] = t10 * m10 + t11 * m11;
    }
}

BMB_FN void docs_apply_warp(const DocsCfg& c, DocsStream& s, int t, const double* W) {
    double* lo = s.last_obs + t * 5;
    const int n = s.obs_n[t];
    int newest = -1;   // the ring entry that aliases last_observation in the reference
    for (int k = 0; k < DOCS_RING && k < n; ++k)
        if (newest < 0 || s.obs_age[t * DOCS_RING + k] > s.obs_age[t * DOCS_RING + newest]) newest = k;
    if (lo[0] + lo[1] + lo[2] + lo[3] + lo[4] > 0) {
        docs_warp_box(W, lo);
        if (newest >= 0) for (int i = 0; i < 4; ++i) s.obs_box[(t * DOCS_RING + newest) * 5 + i] = lo[i];
    }
    for (int dt = c.delta_t; dt >= 0; --dt) {
        const int want = s.age[t] - dt;
        for (int k = 0; k < DOCS_RING && k < n; ++k)
            if (s.obs_age[t * DOCS_RING + k] == want) {
                double* ob = s.obs_box + (t * DOCS_RING + k) * 5;
                docs_warp_box(W, ob);
                if (k == newest) for (int i = 0; i < 4; ++i) lo[i] = ob[i];
                break;
            }
    }
    docs_warp_state(W, s.x + t * 8, s.P + t * 56);
    if (!s.observed[t] && s.has_saved[t]) docs_warp_state(W, s.xs + t * 8, s.Ps + t * 56);
    xysr_enforce(s.x + t * 8, s.P + t * 56);
}

// k_previous_obs (deepocsort.py:13-22)
BMB_FN void docs_k_prev(const DocsCfg& c, const DocsStream& s, int t, double* o) {
    const int n = s.obs_n[t];
    if (n == 0) { for (int i = 0; i < 5; ++i) o[i] = -1.0; return; }
    for (int i = 0; i < c.delta_t; ++i) {
        const int want = s.age[t] - (c.delta_t - i);
        for (int k = 0; k < DOCS_RING && k < n; ++k)
            if (s.obs_age[t * DOCS_RING + k] == want) {
                for (int q = 0; q < 5; ++q) o[q] = s.obs_box[(t * DOCS_RING + k) * 5 + q];
                return;
            }
    }
    int best = 0;
    for (int k = 1; k < DOCS_RING && k < n; ++k)
        if (s.obs_age[t * DOCS_RING + k] > s.obs_age[t * DOCS_RING + best]) best = k;
    for (int q = 0; q < 5; ++q) o[q] = s.obs_box[(t * DOCS_RING + best) * 5 + q];
}

// Dense rectangular assignment with lapjv(extend_cost=True) semantics: the smaller side is fully matched.
// cost is (R, C) row major with leading dimension ld in s.iou-sized scratch `src`; result mrow[r] = c or -1.
template <typename CostAt>
BMB_FN void docs_assign(DocsStream& s, int R, int C, int ld2, int* result, CostAt cost_at) {
    // lapjv pads the (R, C) matrix with zeros to n = max(R, C) and solves the square problem; rows assigned to a
    // padded column come back as -1 (lap.lapjv: x[x >= n_cols] = -1)
    const int n = R > C ? R : C;
    for (int e = BMB_TID; e < n * n; e += BMB_NT) {
        const int i = e / n, j = e - i * n;
        s.cost[(size_t)i * ld2 + j] = (i < R && j < C) ? cost_at(i, j) : 0.0;
    }
    BMB_SYNC();
    jv_dense_solve(s, n, ld2, R, s.jv_wide);
    for (int r = BMB_TID; r < R; r += BMB_NT) result[r] = s.lap_x[r] < C ? s.lap_x[r] : -1;
    BMB_SYNC();
}

BMB_FN void docs_frame(const DocsCfg& c, DocsStream& s) {
    const int CT = c.cap_tracks, CD = c.cap_dets, F = c.feat_dim;
    const int LD = CT;                       // leading dimension of the (det, track) matrices
    const int LD2 = CT > CD ? CT : CD;       // leading dimension of the assignment scratch
    int* mb = s.free_l;
    int* free_slots = s.free_l + MB_COUNT;
    int D = *s.n_dets;
    if (D > CD) { if (BMB_TID == 0) s.scalars[SC_ERROR] = ERR_DET_CAPACITY; D = CD; }
    const int frame = s.scalars[SC_FRAME] + 1;
    const int n_trk0 = s.scalars[SC_N_ACTIVE];
    const bool use_emb = !c.embedding_off && s.embs != nullptr;
    long long _t_prev = BMB_CLOCK();   // phase clocks: 0 dets+predict, 1 iou/appearance, 2 first assignment (cost + solver:
    BMB_SYNC();                        // 8..10 inside), 3 updates, 4 second round, 5 misses+births, 6 emit+cull

    // ---- kept detections (scores > det_thresh in float32), trust -> alpha (deepocsort.py:330-354) ----
    if (BMB_WARP == 0) {
        const int nk = warp_append(D, s.kdet, 0, [&](int d) { return s.dets[d * 6 + 4] > c.det_thresh_f32; },
                                   [&](int d) { return d; });
        if (BMB_LANE == 0) mb[0] = nk;
    }
    BMB_SYNC();
    const int nk = mb[0];
    for (int k = BMB_TID; k < nk; k += BMB_NT) {
        const float* r = s.dets + s.kdet[k] * 6;
        double* b = s.dbox + k * 5;
        for (int i = 0; i < 5; ++i) b[i] = (double)r[i];
        const double trust = (b[4] - c.det_thresh) / (1 - c.det_thresh);
        s.dalpha[k] = c.alpha_fixed + (1 - c.alpha_fixed) * (1 - trust);
    }
    // ---- predict every track (deepocsort.py:361-368, :208-223), drop NaN boxes ----
    const bool warp_pending = s.warp != nullptr && s.warp[6] != 0.0;   // deepocsort.py:345-348, estimator replaced by input
    for (int k = BMB_TID; k < n_trk0; k += BMB_NT) {
        const int t = s.tracks[k];
        if (warp_pending) docs_apply_warp(c, s, t, s.warp);
        double* x = s.x + t * 8;
        if (x[6] + x[2] <= 0) x[6] *= 0.0;
        xysr_predict(c, x, s.P + t * 56);
        s.age[t] += 1;
        if (s.tsu[t] > 0) s.hit_streak[t] = 0;
        s.tsu[t] += 1;
        double b[4];
        xysr_to_box(x, b);
        s.tmp_b[k] = (b[0] != b[0] || b[1] != b[1] || b[2] != b[2] || b[3] != b[3]) ? 1 : 0;
    }
    BMB_SYNC();
    if (BMB_WARP == 0) {
        const int n = warp_append(n_trk0, s.tmp_a, 0, [&](int k) { return !s.tmp_b[k]; }, [&](int k) { return s.tracks[k]; });
        if (BMB_LANE == 0) mb[1] = n;
    }
    BMB_SYNC();
    const int T = mb[1];
    for (int k = BMB_TID; k < T; k += BMB_NT) s.tracks[k] = s.tmp_a[k];
    BMB_SYNC();
    for (int k = BMB_TID; k < T; k += BMB_NT) {
        const int t = s.tracks[k];
        xysr_to_box(s.x + t * 8, s.tbox + k * 4);
        docs_k_prev(c, s, t, s.kobs + k * 5);
    }
    BMB_SYNC();

    BMB_PHASE(0);
    // ---- first association (association.py:61-152) ----
    int n_und = 0, n_unt = 0;
    if (T == 0) {
        for (int k = BMB_TID; k < nk; k += BMB_NT) { s.mrow[k] = -1; s.und[k] = k; }
        n_und = nk;
        BMB_SYNC();
    } else {
        // iou, appearance similarity (float64 accumulate), per-row / per-column threshold counts
        for (int e = BMB_TID; e < nk * T; e += BMB_NT) {
            const int d = e / T, k = e - d * T;
            s.iou[(size_t)d * LD + k] = iou_ff(s.dbox + d * 5, s.tbox + k * 4);
        }
        if (use_emb) {
            // dets_embs @ trk_embs.T (deepocsort.py:391) was computed for every (detection, slot) by the wide grid
            for (int e = BMB_TID; e < nk * T; e += BMB_NT) {
                const int d = e / T, k = e - d * T;
                s.embc[(size_t)d * LD + k] = s.embq[(size_t)s.kdet[d] * CT + s.tracks[k]];
            }
        }
        BMB_SYNC();
        BMB_PHASE(1);
        // 1-1 shortcut test: every row and column has at most one entry above the threshold, and some row has one
        if (BMB_TID == 0) { mb[2] = 0; mb[3] = 0; }
        BMB_SYNC();
        for (int d = BMB_TID; d < nk; d += BMB_NT) {
            int n = 0;
            for (int k = 0; k < T; ++k) n += s.iou[(size_t)d * LD + k] > c.iou_threshold ? 1 : 0;
            BMB_ATOMIC_MAX(&mb[2], n > 1 ? 2 : n);
        }
        for (int k = BMB_TID; k < T; k += BMB_NT) {
            int n = 0;
            for (int d = 0; d < nk; ++d) n += s.iou[(size_t)d * LD + k] > c.iou_threshold ? 1 : 0;
            BMB_ATOMIC_MAX(&mb[3], n > 1 ? 2 : n);
        }
        BMB_SYNC();
        const bool have = nk > 0;
        // numpy: a.sum(1).max() == 1 and a.sum(0).max() == 1 ; a racing "1" may be overwritten by "2" only
        int rmax = mb[2], cmax = mb[3];
        BMB_SYNC();
        if (!have) {
            for (int k = BMB_TID; k < nk; k += BMB_NT) s.mrow[k] = -1;
            BMB_SYNC();
        } else if (rmax == 1 && cmax == 1) {
            for (int d = BMB_TID; d < nk; d += BMB_NT) {
                int m = -1;
                for (int k = 0; k < T; ++k)
                    if (s.iou[(size_t)d * LD + k] > c.iou_threshold) m = k;
                s.mrow[d] = m;
            }
            BMB_SYNC();
        } else {
            // emb_cost[iou <= 0] = 0 ; adaptive weighting by the top-2 ratio of every row and column
            double* rtop = s.top;                 // [nk][2]
            double* ctop = s.top + 2 * (size_t)CD;  // [T][2]
            if (use_emb) {
                for (int e = BMB_TID; e < nk * T; e += BMB_NT) {
                    const int d = e / T, k = e - d * T;
                    if (s.iou[(size_t)d * LD + k] <= 0) s.embc[(size_t)d * LD + k] = 0.0;
                }
                BMB_SYNC();
                if (!c.aw_off) {
                    for (int d = BMB_TID; d < nk; d += BMB_NT) {
                        double a = -INFINITY, b = -INFINITY;
                        for (int k = 0; k < T; ++k) {
                            const double v = s.embc[(size_t)d * LD + k];
                            if (v > a) { b = a; a = v; } else if (v > b) b = v;
                        }
                        rtop[d * 2] = a; rtop[d * 2 + 1] = b;
                    }
                    for (int k = BMB_TID; k < T; k += BMB_NT) {
                        double a = -INFINITY, b = -INFINITY;
                        for (int d = 0; d < nk; ++d) {
                            const double v = s.embc[(size_t)d * LD + k];
                            if (v > a) { b = a; a = v; } else if (v > b) b = v;
                        }
                        ctop[k * 2] = a; ctop[k * 2 + 1] = b;
                    }
                    BMB_SYNC();
                }
            }
            auto weight = [&](double a, double b, int len) {
                if (len < 2) return 1.0;
                if (a == 0) return 0.0;
                double r = (b / a) - c.aw_param;
                r = r > 0 ? r : 0.0;
                return 1 - r / (1 - c.aw_param);
            };
            docs_assign(s, nk, T, LD2, s.mrow, [&](int d, int k) {
                const double io = s.iou[(size_t)d * LD + k];
                // velocity-direction consistency (association.py:8-17, 83-100)
                const double* po = s.kobs + k * 5;
                const double* db = s.dbox + d * 5;
                const double dxx = (db[0] + db[2]) / 2.0 - (po[0] + po[2]) / 2.0;
                const double dyy = (db[1] + db[3]) / 2.0 - (po[1] + po[3]) / 2.0;
                const double nrm = sqrt(dxx * dxx + dyy * dyy) + 1e-6;
                const double X = dxx / nrm, Y = dyy / nrm;
                const int t = s.tracks[k];
                const double vy = s.has_vel[t] ? s.vel[t * 2] : 0.0, vx = s.has_vel[t] ? s.vel[t * 2 + 1] : 0.0;
                double cs = vx * X + vy * Y;
                cs = cs < -1.0 ? -1.0 : (cs > 1.0 ? 1.0 : cs);
                const double ang = (3.141592653589793 / 2.0 - fabs(acos(cs))) / 3.141592653589793;
                const double valid = po[4] < 0 ? 0.0 : 1.0;
                const double angc = ((valid * ang) * c.inertia) * db[4];
                double em = 0.0;
                if (use_emb) {
                    const double e0 = s.embc[(size_t)d * LD + k];
                    if (c.aw_off) em = e0 * c.w_emb;
                    else em = ((c.w_emb * weight(rtop[d * 2], rtop[d * 2 + 1], T)) * weight(ctop[k * 2], ctop[k * 2 + 1], nk)) * e0;
                }
                return -((io + angc) + em);
            });
        }
        // unmatched lists + IoU filter (association.py:126-152)
        if (BMB_TID == 0) {
            for (int k = 0; k < T; ++k) s.mark[k] = 0;
            for (int d = 0; d < nk; ++d) if (s.mrow[d] >= 0) s.mark[s.mrow[d]] = 1;
            for (int d = 0; d < nk; ++d) if (s.mrow[d] < 0) s.und[n_und++] = d;
            for (int k = 0; k < T; ++k) if (!s.mark[k]) s.unt[n_unt++] = k;
            for (int d = 0; d < nk; ++d) {
                const int k = s.mrow[d];
                if (k >= 0 && s.iou[(size_t)d * LD + k] < c.iou_threshold) {
                    s.und[n_und++] = d;
                    s.unt[n_unt++] = k;
                    s.mrow[d] = -1;
                }
            }
            mb[4] = n_und; mb[5] = n_unt;
        }
        BMB_SYNC();
        n_und = mb[4]; n_unt = mb[5];
    }
    BMB_PHASE(2);
    // apply the first-round matches
    for (int d = BMB_TID; d < nk; d += BMB_NT)
        if (s.mrow[d] >= 0) docs_track_update(c, s, s.tracks[s.mrow[d]], d);
    if (use_emb)
        for (int d = BMB_WARP; d < nk; d += BMB_NW)
            if (s.mrow[d] >= 0) docs_emb_update(c, s, s.tracks[s.mrow[d]], d);
    BMB_SYNC();

    BMB_PHASE(3);
    // ---- second round: observation-centric recovery on the last observations (deepocsort.py:414-450) ----
    if (n_und > 0 && n_unt > 0) {
        // last_boxes were gathered before any update: unmatched tracks have not been touched this frame
        if (BMB_TID == 0) mb[6] = 0;
        BMB_SYNC();
        for (int e = BMB_TID; e < n_und * n_unt; e += BMB_NT) {
            const int a = e / n_unt, b = e - a * n_unt;
            const double v = iou_ff(s.dbox + s.und[a] * 5, s.last_obs + s.tracks[s.unt[b]] * 5);
            s.iou[(size_t)a * LD + b] = v;
            if (v > c.iou_threshold) mb[6] = 1;
        }
        BMB_SYNC();
        const bool any = mb[6] != 0;
        BMB_SYNC();
        if (any) {
            // mrow is reused for the re-match: save the first-round result of the rows involved is not needed
            // (all rows in `und` are unmatched); docs_assign writes mrow[0..n_und)
            for (int d = BMB_TID; d < n_und; d += BMB_NT) s.tmp_a[d] = s.und[d];
            for (int k = BMB_TID; k < n_unt; k += BMB_NT) s.tmp_a[CD + k] = s.unt[k];
            BMB_SYNC();
            // first-round mrow stays intact: the re-match goes to tmp_b
            docs_assign(s, n_und, n_unt, LD2, s.tmp_b, [&](int a, int b) { return -s.iou[(size_t)a * LD + b]; });
            for (int a = BMB_TID; a < n_und; a += BMB_NT) {
                const int b = s.tmp_b[a];
                if (b >= 0 && s.iou[(size_t)a * LD + b] < c.iou_threshold) s.tmp_b[a] = -1;
            }
            BMB_SYNC();
            for (int a = BMB_TID; a < n_und; a += BMB_NT)
                if (s.tmp_b[a] >= 0) docs_track_update(c, s, s.tracks[s.tmp_a[CD + s.tmp_b[a]]], s.tmp_a[a]);
            if (use_emb)
                for (int a = BMB_WARP; a < n_und; a += BMB_NW)
                    if (s.tmp_b[a] >= 0) docs_emb_update(c, s, s.tracks[s.tmp_a[CD + s.tmp_b[a]]], s.tmp_a[a]);
            BMB_SYNC();
            // np.setdiff1d: sorted unique remainder
            if (BMB_TID == 0) {
                for (int k = 0; k < T; ++k) s.mark[k] = 0;
                for (int a = 0; a < n_und; ++a) if (s.tmp_b[a] >= 0) s.mark[s.tmp_a[CD + s.tmp_b[a]]] = 1;
                // detections: collect the still-unmatched positions and sort ascending (insertion sort, short)
                int nd = 0;
                for (int a = 0; a < n_und; ++a) if (s.tmp_b[a] < 0) s.und[nd++] = s.tmp_a[a];
                for (int i = 1; i < nd; ++i) { int v = s.und[i], j = i - 1; while (j >= 0 && s.und[j] > v) { s.und[j + 1] = s.und[j]; --j; } s.und[j + 1] = v; }
                int nt = 0;
                for (int k = 0; k < n_unt; ++k) if (!s.mark[s.tmp_a[CD + k]]) s.unt[nt++] = s.tmp_a[CD + k];
                for (int i = 1; i < nt; ++i) { int v = s.unt[i], j = i - 1; while (j >= 0 && s.unt[j] > v) { s.unt[j + 1] = s.unt[j]; --j; } s.unt[j + 1] = v; }
                // np.setdiff1d also removes duplicates
                int u = 0;
                for (int i = 0; i < nd; ++i) if (i == 0 || s.und[i] != s.und[i - 1]) s.und[u++] = s.und[i];
                nd = u; u = 0;
                for (int i = 0; i < nt; ++i) if (i == 0 || s.unt[i] != s.unt[i - 1]) s.unt[u++] = s.unt[i];
                nt = u;
                mb[4] = nd; mb[5] = nt;
            }
            BMB_SYNC();
            n_und = mb[4]; n_unt = mb[5];
        }
    }
    BMB_PHASE(4);
    // ---- misses, births (deepocsort.py:452-466) ----
    for (int k = BMB_TID; k < n_unt; k += BMB_NT) docs_kf_miss(s, s.tracks[s.unt[k]]);
    for (int k = BMB_TID; k < CT; k += BMB_NT) s.mark[k] = 0;
    BMB_SYNC();
    for (int k = BMB_TID; k < T; k += BMB_NT) s.mark[s.tracks[k]] = 1;
    BMB_SYNC();
    if (BMB_WARP == 0) {
        int nfree = 0;
        for (int k0 = 0; k0 < CT && nfree < n_und; k0 += BMB_NL) {
            const int k = k0 + BMB_LANE;
            const bool p = k < CT && !s.mark[k];
            const unsigned m = BMB_BALLOT(p);
#if BMB_DEVICE
            const int pos = nfree + __popc(m & ((1u << BMB_LANE) - 1u));
#else
            const int pos = nfree;
#endif
            if (p && pos < n_und) free_slots[pos] = k;
            nfree += BMB_POPC(m);
        }
        if (BMB_LANE == 0) {
            if (nfree < n_und) s.scalars[SC_ERROR] = ERR_TRACK_CAPACITY;
            mb[7] = nfree < n_und ? nfree : n_und;
        }
    }
    BMB_SYNC();
    const int n_birth = mb[7];
    {
        const int id0 = s.scalars[SC_NEXT_ID];
        for (int k = BMB_TID; k < n_birth; k += BMB_NT) {
            const int t = free_slots[k], kd = s.und[k];
            const double* bb = s.dbox + kd * 5;
            double* x = s.x + t * 8;
            double* P = s.P + t * 56;
            double z[4];
            xyxy_to_xysr(bb, z);
            for (int i = 0; i < 7; ++i) x[i] = i < 4 ? z[i] : 0.0;
            for (int i = 0; i < 49; ++i) P[i] = 0.0;
            for (int i = 0; i < 7; ++i) P[i * 7 + i] = i < 4 ? 10.0 : 10000.0;
            s.gap[t] = 0; s.has_saved[t] = 0; s.observed[t] = 0;
            s.age[t] = 0; s.hits[t] = 0; s.hit_streak[t] = 0; s.tsu[t] = 0;
            s.id[t] = id0 + 1 + k;   // KalmanBoxTracker.count starts at 1 (deepocsort.py:293)
            s.conf[t] = bb[4];
            s.cls[t] = (double)s.dets[s.kdet[kd] * 6 + 5];
            s.det_ind[t] = (double)s.kdet[kd];
            for (int i = 0; i < 5; ++i) s.last_obs[t * 5 + i] = -1.0;
            s.obs_n[t] = 0;
            s.has_vel[t] = 0;
            s.tracks[T + k] = t;
        }
        if (use_emb)
            for (int k = BMB_WARP; k < n_birth; k += BMB_NW) {
                const int t = free_slots[k], kd = s.und[k];
                const float* e = s.embs + (size_t)s.kdet[kd] * F;
                for (int q = BMB_LANE; q < F; q += BMB_NL) s.emb[(size_t)t * F + q] = (double)e[q];
            }
    }
    BMB_SYNC();
    const int n_all = T + n_birth;
    BMB_PHASE(5);
    // ---- emit (reversed list order) and cull (deepocsort.py:467-489) ----
    for (int k = BMB_TID; k < n_all; k += BMB_NT) {
        const int t = s.tracks[k];
        s.tmp_b[k] = (s.tsu[t] < 1 && (s.hit_streak[t] >= c.min_hits || frame <= c.min_hits)) ? 1 : 0;
    }
    BMB_SYNC();
    if (BMB_WARP == 0) {
        int n_out = warp_append(n_all, s.tmp_a, 0, [&](int r) { return s.tmp_b[n_all - 1 - r] != 0; },
                                [&](int r) { return s.tracks[n_all - 1 - r]; });
        if (n_out > CD) { if (BMB_LANE == 0) s.scalars[SC_ERROR] = ERR_DET_CAPACITY; n_out = CD; }
        const int keep = warp_append(n_all, s.unt, 0, [&](int k) { return s.tsu[s.tracks[k]] <= c.max_age; },
                                     [&](int k) { return s.tracks[k]; });
        if (BMB_LANE == 0) { mb[8] = n_out; mb[9] = keep; }
    }
    BMB_SYNC();
    for (int k = BMB_TID; k < mb[8]; k += BMB_NT) {
        const int t = s.tmp_a[k];
        const double* lo = s.last_obs + t * 5;
        double b[4];
        if (lo[0] + lo[1] + lo[2] + lo[3] + lo[4] < 0) xysr_to_box(s.x + t * 8, b);
        else { b[0] = lo[0]; b[1] = lo[1]; b[2] = lo[2]; b[3] = lo[3]; }
        float* o = s.out + k * 8;
        o[0] = (float)b[0]; o[1] = (float)b[1]; o[2] = (float)b[2]; o[3] = (float)b[3];
        o[4] = (float)s.id[t]; o[5] = (float)s.conf[t]; o[6] = (float)s.cls[t]; o[7] = (float)s.det_ind[t];
    }
    for (int k = BMB_TID; k < mb[9]; k += BMB_NT) s.tmp_b[k] = s.unt[k];
    BMB_SYNC();
    for (int k = BMB_TID; k < mb[9]; k += BMB_NT) s.tracks[k] = s.tmp_b[k];
    if (BMB_TID == 0) {
        s.scalars[SC_N_ACTIVE] = mb[9];
        s.scalars[SC_FRAME] = frame;
        s.scalars[SC_N_OUT] = mb[8];
        s.scalars[SC_NEXT_ID] += n_birth;
    }
    BMB_SYNC();
    BMB_PHASE(6);
}

}  // namespace bmb
