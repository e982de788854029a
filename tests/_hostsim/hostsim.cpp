// tests/_hostsim/hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles boxmot_b200/csrc/tracker_core.cuh for the host (BMB_HOSTSIM: one "thread") so the tracker
// control flow can be checked against the golden vectors in a container without a GPU.  It is built by
// tests/test_hostsim_tracker.py with g++, loaded only by that test, and never linked into libboxmot_b200.so.
#define BMB_HOSTSIM 1
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tracker_core.cuh"
#include "tracker_layout.h"
#include "cmc_ecc.cuh"

using namespace bmb;

struct HostSim {
    TrkCfg cfg;
    TrkStream s;
    std::vector<uint8_t> mem;
    std::vector<float> dets;
    std::vector<float> embs;
    int n_dets;
    double warp[8];
};

extern "C" {

HostSim* hostsim_create(const TrkCfg* cfg) {
    HostSim* h = new HostSim();
    h->cfg = *cfg;
    size_t bytes = carve_stream(h->cfg, nullptr, nullptr, nullptr);
    h->mem.assign(bytes, 0);
    carve_stream(h->cfg, h->mem.data(), &h->s, nullptr);
    h->dets.assign((size_t)cfg->cap_dets * 6, 0.f);
    h->embs.assign((size_t)cfg->cap_dets * (cfg->feat_dim > 0 ? cfg->feat_dim : 1), 0.f);
    h->s.dets = h->dets.data();
    h->s.n_dets = &h->n_dets;
    for (double& w : h->warp) w = 0.0;
    h->s.warp = h->warp;
    return h;
}

void hostsim_destroy(HostSim* h) { delete h; }

int hostsim_cfg_size() { return (int)sizeof(TrkCfg); }

void hostsim_set_warp(HostSim* h, const double* w6) {
    for (int i = 0; i < 6; ++i) h->warp[i] = w6[i];
    h->warp[6] = 1.0;
}

// returns number of output rows, or -(error code)
int hostsim_update(HostSim* h, const float* dets, int n, const float* embs, float* out, int* lap_steps) {
    const TrkCfg& c = h->cfg;
    if (n > c.cap_dets) return -ERR_DET_CAPACITY;
    h->n_dets = n;
    if (n) memcpy(h->dets.data(), dets, sizeof(float) * 6 * n);
    if (c.with_reid && embs) {
        const int F = c.feat_dim;
        for (int d = 0; d < n; ++d) feat_prepare(embs + (size_t)d * F, h->s.dfeat + (size_t)d * F, F);
        // the wide appearance kernel: every occupied slot against every detection
        std::vector<char> live(c.cap_tracks, 0);
        for (int k = 0; k < h->s.scalars[SC_N_ACTIVE]; ++k) live[h->s.active[k]] = 1;
        for (int k = 0; k < h->s.scalars[SC_N_LOST]; ++k) live[h->s.lost[k]] = 1;
        for (int t = 0; t < c.cap_tracks; ++t) {
            if (!live[t]) continue;
            for (int d = 0; d < n; ++d)
                h->s.embd[(size_t)t * c.cap_dets + d] =
                    cosine_cost_f64(h->s.smooth + (size_t)t * F, h->s.dfeat + (size_t)d * F, F);
        }
    }
    h->s.scalars[SC_LAP_STEPS] = 0;
    tracker_frame(c, h->s);
    h->warp[6] = 0.0;  // a warp applies to one frame
    if (c.with_reid)
        for (int k = 0; k < h->s.scalars[SC_N_EMA]; ++k) apply_feature_ema(c, h->s, k);
    if (lap_steps) *lap_steps = h->s.scalars[SC_LAP_STEPS];
    if (h->s.scalars[SC_ERROR]) return -h->s.scalars[SC_ERROR];
    int m = h->s.scalars[SC_N_OUT];
    memcpy(out, h->s.out, sizeof(float) * 8 * m);
    return m;
}

// snapshot of live tracks: ids, means, covs; returns count
int hostsim_snapshot(HostSim* h, int* ids, double* means, double* covs, int cap) {
    int n = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int cnt = h->s.scalars[pass == 0 ? SC_N_ACTIVE : SC_N_LOST];
        const int* lst = pass == 0 ? h->s.active : h->s.lost;
        for (int k = 0; k < cnt && n < cap; ++k, ++n) {
            int t = lst[k];
            ids[n] = h->s.id[t];
            memcpy(means + (size_t)n * 8, h->s.mean + (size_t)t * 8, sizeof(double) * 8);
            memcpy(covs + (size_t)n * 64, h->s.cov + (size_t)t * 64, sizeof(double) * 64);
        }
    }
    return n;
}

// ---- DeepOCSORT host simulation -------------------------------------------------------------------------------
struct DocsSim {
    DocsCfg cfg;
    DocsStream s;
    double warp[8];
    std::vector<uint8_t> mem;
    std::vector<float> dets, embs;
    int n_dets;
};

DocsSim* docs_create(const DocsCfg* cfg) {
    DocsSim* h = new DocsSim();
    h->cfg = *cfg;
    size_t bytes = carve_docs(h->cfg, nullptr, nullptr, nullptr);
    h->mem.assign(bytes, 0);
    carve_docs(h->cfg, h->mem.data(), &h->s, nullptr);
    h->dets.assign((size_t)cfg->cap_dets * 6, 0.f);
    h->embs.assign((size_t)cfg->cap_dets * (cfg->feat_dim > 0 ? cfg->feat_dim : 1), 0.f);
    h->s.dets = h->dets.data();
    h->s.n_dets = &h->n_dets;
    h->s.embs = nullptr;
    for (double& w : h->warp) w = 0.0;
    h->s.warp = h->warp;
    return h;
}
void docs_destroy(DocsSim* h) { delete h; }
void docs_set_warp(DocsSim* h, const double* w6) {
    for (int i = 0; i < 6; ++i) h->warp[i] = w6[i];
    h->warp[6] = 1.0;
}
int docs_cfg_size() { return (int)sizeof(DocsCfg); }

int docs_update(DocsSim* h, const float* dets, int n, const float* embs, float* out) {
    const DocsCfg& c = h->cfg;
    if (n > c.cap_dets) return -ERR_DET_CAPACITY;
    h->n_dets = n;
    if (n) memcpy(h->dets.data(), dets, sizeof(float) * 6 * n);
    if (embs && n) {
        memcpy(h->embs.data(), embs, sizeof(float) * (size_t)c.feat_dim * n);
        h->s.embs = h->embs.data();
    } else {
        h->s.embs = embs ? h->embs.data() : nullptr;
    }
    if (h->s.embs && !c.embedding_off)   // the wide appearance kernel: every live slot against every detection
        for (int k = 0; k < h->s.scalars[SC_N_ACTIVE]; ++k)
            for (int d = 0; d < n; ++d)
                h->s.embq[(size_t)d * c.cap_tracks + h->s.tracks[k]] = docs_emb_dot(c, h->s, d, h->s.tracks[k]);
    docs_frame(c, h->s);
    h->warp[6] = 0.0;  // a warp applies to one frame
    if (h->s.scalars[SC_ERROR]) return -h->s.scalars[SC_ERROR];
    const int m = h->s.scalars[SC_N_OUT];
    memcpy(out, h->s.out, sizeof(float) * 8 * m);
    return m;
}

int docs_snapshot(DocsSim* h, int* ids, double* xs, double* Ps, int cap) {
    int n = 0;
    for (int k = 0; k < h->s.scalars[SC_N_ACTIVE] && n < cap; ++k, ++n) {
        const int t = h->s.tracks[k];
        ids[n] = h->s.id[t];
        memcpy(xs + (size_t)n * 7, h->s.x + (size_t)t * 8, sizeof(double) * 7);
        memcpy(Ps + (size_t)n * 49, h->s.P + (size_t)t * 56, sizeof(double) * 49);
    }
    return n;
}

void docs_set_jv_wide(DocsSim* h, int wide) { h->s.jv_wide = wide == 3 ? (3 | (0xF << 2)) : wide; }   // as Engine::set_jv_wide

int docs_jv(DocsSim* h, const double* cost, int R, int C, int* x, int* y) {
    const int n = R > C ? R : C;
    const int ld = h->cfg.cap_tracks > h->cfg.cap_dets ? h->cfg.cap_tracks : h->cfg.cap_dets;
    if (n > ld) return -1;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) h->s.cost[(size_t)i * ld + j] = (i < R && j < C) ? cost[(size_t)i * C + j] : 0.0;
    jv_dense_solve(h->s, n, ld, R, h->s.jv_wide);
    for (int i = 0; i < R; ++i) x[i] = h->s.lap_x[i] < C ? h->s.lap_x[i] : -1;
    for (int j = 0; j < C; ++j) y[j] = h->s.lap_y[j] < R ? h->s.lap_y[j] : -1;
    return 0;
}

// standalone LAP entry for solver tests: cost is (T, D) row-major
int hostsim_lap(HostSim* h, const double* cost, int T, int D, double thresh, int* x, int* y) {
    const TrkCfg& c = h->cfg;
    if (T > c.cap_tracks || D > c.cap_dets) return -1;
    for (int i = 0; i < T; ++i) memcpy(h->s.cost + (size_t)i * c.cap_dets, cost + (size_t)i * D, sizeof(double) * D);
    lap_solve(h->s, T, D, c.cap_dets, thresh);
    memcpy(x, h->s.lap_x, sizeof(int) * T);
    memcpy(y, h->s.lap_y, sizeof(int) * D);
    return 0;
}

// ---- StrongSORT host simulation ---------------------------------------------------------------------------------
struct SsSim {
    SsCfg cfg;
    SsStream s;
    std::vector<uint8_t> mem;
    std::vector<float> dets, embs;
    int n_dets;
    double warp[8];
};

SsSim* ss_create(const SsCfg* cfg) {
    SsSim* h = new SsSim();
    h->cfg = *cfg;
    size_t bytes = carve_ss(h->cfg, nullptr, nullptr, nullptr);
    h->mem.assign(bytes, 0);
    carve_ss(h->cfg, h->mem.data(), &h->s, nullptr);
    h->dets.assign((size_t)cfg->cap_dets * 6, 0.f);
    h->embs.assign((size_t)cfg->cap_dets * (cfg->feat_dim > 0 ? cfg->feat_dim : 1), 0.f);
    h->s.dets = h->dets.data();
    h->s.n_dets = &h->n_dets;
    h->s.embs = h->embs.data();
    for (double& w : h->warp) w = 0.0;
    h->s.warp = h->warp;
    return h;
}
void ss_destroy(SsSim* h) { delete h; }
int ss_cfg_size() { return (int)sizeof(SsCfg); }
void ss_set_warp(SsSim* h, const double* w6) {
    for (int i = 0; i < 6; ++i) h->warp[i] = w6[i];
    h->warp[6] = 1.0;
}

int ss_update(SsSim* h, const float* dets, int n, const float* embs, float* out) {
    const SsCfg& c = h->cfg;
    if (n > c.cap_dets) return -ERR_DET_CAPACITY;
    h->n_dets = n;
    if (n) memcpy(h->dets.data(), dets, sizeof(float) * 6 * n);
    if (n) memcpy(h->embs.data(), embs, sizeof(float) * (size_t)c.feat_dim * n);
    for (int d = 0; d < n; ++d) ss_prepare_row(c, h->s, d);
    for (int k = 0; k < h->s.scalars[SC_N_ACTIVE]; ++k) {
        const int t = h->s.tracks[k];
        if (h->s.state[t] != SS_CONFIRMED) continue;
        for (int d = 0; d < n; ++d) h->s.appc[(size_t)t * c.cap_dets + d] = ss_nn_cosine(c, h->s, t, d);
    }
    ss_frame(c, h->s);
    h->warp[6] = 0.0;
    for (int k = 0; k < h->s.scalars[SC_N_ACTIVE]; ++k) ss_features_pos(c, h->s, k);
    if (h->s.scalars[SC_ERROR]) return -h->s.scalars[SC_ERROR];
    const int m = h->s.scalars[SC_N_OUT];
    memcpy(out, h->s.out, sizeof(float) * 8 * m);
    return m;
}

int ss_snapshot(SsSim* h, int* ids, double* means, double* covs, int cap) {
    int n = 0;
    for (int k = 0; k < h->s.scalars[SC_N_ACTIVE] && n < cap; ++k, ++n) {
        const int t = h->s.tracks[k];
        ids[n] = h->s.id[t];
        memcpy(means + (size_t)n * 8, h->s.mean + (size_t)t * 8, sizeof(double) * 8);
        memcpy(covs + (size_t)n * 64, h->s.cov + (size_t)t * 64, sizeof(double) * 64);
    }
    return n;
}

// scipy.optimize.linear_sum_assignment(cost (R, C)) -> row_ind, col_ind (min(R, C) pairs, rows ascending)
// diagnostics for the soak tools: the cost matrix of the last matching stage, in solver orientation
int ss_debug_cost(SsSim* h, double* out, int n) {
    memcpy(out, h->s.cost, sizeof(double) * (size_t)n);
    return 0;
}

int ss_lsa(SsSim* h, const double* cost, int R, int C, int* row_ind, int* col_ind) {
    const SsCfg& c = h->cfg;
    if ((size_t)R * C > (size_t)c.cap_tracks * c.cap_dets) return -1;
    const int MX = c.cap_tracks > c.cap_dets ? c.cap_tracks : c.cap_dets;
    if (R > MX || C > MX) return -1;
    const bool tr = C < R;
    const int nr = tr ? C : R, nc = tr ? R : C;
    for (int r = 0; r < R; ++r)
        for (int q = 0; q < C; ++q) h->s.cost[tr ? (size_t)q * nc + r : (size_t)r * nc + q] = cost[(size_t)r * C + q];
    lsa_solve(h->s, h->s.cost, nr, nc, nc);
    if (h->s.scalars[SC_ERROR]) { h->s.scalars[SC_ERROR] = 0; return -2; }
    int n = 0;
    if (!tr) {
        for (int i = 0; i < nr; ++i) { row_ind[n] = i; col_ind[n] = h->s.lsa_col4row[i]; ++n; }
    } else {
        for (int r = 0; r < R; ++r)
            if (h->s.lsa_row4col[r] >= 0) { row_ind[n] = r; col_ind[n] = h->s.lsa_row4col[r]; ++n; }
    }
    return n;
}

// list(set(a) - set(b)): a ascending, in_b[k] flags the members of b; returns the count
int ss_pyset_difference(SsSim* h, const int* a, int na, const unsigned char* in_b, int* out) {
    const SsCfg& c = h->cfg;
    if (na > c.cap_tracks) return -1;
    std::vector<unsigned char> flag(na ? a[na - 1] + 1 : 1, 0);
    int nb = 0, n = 0, umax = -1;
    for (int k = 0; k < na; ++k) {
        if (in_b[k]) { flag[a[k]] = 1; ++nb; }
        else { out[n++] = a[k]; umax = a[k]; }
    }
    if (na == 0 || pyset_difference_is_ascending(na, a[na - 1], nb, umax)) return n;
    const size_t stride = (size_t)8 * c.cap_tracks + 16;
    return pyset_difference(a, na, nb, [&](int key) { return flag[key] != 0; }, out, h->s.tmp_a, h->s.set_buf,
                            h->s.set_buf + stride, h->s.set_buf + 2 * stride, h->s.set_buf + 3 * stride);
}

// ---- camera-motion estimation (cmc_ecc.cuh) ---------------------------------------------------------------------
// BaseCMC.preprocess on one BGR frame: out is rint(rows * scale) x rint(cols * scale) uint8
void hostsim_cmc_prepare(const uint8_t* img, int rows, int cols, double scale, uint8_t* out, int h, int w) {
    const double inv = 1.0 / scale;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) out[(size_t)y * w + x] = cmc_prepare_pixel(img, rows, cols, inv, y, x);
}

int hostsim_ecc(const uint8_t* T, const uint8_t* I, int h, int w, double eps, int max_iter, float* txy) {
    double red[8];
    txy[0] = txy[1] = 0.f;
    return ecc_translation(T, I, h, w, eps, max_iter, red, txy);
}
}
