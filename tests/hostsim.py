"""Host simulation of the device tracker code (tests only; never a product path).

Builds tests/_hostsim/hostsim.cpp with g++ (BMB_HOSTSIM: boxmot_b200/csrc/tracker_core.cuh compiled for one
host "thread") and exposes it through ctypes so `-m "not gpu"` tests can check the control flow of the CUDA
tracker against the reference goldens without a GPU.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "_hostsim" / "hostsim.cpp"
OUT = ROOT / "tests" / "_hostsim" / "hostsim.so"
CSRC = ROOT / "boxmot_b200" / "csrc"


class TrkCfg(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int), ("with_reid", ctypes.c_int), ("fuse_first", ctypes.c_int),
        ("proximity_mask", ctypes.c_int), ("max_time_lost", ctypes.c_int), ("removed_cap", ctypes.c_int),
        ("feat_dim", ctypes.c_int), ("cap_tracks", ctypes.c_int), ("cap_dets", ctypes.c_int),
        ("vote_cls", ctypes.c_int),
        ("high_thresh", ctypes.c_double), ("low_thresh", ctypes.c_double),
        ("new_thresh_f32", ctypes.c_float),
        ("match1", ctypes.c_double), ("match2", ctypes.c_double), ("match3", ctypes.c_double),
        ("proximity", ctypes.c_double), ("appearance", ctypes.c_double), ("unc_emb_scale", ctypes.c_double),
    ]


def build():
    deps = [SRC] + sorted(CSRC.glob("*.cuh")) + [CSRC / "tracker_layout.h"]
    if OUT.exists() and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return OUT
    tmp = OUT.with_suffix(f".{os.getpid()}.tmp.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                           f"-I{CSRC}", "-o", str(tmp), str(SRC)])
    os.replace(tmp, OUT)
    return OUT


def bytetrack_cfg(min_conf=0.1, track_thresh=0.45, match_thresh=0.8, track_buffer=25, frame_rate=30,
                  cap_tracks=512, cap_dets=256):
    c = TrkCfg()
    c.kind, c.with_reid, c.fuse_first, c.proximity_mask = 0, 0, 1, 0
    c.max_time_lost = int(frame_rate / 30.0 * track_buffer)
    c.removed_cap, c.feat_dim, c.cap_tracks, c.cap_dets, c.vote_cls = 0, 0, cap_tracks, cap_dets, 0
    c.high_thresh, c.low_thresh = track_thresh, min_conf
    c.new_thresh_f32 = np.float32(track_thresh)
    c.match1, c.match2, c.match3 = match_thresh, 0.5, 0.7
    c.proximity, c.appearance, c.unc_emb_scale = 0.0, 0.0, 1.0
    return c


def botsort_cfg(track_high_thresh=0.5, track_low_thresh=0.1, new_track_thresh=0.6, track_buffer=30,
                match_thresh=0.8, proximity_thresh=0.5, appearance_thresh=0.25, frame_rate=30,
                fuse_first_associate=False, with_reid=True, second_match_thresh=0.5,
                unconfirmed_match_thresh=0.7, unconfirmed_emb_scale=2.0, removed_stracks_buffer=100,
                feat_dim=512, cap_tracks=512, cap_dets=256):
    c = TrkCfg()
    c.kind, c.with_reid, c.fuse_first, c.proximity_mask = 1, int(with_reid), int(fuse_first_associate), 1
    c.max_time_lost = int(frame_rate / 30.0 * track_buffer)
    c.removed_cap, c.feat_dim = removed_stracks_buffer, (feat_dim if with_reid else 0)
    c.cap_tracks, c.cap_dets, c.vote_cls = cap_tracks, cap_dets, 1
    c.high_thresh, c.low_thresh = track_high_thresh, track_low_thresh
    c.new_thresh_f32 = np.float32(new_track_thresh)
    c.match1, c.match2, c.match3 = match_thresh, second_match_thresh, unconfirmed_match_thresh
    c.proximity, c.appearance, c.unc_emb_scale = proximity_thresh, appearance_thresh, unconfirmed_emb_scale
    return c


class HostSimTracker:
    def __init__(self, cfg: TrkCfg):
        self.lib = ctypes.CDLL(str(build()))
        assert self.lib.hostsim_cfg_size() == ctypes.sizeof(TrkCfg)
        self.lib.hostsim_create.restype = ctypes.c_void_p
        self.lib.hostsim_create.argtypes = [ctypes.POINTER(TrkCfg)]
        self.lib.hostsim_update.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p]
        self.lib.hostsim_snapshot.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int]
        self.lib.hostsim_lap.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
        self.lib.hostsim_destroy.argtypes = [ctypes.c_void_p]
        self.cfg = cfg
        self.h = self.lib.hostsim_create(ctypes.byref(cfg))
        self.lap_steps = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.hostsim_destroy(self.h)
            self.h = None

    def update(self, dets, img=None, embs=None, warp=None):
        if warp is not None:
            w = np.ascontiguousarray(warp, dtype=np.float64).reshape(6)
            self.lib.hostsim_set_warp.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            self.lib.hostsim_set_warp(self.h, w.ctypes.data)
        dets = np.ascontiguousarray(dets, dtype=np.float32).reshape(-1, 6)
        n = len(dets)
        out = np.empty((max(n, 1), 8), np.float32)
        e = None
        if embs is not None:
            e = np.ascontiguousarray(embs, dtype=np.float32)
            assert len(e) == n
        steps = ctypes.c_int(0)
        m = self.lib.hostsim_update(self.h, dets.ctypes.data, n, e.ctypes.data if e is not None else None,
                                    out.ctypes.data, ctypes.byref(steps))
        if m < 0:
            raise RuntimeError(f"hostsim error {-m}")
        self.lap_steps += steps.value
        return out[:m].copy()

    def state_snapshot(self):
        cap = self.cfg.cap_tracks
        ids = np.empty(cap, np.int32)
        means = np.empty((cap, 8))
        covs = np.empty((cap, 8, 8))
        n = self.lib.hostsim_snapshot(self.h, ids.ctypes.data, means.ctypes.data, covs.ctypes.data, cap)
        return {int(ids[i]): (means[i].copy(), covs[i].copy()) for i in range(n)}

    def lap(self, cost, thresh):
        cost = np.ascontiguousarray(cost, dtype=np.float64)
        T, D = cost.shape
        x = np.empty(T, np.int32)
        y = np.empty(D, np.int32)
        assert self.lib.hostsim_lap(self.h, cost.ctypes.data, T, D, thresh, x.ctypes.data, y.ctypes.data) == 0
        return x, y


class DocsCfg(ctypes.Structure):
    _fields_ = [("cap_tracks", ctypes.c_int), ("cap_dets", ctypes.c_int), ("feat_dim", ctypes.c_int),
                ("delta_t", ctypes.c_int), ("max_age", ctypes.c_int), ("min_hits", ctypes.c_int),
                ("embedding_off", ctypes.c_int), ("aw_off", ctypes.c_int), ("det_thresh_f32", ctypes.c_float),
                ("det_thresh", ctypes.c_double), ("iou_threshold", ctypes.c_double), ("inertia", ctypes.c_double),
                ("w_emb", ctypes.c_double), ("alpha_fixed", ctypes.c_double), ("aw_param", ctypes.c_double),
                ("q_xy", ctypes.c_double), ("q_s", ctypes.c_double)]


def deepocsort_cfg(delta_t=3, inertia=0.2, w_association_emb=0.5, alpha_fixed_emb=0.95, aw_param=0.5,
                   embedding_off=False, aw_off=False, Q_xy_scaling=0.01, Q_s_scaling=0.0001, det_thresh=0.3,
                   max_age=30, min_hits=3, iou_threshold=0.3, feat_dim=512, cap_tracks=512, cap_dets=256):
    c = DocsCfg()
    c.cap_tracks, c.cap_dets, c.feat_dim, c.delta_t = cap_tracks, cap_dets, feat_dim, delta_t
    c.max_age, c.min_hits, c.embedding_off, c.aw_off = max_age, min_hits, int(embedding_off), int(aw_off)
    c.det_thresh_f32 = np.float32(det_thresh)
    c.det_thresh, c.iou_threshold, c.inertia, c.w_emb = det_thresh, iou_threshold, inertia, w_association_emb
    c.alpha_fixed, c.aw_param, c.q_xy, c.q_s = alpha_fixed_emb, aw_param, Q_xy_scaling, Q_s_scaling
    return c


class HostSimDeepOcSort:
    def __init__(self, cfg: DocsCfg):
        self.lib = ctypes.CDLL(str(build()))
        assert self.lib.docs_cfg_size() == ctypes.sizeof(DocsCfg)
        self.lib.docs_create.restype = ctypes.c_void_p
        self.lib.docs_create.argtypes = [ctypes.POINTER(DocsCfg)]
        self.lib.docs_update.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        self.lib.docs_snapshot.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int]
        self.lib.docs_destroy.argtypes = [ctypes.c_void_p]
        self.cfg = cfg
        self.h = self.lib.docs_create(ctypes.byref(cfg))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.docs_destroy(self.h)
            self.h = None

    def update(self, dets, img=None, embs=None, warp=None):
        if warp is not None:
            w = np.ascontiguousarray(warp, dtype=np.float64).reshape(6)
            self.lib.docs_set_warp.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            self.lib.docs_set_warp(self.h, w.ctypes.data)
        dets = np.ascontiguousarray(dets, dtype=np.float32).reshape(-1, 6)
        n = len(dets)
        out = np.empty((max(n, 1), 8), np.float32)
        e = None
        if embs is not None:
            e = np.ascontiguousarray(embs, dtype=np.float32).reshape(n, -1) if n else np.zeros((1, self.cfg.feat_dim), np.float32)
        m = self.lib.docs_update(self.h, dets.ctypes.data, n, e.ctypes.data if e is not None else None, out.ctypes.data)
        if m < 0:
            raise RuntimeError(f"hostsim error {-m}")
        return out[:m].copy()

    def set_jv_wide(self, mode: int):
        """Augmentation variant of jv_dense.cuh: 0 one warp, 1 CTA-wide over list positions, 2 CTA-wide with owned
        columns (one "thread" here: same control flow, masks of 1 bit)."""
        self.lib.docs_set_jv_wide.argtypes = [ctypes.c_void_p, ctypes.c_int]
        self.lib.docs_set_jv_wide(self.h, int(mode))

    def jv(self, cost):
        cost = np.ascontiguousarray(cost, dtype=np.float64)
        r, c = cost.shape
        x = np.empty(max(r, 1), np.int32)
        y = np.empty(max(c, 1), np.int32)
        self.lib.docs_jv.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        assert self.lib.docs_jv(self.h, cost.ctypes.data, r, c, x.ctypes.data, y.ctypes.data) == 0
        return x[:r], y[:c]

    def state_snapshot(self):
        cap = self.cfg.cap_tracks
        ids = np.empty(cap, np.int32)
        xs = np.empty((cap, 7))
        ps = np.empty((cap, 7, 7))
        n = self.lib.docs_snapshot(self.h, ids.ctypes.data, xs.ctypes.data, ps.ctypes.data, cap)
        return {int(ids[i]): (xs[i].copy(), ps[i].copy()) for i in range(n)}


class SsCfg(ctypes.Structure):
    _fields_ = [("cap_tracks", ctypes.c_int), ("cap_dets", ctypes.c_int), ("feat_dim", ctypes.c_int),
                ("n_init", ctypes.c_int), ("max_age", ctypes.c_int), ("budget", ctypes.c_int),
                ("min_conf", ctypes.c_double), ("max_cos_dist", ctypes.c_double), ("max_iou_dist", ctypes.c_double),
                ("mc_lambda", ctypes.c_double), ("ema_alpha", ctypes.c_double)]


def strongsort_cfg(min_conf=0.1, max_cos_dist=0.2, max_iou_dist=0.7, n_init=3, nn_budget=100, mc_lambda=0.98,
                   ema_alpha=0.9, max_age=30, feat_dim=96, cap_tracks=256, cap_dets=128):
    c = SsCfg()
    c.cap_tracks, c.cap_dets, c.feat_dim, c.n_init, c.max_age, c.budget = cap_tracks, cap_dets, feat_dim, n_init, max_age, nn_budget
    c.min_conf, c.max_cos_dist, c.max_iou_dist, c.mc_lambda, c.ema_alpha = min_conf, max_cos_dist, max_iou_dist, mc_lambda, ema_alpha
    return c


class HostSimStrongSort:
    def __init__(self, cfg: SsCfg):
        self.lib = ctypes.CDLL(str(build()))
        assert self.lib.ss_cfg_size() == ctypes.sizeof(SsCfg)
        self.lib.ss_create.restype = ctypes.c_void_p
        self.lib.ss_create.argtypes = [ctypes.POINTER(SsCfg)]
        self.lib.ss_update.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        self.lib.ss_snapshot.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int]
        self.lib.ss_set_warp.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        self.lib.ss_lsa.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        self.lib.ss_pyset_difference.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        self.lib.ss_destroy.argtypes = [ctypes.c_void_p]
        self.cfg = cfg
        self.h = self.lib.ss_create(ctypes.byref(cfg))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ss_destroy(self.h)
            self.h = None

    def update(self, dets, img=None, embs=None, warp=None):
        if warp is not None:
            w = np.ascontiguousarray(warp, dtype=np.float64).reshape(6)
            self.lib.ss_set_warp(self.h, w.ctypes.data)
        dets = np.ascontiguousarray(dets, dtype=np.float32).reshape(-1, 6)
        n = len(dets)
        out = np.empty((max(n, 1), 8), np.float32)
        e = np.ascontiguousarray(embs, dtype=np.float32).reshape(n, -1) if n else np.zeros((1, self.cfg.feat_dim), np.float32)
        m = self.lib.ss_update(self.h, dets.ctypes.data, n, e.ctypes.data, out.ctypes.data)
        if m < 0:
            raise RuntimeError(f"hostsim error {-m}")
        return out[:m].copy()

    def lsa(self, cost):
        cost = np.ascontiguousarray(cost, dtype=np.float64)
        r, c = cost.shape
        ri = np.empty(max(min(r, c), 1), np.int32)
        ci = np.empty(max(min(r, c), 1), np.int32)
        n = self.lib.ss_lsa(self.h, cost.ctypes.data, r, c, ri.ctypes.data, ci.ctypes.data)
        assert n >= 0, n
        return ri[:n].copy(), ci[:n].copy()

    def set_difference(self, a, in_b):
        a = np.ascontiguousarray(a, dtype=np.int32)
        f = np.ascontiguousarray(in_b, dtype=np.uint8)
        out = np.empty(max(len(a), 1), np.int32)
        n = self.lib.ss_pyset_difference(self.h, a.ctypes.data, len(a), f.ctypes.data, out.ctypes.data)
        assert n >= 0
        return out[:n].tolist()

    def state_snapshot(self):
        cap = self.cfg.cap_tracks
        ids = np.empty(cap, np.int32)
        means = np.empty((cap, 8))
        covs = np.empty((cap, 8, 8))
        n = self.lib.ss_snapshot(self.h, ids.ctypes.data, means.ctypes.data, covs.ctypes.data, cap)
        return {int(ids[i]): (means[i].copy(), covs[i].copy()) for i in range(n)}


# ---- camera-motion estimation (cmc_ecc.cuh) ------------------------------------------------------------------
def cmc_prepare(img: np.ndarray, scale: float = 0.15) -> np.ndarray:
    """BaseCMC.preprocess through the device source compiled for the host."""
    lib = ctypes.CDLL(str(build()))
    img = np.ascontiguousarray(img, np.uint8)
    rows, cols = img.shape[:2]
    h, w = int(np.rint(rows * scale)), int(np.rint(cols * scale))
    out = np.empty((h, w), np.uint8)
    lib.hostsim_cmc_prepare.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p,
                                        ctypes.c_int, ctypes.c_int]
    lib.hostsim_cmc_prepare(img.ctypes.data, rows, cols, float(scale), out.ctypes.data, h, w)
    return out


def ecc_translation(template: np.ndarray, image: np.ndarray, eps: float = 1e-5, max_iter: int = 100):
    """(status, tx, ty): ecc_translation() of cmc_ecc.cuh on two uint8 registration images."""
    lib = ctypes.CDLL(str(build()))
    t = np.ascontiguousarray(template, np.uint8)
    i = np.ascontiguousarray(image, np.uint8)
    txy = np.zeros(2, np.float32)
    lib.hostsim_ecc.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                ctypes.c_void_p]
    st = lib.hostsim_ecc(t.ctypes.data, i.ctypes.data, t.shape[0], t.shape[1], float(eps), int(max_iter), txy.ctypes.data)
    return st, txy[0], txy[1]
