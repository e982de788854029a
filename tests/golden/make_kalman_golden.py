"""Known answers for the Kalman filters of the path, from the UNMODIFIED reference classes
(`boxmot/motion/kalman_filters/{xywh,xyah,xysr}.py`) on
  (a) the exact inputs of the reference's own unit tests (`tests/unit/test_kalman_filters_modes.py:15-40, 77-93`:
      initiate -> predict -> update(confidence=0.9), which hold no numeric expectations), and
  (b) a seeded batch of boxes: initiate, three predict/update rounds (NSA confidences included), multi_predict.
Written to tests/golden/kalman_reference.npz;  run: python tests/golden/make_kalman_golden.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1]))

import refharness  # noqa: E402

REF_CASES = {
    "xywh": (np.array([100.0, 80.0, 40.0, 20.0]), np.array([101.0, 79.5, 40.5, 20.5])),
    "xyah": (np.array([100.0, 80.0, 1.6, 60.0]), np.array([100.5, 80.2, 1.58, 60.1])),
}
UNFREEZE_OBS = [np.array([[300.0], [200.0], [50000.0], [1.5]]), np.array([[320.0], [210.0], [51000.0], [1.4]]),
                np.array([[350.0], [230.0], [52000.0], [1.3]])]
XYSR_CASE = (np.array([[300.0], [200.0], [50000.0], [1.5]]), np.array([[305.0], [202.0], [50500.0], [1.45]]))


def seeded_batch(kind, n=24, seed=3):
    rng = np.random.default_rng(seed)
    cx, cy = rng.uniform(50, 1800, n), rng.uniform(50, 1000, n)
    w, h = rng.uniform(15, 300, n), rng.uniform(30, 600, n)
    z0 = np.stack([cx, cy, w / h, h], 1) if kind == "xyah" else np.stack([cx, cy, w, h], 1)
    steps = []
    for _ in range(3):
        dz = rng.normal(0, 1.5, (n, 4))
        if kind == "xyah":
            dz[:, 2] *= 0.01
        steps.append((dz, rng.uniform(0.0, 0.98, n)))
    return z0, steps


def main():
    refharness.install_reference()
    from boxmot.motion.kalman_filters.xyah import KalmanFilterXYAH
    from boxmot.motion.kalman_filters.xysr import KalmanFilterXYSR
    from boxmot.motion.kalman_filters.xywh import KalmanFilterXYWH

    out = {}
    for kind, cls in (("xywh", KalmanFilterXYWH), ("xyah", KalmanFilterXYAH)):
        kf = cls(ndim=4)
        z0, z1 = REF_CASES[kind]
        m, c = kf.initiate(z0)
        out[f"{kind}_ref_init_mean"], out[f"{kind}_ref_init_cov"] = m.copy(), c.copy()
        m, c = kf.predict(m, c)
        out[f"{kind}_ref_pred_mean"], out[f"{kind}_ref_pred_cov"] = m.copy(), c.copy()
        m2, c2 = kf.update(m, c, z1, confidence=0.9)
        out[f"{kind}_ref_upd_mean"], out[f"{kind}_ref_upd_cov"] = m2.copy(), c2.copy()
        out[f"{kind}_ref_gate"] = kf.gating_distance(m2, c2, z1[None, :])
        # seeded batch
        z0b, steps = seeded_batch(kind)
        means, covs = zip(*[kf.initiate(z) for z in z0b])
        means, covs = np.asarray(means), np.asarray(covs)
        out[f"{kind}_b_init_mean"], out[f"{kind}_b_init_cov"] = means.copy(), covs.copy()
        z = z0b.copy()
        for r, (dz, conf) in enumerate(steps):
            means, covs = kf.multi_predict(means, covs)
            out[f"{kind}_b_pred{r}_mean"], out[f"{kind}_b_pred{r}_cov"] = means.copy(), covs.copy()
            z = z + dz
            upd = [kf.update(means[i], covs[i], z[i], confidence=float(conf[i]) if r == 2 else 0.0) for i in range(len(z))]
            means, covs = np.asarray([u[0] for u in upd]), np.asarray([u[1] for u in upd])
            out[f"{kind}_b_upd{r}_mean"], out[f"{kind}_b_upd{r}_cov"] = means.copy(), covs.copy()
    kf = KalmanFilterXYSR(dim_x=7, dim_z=4, max_obs=50)
    z0, z1 = XYSR_CASE
    m, c = kf.initiate(z0)
    out["xysr_ref_init_x"], out["xysr_ref_init_P"] = m.copy(), c.copy()
    kf.x, kf.P = m.copy(), c.copy()
    out["xysr_ref_Q"], out["xysr_ref_R"], out["xysr_ref_F"], out["xysr_ref_H"] = kf.Q.copy(), kf.R.copy(), kf.F.copy(), kf.H.copy()
    kf.predict()
    out["xysr_ref_pred_x"], out["xysr_ref_pred_P"] = kf.x.copy(), kf.P.copy()
    kf.update(z1)
    out["xysr_ref_upd_x"], out["xysr_ref_upd_P"] = kf.x.copy(), kf.P.copy()
    # the reference's un-freeze regression scenario (test_kalman_filters_modes.py:164-193, issue #2207): two observations,
    # five missed frames, a third observation -> freeze, virtual-trajectory replay, update
    kf = KalmanFilterXYSR(dim_x=7, dim_z=4, max_obs=50)
    kf.F = np.eye(7)
    kf.F[:4, 4:] = np.pad(np.eye(3), ((0, 1), (0, 0)))
    kf.H = np.zeros((4, 7))
    kf.H[:4, :4] = np.eye(4)
    kf.R *= 10.0
    out["unfreeze_x0"], out["unfreeze_P0"] = kf.x.copy(), kf.P.copy()
    out["unfreeze_Q"], out["unfreeze_R"] = kf.Q.copy(), kf.R.copy()
    for k, obs in enumerate(UNFREEZE_OBS[:2]):
        kf.predict()
        kf.update(obs)
    for _ in range(5):
        kf.predict()
        kf.update(None)
    out["unfreeze_gap_x"], out["unfreeze_gap_P"] = kf.x.copy(), kf.P.copy()
    kf.predict()
    kf.update(UNFREEZE_OBS[2])
    out["unfreeze_x"], out["unfreeze_P"] = kf.x.copy(), kf.P.copy()
    np.savez_compressed(HERE / "kalman_reference.npz", **out)
    print("wrote kalman_reference.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
