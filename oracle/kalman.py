"""Oracle restatement of the stateless box Kalman filters -- TEST INFRASTRUCTURE ONLY.

Follows (file:line relative to /root/reference/boxmot):
  * motion/kalman_filters/base.py:234-244  initiate
  * motion/kalman_filters/base.py:311-327  multi_predict  (process noise from the PRIOR mean)
  * motion/kalman_filters/base.py:286-309  project        (NSA scaling (1 - confidence))
  * motion/kalman_filters/base.py:329-355  update         (cho_factor / cho_solve, non-Joseph form)
  * motion/kalman_filters/xyah.py:22-88,112-120   XYAH std tables + clamp a,h >= 1e-4
  * motion/kalman_filters/xywh.py:22-85,149-160   XYWH std tables + clamp w,h >= 1e-4
Only the AABB (ndim=4) filters are restated; OBB is out of scope.
State is float64, exactly as the reference keeps it; the same numpy/scipy calls are used so that on one
machine the oracle and the reference agree to the last bit.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

W_POS = 1.0 / 20
W_VEL = 1.0 / 160

_F = np.eye(8)
for _i in range(4):
    _F[_i, 4 + _i] = 1.0
_H = np.eye(4, 8)


def _sizes(kind: str, vec):
    """Per-coordinate scale used by the std tables: XYAH uses h everywhere, XYWH alternates w,h."""
    if kind == "xyah":
        return vec[..., 3], vec[..., 3], vec[..., 3], vec[..., 3]
    if kind == "xywh":
        return vec[..., 2], vec[..., 3], vec[..., 2], vec[..., 3]
    raise ValueError(kind)


def initiate(kind: str, measurement):
    m = np.asarray(measurement, dtype=float).copy()
    mean = np.r_[m, np.zeros_like(m)]
    s = _sizes(kind, m)
    if kind == "xyah":
        std = [2 * W_POS * s[0], 2 * W_POS * s[1], 1e-2, 2 * W_POS * s[3],
               10 * W_VEL * s[0], 10 * W_VEL * s[1], 1e-5, 10 * W_VEL * s[3]]
    else:
        std = [2 * W_POS * s[0], 2 * W_POS * s[1], 2 * W_POS * s[2], 2 * W_POS * s[3],
               10 * W_VEL * s[0], 10 * W_VEL * s[1], 10 * W_VEL * s[2], 10 * W_VEL * s[3]]
    cov = np.diag(np.square(std))
    mean[2] = max(float(mean[2]), 1e-4)
    mean[3] = max(float(mean[3]), 1e-4)
    return mean, cov


def multi_predict(kind: str, mean: np.ndarray, cov: np.ndarray):
    """mean (T,8), cov (T,8,8) -> predicted copies."""
    if len(mean) == 0:
        return mean, cov
    s = _sizes(kind, mean)
    if kind == "xyah":
        std_pos = [W_POS * s[0], W_POS * s[1], 1e-2 * np.ones_like(s[0]), W_POS * s[3]]
        std_vel = [W_VEL * s[0], W_VEL * s[1], 1e-5 * np.ones_like(s[0]), W_VEL * s[3]]
    else:
        std_pos = [W_POS * s[0], W_POS * s[1], W_POS * s[2], W_POS * s[3]]
        std_vel = [W_VEL * s[0], W_VEL * s[1], W_VEL * s[2], W_VEL * s[3]]
    sqr = np.square(np.r_[std_pos, std_vel]).T
    motion_cov = np.asarray([np.diag(sqr[i]) for i in range(len(mean))])
    mean = np.dot(mean, _F.T)
    left = np.dot(_F, cov).transpose((1, 0, 2))
    cov = np.dot(left, _F.T) + motion_cov
    mean[:, 2] = np.maximum(mean[:, 2], 1e-4)
    mean[:, 3] = np.maximum(mean[:, 3], 1e-4)
    return mean, cov


def project(kind: str, mean, cov, confidence: float = 0.0):
    s = _sizes(kind, mean)
    if kind == "xyah":
        std = [W_POS * s[0], W_POS * s[1], 1e-1, W_POS * s[3]]
    else:
        std = [W_POS * s[0], W_POS * s[1], W_POS * s[2], W_POS * s[3]]
    std = [(1 - confidence) * x for x in std]
    innovation_cov = np.diag(np.square(std))
    pm = np.dot(_H, mean)
    pc = np.linalg.multi_dot((_H, cov, _H.T))
    return pm, pc + innovation_cov


def update(kind: str, mean, cov, measurement, confidence: float = 0.0):
    pm, pc = project(kind, mean, cov, confidence)
    chol, lower = scipy.linalg.cho_factor(pc, lower=True, check_finite=False)
    gain = scipy.linalg.cho_solve((chol, lower), np.dot(cov, _H.T).T, check_finite=False).T
    innovation = measurement - pm
    new_mean = mean + np.dot(innovation, gain.T)
    new_cov = cov - np.linalg.multi_dot((gain, pc, gain.T))
    new_mean[2] = max(float(new_mean[2]), 1e-4)
    new_mean[3] = max(float(new_mean[3]), 1e-4)
    return new_mean, new_cov
