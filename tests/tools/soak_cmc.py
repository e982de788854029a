"""Randomised soak of the ECC estimator: the device source compiled for the host (cmc_ecc.cuh via tests/_hostsim) and the
oracle restatement against the installed OpenCV on random panning-camera pairs -- random frame sizes, pan steps up to 30 px,
sensor noise, occasional unrelated / inverted second frames (the StsNoConv exits).
    python tests/tools/soak_cmc.py [n_pairs=300]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import cv2  # noqa: E402

from boxmot_b200.synthetic import camera_pan_sequence  # noqa: E402
from oracle import cmc  # noqa: E402
from tests import hostsim as hs  # noqa: E402

CRIT = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 100, 1e-5)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(2024)
worst, noconv, bad = 0.0, 0, []
for k in range(n):
    hw = (int(rng.integers(200, 760)), int(rng.integers(300, 1300)))
    step = float(rng.choice([1.5, 4.0, 10.0, 30.0]))
    frames, _, _, _ = camera_pan_sequence(2, hw=hw, seed=int(rng.integers(1, 10**6)), max_step=step, n_objects=1)
    a, b = frames
    mode = int(rng.integers(0, 12))
    if mode == 0:
        b = 255 - b
    elif mode == 1:
        b = camera_pan_sequence(1, hw=hw, seed=int(rng.integers(1, 10**6)), n_objects=1)[0][0]
    p0, p1 = cmc.preprocess(a), cmc.preprocess(b)
    assert np.array_equal(hs.cmc_prepare(a), cv2.resize(cv2.cvtColor(a, cv2.COLOR_BGR2GRAY), (0, 0), fx=0.15, fy=0.15,
                                                       interpolation=cv2.INTER_LINEAR))
    try:
        _, w = cv2.findTransformECC(p0, p1, np.eye(2, 3, dtype=np.float32), cv2.MOTION_TRANSLATION, CRIT, None, 1)
        want = (0, float(w[0, 2]), float(w[1, 2]))
    except cv2.error:
        want = (1, 0.0, 0.0)
    st, hx, hy = hs.ecc_translation(p0, p1)
    try:
        _, ox, oy = cmc.find_transform_ecc_translation(p0, p1)
        ost = 0
    except cmc.NoConvergence:
        ost, ox, oy = 1, 0.0, 0.0
    noconv += want[0]
    if st != want[0] or ost != want[0]:
        bad.append((k, hw, mode, want, (st, float(hx), float(hy)), (ost, float(ox), float(oy))))
        continue
    err = max(abs(hx - want[1]), abs(hy - want[2]), abs(ox - want[1]), abs(oy - want[2])) / max(1.0, abs(want[1]), abs(want[2]))
    worst = max(worst, err)
    if err > 1e-4:
        bad.append((k, hw, mode, want, (st, float(hx), float(hy)), (ost, float(ox), float(oy))))
print(f"{n} pairs, {noconv} where OpenCV raised StsNoConv, worst relative warp error {worst:.2e}, disagreements: {len(bad)}", bad[:5])
sys.exit(1 if bad else 0)
