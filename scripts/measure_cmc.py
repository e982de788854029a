"""Device time of the on-device ECC camera-motion estimator next to OpenCV's on the host cores:
BoT-SORT (no ReID) on the synthetic panning-camera sequence with use_cmc on / off, per resolution.  Prints ms per frame of
update() (host inputs, synchronous) and the difference = frame H2D + gray/resize + ECC iterations."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import cv2  # noqa: E402

import boxmot_b200 as bb  # noqa: E402
from boxmot_b200.synthetic import camera_pan_sequence  # noqa: E402

CRIT = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 100, 1e-5)
for hw in ((720, 1280), (1080, 1920)):
    frames, dets, _, _ = camera_pan_sequence(40, hw=hw, seed=21)
    res = {}
    for cmc in (False, True):
        trk = bb.BotSort(use_cmc=cmc, cmc_method="ecc", with_reid=False, cap_tracks=128, cap_dets=64)
        for f in range(8):
            trk.update(dets[f], frames[f])
        t0 = time.perf_counter()
        for f in range(8, 40):
            trk.update(dets[f], frames[f])
        res[cmc] = (time.perf_counter() - t0) / 32 * 1e3
    prev, t_cv, iters = None, 0.0, 0
    for f in range(8, 40):
        t0 = time.perf_counter()
        cur = cv2.resize(cv2.cvtColor(frames[f], cv2.COLOR_BGR2GRAY), (0, 0), fx=0.15, fy=0.15, interpolation=cv2.INTER_LINEAR)
        if prev is not None:
            cv2.findTransformECC(prev, cur, np.eye(2, 3, dtype=np.float32), cv2.MOTION_TRANSLATION, CRIT, None, 1)
        prev = cur
        t_cv += time.perf_counter() - t0
    print(f"{hw[1]}x{hw[0]}: update() {res[False]:.3f} ms without CMC, {res[True]:.3f} ms with on-device ECC "
          f"(+{res[True] - res[False]:.3f} ms incl. the frame upload); OpenCV preprocess + findTransformECC on the host: "
          f"{t_cv / 32 * 1e3:.3f} ms/frame")
