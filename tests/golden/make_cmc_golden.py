"""Golden vectors for the camera-motion estimator, dumped from the UNMODIFIED reference class
boxmot.motion.cmc.ecc.ECC (tests/golden/refharness.py imports /root/reference; OpenCV is the installed cv2):

* `pan_warps`   -- ECC().apply on every frame of boxmot_b200.synthetic.camera_pan_sequence(12) (frames are regenerated
                   by the tests from the same seed);
* `mot17_reg`   -- BaseCMC.preprocess of the first frames of the reference's own assets/MOT17-mini sequences
                   (162 x 288 uint8 registration images; the 1080p frames themselves are not committed) and
  `mot17_warps` -- what ECC().apply returned on them, per sequence.

Run here (needs /root/reference):  python tests/golden/make_cmc_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from tests.golden.refharness import REFERENCE_ROOT, install_reference  # noqa: E402

install_reference()
import cv2  # noqa: E402
from boxmot.motion.cmc.ecc import ECC  # noqa: E402

from boxmot_b200.synthetic import camera_pan_sequence  # noqa: E402

frames, _, offs, _ = camera_pan_sequence(12)
ecc = ECC()
pan = np.stack([ecc.apply(f) for f in frames])
out = dict(pan_warps=pan.astype(np.float32), pan_offsets=offs)
regs, warps = [], []
for seq in ("MOT17-02-FRCNN", "MOT17-04-FRCNN"):
    files = sorted((REFERENCE_ROOT / "assets" / "MOT17-mini" / "train" / seq / "img1").glob("*.jpg"))[:5]
    ecc = ECC()
    for f in files:
        img = cv2.imread(str(f))
        warps.append(ecc.apply(img))
        regs.append(ecc.prev_img.copy())
out["mot17_reg"] = np.stack(regs)
out["mot17_warps"] = np.stack(warps).astype(np.float32)
np.savez_compressed(ROOT / "tests" / "golden" / "cmc_ecc.npz", **out)
print({k: v.shape for k, v in out.items()})
print(pan[:, :, 2])
