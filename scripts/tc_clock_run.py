import sys, tempfile
from pathlib import Path
import numpy as np
sys.path.insert(0, '.')
from oracle import reid as orid
from boxmot_b200.reid import B200ReID
from boxmot_b200.weights import export_blob
sd = orid.make_osnet_state("osnet_x0_25", seed=11)
reid = B200ReID(export_blob(sd, Path(tempfile.mkdtemp()) / "m.b200reid"))
rng = np.random.default_rng(0)
img = rng.integers(0, 255, size=(720, 1280, 3), dtype=np.uint8)
n = 208
cx, cy = rng.uniform(0, 1280, n), rng.uniform(0, 720, n)
w, h = rng.uniform(20, 120, n), rng.uniform(40, 240, n)
boxes = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(np.float32)
f = reid.get_features(boxes, img)
print("----- second pass (warm) -----", flush=True)
f = reid.get_features(boxes, img)
