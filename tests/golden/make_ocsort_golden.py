"""Golden vectors for OC-SORT from the UNMODIFIED reference class (boxmot/trackers/bbox/ocsort/ocsort.py), same file
format as make_golden.py (rows per frame, Kalman snapshots of the live tracks at a few frames; ids as emitted, i.e. the
class's 0-based id + 1).   python tests/golden/make_ocsort_golden.py"""
import importlib.util
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
from tests.golden.refharness import install_reference  # noqa: E402

install_reference()
from boxmot.trackers.bbox.ocsort.ocsort import OcSort  # noqa: E402

spec = importlib.util.spec_from_file_location("b200_tests_common", HERE.parent / "common.py")
common = importlib.util.module_from_spec(spec)
spec.loader.exec_module(common)
SNAP_FRAMES = (1, 2, 10, 50, 150, 299)


def run(tracker, frames, img):
    rows, offsets, snaps = [], [0], {}
    for f, dets in enumerate(frames):
        out = np.asarray(tracker.update(dets.copy(), img))
        out = out.reshape(-1, 8) if out.size else np.empty((0, 8), np.float32)
        rows.append(out.astype(np.float32))
        offsets.append(offsets[-1] + len(out))
        if (f + 1) in SNAP_FRAMES:
            ids, means, covs = [], [], []
            for t in tracker.active_tracks:
                ids.append(t.id + 1)
                means.append(np.r_[np.asarray(t.kf.x, dtype=np.float64).reshape(-1), 0.0])
                c = np.zeros((8, 8))
                c[:7, :7] = t.kf.P
                covs.append(c)
            snaps[f"snap{f + 1}_ids"] = np.asarray(ids, dtype=np.int64)
            snaps[f"snap{f + 1}_mean"] = np.asarray(means).reshape(-1, 8)
            snaps[f"snap{f + 1}_cov"] = np.asarray(covs).reshape(-1, 8, 8)
    return dict(rows=np.concatenate(rows, 0), offsets=np.asarray(offsets, np.int64), **snaps)


if __name__ == "__main__":
    img = np.zeros((360, 640, 3), np.uint8)
    for name, (kind, kwargs, make_frames, _) in common.OCSORT_CASES.items():
        out = run(OcSort(**kwargs), make_frames(), img)
        np.savez_compressed(HERE / f"{name}.npz", **out)
        print(name, out["rows"].shape)
