"""Per-tracker end-to-end frame rates on the bench stream (1 stream x N detections, OSNet_x0_25 on the device,
host numpy inputs through MultiStreamTracker.update) -- the four frontends side by side.  Not the bench contract
(bench.py is); used for DESIGN.md section 6."""
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import boxmot_b200 as bb  # noqa: E402
from boxmot_b200.synthetic import bench_stream, make_osnet_state  # noqa: E402
from boxmot_b200.trackers import TRACKER_DEFAULTS  # noqa: E402
from boxmot_b200.weights import export_blob  # noqa: E402


def main():
    n_dets = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 120
    img, dets = bench_stream(n_dets, frames + 20)
    blob = export_blob(make_osnet_state("osnet_x0_25", seed=0), Path(tempfile.mkdtemp()) / "m.b200reid")
    out = {}
    for kind in ("bytetrack", "botsort", "deepocsort", "strongsort"):
        p = dict(TRACKER_DEFAULTS[kind])
        for k in ("use_cmc", "cmc_method", "cmc_off"):
            p.pop(k, None)
        trk = bb.MultiStreamTracker(kind, n_streams=1, cap_tracks=1024, cap_dets=n_dets, feat_dim=512,
                                    reid_blob=None if kind == "bytetrack" else str(blob), **p)
        for f in range(20):
            trk.update([dets[f]], [img])
        t0 = time.perf_counter()
        rows = 0
        for f in range(20, 20 + frames):
            rows += len(trk.update([dets[f]], [img])[0])
        dt = time.perf_counter() - t0
        out[kind] = {"frames_per_s": frames / dt, "ms_per_frame": 1e3 * dt / frames, "rows_per_frame": rows / frames,
                     "launches": trk.last_launches(), "device_ms_reid_assoc": trk.last_device_ms()}
        trk.close()
    print(json.dumps({"dets_per_frame": n_dets, "frames": frames, "trackers": out}))


if __name__ == "__main__":
    main()
