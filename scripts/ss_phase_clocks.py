"""StrongSORT association phase clocks (ss_frame: 0 dets+camera+predict, 1 appearance stage, 2 set order + IoU stage,
3 updates/births/emit) and the device time of the four association launches, on the bench stream."""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import boxmot_b200 as bb  # noqa: E402
from boxmot_b200 import _lib  # noqa: E402
from boxmot_b200.synthetic import bench_stream  # noqa: E402
from oracle.streams import stress_embeddings  # noqa: E402

lib = _lib.require_device()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
img, dets = bench_stream(n, 160)
embs = stress_embeddings(dets, n, seed=3)
trk = bb.MultiStreamTracker("strongsort", n_streams=1, cap_tracks=1024, cap_dets=n, feat_dim=512, min_conf=0.6, max_cos_dist=0.4)
ph = (ctypes.c_longlong * 16)()
for f in range(120):
    trk.update([dets[f]], None, [embs[f]])
lib.boxmot_b200_tracker_phase_clocks(trk.handle, 0, ph, 1)
for f in range(120, 160):
    trk.update([dets[f]], None, [embs[f]])
lib.boxmot_b200_tracker_phase_clocks(trk.handle, 0, ph, 1)
print({i: ph[i] / 40 for i in range(6)}, trk.last_device_ms())
