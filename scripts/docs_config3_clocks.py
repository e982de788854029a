"""Phase clocks of the DeepOCSORT frame kernel on the BASELINE config-3 shape (512 detections / frame drawn from 2048
objects in 4 cohorts): SM-clock cycles per phase of the last frame (0 dets+predict, 1 iou/appearance, 2 first
assignment incl. solver, 3 updates, 4 second round, 5 misses+births, 6 emit; solver: 8 column reduction + transfer,
9 row reduction, 10 augmentation (11 _find_dense, 14 hit replays, 15 init + prices + path inside it), 12 free rows entering
augmentation, 13 band columns relaxed, 7 no-op band columns walked over (mode 3)).  BOXMOT_B200_JV_WIDE selects the solver variant."""
import ctypes
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import boxmot_b200 as bb  # noqa: E402
from boxmot_b200 import _lib  # noqa: E402
from tests.test_gpu_deepocsort_scale import cohort_stream  # noqa: E402

lib = _lib.require_device()
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dets, embs = cohort_stream(frames=frames)
gpu = bb.DeepOcSort(cap_tracks=2600, cap_dets=512)
ph = (ctypes.c_longlong * 16)()
for f, (d, e) in enumerate(zip(dets, embs)):
    lib.boxmot_b200_tracker_phase_clocks(gpu._engine.handle, 0, ph, 1)   # reset
    t0 = time.perf_counter()
    gpu.update(d, None, e)
    ms = (time.perf_counter() - t0) * 1e3
    lib.boxmot_b200_tracker_phase_clocks(gpu._engine.handle, 0, ph, 0)
    print(f"frame {f}: {ms:7.1f} ms  clocks(M): " + " ".join(f"{i}:{ph[i] / 1e6:.1f}" for i in range(16) if ph[i]))
