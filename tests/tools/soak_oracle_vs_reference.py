"""Randomised soak of the ORACLE against the UNMODIFIED reference (needs /root/reference; run in the build container):
the same seeded stress streams and randomised parameters as soak_hostsim.py, reference trackers constructed with the same
keyword arguments (CMC off / identity warp, ids from 1 per stream).  A disagreement is a restatement bug in oracle/.
    python tests/tools/soak_oracle_vs_reference.py [n_cases] [first_seed]"""
from __future__ import annotations

import os
import sys
import time

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):   # tiny matrices: BLAS threads only contend
    os.environ.setdefault(_v, "1")
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import refharness  # noqa: E402

refharness.install_reference()
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("b200_soak", ROOT / "tests" / "tools" / "soak_hostsim.py")
soak = importlib.util.module_from_spec(spec)
spec.loader.exec_module(soak)
spec = importlib.util.spec_from_file_location("b200_make_golden", ROOT / "tests" / "golden" / "make_golden.py")
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)

from boxmot.trackers.bbox.botsort.botsort import BotSort  # noqa: E402
from boxmot.trackers.bbox.bytetrack import basetrack as bt_base  # noqa: E402
from boxmot.trackers.bbox.bytetrack.bytetrack import ByteTrack  # noqa: E402
from boxmot.trackers.bbox.deepocsort.deepocsort import DeepOcSort  # noqa: E402
from boxmot.trackers.bbox.strongsort.strongsort import StrongSort  # noqa: E402


def reference_tracker(kind, kw, n_frames, warps):
    """The reference tracker with its camera-motion estimator replaced by the supplied matrices (or switched off)."""
    if kind == "bytetrack":
        bt_base.BaseTrack._count = 0
        return ByteTrack(**kw)
    if kind == "botsort":
        trk = BotSort(reid_model=None, use_cmc=warps is not None, **kw)
        if warps is not None:
            trk.cmc = mg._GivenWarps(warps)
        return trk
    if kind == "deepocsort":
        trk = DeepOcSort(reid_model=None, cmc_off=warps is None, **kw)
        if warps is not None:
            trk.cmc = mg._GivenWarps(warps)
        return trk
    trk = StrongSort(reid_model=None, **kw)
    trk.cmc = mg._FrameWarps(warps if warps is not None else [np.eye(2, 3)] * n_frames)
    return trk


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    img = np.zeros((360, 640, 3), np.uint8)
    bad = 0
    t0 = time.time()
    for seed in range(first, first + n):
        kind, kw, frames, embs, _sim, orc, warps = soak.case_with_warps(seed)
        ref = reference_tracker(kind, kw, len(frames), warps)
        try:
            for f, d in enumerate(frames):
                e = None if embs is None else embs[f]
                if isinstance(getattr(ref, "cmc", None), mg._FrameWarps):
                    ref.cmc.f = f
                x = {} if warps is None else {"warp": warps[f]}
                want = ref.update(d.copy(), img) if e is None else ref.update(d.copy(), img, e.copy())
                got = orc.update(d, None) if e is None else orc.update(d, None, e.copy(), **x)
                want = np.asarray(want, np.float32)
                want = want.reshape(-1, 8) if want.size else np.empty((0, 8), np.float32)
                soak.assert_rows_match(got, want, f, exact_boxes=True)
        except AssertionError as ex:
            bad += 1
            print(f"seed {seed} {kind} warps={warps is not None} DIVERGED: {str(ex).splitlines()[0]}  kw={kw}")
        if (seed - first + 1) % 50 == 0:
            print(f"  ... {seed - first + 1} cases, {bad} diverged, {time.time() - t0:.0f} s", flush=True)
    print(f"{n} cases, {bad} diverged, {time.time() - t0:.0f} s")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
