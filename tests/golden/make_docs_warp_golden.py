"""DeepOCSORT with a SUPPLIED camera-motion warp per frame: the reference tracker is built with
`cmc_off=False` and its estimator replaced by one that returns the given 2x3 matrices, so `apply_affine_correction`
(deepocsort.py:189-206, xysr.py:311-366) runs on every track every frame.  Writes tests/golden/deepocsort_warp_stress64.npz.
Run: python tests/golden/make_docs_warp_golden.py"""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1]))
import refharness  # noqa: E402


def main():
    refharness.install_reference()
    spec = importlib.util.spec_from_file_location("b200_tests_common", HERE.parent / "common.py")
    common = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(common)
    spec = importlib.util.spec_from_file_location("b200_make_golden", HERE / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    from boxmot.trackers.bbox.deepocsort.deepocsort import DeepOcSort

    name = "deepocsort_warp_stress64"
    kind, kwargs, make_frames, make_embs = common.CASES[name]
    frames = make_frames()
    embs = make_embs(frames)
    trk = DeepOcSort(reid_model=None, cmc_off=False, **kwargs)
    trk.cmc = mg._GivenWarps(common.WARPS[name]())
    res = mg.run(trk, frames, np.zeros((360, 640, 3), np.uint8), embs)
    np.savez_compressed(HERE / f"{name}.npz", **res)
    print(name, "frames", len(frames), "rows", len(res["rows"]), "ids", len(np.unique(res["rows"][:, 4])))


if __name__ == "__main__":
    main()
