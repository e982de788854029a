"""Digests of the ORACLE's output on the randomised soak cases (tests/tools/soak_hostsim.py::case_with_warps): per case a
sha1 over, frame by frame, the row count and the bit patterns of the [id, conf, cls, det_ind] columns -- the part of the
output the parity bar requires bit-exact.  tests/test_gpu_soak.py replays the same cases on the device and compares.
The oracle itself is pinned to the unmodified reference on these generators (2 934 streams).

    python tests/golden/make_soak_digests.py [n_cases=2000] [workers=8]   ->  tests/golden/soak_digests.json
"""
import hashlib
import json
import os
import sys
from pathlib import Path

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")
import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def frame_digest(h, rows):
    r = np.ascontiguousarray(np.asarray(rows, np.float32).reshape(-1, 8)[:, 4:8])
    h.update(np.int32(len(r)).tobytes())
    h.update(r.tobytes())


def oracle_case(seed):
    from tests.tools.soak_hostsim import case_with_warps

    kind, kw, frames, embs, sim, orc, warps = case_with_warps(seed)
    h = hashlib.sha1()
    for f, d in enumerate(frames):
        e = None if embs is None else embs[f]
        x = {} if warps is None else {"warp": warps[f]}
        want = orc.update(d, None) if embs is None else orc.update(d, None, e.copy(), **x)
        frame_digest(h, want)
    return seed, kind, len(frames), warps is not None, h.hexdigest()


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    import multiprocessing as mp

    with mp.get_context("spawn").Pool(workers) as pool:
        res = pool.map(oracle_case, range(n), chunksize=8)
    out = {"n_cases": n, "generator": "tests/tools/soak_hostsim.py::case_with_warps", "columns": "id, conf, cls, det_ind (float32 bits) + row count",
           "cases": {str(s): {"kind": k, "frames": nf, "warps": w, "sha1": d} for s, k, nf, w, d in res}}
    (Path(__file__).parent / "soak_digests.json").write_text(json.dumps(out))
    print("wrote", n, "digests")
