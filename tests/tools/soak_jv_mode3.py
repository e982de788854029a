"""DeepOCSORT cases of the randomised soak (tests/tools/soak_hostsim.py) with the dense-JV variant forced to mode 3 (the
default: no-op band columns, parallel _find_dense tail, hit list, CTA-wide row reduction), host simulation vs oracle, in
parallel worker processes.   python tests/tools/soak_jv_mode3.py [n_deepocsort_cases=400] [workers=8] [first_seed=0]"""
import os
import sys
from pathlib import Path

os.environ["SOAK_JV_MODE"] = sys.argv[4] if len(sys.argv) > 4 else "3"
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def run(seed):
    from tests.common import assert_rows_match
    from tests.tools.soak_hostsim import case_with_warps

    kind, kw, frames, embs, sim, orc, warps = case_with_warps(seed)
    try:
        for f, d in enumerate(frames):
            x = {} if warps is None else {"warp": warps[f]}
            assert_rows_match(sim.update(d, None, embs[f], **x), orc.update(d, None, embs[f].copy(), **x), f, box_rtol=1e-4)
    except AssertionError as ex:
        return seed, str(ex).splitlines()[0]
    return seed, None


if __name__ == "__main__":
    import multiprocessing as mp

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    seeds = [s for s in range(first, first + 4 * n) if s % 4 == 2]
    from tests.hostsim import build

    build()
    with mp.get_context("spawn").Pool(workers) as pool:
        res = pool.map(run, seeds, chunksize=4)
    bad = [(s, m) for s, m in res if m]
    print(f"{len(seeds)} DeepOCSORT cases in JV mode {os.environ['SOAK_JV_MODE']}: {len(bad)} diverged", bad[:10])
    sys.exit(1 if bad else 0)
