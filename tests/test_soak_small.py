"""A small slice of the randomised soak (tests/tools/soak_hostsim.py) in the regular CPU suite: seeded stress streams with
randomised tracker parameters, warps in half of the cases, host simulation of the device source against the oracle.
The long runs (thousands of streams, and oracle vs the unmodified reference) are in DESIGN.md section 1."""
import pytest

from tests.common import assert_rows_match
from tests.tools.soak_hostsim import case_with_warps


@pytest.mark.parametrize("seed", list(range(40, 56)))
def test_random_stream_hostsim_equals_oracle(seed):
    kind, kw, frames, embs, sim, orc, warps = case_with_warps(seed)
    for f, d in enumerate(frames):
        e = None if embs is None else embs[f]
        x = {} if warps is None else {"warp": warps[f]}
        got = sim.update(d, None, e, **x)
        want = orc.update(d, None) if embs is None else orc.update(d, None, e.copy(), **x)
        assert_rows_match(got, want, f, box_rtol=1e-4)
