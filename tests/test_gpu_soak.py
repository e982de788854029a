"""GPU soak: the randomised stress cases of tests/tools/soak_hostsim.py (seeded streams, randomised tracker parameters,
supplied camera-motion warps in half of the BoT-SORT / DeepOCSORT / StrongSORT cases, windows of the MOT17-mini public
detections in every third block) replayed on the DEVICE, >= 500 streams per tracker, against digests of the oracle's
[id, conf, cls, det_ind] output (tests/golden/soak_digests.json, made on the host by tests/golden/make_soak_digests.py).
The only tolerated divergence is the documented StrongSORT birth-order swap (ids permuted, everything else equal)."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DIGESTS = Path(__file__).parent / "golden" / "soak_digests.json"


def _device_digest(seed):
    import boxmot_b200 as bb
    from tests.golden.make_soak_digests import frame_digest
    from tests.tools.soak_hostsim import case_with_warps

    kind, kw, frames, embs, _, _, warps = case_with_warps(seed, build=False)
    cls = {"bytetrack": bb.ByteTrack, "botsort": bb.BotSort, "deepocsort": bb.DeepOcSort, "strongsort": bb.StrongSort}[kind]
    extra = {} if kind == "bytetrack" else dict(feat_dim=64)
    trk = cls(cap_tracks=512, cap_dets=256, **extra, **kw)
    h = hashlib.sha1()
    for f, d in enumerate(frames):
        if kind == "bytetrack":
            out = trk.update(d, None)
        else:
            x = {} if warps is None else {"warp": warps[f]}
            out = trk.update(d, None, embs[f], **x)
        frame_digest(h, out)
    return kind, warps is not None, h.hexdigest()


@pytest.mark.skipif(not DIGESTS.exists(), reason="soak digests not generated")
def test_device_soak_against_oracle_digests():
    want = json.loads(DIGESTS.read_text())["cases"]
    seeds = sorted(int(s) for s in want)
    per_kind, bad = {}, []
    for seed in seeds:
        kind, warped, got = _device_digest(seed)
        per_kind[kind] = per_kind.get(kind, 0) + 1
        if got != want[str(seed)]["sha1"]:
            bad.append((seed, kind, warped))
    print("soak:", per_kind, "diverged:", bad)
    assert min(per_kind.values()) >= 500 or len(seeds) < 2000
    # Tolerated: StrongSORT streams whose output equals the oracle's up to a consistent renaming of ids -- the order in
    # which two simultaneously born tracks get their ids follows scipy's tie-break among fully gated entries, which runs
    # through the last bits of LAPACK's Cholesky solve (the reference's own choice depends on its BLAS build).
    from tests.tools.soak_hostsim import same_up_to_an_id_permutation

    def make_device(kind, kw):
        import boxmot_b200 as bb

        return bb.StrongSort(cap_tracks=512, cap_dets=256, feat_dim=64, **kw)

    real = [b for b in bad if b[1] != "strongsort" or not same_up_to_an_id_permutation(b[0], make_device)]
    assert not real, f"device output differs from the oracle on {real}"
    assert len(bad) <= max(3, len(seeds) // 300), f"more StrongSORT birth-order swaps than documented: {bad}"
