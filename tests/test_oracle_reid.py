"""Pin the ReID oracle (oracle/reid.py): the fixed-point resize against cv2.resize bit for bit, crops and the
functional OSNet against goldens dumped from the reference classes, and the BN-folded blob (weights.py)
against the oracle forward."""
import numpy as np
import pytest
import torch

from oracle import reid as orid
from tests.common import GOLDEN


@pytest.mark.parametrize("seed", range(6))
def test_resize_matches_cv2_bit_exact(seed):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(seed)
    sizes = [(int(rng.integers(1, 500)), int(rng.integers(1, 300))) for _ in range(25)] + [(256, 128), (512, 256), (1, 1), (2, 700)]
    for h, w in sizes:
        src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(cv2.resize(src, (128, 256), interpolation=cv2.INTER_LINEAR),
                              orid.resize_linear_u8(src, 256, 128)), (h, w)


def _golden():
    z = np.load(GOLDEN / "reid_osnet_x0_25.npz")
    return z


def test_crops_match_reference_golden():
    z = _golden()
    rng = np.random.default_rng(int(z["image_seed"]))
    img = rng.integers(0, 255, size=(720, 1280, 3), dtype=np.uint8)
    import hashlib

    crops = orid.get_crops(z["boxes"], img).numpy()
    assert np.array_equal(crops[z["crops_sample_index"]], z["crops_sample"])
    assert hashlib.sha256(np.ascontiguousarray(crops).tobytes()).hexdigest() == str(z["crops_sha256"]), \
        "staged crops must be bit-exact"


def test_osnet_matches_reference_golden():
    z = _golden()
    sd = orid.make_osnet_state("osnet_x0_25", seed=int(z["weight_seed"]))
    rng = np.random.default_rng(int(z["image_seed"]))
    img = rng.integers(0, 255, size=(720, 1280, 3), dtype=np.uint8)
    feats = orid.get_features(sd, z["boxes"], img)
    # same torch build -> identical; other BLAS/oneDNN builds may differ in the last ulps
    np.testing.assert_allclose(feats, z["features"], rtol=0, atol=2e-6)
    assert abs(np.linalg.norm(feats, axis=1) - 1).max() < 1e-6


@pytest.mark.parametrize("arch", ["osnet_x0_25", "osnet_x1_0"])
def test_folded_blob_equals_oracle(arch, tmp_path):
    from boxmot_b200.weights import export_blob
    from tests.blobsim import blob_forward

    sd = orid.make_osnet_state(arch, seed=3)
    blob = export_blob(sd, tmp_path / f"{arch}.b200reid")
    x = torch.randn(3, 3, 256, 128)
    want, st_w = orid.osnet_forward(sd, x, return_stages=True)
    got, st_g = blob_forward(blob, x.permute(0, 2, 3, 1).contiguous(), return_stages=True)
    for k in ("pool", "conv2.0", "conv2.2", "conv3.1", "conv4.1"):
        np.testing.assert_allclose(st_g[k].permute(0, 3, 1, 2).numpy(), st_w[k].numpy(), rtol=1e-4, atol=2e-5)
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) < 2e-6 * max(scale, 1.0)


def test_pt_roundtrip(tmp_path):
    from boxmot_b200.weights import export_blob, read_blob

    sd = orid.make_osnet_state("osnet_x0_25", seed=4)
    torch.save({"state_dict": {"module." + k: v for k, v in sd.items()}}, tmp_path / "osnet_x0_25_msmt17.pt")
    blob = export_blob(tmp_path / "osnet_x0_25_msmt17.pt")
    header, payload = read_blob(blob)
    assert header[3:8] == (16, 64, 96, 128, 512) and payload.size == header[8]


def test_mobilenetv2_folded_blob_equals_oracle(tmp_path):
    from boxmot_b200.synthetic import make_mobilenetv2_state
    from boxmot_b200.weights import export_blob, read_blob, read_block_table
    from tests.blobsim import blob_forward

    sd = make_mobilenetv2_state(1.4, seed=2)
    blob = export_blob(sd, tmp_path / "mobilenetv2_x1_4.b200reid")
    header, _ = read_blob(blob)
    assert header[2] == 2 and header[3] == 44 and header[7] == 1792 and len(read_block_table(blob)) == 17
    x = torch.randn(2, 3, 256, 128)
    want = orid.mobilenetv2_forward(sd, x)
    got = blob_forward(blob, x.permute(0, 2, 3, 1).contiguous())
    assert float((got - want).abs().max()) < 5e-6 * float(want.abs().max())


@pytest.mark.parametrize("arch", ["osnet_x1_0", "mobilenetv2_x1_4"])
def test_other_architectures_match_the_reference_classes(arch):
    """`osnet_forward` at x1_0 width and `mobilenetv2_forward` against embeddings the reference's own model classes
    produced through `BaseModelBackend.get_features` (tests/golden/make_reid_arch_golden.py) -- these functional
    restatements are the oracle of the OSNet_x1_0 / MobileNetV2_x1_4 GPU tests."""
    from boxmot_b200.synthetic import make_mobilenetv2_state, make_osnet_state
    from tests.golden.make_reid_arch_golden import IMAGE_SEED, MBV2_SEED, OSNET_SEED

    z = np.load(GOLDEN / "reid_arch_reference.npz")
    img = np.random.default_rng(IMAGE_SEED).integers(0, 255, size=(540, 960, 3), dtype=np.uint8)
    sd = make_osnet_state("osnet_x1_0", seed=OSNET_SEED) if arch == "osnet_x1_0" else make_mobilenetv2_state(1.4, seed=MBV2_SEED)
    feats = orid.get_features(sd, z["boxes"], img)
    assert feats.shape == z[arch].shape and feats.shape[1] == (512 if arch == "osnet_x1_0" else 1792)
    np.testing.assert_allclose(feats, z[arch], rtol=0, atol=2e-6)
    assert abs(np.linalg.norm(feats, axis=1) - 1).max() < 1e-6


@pytest.mark.parametrize("seed", range(6))
def test_resize_pad_matches_the_reference_function(seed):
    """oracle.reid.resize_pad_u8 against reid/core/preprocessing.py:21-45 restated with the installed OpenCV
    (cv2.resize INTER_LINEAR + cv2.copyMakeBorder with IMAGENET_MEAN_BGR): bit for bit over random crop sizes."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(100 + seed)
    for _ in range(40):
        h, w = int(rng.integers(2, 400)), int(rng.integers(2, 300))
        crop = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        th, tw = 256, 128
        scale = min(tw / w, th / h)
        nw, nh = int(w * scale), int(h * scale)
        if nw < 1 or nh < 1:
            continue
        resized = cv2.resize(crop, (nw, nh), interpolation=cv2.INTER_LINEAR)
        top, left = (th - nh) // 2, (tw - nw) // 2
        want = cv2.copyMakeBorder(resized, top, th - nh - top, left, tw - nw - left, cv2.BORDER_CONSTANT, value=(104, 116, 124))
        assert np.array_equal(orid.resize_pad_u8(crop, th, tw), want)
