"""Embeddings of the reference's OSNet_x1_0 and MobileNetV2_x1_4 classes (`reid/backbones/osnet.py:488`,
`reid/backbones/mobilenetv2.py:233`) with seeded weights loaded by `load_state_dict(strict=True)`, on a handful of
boxes of a seeded frame, through the reference backend's own `get_features` (crop, normalise, forward, L2 norm).
Pins `oracle.reid.osnet_forward` / `mobilenetv2_forward` for the architectures the x0_25 golden does not cover.
Writes tests/golden/reid_arch_reference.npz.   Run: python tests/golden/make_reid_arch_golden.py"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1]))
import refharness  # noqa: E402

IMAGE_SEED, OSNET_SEED, MBV2_SEED = 321, 11, 12


def boxes_for(n=6, seed=9):
    r = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        cx, cy = r.uniform(0, 960), r.uniform(0, 540)
        w, h = r.uniform(12, 140), r.uniform(24, 300)
        out.append([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2])
    return np.asarray(out + [[-30, -30, -5, -5], [900, 400, 1100, 700]], np.float32)   # outside / clipped boxes too


def main():
    refharness.install_reference()
    import torch
    from boxmot.reid.backbones.mobilenetv2 import mobilenetv2_x1_4
    from boxmot.reid.backbones.osnet import osnet_x1_0
    from boxmot.reid.backends.base_backend import BaseModelBackend
    from boxmot.reid.core.preprocessing import get_preprocess_fn

    from boxmot_b200.synthetic import make_mobilenetv2_state, make_osnet_state

    class RefBackend(BaseModelBackend):
        def __init__(self, model):
            self.device = torch.device("cpu")
            self.half = False
            self.input_shape = (256, 128)
            self.nhwc = False
            self.preprocess_fn = get_preprocess_fn(None)
            self.mean_array = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
            self.std_array = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
            self.model = model

        def forward(self, x):
            return self.model(x)

        def load_model(self, w):
            pass

    img = np.random.default_rng(IMAGE_SEED).integers(0, 255, size=(540, 960, 3), dtype=np.uint8)
    boxes = boxes_for()
    out = {"boxes": boxes}
    m = osnet_x1_0(num_classes=1041, pretrained=False)
    m.load_state_dict(make_osnet_state("osnet_x1_0", seed=OSNET_SEED), strict=True)
    out["osnet_x1_0"] = np.asarray(RefBackend(m.eval()).get_features(boxes, img), np.float32)
    m = mobilenetv2_x1_4(num_classes=1041, loss="softmax", pretrained=False)
    m.load_state_dict(make_mobilenetv2_state(1.4, seed=MBV2_SEED), strict=True)
    out["mobilenetv2_x1_4"] = np.asarray(RefBackend(m.eval()).get_features(boxes, img), np.float32)
    np.savez_compressed(HERE / "reid_arch_reference.npz", **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
