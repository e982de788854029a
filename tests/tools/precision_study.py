"""Which tensor-core operand precision keeps the ReID embedding inside the parity bound?

TEST-SIDE ANALYSIS TOOL (imports the oracle; nothing here is product code).  The GEMM-shaped layers of OSNet (every 1x1
convolution with groups == 1 and the final fc) are re-run with their two operands rounded the way a tcgen05 MMA kind would
read them, accumulation kept in float32 like the tensor core's accumulator; depthwise 3x3, the 7x7 stem, pools, gates and
the normalisations stay float32 (they are CUDA-core work in every design).  Output: max |e - e_fp32| / max |e_fp32| per
crop set, the quantity tests/test_gpu_reid.py bounds by 1e-4 (BASELINE.json north_star).

    python tests/tools/precision_study.py [osnet_x0_25|osnet_x1_0|mobilenetv2_x1_4] [n_crops]
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from boxmot_b200.synthetic import make_osnet_state  # noqa: E402
from oracle import reid as orid  # noqa: E402


def _drop_mantissa(x: torch.Tensor, keep_bits: int, nearest: bool) -> torch.Tensor:
    """float32 -> float32 with only `keep_bits` explicit mantissa bits (10 = TF32, 7 = BF16)."""
    i = x.contiguous().view(torch.int32)
    drop = 23 - keep_bits
    if nearest:
        i = i + ((1 << (drop - 1)) - 1) + ((i >> drop) & 1)   # round to nearest even
    return (i & ~((1 << drop) - 1)).view(torch.float32)


def tf32(x, nearest=True):
    return _drop_mantissa(x, 10, nearest)


def bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def fp16(x):
    return x.to(torch.float16).to(torch.float32)


def split(x, rnd):
    hi = rnd(x)
    return hi, rnd(x - hi)


MODES = {
    "fp32": lambda conv, a, w: conv(a, w),
    "tf32 (truncated operands, what kind::tf32 reads from fp32 data)": lambda conv, a, w: conv(tf32(a, False), tf32(w, False)),
    "tf32 (operands rounded to nearest)": lambda conv, a, w: conv(tf32(a), tf32(w)),
    "3xTF32 (hi*hi + hi*lo + lo*hi)": lambda conv, a, w: (lambda ah, al, wh, wl: conv(ah, wh) + (conv(ah, wl) + conv(al, wh)))(*split(a, tf32), *split(w, tf32)),
    "bf16": lambda conv, a, w: conv(bf16(a), bf16(w)),
    "fp16": lambda conv, a, w: conv(fp16(a), fp16(w)),
    "3xBF16 (hi*hi + hi*lo + lo*hi)": lambda conv, a, w: (lambda ah, al, wh, wl: conv(ah, wh) + (conv(ah, wl) + conv(al, wh)))(*split(a, bf16), *split(w, bf16)),
    "fp16 activations x (fp16 hi + fp16 lo) weights": lambda conv, a, w: (lambda wh, wl: conv(fp16(a), wh) + conv(fp16(a), wl))(*split(w, fp16)),
}


def run(arch="osnet_x0_25", n=24, seed=5):
    if arch.startswith("mobilenetv2"):
        from boxmot_b200.synthetic import make_mobilenetv2_state

        sd = make_mobilenetv2_state(1.4, seed=seed)
    else:
        sd = make_osnet_state(arch, seed=seed)
    rng = np.random.default_rng(3)
    img = rng.integers(0, 255, size=(480, 640, 3), dtype=np.uint8)
    cx, cy = rng.uniform(0, 640, n), rng.uniform(0, 480, n)
    w, h = rng.uniform(20, 120, n), rng.uniform(40, 240, n)
    boxes = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(np.float32)
    real_conv2d, real_linear = orid.F.conv2d, orid.F.linear
    results = {}
    try:
        for name, mode in MODES.items():
            def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, _m=mode):
                if groups == 1 and weight.shape[2] == 1 and weight.shape[3] == 1 and x.shape[1] > 3:
                    y = _m(lambda a, b: real_conv2d(a, b, None, stride, padding, dilation, 1), x, weight)
                    return y if bias is None else y + bias.view(1, -1, 1, 1)
                return real_conv2d(x, weight, bias, stride, padding, dilation, groups)

            def linear(x, weight, bias=None, _m=mode):
                y = _m(lambda a, b: real_linear(a, b), x, weight)
                return y if bias is None else y + bias

            orid.F.conv2d, orid.F.linear = conv2d, linear
            with torch.no_grad():
                results[name] = orid.get_features(sd, boxes, img)
    finally:
        orid.F.conv2d, orid.F.linear = real_conv2d, real_linear
    ref = results["fp32"]
    rows = []
    for name, e in results.items():
        err = np.abs(e - ref).max(axis=1) / np.abs(ref).max(axis=1)
        rows.append((name, float(err.max()), float(np.median(err)),
                     float((1 - (e * ref).sum(1) / (np.linalg.norm(e, axis=1) * np.linalg.norm(ref, axis=1))).max())))
    return rows


if __name__ == "__main__":
    arch = sys.argv[1] if len(sys.argv) > 1 else "osnet_x0_25"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    print(f"{arch}, {n} crops: max_i |e_i - e_i^fp32| / max_i |e_i^fp32| (bound 1e-4)")
    for name, worst, med, cos in run(arch, n):
        print(f"  {name:72s} worst {worst:.2e}  median {med:.2e}  worst cosine distance {cos:.1e}  "
              f"{'OK' if worst <= 1e-4 else 'exceeds the bound'}")
