"""The design constraint behind the tensor-core ReID path, kept as a test: with float32 accumulation, single-pass
TF32 / FP16 / BF16 operands in the GEMM-shaped OSNet layers put the embedding outside the 1e-4 parity bound, the
three-pass hi/lo splits (3xTF32, 3xBF16) keep it inside (tests/tools/precision_study.py,
profiles/r1d_reid_operand_precision.md)."""
from tests.tools.precision_study import run


def test_operand_precision_vs_parity_bound():
    rows = {name.split(" ")[0]: worst for name, worst, _, _ in run("osnet_x0_25", n=6)}
    assert rows["fp32"] == 0.0
    assert rows["3xTF32"] < 1e-5 and rows["3xBF16"] < 5e-5
    for single in ("tf32", "bf16", "fp16"):
        assert rows[single] > 1e-4, f"single-pass {single} unexpectedly inside the bound"
