"""The C-ABI shared library loads in a GPU-less container and exports every symbol include/boxmot_b200.h declares
(no compute entry point is called here); the ctypes table in boxmot_b200/_lib.py covers the header."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "boxmot_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"BOXMOT_B200_API\s+[\w\s\*]+?\b(boxmot_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from boxmot_b200.build import build_library

    lib = ctypes.CDLL(str(build_library()))
    names = _declared()
    assert len(names) >= 40 and "boxmot_b200_tracker_update_batch" in names and "boxmot_reid_capi_create" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/boxmot_b200.h but not exported: {missing}"


def test_ctypes_table_matches_header():
    from boxmot_b200 import _lib

    lib = _lib.load_library()   # raises if the library is absent: there is no fallback
    declared = set(_declared())
    table = set(_lib.SYMBOLS)
    assert table <= declared, f"bound in _lib.py but not declared in the header: {sorted(table - declared)}"
    for name in table:
        assert getattr(lib, name).argtypes is not None
    # the reference's own ABI names for this path are all present
    for prefix in ("boxmot_reid_capi_", "boxmot_bytetrack_", "boxmot_botsort_"):
        assert any(n.startswith(prefix) for n in declared)


def test_no_device_means_a_loud_error_not_a_fallback():
    import pytest
    import torch

    from boxmot_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.B200Error, match="no CPU fallback"):
        _lib.require_device()
    import boxmot_b200 as bb

    with pytest.raises(_lib.B200Error):
        bb.ByteTrack()
