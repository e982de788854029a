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


def _struct_fields(name):
    text = (ROOT / "include" / "boxmot_b200.h").read_text()
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(const char\*|float|int|double) (\w+)$", decl)
        assert m, f"unparsed member of {name}: {decl!r}"
        out.append([m.group(2), {"const char*": "char*"}.get(m.group(1), m.group(1))])
    return out


def _arg_kinds(func):
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "boxmot_b200.h").read_text(), flags=re.S)
    args = re.search(r"\b%s\s*\((.*?)\);" % func, text, flags=re.S).group(1)
    kinds = []
    for a in args.split(","):
        a = " ".join(a.split())
        kinds.append("int*" if a.startswith("int*") else ("void*" if "*" in a else a.split()[0]))
    return kinds


def test_reference_abi_part_matches_the_references_own_ctypes_bindings():
    """include/boxmot_b200.h Part 1 against what boxmot/native/trackers/{botsort,bytetrack,_common}.py and
    boxmot/native/reid/capi.py bind (tests/golden/make_native_abi_golden.py): config struct members in order with
    their C types, the update argument lists, and every bound symbol exported by the library."""
    import json

    from boxmot_b200.build import build_library

    g = json.loads((ROOT / "tests" / "golden" / "native_abi.json").read_text())
    for name in ("BoxMOTBotSortConfig", "BoxMOTByteTrackConfig"):
        assert _struct_fields(name) == g[name], name
    assert _arg_kinds("boxmot_bytetrack_update") == g["update_args"]
    assert _arg_kinds("boxmot_botsort_update") == g["update_with_embs_args"]
    lib = ctypes.CDLL(str(build_library()))
    declared = set(_declared())
    for sym in g["symbols"]:
        assert sym in declared and hasattr(lib, sym), sym

    # the ctypes mirrors used by the Python frontends have the reference's struct sizes
    class BotSortCfg(ctypes.Structure):
        _fields_ = [(n, {"float": ctypes.c_float, "int": ctypes.c_int, "char*": ctypes.c_char_p}[t]) for n, t in g["BoxMOTBotSortConfig"]]

    assert ctypes.sizeof(BotSortCfg) == g["sizeof"]["BoxMOTBotSortConfig"]
