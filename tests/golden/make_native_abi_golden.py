"""What the reference's own ctypes wrappers bind for this path (`boxmot/native/trackers/{botsort,bytetrack,_common}.py`,
`boxmot/native/reid/capi.py`): config struct layouts, the update argument lists and the symbol names.  Dumped to
tests/golden/native_abi.json -- the contract include/boxmot_b200.h Part 1 must honour.
Run: python tests/golden/make_native_abi_golden.py"""
import ctypes
import json
import re
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1]))
import refharness  # noqa: E402

refharness.install_reference()
from boxmot.native.trackers import _common as c  # noqa: E402
from boxmot.native.trackers import botsort as bs  # noqa: E402
from boxmot.native.trackers import bytetrack as bt  # noqa: E402

NAMES = {ctypes.c_float: "float", ctypes.c_int: "int", ctypes.c_char_p: "char*", ctypes.c_void_p: "void*",
         ctypes.c_double: "double", ctypes.POINTER(ctypes.c_int): "int*"}


def fields(struct):
    return [[n, NAMES[t]] for n, t in struct._fields_]


ref = refharness.REFERENCE_ROOT / "boxmot" / "native"
text = "".join((ref / p).read_text() for p in ("reid/capi.py", "trackers/botsort.py", "trackers/bytetrack.py"))
symbols = sorted(s for s in set(re.findall(r"boxmot_(?:reid_capi|botsort|bytetrack)_[a-z_]+", text)) if not s.endswith("_"))
out = {
    "BoxMOTBotSortConfig": fields(bs._BotSortCConfig),
    "BoxMOTByteTrackConfig": fields(bt._ByteTrackCConfig),
    "sizeof": {"BoxMOTBotSortConfig": ctypes.sizeof(bs._BotSortCConfig), "BoxMOTByteTrackConfig": ctypes.sizeof(bt._ByteTrackCConfig)},
    "update_args": [NAMES[t] for t in c.LIVE_UPDATE_ARGTYPES],
    "update_with_embs_args": [NAMES[t] for t in c.LIVE_UPDATE_WITH_EMBS_ARGTYPES],
    "symbols": symbols,
}
(HERE / "native_abi.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out)[:400])
