"""Constructor signatures of the four reference trackers (parameter names and default values), dumped from the
UNMODIFIED reference to tests/golden/tracker_ctor_defaults.json.   Run: python tests/golden/make_ctor_defaults.py"""
import inspect
import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1]))
import refharness  # noqa: E402

refharness.install_reference()
from boxmot.trackers.basetracker import BaseTracker  # noqa: E402
from boxmot.trackers.bbox.botsort.botsort import BotSort  # noqa: E402
from boxmot.trackers.bbox.bytetrack.bytetrack import ByteTrack  # noqa: E402
from boxmot.trackers.bbox.deepocsort.deepocsort import DeepOcSort  # noqa: E402
from boxmot.trackers.bbox.strongsort.strongsort import StrongSort  # noqa: E402


def sig(cls):
    out = {}
    for name, p in inspect.signature(cls.__init__).parameters.items():
        if name == "self" or p.kind in (p.VAR_KEYWORD, p.VAR_POSITIONAL):
            continue
        d = p.default
        out[name] = None if d is inspect._empty else (d if isinstance(d, (int, float, str, bool, type(None))) else repr(d))
    return out


data = {"BaseTracker": sig(BaseTracker), "ByteTrack": sig(ByteTrack), "BotSort": sig(BotSort),
        "DeepOcSort": sig(DeepOcSort), "StrongSort": sig(StrongSort)}
(HERE / "tracker_ctor_defaults.json").write_text(json.dumps(data, indent=1, sort_keys=True) + "\n")
for k, v in data.items():
    print(k, v)
