"""Exports of the reference's `TrackResults` (boxmot/trackers/track_results.py:96-200) for a fixed (M, 8) array, dumped to
tests/golden/trackresults_exports.json.   Run: python tests/golden/make_trackresults_golden.py"""
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1]))
import refharness  # noqa: E402

ROWS = [[10.25, 20.5, 110.75, 220.125, 3, 0.91234567, 0, 5], [0.0, 1.5, 33.333, 44.4444, 12, 0.5, 2, 0],
        [640.1, 360.9, 700.0, 480.0, 7, 0.123456789, 1, 2]]


def main():
    refharness.install_reference()
    from boxmot.trackers.track_results import TrackResults

    tr = TrackResults(np.asarray(ROWS, dtype=np.float64))
    out = {"summary": tr.summary(), "json": tr.to_json(), "json_indent": tr.to_json(indent=1), "csv": tr.to_csv(),
           "csv_frame": tr.to_csv(frame_id=17), "xywh": tr.xywh.tolist(), "is_obb": bool(tr.is_obb)}
    with tempfile.TemporaryDirectory() as td:
        p = Path(td) / "a" / "t.csv"
        tr.save_csv(p, frame_id=3)
        tr.save_csv(p, frame_id=4)
        out["save_csv"] = p.read_text()
        m = Path(td) / "b" / "t.txt"
        tr.save_mot(m, frame_id=9)
        out["save_mot"] = m.read_text()
    (HERE / "trackresults_exports.json").write_text(json.dumps(out, indent=1) + "\n")
    print("wrote trackresults_exports.json")


if __name__ == "__main__":
    main()
