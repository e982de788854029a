"""Host logic of boxmot_b200.runtime with a stand-in tracker: keyword forwarding, 2-d normalisation,
timing accessors, MOT formatting -- the behaviours of engine/tracking/runtime.py:15-128 that do not need a GPU."""
import numpy as np

from boxmot_b200.runtime import TimingStats, TrackerRuntime


class _Stub:
    def __init__(self):
        self.seen = []

    def update(self, dets, img, **kw):
        self.seen.append(sorted(kw))
        return np.asarray(dets, np.float32)[:, [0, 1, 2, 3, 4, 4, 5, 5]] if len(dets) else np.empty((0, 8), np.float32)


def test_update_forwards_only_given_keywords_and_times_the_call():
    ts = TimingStats()
    rt = TrackerRuntime(_Stub(), ts)
    d = np.array([[1, 2, 30, 40, 0.9, 0]], np.float32)
    out, ms = rt.update(d, None)
    assert out.shape == (1, 8) and ms >= 0 and ms == ts.get_last_track_time()
    rt.update(d, None, embs=np.zeros((1, 4), np.float32))
    rt.update(d, None, masks=np.zeros((1, 2, 2), np.uint8))
    assert rt.tracker.seen == [[], ["embs"], ["masks"]]
    assert ts.totals["track"] > 0 and ts.get_last_reid_time() == 0
    out, _ = TrackerRuntime(_Stub()).update(np.empty((0, 6), np.float32), None)
    assert out.ndim == 2 and out.size == 0


def test_timing_stats_accumulates_device_split():
    ts = TimingStats()
    for _ in range(3):
        ts.start_frame()
        ts.reset_frame_reid()
        ts.add_device_times(1.5, 0.5)
        ts.end_frame()
    assert ts.frames == 3 and ts.get_last_reid_time() == 1.5
    s = ts.summary()
    assert abs(s["reid_device"] - 1.5) < 1e-12 and abs(s["assoc_device"] - 0.5) < 1e-12 and s["reid"] == s["reid_device"]


def test_format_for_mot_matches_reference_golden():
    from pathlib import Path

    g = np.load(Path(__file__).parent / "golden" / "replay_frames.npz")
    assert np.array_equal(TrackerRuntime.format_for_mot(g["mot_tracks"], 17), g["mot_rows"])
    assert TrackerRuntime.format_for_mot(np.empty((0, 8), np.float32), 1).shape == (0, 0)
    one = TrackerRuntime.format_for_mot(g["mot_tracks"][0], 17)   # a single 1-d row
    assert np.array_equal(one, g["mot_rows"][:1])
