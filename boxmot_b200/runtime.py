"""Frame-loop callers of the hot path: the thin layer that sits between a detector / decoder and
`tracker.update`, restated for inputs that stay in HBM.

* ``TimingStats``     -- `boxmot/utils/timing.py:259-360`: per-phase totals in milliseconds (`reid`, `track`, `total`,
  `frames`, last-frame accessors).  Here the ReID / association split of a frame comes from CUDA events recorded by
  the engine around the two halves of `update` (`boxmot_b200_tracker_last_device_ms`), not from host clocks wrapped
  around a Python ReID call (`wrap_tracker_reid`).
* ``TrackerRuntime``  -- `boxmot/engine/tracking/runtime.py:15-128`: `create(...)`, `update(dets, img, embs, masks) ->
  (tracks, elapsed_ms)` with the same keyword forwarding and 2-d normalisation, `format_for_mot`.
* ``DeviceFrameLoop`` -- the loop of `Results._run_tracker` (`engine/tracking/results.py:467-495`) for S streams whose
  frames and detector boxes are CUDA tensors: no host hop on the way in, rows copied out only when asked for.
"""
from __future__ import annotations

import time
from typing import Any, List, Optional, Sequence

import numpy as np

from .replay import to_mot_rows
from .trackers import MultiStreamTracker, TrackResults, create_tracker


class TimingStats:
    """Totals in milliseconds; same accessor names as the reference class for the phases this path has."""

    KEYS = ("reid", "reid_device", "assoc_device", "track", "total")

    def __init__(self):
        self.reset()

    def reset(self):
        self.totals = {k: 0.0 for k in self.KEYS}
        self.frames = 0
        self._frame_start = None
        self._track_start = None
        self._last_track_time = 0.0
        self._last_reid_time = 0.0

    def start_frame(self):
        self._frame_start = time.perf_counter()

    def end_frame(self):
        if self._frame_start is not None:
            self.totals["total"] += (time.perf_counter() - self._frame_start) * 1000
            self.frames += 1
            self._frame_start = None

    def start_tracking(self):
        self._track_start = time.perf_counter()

    def end_tracking(self):
        if self._track_start is not None:
            elapsed = (time.perf_counter() - self._track_start) * 1000
            self.totals["track"] += elapsed
            self._last_track_time = elapsed
            self._track_start = None

    def get_last_track_time(self):
        return self._last_track_time

    def get_last_reid_time(self):
        return self._last_reid_time

    def reset_frame_reid(self):
        self._last_reid_time = 0.0

    def add_reid_time(self, time_ms):
        self.totals["reid"] += time_ms
        self._last_reid_time += time_ms

    def add_device_times(self, reid_ms: float, assoc_ms: float):
        """CUDA-event durations of the frame's two halves on the engine stream."""
        self.totals["reid_device"] += reid_ms
        self.totals["assoc_device"] += assoc_ms
        self.add_reid_time(reid_ms)

    def summary(self) -> dict:
        n = max(self.frames, 1)
        out = {k: v / n for k, v in self.totals.items()}
        out["frames"] = self.frames
        out["fps"] = 1000.0 / out["total"] if out["total"] > 0 else 0.0
        return out


class TrackerRuntime:
    """Wrap one tracker with timing and formatting helpers (`engine/tracking/runtime.py:15-128`)."""

    def __init__(self, tracker: Any, timing_stats: Optional[TimingStats] = None) -> None:
        self.tracker = tracker
        self.timing_stats = timing_stats

    @classmethod
    def create(cls, tracker_name: str, reid_weights=None, device=None, half: bool = False, per_class: bool = False,
               evolve_param_dict: Optional[dict] = None, timing_stats: Optional[TimingStats] = None,
               **kwargs: Any) -> "TrackerRuntime":
        tracker = create_tracker(str(tracker_name).lower(), reid_weights=reid_weights, device=device, half=half,
                                 per_class=per_class, evolve_param_dict=evolve_param_dict, **kwargs)
        return cls(tracker, timing_stats=timing_stats)

    @staticmethod
    def _ensure_2d_tracks(tracks) -> np.ndarray:
        arr = np.asarray(tracks, dtype=np.float32)
        if arr.size == 0:
            return arr if arr.ndim == 2 else np.empty((0, 0), dtype=np.float32)
        return arr.reshape(1, -1) if arr.ndim == 1 else arr

    @staticmethod
    def format_for_mot(tracks, frame_idx: int) -> np.ndarray:
        arr = TrackerRuntime._ensure_2d_tracks(tracks)
        if arr.size == 0:
            return np.empty((0, 0), dtype=np.float32)
        return to_mot_rows(arr, frame_idx)

    def update(self, dets, img, embs=None, masks=None):
        ts = self.timing_stats
        if ts is not None:
            ts.reset_frame_reid()
            ts.start_tracking()
        else:
            t0 = time.perf_counter()
        try:
            kwargs = {}
            if embs is not None:
                kwargs["embs"] = embs
            if masks is not None:
                kwargs["masks"] = masks
            tracks = self.tracker.update(dets, img, **kwargs)
        finally:
            if ts is not None:
                ts.end_tracking()
                elapsed_ms = ts.get_last_track_time()
                eng = getattr(self.tracker, "_engine", None)
                if eng is not None:
                    ts.add_device_times(*eng.last_device_ms())
            else:
                elapsed_ms = (time.perf_counter() - t0) * 1000
        return self._ensure_2d_tracks(tracks), elapsed_ms


class DeviceFrameLoop:
    """S streams advanced from device-resident frames and detector boxes.

    `step()` only enqueues work (ReID of the next frame overlaps the association of this one on the engine's two
    CUDA streams); `rows()` waits and copies the last frame's rows out; `run()` is the loop of `_run_tracker`
    returning MOT rows per stream.

    Stream-ordering contract.  The engine runs on its own CUDA streams, which are not ordered with the producer's
    (the detector's / decoder's torch stream).  `det_rows` is a host array, so the producer has already been waited
    for when the counts were read; `step()` makes that explicit by synchronising torch's current stream when it is
    handed torch tensors.  On the consumer side the loop keeps a reference to every input it has enqueued until the
    engine has provably finished with it (a `fetch()`), and never lets more than `max_in_flight` frames queue up, so a
    caching allocator cannot hand an input buffer to the next frame while a queued kernel still reads it."""

    def __init__(self, tracker: MultiStreamTracker, timing_stats: Optional[TimingStats] = None,
                 max_in_flight: int = 4):
        self.tracker = tracker
        self.timing_stats = timing_stats
        self.frame_idx = 0
        self.max_in_flight = max(1, int(max_in_flight))
        self._in_flight: list = []

    def _producer_ready(self, *tensors) -> None:
        for t in tensors:
            if t is not None and hasattr(t, "data_ptr") and getattr(t, "is_cuda", False):
                import torch

                torch.cuda.current_stream(t.device).synchronize()
                return

    def step(self, d_dets, det_rows: Sequence[int], d_frames=None, d_embs=None, sync: bool = False) -> None:
        ts = self.timing_stats
        if ts is not None:
            ts.start_frame()
            ts.reset_frame_reid()
            ts.start_tracking()
        if len(self._in_flight) >= self.max_in_flight:
            self.tracker.fetch()          # waits for everything queued so far
            self._in_flight.clear()
        self._producer_ready(d_dets, d_frames, d_embs)
        self.tracker.update_device(d_dets, det_rows, d_embs=d_embs, d_images=d_frames, sync=sync)
        if sync:
            self._in_flight.clear()
        else:
            self._in_flight.append((d_dets, d_frames, d_embs))
        if ts is not None:
            ts.end_tracking()
            if sync:
                ts.add_device_times(*self.tracker.last_device_ms())
            ts.end_frame()
        self.frame_idx += 1

    def rows(self) -> List[TrackResults]:
        out = self.tracker.fetch()
        self._in_flight.clear()
        return out

    def run(self, frames, every_frame: bool = True) -> List[np.ndarray]:
        """`frames`: iterable of `(d_dets, det_rows, d_frames_or_None, d_embs_or_None)`.  With `every_frame` the rows of
        each frame are fetched (one sync per frame, what a display / writer needs); without it only the last
        frame's rows are returned and the loop never waits for the device."""
        S = self.tracker.n_streams
        out: List[List[np.ndarray]] = [[] for _ in range(S)]
        for item in frames:
            d_dets, det_rows, d_frames, d_embs = item
            self.step(d_dets, det_rows, d_frames, d_embs, sync=False)
            if every_frame:
                for i, r in enumerate(self.rows()):
                    if len(r):
                        out[i].append(to_mot_rows(np.asarray(r), self.frame_idx))
        if not every_frame:
            for i, r in enumerate(self.rows()):
                if len(r):
                    out[i].append(to_mot_rows(np.asarray(r), self.frame_idx))
        return [np.concatenate(o, 0) if o else np.empty((0, 9), np.float32) for o in out]
