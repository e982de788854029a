"""Device-resident frame loop: `MultiStreamTracker.update_device` / `fetch`, `DeviceFrameLoop` and
`TrackerRuntime` on the GPU against the host-input path and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _streams(S, F, n, dim):
    from oracle.streams import stress_embeddings, stress_stream

    frames = [stress_stream(n, F, seed=70 + s) for s in range(S)]
    embs = [stress_embeddings(frames[s], n, dim=dim, seed=80 + s) for s in range(S)]
    return frames, embs


def test_device_loop_equals_host_input_path_and_oracle():
    import torch

    import boxmot_b200 as bb
    from boxmot_b200.replay import to_mot_rows
    from boxmot_b200.runtime import DeviceFrameLoop, TimingStats
    from oracle.trackers import BotSortOracle

    S, F, n, dim, CD = 3, 30, 40, 64, 64
    frames, embs = _streams(S, F, n, dim)
    dev = bb.MultiStreamTracker("botsort", n_streams=S, cap_tracks=256, cap_dets=CD, feat_dim=dim)
    host = bb.MultiStreamTracker("botsort", n_streams=S, cap_tracks=256, cap_dets=CD, feat_dim=dim)

    def device_frames():
        for f in range(F):
            d = np.zeros((S, CD, 6), np.float32)
            e = np.zeros((S, CD, dim), np.float32)
            rows = []
            for s in range(S):
                k = len(frames[s][f])
                d[s, :k], e[s, :k] = frames[s][f], embs[s][f]
                rows.append(k)
            yield torch.from_numpy(d).cuda(), rows, None, torch.from_numpy(e).cuda()

    ts = TimingStats()
    got = DeviceFrameLoop(dev, ts).run(device_frames(), every_frame=True)
    assert ts.frames == F and ts.totals["track"] > 0
    oracles = [BotSortOracle() for _ in range(S)]
    want = [[] for _ in range(S)]
    for f in range(F):
        h = host.update([frames[s][f] for s in range(S)], None, [embs[s][f] for s in range(S)])
        for s in range(S):
            o = oracles[s].update(frames[s][f], None, embs[s][f].copy())
            assert np.array_equal(np.asarray(h[s])[:, 4:], np.asarray(o, np.float32).reshape(-1, 8)[:, 4:])
            if len(h[s]):
                want[s].append(to_mot_rows(np.asarray(h[s]), f + 1))
    for s in range(S):
        assert np.array_equal(got[s], np.concatenate(want[s])), f"stream {s}"
    # without per-frame fetch the loop never waits; the final rows are those of the last frame
    dev.reset()
    last = DeviceFrameLoop(dev).run(device_frames(), every_frame=False)
    for s in range(S):
        assert np.array_equal(last[s], want[s][-1])
    with pytest.raises(bb.B200Error, match="device memory"):
        dev.update_device(torch.zeros(S, CD, 6), [0] * S)
    with pytest.raises(bb.B200Error, match="exceed cap_dets"):
        dev.update_device(torch.zeros(S, CD, 6).cuda(), [CD + 1] + [0] * (S - 1))


def test_tracker_runtime_reports_device_split(tmp_path):
    from boxmot_b200.reid import B200ReID
    from boxmot_b200.runtime import TimingStats, TrackerRuntime
    from boxmot_b200.synthetic import bench_stream, make_osnet_state
    from boxmot_b200.weights import export_blob
    from oracle import reid as orid
    from oracle.trackers import BotSortOracle
    from tests.common import BOTSORT_YAML

    sd = make_osnet_state("osnet_x0_25", seed=9)
    blob = export_blob(sd, tmp_path / "rt.b200reid")
    img, frames = bench_stream(24, 8, hw=(360, 640))
    ts = TimingStats()
    rt = TrackerRuntime.create("botsort", reid_model=B200ReID(blob), timing_stats=ts, cap_tracks=128, cap_dets=64)
    orc = BotSortOracle(reid_model=orid.OracleReID(sd), **BOTSORT_YAML)
    for f, d in enumerate(frames):
        ts.start_frame()
        tracks, ms = rt.update(d, img)
        ts.end_frame()
        want = np.asarray(orc.update(d, img), np.float32).reshape(-1, 8)
        assert np.array_equal(tracks[:, 4:], want[:, 4:])
        assert 0 < ts.get_last_reid_time() <= ms * 1.5 and ms == ts.get_last_track_time()
        mot = rt.format_for_mot(tracks, f + 1)
        assert mot.shape == (len(tracks), 9) if len(tracks) else mot.shape == (0, 0)
    s = ts.summary()
    assert s["frames"] == len(frames) and s["reid_device"] > 0 and s["assoc_device"] > 0 and s["fps"] > 0


@pytest.mark.parametrize("kind", ["deepocsort", "strongsort"])
def test_device_resident_embeddings_reach_every_family(kind):
    """`update_device(d_embs=...)` must associate on the caller's embeddings for DeepOCSORT and StrongSORT too (round 1
    copied them only on the StrongSORT branch; DeepOCSORT silently used its own stale buffer)."""
    import torch

    import boxmot_b200 as bb

    S, F, n, dim, CD = 2, 24, 30, 64, 64
    frames, embs = _streams(S, F, n, dim)
    kw = dict(n_streams=S, cap_tracks=256, cap_dets=CD, feat_dim=dim)
    dev, host = bb.MultiStreamTracker(kind, **kw), bb.MultiStreamTracker(kind, **kw)
    for f in range(F):
        d = np.zeros((S, CD, 6), np.float32)
        e = np.zeros((S, CD, dim), np.float32)
        rows = []
        for s in range(S):
            k = len(frames[s][f])
            d[s, :k], e[s, :k] = frames[s][f], embs[s][f]
            rows.append(k)
        dev.update_device(torch.from_numpy(d).cuda(), rows, torch.from_numpy(e).cuda())
        got = dev.fetch()
        want = host.update([frames[s][f] for s in range(S)], None, [embs[s][f] for s in range(S)])
        for s in range(S):
            assert np.array_equal(np.asarray(got[s]), np.asarray(want[s])), (kind, f, s)
