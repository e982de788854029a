"""Cached-replay format: `boxmot_b200.replay` against files written / read by the UNMODIFIED reference
(tests/golden/make_replay_golden.py) and against numpy's own NPY reader; the runner's host logic against per-sequence
oracle trackers (the GPU run of the same runner is tests/test_gpu_replay.py)."""
from pathlib import Path

import numpy as np
import pytest

from boxmot_b200 import replay as rp
from tests.golden.make_replay_golden import synthetic_sequence

GOLD = Path(__file__).parent / "golden"


def _write_like_generator(dets_path, embs_path, frames, reopen_at=20):
    def writers():
        return (rp.NpyAppender(dets_path, dtype=np.float32, trailing_shape=(7,), empty_trailing_shape=(7,)),
                rp.NpyAppender(embs_path, dtype=np.float32, trailing_shape=None, empty_trailing_shape=(0,)))

    dw, ew = writers()
    for k, (f, d, e) in enumerate(frames):
        if k == reopen_at:
            dw.close(); ew.close()
            dw, ew = writers()
            assert dw.rows == sum(len(x[1]) for x in frames[:k])
        if len(d) == 0:
            continue
        ew.append(e)
        dw.append(np.column_stack([np.full((len(d), 1), f, np.float32), d]).astype(np.float32))
    dw.close(); ew.close()


def test_appender_files_are_byte_identical_to_the_reference_writer(tmp_path):
    frames = synthetic_sequence()
    _write_like_generator(tmp_path / "d.npy", tmp_path / "e.npy", frames)
    assert (tmp_path / "d.npy").read_bytes() == (GOLD / "replay_dets.npy").read_bytes()
    assert (tmp_path / "e.npy").read_bytes() == (GOLD / "replay_embs.npy").read_bytes()
    w = rp.NpyAppender(tmp_path / "empty.npy", dtype=np.float32, trailing_shape=None, empty_trailing_shape=(0,))
    w.close()
    assert (tmp_path / "empty.npy").read_bytes() == (GOLD / "replay_empty_embs.npy").read_bytes()
    # write_cache is the same sequence of appends without the resume
    rp.write_cache(tmp_path / "d2.npy", tmp_path / "e2.npy", frames)
    assert (tmp_path / "d2.npy").read_bytes() == (GOLD / "replay_dets.npy").read_bytes()
    assert (tmp_path / "e2.npy").read_bytes() == (GOLD / "replay_embs.npy").read_bytes()


@pytest.mark.parametrize("rows,trailing,dtype", [(0, (7,), np.float32), (1, (7,), np.float32), (123456789, (512,), np.float32),
                                                 (5, (), np.float64), (9, (3, 4), np.uint8), (0, (0,), np.float32)])
def test_header_equals_numpy(rows, trailing, dtype, tmp_path):
    import io

    from numpy.lib import format as npf

    for version, writer in (((1, 0), npf.write_array_header_1_0), ((2, 0), npf.write_array_header_2_0)):
        buf = io.BytesIO()
        writer(buf, {"descr": npf.dtype_to_descr(np.dtype(dtype)), "fortran_order": False, "shape": (rows, *trailing)})
        assert rp.npy_header(rows, trailing, dtype, version) == buf.getvalue()


def test_appender_grows_in_place_and_numpy_reads_it(tmp_path):
    p = tmp_path / "a.npy"
    w = rp.NpyAppender(p, dtype=np.float32, trailing_shape=(3,))
    chunks = [np.arange(6, dtype=np.float32).reshape(2, 3) + 10 * i for i in range(50)]
    for c in chunks:
        w.append(c)
        assert np.load(p).shape == (w.rows, 3)   # valid after every append (no close needed)
    w.append(np.array([1, 2, 3], np.float32))     # a 1-d row
    w.append(np.empty((0, 3), np.float32))        # ignored
    with pytest.raises(ValueError):
        w.append(np.zeros((2, 4), np.float32))
    w.close()
    assert np.array_equal(np.load(p), np.concatenate(chunks + [np.array([[1, 2, 3]], np.float32)]))
    # an "empty placeholder" embeddings file is replaced by the real width on the next run
    e = tmp_path / "e.npy"
    rp.NpyAppender(e, trailing_shape=None, empty_trailing_shape=(0,)).close()
    assert np.load(e).shape == (0, 0)
    w = rp.NpyAppender(e, trailing_shape=None, empty_trailing_shape=(0,))
    w.append(np.ones((2, 8), np.float32))
    w.close()
    assert np.load(e).shape == (2, 8)


def test_sequence_cache_yields_what_the_reference_dataset_yields():
    g = np.load(GOLD / "replay_frames.npz")
    ids = np.arange(1, 41)
    for tag, fps in (("full", None), ("fps10", 10)):
        sc = rp.SequenceCache(GOLD / "replay_dets.npy", GOLD / "replay_embs.npy", frame_ids=ids, name="SEQ",
                              orig_fps=30, target_fps=fps)
        got = list(sc.frames())
        assert [f for f, _, _ in got] == g[f"{tag}_fids"].tolist()
        assert [len(d) for _, d, _ in got] == g[f"{tag}_counts"].tolist()
        np.testing.assert_array_equal([float(np.asarray(d, np.float64).sum()) for _, d, _ in got], g[f"{tag}_dsum"])
        np.testing.assert_array_equal([float(np.asarray(e, np.float64).sum()) for _, _, e in got], g[f"{tag}_esum"])
        for _, d, e in got:
            assert d.shape[1] == 6 and (len(d) == 0 or e.shape == (len(d), 16))


def test_mot_rows_equal_the_reference_formatter():
    g = np.load(GOLD / "replay_frames.npz")
    got = rp.to_mot_rows(g["mot_tracks"], 17)
    assert got.dtype == g["mot_rows"].dtype and np.array_equal(got, g["mot_rows"])
    assert rp.to_mot_rows(np.empty((0, 8), np.float32), 3).shape == (0, 9)


def test_cache_paths_layout(tmp_path):
    d, e = rp.cache_paths(tmp_path, "yolox_x.pt", "MOT17-02", benchmark="MOT17-ablation", split="train",
                          reid_key="osnet_x0_25_msmt17_pt_pytorch_py")
    assert d == tmp_path / "dets_n_embs/MOT17-ablation/train/yolox_x/dets/MOT17-02.npy"
    assert e == tmp_path / "dets_n_embs/MOT17-ablation/train/yolox_x/embs/osnet_x0_25_msmt17_pt_pytorch_py/resize/MOT17-02.npy"
    assert rp.cache_paths(tmp_path, "public", "S")[1] is None


class _OracleStreams:
    """S oracle trackers behind the multi-stream `update(dets_list, imgs, embs_list)` interface."""

    def __init__(self, make, n, feat_dim, with_reid):
        self.t = [make() for _ in range(n)]
        self.n_streams, self.feat_dim, self.with_reid = n, feat_dim, with_reid
        self.calls = [0] * n

    def last_device_ms(self):
        return 0.25, 0.75

    def update(self, dets, imgs, embs):
        out = []
        for i, t in enumerate(self.t):
            self.calls[i] += 1
            e = None if embs is None else (np.zeros((0, self.feat_dim), np.float32) if embs[i] is None else embs[i].copy())
            out.append(t.update(dets[i], None, e) if self.with_reid else t.update(dets[i], None))
        return out


def _cohort_cache(tmp_path, name, seed, n_frames, dim=32, gaps=()):
    from oracle.streams import stress_embeddings, stress_stream

    frames = stress_stream(24, n_frames, seed=seed)
    embs = stress_embeddings(frames, 24, dim=dim, seed=seed + 1)
    seq = [(f + 1, (np.empty((0, 6), np.float32) if f in gaps else frames[f]), embs[f] if f not in gaps else None)
           for f in range(n_frames)]
    dp, ep = rp.cache_paths(tmp_path, "det", name, reid_key="reid")
    rp.write_cache(dp, ep, seq)
    return rp.SequenceCache(dp, ep, frame_ids=np.arange(1, n_frames + 1), name=name), seq


def test_runner_equals_one_reference_style_loop_per_sequence(tmp_path):
    """Lock-step multi-sequence replay == the per-sequence loop of process_sequence (conf filter, empty frames skipped)
    for sequences of different lengths, with the oracle BoT-SORT standing in for the GPU streams."""
    from oracle.trackers import BotSortOracle

    caches, seqs = zip(*[_cohort_cache(tmp_path, f"S{i}", 20 + i, n, gaps=g)
                         for i, (n, g) in enumerate([(30, (4, 5)), (18, ()), (25, (0, 24))])])
    thr = 0.3
    from boxmot_b200.runtime import TimingStats

    multi = _OracleStreams(BotSortOracle, 3, 32, True)
    ts = TimingStats()
    got = rp.replay_sequences(multi, caches, conf_threshold=thr, timing=ts)
    assert ts.frames == max(multi.calls) and ts.totals["reid_device"] == 0.25 * ts.frames
    assert ts.totals["assoc_device"] == 0.75 * ts.frames
    for i, seq in enumerate(seqs):
        ref, rows = BotSortOracle(), []
        n_calls = 0
        for fid, d, e in seq:
            keep = d[:, 4] >= thr if len(d) else np.zeros(0, bool)
            if not keep.any():
                continue
            n_calls += 1
            out = ref.update(d[keep], None, e[keep].copy())
            if len(out):
                rows.append(rp.to_mot_rows(out, fid))
        want = np.concatenate(rows)
        assert np.array_equal(got[i], want)
        assert multi.calls[i] == max(multi.calls)   # finished streams keep receiving (empty) steps
    assert max(multi.calls) <= 30


def test_runner_rejects_misaligned_or_missing_embeddings(tmp_path):
    from oracle.trackers import BotSortOracle

    cache, _ = _cohort_cache(tmp_path, "S", 3, 5)
    bad = rp.SequenceCache(np.asarray(cache.dets), None, name="S")
    with pytest.raises(ValueError, match="no embeddings"):
        rp.replay_sequences(_OracleStreams(BotSortOracle, 1, 32, True), [bad])
    with pytest.raises(ValueError, match="Row mismatch"):
        rp.SequenceCache(np.asarray(cache.dets), np.asarray(cache.embs)[:-1], name="S")
    with pytest.raises(ValueError, match="32-d"):
        rp.replay_sequences(_OracleStreams(BotSortOracle, 1, 64, True), [cache])


def test_fill_embeddings_is_row_aligned_with_the_cached_detections(tmp_path):
    """Embeddings-only fill (engine/eval/cache.py:252-308) with a stand-in ReID model: one get_features call per
    frame run in stored order, rows aligned with the dets cache, unreadable frames skipped, count mismatch raised."""
    frames = synthetic_sequence(n_frames=12, dim=4)
    dp = tmp_path / "dets" / "S.npy"
    rp.write_cache(dp, None, frames)
    dets = np.load(dp)
    calls = []

    class Model:
        def get_features(self, boxes, img):
            calls.append((len(boxes), int(img[0, 0, 0])))
            return np.column_stack([boxes[:, 0], boxes[:, 3], np.full(len(boxes), img[0, 0, 0]), np.ones(len(boxes))])

    def load(fid):
        return np.full((4, 4, 3), fid, np.uint8)

    ep = tmp_path / "embs" / "reid" / "resize" / "S.npy"
    n = rp.fill_embeddings(dp, ep, Model(), load)
    embs = np.load(ep)
    assert n == len(dets) == len(embs) and embs.dtype == np.float32
    assert np.array_equal(embs[:, 0], dets[:, 1]) and np.array_equal(embs[:, 1], dets[:, 4])
    assert np.array_equal(embs[:, 2], dets[:, 0])                       # each run saw its own frame
    assert [c[1] for c in calls] == sorted({int(f) for f in dets[:, 0]})
    rp.SequenceCache(dp, ep, name="S")                                  # loads: rows match
    # an unreadable frame is skipped (the reference guards the same way) -> fewer rows, which SequenceCache refuses
    ep2 = tmp_path / "e2.npy"
    n2 = rp.fill_embeddings(dp, ep2, Model(), lambda fid: None if fid == 3 else load(fid))
    assert n2 == len(dets) - int((dets[:, 0] == 3).sum())
    with pytest.raises(ValueError, match="Row mismatch"):
        rp.SequenceCache(dp, ep2, name="S")

    class Bad:
        def get_features(self, boxes, img):
            return np.zeros((len(boxes) + 1, 4), np.float32)

    with pytest.raises(RuntimeError, match="Embedding count mismatch"):
        rp.fill_embeddings(dp, tmp_path / "e3.npy", Bad(), load)


@pytest.mark.parametrize("tracker", ["bytetrack", "botsort"])
def test_replay_end_to_end_equals_the_references_process_sequence(tracker, tmp_path):
    """The whole cached-replay path against the result files the reference's own `process_sequence` wrote
    (tests/golden/make_replay_e2e_golden.py: MOTDataset -> TrackerRuntime -> convert_to_mot_format ->
    write_mot_results) for two MOT17-mini sequences of different length replayed TOGETHER here: cache written by
    `write_cache`, read by `SequenceCache`, confidence filter 0.2, lock-step runner, `write_mot_results`."""
    from oracle.trackers import BotSortOracle, ByteTrackOracle
    from tests import common
    from tests.golden.make_replay_e2e_golden import N_FRAMES, REID_KEY, SEQ, build_tree

    with_embs = tracker == "botsort"
    build_tree(tmp_path, common, with_embs)
    caches = []
    for key, seq in SEQ.items():
        dp, ep = rp.cache_paths(tmp_path / "proj", "public", seq, reid_key=REID_KEY if with_embs else None)
        caches.append(rp.SequenceCache(dp, ep, frame_ids=np.arange(1, N_FRAMES[key] + 1), name=seq))
    make = (lambda: ByteTrackOracle(**common.BYTETRACK_YAML)) if tracker == "bytetrack" else \
        (lambda: BotSortOracle(**common.BOTSORT_YAML))
    multi = _OracleStreams(make, len(caches), 512, with_embs)
    rows = rp.replay_sequences(multi, caches, conf_threshold=0.2)
    got = ""
    for (key, seq), r in zip(SEQ.items(), rows):
        out = tmp_path / "exp" / f"{seq}.txt"
        rp.write_mot_results(out, r)
        got += f"# {seq} frames={N_FRAMES[key]}\n" + out.read_text()
    want = (GOLD / f"replay_e2e_{tracker}.txt").read_text()
    assert got == want
