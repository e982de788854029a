"""Cached-replay runner on the GPU: S cached sequences of different lengths replayed as the S streams of
one MultiStreamTracker must give, per sequence, exactly the MOT rows of a per-sequence oracle loop
(process_sequence semantics: confidence filter, frames without detections skipped)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["botsort", "bytetrack", "strongsort", "deepocsort"])
def test_multi_sequence_replay_equals_per_sequence_oracle(kind, tmp_path):
    import boxmot_b200 as bb
    from boxmot_b200 import replay as rp
    from oracle.deepocsort import DeepOcSortOracle
    from oracle.streams import stress_embeddings, stress_stream, unit_embeddings
    from oracle.strongsort import StrongSortOracle
    from oracle.trackers import BotSortOracle, ByteTrackOracle

    dim, thr = 64, 0.15
    lengths_gaps = [(40, (4, 5)), (22, ()), (31, (0, 30)), (40, (17,))]
    caches, seqs = [], []
    for i, (n, gaps) in enumerate(lengths_gaps):
        frames = stress_stream(32, n, seed=40 + i)
        embs = (unit_embeddings if kind == "deepocsort" else stress_embeddings)(frames, 32, dim=dim, seed=60 + i)
        seq = [(f + 1, np.empty((0, 6), np.float32) if f in gaps else frames[f], None if f in gaps else embs[f])
               for f in range(n)]
        dp, ep = rp.cache_paths(tmp_path, "det", f"S{i}", reid_key="reid")
        rp.write_cache(dp, ep, seq)
        caches.append(rp.SequenceCache(dp, ep, frame_ids=np.arange(1, n + 1), name=f"S{i}"))
        seqs.append(seq)
    kw = dict(strongsort=dict(min_conf=0.3, max_cos_dist=0.4, n_init=2)).get(kind, {})
    make = dict(botsort=BotSortOracle, bytetrack=ByteTrackOracle, strongsort=StrongSortOracle,
                deepocsort=DeepOcSortOracle)[kind]
    trk = bb.MultiStreamTracker(kind, n_streams=len(caches), cap_tracks=256, cap_dets=64, feat_dim=dim, **kw)
    got = rp.replay_sequences(trk, caches, conf_threshold=thr)
    for i, seq in enumerate(seqs):
        ref, rows = make(**kw), []
        for fid, d, e in seq:
            keep = d[:, 4] >= thr if len(d) else np.zeros(0, bool)
            if not keep.any():
                continue
            out = ref.update(d[keep], None) if kind == "bytetrack" else ref.update(d[keep], None, e[keep].copy())
            if len(out):
                rows.append(rp.to_mot_rows(np.asarray(out, np.float32), fid))
        want = np.concatenate(rows)
        assert got[i].shape == want.shape, f"sequence {i}"
        # frame, id, conf, cls, det_ind exact; the integer-rounded box may differ by one unit where the float box
        # sits within 1e-4 of a rounding boundary (boxes are held to 1e-4 relative, ids to exact)
        assert np.array_equal(got[i][:, [0, 1, 6, 7, 8]], want[:, [0, 1, 6, 7, 8]]), f"sequence {i}"
        assert np.abs(got[i][:, 2:6] - want[:, 2:6]).max() <= 1, f"sequence {i}"
        assert (got[i][:, 2:6] != want[:, 2:6]).mean() < 0.01
