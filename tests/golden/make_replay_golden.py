"""Golden files for boxmot_b200/replay.py, produced by the UNMODIFIED reference in this container:

* replay_dets.npy / replay_embs.npy / replay_empty_embs.npy -- written by `AppendableNpyWriter`
  (boxmot/data/cache.py:140-260) chunk by chunk, re-opened once in the middle (resume path), closed;
* replay_frames.npz -- what `MOTSequence` (boxmot/data/dataset.py:280-430) yields for those files over an image list
  with gaps, with and without `target_fps` thinning, and `convert_to_mot_format` rows for a fixed tracker output.

Run:  python tests/golden/make_replay_golden.py      (needs /root/reference; the outputs are committed)
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
from tests.golden.refharness import install_reference  # noqa: E402


def synthetic_sequence(seed=11, n_frames=40, dim=16):
    """(frame_id, dets (n,6), embs (n,dim)); frames 7, 8 and 23 have no detections, frame ids start at 1."""
    rng = np.random.default_rng(seed)
    out = []
    for f in range(1, n_frames + 1):
        n = 0 if f in (7, 8, 23) else int(rng.integers(1, 9))
        xy = rng.uniform(0, 500, (n, 2))
        wh = rng.uniform(20, 120, (n, 2))
        d = np.column_stack([xy, xy + wh, rng.uniform(0.05, 0.99, n), rng.integers(0, 3, n)]).astype(np.float32)
        e = rng.normal(size=(n, dim)).astype(np.float32)
        out.append((f, d, e))
    return out


def main():
    install_reference()
    import tempfile

    from boxmot.data.cache import AppendableNpyWriter
    from boxmot.data.dataset import MOTSequence
    from boxmot.engine.tracking.mot import convert_to_mot_format

    frames = synthetic_sequence()
    dets_path, embs_path, empty_path = HERE / "replay_dets.npy", HERE / "replay_embs.npy", HERE / "replay_empty_embs.npy"
    for p in (dets_path, embs_path, empty_path):
        p.unlink(missing_ok=True)

    def writers():
        return (AppendableNpyWriter(dets_path, dtype=np.float32, trailing_shape=(7,), empty_trailing_shape=(7,)),
                AppendableNpyWriter(embs_path, dtype=np.float32, trailing_shape=None, empty_trailing_shape=(0,)))

    dw, ew = writers()
    for k, (f, d, e) in enumerate(frames):
        if k == 20:  # resume: close and re-open the existing files
            dw.close(); ew.close()
            dw, ew = writers()
        if len(d) == 0:
            continue
        ew.append(e)
        dw.append(np.column_stack([np.full((len(d), 1), f, np.float32), d]).astype(np.float32))
    dw.close(); ew.close()
    w = AppendableNpyWriter(empty_path, dtype=np.float32, trailing_shape=None, empty_trailing_shape=(0,))
    w.close()

    out = {}
    with tempfile.TemporaryDirectory() as td:
        seq_dir = Path(td) / "SEQ"
        (seq_dir / "img1").mkdir(parents=True)
        (seq_dir / "seqinfo.ini").write_text("[Sequence]\nname=SEQ\nframeRate=30\nseqLength=40\n")
        stub = np.zeros((4, 4, 3), np.uint8)
        frame_ids = np.arange(1, 41)
        paths = []
        for f in frame_ids:
            p = seq_dir / "img1" / f"{f:06d}.npy"
            np.save(p, stub)
            paths.append(p)
        for tag, fps in (("full", None), ("fps10", 10)):
            meta = dict(frame_ids=frame_ids.copy(), frame_paths=list(paths), det_path=str(dets_path),
                        emb_path=str(embs_path), seq_dir=seq_dir, mask_path=None)
            seq = MOTSequence("SEQ", meta, target_fps=fps, show_progress=False, skip_image_load=True)
            fids, counts, dsum, esum = [], [], [], []
            for fr in seq:
                fids.append(int(fr["frame_id"]))
                counts.append(len(fr["dets"]))
                dsum.append(float(np.asarray(fr["dets"], np.float64).sum()))
                esum.append(float(np.asarray(fr["embs"], np.float64).sum()))
            out[f"{tag}_fids"] = np.asarray(fids)
            out[f"{tag}_counts"] = np.asarray(counts)
            out[f"{tag}_dsum"] = np.asarray(dsum)
            out[f"{tag}_esum"] = np.asarray(esum)
    rng = np.random.default_rng(5)
    tracks = np.column_stack([rng.uniform(0, 300, (6, 2)), rng.uniform(300, 600, (6, 2)), np.arange(1, 7),
                              rng.uniform(0.3, 0.9, 6), rng.integers(0, 3, 6), np.arange(6)[::-1]]).astype(np.float32)
    tracks[0, :4] = [10.5, 20.5, 31.0, 41.0]   # .5 cases: numpy rounds half to even AFTER the subtraction
    out["mot_tracks"] = tracks
    out["mot_rows"] = convert_to_mot_format(tracks, 17)
    np.savez(HERE / "replay_frames.npz", **out)
    print("wrote", dets_path.name, embs_path.name, empty_path.name, "replay_frames.npz")


if __name__ == "__main__":
    main()
