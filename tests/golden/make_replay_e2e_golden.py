"""End-to-end golden for the cached-replay path: the reference's own `process_sequence`
(boxmot/engine/eval/replay.py:216-368: MOTDataset -> TrackerRuntime(bytetrack | botsort with cached embeddings) ->
convert_to_mot_format -> write_mot_results) over a dets / embs cache built from the MOT17-mini fixture.
Writes tests/golden/replay_e2e_<tracker>.txt (the result file the reference writes) for two sequences.

    python tests/golden/make_replay_e2e_golden.py
"""
from __future__ import annotations

import importlib.util
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1]))
import refharness  # noqa: E402

N_FRAMES = {"04": 90, "02": 140}
SEQ = {"04": "MOT17-04-FRCNN", "02": "MOT17-02-FRCNN"}


REID_FILE = "osnet_x0_25_msmt17.pt"               # the file name drives the reference's model registry
REID_KEY = "osnet_x0_25_msmt17_pt_pytorch_py"     # reid_cache_key(REID_FILE): the embeddings bucket name


def build_tree(root: Path, common, with_embs: bool):
    """<root>/mot/<seq>/{img1/<frame>.npy stubs, seqinfo.ini} and <root>/proj/dets_n_embs/public/{dets,embs/...}/<seq>.npy"""
    from boxmot_b200 import replay as rp

    for key, seq in SEQ.items():
        n = N_FRAMES[key]
        d = root / "mot" / seq
        (d / "img1").mkdir(parents=True)
        for f in range(1, n + 1):   # MOTDataset lists the frames from the image files: tiny stub frames
            np.save(d / "img1" / f"{f:06d}.npy", np.zeros((4, 4, 3), np.uint8))
        (d / "seqinfo.ini").write_text(f"[Sequence]\nname={seq}\nimDir=img1\nframeRate=30\nseqLength={n}\n"
                                       "imWidth=1920\nimHeight=1080\nimExt=.jpg\n")
        frames = common.mot17_stream(key)[:n]
        embs = common.mot17_embeddings(key, common.mot17_stream(key), dim=512, seed=21)[:n] if with_embs else [None] * n
        dp, ep = rp.cache_paths(root / "proj", "public", seq, reid_key=REID_KEY if with_embs else None)
        rp.write_cache(dp, ep, [(f + 1, frames[f], embs[f]) for f in range(n)])


def main():
    refharness.install_reference()
    spec = importlib.util.spec_from_file_location("b200_tests_common", HERE.parent / "common.py")
    common = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(common)
    from boxmot.engine.eval.replay import process_sequence
    from boxmot.trackers.bbox.bytetrack import basetrack as bt_base

    for tracker, with_embs in (("bytetrack", False), ("botsort", True)):
        with tempfile.TemporaryDirectory() as td:
            root = Path(td)
            build_tree(root, common, with_embs)
            cfg = dict(common.BYTETRACK_YAML) if tracker == "bytetrack" else dict(common.BOTSORT_YAML, use_cmc=False)
            out = []
            reid_path = None
            if with_embs:   # process_sequence builds the backend from the weights file; the cached rows mean it never runs
                import torch

                from boxmot_b200.synthetic import make_osnet_state

                reid_path = root / REID_FILE
                torch.save(make_osnet_state("osnet_x0_25", seed=1), reid_path)
            for key, seq in SEQ.items():
                bt_base.BaseTrack._count = 0   # ids from 1 per sequence (one process per sequence upstream)
                name, kept, timing = process_sequence(
                    seq, str(root / "mot"), str(root / "proj"), "public", str(reid_path) if with_embs else None,
                    tracker, str(root / "exp"), None, cfg_dict=cfg, conf_threshold=0.2)
                txt = (root / "exp" / f"{seq}.txt").read_text()
                out.append(f"# {seq} frames={len(kept)}\n" + txt)
                print(tracker, seq, "frames", len(kept), "rows", txt.count("\n"), timing["num_frames"])
            (HERE / f"replay_e2e_{tracker}.txt").write_text("".join(out))


if __name__ == "__main__":
    main()
