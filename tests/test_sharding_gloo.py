"""The N>1 path on CPU: two gloo ranks shard independent streams (stream s -> rank s % world), run them with the
oracle trackers, and the gathered per-stream results equal a single-process run; timing reduction is MAX."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_stream(stream_index, frames=25):
    from oracle.streams import bench_stream
    from oracle.trackers import ByteTrackOracle

    _, dets = bench_stream(32, frames, hw=(360, 640), stream=stream_index)
    trk = ByteTrackOracle(track_thresh=0.6, match_thresh=0.9, track_buffer=30)
    digest = 0.0
    for d in dets:
        out = trk.update(d, None)
        digest += float(out[:, 4].sum()) + 1e-3 * float(out[:, :4].sum())
    return digest


def _worker(rank, world, port, n_streams, q):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from boxmot_b200 import sharding

    mine = sharding.streams_for_rank(n_streams, rank, world)
    sharding.barrier(dist)
    results = {s: _run_stream(s) for s in mine}
    elapsed = 1.0 + rank  # fake per-rank timing: the reduction must report the slowest rank
    worst = sharding.reduce_max([elapsed, 10.0 - rank], dist)
    counts = sharding.gather_counts(len(mine) * 25, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, results)
    if rank == 0:
        merged = {}
        for g in gathered:
            merged.update(g)
        q.put((merged, worst, counts))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stream_sharding_matches_single_process():
    from boxmot_b200 import sharding

    n_streams, world = 5, 2
    assert sharding.streams_for_rank(5, 0, 2) == [0, 2, 4] and sharding.streams_for_rank(5, 1, 2) == [1, 3]
    assert sorted(sum((sharding.streams_for_rank(7, r, 3) for r in range(3)), [])) == list(range(7))
    assert sharding.reduce_max([1.5, 2.0]) == [1.5, 2.0] and sharding.gather_counts(7) == [7]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged, worst, counts = q.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert worst == [2.0, 10.0]
    assert counts == [75, 50]
    for s in range(n_streams):
        assert merged[s] == pytest.approx(_run_stream(s), rel=0, abs=0)
