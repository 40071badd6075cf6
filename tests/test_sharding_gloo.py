"""CPU, world_size 2 over gloo: the N>1 host logic of bench.py (frame sharding, max-over-ranks timing,
whole-job aggregation).  The data path itself has no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_frames, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "rt-depth-map_b200")); sys.path.insert(0, root)
    from rtdm_b200 import sharding, synth
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = sharding.shard_range(n_frames, rank, world)
    # each rank processes its own frames with the CPU oracle standing in for the device call
    p = oracle.make_params(blockSize=9, numDisparities=16, speckleWindowSize=0)
    sums = []
    for i in range(a, b):
        L, R, _ = synth.stereo_pair(96, 64, 16, 1000 + i)
        sums.append(int(oracle.bm_compute(L, R, p).astype(np.int64).sum()))
    ms = 10.0 * (rank + 1)                       # pretend rank 1 is slower
    value, ms_max, total = sharding.whole_job_throughput(b - a, ms, 100.0, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, (a, b, sums))
    q.put((rank, value, ms_max, total, gathered))
    dist.destroy_process_group()


def test_shard_range_partitions():
    from rtdm_b200 import sharding
    for n in (0, 1, 7, 64, 257):
        for w in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_sharding_over_gloo():
    world, n_frames = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, value, ms_max, total, gathered in res:
        assert ms_max == 20.0 and total == n_frames              # max over ranks, all frames counted
        assert abs(value - n_frames * 100.0 / 0.020) < 1e-6
        covered = sorted(i for a, b, _ in gathered for i in range(a, b))
        assert covered == list(range(n_frames))                  # every frame exactly once
    # per-frame results are independent of the sharding: compare with a single-rank pass
    import sys
    from rtdm_b200 import synth
    from oracle import oracle
    p = oracle.make_params(blockSize=9, numDisparities=16, speckleWindowSize=0)
    single = [int(oracle.bm_compute(*synth.stereo_pair(96, 64, 16, 1000 + i)[:2], p).astype(np.int64).sum()) for i in range(n_frames)]
    sharded = [s for a, b, sums in sorted(res[0][4]) for s in sums]
    assert sharded == single
