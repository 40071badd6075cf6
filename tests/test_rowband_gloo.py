"""CPU tests of the row-band split of one frame over several ranks (rtdm_b200/rowband.py):
the partition / halo / stitch / global-speckle logic with the CPU oracle standing in for the device matcher,
single-process for many geometries and world_size 2 over gloo for the gather."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp


def _fns(orc, p):
    band_p = dict(p, speckleWindowSize=0, speckleRange=0)

    def band_fn(L, R, roi1, roi2):
        return orc.bm_compute(np.ascontiguousarray(L), np.ascontiguousarray(R), orc.make_params(**dict(band_p, roi1=roi1, roi2=roi2)))

    def speckle_fn(full):
        if p["speckleWindowSize"] > 0 and p["speckleRange"] >= 0:
            return orc.filter_speckles(full, -16, p["speckleWindowSize"], p["speckleRange"])
        return full
    return band_fn, speckle_fn


@pytest.mark.parametrize("case", [
    dict(W=160, H=96, world=2, p=dict(blockSize=9, numDisparities=32, speckleWindowSize=60, speckleRange=16, disp12MaxDiff=1)),
    dict(W=160, H=97, world=3, p=dict(blockSize=13, numDisparities=32, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)),
    dict(W=150, H=61, world=4, p=dict(blockSize=5, numDisparities=16, speckleWindowSize=0, speckleRange=0, disp12MaxDiff=-1)),
    dict(W=140, H=80, world=8, p=dict(blockSize=15, numDisparities=16, speckleWindowSize=30, speckleRange=8, disp12MaxDiff=2)),
    dict(W=160, H=90, world=2, p=dict(blockSize=11, numDisparities=32, speckleWindowSize=50, speckleRange=16, disp12MaxDiff=1,
                                      preFilterType=0, preFilterSize=9, preFilterCap=25)),
    dict(W=160, H=96, world=3, roi1=(20, 10, 120, 70), roi2=(5, 20, 150, 70),
         p=dict(blockSize=9, numDisparities=32, speckleWindowSize=60, speckleRange=16, disp12MaxDiff=1)),
])
def test_bands_stitch_to_the_whole_frame(case):
    from oracle import oracle as orc
    from rtdm_b200 import rowband, synth
    W, H, world, p = case["W"], case["H"], case["world"], case["p"]
    roi1, roi2 = case.get("roi1"), case.get("roi2")
    L, R, _ = synth.stereo_pair(W, H, p["numDisparities"], 600 + H)
    band_fn, speckle_fn = _fns(orc, p)
    halo = rowband.band_halo(p["blockSize"], p.get("preFilterType", 1), p.get("preFilterSize", 9))
    bands = [rowband.compute_band(band_fn, L, R, H, world, r, halo, roi1, roi2) for r in range(world)]
    got = speckle_fn(rowband.stitch_rows(bands))
    ref = orc.bm_compute(L, R, orc.make_params(**dict(p, roi1=roi1, roi2=roi2)))
    assert got.shape == ref.shape and np.array_equal(got, ref), int((got != ref).sum())
    # the partition: every row once, even first input row, halo inside the image
    rows = [rowband.band_rows(H, world, r, halo) for r in range(world)]
    assert rows[0][0] == 0 and rows[-1][1] == H and all(rows[k][1] == rows[k + 1][0] for k in range(world - 1))
    assert all(i0 % 2 == 0 and 0 <= i0 <= y0 and y1 <= i1 <= H for (y0, y1, i0, i1) in rows)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "rt-depth-map_b200")); sys.path.insert(0, root)
    from oracle import oracle as orc
    from rtdm_b200 import rowband, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = dict(blockSize=9, numDisparities=32, speckleWindowSize=60, speckleRange=16, disp12MaxDiff=1)
    L, R, _ = synth.stereo_pair(192, 101, 32, 4242)
    band_fn, speckle_fn = _fns(orc, p)
    out = rowband.compute_frame_distributed(band_fn, speckle_fn, L, R, dist, rowband.band_halo(9))
    q.put((rank, None if out is None else out.copy()))
    dist.destroy_process_group()


def test_two_rank_row_bands_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    from oracle import oracle as orc
    from rtdm_b200 import synth
    L, R, _ = synth.stereo_pair(192, 101, 32, 4242)
    ref = orc.bm_compute(L, R, orc.make_params(blockSize=9, numDisparities=32, speckleWindowSize=60, speckleRange=16, disp12MaxDiff=1))
    assert res[1] is None and np.array_equal(res[0], ref)
