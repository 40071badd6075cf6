"""Row-band split of ONE large frame over several GPUs for the Konolige matcher (optional part).

Every stage of the BM pipeline except the speckle filter is local to a few rows:
  prefilter   x-Sobel needs source rows y-1 .. y+1 (normalized response: y - ws/2 .. y + ws/2)
  SAD / tex   prefiltered rows y-h .. y+h  (h = blockSize / 2)
  disp12      one row
so rank r computes output rows [y0, y1) from input rows [i0, i1) = [y0 - halo, y1 + halo) clipped to the image,
halo = h + 1 (+ ws/2 for the normalized-response prefilter).  i0 is rounded down to an even row so that the
x-Sobel's "odd last row = cap" rule (OpenCV pairs rows) lands on the same absolute row as in the whole frame.
Image borders keep their reflect / clamp behaviour because a border band starts or ends at the true border.
filterSpeckles is global (components cross bands): the bands run with it switched off, rank 0 gathers them
(torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests) and filters the stitched frame.
SGBM does not split this way (its vertical paths run the full height): replicas only.

The numerical core is injected (`band_fn`, `speckle_fn`), so the same partition / stitch logic runs over the
CUDA matcher in production and over any other implementation of the two calls in tests.
"""
from __future__ import annotations

import numpy as np

from .sharding import shard_range


def band_halo(blockSize: int, preFilterType: int = 1, preFilterSize: int = 9) -> int:
    return blockSize // 2 + (1 if preFilterType == 1 else preFilterSize // 2 + 1)


def band_rows(H: int, world: int, rank: int, halo: int):
    """-> (y0, y1, i0, i1): output rows [y0, y1) of `rank`, input rows [i0, i1) it needs."""
    y0, y1 = shard_range(H, rank, world)
    i0 = max(0, y0 - halo) & ~1
    i1 = min(H, y1 + halo)
    return y0, y1, i0, i1


def shift_roi(roi, i0: int):
    """A ROI rectangle (x, y, w, h) in band coordinates (not clipped: only its row range matters)."""
    return None if roi is None else (roi[0], roi[1] - i0, roi[2], roi[3])


def compute_band(band_fn, left, right, H: int, world: int, rank: int, halo: int, roi1=None, roi2=None, filtered=-16):
    """Runs `band_fn(left_band, right_band, roi1_band, roi2_band, full_rows)` (the matcher WITHOUT the speckle
    filter) on this rank's input rows and returns its output rows [y0, y1) (numpy or torch, whatever band_fn
    returns).  An empty shard (more ranks than rows) returns a 0-row array."""
    y0, y1, i0, i1 = band_rows(H, world, rank, halo)
    if y1 <= y0:
        return left[0:0].astype(np.int16) if isinstance(left, np.ndarray) else None
    # a ROI-less matcher derives its valid rows from its image height; inside a band that must be the FULL image's
    # rows, so the whole-image rectangle is passed explicitly, shifted into band coordinates
    W = left.shape[1]
    r1 = shift_roi(roi1 if roi1 is not None else (0, 0, W, H), i0)
    r2 = shift_roi(roi2 if roi2 is not None else (0, 0, W, H), i0)
    out = band_fn(left[i0:i1], right[i0:i1], r1, r2)
    return out[y0 - i0:y1 - i0]


def stitch_rows(bands):
    return np.concatenate([b for b in bands if b is not None and len(b)], axis=0)


def compute_frame_distributed(band_fn, speckle_fn, left, right, dist, halo: int, roi1=None, roi2=None):
    """One frame, one band per rank.  Returns the finished disparity map on rank 0 (None elsewhere).
    Host-array flavour (gloo or NCCL through pinned staging); the device flavour is RowBandKonolige."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    H, W = left.shape
    band = compute_band(band_fn, left, right, H, world, rank, halo, roi1, roi2)
    rows = [shard_range(H, r, world) for r in range(world)]
    maxr = max(b - a for a, b in rows)
    send = torch.full((maxr, W), 0, dtype=torch.int16)
    if band is not None and len(band):
        send[: len(band)] = torch.from_numpy(np.ascontiguousarray(band))
    # neither gloo nor NCCL moves int16: the rows travel as bytes
    recv = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
    dist.gather(send.view(torch.uint8), [t.view(torch.uint8) for t in recv] if rank == 0 else None, dst=0)
    if rank != 0:
        return None
    full = stitch_rows([recv[r][: rows[r][1] - rows[r][0]].numpy() for r in range(world)])
    return speckle_fn(full)


class RowBandKonolige:
    """torch.distributed flavour (one process per GPU, NCCL gather); the single-process flavour behind the C ABI, with
    peer copies instead of a collective, is rtdm_b200.CUDARowBandMatcherKonolige (rtdm_bm_rowband_*).  Device flavour: every rank holds the frame's input rows on its GPU, computes its band with
    CUDAMatcherKonolige (speckle off), rank 0 gathers the bands over NCCL and runs the speckle stage of a
    full-size handle on the stitched frame.  Constructor arguments as CUDAMatcherKonolige; `dist` is an initialised
    torch.distributed module (None = single process, all bands computed locally one after the other)."""

    def __init__(self, rt, W, H, preFilterCap, blockSize, textureThreshold, numOfDisparities, uniquenessRatio,
                 speckleWindowSize, speckleRange, disp12MaxDiff, *, dist=None, nbands=None, device=0):
        import torch
        self.rt, self.W, self.H, self.dist, self.torch = rt, W, H, dist, torch
        self.world = dist.get_world_size() if dist is not None else int(nbands or 1)
        self.rank = dist.get_rank() if dist is not None else 0
        self.halo = band_halo(blockSize)
        self.dev = torch.device("cuda", device)
        hb = max(i1 - i0 for (_, _, i0, i1) in (band_rows(H, self.world, r, self.halo) for r in range(self.world)))
        self.band = rt.CUDAMatcherKonolige(None, None, preFilterCap, blockSize, 0, textureThreshold, numOfDisparities,
                                           numOfDisparities, uniquenessRatio, 0, 0, disp12MaxDiff,
                                           max_width=W, max_height=max(hb, 1), device=device)
        self.full = None
        if self.rank == 0:
            self.full = rt.CUDAMatcherKonolige(None, None, preFilterCap, blockSize, 0, textureThreshold, numOfDisparities,
                                               numOfDisparities, uniquenessRatio, speckleWindowSize, speckleRange,
                                               disp12MaxDiff, max_width=W, max_height=H, device=device)
        self.stream = torch.cuda.Stream(device=self.dev)

    def _band_device(self, L, R, rank, out_rows):
        """L, R: (H, W) uint8 device tensors; writes this rank's rows into out_rows (a (y1 - y0, W) int16 view)."""
        y0, y1, i0, i1 = band_rows(self.H, self.world, rank, self.halo)
        if y1 <= y0:
            return
        W, hb = self.W, i1 - i0
        tmp = self.torch.empty((hb, W), dtype=self.torch.int16, device=self.dev)
        tmp.record_stream(self.stream)                    # freed by the caching allocator only after self.stream is done with it
        self.band.setROI1((0, -i0, W, self.H)); self.band.setROI2((0, -i0, W, self.H))
        self.band.compute_device(1, L[i0:i1].data_ptr(), W, W * hb, R[i0:i1].data_ptr(), W, W * hb, W, hb,
                                 tmp.data_ptr(), W * 2, W * hb * 2, self.stream.cuda_stream)
        with self.torch.cuda.stream(self.stream):
            out_rows.copy_(tmp[y0 - i0:y1 - i0])

    def compute(self, L, R):
        torch, W, H = self.torch, self.W, self.H
        rows = [shard_range(H, r, self.world) for r in range(self.world)]
        # L / R (and everything allocated below) are produced on torch's current stream; the matcher runs on self.stream
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))
        if self.dist is None:
            full = torch.empty((H, W), dtype=torch.int16, device=self.dev)
            for r in range(self.world):
                self._band_device(L, R, r, full[rows[r][0]:rows[r][1]])
        else:
            maxr = max(b - a for a, b in rows)
            send = torch.zeros((maxr, W), dtype=torch.int16, device=self.dev)
            self.stream.wait_stream(torch.cuda.current_stream(self.dev))     # the zero fill
            send.record_stream(self.stream)
            y0, y1 = rows[self.rank]
            self._band_device(L, R, self.rank, send[: y1 - y0])
            self.stream.synchronize()
            recv = [torch.empty_like(send) for _ in range(self.world)] if self.rank == 0 else None
            self.dist.gather(send.view(torch.uint8), [t.view(torch.uint8) for t in recv] if self.rank == 0 else None, dst=0)
            if self.rank != 0:
                return None
            full = torch.cat([recv[r][: rows[r][1] - rows[r][0]] for r in range(self.world)], dim=0).contiguous()
            self.stream.wait_stream(torch.cuda.current_stream(self.dev))     # the gather and the cat ran on the current stream
            full.record_stream(self.stream)
        self.full.speckle_device(1, full.data_ptr(), W * 2, W * H * 2, W, H, self.stream.cuda_stream)
        self.stream.synchronize()
        return full
