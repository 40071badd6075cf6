"""GPU: the row-band split of one frame (rtdm_b200/rowband.py) with the CUDA matcher computing every band and the
CUDA speckle stage on the stitched frame must reproduce the whole-frame result bit for bit.  One process and one
GPU compute the bands one after the other here; tools/rowband_multi_gpu.py runs one band per GPU over NCCL."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H,nd,bs,nbands", [(1280, 720, 128, 13, 2), (1280, 720, 128, 13, 8), (640, 481, 64, 15, 3), (320, 100, 32, 5, 4)])
def test_row_bands_equal_whole_frame(gpu, orc, W, H, nd, bs, nbands):
    import torch
    from rtdm_b200 import rowband, synth
    L, R, _ = synth.stereo_pair(W, H, nd, 900 + H)
    rb = rowband.RowBandKonolige(gpu, W, H, 31, bs, 10, nd, 10, 100, 32, 1, nbands=nbands)
    out = rb.compute(torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()).cpu().numpy()
    whole = gpu.CUDAMatcherKonolige(None, None, 31, bs, 0, 10, nd, nd, 10, 100, 32, 1, max_width=W, max_height=H).compute(L, R)
    assert np.array_equal(out, whole), int((out != whole).sum())
    if W * H <= 640 * 481:
        ref = orc.bm_compute(L, R, orc.make_params(preFilterCap=31, blockSize=bs, minDisparity=0, textureThreshold=10, numDisparities=nd,
                                                  uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1))
        assert np.array_equal(out, ref)


def test_speckle_device_entry(gpu, orc):
    import torch
    rng = np.random.default_rng(5)
    a = (rng.integers(0, 5, (2, 120, 160)) * 16).astype(np.int16)
    a[rng.random(a.shape) < 0.25] = -16
    m = gpu.CUDAMatcherKonolige(None, None, 31, 9, 0, 10, 32, 32, 10, 40, 16, 1, max_width=160, max_height=120, max_batch=2)
    d = torch.from_numpy(a).cuda()
    m.speckle_device(2, d.data_ptr(), 160 * 2, 160 * 120 * 2, 160, 120)
    torch.cuda.synchronize()
    for i in range(2):
        assert np.array_equal(d[i].cpu().numpy(), orc.filter_speckles(a[i], -16, 40, 16))
    with pytest.raises(gpu.RtdmError):
        m.speckle_device(3, d.data_ptr(), 320, 160 * 120 * 2, 160, 120)      # more frames than the handle holds
