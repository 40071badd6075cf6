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


@pytest.mark.parametrize("W,H,nd,bs,nbands", [(1280, 720, 128, 13, 2), (1280, 720, 128, 13, 5), (640, 481, 64, 15, 3), (934, 404, 192, 13, 4),
                                               (333, 201, 48, 7, 2)])
def test_rowband_c_abi_equals_whole_frame(gpu, orc, W, H, nd, bs, nbands):
    """rtdm_bm_rowband_compute (host pointers) and rtdm_bm_rowband_compute_device: bands computed by separate handles, peer
    copies into the stitched frame on devices[0], speckle filter there.  One GPU listed `nbands` times here (the bands of one
    device run one after the other); with more GPUs visible the same call spreads over them (next test)."""
    import torch
    from rtdm_b200 import synth
    L, R, _ = synth.stereo_pair(W, H, nd, 700 + H)
    args = (None, None, 31, bs, 0, 10, nd, nd, 10, 100, 32, 1)
    whole = gpu.CUDAMatcherKonolige(*args, max_width=W, max_height=H)
    ref = whole.compute(L, R)
    rb = gpu.CUDARowBandMatcherKonolige(*args, devices=[0] * nbands, max_width=W, max_height=H)
    out = rb.compute(L, R)
    assert np.array_equal(out, ref), int((out != ref).sum())
    assert rb.last_launches() > 0
    # a ROI (estimator.cpp:54) and strided views
    roi = (W // 8, H // 6, W - W // 4, H - H // 3)
    whole.setROI1(roi); rb.setROI1(roi)
    Lb = np.zeros((H, W + 24), np.uint8); Rb = np.zeros((H, W + 24), np.uint8)
    Lb[:, 8:8 + W] = L; Rb[:, 8:8 + W] = R
    ref2 = whole.compute(L, R)
    out2 = rb.compute(Lb[:, 8:8 + W], Rb[:, 8:8 + W])
    assert np.array_equal(out2, ref2), int((out2 != ref2).sum())
    # device-resident flavour
    dL, dR = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    dD = torch.empty((H, W), dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    rb.compute_device(dL.data_ptr(), W, dR.data_ptr(), W, W, H, dD.data_ptr(), W * 2)
    assert np.array_equal(dD.cpu().numpy(), ref2)
    if W * H <= 640 * 481:
        p = orc.make_params(preFilterCap=31, blockSize=bs, minDisparity=0, textureThreshold=10, numDisparities=nd, uniquenessRatio=10,
                            speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
        assert np.array_equal(out, orc.bm_compute(L, R, p))


def test_rowband_c_abi_over_all_gpus(gpu, orc):
    """One band per visible GPU (skipped on a single-GPU box): the bands travel over cudaMemcpyPeerAsync."""
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    from rtdm_b200 import synth
    W, H, nd, bs = 1920, 1080, 128, 13
    L, R, _ = synth.stereo_pair(W, H, nd, 4242)
    args = (None, None, 31, bs, 0, 10, nd, nd, 10, 100, 32, 1)
    ref = gpu.CUDAMatcherKonolige(*args, max_width=W, max_height=H).compute(L, R)
    out = gpu.CUDARowBandMatcherKonolige(*args, devices=list(range(n)), max_width=W, max_height=H).compute(L, R)
    assert np.array_equal(out, ref), int((out != ref).sum())


def test_rowband_c_abi_argument_checks(gpu):
    args = (None, None, 31, 13, 16, 10, 64, 64, 10, 100, 32, 1)              # minDisparity 16
    with pytest.raises(gpu.RtdmError) as e:
        gpu.CUDARowBandMatcherKonolige(*args, devices=[0, 0], max_width=320, max_height=240)
    assert e.value.code == -gpu.EINVAL
    rb = gpu.CUDARowBandMatcherKonolige(None, None, 31, 13, 0, 10, 64, 64, 10, 100, 32, 1, devices=[0, 0], max_width=320, max_height=240)
    with pytest.raises(gpu.RtdmError):
        rb.compute(np.zeros((300, 320), np.uint8), np.zeros((300, 320), np.uint8))     # taller than the handle
