"""GPU parity of the depth epilogue (/16, reprojectImageTo3D, calc_depth) against the cv2 golden
fixture and the numpy oracle.  Integer and float32 stages are bit-exact; the per-rectangle mean is a double sum
taken in parallel order, so it is compared at 1e-12 relative (a serial double sum of <= 1e6 float32 values
differs from any other order by far less) and the printed centimetre label must be identical."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
MEAN_RTOL = 1e-12


def _same_f32(a, b):
    return np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_depth_matches_cv2_golden(gpu):
    g = load_golden("depth_320x240")
    H, W = g["disp"].shape
    e = gpu.CUDADepthEpilogue(W, H, 16)
    mean, cnt, xyz = e.run(g["disp"], g["Q"], g["mask"], g["rects"], want_xyz=True)
    assert _same_f32(xyz, g["xyz"])
    assert np.array_equal(cnt, g["count"])
    assert np.allclose(mean, g["mean_z"], rtol=MEAN_RTOL, atol=0)
    for a, b in zip(mean, g["mean_z"]):
        assert gpu.distance_cm(a, 1.0) == gpu.distance_cm(b, 1.0)
    assert e.last_launches() == 3
    # no mask == every pixel; no rectangles == only the xyz image
    mean2, cnt2 = e.run(g["disp"], g["Q"], None, g["rects"][2:3])
    assert cnt2[0] >= cnt[2]
    m0, c0, xyz0 = e.run(g["disp"], g["Q"], None, np.zeros((0, 4), np.int32), want_xyz=True)
    assert len(m0) == 0 and _same_f32(xyz0, g["xyz"])


def test_depth_random_inputs_match_oracle(gpu, orc):
    rng = np.random.default_rng(8)
    for (W, H) in [(1, 1), (17, 9), (333, 211), (1280, 720)]:
        disp = (rng.integers(-16, 128 * 16, (H, W))).astype(np.int16)
        disp[rng.random((H, W)) < 0.25] = -16
        Q = np.array([[1, 0, 0, -W / 2 + 0.37], [0, 1, 0, -H / 2 - 0.21], [0, 0, 0, 0.8 * W], [0, 0, 1 / 119.87, 0.004]], np.float64)
        mask = (rng.random((H, W)) < 0.6).astype(np.uint8) * 255
        rects = [(0, 0, W, H)]
        if W > 20:
            rects += [(3, 2, W // 2, H // 3), (W // 3, H // 2, W // 2, H // 4), (W - 1, H - 1, 1, 1)]
        e = gpu.CUDADepthEpilogue(W, H, 8)
        mean, cnt, xyz = e.run(disp, Q, mask, np.array(rects, np.int32), want_xyz=True)
        ref_xyz = orc.reproject_to_3d(orc.disp_div16(disp), Q)
        assert _same_f32(xyz, ref_xyz), (W, H)
        rm, rc = orc.calc_depth(ref_xyz, mask, rects)
        assert np.array_equal(cnt, rc), (W, H)
        assert np.allclose(mean, rm, rtol=MEAN_RTOL, atol=0), (W, H)


def test_depth_on_matcher_output_stays_on_device(gpu, orc):
    """Matcher -> filter -> depth epilogue without leaving the GPU (torch only holds the device memory)."""
    import torch
    from rtdm_b200 import synth
    W, H, nd = 640, 480, 64
    L, R, _ = synth.stereo_pair(W, H, nd, 77)
    M = synth.binary_mask(W, H, 78)
    bm = gpu.CUDAMatcherKonolige(None, None, 31, 13, 0, 10, nd, nd, 10, 100, 32, 1, max_width=W, max_height=H)
    fl = gpu.CUDAMorphologicalFilter(W, H, 8)
    ep = gpu.CUDADepthEpilogue(W, H, 4)
    dL, dR, dM = (torch.from_numpy(a).cuda() for a in (L, R, M))
    dD = torch.empty((H, W), dtype=torch.int16, device="cuda"); dMo = torch.empty_like(dM)
    st = torch.cuda.Stream()
    bm.compute_device(1, dL.data_ptr(), W, W * H, dR.data_ptr(), W, W * H, W, H, dD.data_ptr(), W * 2, W * H * 2, st.cuda_stream)
    fl.run_device(1, dM.data_ptr(), dMo.data_ptr(), st.cuda_stream)
    Q = np.array([[1, 0, 0, -320.5], [0, 1, 0, -240.25], [0, 0, 0, 520.0], [0, 0, 1 / 60.0, 0]], np.float64)
    rects = np.array([[100, 80, 300, 200], [0, 0, W, H]], np.int32)
    mean, cnt = ep.run_device(dD.data_ptr(), W * 2, W, H, Q, dMo.data_ptr(), W, rects, st.cuda_stream)
    disp = orc.bm_compute(L, R, orc.make_params(preFilterCap=31, blockSize=13, minDisparity=0, textureThreshold=10,
                                                numDisparities=nd, uniquenessRatio=10, speckleWindowSize=100,
                                                speckleRange=32, disp12MaxDiff=1))
    rm, rc = orc.calc_depth(orc.reproject_to_3d(orc.disp_div16(disp), Q), orc.morph_open_close(M), rects)
    assert np.array_equal(cnt, rc)
    assert np.allclose(mean, rm, rtol=MEAN_RTOL, atol=0)


def test_depth_errors(gpu):
    e = gpu.CUDADepthEpilogue(64, 48, 2)
    d = np.zeros((48, 64), np.int16); Q = np.eye(4)
    with pytest.raises(gpu.RtdmError):
        e.run(d, Q, None, np.array([[60, 0, 10, 10]], np.int32))        # rectangle leaves the image
    with pytest.raises(gpu.RtdmError):
        e.run(d, Q, None, np.zeros((3, 4), np.int32))                    # more rectangles than the handle holds
    with pytest.raises(gpu.RtdmError):
        e.run(np.zeros((100, 100), np.int16), Q, None, np.zeros((1, 4), np.int32))
