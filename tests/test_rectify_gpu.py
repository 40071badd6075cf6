"""GPU parity of the rectification front-end (cvtColor RGB2GRAY + remap INTER_LINEAR with the
fixed-point maps + ROI crop) against the cv2 golden fixture and the numpy oracle.  Integer arithmetic: bit-exact."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_rectify_matches_cv2_golden(gpu):
    g = load_golden("rectify_320x240")
    roi = tuple(int(v) for v in g["roi"])
    r = gpu.CUDARectifier(g["map1"], g["map2"], roi, max_batch=2)
    assert np.array_equal(r.run(g["rgb"]), g["crop"])
    both = r.run(np.stack([g["rgb"], g["rgb"][::-1].copy()]))
    assert np.array_equal(both[0], g["crop"]) and r.last_launches() == 1
    # whole image as the ROI
    H, W = g["gray"].shape
    assert np.array_equal(gpu.CUDARectifier(g["map1"], g["map2"], (0, 0, W, H)).run(g["rgb"]), g["rect"])


def test_rectify_random_maps_match_oracle(gpu, orc):
    """Arbitrary maps, including entries far outside the source (BORDER_CONSTANT 0) and every fraction pair."""
    rng = np.random.default_rng(17)
    for (W, H) in [(7, 5), (64, 48), (333, 211)]:
        rgb = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
        m1 = np.stack([rng.integers(-3, W + 3, (H, W)), rng.integers(-3, H + 3, (H, W))], -1).astype(np.int16)
        m2 = rng.integers(0, 1024, (H, W)).astype(np.uint16)
        roi = (1, 1, W - 2, H - 2)
        assert np.array_equal(gpu.CUDARectifier(m1, m2, roi).run(rgb), orc.rectify(rgb, m1, m2, roi)), (W, H)


def test_rectify_feeds_the_matcher_on_device(gpu, orc):
    """Decoder output (RGB) -> rectifier -> matcher, all on the GPU; equal to the oracle chain."""
    import torch
    from rtdm_b200 import synth
    W, H, nd = 320, 240, 32
    L, R, _ = synth.stereo_pair(W, H, nd, 55)
    rgbL, rgbR = np.repeat(L[..., None], 3, 2), np.repeat(R[..., None], 3, 2)
    rgbL[..., 0] = np.clip(rgbL[..., 0].astype(int) + 7, 0, 255); rgbR[..., 2] = np.clip(rgbR[..., 2].astype(int) - 5, 0, 255)
    # a mild shear / shift map with sub-pixel fractions
    ys, xs = np.mgrid[0:H, 0:W]
    fx = (xs * 32 + ys * 3 + 40) ; fy = (ys * 32 + 17)
    m1 = np.stack([fx >> 5, fy >> 5], -1).astype(np.int16); m2 = (((fy & 31) << 5) | (fx & 31)).astype(np.uint16)
    roi = (8, 6, 296, 224)
    rw, rh = roi[2], roi[3]
    rect = gpu.CUDARectifier(m1, m2, roi)
    bm = gpu.CUDAMatcherKonolige(None, None, 31, 9, 0, 10, nd, nd, 10, 100, 32, 1, max_width=rw, max_height=rh)
    dIn = [torch.from_numpy(a).cuda() for a in (rgbL, rgbR)]
    dRect = [torch.empty((rh, rw), dtype=torch.uint8, device="cuda") for _ in range(2)]
    dD = torch.empty((rh, rw), dtype=torch.int16, device="cuda")
    st = torch.cuda.Stream()
    for i in range(2):
        rect.run_device(1, dIn[i].data_ptr(), W * 3, W * H * 3, dRect[i].data_ptr(), rw, rw * rh, st.cuda_stream)
    bm.compute_device(1, dRect[0].data_ptr(), rw, rw * rh, dRect[1].data_ptr(), rw, rw * rh, rw, rh, dD.data_ptr(), rw * 2, rw * rh * 2, st.cuda_stream)
    st.synchronize()
    oL, oR = orc.rectify(rgbL, m1, m2, roi), orc.rectify(rgbR, m1, m2, roi)
    assert np.array_equal(dRect[0].cpu().numpy(), oL) and np.array_equal(dRect[1].cpu().numpy(), oR)
    ref = orc.bm_compute(oL, oR, orc.make_params(preFilterCap=31, blockSize=9, minDisparity=0, textureThreshold=10, numDisparities=nd,
                                                uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1))
    assert np.array_equal(dD.cpu().numpy(), ref)


def test_rectify_errors(gpu):
    m1 = np.zeros((10, 12, 2), np.int16); m2 = np.zeros((10, 12), np.uint16)
    with pytest.raises(gpu.RtdmError):
        gpu.CUDARectifier(m1, m2, (4, 4, 12, 4))                      # ROI leaves the image
    r = gpu.CUDARectifier(m1, m2, (0, 0, 12, 10))
    with pytest.raises(gpu.RtdmError):
        r.run(np.zeros((2, 10, 12, 3), np.uint8))                     # batch larger than the handle
