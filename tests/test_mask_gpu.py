"""GPU parity of the mask front-end (remap of the colour frame + ROI crop + RGB2BGR + BGR2HSV +
inRange, estimator.cpp:38-43) and back-end (findContours RETR_EXTERNAL + boundingRect + size filter + spanning
rectangle, estimator.cpp:46-53 / :164-204) against the cv2 golden fixture and the oracle.  Integer work: bit-exact."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_colormask_matches_cv2_golden(gpu):
    g = load_golden("mask_320x240")
    roi = tuple(int(v) for v in g["roi"])
    cm = gpu.CUDAColorMask(g["map1"], g["map2"], roi, max_batch=2)
    mask, bgr = cm.run(g["rgb"], g["low"], g["high"], want_bgr=True)
    assert np.array_equal(bgr, g["bgr"])
    assert np.array_equal(mask, g["filter_in"])
    assert cm.last_launches() == 1
    both = cm.run(np.stack([g["rgb"], g["rgb"][:, ::-1].copy()]), g["low"], g["high"])
    assert np.array_equal(both[0], g["filter_in"]) and not np.array_equal(both[1], g["filter_in"])


def test_colormask_random_maps_and_ranges_match_oracle(gpu, orc):
    """Arbitrary maps (entries far outside the source: BORDER_CONSTANT 0), every fraction pair, saturated and gray
    colours, empty / full / inverted HSV ranges."""
    rng = np.random.default_rng(23)
    for i, (W, H) in enumerate([(7, 5), (64, 48), (333, 211)]):
        rgb = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
        rgb[: H // 3] = rgb[: H // 3, :, :1]                      # gray rows: diff = 0
        rgb[H // 3: H // 2, :, 1] = 255                           # saturated rows
        m1 = np.stack([rng.integers(-3, W + 3, (H, W)), rng.integers(-3, H + 3, (H, W))], -1).astype(np.int16)
        m2 = rng.integers(0, 1024, (H, W)).astype(np.uint16)
        roi = (1, 1, W - 2, H - 2)
        cm = gpu.CUDAColorMask(m1, m2, roi)
        for low, high in [((0, 0, 0), (179, 255, 255)), ((20, 30, 40), (120, 200, 220)), ((100, 0, 0), (50, 255, 255)),
                          ((0, 0, 0), (0, 0, 255)), (tuple(int(v) for v in rng.integers(0, 100, 3)), tuple(int(v) for v in rng.integers(100, 256, 3)))]:
            got, gb = cm.run(rgb, low, high, want_bgr=True)
            ref, rb = orc.color_mask(rgb, m1, m2, roi, low, high)
            assert np.array_equal(gb, rb), (W, H)
            assert np.array_equal(got, ref), (W, H, low, high)


def test_regions_match_cv2_golden(gpu):
    g = load_golden("mask_320x240")
    H, W = g["filter_out"].shape
    reg = gpu.CUDAObjectRegions(W, H, 2048)
    for name in ("filter_in", "filter_out"):
        rects, roi, nc = reg.run(g[name], 0)
        ref = g["boxes_" + name[7:]]
        assert nc == len(ref) and np.array_equal(rects, ref), name
    rects, roi, nc = reg.run(g["filter_out"], 400)
    ref = np.array([b for b in g["boxes_out"] if b[2] * b[3] >= 400], np.int32)
    assert np.array_equal(rects, ref) and nc == len(g["boxes_out"])
    assert roi == (ref[:, 0].min(), ref[:, 1].min(), (ref[:, 0] + ref[:, 2]).max() - ref[:, 0].min(), (ref[:, 1] + ref[:, 3]).max() - ref[:, 1].min())
    assert reg.last_launches() == 5


def test_regions_random_masks_match_oracle(gpu, orc):
    """Nested components (not listed), border contact, diagonal links, single pixels, non-255 values, strided views,
    empty and full masks, the reference's 'no box' rectangle, overflow of max_regions."""
    from rtdm_b200 import synth
    rng = np.random.default_rng(31)
    reg = gpu.CUDAObjectRegions(1280, 720, 65536)
    sizes = [(1, 1), (2, 3), (17, 9), (160, 120), (333, 211), (934, 404), (1280, 720)]
    for i, (W, H) in enumerate(sizes):
        if W >= 64:
            m = synth.binary_mask(W, H, 7000 + i)
            if i % 2:
                m = np.where(rng.integers(0, 40, (H, W)) == 0, 255 - m, m).astype(np.uint8)      # speckles, holes, nesting
        else:
            m = (rng.integers(0, 3, (H, W)) == 0).astype(np.uint8) * int(rng.integers(1, 256))
        for min_size in (0, 50):
            rects, roi, nc = reg.run(m, min_size)
            rb, rroi = orc.object_regions(m, min_size)
            assert nc == len(orc.contour_boxes(m)), (W, H)
            assert np.array_equal(rects, np.array(rb, np.int32).reshape(-1, 4)), (W, H, min_size)
            assert roi == rroi, (W, H, min_size)
    # ring with a nested blob, a diagonal chain, border contact
    m = np.zeros((40, 60), np.uint8)
    m[20:35, 10:40] = 255; m[23:32, 13:37] = 0; m[26:29, 20:25] = 255; m[0:3, 50:60] = 9; m[10, 30] = 1; m[11, 31] = 1; m[36:40, 0:4] = 255
    rects, roi, nc = reg.run(m, 0)
    assert [tuple(r) for r in rects] == orc.contour_boxes(m) and nc == 4
    big = np.zeros((64, 64), np.uint8); big[10:50, 12:60] = 255
    view = big[:, ::1][5:60, 3:63]                                  # a strided ROI view
    assert [tuple(r) for r in reg.run(view, 0)[0]] == orc.contour_boxes(np.ascontiguousarray(view))
    r0, sp0, nc0 = reg.run(np.zeros((9, 9), np.uint8), 0)
    assert len(r0) == 0 and sp0 == (1000000, 1000000, -2000000, -2000000) and nc0 == 0
    assert [tuple(r) for r in reg.run(np.full((9, 11), 255, np.uint8), 0)[0]] == [(0, 0, 11, 9)]
    checker = np.zeros((20, 20), np.uint8)
    checker[::2, ::2] = 255                                         # 100 isolated pixels
    small = gpu.CUDAObjectRegions(20, 20, 16)
    with pytest.raises(gpu.RtdmError):
        small.run(checker, 0)


def test_mask_chain_on_device(gpu, orc):
    """Decoder output -> colour mask -> open/close filter -> object boxes -> setROI1 + matcher -> depth epilogue,
    with every image staying on the GPU; equal to the oracle chain (estimator.cpp:38-77)."""
    import torch
    g = load_golden("mask_320x240")
    roi = tuple(int(v) for v in g["roi"])
    rw, rh = roi[2], roi[3]
    H, W, _ = g["rgb"].shape
    cm = gpu.CUDAColorMask(g["map1"], g["map2"], roi)
    filt = gpu.CUDAMorphologicalFilter(rw, rh, 8)
    reg = gpu.CUDAObjectRegions(rw, rh, 256)
    dRgb = torch.from_numpy(g["rgb"]).cuda()
    dIn = torch.empty((rh, rw), dtype=torch.uint8, device="cuda"); dOut = torch.empty_like(dIn)
    st = torch.cuda.Stream()
    cm.run_device(1, dRgb.data_ptr(), W * 3, W * H * 3, g["low"], g["high"], dIn.data_ptr(), rw, rw * rh, stream=st.cuda_stream)
    filt.run_device(1, dIn.data_ptr(), dOut.data_ptr(), st.cuda_stream)
    rects, span, nc = reg.run_device(dOut.data_ptr(), rw, rw, rh, 400, st.cuda_stream)
    assert np.array_equal(dIn.cpu().numpy(), g["filter_in"]) and np.array_equal(dOut.cpu().numpy(), g["filter_out"])
    rb, rspan = orc.object_regions(g["filter_out"], 400)
    assert [tuple(r) for r in rects] == rb and span == rspan and nc == len(g["boxes_out"])
    # the spanning rectangle is what the matcher receives as ROI1 (estimator.cpp:54)
    from rtdm_b200 import synth
    nd = 32
    L, R, _ = synth.stereo_pair(rw, rh, nd, 77)
    bm = gpu.CUDAMatcherKonolige(None, None, 31, 9, 0, 10, nd, nd, 10, 100, 32, 1, max_width=rw, max_height=rh)
    bm.setROI1(span)
    got = bm.compute(L, R)
    ref = orc.bm_compute(L, R, orc.make_params(preFilterCap=31, blockSize=9, minDisparity=0, textureThreshold=10, numDisparities=nd,
                                               uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1, roi1=span))
    assert np.array_equal(got, ref)
