"""CPU: the oracle (oracle/stereo_oracle.c) against the committed cv2-4.13.0 golden vectors, and
live against cv2 when it imports.  The reference holds no tests/golden vectors of its own
(SURVEY.md section 4), so these fixtures are the pin."""
import json

import numpy as np
import pytest

from conftest import golden_names, load_golden


def _bm_params(orc, p):
    return orc.make_params(
        preFilterType=p.get("preFilterType", 1) if p.get("preFilterType") is not None else 1,
        preFilterSize=p.get("preFilterSize", 9) if p.get("preFilterSize") is not None else 9,
        preFilterCap=p["preFilterCap"], blockSize=p["blockSize"], minDisparity=p["minDisparity"],
        numDisparities=p["numDisparities"], textureThreshold=p["textureThreshold"],
        uniquenessRatio=p["uniquenessRatio"], speckleWindowSize=p["speckleWindowSize"],
        speckleRange=p["speckleRange"], disp12MaxDiff=p["disp12MaxDiff"],
        roi1=p.get("roi1"), roi2=p.get("roi2"))


@pytest.mark.parametrize("name", golden_names("bm_"))
def test_bm_oracle_matches_golden(orc, name):
    g = load_golden(name)
    p = json.loads(str(g["params"]))
    got = orc.bm_compute(g["left"], g["right"], _bm_params(orc, p))
    assert np.array_equal(got, g["disp"]), f"{name}: {(got != g['disp']).sum()} pixels differ"


@pytest.mark.parametrize("name", golden_names("sgbm_"))
def test_sgbm_oracle_matches_golden(orc, name):
    g = load_golden(name)
    p = json.loads(str(g["params"]))
    got, outside = orc.sgbm_compute(g["left"], g["right"], orc.sgbm_params(**p), return_domain_flag=True)
    assert not outside
    assert np.array_equal(got, g["disp"]), f"{name}: {(got != g['disp']).sum()} pixels differ"


@pytest.mark.parametrize("name", golden_names("morph_"))
def test_morph_oracle_matches_golden(orc, name):
    g = load_golden(name)
    assert np.array_equal(orc.morph(g["src"], 0), g["erode"])
    assert np.array_equal(orc.morph(g["src"], 1), g["dilate"])
    assert np.array_equal(orc.morph_open_close(g["src"]), g["openclose"])


def test_ellipse_matches_golden(orc):
    se = load_golden("morph_mask_10x10")["se"]
    j1, j2 = orc.ellipse_rows(10, 10)
    for i in range(10):
        cols = np.nonzero(se[i])[0]
        assert (cols.min(), cols.max() + 1) == (j1[i], j2[i])
    assert int(se.sum()) == 83


def test_speckle_oracle_matches_golden(orc):
    g = load_golden("post_speckle_320x240")
    for key in g.files:
        if key.startswith("sp_"):
            _, ms, md = key.split("_")
            assert np.array_equal(orc.filter_speckles(g["raw"], -16, int(ms), int(md)), g[key]), key


def test_median_oracle_matches_golden(orc):
    g = load_golden("post_median_131x97")
    assert np.array_equal(orc.median3_s16(g["src"]), g["median"])


def test_validate_oracle_matches_golden(orc):
    g = load_golden("post_validate_320x240")
    for d12 in (0, 1, 3):
        assert np.array_equal(orc.validate_disparity(g["raw"], g["cost"], 0, 64, d12), g[f"d12_{d12}"])


def test_synth_is_deterministic():
    """The fixtures store their inputs, but bench/test inputs are regenerated from seeds: guard that."""
    from rtdm_b200 import synth
    g = load_golden("bm_cfg0_320x240_nd64_bs15")
    L, R, _ = synth.stereo_pair(320, 240, 64, 1000)
    assert np.array_equal(L, g["left"]) and np.array_equal(R, g["right"])


# ---- live differential checks against cv2 (skipped only if cv2 cannot be imported) ----------------
def _cv2():
    from oracle import cv2_ref
    if not cv2_ref.have_cv2():
        pytest.skip("cv2 not importable")
    return cv2_ref


def test_bm_oracle_vs_cv2_random_params(orc):
    cv2_ref = _cv2()
    from rtdm_b200 import synth
    rng = np.random.default_rng(7)
    checked = 0
    for i in range(10):
        W, H = int(rng.integers(100, 360)), int(rng.integers(60, 240))
        nd = 16 * int(rng.integers(1, 5)); bs = 2 * int(rng.integers(2, 11)) + 1
        if bs >= min(W, H) or nd + bs >= W:
            continue
        p = dict(preFilterCap=int(rng.integers(1, 32)), blockSize=bs, minDisparity=0,
                 textureThreshold=int(rng.integers(0, 50)), numDisparities=nd,
                 uniquenessRatio=int(rng.integers(0, 30)), speckleWindowSize=int(rng.integers(0, 200)),
                 speckleRange=int(rng.integers(0, 64)), disp12MaxDiff=int(rng.integers(-1, 4)))
        L, R, _ = synth.stereo_pair(W, H, nd, 500 + i)
        ref = cv2_ref.make_bm(**p).compute(L, R)
        got = orc.bm_compute(L, R, _bm_params(orc, p))
        assert np.array_equal(ref, got), (p, int((ref != got).sum()))
        checked += 1
    assert checked >= 5


def test_bm_oracle_vs_cv2_720p(orc):
    cv2_ref = _cv2()
    from rtdm_b200 import synth
    p = dict(preFilterCap=31, blockSize=13, minDisparity=0, textureThreshold=10, numDisparities=128,
             uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
    L, R, _ = synth.stereo_pair(1280, 720, 128, 1000)
    assert np.array_equal(cv2_ref.make_bm(**p).compute(L, R), orc.bm_compute(L, R, _bm_params(orc, p)))


def test_sgbm_oracle_vs_cv2_random_params(orc):
    cv2_ref = _cv2()
    from rtdm_b200 import synth
    rng = np.random.default_rng(8)
    checked = 0
    for i in range(8):
        W, H = int(rng.integers(120, 360)), int(rng.integers(60, 200))
        nd = 16 * int(rng.integers(1, 5)); bs = 2 * int(rng.integers(0, 4)) + 1
        if nd + 2 * bs >= W:
            continue
        p = dict(blockSize=bs, minDisparity=0, numDisparities=nd, uniquenessRatio=int(rng.integers(0, 25)),
                 speckleWindowSize=int(rng.integers(0, 150)), speckleRange=int(rng.integers(0, 8)),
                 disp12MaxDiff=int(rng.integers(-1, 4)), mode=int(rng.integers(0, 2)))
        L, R, _ = synth.stereo_pair(W, H, nd, 600 + i)
        ref = cv2_ref.make_sgbm(**p).compute(L, R)
        got, outside = orc.sgbm_compute(L, R, orc.sgbm_params(**p), return_domain_flag=True)
        if outside:
            continue          # S saturated: cv2 itself is implementation-defined there (App. B.5)
        assert np.array_equal(ref, got), (p, int((ref != got).sum()))
        checked += 1
    assert checked >= 4


def test_depth_oracle_matches_cv2_golden():
    """Depth epilogue restatement (oracle.py: disp_div16, reproject_to_3d, calc_depth) against the fixture made
    with cv2.divide / cv2.reprojectImageTo3D (tests/golden/make_golden.py)."""
    from oracle import oracle as orc
    g = load_golden("depth_320x240")
    d16 = orc.disp_div16(g["disp"])
    assert np.array_equal(d16, g["div16"])
    xyz = orc.reproject_to_3d(d16, g["Q"])
    assert np.array_equal(xyz.view(np.uint32), g["xyz"].view(np.uint32))
    mean, cnt = orc.calc_depth(xyz, g["mask"], g["rects"])
    assert np.array_equal(cnt, g["count"]) and np.array_equal(mean, g["mean_z"])
    # round half to even on the x16 fixed point, negatives included
    assert list(orc.disp_div16(np.array([[-24, -16, -8, 8, 24, 40, 7, 9]], np.int16))[0]) == [-2, -1, 0, 0, 2, 2, 0, 1]
    assert orc.distance_cm(1607.4357, 1.0) == "161 cm"


def test_rectify_oracle_matches_cv2_golden():
    """Rectification restatement (oracle.py: rgb2gray, remap_linear_fixed, rectify) against the fixture made with
    cv2.cvtColor / cv2.remap and maps from cv2.initUndistortRectifyMap(CV_16SC2)."""
    from oracle import oracle as orc
    g = load_golden("rectify_320x240")
    assert np.array_equal(orc.rgb2gray(g["rgb"]), g["gray"])
    assert np.array_equal(orc.remap_linear_fixed(g["gray"], g["map1"], g["map2"]), g["rect"])
    assert np.array_equal(orc.rectify(g["rgb"], g["map1"], g["map2"], tuple(g["roi"])), g["crop"])


def test_mask_oracle_matches_cv2_golden():
    """Mask front-end / back-end restatement (oracle.py: color_mask, contour_boxes, object_regions) against the fixture
    made with cv2 4.13.0 (remap of the 3-channel frame, RGB2BGR, BGR2HSV, inRange, findContours + boundingRect)."""
    from oracle import oracle as orc
    g = load_golden("mask_320x240")
    roi = tuple(int(v) for v in g["roi"])
    mask, bgr = orc.color_mask(g["rgb"], g["map1"], g["map2"], roi, g["low"], g["high"])
    assert np.array_equal(bgr, g["bgr"])
    assert np.array_equal(orc.bgr2hsv(g["bgr"]), g["hsv"])
    assert np.array_equal(mask, g["filter_in"])
    assert np.array_equal(orc.morph_open_close(g["filter_in"]), g["filter_out"])
    for name in ("filter_in", "filter_out"):
        boxes = np.array(orc.contour_boxes(g[name]), np.int32).reshape(-1, 4)
        assert np.array_equal(boxes, g["boxes_" + name[7:]]), name
    bounds, span = orc.object_regions(g["filter_out"], 400)
    ref = [tuple(b) for b in g["boxes_out"] if b[2] * b[3] >= 400]
    assert bounds == ref and len(bounds) >= 3
    assert span == (min(b[0] for b in ref), min(b[1] for b in ref), max(b[0] + b[2] for b in ref) - min(b[0] for b in ref),
                    max(b[1] + b[3] for b in ref) - min(b[1] for b in ref))
    assert orc.object_regions(np.zeros((5, 7), np.uint8), 1) == ([], (1000000, 1000000, -2000000, -2000000))


def test_mask_oracle_vs_cv2_live():
    """Live against cv2 when it imports: every BGR colour through BGR2HSV, and contour boxes of random masks
    (nested components, border contact, diagonal links)."""
    cv2 = pytest.importorskip("cv2")
    assert cv2.__version__ == "4.13.0"
    from oracle import oracle as orc
    from golden.make_golden import cv_boxes
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (512, 512, 3)).astype(np.uint8)
    a[:256, :256, 0] = np.arange(256)[None, :]; a[:256, :256, 1] = np.arange(256)[:, None]; a[:256, :256, 2] = 200
    assert np.array_equal(orc.bgr2hsv(a), cv2.cvtColor(a, cv2.COLOR_BGR2HSV))
    gray = np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 3, 2)
    assert np.array_equal(orc.bgr2hsv(gray), cv2.cvtColor(gray, cv2.COLOR_BGR2HSV))
    for t in range(25):
        W, H = int(rng.integers(3, 160)), int(rng.integers(3, 120))
        n = cv2.GaussianBlur(rng.integers(0, 256, (H, W)).astype(np.uint8), (0, 0), float(rng.uniform(0.6, 4)))
        m = ((n > np.median(n) + int(rng.integers(-3, 4))) * int(rng.integers(1, 256))).astype(np.uint8)
        assert cv_boxes(m) == orc.contour_boxes(m), (t, W, H)


def test_oracle_vs_cv2_min_disparity(orc):
    """minDisparity != 0, both matchers, live against cv2 (BM: whole map incl. the App. B.3 spill row)."""
    cv2_ref = _cv2()
    from rtdm_b200 import synth
    for minD in (16, -16, 5, -37, 2, 1):
        for (W, H, nd, bs) in [(320, 240, 64, 15), (233, 157, 48, 9)]:
            L, R, _ = synth.stereo_pair(W, H, nd, 1234 + minD)
            for roi in (None, (30, 20, W - 70, H - 50)):
                p = dict(preFilterCap=31, blockSize=bs, minDisparity=minD, textureThreshold=10, numDisparities=nd,
                         uniquenessRatio=10, speckleWindowSize=100 if roi is None else 0, speckleRange=32,
                         disp12MaxDiff=1 if roi is None else -1, roi1=roi)
                assert np.array_equal(cv2_ref.make_bm(**p).compute(L, R), orc.bm_compute(L, R, _bm_params(orc, p))), (minD, W, roi)
            for mode in (0, 1):
                ps = dict(blockSize=5, minDisparity=minD, numDisparities=nd, uniquenessRatio=10, speckleWindowSize=100,
                          speckleRange=32, disp12MaxDiff=1, mode=mode)
                got, outside = orc.sgbm_compute(L, R, orc.sgbm_params(**ps), return_domain_flag=True)
                assert not outside
                assert np.array_equal(cv2_ref.make_sgbm(**ps).compute(L, R), got), (minD, W, mode)
