"""GPU parity: the CUDA Konolige matcher (through the C ABI) against the oracle and the cv2 golden
vectors.  Bit-exact: CV_16S disparity maps must be identical."""
import json

import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["fast", "fast3", "fast2", "generic"])
def bm_kernel(request, monkeypatch):
    """Runs a test once per SAD/WTA kernel: the default selection (TMA-staged bm_sad4.cu where it applies, else
    bm_sad2.cu, else generic), bm_sad3.cu / bm_sad2.cu kept where bm_sad4.cu would run (RTDM_BM_KERNEL=3 / 2), and the
    generic kernel (bm_sad.cu, RTDM_BM_KERNEL=1)."""
    if request.param == "generic":
        monkeypatch.setenv("RTDM_BM_KERNEL", "1")
    elif request.param == "fast2":
        monkeypatch.setenv("RTDM_BM_KERNEL", "2")
    elif request.param == "fast3":
        monkeypatch.setenv("RTDM_BM_KERNEL", "3")
    else:
        monkeypatch.delenv("RTDM_BM_KERNEL", raising=False)
    return request.param


def _expected_kernel(bm_kernel, p):
    if bm_kernel == "generic" or p["blockSize"] > 15 or p["minDisparity"] != 0:
        return 1
    if bm_kernel in ("fast", "fast3") and p["numDisparities"] in (32, 48, 64, 96, 128, 192, 256):
        return 4 if bm_kernel == "fast" else 3
    return 2


def _mk(rt, p, W, H, **kw):
    m = rt.CUDAMatcherKonolige(None, None, p["preFilterCap"], p["blockSize"], p["minDisparity"],
                               p["textureThreshold"], p["numDisparities"], p["numDisparities"],
                               p["uniquenessRatio"], p["speckleWindowSize"], p["speckleRange"],
                               p["disp12MaxDiff"],
                               preFilterType=1 if p.get("preFilterType") is None else p["preFilterType"],
                               preFilterSize=9 if p.get("preFilterSize") is None else p["preFilterSize"],
                               max_width=W, max_height=H, **kw)
    if p.get("roi1") is not None:
        m.setROI1(p["roi1"])
    if p.get("roi2") is not None:
        m.setROI2(p["roi2"])
    return m


def _orc_params(orc, p):
    return orc.make_params(
        preFilterType=1 if p.get("preFilterType") is None else p["preFilterType"],
        preFilterSize=9 if p.get("preFilterSize") is None else p["preFilterSize"],
        preFilterCap=p["preFilterCap"], blockSize=p["blockSize"], minDisparity=p["minDisparity"],
        numDisparities=p["numDisparities"], textureThreshold=p["textureThreshold"],
        uniquenessRatio=p["uniquenessRatio"], speckleWindowSize=p["speckleWindowSize"],
        speckleRange=p["speckleRange"], disp12MaxDiff=p["disp12MaxDiff"], roi1=p.get("roi1"), roi2=p.get("roi2"))


@pytest.mark.parametrize("name", golden_names("bm_"))
def test_bm_matches_cv2_golden(gpu, name, bm_kernel):
    g = load_golden(name)
    p = json.loads(str(g["params"]))
    H, W = g["left"].shape
    m = _mk(gpu, p, W, H)
    got = m.compute(g["left"], g["right"])
    assert got.dtype == np.int16 and got.shape == (H, W)
    assert np.array_equal(got, g["disp"]), f"{name} [{bm_kernel}]: {(got != g['disp']).sum()} pixels differ"
    assert m.last_launches() > 0
    assert m.last_kernel() == _expected_kernel(bm_kernel, p), "kernel selection (generic / bm_sad2 / bm_sad3 / bm_sad4)"


def test_bm_stages_match_oracle(gpu, orc, bm_kernel):
    """Intermediates: prefiltered images, raw WTA disparity and cost."""
    from rtdm_b200 import synth
    W, H, nd, bs, cap = 320, 240, 64, 15, 31
    L, R, _ = synth.stereo_pair(W, H, nd, 77)
    p = dict(preFilterCap=cap, blockSize=bs, minDisparity=0, textureThreshold=10, numDisparities=nd,
             uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
    m = _mk(gpu, p, W, H)
    m.compute(L, R)
    Lp, Rp = orc.prefilter_xsobel(L, cap), orc.prefilter_xsobel(R, cap)
    assert np.array_equal(m.debug_fetch(0, W, H), Lp)
    assert np.array_equal(m.debug_fetch(1, W, H), Rp)
    h = bs // 2
    rd, rc = orc.bm_core(Lp, Rp, h, H - h, cap, bs, 0, nd, 10, 10)
    gd, gc = m.debug_fetch(2, W, H), m.debug_fetch(3, W, H)
    lofs = nd - 1
    assert np.array_equal(gd[h:H - h, lofs:], rd[h:H - h, lofs:])
    valid = rd[h:H - h, lofs:] >= 0
    assert np.array_equal(gc[h:H - h, lofs:][valid], rc[h:H - h, lofs:][valid])


def test_bm_random_params_match_oracle(gpu, orc, bm_kernel):
    from rtdm_b200 import synth
    rng = np.random.default_rng(11)
    checked = 0
    for i in range(16):
        W, H = int(rng.integers(70, 420)), int(rng.integers(40, 260))
        nd = 16 * int(rng.integers(1, 9)); bs = 2 * int(rng.integers(2, 11)) + 1
        if bs >= min(W, H):
            continue
        p = dict(preFilterCap=int(rng.integers(1, 32)), blockSize=bs, minDisparity=0,
                 textureThreshold=int(rng.integers(0, 50)), numDisparities=nd,
                 uniquenessRatio=int(rng.integers(0, 30)), speckleWindowSize=int(rng.integers(0, 200)),
                 speckleRange=int(rng.integers(0, 64)), disp12MaxDiff=int(rng.integers(-1, 4)),
                 preFilterType=int(rng.integers(0, 2)), preFilterSize=2 * int(rng.integers(2, 8)) + 1)
        L, R, _ = synth.stereo_pair(W, H, nd, 900 + i)
        ref = orc.bm_compute(L, R, _orc_params(orc, p))
        got = _mk(gpu, p, W, H).compute(L, R)
        assert np.array_equal(ref, got), (p, W, H, int((ref != got).sum()))
        checked += 1
    assert checked >= 10


@pytest.mark.parametrize("kern", [4, 3])
@pytest.mark.parametrize("bs", [5, 7, 9, 11, 13, 15])
@pytest.mark.parametrize("nd", [32, 48, 64, 96, 128, 192, 256])
def test_bm_warp_specialised_kernel_matrix(gpu, orc, nd, bs, kern, monkeypatch):
    """bm_sad4.cu (TMA-staged) and bm_sad3.cu over their whole domain (blockSize 5 .. 15 x numDisparities 32 / 48 / 64 / 96 / 128 / 192 / 256;
    192, 96 and 48 are what the reference's default -nd 192 scales to at 1280, 640 and 320 pixels of width): stripe borders (clamped
    columns on both image sides), widths that are not a multiple of anything, odd heights, ROIs, texture and
    uniqueness thresholds on and off, and every prefilter cap parity; raw WTA output and cost against the oracle's
    core as well as the final map."""
    from rtdm_b200 import synth
    if kern == 3:
        monkeypatch.setenv("RTDM_BM_KERNEL", "3")
    else:
        monkeypatch.delenv("RTDM_BM_KERNEL", raising=False)
    rng = np.random.default_rng(100 * nd + bs)
    for i, (W, H) in enumerate([(nd + 40, 61), (333, 127), (640, 203), (nd + bs + 3, 40)]):
        p = dict(preFilterCap=int(rng.integers(1, 32)), blockSize=bs, minDisparity=0,
                 textureThreshold=int(rng.integers(0, 40)) * (i % 2), numDisparities=nd,
                 uniquenessRatio=int(rng.integers(0, 25)) * ((i + 1) % 2 + i // 2), speckleWindowSize=100, speckleRange=32,
                 disp12MaxDiff=int(rng.integers(-1, 3)))
        if i == 2:
            p["roi1"] = (17, 9, W - 40, H - 20)
        L, R, _ = synth.stereo_pair(W, H, nd, 4000 + 10 * nd + bs + i)
        m = _mk(gpu, p, W, H)
        got = m.compute(L, R)
        assert m.last_kernel() == kern
        assert np.array_equal(got, orc.bm_compute(L, R, _orc_params(orc, p))), (p, W, H)
        if p.get("roi1") is None:
            Lp, Rp = orc.prefilter_xsobel(L, p["preFilterCap"]), orc.prefilter_xsobel(R, p["preFilterCap"])
            h = bs // 2
            rd, rc = orc.bm_core(Lp, Rp, h, H - h, p["preFilterCap"], bs, 0, nd, p["textureThreshold"], p["uniquenessRatio"])
            gd, gc = m.debug_fetch(2, W, H), m.debug_fetch(3, W, H)
            lofs = nd - 1
            assert np.array_equal(gd[h:H - h, lofs:], rd[h:H - h, lofs:]), (p, W, H)
            ok = rd[h:H - h, lofs:] >= 0
            assert np.array_equal(gc[h:H - h, lofs:][ok], rc[h:H - h, lofs:][ok]), (p, W, H)


def test_bm_degenerate_width(gpu, orc):
    """numDisparities wider than the image: the whole map is FILTERED."""
    from rtdm_b200 import synth
    L, R, _ = synth.stereo_pair(100, 60, 16, 5)
    p = dict(preFilterCap=31, blockSize=9, minDisparity=0, textureThreshold=10, numDisparities=128,
             uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
    got = _mk(gpu, p, 100, 60).compute(L, R)
    assert np.array_equal(got, orc.bm_compute(L, R, _orc_params(orc, p)))
    assert (got == -16).all()


def test_bm_strided_roi_views(gpu, orc, bm_kernel):
    """Estimator feeds non-contiguous ROI views (estimator.cpp:33,36): row step = full image width."""
    from rtdm_b200 import synth
    Lf, Rf, _ = synth.stereo_pair(400, 300, 64, 21)
    L, R = Lf[20:260, 30:350], Rf[20:260, 30:350]
    assert not L.flags.c_contiguous
    p = dict(preFilterCap=31, blockSize=13, minDisparity=0, textureThreshold=10, numDisparities=64,
             uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
    got = _mk(gpu, p, 320, 240).compute(L, R)
    ref = orc.bm_compute(np.ascontiguousarray(L), np.ascontiguousarray(R), _orc_params(orc, p))
    assert np.array_equal(got, ref)


def test_bm_set_roi_per_frame(gpu, orc, bm_kernel):
    """setROI1 is called every frame (estimator.cpp:54); ROI only changes the valid rectangle."""
    from rtdm_b200 import synth
    L, R, _ = synth.stereo_pair(320, 240, 64, 31)
    p = dict(preFilterCap=31, blockSize=13, minDisparity=0, textureThreshold=10, numDisparities=64,
             uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
    m = _mk(gpu, p, 320, 240)
    for roi in [(50, 40, 200, 150), (0, 0, 320, 240), (100, 100, 150, 100), None]:
        m.setROI1(roi)
        q = dict(p, roi1=roi)
        assert np.array_equal(m.compute(L, R), orc.bm_compute(L, R, _orc_params(orc, q))), roi


def test_bm_720p_full_size_and_batch(gpu, orc, bm_kernel):
    """BASELINE config 2 size: 1280x720, nd=128, reference parameters; batched call == per-frame."""
    from rtdm_b200 import synth
    p = dict(preFilterCap=31, blockSize=13, minDisparity=0, textureThreshold=10, numDisparities=128,
             uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
    frames = [synth.stereo_pair(1280, 720, 128, 1000 + i) for i in range(3)]
    Ls = np.stack([f[0] for f in frames]); Rs = np.stack([f[1] for f in frames])
    m = _mk(gpu, p, 1280, 720, max_batch=3)
    out = m.compute_batch(Ls, Rs)
    for i in range(3):
        ref = orc.bm_compute(Ls[i], Rs[i], _orc_params(orc, p))
        assert np.array_equal(out[i], ref), (i, int((out[i] != ref).sum()))
        # size-independent properties: invalid marker, range, left border
        assert (out[i][:, :127 + 6] == -16).all()
        v = out[i][out[i] != -16]
        assert v.min() >= 0 and v.max() <= 127 * 16 + 15
    assert np.array_equal(m.compute(Ls[1], Rs[1]), out[1])


def test_bm_streaming_submissions(gpu, orc):
    """rtdm_bm_submit_batch keeps two batches in flight on alternating staging buffers: every submission's output
    must equal the blocking call's (and the oracle's), whatever the order of waits."""
    from rtdm_b200 import synth
    p = dict(preFilterCap=31, blockSize=9, minDisparity=0, textureThreshold=10, numDisparities=64,
             uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
    W, H, B, S = 320, 240, 4, 5
    batches = []
    for s in range(S):
        fr = [synth.stereo_pair(W, H, 64, 7000 + 10 * s + i) for i in range(B)]
        batches.append((np.stack([f[0] for f in fr]), np.stack([f[1] for f in fr])))
    m = _mk(gpu, p, W, H, max_batch=B)
    want = [m.compute_batch(L, R) for L, R in batches]
    assert np.array_equal(want[0][0], orc.bm_compute(batches[0][0][0], batches[0][1][0], _orc_params(orc, p)))
    outs = [np.full((B, H, W), 12345, np.int16) for _ in range(S)]
    for s, (L, R) in enumerate(batches):
        m.submit_batch(L, R, outs[s])
        if s >= 1:
            m.wait_oldest()
            assert np.array_equal(outs[s - 1], want[s - 1]), s - 1
    m.wait()
    assert np.array_equal(outs[S - 1], want[S - 1])
    # a blocking call after streaming, and wait() with nothing in flight
    assert np.array_equal(m.compute_batch(*batches[2]), want[2])
    m.wait(); m.wait_oldest()
    # three submissions back to back: the third blocks on the first internally
    for s in range(3):
        outs[s][:] = 0
        m.submit_batch(batches[s][0], batches[s][1], outs[s])
    m.wait()
    for s in range(3):
        assert np.array_equal(outs[s], want[s]), s


def test_bm_errors(gpu):
    with pytest.raises(gpu.RtdmError) as e:
        gpu.CUDAMatcherKonolige(None, None, 31, 12, 0, 10, 128, 128, 10, 100, 32, 1)
    assert e.value.code == -gpu.EINVAL
    m = gpu.CUDAMatcherKonolige(None, None, 31, 13, 0, 10, 64, 64, 10, 100, 32, 1, max_width=64, max_height=64)
    with pytest.raises(gpu.RtdmError):
        m.compute(np.zeros((100, 100), np.uint8), np.zeros((100, 100), np.uint8))   # larger than the handle
    with pytest.raises(gpu.RtdmError):
        m.compute(np.zeros((10, 10), np.uint8), np.zeros((10, 10), np.uint8))       # block larger than image


def test_bm_side_by_side_frame(gpu, orc):
    """BASELINE config 2 input shape: a ZED-style side-by-side frame (left | right in one 2W x H buffer).  The
    two halves are passed as views with row step 2W -- no repacking on the host."""
    from rtdm_b200 import synth
    W, H, nd = 640, 360, 64
    L, R, _ = synth.stereo_pair(W, H, nd, 77)
    sbs = np.concatenate([L, R], axis=1)
    assert sbs.shape == (H, 2 * W)
    p = dict(preFilterCap=31, blockSize=13, minDisparity=0, textureThreshold=10, numDisparities=nd,
             uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
    got = _mk(gpu, p, W, H).compute(sbs[:, :W], sbs[:, W:])
    assert np.array_equal(got, orc.bm_compute(L, R, _orc_params(orc, p)))


def test_bm_large_frame_256_disparities(gpu, orc):
    """Largest disparity count the kernels accept (256) on a 2560x1440 frame: checks index arithmetic far from the
    720p case (stripes, bands, 32-octet tasks) against the oracle."""
    from rtdm_b200 import synth
    W, H, nd = 2560, 1440, 256
    L, R, _ = synth.stereo_pair(W, H, nd, 4242)
    for bs in (9, 21):
        p = dict(preFilterCap=31, blockSize=bs, minDisparity=0, textureThreshold=10, numDisparities=nd,
                 uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
        got = _mk(gpu, p, W, H).compute(L, R)
        ref = orc.bm_compute(L, R, _orc_params(orc, p))
        assert np.array_equal(got, ref), (bs, int((got != ref).sum()))


@pytest.mark.parametrize("shape", ["tma", "wide", "pair"])
def test_bm_warp_specialised_kernel_is_deterministic_under_load(gpu, orc, shape, monkeypatch):
    """The producer / winner-take-all warps of bm_sad4.cu / bm_sad3.cu hand rows over through named barriers, mbarriers and
    a shared-memory ring (fed by TMA in bm_sad4.cu, by loader warps in bm_sad3.cu): a protocol slip would show up as a
    rare, timing-dependent difference.  24 launches of a 12-frame 720p batch (bm_sad4.cu and both CTA shapes of
    bm_sad3.cu) must all equal the oracle's maps."""
    from rtdm_b200 import synth
    if shape == "tma":
        monkeypatch.delenv("RTDM_BM_KERNEL", raising=False)
    else:
        monkeypatch.setenv("RTDM_BM_KERNEL", "3")
    monkeypatch.setenv("RTDM_BM3_SHAPE", "1" if shape == "pair" else "0")
    p = dict(preFilterCap=31, blockSize=13, minDisparity=0, textureThreshold=10, numDisparities=128,
             uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
    frames = [synth.stereo_pair(1280, 720, 128, 1000 + i) for i in range(3)]
    ref = [orc.bm_compute(f[0], f[1], _orc_params(orc, p)) for f in frames]
    B = 12
    Ls = np.stack([frames[i % 3][0] for i in range(B)]); Rs = np.stack([frames[i % 3][1] for i in range(B)])
    m = _mk(gpu, p, 1280, 720, max_batch=B)
    for rep in range(24):
        out = m.compute_batch(Ls, Rs)
        assert m.last_kernel() == (4 if shape == "tma" else 3)
        for i in range(B):
            assert np.array_equal(out[i], ref[i % 3]), (shape, rep, i, int((out[i] != ref[i % 3]).sum()))


@pytest.mark.parametrize("minD", [16, -16, 5, -37, 40])
def test_bm_min_disparity_matches_oracle(gpu, orc, minD):
    """minDisparity is a constructor argument of the reference peer (bm-sw.h:28-30) that main.cpp leaves at 0; the
    library accepts any value (generic kernel), so every value must be right.  The oracle is pinned against cv2 for
    these values by the bm_mind* fixtures, including cv2's row spill for minD > 0 (the last minD
    computed columns of the last valid row land in the first pixels of the row below the valid rectangle) -- the
    whole map is compared, not only the valid rectangle."""
    from rtdm_b200 import synth
    rng = np.random.default_rng(1000 + minD)
    for i, (W, H, nd, bs) in enumerate([(320, 240, 64, 15), (233, 157, 48, 9), (400, 200, 128, 13), (640, 120, 32, 21)]):
        p = dict(preFilterCap=int(rng.integers(1, 32)), blockSize=bs, minDisparity=minD, textureThreshold=10 * (i % 2),
                 numDisparities=nd, uniquenessRatio=int(rng.integers(0, 20)), speckleWindowSize=100 * ((i + 1) % 2),
                 speckleRange=32, disp12MaxDiff=int(rng.integers(-1, 3)))
        if i in (1, 2):
            p["roi1"] = (30, 20, W - 70, H - 50)        # valid rectangle ends above the last row: the spill row stays visible
        L, R, _ = synth.stereo_pair(W, H, nd, 8000 + 10 * i + minD)
        m = _mk(gpu, p, W, H)
        got = m.compute(L, R)
        ref = orc.bm_compute(L, R, _orc_params(orc, p))
        assert m.last_kernel() == 1
        assert np.array_equal(got, ref), (p, W, H, int((got != ref).sum()), np.argwhere(got != ref)[:5])
        # batch of 3 (one spill row per frame)
        if i == 1:
            mb = _mk(gpu, p, W, H, max_batch=3)
            fr = [synth.stereo_pair(W, H, nd, 8100 + k) for k in range(3)]
            out = mb.compute_batch(np.stack([f[0] for f in fr]), np.stack([f[1] for f in fr]))
            for k in range(3):
                assert np.array_equal(out[k], orc.bm_compute(fr[k][0], fr[k][1], _orc_params(orc, p))), (minD, k)


def test_bm_reference_operating_point(gpu, orc, bm_kernel):
    """What Estimator::run really does per frame (estimator.cpp:54-56, main.cpp:131-135): the 934x404 calibrated ROI crop
    as a strided view of the 1280x720 rectified image, -nd 192, setROI1 with a fresh rectangle every frame, one frame,
    synchronous.  Against the cv2 fixture for one rectangle and the oracle for the others."""
    g = load_golden("bm_op_934x404_nd192_bs13")
    p = json.loads(str(g["params"]))
    H, W = g["left"].shape
    fullL = np.zeros((720, 1280), np.uint8); fullR = np.zeros((720, 1280), np.uint8)
    fullL[150:150 + H, 170:170 + W] = g["left"]; fullR[150:150 + H, 170:170 + W] = g["right"]
    Lv, Rv = fullL[150:150 + H, 170:170 + W], fullR[150:150 + H, 170:170 + W]      # step = full image width (estimator.cpp:33,36)
    m = _mk(gpu, dict(p, roi1=None), W, H)
    m.setROI1(p["roi1"])
    got = m.compute(Lv, Rv)
    assert np.array_equal(got, g["disp"]), int((got != g["disp"]).sum())
    assert m.last_kernel() == _expected_kernel(bm_kernel, p)
    for roi in [(0, 0, W, H), (400, 100, 300, 200), (600, 10, 334, 390), (10, 300, 900, 104)]:
        m.setROI1(roi)
        ref = orc.bm_compute(g["left"], g["right"], _orc_params(orc, dict(p, roi1=roi)))
        assert np.array_equal(m.compute(Lv, Rv), ref), roi


def test_bm_streaming_63_frame_720p_batches(gpu, orc):
    """The exact call pattern bench.py times: rtdm_bm_submit_batch with 63-frame 1280x720 batches (the host pipeline
    splits them into wave-sized chunks), two submissions in flight, results read one submission later.  Three
    submissions of the same 63 DISTINCT frames in rotated order: all must agree frame by frame, and the frames next to
    every possible chunk boundary must equal the oracle's maps."""
    from rtdm_b200 import synth
    p = dict(preFilterCap=31, blockSize=13, minDisparity=0, textureThreshold=10, numDisparities=128,
             uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
    W, H, B = 1280, 720, 63
    fr = [synth.stereo_pair(W, H, 128, 1000 + i) for i in range(B)]
    Ls = np.stack([f[0] for f in fr]); Rs = np.stack([f[1] for f in fr])
    m = _mk(gpu, p, W, H, max_batch=B)
    outs, ins = [], []
    for s in range(3):
        k = 7 * s
        Lr, Rr = np.roll(Ls, k, axis=0), np.roll(Rs, k, axis=0)        # frame i of the batch = frame (i - k) % B of the set
        ins.append((Lr, Rr))
        outs.append(np.full((B, H, W), 12345, np.int16))
        m.submit_batch(Lr, Rr, outs[s])
        if s >= 1:
            m.wait_oldest()
    m.wait()
    assert m.last_kernel() == 4
    for s in (1, 2):
        assert np.array_equal(np.roll(outs[s], -7 * s, axis=0), outs[0]), s
    for i in (0, 1, 15, 16, 20, 21, 27, 28, 31, 32, 34, 35, 41, 42, 47, 48, 62):
        ref = orc.bm_compute(Ls[i], Rs[i], _orc_params(orc, p))
        assert np.array_equal(outs[0][i], ref), (i, int((outs[0][i] != ref).sum()))


def test_bm_wide_frame_fused_row_kernel(gpu, orc):
    """Rows wider than 4912 pixels: the fused validate + row-run kernel needs more than 48 KB of dynamic shared memory."""
    from rtdm_b200 import synth
    W, H, nd = 5120, 48, 64
    L, R, _ = synth.stereo_pair(W, H, nd, 31337)
    p = dict(preFilterCap=31, blockSize=9, minDisparity=0, textureThreshold=10, numDisparities=nd,
             uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
    got = _mk(gpu, p, W, H).compute(L, R)
    assert np.array_equal(got, orc.bm_compute(L, R, _orc_params(orc, p)))
