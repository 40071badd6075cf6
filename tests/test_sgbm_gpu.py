"""GPU parity: the CUDA semi-global matcher (through the C ABI) against the cv2 golden vectors and
the oracle.  Bit-exact CV_16S."""
import json

import numpy as np
import pytest

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu


def _mk(rt, p, W, H, **kw):
    return rt.CUDASemiGlobalMatcher(p["blockSize"], p["minDisparity"], p["numDisparities"], p["uniquenessRatio"],
                                    p["speckleWindowSize"], p["speckleRange"], p["disp12MaxDiff"],
                                    mode=p.get("mode", 0), P1=p.get("P1"), P2=p.get("P2"), max_width=W, max_height=H, **kw)


@pytest.mark.parametrize("name", golden_names("sgbm_"))
def test_sgbm_matches_cv2_golden(gpu, name):
    g = load_golden(name)
    p = json.loads(str(g["params"]))
    H, W = g["left"].shape
    m = _mk(gpu, p, W, H)
    got = m.compute(g["left"], g["right"])
    assert np.array_equal(got, g["disp"]), f"{name}: {(got != g['disp']).sum()} pixels differ"
    assert m.last_launches() >= 2 + 5 + 1 + 1        # planes, fused cost, paths, WTA (or LR check), median
    m.setROI1((1, 2, 3, 4)); m.setROI2(None)          # no-ops like the reference (sgbm-sw.h:32-33)
    assert np.array_equal(m.compute(g["left"], g["right"]), g["disp"])


def test_sgbm_random_params_match_oracle(gpu, orc):
    from rtdm_b200 import synth
    rng = np.random.default_rng(12)
    checked = 0
    for i in range(14):
        W, H = int(rng.integers(100, 400)), int(rng.integers(40, 220))
        nd = 16 * int(rng.integers(1, 11)); bs = 2 * int(rng.integers(0, 4)) + 1
        if nd + 8 >= W:
            continue
        p = dict(blockSize=bs, minDisparity=0, numDisparities=nd, uniquenessRatio=int(rng.integers(0, 25)),
                 speckleWindowSize=int(rng.integers(0, 150)), speckleRange=int(rng.integers(0, 8)),
                 disp12MaxDiff=int(rng.integers(-1, 4)), mode=int(rng.integers(0, 2)))
        L, R, _ = synth.stereo_pair(W, H, nd, 700 + i)
        ref, outside = orc.sgbm_compute(L, R, orc.sgbm_params(**p), return_domain_flag=True)
        if outside:
            continue
        got = _mk(gpu, p, W, H).compute(L, R)
        assert np.array_equal(ref, got), (p, W, H, int((ref != got).sum()))
        checked += 1
    assert checked >= 8


def test_sgbm_degenerate_width(gpu, orc):
    from rtdm_b200 import synth
    L, R, _ = synth.stereo_pair(100, 60, 16, 5)
    p = dict(blockSize=5, minDisparity=0, numDisparities=112, uniquenessRatio=10, speckleWindowSize=100,
             speckleRange=32, disp12MaxDiff=1, mode=0)
    got = _mk(gpu, p, 100, 60).compute(L, R)
    assert (got == -16).all()
    assert np.array_equal(got, orc.sgbm_compute(L, R, orc.sgbm_params(**p)))


@pytest.mark.parametrize("vpass", [False, True])
@pytest.mark.parametrize("mode", [0, 1])
def test_sgbm_720p_full_size(gpu, orc, mode, vpass, monkeypatch):
    """BASELINE config 3: 1280x720, nd=128, bs=5, P1=600, P2=2400; MODE_SGBM and MODE_HH; batch of 2, as tiled sweeps
    (what a batch this small takes) and forced through the whole-height pass (a cluster of 9 CTAs per frame)."""
    from rtdm_b200 import synth
    monkeypatch.delenv("RTDM_SGBM_NOVPASS", raising=False)
    if vpass:
        monkeypatch.setenv("RTDM_SGBM_VPASS_MIN", "2")
    else:
        monkeypatch.delenv("RTDM_SGBM_VPASS_MIN", raising=False)
    p = dict(blockSize=5, minDisparity=0, numDisparities=128, uniquenessRatio=10, speckleWindowSize=100,
             speckleRange=32, disp12MaxDiff=1, mode=mode)
    frames = [synth.stereo_pair(1280, 720, 128, 1000 + i) for i in range(2)]
    Ls = np.stack([f[0] for f in frames]); Rs = np.stack([f[1] for f in frames])
    m = _mk(gpu, p, 1280, 720, max_batch=2)
    out = m.compute_batch(Ls, Rs)
    for i in range(2):
        ref, outside = orc.sgbm_compute(Ls[i], Rs[i], orc.sgbm_params(**p), return_domain_flag=True)
        assert not outside
        assert np.array_equal(out[i], ref), (i, int((out[i] != ref).sum()))
        v = out[i][out[i] != -16]
        assert v.min() >= 0 and v.max() <= 127 * 16 + 15
        assert (out[i][:, :127] == -16).all()
    assert (m.last_launches() <= 12) == vpass


@pytest.mark.parametrize("env", [None, "RTDM_SGBM_NOSWEEP", "RTDM_SGBM_NOFUSE", "RTDM_SGBM_OLDCOST", "RTDM_SGBM_OLDPATH"])
def test_sgbm_kernel_variants_agree_with_oracle(gpu, orc, env, monkeypatch):
    """Every kernel choice of the matching stage (row sweeps / per-direction chains, WTA fused into the last path or
    separate, fused or two-pass cost volume, 4-word or generic path kernel) is the same arithmetic."""
    from rtdm_b200 import synth
    for v in ("RTDM_SGBM_NOSWEEP", "RTDM_SGBM_NOFUSE", "RTDM_SGBM_OLDCOST", "RTDM_SGBM_OLDPATH"):
        monkeypatch.delenv(v, raising=False)
    if env:
        monkeypatch.setenv(env, "1")
    for (W, H, nd, bs, mode) in [(400, 150, 128, 5, 1), (331, 97, 64, 3, 0), (520, 60, 64, 7, 1)]:
        p = dict(blockSize=bs, minDisparity=0, numDisparities=nd, uniquenessRatio=12, speckleWindowSize=50,
                 speckleRange=2, disp12MaxDiff=1, mode=mode)
        L, R, _ = synth.stereo_pair(W, H, nd, 4100 + W)
        ref, outside = orc.sgbm_compute(L, R, orc.sgbm_params(**p), return_domain_flag=True)
        assert not outside
        got = _mk(gpu, p, W, H).compute(L, R)
        assert np.array_equal(got, ref), (env, W, H, nd, bs, mode, int((got != ref).sum()))


@pytest.mark.parametrize("case", [
    # W, H, nd, bs, mode, batch, cap on clusters (None: what the device keeps resident), P2
    (400, 150, 128, 5, 1, 3, None, None),      # 3 CTAs per cluster (128 columns each), one frame per cluster
    (331, 97, 64, 3, 0, 2, None, None),        # D = 64: 256 columns per CTA, 2 CTAs
    (700, 64, 128, 5, 1, 5, 2, None),          # 5 CTAs, 5 frames over 2 clusters: 3 frames, then 2, per cluster
    (200, 51, 64, 5, 1, 4, 1, None),           # 1 CTA, odd height: the buffer parity flips from frame to frame
    (453, 33, 128, 3, 0, 7, 3, None),          # last CTA holds 69 of its 128 columns
    (360, 120, 64, 5, 0, 3, 1, 4400),          # step-wise clamp instantiation
    (500, 40, 192, 5, 1, 4, 2, None),          # D = 192 (the reference's default -nd): one pixel per warp, 6 disparities per lane, 4 columns per thread
    (420, 50, 96, 3, 0, 3, 1, None),           # D = 96: 16 lanes x 6 disparities
    (400, 45, 48, 5, 1, 5, 2, None),           # D = 48: 8 lanes x 6 disparities
    (330, 36, 192, 3, 0, 2, None, 4400),       # D = 192, step-wise clamp
])
def test_sgbm_cluster_pass_matches_oracle(gpu, orc, case, monkeypatch):
    """Batches take sgbm_vpass_kernel (one thread-block cluster per frame, neighbours exchanging boundary columns through
    distributed shared memory); single frames take the tiled sweeps.  Both must be the oracle's arithmetic, for every
    cluster size, with several frames per cluster, and with the last CTA partly outside the image."""
    from rtdm_b200 import synth
    W, H, nd, bs, mode, B, maxcl, P2 = case
    for v in ("RTDM_SGBM_NOSWEEP", "RTDM_SGBM_NOFUSE", "RTDM_SGBM_OLDCOST", "RTDM_SGBM_OLDPATH", "RTDM_SGBM_NOVPASS", "RTDM_SGBM_VPASS_MAXCL", "RTDM_SGBM_VPASS_MIN"):
        monkeypatch.delenv(v, raising=False)
    monkeypatch.setenv("RTDM_SGBM_VPASS_MIN", "2")               # by default small batches take the tiled sweeps
    if maxcl:
        monkeypatch.setenv("RTDM_SGBM_VPASS_MAXCL", str(maxcl))
    p = dict(blockSize=bs, minDisparity=0, numDisparities=nd, uniquenessRatio=12, speckleWindowSize=0 if P2 else 50,
             speckleRange=2, disp12MaxDiff=1, mode=mode)
    if P2:
        p.update(P1=700, P2=P2, uniquenessRatio=5)
    fr = [synth.stereo_pair(W, H, nd, 5200 + 7 * i + W) for i in range(B)]
    L = np.stack([f[0] for f in fr]); R = np.stack([f[1] for f in fr])
    m = _mk(gpu, p, W, H, max_batch=B)
    out = m.compute_batch(L, R)
    assert m.last_launches() <= 2 + 2 + 2 + 1 + 1 + 4            # planes, cost, 2 horizontal paths, 2 passes, LR, median, speckle: no tiled sweeps
    checked = 0
    for i in range(B):
        ref, outside = orc.sgbm_compute(L[i], R[i], orc.sgbm_params(**p), return_domain_flag=True)
        if outside:
            continue
        assert np.array_equal(out[i], ref), (case, i, int((out[i] != ref).sum()))
        assert np.array_equal(m.compute(L[i], R[i]), ref), (case, i, "tiled sweeps")
        checked += 1
    assert checked >= 1


def test_sgbm_cluster_pass_equals_tiled_sweeps_on_random_geometries(gpu, monkeypatch):
    """The two aggregation variants against each other (both are pinned to the oracle elsewhere): random widths (1 .. 5
    CTAs per cluster, partly filled last CTA), heights from one row up (the first row of a pass needs no neighbour data, the
    last row hands zeros to the next frame), D = 64 and 128, both modes, batches that give a cluster 1 .. 4 frames."""
    from rtdm_b200 import synth
    rng = np.random.default_rng(77)
    cases = [(300, 1, 128, 5, 1, 3, 2), (300, 2, 128, 5, 1, 3, 1), (200, 3, 64, 3, 0, 5, 2)]
    for _ in range(14):
        nd = int(rng.choice([48, 64, 96, 128, 192]))
        W = nd + int(rng.integers(8, 640)); H = int(rng.integers(4, 70))
        cases.append((W, H, nd, int(rng.choice([3, 5, 7])), int(rng.integers(0, 2)), int(rng.integers(2, 9)), int(rng.integers(1, 4))))
    for (W, H, nd, bs, mode, B, maxcl) in cases:
        fr = [synth.stereo_pair(W, H, nd, 9100 + 3 * i + W + H) for i in range(B)]
        L = np.stack([f[0] for f in fr]); R = np.stack([f[1] for f in fr])
        outs = []
        for novpass in (False, True):
            for v in ("RTDM_SGBM_NOVPASS", "RTDM_SGBM_VPASS_MIN", "RTDM_SGBM_VPASS_MAXCL"):
                monkeypatch.delenv(v, raising=False)
            if novpass:
                monkeypatch.setenv("RTDM_SGBM_NOVPASS", "1")
            else:
                monkeypatch.setenv("RTDM_SGBM_VPASS_MIN", "2"); monkeypatch.setenv("RTDM_SGBM_VPASS_MAXCL", str(maxcl))
            m = gpu.CUDASemiGlobalMatcher(bs, 0, nd, 10, 60, 4, 1, mode=mode, max_width=W, max_height=H, max_batch=B)
            outs.append(m.compute_batch(L, R).copy())
            if not novpass:
                assert m.last_launches() <= 12, (W, H, nd, bs, mode, B, maxcl, m.last_launches())     # one launch per pass, no tiles
        assert np.array_equal(outs[0], outs[1]), (W, H, nd, bs, mode, B, maxcl, int((outs[0] != outs[1]).sum()))


def test_sgbm_batch_quantum(gpu):
    """rtdm_sgbm_batch_quantum: the clusters in flight for sizes that take the whole-height pass (one frame each), 1 otherwise."""
    m = gpu.CUDASemiGlobalMatcher(5, 0, 128, 10, 100, 32, 1, mode=1, max_width=1280, max_height=720, max_batch=45)
    q = m.batch_quantum(1280, 720)
    assert 1 <= q <= 148 // 9 + 1
    assert m.batch_quantum(320, 240) >= q                      # fewer CTAs per cluster: at least as many clusters
    m96 = gpu.CUDASemiGlobalMatcher(5, 0, 96, 10, 100, 32, 1, max_width=640, max_height=480, max_batch=4)
    assert m96.batch_quantum(640, 480) == 1                    # D = 96: tiled / chain kernels
    with pytest.raises(gpu.RtdmError):
        m.batch_quantum(4000, 720)


@pytest.mark.parametrize("case", [(560, 70, 256, 5, 0), (700, 48, 256, 3, 1), (430, 64, 256, 7, 0), (300, 90, 192, 5, 1), (934, 404, 192, 5, 0),
                                  (320, 240, 48, 5, 1), (640, 200, 96, 7, 0), (250, 80, 96, 1, 1), (200, 60, 48, 3, 0)])
def test_sgbm_wide_disparity_ranges_match_oracle(gpu, orc, case):
    """numDisparities 256 (fused cost kernel with two 8-column segments per CTA, generic chain kernel for the paths) and the
    reference's default -nd 192 scaled to the frame width (192 / 96 / 48 at 1280 / 640 / 320 pixels; the operating point
    934x404x192): 6 disparities per lane in the specialised chain kernels, on a batch of 3."""
    from rtdm_b200 import synth
    W, H, nd, bs, mode = case
    p = dict(blockSize=bs, minDisparity=0, numDisparities=nd, uniquenessRatio=10, speckleWindowSize=60, speckleRange=4,
             disp12MaxDiff=1, mode=mode)
    fr = [synth.stereo_pair(W, H, nd, 7700 + i + W) for i in range(3)]
    L = np.stack([f[0] for f in fr]); R = np.stack([f[1] for f in fr])
    out = _mk(gpu, p, W, H, max_batch=3).compute_batch(L, R)
    checked = 0
    for i in range(3):
        ref, outside = orc.sgbm_compute(L[i], R[i], orc.sgbm_params(**p), return_domain_flag=True)
        if outside:
            continue
        assert np.array_equal(out[i], ref), (case, i, int((out[i] != ref).sum()))
        checked += 1
    assert checked >= 1


def test_sgbm_large_penalties_take_the_stepwise_clamp(gpu, orc):
    """P2 large enough that three path costs next to S could overflow 16 bits (2*P2 + bs^2*93 > 10922): the sweep
    kernel then clamps after every addition (cv::StereoSGBM's saturating adds) instead of once.  5-path mode, so
    that S itself stays below the saturation value (the oracle's domain)."""
    from rtdm_b200 import synth
    checked = 0
    for P2 in (4400, 4800):
        p = dict(blockSize=5, minDisparity=0, numDisparities=64, uniquenessRatio=5, speckleWindowSize=0,
                 speckleRange=0, disp12MaxDiff=1, mode=0, P1=700, P2=P2)
        L, R, _ = synth.stereo_pair(360, 120, 64, 99)
        ref, outside = orc.sgbm_compute(L, R, orc.sgbm_params(**p), return_domain_flag=True)
        if outside:
            continue
        got = _mk(gpu, p, 360, 120).compute(L, R)
        assert np.array_equal(got, ref), (P2, int((got != ref).sum()))
        checked += 1
    assert checked >= 1


def test_sgbm_errors(gpu):
    with pytest.raises(gpu.RtdmError) as e:
        gpu.CUDASemiGlobalMatcher(5, 0, 100, 10, 100, 32, 1)
    assert e.value.code == -gpu.EINVAL
    with pytest.raises(gpu.RtdmError):
        gpu.CUDASemiGlobalMatcher(4, 0, 64, 10, 100, 32, 1)


def test_sgbm_streaming_submissions(gpu, orc):
    """rtdm_sgbm_submit_batch keeps two batches in flight on alternating staging buffers: every submission's output
    equals the blocking call's, in any interleaving of wait_oldest / wait."""
    import torch
    from rtdm_b200 import synth
    W, H, nd, B = 320, 240, 64, 3
    p = dict(blockSize=5, minDisparity=0, numDisparities=nd, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1)
    batches = []
    for b in range(4):
        fr = [synth.stereo_pair(W, H, nd, 8100 + 10 * b + i) for i in range(B)]
        batches.append((torch.from_numpy(np.stack([f[0] for f in fr])).pin_memory().numpy(),
                        torch.from_numpy(np.stack([f[1] for f in fr])).pin_memory().numpy()))
    m = gpu.CUDASemiGlobalMatcher(p["blockSize"], 0, nd, 10, 100, 32, 1, mode=1, max_width=W, max_height=H, max_batch=B)
    ref = [m.compute_batch(L, R).copy() for L, R in batches]
    outs = [torch.empty((B, H, W), dtype=torch.int16).pin_memory().numpy() for _ in batches]
    for i, (L, R) in enumerate(batches):
        m.submit_batch(L, R, outs[i])
        if i > 0:
            m.wait_oldest()
            assert np.array_equal(outs[i - 1], ref[i - 1]), i
    m.wait()
    assert np.array_equal(outs[-1], ref[-1])
    op = orc.make_params(P1=600, P2=2400, preFilterCap=0, mode=1, **p)
    assert np.array_equal(ref[0][0], orc.sgbm_compute(batches[0][0][0], batches[0][1][0], op))


def test_sgbm_wave_sized_sub_batches_equal_single_frames(gpu):
    """A 40-frame 720p call runs as wave-sized sub-batches (37 + 3 on 148 SMs): every frame must equal the map the
    same matcher computes for it alone (sub-batch boundaries, frame offsets of all work buffers)."""
    from rtdm_b200 import synth
    W, H, nd, B = 1280, 720, 128, 40
    fr = [synth.stereo_pair(W, H, nd, 8300 + i) for i in range(5)]
    L = np.stack([fr[i % 5][0] for i in range(B)]); R = np.stack([fr[i % 5][1] for i in range(B)])
    m = gpu.CUDASemiGlobalMatcher(5, 0, nd, 10, 100, 32, 1, mode=1, max_width=W, max_height=H, max_batch=B)
    out = m.compute_batch(L, R)
    single = [m.compute(fr[i][0], fr[i][1]) for i in range(5)]
    for k in (0, 1, 17, 36, 37, 38, 39):
        assert np.array_equal(out[k], single[k % 5]), k
    assert len({out[k].tobytes() for k in range(0, B, 5)}) == 1            # the same frame at eight batch positions


@pytest.mark.parametrize("minD", [16, -16, 5, -37])
def test_sgbm_min_disparity_matches_oracle(gpu, orc, minD):
    """minDisparity != 0 (sgbm-sw.h:28-29).  The oracle is pinned against cv2 for it by the sgbm_*mind* fixtures: for
    minD >= 2 cv2's LR check lets never-written disp2 entries (they hold (minD - 1) * 16) pass its `>= minD` test."""
    from rtdm_b200 import synth
    rng = np.random.default_rng(2000 + minD)
    checked = 0
    for i, (W, H, nd, bs) in enumerate([(320, 240, 64, 5), (233, 157, 48, 3), (400, 200, 128, 5), (300, 90, 32, 7), (500, 64, 96, 1)]):
        p = dict(blockSize=bs, minDisparity=minD, numDisparities=nd, uniquenessRatio=int(rng.integers(0, 20)),
                 speckleWindowSize=100 * (i % 2), speckleRange=int(rng.integers(1, 8)), disp12MaxDiff=int(rng.integers(-1, 3)),
                 mode=i % 2)
        L, R, _ = synth.stereo_pair(W, H, nd, 9000 + 10 * i + minD)
        ref, outside = orc.sgbm_compute(L, R, orc.sgbm_params(**p), return_domain_flag=True)
        if outside:
            continue
        got = _mk(gpu, p, W, H).compute(L, R)
        assert np.array_equal(ref, got), (p, W, H, int((ref != got).sum()))
        checked += 1
    assert checked >= 4
