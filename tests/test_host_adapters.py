"""The C++ plugin peers (rt-depth-map_b200/host/rtdm_plugins.h) compiled with g++ against the C ABI
and driven like Estimator drives the reference plugins."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "rt-depth-map_b200")
EXE = os.path.join(PKG, "build", "host_adapter_main")


@pytest.fixture(scope="module")
def exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(PKG, "host"),
                           os.path.join(ROOT, "tests", "cpp", "host_adapter_main.cpp"), "-o", EXE,
                           "-L", PKG, "-lrtdm_b200", "-Wl,-rpath," + PKG])
    return EXE


def test_adapters_compile_and_fail_loudly_without_device(exe, rt):
    r = subprocess.run([exe, "probe"], capture_output=True, text=True)
    if rt.device_count() == 0:
        assert r.returncode == 3 and "nodevice" in r.stdout and "no CUDA device" in r.stdout
    else:
        assert r.returncode == 0 and "ok" in r.stdout


@pytest.mark.gpu
def test_adapters_match_oracle(exe, gpu, orc, tmp_path):
    from rtdm_b200 import synth
    W, H, nd, bs = 320, 240, 64, 13
    L, R, _ = synth.stereo_pair(W, H, nd, 1234)
    lp, rp, op = (str(tmp_path / n) for n in ("l.raw", "r.raw", "o.raw"))
    L.tofile(lp); R.tofile(rp)
    roi = (40, 30, 240, 180)
    subprocess.check_call([exe, "bm", str(W), str(H), str(nd), str(bs), lp, rp, op] + [str(v) for v in roi])
    got = np.fromfile(op, np.int16).reshape(H, W)
    ref = orc.bm_compute(L, R, orc.make_params(blockSize=bs, numDisparities=nd, roi1=roi))
    assert np.array_equal(got, ref)
    # the row-band peer (rtdm_bm_rowband_*): three bands of the same frame
    subprocess.check_call([exe, "bm", str(W), str(H), str(nd), str(bs), lp, rp, op] + [str(v) for v in roi],
                          env=dict(os.environ, RTDM_TEST_ROWBANDS="3"))
    assert np.array_equal(np.fromfile(op, np.int16).reshape(H, W), ref)
    for mode in (0, 1):
        subprocess.check_call([exe, "sgbm", str(W), str(H), str(nd), "5", lp, rp, op, str(mode)])
        got = np.fromfile(op, np.int16).reshape(H, W)
        assert np.array_equal(got, orc.sgbm_compute(L, R, orc.sgbm_params(numDisparities=nd, mode=mode)))
    m = synth.binary_mask(W, H, 99)
    mp = str(tmp_path / "m.raw"); m.tofile(mp)
    subprocess.check_call([exe, "morph", str(W), str(H), mp, op])
    assert np.array_equal(np.fromfile(op, np.uint8).reshape(H, W), orc.morph_open_close(m))
    # depth epilogue peer (estimator.cpp:75-77) on the matcher's and the filter's outputs
    dp, tp = str(tmp_path / "d.raw"), str(tmp_path / "depth.txt")
    ref.tofile(dp); orc.morph_open_close(m).tofile(mp)
    rect = (50, 40, 200, 150)
    subprocess.check_call([exe, "depth", str(W), str(H), dp, mp, tp] + [str(v) for v in rect])
    Q = np.array([[1, 0, 0, -W / 2.0 + 0.37], [0, 1, 0, -H / 2.0 - 0.21], [0, 0, 0, 0.8 * W], [0, 0, 1 / 119.87, 0.004]])
    rm, rc = orc.calc_depth(orc.reproject_to_3d(orc.disp_div16(ref), Q), orc.morph_open_close(m), [rect, (0, 0, W, H)])
    rows = [l.split() for l in open(tp).read().splitlines()]
    assert [int(r[1]) for r in rows] == list(rc)
    assert np.allclose([float(r[0]) for r in rows], rm, rtol=1e-12, atol=0)
    # rectifier peer (estimator.cpp:29-36)
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
    m1 = np.stack([np.clip(np.arange(W)[None, :] + rng.integers(-2, 3, (H, W)), -1, W), np.clip(np.arange(H)[:, None] + rng.integers(-2, 3, (H, W)), -1, H)], -1).astype(np.int16)
    m2 = rng.integers(0, 1024, (H, W)).astype(np.uint16)
    paths = [str(tmp_path / n) for n in ("rgb.raw", "m1.raw", "m2.raw", "rect.raw")]
    rgb.tofile(paths[0]); m1.tofile(paths[1]); m2.tofile(paths[2])
    subprocess.check_call([exe, "rectify", str(W), str(H)] + paths)
    got = np.fromfile(paths[3], np.uint8).reshape(H - 4, W - 4)
    assert np.array_equal(got, orc.rectify(rgb, m1, m2, (2, 2, W - 4, H - 4)))


@pytest.mark.gpu
def test_mask_adapters_match_oracle(exe, gpu, orc, tmp_path):
    """CUDAColorMask -> CUDAMorphologicalFilter (plugin-owned buffers) -> CUDAObjectRegions, driven from C++ the way
    Estimator::run would (estimator.cpp:38-53)."""
    import cv2
    rng = np.random.default_rng(9)
    W, H = 200, 150
    base = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3)).astype(np.float32), (0, 0), 7.0)
    rgb = np.clip((base - base.mean()) / base.std() * 70 + 128 + rng.integers(-5, 6, (H, W, 3)), 0, 255).astype(np.uint8)
    m1 = np.stack([np.clip(np.arange(W)[None, :] + rng.integers(-1, 2, (H, W)), -1, W), np.clip(np.arange(H)[:, None] + rng.integers(-1, 2, (H, W)), -1, H)], -1).astype(np.int16)
    m2 = rng.integers(0, 1024, (H, W)).astype(np.uint16)
    paths = [str(tmp_path / n) for n in ("rgb.raw", "m1.raw", "m2.raw", "mask.raw", "boxes.txt")]
    rgb.tofile(paths[0]); m1.tofile(paths[1]); m2.tofile(paths[2])
    subprocess.check_call([exe, "mask", str(W), str(H)] + paths)
    roi = (2, 2, W - 4, H - 4)
    ref_mask, _ = orc.color_mask(rgb, m1, m2, roi, (20, 40, 40), (130, 255, 255))
    assert np.array_equal(np.fromfile(paths[3], np.uint8).reshape(H - 4, W - 4), ref_mask)
    out = orc.morph_open_close(ref_mask)
    bounds, span = orc.object_regions(out, 60)
    rows = [tuple(int(v) for v in l.split()) for l in open(paths[4]).read().splitlines()]
    assert rows[0] == (len(orc.contour_boxes(out)),) + span
    assert rows[1:] == bounds and len(bounds) >= 1
