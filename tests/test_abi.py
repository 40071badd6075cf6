"""CPU: the C-ABI library loads, exports every symbol include/rtdm_b200.h declares, validates
parameters like OpenCV would, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "rtdm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rtdm_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound(rt):
    names = _declared_symbols()
    assert len(names) >= 30
    l = rt.lib()
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/rtdm_b200.h but not exported"
        assert n in rt.SIGNATURES, f"{n} has no ctypes signature"
    assert l.rtdm_abi_version() == 1


def test_defaults_match_reference_main(rt):
    p = rt.RtdmParams()
    rt.lib().rtdm_params_default_bm(C.byref(p))
    # main.cpp:134-135: SWMatcherKonolige(roif, roif, 31, 13, 0, 10, nd, nd, 10, 100, 32, 1)
    assert (p.preFilterCap, p.blockSize, p.minDisparity, p.textureThreshold, p.uniquenessRatio,
            p.speckleWindowSize, p.speckleRange, p.disp12MaxDiff) == (31, 13, 0, 10, 10, 100, 32, 1)
    assert p.preFilterType == rt.PREFILTER_XSOBEL and p.preFilterSize == 9
    rt.lib().rtdm_params_default_sgbm(C.byref(p))
    assert (p.P1, p.P2, p.blockSize, p.mode) == (600, 2400, 5, rt.MODE_SGBM)   # sgbm-sw.cpp:15-18


@pytest.mark.parametrize("field,value", [
    ("numDisparities", 100), ("numDisparities", 0), ("blockSize", 12), ("blockSize", 3),
    ("preFilterCap", 0), ("preFilterCap", 64), ("textureThreshold", -1), ("uniquenessRatio", -1),
    ("preFilterType", 2), ("preFilterSize", 4),
])
def test_bm_create_rejects_what_opencv_asserts_on(rt, field, value):
    p = rt.RtdmParams()
    rt.lib().rtdm_params_default_bm(C.byref(p))
    setattr(p, field, value)
    h = C.c_void_p()
    rc = rt.lib().rtdm_bm_create(C.byref(h), C.byref(p), 640, 480, 1, 0)
    assert rc == -rt.EINVAL and not h.value
    assert rt.lib().rtdm_last_error()


def test_bm_create_rejects_non_bitexact_domain(rt):
    p = rt.RtdmParams()
    rt.lib().rtdm_params_default_bm(C.byref(p))
    p.preFilterCap = 63                      # with disp12MaxDiff = 1: SURVEY.md App. B.2
    h = C.c_void_p()
    assert rt.lib().rtdm_bm_create(C.byref(h), C.byref(p), 640, 480, 1, 0) == -rt.EINVAL


def test_no_cpu_fallback(rt):
    if rt.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(rt.RtdmError) as e:
        rt.CUDAMatcherKonolige(None, None, 31, 13, 0, 10, 128, 128, 10, 100, 32, 1)
    assert e.value.code == -rt.ENODEV
    with pytest.raises(rt.RtdmError) as e:
        rt.CUDAMorphologicalFilter(64, 48, 8)
    assert e.value.code == -rt.ENODEV
    import numpy as np
    with pytest.raises(rt.RtdmError):
        rt.filter_speckles(np.zeros((8, 8), np.int16), -16, 10, 1)
    # the steps either side of the plugins have no CPU path either
    m1 = np.zeros((8, 8, 2), np.int16); m2 = np.zeros((8, 8), np.uint16)
    for make in (lambda: rt.CUDAObjectRegions(64, 48), lambda: rt.CUDAColorMask(m1, m2, (0, 0, 8, 8)),
                 lambda: rt.CUDARectifier(m1, m2, (0, 0, 8, 8)), lambda: rt.CUDADepthEpilogue(64, 48, 4),
                 lambda: rt.CUDASemiGlobalMatcher(5, 0, 64, 10, 100, 32, 1)):
        with pytest.raises(rt.RtdmError) as e:
            make()
        assert e.value.code == -rt.ENODEV


def test_mask_entry_points_validate_arguments(rt):
    """Argument checks of the mask front-end / back-end happen before any device work (same -EINVAL convention)."""
    import numpy as np
    m1 = np.zeros((8, 8, 2), np.int16); m2 = np.zeros((8, 8), np.uint16)
    for roi in [(-1, 0, 4, 4), (0, 0, 9, 8), (0, 0, 0, 4), (6, 6, 4, 4)]:
        with pytest.raises(rt.RtdmError) as e:
            rt.CUDAColorMask(m1, m2, roi)
        assert e.value.code == -rt.EINVAL, roi
    with pytest.raises(rt.RtdmError) as e:
        rt.CUDAColorMask(np.zeros((8, 8), np.int16), m2, (0, 0, 8, 8))
    assert e.value.code == -rt.EINVAL
    for args in [(0, 48, 16), (64, 0, 16), (64, 48, 0)]:
        with pytest.raises(rt.RtdmError) as e:
            rt.CUDAObjectRegions(*args)
        assert e.value.code == -rt.EINVAL, args


def test_product_does_not_reference_oracle():
    """The product tree must not import, link or call anything under oracle/."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "rt-depth-map_b200")):
        if os.sep + "build" in d:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                t = open(os.path.join(d, f), errors="replace").read()
                if re.search(r"(import\s+oracle|from\s+oracle|liboracle|orc_[a-z_]+\s*\()", t):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
