"""ctypes front-end for oracle/liboracle.so (plain-C restatement of the OpenCV routines that
rt-depth-map's SW plugins call; see stereo_oracle.c for file:line citations).

TEST INFRASTRUCTURE ONLY -- never imported by the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "preFilterType", "preFilterSize", "preFilterCap", "blockSize", "minDisparity",
        "numDisparities", "textureThreshold", "uniquenessRatio", "speckleWindowSize",
        "speckleRange", "disp12MaxDiff", "mode", "P1", "P2")] + [
        ("roi1", C.c_int * 4), ("roi2", C.c_int * 4)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("stereo_oracle.c", "sgbm_oracle.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def make_params(preFilterType=1, preFilterSize=9, preFilterCap=31, blockSize=13, minDisparity=0,
                numDisparities=128, textureThreshold=10, uniquenessRatio=10, speckleWindowSize=100,
                speckleRange=32, disp12MaxDiff=1, mode=0, P1=0, P2=0, roi1=None, roi2=None) -> OrcParams:
    p = OrcParams()
    p.preFilterType, p.preFilterSize, p.preFilterCap = preFilterType, preFilterSize, preFilterCap
    p.blockSize, p.minDisparity, p.numDisparities = blockSize, minDisparity, numDisparities
    p.textureThreshold, p.uniquenessRatio = textureThreshold, uniquenessRatio
    p.speckleWindowSize, p.speckleRange, p.disp12MaxDiff = speckleWindowSize, speckleRange, disp12MaxDiff
    p.mode, p.P1, p.P2 = mode, P1, P2
    for i in range(4):
        p.roi1[i] = int(roi1[i]) if roi1 is not None else 0
        p.roi2[i] = int(roi2[i]) if roi2 is not None else 0
    return p


def _u8(a):
    a = np.asarray(a)
    assert a.dtype == np.uint8 and a.ndim == 2 and a.strides[1] == 1
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def prefilter_xsobel(img, cap):
    img = _u8(img); H, W = img.shape
    out = np.empty((H, W), np.uint8)
    lib().orc_prefilter_xsobel(_p(img), C.c_int(img.strides[0]), _p(out), C.c_int(W), W, H, int(cap))
    return out


def prefilter_norm(img, winsize, cap):
    img = _u8(img); H, W = img.shape
    out = np.empty((H, W), np.uint8)
    lib().orc_prefilter_norm(_p(img), C.c_int(img.strides[0]), _p(out), C.c_int(W), W, H, int(winsize), int(cap))
    return out


def valid_roi(W, H, minD, nd, bs, roi1=None, roi2=None):
    r1 = (C.c_int * 4)(*(roi1 or (0, 0, 0, 0))); r2 = (C.c_int * 4)(*(roi2 or (0, 0, 0, 0)))
    out = (C.c_int * 4)()
    lib().orc_valid_roi(r1, r2, W, H, minD, nd, bs, out)
    return tuple(out)


def bm_core(Lp, Rp, row0, row1, cap, bs, minD, nd, texThr, uniq):
    """Raw WTA disparity + cost on PREFILTERED images for rows [row0,row1)."""
    Lp = _u8(Lp); Rp = _u8(Rp); H, W = Lp.shape
    assert Lp.strides[0] == Rp.strides[0]
    disp = np.full((H, W), (minD - 1) * 16, np.int16)
    cost = np.zeros((H, W), np.int16)
    lib().orc_bm_core(_p(Lp), _p(Rp), C.c_int(Lp.strides[0]), W, H, row0, row1, cap, bs, minD, nd,
                      texThr, uniq, _p(disp), W, _p(cost), W)
    return disp, cost


def bm_sad_row(Lp, Rp, y, bs, minD, nd):
    Lp = _u8(Lp); Rp = _u8(Rp); H, W = Lp.shape
    lofs = max(nd - 1 + minD, 0); rofs = -min(nd - 1 + minD, 0); W1 = W - rofs - nd + 1
    out = np.zeros((W1, nd), np.int32)
    lib().orc_bm_sad_row(_p(Lp), _p(Rp), C.c_int(Lp.strides[0]), W, H, y, bs, minD, nd, _p(out))
    return out


def validate_disparity(disp, cost, minD, nd, d12):
    disp = np.ascontiguousarray(disp, np.int16).copy(); cost = np.ascontiguousarray(cost, np.int16)
    H, W = disp.shape
    lib().orc_validate_disparity(_p(disp), W, _p(cost), W, W, H, minD, nd, d12)
    return disp


def filter_speckles(disp, newVal, maxSize, maxDiff):
    disp = np.ascontiguousarray(disp, np.int16).copy(); H, W = disp.shape
    lib().orc_filter_speckles(_p(disp), W, W, H, int(newVal), int(maxSize), int(maxDiff))
    return disp


def bm_compute(left, right, params: OrcParams):
    left = _u8(left); right = _u8(right); H, W = left.shape
    disp = np.empty((H, W), np.int16)
    rc = lib().orc_bm_compute(_p(left), C.c_int(left.strides[0]), _p(right), C.c_int(right.strides[0]),
                              W, H, C.byref(params), _p(disp), W)
    if rc:
        raise ValueError(f"orc_bm_compute: {rc}")
    return disp


def sgbm_compute(left, right, params: OrcParams, return_domain_flag=False):
    """-> disparity (and, optionally, True when the aggregated cost S saturated int16, i.e. the input
    is outside the bit-exact domain of SURVEY.md App. B.5)."""
    left = _u8(left); right = _u8(right); H, W = left.shape
    disp = np.empty((H, W), np.int16)
    rc = lib().orc_sgbm_compute(_p(left), C.c_int(left.strides[0]), _p(right), C.c_int(right.strides[0]),
                                W, H, C.byref(params), _p(disp), W)
    if rc < 0:
        raise ValueError(f"orc_sgbm_compute: {rc}")
    return (disp, bool(rc)) if return_domain_flag else disp


def sgbm_params(blockSize=5, minDisparity=0, numDisparities=128, uniquenessRatio=10, speckleWindowSize=100,
                speckleRange=32, disp12MaxDiff=1, mode=0, P1=8 * 3 * 5 * 5, P2=32 * 3 * 5 * 5) -> OrcParams:
    """Parameters as SWSemiGlobalMatcher's constructor sets them (sgbm-sw.cpp:15-24)."""
    return make_params(preFilterCap=0, blockSize=blockSize, minDisparity=minDisparity,
                       numDisparities=numDisparities, uniquenessRatio=uniquenessRatio,
                       speckleWindowSize=speckleWindowSize, speckleRange=speckleRange,
                       disp12MaxDiff=disp12MaxDiff, mode=mode, P1=P1, P2=P2)


def morph(img, op, kw=10, kh=10):
    img = _u8(img); H, W = img.shape
    out = np.empty((H, W), np.uint8)
    lib().orc_morph(_p(img), C.c_int(img.strides[0]), _p(out), W, W, H, kw, kh, int(op))
    return out


def morph_open_close(img, kw=10, kh=10):
    img = _u8(img); H, W = img.shape
    out = np.empty((H, W), np.uint8)
    lib().orc_morph_open_close(_p(img), C.c_int(img.strides[0]), _p(out), W, W, H, kw, kh)
    return out


def ellipse_rows(kw=10, kh=10):
    j1 = (C.c_int * kh)(); j2 = (C.c_int * kh)()
    lib().orc_ellipse_rows(kw, kh, j1, j2)
    return list(j1), list(j2)


def median3_s16(img):
    img = np.ascontiguousarray(img, np.int16); H, W = img.shape
    out = np.empty((H, W), np.int16)
    lib().orc_median3_s16(_p(img), W, _p(out), W, W, H)
    return out
