"""cv2 (OpenCV) configured exactly as rt-depth-map's SW plugins configure it.

TEST INFRASTRUCTURE ONLY.  Used (a) to pin oracle/stereo_oracle.c, (b) to generate the committed
golden fixtures under tests/golden/, (c) as the `reference` CPU arm of bench.py when importable
(this is the real OpenCV code the reference's bm-sw.cpp / sgbm-sw.cpp / mf-sw.cpp call into).
Parity is defined against cv2 4.13.0 only.
"""
from __future__ import annotations

import numpy as np

PINNED_VERSION = "4.13.0"


def have_cv2() -> bool:
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def cv2_pinned():
    import cv2
    if cv2.__version__ != PINNED_VERSION:
        raise RuntimeError(f"parity is pinned to cv2 {PINNED_VERSION}, found {cv2.__version__}")
    return cv2


def make_bm(preFilterCap=31, blockSize=13, minDisparity=0, textureThreshold=10, numDisparities=128,
            uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1,
            preFilterType=None, preFilterSize=None, roi1=None, roi2=None):
    """Mirror of SWMatcherKonolige::SWMatcherKonolige (bm-sw.cpp:12-26), same setter order."""
    cv2 = cv2_pinned()
    m = cv2.StereoBM_create(numDisparities, blockSize)
    m.setPreFilterCap(preFilterCap)
    m.setMinDisparity(minDisparity)
    m.setNumDisparities(numDisparities)
    m.setTextureThreshold(textureThreshold)
    m.setUniquenessRatio(uniquenessRatio)
    m.setSpeckleWindowSize(speckleWindowSize)
    m.setSpeckleRange(speckleRange)
    m.setDisp12MaxDiff(disp12MaxDiff)
    if preFilterType is not None:
        m.setPreFilterType(preFilterType)
    if preFilterSize is not None:
        m.setPreFilterSize(preFilterSize)
    if roi1 is not None:
        m.setROI1(tuple(int(v) for v in roi1))      # bm-sw.cpp:40-43
    if roi2 is not None:
        m.setROI2(tuple(int(v) for v in roi2))      # bm-sw.cpp:45-48
    return m


def make_sgbm(blockSize=5, minDisparity=0, numDisparities=128, uniquenessRatio=10,
              speckleWindowSize=100, speckleRange=32, disp12MaxDiff=1, mode=None):
    """Mirror of SWSemiGlobalMatcher::SWSemiGlobalMatcher (sgbm-sw.cpp:12-25)."""
    cv2 = cv2_pinned()
    m = cv2.StereoSGBM_create(0, numDisparities, blockSize)
    m.setP1(8 * 3 * 5 * 5)
    m.setP2(32 * 3 * 5 * 5)
    m.setMinDisparity(minDisparity)
    m.setNumDisparities(numDisparities)
    m.setUniquenessRatio(uniquenessRatio)
    m.setSpeckleWindowSize(speckleWindowSize)
    m.setSpeckleRange(speckleRange)
    m.setDisp12MaxDiff(disp12MaxDiff)
    if mode is not None:
        m.setMode(mode)        # 0 = MODE_SGBM (reference default), 1 = MODE_HH
    return m


def morph_open_close(img: np.ndarray, kw=10, kh=10) -> np.ndarray:
    """SWMorphologicalFilter::run (mf-sw.cpp:19-28): erode, dilate, dilate, erode."""
    cv2 = cv2_pinned()
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (kw, kh))
    out = cv2.erode(img, k)
    out = cv2.dilate(out, k)
    out = cv2.dilate(out, k)
    out = cv2.erode(out, k)
    return out
