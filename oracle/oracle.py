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


# ---------------------------------------------------------------------------------------------------
# Depth epilogue: numpy restatement of estimator.cpp:75-77 and calc_depth (:206-263)
# ---------------------------------------------------------------------------------------------------
DEPTH_BIG_Z = 10000.0


def disp_div16(disp):
    """`left_disp /= 16.` on a CV_16S Mat (estimator.cpp:75): saturate_cast<short>(cvRound(d / 16.)), i.e.
    round-half-to-even of the x16 fixed-point disparity."""
    d = np.asarray(disp, np.int16).astype(np.int32)
    q, r = d >> 4, d & 15
    q = q + (r > 8) + ((r == 8) & ((q & 1) == 1))
    return q.astype(np.int16)


def reproject_to_3d(disp16, Q):
    """cv::reprojectImageTo3D(disp, xyz, Q, handleMissingValues = true, CV_32F) for a CV_16S disparity
    (estimator.cpp:76).  OpenCV's arithmetic: the disparity goes through float, the homogeneous point is the
    4x4 double product accumulated left to right from zero, its first three components are narrowed to float,
    then multiplied by the double reciprocal of the fourth and narrowed again; pixels at the minimum disparity
    of the whole image get Z = 10000."""
    d16 = np.asarray(disp16, np.int16)
    Q = np.asarray(Q, np.float64).reshape(4, 4)
    H, W = d16.shape
    x = np.arange(W, dtype=np.float64)[None, :].repeat(H, 0)
    y = np.arange(H, dtype=np.float64)[:, None].repeat(W, 1)
    d = d16.astype(np.float32).astype(np.float64)
    hom = []
    for i in range(4):
        s = Q[i, 0] * x
        s = s + Q[i, 1] * y
        s = s + Q[i, 2] * d
        s = s + Q[i, 3] * 1.0
        hom.append(s)
    out = np.empty((H, W, 3), np.float32)
    with np.errstate(all="ignore"):
        iw = 1.0 / hom[3]
        for i in range(3):
            out[..., i] = (hom[i].astype(np.float32).astype(np.float64) * iw).astype(np.float32)
    out[..., 2][np.abs(d - float(d16.min())) <= np.finfo(np.float32).eps] = np.float32(DEPTH_BIG_Z)
    return out


def calc_depth(xyz, mask, rects):
    """Estimator::calc_depth (estimator.cpp:206-263): per rectangle the mean Z over the pixels with mask != 0,
    Z != 10000 (within FLT_EPSILON) and |Z| <= 10000, accumulated in double in row-major order.
    Returns (mean_z [n] float64 -- 0 where no pixel qualifies --, count [n] int32)."""
    means, counts = [], []
    for (rx, ry, rw, rh) in rects:
        z = xyz[ry:ry + rh, rx:rx + rw, 2].astype(np.float64)
        m = np.asarray(mask)[ry:ry + rh, rx:rx + rw]
        with np.errstate(all="ignore"):
            skip = (np.abs(z - DEPTH_BIG_Z) < np.finfo(np.float32).eps) | (np.abs(z) > DEPTH_BIG_Z) | (m == 0)
        res, cnt = 0.0, 0
        for v in z[~skip].ravel():            # row-major, like the reference's two loops
            res += float(v); cnt += 1
        means.append(res / cnt if cnt else 0.0)
        counts.append(cnt)
    return np.asarray(means, np.float64), np.asarray(counts, np.int32)


def distance_cm(mean_z, calibration_unit):
    """The label the reference prints (estimator.cpp:252-254): res * calibrationUnit / 10, fixed, 0 decimals."""
    return f"{mean_z * calibration_unit / 10.0:.0f} cm"


# ---------------------------------------------------------------------------------------------------
# Rectification front-end: numpy restatement of estimator.cpp:29-36
# ---------------------------------------------------------------------------------------------------
def rgb2gray(rgb):
    """cvtColor(img, gray, CV_RGB2GRAY) on CV_8UC3 (estimator.cpp:29-30): OpenCV 4.x 15-bit fixed point,
    (R*9798 + G*19235 + B*3735 + 2^14) >> 15."""
    a = np.asarray(rgb, np.uint8).astype(np.int32)
    return ((a[..., 0] * 9798 + a[..., 1] * 19235 + a[..., 2] * 3735 + (1 << 14)) >> 15).astype(np.uint8)


def remap_linear_fixed(src, map1, map2):
    """remap(src, dst, map1, map2, INTER_LINEAR) for CV_8UC1 with the fixed-point maps initUndistortRectifyMap(...,
    CV_16SC2, ...) makes (estimator.cpp:32,35; main.cpp:95-96): map1 = integer source (x, y), map2 = fy << 5 | fx
    (5-bit fractions).  Weights (32-fy)(32-fx)*32 ... sum to 2^15 exactly, result (sum + 2^14) >> 15;
    BORDER_CONSTANT 0 outside the source."""
    src = np.asarray(src, np.uint8); sh, sw = src.shape
    m1 = np.asarray(map1, np.int16).astype(np.int64); m2 = np.asarray(map2, np.uint16).astype(np.int64)
    sx, sy = m1[..., 0], m1[..., 1]
    fx, fy = m2 & 31, (m2 >> 5) & 31

    def at(y, x):
        ok = (x >= 0) & (x < sw) & (y >= 0) & (y < sh)
        return np.where(ok, src[np.clip(y, 0, sh - 1), np.clip(x, 0, sw - 1)].astype(np.int64), 0)
    v = ((32 - fy) * (32 - fx) * 32) * at(sy, sx) + ((32 - fy) * fx * 32) * at(sy, sx + 1) \
        + (fy * (32 - fx) * 32) * at(sy + 1, sx) + (fy * fx * 32) * at(sy + 1, sx + 1)
    return ((v + (1 << 14)) >> 15).astype(np.uint8)


def rectify(rgb, map1, map2, roi):
    """gray -> remap -> crop to roi = (x, y, w, h): what the matcher receives as left_rect / right_rect."""
    x, y, w, h = roi
    return np.ascontiguousarray(remap_linear_fixed(rgb2gray(rgb), map1, map2)[y:y + h, x:x + w])


# ---------------------------------------------------------------------------------------------------
# Mask front-end and back-end: restatement of estimator.cpp:38-53 and :164-204
# ---------------------------------------------------------------------------------------------------
def _hsv_tables():
    """OpenCV's RGB2HSV_b tables (imgproc color_hsv: hsv_shift = 12): sdiv[i] = cvRound((255 << 12) / i),
    hdiv[i] = cvRound((180 << 12) / (6 i)); entry 0 is 0."""
    sd = np.zeros(256, np.int64); hd = np.zeros(256, np.int64)
    i = np.arange(1, 256, dtype=np.float64)
    sd[1:] = np.rint((255 << 12) / i).astype(np.int64)
    hd[1:] = np.rint((180 << 12) / (6.0 * i)).astype(np.int64)
    return sd, hd


def bgr2hsv(bgr):
    """cvtColor(img, hsv, COLOR_BGR2HSV) on CV_8UC3 (estimator.cpp:42): H in [0, 180), integer algorithm
    (v = max, diff = max - min, s = (diff * sdiv[v] + 2^11) >> 12, h from the sextant * hdiv[diff], + 180 if negative).
    Pinned over all 2^24 colours against cv2 4.13.0 (tests/test_oracle_golden.py)."""
    sd, hd = _hsv_tables()
    a = np.asarray(bgr, np.uint8).astype(np.int64)
    b, g, r = a[..., 0], a[..., 1], a[..., 2]
    v = np.maximum(np.maximum(b, g), r); diff = v - np.minimum(np.minimum(b, g), r)
    s = (diff * sd[v] + (1 << 11)) >> 12
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * hd[diff] + (1 << 11)) >> 12
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], -1).astype(np.uint8)


def in_range(img, low, high):
    """inRange(img, Scalar(low), Scalar(high), dst) on CV_8UC3 (estimator.cpp:43): 255 where every channel lies in
    [low, high], else 0."""
    a = np.asarray(img, np.uint8).astype(np.int64)
    lo = np.asarray(low, np.int64).reshape(1, 1, 3); hi = np.asarray(high, np.int64).reshape(1, 1, 3)
    return (((a >= lo) & (a <= hi)).all(-1) * 255).astype(np.uint8)


def color_mask(rgb, map1, map2, roi, low, high):
    """estimator.cpp:38-43: remap(img[0], INTER_LINEAR) on the 3-channel frame (same fixed-point weights per channel as
    remap_linear_fixed), crop to roif, RGB -> BGR, BGR -> HSV, inRange -> filter_in.  Returns (mask, bgr_rectified)."""
    x, y, w, h = roi
    rgb = np.asarray(rgb, np.uint8)
    rect = np.stack([remap_linear_fixed(rgb[..., c], map1, map2) for c in range(3)], -1)[y:y + h, x:x + w]
    bgr = np.ascontiguousarray(rect[..., ::-1])
    return in_range(bgr2hsv(bgr), low, high), bgr


def contour_boxes(mask):
    """findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) + boundingRect per top-level contour
    (estimator.cpp:47, :164-175), as (x, y, w, h) in OpenCV's order.
    A top-level contour is the outer border of an 8-connected component of non-zero pixels whose surrounding
    background (4-connected, with a virtual zero frame around the image) is the frame's; components inside a hole of
    another component are not listed.  OpenCV lists the contours in reverse order of discovery, discovery = raster
    order of the component's first pixel.  Pinned against cv2 4.13.0 on random masks (tests/test_oracle_golden.py)."""
    from scipy import ndimage
    m = np.asarray(mask) != 0
    H, W = m.shape
    pad = np.zeros((H + 2, W + 2), bool); pad[1:-1, 1:-1] = m
    fg, nf = ndimage.label(pad, structure=np.ones((3, 3), int))
    bg, _ = ndimage.label(~pad, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    frame = bg == bg[0, 0]
    near = np.zeros_like(pad)
    near[1:, :] |= frame[:-1, :]; near[:-1, :] |= frame[1:, :]; near[:, 1:] |= frame[:, :-1]; near[:, :-1] |= frame[:, 1:]
    ext = np.zeros(nf + 1, bool)
    ext[np.unique(fg[near & pad])] = True
    out = []
    for sl, lab in zip(ndimage.find_objects(fg), range(1, nf + 1)):
        if not ext[lab]:
            continue
        ys, xs = sl
        first = int(np.flatnonzero((fg[ys, xs] == lab).ravel())[0])
        fy, fx = divmod(first, xs.stop - xs.start)
        out.append((((ys.start - 1 + fy) * W + xs.start - 1 + fx), (xs.start - 1, ys.start - 1, xs.stop - xs.start, ys.stop - ys.start)))
    out.sort(key=lambda t: -t[0])
    return [b for _, b in out]


def object_regions(mask, min_obj_size):
    """fill_bounding_rects_of_contours + find_relevant_matching_region (estimator.cpp:164-204): the bounding boxes
    with area (w * h) >= min_obj_size, and the rectangle spanning them all (what bm->setROI1 receives).  With no
    box left the reference's roi is (1000000, 1000000, -2000000, -2000000); kept."""
    bounds = [b for b in contour_boxes(mask) if b[2] * b[3] >= min_obj_size]
    min_x = min_y = 1000000; max_x = max_y = -1000000
    for x, y, w, h in bounds:
        min_x = min(min_x, x); min_y = min(min_y, y); max_x = max(max_x, x + w); max_y = max(max_y, y + h)
    return bounds, (min_x, min_y, max_x - min_x, max_y - min_y)
