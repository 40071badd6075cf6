"""rtdm_b200 -- Python mirror of rt-depth-map's matcher / filter plugin interface over the C ABI
of librtdm_b200.so (include/rtdm_b200.h).

The classes keep the reference's names, constructor argument lists and method names:

  CUDAMatcherKonolige      peer of SWMatcherKonolige      (reference include/stereo-matcher/bm-sw.h:25-37)
  CUDASemiGlobalMatcher    peer of SWSemiGlobalMatcher    (reference include/stereo-matcher/sgbm-sw.h:25-36)
  CUDAMorphologicalFilter  peer of SWMorphologicalFilter  (reference include/filter/mf-sw.h:16-21,
                                                           base class include/filter/filter.h:13-37)

Everything computes on the GPU through the shared library.  There is NO CPU fallback: if the library
is missing, or no CUDA device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_PKG_DIR), "librtdm_b200.so")

EIO, ENOMEM, ENODEV, EINVAL, ENOSYS = 5, 12, 19, 22, 38
PREFILTER_NORMALIZED_RESPONSE, PREFILTER_XSOBEL = 0, 1
MODE_SGBM, MODE_HH = 0, 1


class RtdmParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "preFilterType", "preFilterSize", "preFilterCap", "blockSize", "minDisparity",
        "numDisparities", "textureThreshold", "uniquenessRatio", "speckleWindowSize",
        "speckleRange", "disp12MaxDiff", "mode", "P1", "P2")] + [
        ("roi1", C.c_int * 4), ("roi2", C.c_int * 4)]


class RtdmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rtdm error {code}: {msg}")
        self.code = code


_lib = None

# name -> (restype, argtypes).  Must list every symbol include/rtdm_b200.h declares.
_vp, _sz, _i = C.c_void_p, C.c_size_t, C.c_int
SIGNATURES = {
    "rtdm_abi_version": (_i, []),
    "rtdm_device_count": (_i, []),
    "rtdm_last_error": (C.c_char_p, []),
    "rtdm_params_default_bm": (None, [C.POINTER(RtdmParams)]),
    "rtdm_params_default_sgbm": (None, [C.POINTER(RtdmParams)]),
    "rtdm_bm_create": (_i, [C.POINTER(_vp), C.POINTER(RtdmParams), _i, _i, _i, _i]),
    "rtdm_bm_destroy": (None, [_vp]),
    "rtdm_bm_set_roi1": (_i, [_vp, _i, _i, _i, _i]),
    "rtdm_bm_set_roi2": (_i, [_vp, _i, _i, _i, _i]),
    "rtdm_bm_compute": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _i, _vp, _sz]),
    "rtdm_bm_compute_batch": (_i, [_vp, _i, _vp, _sz, _sz, _vp, _sz, _sz, _i, _i, _vp, _sz, _sz]),
    "rtdm_bm_submit_batch": (_i, [_vp, _i, _vp, _sz, _sz, _vp, _sz, _sz, _i, _i, _vp, _sz, _sz]),
    "rtdm_bm_wait": (_i, [_vp]),
    "rtdm_bm_wait_oldest": (_i, [_vp]),
    "rtdm_bm_speckle_device": (_i, [_vp, _i, _vp, _sz, _sz, _i, _i, _vp]),
    "rtdm_bm_compute_device": (_i, [_vp, _i, _vp, _sz, _sz, _vp, _sz, _sz, _i, _i, _vp, _sz, _sz, _vp]),
    "rtdm_bm_last_launches": (_i, [_vp]),
    "rtdm_bm_debug_fetch": (_i, [_vp, _i, _vp, _sz]),
    "rtdm_bm_last_kernel": (_i, [_vp]),
    "rtdm_bm_rowband_create": (_i, [C.POINTER(_vp), C.POINTER(RtdmParams), _i, _i, _i, C.POINTER(_i)]),
    "rtdm_bm_rowband_destroy": (None, [_vp]),
    "rtdm_bm_rowband_set_roi1": (_i, [_vp, _i, _i, _i, _i]),
    "rtdm_bm_rowband_set_roi2": (_i, [_vp, _i, _i, _i, _i]),
    "rtdm_bm_rowband_compute": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _i, _vp, _sz]),
    "rtdm_bm_rowband_compute_device": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _i, _vp, _sz]),
    "rtdm_bm_rowband_last_launches": (_i, [_vp]),
    "rtdm_bm_set_profiling": (_i, [_vp, _i]),
    "rtdm_bm_stage_times": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(_i)]),
    "rtdm_sgbm_create": (_i, [C.POINTER(_vp), C.POINTER(RtdmParams), _i, _i, _i, _i]),
    "rtdm_sgbm_destroy": (None, [_vp]),
    "rtdm_sgbm_compute": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _i, _vp, _sz]),
    "rtdm_sgbm_compute_batch": (_i, [_vp, _i, _vp, _sz, _sz, _vp, _sz, _sz, _i, _i, _vp, _sz, _sz]),
    "rtdm_sgbm_submit_batch": (_i, [_vp, _i, _vp, _sz, _sz, _vp, _sz, _sz, _i, _i, _vp, _sz, _sz]),
    "rtdm_sgbm_wait": (_i, [_vp]),
    "rtdm_sgbm_wait_oldest": (_i, [_vp]),
    "rtdm_sgbm_compute_device": (_i, [_vp, _i, _vp, _sz, _sz, _vp, _sz, _sz, _i, _i, _vp, _sz, _sz, _vp]),
    "rtdm_sgbm_last_launches": (_i, [_vp]),
    "rtdm_sgbm_batch_quantum": (_i, [_vp, _i, _i]),
    "rtdm_sgbm_set_profiling": (_i, [_vp, _i]),
    "rtdm_sgbm_stage_times": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(_i)]),
    "rtdm_morph_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i]),
    "rtdm_morph_destroy": (None, [_vp]),
    "rtdm_morph_in_buffer": (_vp, [_vp]),
    "rtdm_morph_out_buffer": (_vp, [_vp]),
    "rtdm_morph_run": (_i, [_vp, _vp, _vp]),
    "rtdm_morph_run_batch": (_i, [_vp, _i, _vp, _vp]),
    "rtdm_morph_run_batch_async": (_i, [_vp, _i, _vp, _vp]),
    "rtdm_morph_sync": (_i, [_vp]),
    "rtdm_morph_run_device": (_i, [_vp, _i, _vp, _vp, _vp]),
    "rtdm_morph_last_launches": (_i, [_vp]),
    "rtdm_filter_speckles": (_i, [_vp, _sz, _i, _i, _i, _i, _i, _i]),
    "rtdm_median3_s16": (_i, [_vp, _sz, _vp, _sz, _i, _i, _i]),
    "rtdm_morph_op": (_i, [_vp, _sz, _vp, _sz, _i, _i, _i, _i, _i, _i]),
    "rtdm_validate_disparity": (_i, [_vp, _sz, _vp, _sz, _i, _i, _i, _i, _i, _i]),
    "rtdm_rectify_create": (_i, [C.POINTER(_vp), _i, _i, _vp, _sz, _vp, _sz, _i, _i, _i, _i, _i, _i]),
    "rtdm_rectify_destroy": (None, [_vp]),
    "rtdm_rectify_run": (_i, [_vp, _i, _vp, _sz, _sz, _vp, _sz, _sz]),
    "rtdm_rectify_run_device": (_i, [_vp, _i, _vp, _sz, _sz, _vp, _sz, _sz, _vp]),
    "rtdm_rectify_last_launches": (_i, [_vp]),
    "rtdm_colormask_create": (_i, [C.POINTER(_vp), _i, _i, _vp, _sz, _vp, _sz, _i, _i, _i, _i, _i, _i]),
    "rtdm_colormask_destroy": (None, [_vp]),
    "rtdm_colormask_run": (_i, [_vp, _i, _vp, _sz, _sz, _vp, _vp, _vp, _sz, _sz, _vp, _sz, _sz]),
    "rtdm_colormask_run_device": (_i, [_vp, _i, _vp, _sz, _sz, _vp, _vp, _vp, _sz, _sz, _vp, _sz, _sz, _vp]),
    "rtdm_colormask_last_launches": (_i, [_vp]),
    "rtdm_regions_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i]),
    "rtdm_regions_destroy": (None, [_vp]),
    "rtdm_regions_run": (_i, [_vp, _vp, _sz, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "rtdm_regions_run_device": (_i, [_vp, _vp, _sz, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "rtdm_regions_last_launches": (_i, [_vp]),
    "rtdm_depth_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i]),
    "rtdm_depth_destroy": (None, [_vp]),
    "rtdm_depth_run": (_i, [_vp, _vp, _sz, _i, _i, _vp, _vp, _sz, _i, _vp, _vp, _vp, _vp, _sz]),
    "rtdm_depth_run_device": (_i, [_vp, _vp, _sz, _i, _i, _vp, _vp, _sz, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rtdm_depth_last_launches": (_i, [_vp]),
    "rtdm_measure_int_peak": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}


def lib():
    """Loads librtdm_b200.so (built in-tree by `make -C rt-depth-map_b200`).  Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RtdmError(-ENODEV, f"{LIB_PATH} is missing: build it with `make -C rt-depth-map_b200` "
                                     "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def _check(rc: int):
    if rc != 0:
        raise RtdmError(rc, lib().rtdm_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    return lib().rtdm_device_count()


def _rect(r):
    """Accepts a cv::Rect-like (x, y, w, h) tuple or None (= empty Rect)."""
    return (0, 0, 0, 0) if r is None else tuple(int(v) for v in r)


def _u8_2d(a, name):
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise RtdmError(-EINVAL, f"{name}: CV_8UC1 (2-D uint8) image required")
    if a.strides[1] != 1:
        a = np.ascontiguousarray(a)
    return a


class BlockMatcher:
    """Abstract matcher, same surface as the reference's BlockMatcher (stereo-matcher.h:13-19)."""

    def compute(self, left, right, out=None):
        raise NotImplementedError

    def setROI1(self, roi1):
        raise NotImplementedError

    def setROI2(self, roi2):
        raise NotImplementedError


class _MatcherBase(BlockMatcher):
    _prefix = ""

    def __init__(self, params: RtdmParams, max_width: int, max_height: int, max_batch: int, device: int):
        self._l = lib()
        self._h = _vp()
        self.params = params
        self.max_width, self.max_height, self.max_batch, self.device = max_width, max_height, max_batch, device
        _check(getattr(self._l, self._prefix + "_create")(C.byref(self._h), C.byref(params), max_width,
                                                          max_height, max_batch, device))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and _vp is not None:
            try:
                getattr(self._l, self._prefix + "_destroy")(h)
            except Exception:       # interpreter shutdown
                pass
            self._h.value = None

    close = __del__

    def compute(self, left, right, out=None):
        """compute(left, right[, out]) -> CV_16SC1 disparity (x16).  Returns the disparity array
        (the reference returns 0 and fills `out`; bm-sw.cpp:33-38)."""
        left, right = _u8_2d(left, "left"), _u8_2d(right, "right")
        if left.shape != right.shape:
            raise RtdmError(-EINVAL, "left and right must have the same size")
        H, W = left.shape
        if out is None or out.shape != (H, W) or out.dtype != np.int16 or out.strides[1] != 2:
            out = np.empty((H, W), np.int16)
        _check(getattr(self._l, self._prefix + "_compute")(
            self._h, left.ctypes.data, left.strides[0], right.ctypes.data, right.strides[0], W, H,
            out.ctypes.data, out.strides[0]))
        return out

    def compute_batch(self, left, right, out=None):
        """left/right: (N, H, W) uint8 host arrays -> (N, H, W) int16."""
        left, right = np.ascontiguousarray(left, np.uint8), np.ascontiguousarray(right, np.uint8)
        N, H, W = left.shape
        if out is None:
            out = np.empty((N, H, W), np.int16)
        _check(getattr(self._l, self._prefix + "_compute_batch")(
            self._h, N, left.ctypes.data, W, W * H, right.ctypes.data, W, W * H, W, H,
            out.ctypes.data, W * 2, W * H * 2))
        return out

    def compute_device(self, n, left_ptr, lstep, lframe, right_ptr, rstep, rframe, width, height,
                       disp_ptr, dstep, dframe, stream=0):
        """Raw device-pointer call (asynchronous on `stream`)."""
        _check(getattr(self._l, self._prefix + "_compute_device")(
            self._h, n, left_ptr, lstep, lframe, right_ptr, rstep, rframe, width, height,
            disp_ptr, dstep, dframe, stream))

    def last_launches(self) -> int:
        return getattr(self._l, self._prefix + "_last_launches")(self._h)


class CUDAMatcherKonolige(_MatcherBase):
    """B200 peer of SWMatcherKonolige; same constructor arguments (bm-sw.h:28-30).  As in the
    reference (bm-sw.cpp:12-14 vs :16-25) `roi1`, `roi2` and `maxDisparity` are accepted and unused."""
    _prefix = "rtdm_bm"

    def __init__(self, roi1, roi2, preFilterCap, blockSize, minDisparity, textureThreshold,
                 numOfDisparities, maxDisparity, uniquenessRatio, speckleWindowSize, speckleRange,
                 disp12MaxDiff, *, preFilterType=PREFILTER_XSOBEL, preFilterSize=9,
                 max_width=1280, max_height=720, max_batch=1, device=0):
        p = RtdmParams()
        lib().rtdm_params_default_bm(C.byref(p))
        p.preFilterType, p.preFilterSize, p.preFilterCap = preFilterType, preFilterSize, preFilterCap
        p.blockSize, p.minDisparity, p.numDisparities = blockSize, minDisparity, numOfDisparities
        p.textureThreshold, p.uniquenessRatio = textureThreshold, uniquenessRatio
        p.speckleWindowSize, p.speckleRange, p.disp12MaxDiff = speckleWindowSize, speckleRange, disp12MaxDiff
        super().__init__(p, max_width, max_height, max_batch, device)

    def setROI1(self, roi1):
        _check(self._l.rtdm_bm_set_roi1(self._h, *_rect(roi1)))

    def setROI2(self, roi2):
        _check(self._l.rtdm_bm_set_roi2(self._h, *_rect(roi2)))

    STAGES = ("prefilter", "sad_wta", "validate_mask", "speckle")

    def set_profiling(self, on: bool):
        _check(self._l.rtdm_bm_set_profiling(self._h, int(on)))

    def stage_times(self):
        """-> ({stage: total ms}, number of profiled calls) since the last query."""
        ms = (C.c_double * 4)()
        calls = _i()
        _check(self._l.rtdm_bm_stage_times(self._h, ms, C.byref(calls)))
        return dict(zip(self.STAGES, list(ms))), calls.value

    def speckle_device(self, n, disp_ptr, dstep, dframe, W, H, stream=0):
        """filterSpeckles with this matcher's parameters on n device frames in place (rtdm_bm_speckle_device)."""
        _check(self._l.rtdm_bm_speckle_device(self._h, n, disp_ptr, dstep, dframe, W, H, stream))

    def submit_batch(self, left, right, out):
        """Streaming variant of compute_batch: (N, H, W) uint8 pinned arrays in, (N, H, W) int16 pinned array out;
        returns immediately.  At most two submissions are in flight; call wait() before reading `out`."""
        N, H, W = left.shape
        if not (left.flags.c_contiguous and right.flags.c_contiguous and out.flags.c_contiguous):
            raise RtdmError(-EINVAL, "submit_batch: contiguous arrays required")
        _check(self._l.rtdm_bm_submit_batch(self._h, N, left.ctypes.data, W, W * H, right.ctypes.data, W, W * H, W, H,
                                            out.ctypes.data, W * 2, W * H * 2))

    def wait(self):
        """Drains every submission."""
        _check(self._l.rtdm_bm_wait(self._h))

    def wait_oldest(self):
        """Waits for the older of the (at most two) submissions in flight; its output array is then complete."""
        _check(self._l.rtdm_bm_wait_oldest(self._h))

    def last_kernel(self) -> int:
        return self._l.rtdm_bm_last_kernel(self._h)

    def debug_fetch(self, what: int, width: int, height: int):
        dt = np.uint8 if what in (0, 1) else np.int16
        a = np.empty((height, width), dt)
        _check(self._l.rtdm_bm_debug_fetch(self._h, what, a.ctypes.data, a.nbytes))
        return a


class CUDARowBandMatcherKonolige(BlockMatcher):
    """One large frame split into row bands over several GPUs of this process (rtdm_bm_rowband_*: SURVEY.md 8(e),
    optional part).  Constructor arguments as CUDAMatcherKonolige plus `devices` (a device may be listed more than
    once: its bands then run one after the other).  The bands reach devices[0] as peer copies; the speckle filter runs
    there on the stitched frame.  Bit-exact against CUDAMatcherKonolige."""

    def __init__(self, roi1, roi2, preFilterCap, blockSize, minDisparity, textureThreshold,
                 numOfDisparities, maxDisparity, uniquenessRatio, speckleWindowSize, speckleRange,
                 disp12MaxDiff, *, devices, preFilterType=PREFILTER_XSOBEL, preFilterSize=9,
                 max_width=1280, max_height=720):
        self._l = lib()
        self._h = _vp()
        p = RtdmParams()
        self._l.rtdm_params_default_bm(C.byref(p))
        p.preFilterType, p.preFilterSize, p.preFilterCap = preFilterType, preFilterSize, preFilterCap
        p.blockSize, p.minDisparity, p.numDisparities = blockSize, minDisparity, numOfDisparities
        p.textureThreshold, p.uniquenessRatio = textureThreshold, uniquenessRatio
        p.speckleWindowSize, p.speckleRange, p.disp12MaxDiff = speckleWindowSize, speckleRange, disp12MaxDiff
        self.params, self.devices = p, list(devices)
        arr = (_i * len(self.devices))(*self.devices)
        _check(self._l.rtdm_bm_rowband_create(C.byref(self._h), C.byref(p), max_width, max_height, len(self.devices), arr))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and _vp is not None:
            try:
                self._l.rtdm_bm_rowband_destroy(h)
            except Exception:       # interpreter shutdown
                pass
            self._h.value = None

    close = __del__

    def setROI1(self, roi1):
        _check(self._l.rtdm_bm_rowband_set_roi1(self._h, *_rect(roi1)))

    def setROI2(self, roi2):
        _check(self._l.rtdm_bm_rowband_set_roi2(self._h, *_rect(roi2)))

    def compute(self, left, right, out=None):
        left, right = _u8_2d(left, "left"), _u8_2d(right, "right")
        if left.shape != right.shape:
            raise RtdmError(-EINVAL, "left and right must have the same size")
        H, W = left.shape
        if out is None or out.shape != (H, W) or out.dtype != np.int16 or out.strides[1] != 2:
            out = np.empty((H, W), np.int16)
        _check(self._l.rtdm_bm_rowband_compute(self._h, left.ctypes.data, left.strides[0], right.ctypes.data,
                                               right.strides[0], W, H, out.ctypes.data, out.strides[0]))
        return out

    def compute_device(self, left_ptr, lstep, right_ptr, rstep, width, height, disp_ptr, dstep):
        """Inputs and output resident on devices[0]; blocking."""
        _check(self._l.rtdm_bm_rowband_compute_device(self._h, left_ptr, lstep, right_ptr, rstep, width, height, disp_ptr, dstep))

    def last_launches(self) -> int:
        return self._l.rtdm_bm_rowband_last_launches(self._h)


class CUDASemiGlobalMatcher(_MatcherBase):
    """B200 peer of SWSemiGlobalMatcher; same constructor arguments (sgbm-sw.h:28-29).  P1/P2 are the
    reference's hard-coded 8*3*5*5 / 32*3*5*5 (sgbm-sw.cpp:17-18); ROI setters are no-ops
    (sgbm-sw.h:32-33).  `mode` defaults to MODE_SGBM like the reference; MODE_HH selects 8 paths; the
    keyword-only P1 / P2 override the two penalties (cv::StereoSGBM::setP1 / setP2)."""
    _prefix = "rtdm_sgbm"

    def __init__(self, blockSize, minDisparity, numOfDisparities, uniquenessRatio, speckleWindowSize,
                 speckleRange, disp12MaxDiff, *, mode=MODE_SGBM, P1=None, P2=None, max_width=1280, max_height=720,
                 max_batch=1, device=0):
        p = RtdmParams()
        lib().rtdm_params_default_sgbm(C.byref(p))
        if P1 is not None:
            p.P1 = P1
        if P2 is not None:
            p.P2 = P2
        p.blockSize, p.minDisparity, p.numDisparities = blockSize, minDisparity, numOfDisparities
        p.uniquenessRatio, p.speckleWindowSize, p.speckleRange = uniquenessRatio, speckleWindowSize, speckleRange
        p.disp12MaxDiff, p.mode = disp12MaxDiff, mode
        super().__init__(p, max_width, max_height, max_batch, device)

    def setROI1(self, roi1):
        pass

    def setROI2(self, roi2):
        pass

    STAGES = ("matching", "median_speckle")

    def submit_batch(self, left, right, out):
        """Streaming variant of compute_batch (see CUDAMatcherKonolige.submit_batch): returns immediately, at most two
        submissions in flight; wait_oldest() / wait() before reading `out`."""
        N, H, W = left.shape
        if not (left.flags.c_contiguous and right.flags.c_contiguous and out.flags.c_contiguous):
            raise RtdmError(-EINVAL, "submit_batch: contiguous arrays required")
        _check(self._l.rtdm_sgbm_submit_batch(self._h, N, left.ctypes.data, W, W * H, right.ctypes.data, W, W * H, W, H,
                                              out.ctypes.data, W * 2, W * H * 2))

    def wait(self):
        _check(self._l.rtdm_sgbm_wait(self._h))

    def wait_oldest(self):
        _check(self._l.rtdm_sgbm_wait_oldest(self._h))

    def batch_quantum(self, width: int, height: int) -> int:
        """Frames the aggregation passes work on at once for this size: batches that are multiples of it run fullest."""
        q = self._l.rtdm_sgbm_batch_quantum(self._h, int(width), int(height))
        _check(min(q, 0))
        return q

    def set_profiling(self, on: bool):
        _check(self._l.rtdm_sgbm_set_profiling(self._h, int(on)))

    def stage_times(self):
        ms = (C.c_double * 2)()
        calls = _i()
        _check(self._l.rtdm_sgbm_stage_times(self._h, ms, C.byref(calls)))
        return dict(zip(self.STAGES, list(ms))), calls.value


class VideoFilterDevice:
    """Same getters as the reference's VideoFilterDevice (filter/filter.cpp:10-53)."""
    img_width = img_height = img_bpp = 0

    def getFrameSize(self):
        return self.img_width * self.img_height * (self.img_bpp >> 3)

    def getBpp(self):
        return self.img_bpp

    def getWidth(self):
        return self.img_width

    def getHeight(self):
        return self.img_height


class CUDAMorphologicalFilter(VideoFilterDevice):
    """B200 peer of SWMorphologicalFilter(w, h, bpp) (mf-sw.cpp:10-28).  The in/out buffers are
    pinned host memory owned by the device handle, exposed as numpy views."""

    def __init__(self, w, h, bpp, *, max_batch=1, device=0):
        self._l = lib()
        self._h = _vp()
        self.img_width, self.img_height, self.img_bpp = w, h, bpp
        self.max_batch = max_batch
        _check(self._l.rtdm_morph_create(C.byref(self._h), w, h, bpp, max_batch, device))
        n = w * h
        self._in = np.ctypeslib.as_array((C.c_uint8 * n).from_address(self._l.rtdm_morph_in_buffer(self._h))).reshape(h, w)
        self._out = np.ctypeslib.as_array((C.c_uint8 * n).from_address(self._l.rtdm_morph_out_buffer(self._h))).reshape(h, w)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and _vp is not None:
            self._in = self._out = None
            try:
                self._l.rtdm_morph_destroy(h)
            except Exception:       # interpreter shutdown
                pass
            self._h.value = None

    close = __del__

    def getVideoInBuffer(self):
        return self._in

    def getVideoOutBuffer(self):
        return self._out

    def run(self, inp=None, out=None):
        """run(in, out): open then close with the 10x10 ellipse.  With no arguments it filters the
        device-owned in buffer into the out buffer (how Estimator uses it, estimator.cpp:45,141-142).
        Returns 0."""
        inp = self._in if inp is None else np.ascontiguousarray(inp, np.uint8)
        out = self._out if out is None else out
        if inp.shape != (self.img_height, self.img_width) or out.shape != inp.shape or not out.flags.c_contiguous:
            raise RtdmError(-EINVAL, "filter: frame size mismatch")
        _check(self._l.rtdm_morph_run(self._h, inp.ctypes.data, out.ctypes.data))
        return 0

    def run_batch(self, inp, out=None):
        """(N, H, W) uint8 host frames -> (N, H, W) filtered frames."""
        inp = np.ascontiguousarray(inp, np.uint8)
        if out is None:
            out = np.empty_like(inp)
        _check(self._l.rtdm_morph_run_batch(self._h, inp.shape[0], inp.ctypes.data, out.ctypes.data))
        return out

    def run_batch_async(self, inp, out):
        """Enqueue-only variant of run_batch (pinned host arrays); call sync() before reading `out`."""
        if not (inp.flags.c_contiguous and out.flags.c_contiguous and inp.dtype == np.uint8 and out.dtype == np.uint8):
            raise RtdmError(-EINVAL, "filter: contiguous uint8 arrays required")
        _check(self._l.rtdm_morph_run_batch_async(self._h, inp.shape[0], inp.ctypes.data, out.ctypes.data))

    def sync(self):
        _check(self._l.rtdm_morph_sync(self._h))

    def run_device(self, n, in_ptr, out_ptr, stream=0):
        _check(self._l.rtdm_morph_run_device(self._h, n, in_ptr, out_ptr, stream))

    def last_launches(self) -> int:
        return self._l.rtdm_morph_last_launches(self._h)


# ---- stand-alone stages ---------------------------------------------------------------------------
class CUDARectifier:
    """The step before the matcher (estimator.cpp:29-36), fused on the GPU for one camera:
    cvtColor(RGB2GRAY) -> remap(INTER_LINEAR, the CV_16SC2 / CV_16UC1 maps of initUndistortRectifyMap) -> crop to
    `roi` = (x, y, w, h).  `run` takes (H, W, 3) or (N, H, W, 3) uint8 RGB host arrays."""

    def __init__(self, map1, map2, roi, *, max_batch=1, device=0):
        self._l = lib()
        self._h = _vp()
        self._vp = _vp
        m1 = np.ascontiguousarray(map1, np.int16); m2 = np.ascontiguousarray(map2, np.uint16)
        H, W = m2.shape
        if m1.shape != (H, W, 2):
            raise RtdmError(-EINVAL, "CUDARectifier: map1 must be (H, W, 2) int16 and map2 (H, W) uint16")
        self.W, self.H, self.roi, self.max_batch = W, H, tuple(int(v) for v in roi), max_batch
        _check(self._l.rtdm_rectify_create(C.byref(self._h), W, H, m1.ctypes.data, W * 4, m2.ctypes.data, W * 2,
                                           *self.roi, max_batch, device))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._vp is not None:
                self._l.rtdm_rectify_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def run(self, rgb):
        a = np.ascontiguousarray(rgb, np.uint8)
        single = a.ndim == 3
        if single:
            a = a[None]
        N, H, W, _ = a.shape
        rw, rh = self.roi[2], self.roi[3]
        out = np.empty((N, rh, rw), np.uint8)
        _check(self._l.rtdm_rectify_run(self._h, N, a.ctypes.data, W * 3, W * H * 3, out.ctypes.data, rw, rw * rh))
        return out[0] if single else out

    def run_device(self, n, rgb_ptr, step, frame, out_ptr, ostep, oframe, stream=0):
        _check(self._l.rtdm_rectify_run_device(self._h, n, rgb_ptr, step, frame, out_ptr, ostep, oframe, stream))

    def last_launches(self) -> int:
        return self._l.rtdm_rectify_last_launches(self._h)


class CUDADepthEpilogue:
    """What Estimator::run does with the matcher's output (estimator.cpp:75-77), fused on the GPU:
    `left_disp /= 16.`, `reprojectImageTo3D(left_disp, xyz, Q, true, CV_32F)` and `calc_depth` (masked mean Z per
    bounding rectangle, estimator.cpp:206-263).  `run` takes host arrays, `run_device` device pointers."""

    def __init__(self, max_width=1280, max_height=720, max_regions=64, device=0):
        self._l = lib()
        self._h = _vp()
        self._vp = _vp
        _check(self._l.rtdm_depth_create(C.byref(self._h), max_width, max_height, max_regions, device))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._vp is not None:
                self._l.rtdm_depth_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def run(self, disp, Q, mask, rects, want_xyz=False):
        """disp: (H, W) int16 x16 disparity as the matcher returns it; Q: 4x4; mask: (H, W) uint8 or None;
        rects: (n, 4) x, y, w, h.  Returns (mean_z [n] float64, count [n] int32[, xyz (H, W, 3) float32])."""
        disp = np.ascontiguousarray(disp, np.int16)
        H, W = disp.shape
        Qa = np.ascontiguousarray(Q, np.float64).reshape(16)
        r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4)
        n = r.shape[0]
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        mean, cnt = np.zeros(max(n, 1), np.float64), np.zeros(max(n, 1), np.int32)
        xyz = np.empty((H, W, 3), np.float32) if want_xyz else None
        _check(self._l.rtdm_depth_run(self._h, disp.ctypes.data, W * 2, W, H, Qa.ctypes.data,
                                      None if m is None else m.ctypes.data, W, n, r.ctypes.data,
                                      mean.ctypes.data, cnt.ctypes.data, None if xyz is None else xyz.ctypes.data, W * 12))
        return (mean[:n], cnt[:n], xyz) if want_xyz else (mean[:n], cnt[:n])

    def run_device(self, disp_ptr, dstep, W, H, Q, mask_ptr, mstep, rects, stream=0, xyz_ptr=None, xstep=0):
        Qa = np.ascontiguousarray(Q, np.float64).reshape(16)
        r = np.ascontiguousarray(rects, np.int32).reshape(-1, 4)
        n = r.shape[0]
        mean, cnt = np.zeros(max(n, 1), np.float64), np.zeros(max(n, 1), np.int32)
        _check(self._l.rtdm_depth_run_device(self._h, disp_ptr, dstep, W, H, Qa.ctypes.data, mask_ptr, mstep, n, r.ctypes.data,
                                             mean.ctypes.data, cnt.ctypes.data, xyz_ptr, xstep, stream))
        return mean[:n], cnt[:n]

    def last_launches(self) -> int:
        return self._l.rtdm_depth_last_launches(self._h)


def distance_cm(mean_z, calibration_unit):
    """The label Estimator::calc_depth prints (estimator.cpp:252-254)."""
    return f"{mean_z * calibration_unit / 10.0:.0f} cm"


def filter_speckles(img, newVal, maxSpeckleSize, maxDiff, device=0):
    a = np.ascontiguousarray(img, np.int16).copy()
    H, W = a.shape
    _check(lib().rtdm_filter_speckles(a.ctypes.data, W * 2, W, H, int(newVal), int(maxSpeckleSize), int(maxDiff), device))
    return a


def median3_s16(img, device=0):
    a = np.ascontiguousarray(img, np.int16)
    H, W = a.shape
    out = np.empty_like(a)
    _check(lib().rtdm_median3_s16(a.ctypes.data, W * 2, out.ctypes.data, W * 2, W, H, device))
    return out


def morph_op(img, op, kw=10, kh=10, device=0):
    a = np.ascontiguousarray(img, np.uint8)
    H, W = a.shape
    out = np.empty_like(a)
    _check(lib().rtdm_morph_op(a.ctypes.data, W, out.ctypes.data, W, W, H, kw, kh, int(op), device))
    return out


def validate_disparity(disp, cost, minD, nd, d12, device=0):
    d = np.ascontiguousarray(disp, np.int16).copy()
    c = np.ascontiguousarray(cost, np.int16)
    H, W = d.shape
    _check(lib().rtdm_validate_disparity(d.ctypes.data, W * 2, c.ctypes.data, W * 2, W, H, minD, nd, d12, device))
    return d


def measure_int_peak(device=0):
    a, b, c, m = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    _check(lib().rtdm_measure_int_peak(device, C.byref(a), C.byref(b), C.byref(c), C.byref(m)))
    return {"iadd3_tiops": a.value, "vimnmx_lop3_tiops": b.value, "vabsdiff4_iadd_tiops": c.value,
            "sm_mhz_if_64_lanes": m.value}


class CUDAColorMask:
    """The step before the morphological filter (estimator.cpp:38-43), fused on the GPU:
    remap(RGB frame, INTER_LINEAR) -> crop to `roi` -> RGB2BGR -> BGR2HSV -> inRange(low, high) -> filter_in.
    `run` takes (H, W, 3) or (N, H, W, 3) uint8 RGB host arrays; returns the 0/255 mask(s), and the rectified BGR
    crop(s) the reference displays when `want_bgr`."""

    def __init__(self, map1, map2, roi, *, max_batch=1, device=0):
        self._l = lib()
        self._h = _vp()
        self._vp = _vp
        m1 = np.ascontiguousarray(map1, np.int16); m2 = np.ascontiguousarray(map2, np.uint16)
        H, W = m2.shape
        if m1.shape != (H, W, 2):
            raise RtdmError(-EINVAL, "CUDAColorMask: map1 must be (H, W, 2) int16 and map2 (H, W) uint16")
        self.W, self.H, self.roi, self.max_batch = W, H, tuple(int(v) for v in roi), max_batch
        _check(self._l.rtdm_colormask_create(C.byref(self._h), W, H, m1.ctypes.data, W * 4, m2.ctypes.data, W * 2,
                                             *self.roi, max_batch, device))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._vp is not None:
                self._l.rtdm_colormask_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def run(self, rgb, low, high, want_bgr=False):
        a = np.ascontiguousarray(rgb, np.uint8)
        single = a.ndim == 3
        if single:
            a = a[None]
        N, H, W, _ = a.shape
        rw, rh = self.roi[2], self.roi[3]
        lo = np.ascontiguousarray(low, np.int32); hi = np.ascontiguousarray(high, np.int32)
        if lo.shape != (3,) or hi.shape != (3,):
            raise RtdmError(-EINVAL, "CUDAColorMask: low and high are (H, S, V) triples")
        mask = np.empty((N, rh, rw), np.uint8)
        bgr = np.empty((N, rh, rw, 3), np.uint8) if want_bgr else None
        _check(self._l.rtdm_colormask_run(self._h, N, a.ctypes.data, W * 3, W * H * 3, lo.ctypes.data, hi.ctypes.data,
                                          mask.ctypes.data, rw, rw * rh, bgr.ctypes.data if want_bgr else None, rw * 3, rw * rh * 3))
        if want_bgr:
            return (mask[0], bgr[0]) if single else (mask, bgr)
        return mask[0] if single else mask

    def run_device(self, n, rgb_ptr, step, frame, low, high, mask_ptr, mstep, mframe, bgr_ptr=None, bstep=0, bframe=0, stream=0):
        lo = np.ascontiguousarray(low, np.int32); hi = np.ascontiguousarray(high, np.int32)
        _check(self._l.rtdm_colormask_run_device(self._h, n, rgb_ptr, step, frame, lo.ctypes.data, hi.ctypes.data,
                                                 mask_ptr, mstep, mframe, bgr_ptr, bstep, bframe, stream))

    def last_launches(self) -> int:
        return self._l.rtdm_colormask_last_launches(self._h)


class CUDAObjectRegions:
    """The step after the morphological filter (estimator.cpp:46-53, :164-204) on the GPU: bounding boxes of the
    top-level contours of the mask (findContours RETR_EXTERNAL + boundingRect, in OpenCV's order) with area >=
    `min_obj_size`, and the rectangle spanning them (what bm->setROI1 receives).
    `run` returns (rects (n, 4) int32 as x, y, w, h; roi (x, y, w, h); ncontours)."""

    def __init__(self, max_width, max_height, max_regions=256, *, device=0):
        self._l = lib()
        self._h = _vp()
        self._vp = _vp
        self.max_regions = max_regions
        _check(self._l.rtdm_regions_create(C.byref(self._h), max_width, max_height, max_regions, device))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._vp is not None:
                self._l.rtdm_regions_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _finish(self, rects, cnt, nc, roi):
        return rects[:cnt.value].copy(), tuple(int(v) for v in roi), int(nc.value)

    def run(self, mask, min_obj_size):
        m = np.asarray(mask)
        if m.dtype != np.uint8 or m.ndim != 2:
            raise RtdmError(-EINVAL, "CUDAObjectRegions: the mask must be a 2-D uint8 array")
        if m.strides[1] != 1:
            m = np.ascontiguousarray(m)
        H, W = m.shape
        rects = np.zeros((self.max_regions, 4), np.int32); roi = np.zeros(4, np.int32)
        cnt, nc = C.c_int(0), C.c_int(0)
        _check(self._l.rtdm_regions_run(self._h, m.ctypes.data, m.strides[0], W, H, int(min_obj_size), rects.ctypes.data,
                                        C.addressof(cnt), C.addressof(nc), roi.ctypes.data))
        return self._finish(rects, cnt, nc, roi)

    def run_device(self, mask_ptr, mstep, width, height, min_obj_size, stream=0):
        rects = np.zeros((self.max_regions, 4), np.int32); roi = np.zeros(4, np.int32)
        cnt, nc = C.c_int(0), C.c_int(0)
        _check(self._l.rtdm_regions_run_device(self._h, mask_ptr, mstep, width, height, int(min_obj_size), rects.ctypes.data,
                                               C.addressof(cnt), C.addressof(nc), roi.ctypes.data, stream))
        return self._finish(rects, cnt, nc, roi)

    def last_launches(self) -> int:
        return self._l.rtdm_regions_last_launches(self._h)
