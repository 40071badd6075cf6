/*
 * rtdm_b200.h -- C ABI of the B200-native stereo disparity engine (librtdm_b200.so).
 *
 * This is the drop-in boundary for rt-depth-map's matcher / filter plugins.  Every entry point
 * names the reference interface it replaces (file:line under the reference checkout).  The C++
 * plugin peers (rt-depth-map_b200/host/ headers: CUDAMatcherKonolige, CUDASemiGlobalMatcher,
 * CUDAMorphologicalFilter) and the Python ctypes mirror (rt-depth-map_b200/rtdm_b200) both call
 * exactly these functions.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Error convention: 0 = ok, negative errno-style code otherwise (the flavour of the reference's
 * include/errors.h:12-46): -RTDM_EINVAL for parameters cv::StereoBM/SGBM would assert on or that
 * are outside the bit-exact domain, -RTDM_ENODEV when no CUDA device / kernel image is usable,
 * -RTDM_ENOMEM on allocation failure, -RTDM_EIO on any other CUDA runtime failure.
 * There is no CPU fallback: without a GPU every compute call fails with -RTDM_ENODEV.
 */
#ifndef RTDM_B200_H_
#define RTDM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTDM_ABI_VERSION 1

#define RTDM_EIO     5
#define RTDM_ENOMEM  12
#define RTDM_ENODEV  19
#define RTDM_EINVAL  22
#define RTDM_ENOSYS  38

/* preFilterType values (cv::StereoBM::PREFILTER_*) */
#define RTDM_PREFILTER_NORMALIZED_RESPONSE 0
#define RTDM_PREFILTER_XSOBEL              1
/* SGBM modes (cv::StereoSGBM::MODE_*) */
#define RTDM_SGBM_MODE_SGBM 0   /* 5 paths; the reference's literal default (sgbm-sw.cpp:15) */
#define RTDM_SGBM_MODE_HH   1   /* 8 paths, two passes */

/* Matcher parameters.  One POD for both matchers; fields a matcher does not use are ignored.
 * Mirrors what the reference's constructors set on the OpenCV objects:
 *   SWMatcherKonolige::SWMatcherKonolige   stereo-matcher/bm-sw.cpp:12-26
 *   SWSemiGlobalMatcher::SWSemiGlobalMatcher stereo-matcher/sgbm-sw.cpp:12-25 */
typedef struct rtdm_params {
    int preFilterType;      /* BM; default XSOBEL                                   */
    int preFilterSize;      /* BM; default 9 (NORMALIZED_RESPONSE only)             */
    int preFilterCap;       /* BM: 1..63 (bit-exact domain with disp12>=0: <=31); SGBM: 0 */
    int blockSize;          /* odd; BM 5..  (bit-exact domain with disp12>=0: <=21) */
    int minDisparity;
    int numDisparities;     /* multiple of 16                                       */
    int textureThreshold;   /* BM                                                   */
    int uniquenessRatio;
    int speckleWindowSize;
    int speckleRange;
    int disp12MaxDiff;
    int mode;               /* SGBM: RTDM_SGBM_MODE_*                               */
    int P1, P2;             /* SGBM                                                 */
    int roi1[4];            /* x, y, w, h; w*h == 0 means "not set" (whole image)   */
    int roi2[4];
} rtdm_params;

typedef struct rtdm_bm rtdm_bm;         /* opaque: Konolige block matcher      */
typedef struct rtdm_sgbm rtdm_sgbm;     /* opaque: semi-global matcher         */
typedef struct rtdm_morph rtdm_morph;   /* opaque: morphological filter device */

/* ---- library ----------------------------------------------------------------------------- */
int rtdm_abi_version(void);
/* number of usable CUDA devices (0 when none; never negative) */
int rtdm_device_count(void);
/* last error message of the calling thread ("" if none) */
const char *rtdm_last_error(void);
/* fills *p with the defaults the reference's main.cpp:134-135 passes to SWMatcherKonolige
 * (cap 31, bs 13, minD 0, tex 10, nd 128, uniq 10, speckle 100/32, disp12 1, XSOBEL, size 9). */
void rtdm_params_default_bm(rtdm_params *p);
/* defaults of SWSemiGlobalMatcher (sgbm-sw.cpp:15-24): bs 5, P1 600, P2 2400, MODE_SGBM. */
void rtdm_params_default_sgbm(rtdm_params *p);

/* ---- SWMatcherKonolige peer -------------------------------------------------------------- */
/* replaces SWMatcherKonolige::SWMatcherKonolige (bm-sw.cpp:12-26).  max_batch frames of at most
 * max_width x max_height can be in flight per call. */
int rtdm_bm_create(rtdm_bm **out, const rtdm_params *p, int max_width, int max_height,
                   int max_batch, int device);
void rtdm_bm_destroy(rtdm_bm *h);
/* replaces SWMatcherKonolige::setROI1 / setROI2 (bm-sw.cpp:40-48) */
int rtdm_bm_set_roi1(rtdm_bm *h, int x, int y, int w, int hgt);
int rtdm_bm_set_roi2(rtdm_bm *h, int x, int y, int w, int hgt);
/* replaces SWMatcherKonolige::compute (bm-sw.cpp:33-38): HOST pointers, CV_8UC1 inputs with
 * arbitrary row steps (bytes), CV_16SC1 output (x16 fixed point, invalid = (minD-1)*16) with row
 * step dstep (bytes).  Synchronous: copies in, runs the CUDA pipeline, copies out. */
int rtdm_bm_compute(rtdm_bm *h, const uint8_t *left, size_t lstep, const uint8_t *right,
                    size_t rstep, int width, int height, int16_t *disp, size_t dstep);
/* batched host variant: n frames, frame k at base + k*frame_stride (bytes). */
int rtdm_bm_compute_batch(rtdm_bm *h, int n, const uint8_t *left, size_t lstep, size_t lframe,
                          const uint8_t *right, size_t rstep, size_t rframe, int width, int height,
                          int16_t *disp, size_t dstep, size_t dframe);
/* streaming variant of rtdm_bm_compute_batch: enqueues the copies and kernels and returns.  Two calls may be in
 * flight (double-buffered staging), so the H2D of batch k+1 and the D2H of batch k-1 overlap the kernels of batch k;
 * a third submit blocks until the oldest call has finished.  left / right / disp should be pinned host memory and
 * must stay valid until rtdm_bm_wait (which drains everything submitted so far) returns. */
int rtdm_bm_submit_batch(rtdm_bm *h, int n, const uint8_t *left, size_t lstep, size_t lframe,
                         const uint8_t *right, size_t rstep, size_t rframe, int width, int height,
                         int16_t *disp, size_t dstep, size_t dframe);
int rtdm_bm_wait(rtdm_bm *h);
/* waits for the OLDER of the submissions still in flight (the only one, if one): its `disp` is then complete */
int rtdm_bm_wait_oldest(rtdm_bm *h);
/* device variant: DEVICE pointers, asynchronous on `cuda_stream` (a cudaStream_t, may be NULL =
 * the handle's own stream).  Inputs/outputs stay resident in HBM. */
int rtdm_bm_compute_device(rtdm_bm *h, int n, const uint8_t *left, size_t lstep, size_t lframe,
                           const uint8_t *right, size_t rstep, size_t rframe, int width, int height,
                           int16_t *disp, size_t dstep, size_t dframe, void *cuda_stream);
/* The last stage of compute alone: cv::filterSpeckles with the handle's speckleWindowSize / speckleRange on n DEVICE
 * frames, in place, asynchronous on `cuda_stream`.  The multi-GPU row-band split (rtdm_b200/rowband.py) runs the
 * bands with the speckle filter off and applies it here to the stitched frame, because components cross bands
 * (SURVEY.md 8(e)).  width * height may not exceed the handle's max_width * max_height. */
int rtdm_bm_speckle_device(rtdm_bm *h, int n, int16_t *disp, size_t dstep, size_t dframe, int width, int height,
                           void *cuda_stream);
/* number of kernel launches issued by the last compute call on this handle */
int rtdm_bm_last_launches(const rtdm_bm *h);
/* which SAD/WTA kernel the last compute call used: 1 = generic (bm_sad.cu), 2 = bm_sad2.cu, 3 = warp-specialised
 * bm_sad3.cu, 4 = TMA-staged warp-specialised bm_sad4.cu (the default wherever it applies: minDisparity 0, blockSize 5..15,
 * numDisparities 32 / 48 / 64 / 96 / 128 / 192 / 256).  The environment variable RTDM_BM_KERNEL = 1 / 2 / 3, read when the
 * handle is created, keeps the older kernel (A/B timing, the tests' kernel matrix). */
int rtdm_bm_last_kernel(const rtdm_bm *h);
/* Per-stage device timing with CUDA events recorded on the launching stream (used by bench.py for
 * the roofline of the dominant kernel).  While enabled every compute call records 5 events.
 * rtdm_bm_stage_times synchronises, sums the elapsed milliseconds of all recorded calls per stage
 * (0 prefilter, 1 SAD+WTA, 2 validate+mask, 3 speckle) into ms_sum[4], and clears the record. */
int rtdm_bm_set_profiling(rtdm_bm *h, int on);
int rtdm_bm_stage_times(rtdm_bm *h, double *ms_sum, int *calls);
/* test hook: copy intermediate planes of frame 0 of the last compute call to host.
 * what: 0 = prefiltered left, 1 = prefiltered right (uint8, width bytes per row),
 *       2 = raw WTA disparity before validate/mask/speckle, 3 = cost (int16, width per row). */
int rtdm_bm_debug_fetch(rtdm_bm *h, int what, void *dst, size_t dst_bytes);

/* ---- one large frame over several GPUs: row bands (optional part) ---------------------------------
 * Same call as SWMatcherKonolige::compute (bm-sw.cpp:33-38) for ONE frame, computed by n_gpus devices of this process:
 * device i computes the output rows [H*i/n, H*(i+1)/n) from its input rows + a halo of blockSize/2 + 1 rows (every stage
 * before the speckle filter is local to a few rows), the int16 bands travel to devices[0] as peer copies
 * (cudaMemcpyPeerAsync: NVLink / NVSwitch when the devices are peers), and devices[0] runs cv::filterSpeckles on the
 * stitched frame, because components cross bands.  Bit-exact against rtdm_bm_compute.  The same device may be listed
 * several times (bands then run one after the other; used by the single-GPU tests).  minDisparity must be <= 0 (the
 * row spill of SURVEY.md App. B.3 crosses bands).  SGBM does not split this way (its vertical paths span the image). */
typedef struct rtdm_bm_rowband rtdm_bm_rowband;
int rtdm_bm_rowband_create(rtdm_bm_rowband **out, const rtdm_params *p, int max_width, int max_height,
                           int n_gpus, const int *devices);
void rtdm_bm_rowband_destroy(rtdm_bm_rowband *h);
/* replace StereoBM::setROI1 / setROI2 for the split matcher (bm-sw.cpp:40-48) */
int rtdm_bm_rowband_set_roi1(rtdm_bm_rowband *h, int x, int y, int w, int hgt);
int rtdm_bm_rowband_set_roi2(rtdm_bm_rowband *h, int x, int y, int w, int hgt);
/* HOST pointers; every device receives only its own input rows + halo; blocking */
int rtdm_bm_rowband_compute(rtdm_bm_rowband *h, const uint8_t *left, size_t lstep, const uint8_t *right, size_t rstep,
                            int width, int height, int16_t *disp, size_t dstep);
/* DEVICE pointers on devices[0] (inputs and output); the other devices fetch their input rows + halo as peer
 * copies.  Blocking (returns when `disp` is complete). */
int rtdm_bm_rowband_compute_device(rtdm_bm_rowband *h, const uint8_t *left, size_t lstep, const uint8_t *right, size_t rstep,
                                   int width, int height, int16_t *disp, size_t dstep);
/* kernel launches of the last call, summed over the devices */
int rtdm_bm_rowband_last_launches(const rtdm_bm_rowband *h);

/* ---- SWSemiGlobalMatcher peer ------------------------------------------------------------ */
/* replaces SWSemiGlobalMatcher::SWSemiGlobalMatcher (sgbm-sw.cpp:12-25) */
int rtdm_sgbm_create(rtdm_sgbm **out, const rtdm_params *p, int max_width, int max_height,
                     int max_batch, int device);
void rtdm_sgbm_destroy(rtdm_sgbm *h);
/* replaces SWSemiGlobalMatcher::compute (sgbm-sw.cpp:32-37); same contracts as the BM calls */
int rtdm_sgbm_compute(rtdm_sgbm *h, const uint8_t *left, size_t lstep, const uint8_t *right,
                      size_t rstep, int width, int height, int16_t *disp, size_t dstep);
int rtdm_sgbm_compute_batch(rtdm_sgbm *h, int n, const uint8_t *left, size_t lstep, size_t lframe,
                            const uint8_t *right, size_t rstep, size_t rframe, int width,
                            int height, int16_t *disp, size_t dstep, size_t dframe);
/* streaming variants, exactly as rtdm_bm_submit_batch / rtdm_bm_wait / rtdm_bm_wait_oldest: two calls may be in flight,
 * the copies of batch k+1 / k-1 run under the kernels of batch k */
int rtdm_sgbm_submit_batch(rtdm_sgbm *h, int n, const uint8_t *left, size_t lstep, size_t lframe,
                           const uint8_t *right, size_t rstep, size_t rframe, int width, int height,
                           int16_t *disp, size_t dstep, size_t dframe);
/* both waits (and the next call on the handle) return -EIO if a whole-height aggregation pass had to give up waiting for a
 * neighbouring thread block's data (a broken launch: the maps of that call are invalid) */
int rtdm_sgbm_wait(rtdm_sgbm *h);
int rtdm_sgbm_wait_oldest(rtdm_sgbm *h);
int rtdm_sgbm_compute_device(rtdm_sgbm *h, int n, const uint8_t *left, size_t lstep, size_t lframe,
                             const uint8_t *right, size_t rstep, size_t rframe, int width,
                             int height, int16_t *disp, size_t dstep, size_t dframe,
                             void *cuda_stream);
int rtdm_sgbm_last_launches(const rtdm_sgbm *h);
/* (numDisparities 48 / 64 / 96 / 128 / 192 -- the reference's default -nd 192 at 1280 pixels, scaled by the width as
 * utils/cmdline-parser.h:85-89 does -- take the specialised kernels; the other multiples of 16 up to 256 take generic ones.)
 * Frames the matching stage works on at the same time for this frame size (one thread-block cluster per frame in the
 * whole-height aggregation passes: 15 on a B200 at 1280x720x128): batches that are multiples of it keep every cluster
 * busy to the end.  1 when the size takes the tiled sweeps (any batch size is as good as another), < 0 on error.
 * No reference peer: SWSemiGlobalMatcher::compute (sgbm-sw.cpp:32-37) is one frame per call. */
int rtdm_sgbm_batch_quantum(rtdm_sgbm *h, int width, int height);
/* CUDA-event stage timing like rtdm_bm_set_profiling: ms_sum[0] = matching (planes, BT cost, box sums, all path
 * launches, WTA), ms_sum[1] = median + speckle */
int rtdm_sgbm_set_profiling(rtdm_sgbm *h, int on);
int rtdm_sgbm_stage_times(rtdm_sgbm *h, double *ms_sum, int *calls);

/* ---- SWMorphologicalFilter peer ---------------------------------------------------------- */
/* replaces SWMorphologicalFilter::SWMorphologicalFilter(w, h, bpp) (filter/mf-sw.cpp:10-17);
 * bpp must be 8.  The in/out frame buffers are PINNED host memory owned by the handle, the
 * analogue of VideoFilterDevice::video_in / video_out (include/filter/filter.h:33-34). */
int rtdm_morph_create(rtdm_morph **out, int width, int height, int bpp, int max_batch, int device);
void rtdm_morph_destroy(rtdm_morph *h);
/* replace VideoFilterDevice::getVideoInBuffer / getVideoOutBuffer (filter/filter.cpp:45-53) */
uint8_t *rtdm_morph_in_buffer(rtdm_morph *h);
uint8_t *rtdm_morph_out_buffer(rtdm_morph *h);
/* replaces SWMorphologicalFilter::run (filter/mf-sw.cpp:19-28): erode, dilate, dilate, erode with
 * the 10x10 MORPH_ELLIPSE (MORPH_FILTER_DX/DY, include/filter/mf-sw.h:11-12).  HOST pointers
 * (tightly packed width*height bytes); in/out may be the handle's own buffers.  Returns 0. */
int rtdm_morph_run(rtdm_morph *h, const uint8_t *in, uint8_t *out);
/* batched host variant: n tightly packed frames (n <= max_batch) */
int rtdm_morph_run_batch(rtdm_morph *h, int n, const uint8_t *in, uint8_t *out);
/* asynchronous batched host variant: enqueues H2D, the kernels and D2H on the handle's stream and returns; in / out
 * should be pinned host memory and must stay valid until rtdm_morph_sync returns.  Lets the filter's transfers
 * overlap a matcher call issued in between. */
int rtdm_morph_run_batch_async(rtdm_morph *h, int n, const uint8_t *in, uint8_t *out);
int rtdm_morph_sync(rtdm_morph *h);
/* device variant: n tightly packed frames, asynchronous on cuda_stream */
int rtdm_morph_run_device(rtdm_morph *h, int n, const uint8_t *in, uint8_t *out, void *cuda_stream);
int rtdm_morph_last_launches(const rtdm_morph *h);

/* ---- stand-alone stages (host pointers; used by the parity tests) -------------------------- */
/* cv::filterSpeckles(img, newVal, maxSpeckleSize, maxDiff) on CV_16SC1, in place */
int rtdm_filter_speckles(int16_t *img, size_t step, int width, int height, int newVal,
                         int maxSpeckleSize, int maxDiff, int device);
/* cv::medianBlur(src, dst, 3) on CV_16SC1 */
int rtdm_median3_s16(const int16_t *src, size_t sstep, int16_t *dst, size_t dstep, int width,
                     int height, int device);
/* cv::erode (op 0) / cv::dilate (op 1) with getStructuringElement(MORPH_ELLIPSE, (kw, kh)),
 * kw, kh <= 31, default anchor and border */
int rtdm_morph_op(const uint8_t *src, size_t sstep, uint8_t *dst, size_t dstep, int width,
                  int height, int kw, int kh, int op, int device);
/* cv::validateDisparity on CV_16SC1 disparity + CV_16SC1 cost, in place */
int rtdm_validate_disparity(int16_t *disp, size_t dstep, const int16_t *cost, size_t cstep,
                            int width, int height, int minDisparity, int numDisparities,
                            int disp12MaxDiff, int device);

/* ---- depth epilogue (the step after the matcher) -------------------------- */
/* Replaces, fused, what Estimator::run does with the matcher's output (estimator.cpp:75-77):
 *     left_disp /= 16.;                                   CV_16S, round half to even
 *     reprojectImageTo3D(left_disp, xyz, Q, true, CV_32F) Z = 10000 at the image's minimum disparity
 *     calc_depth(xyz, ., filter_out, ., obj_boundings, .) estimator.cpp:206-263: per rectangle the mean of Z over
 *                                                         pixels with mask != 0, Z != 10000, |Z| <= 10000
 * disp  : the matcher's x16 CV_16S output (NOT yet divided), dstep in bytes
 * Q     : 16 doubles, row-major 4x4 (cv::stereoRectify's reprojection matrix)
 * mask  : the filter's output (CV_8UC1) or NULL for "every pixel", mstep in bytes
 * rects : nregions x (x, y, width, height), inside the image (else -EINVAL, like cv::Mat::operator()(Rect))
 * mean_z, count : nregions results (mean_z = 0 where count = 0); the reference's label is mean_z * unit / 10 cm
 * xyz   : optional CV_32FC3 output image (height x width x 3 floats, xstep in bytes), NULL to skip
 * The sums are accumulated in double like the reference's, but in parallel order: mean_z agrees with a serial
 * sum to ~1e-13 relative. */
typedef struct rtdm_depth rtdm_depth;
int rtdm_depth_create(rtdm_depth **out, int max_width, int max_height, int max_regions, int device);
void rtdm_depth_destroy(rtdm_depth *h);
/* HOST pointers, synchronous */
int rtdm_depth_run(rtdm_depth *h, const int16_t *disp, size_t dstep, int width, int height, const double *Q,
                   const uint8_t *mask, size_t mstep, int nregions, const int *rects,
                   double *mean_z, int *count, float *xyz, size_t xstep);
/* disp / mask / xyz are DEVICE pointers (e.g. what rtdm_bm_compute_device and rtdm_morph_run_device just wrote);
 * Q, rects, mean_z, count stay host pointers; the call returns after the few result bytes have arrived */
int rtdm_depth_run_device(rtdm_depth *h, const int16_t *disp, size_t dstep, int width, int height, const double *Q,
                          const uint8_t *mask, size_t mstep, int nregions, const int *rects,
                          double *mean_z, int *count, float *xyz, size_t xstep, void *cuda_stream);
int rtdm_depth_last_launches(const rtdm_depth *h);

/* ---- rectification front-end (the step before the matcher) ---------------- */
/* Replaces, fused, estimator.cpp:29-36 for one camera:
 *     cvtColor(img, gray, CV_RGB2GRAY); remap(gray, rect, map1, map2, INTER_LINEAR); rect = rect(roif);
 * map1 (CV_16SC2) and map2 (CV_16UC1) are the fixed-point maps of initUndistortRectifyMap(..., CV_16SC2, ...)
 * (main.cpp:95-96), HOST pointers, full image size (src_height x src_width), steps in bytes; they are copied
 * to the device once.  roi = Estimator's roif; the output is the CV_8UC1 crop the matcher receives. */
typedef struct rtdm_rectify rtdm_rectify;
int rtdm_rectify_create(rtdm_rectify **out, int src_width, int src_height, const int16_t *map1, size_t map1_step,
                        const uint16_t *map2, size_t map2_step, int roi_x, int roi_y, int roi_width, int roi_height,
                        int max_batch, int device);
void rtdm_rectify_destroy(rtdm_rectify *h);
/* n HOST frames CV_8UC3 in R,G,B order (what the decoder delivers, estimator.cpp:24-27) -> n crops */
int rtdm_rectify_run(rtdm_rectify *h, int n, const uint8_t *rgb, size_t step, size_t frame,
                     uint8_t *out, size_t ostep, size_t oframe);
/* DEVICE pointers, asynchronous on cuda_stream: frames can stay on the GPU from here to the depth epilogue */
int rtdm_rectify_run_device(rtdm_rectify *h, int n, const uint8_t *rgb, size_t step, size_t frame,
                            uint8_t *out, size_t ostep, size_t oframe, void *cuda_stream);
int rtdm_rectify_last_launches(const rtdm_rectify *h);

/* ---- mask front-end (the step before the morphological filter) ------------ */
/* Replaces, fused, estimator.cpp:38-43:
 *     remap(img[0], img_rectified, map1, map2, INTER_LINEAR); img_rectified = img_rectified(roif);
 *     cvtColor(img_rectified, img_rectified, COLOR_RGB2BGR); cvtColor(img_rectified, imgHSV, COLOR_BGR2HSV);
 *     inRange(imgHSV, Scalar(iLowH, iLowS, iLowV), Scalar(iHighH, iHighS, iHighV), filter_in);
 * maps / roi as for rtdm_rectify_create.  low / high: 3 ints each, (H, S, V), H in [0, 180).
 * mask: CV_8UC1 roi_height x roi_width, 0 / 255 -- what SWMorphologicalFilter::run receives (it may be the filter
 * plugin's own input buffer).  bgr: optional CV_8UC3 rectified crop in B,G,R order (the image the reference
 * displays and labels), NULL to skip. */
typedef struct rtdm_colormask rtdm_colormask;
int rtdm_colormask_create(rtdm_colormask **out, int src_width, int src_height, const int16_t *map1, size_t map1_step,
                          const uint16_t *map2, size_t map2_step, int roi_x, int roi_y, int roi_width, int roi_height,
                          int max_batch, int device);
void rtdm_colormask_destroy(rtdm_colormask *h);
/* n HOST frames CV_8UC3 in R,G,B order (what the decoder delivers) -> n masks (and n BGR crops) */
int rtdm_colormask_run(rtdm_colormask *h, int n, const uint8_t *rgb, size_t step, size_t frame, const int *low, const int *high,
                       uint8_t *mask, size_t mstep, size_t mframe, uint8_t *bgr, size_t bstep, size_t bframe);
/* DEVICE pointers, asynchronous on cuda_stream (e.g. mask = the device input of rtdm_morph_run_device) */
int rtdm_colormask_run_device(rtdm_colormask *h, int n, const uint8_t *rgb, size_t step, size_t frame, const int *low, const int *high,
                              uint8_t *mask, size_t mstep, size_t mframe, uint8_t *bgr, size_t bstep, size_t bframe, void *cuda_stream);
int rtdm_colormask_last_launches(const rtdm_colormask *h);

/* ---- mask back-end (the step after the morphological filter) -------------- */
/* Replaces estimator.cpp:46-53 with Estimator::fill_bounding_rects_of_contours (:164-175) and
 * Estimator::find_relevant_matching_region (:177-204):
 *     findContours(filter_out, contours, hierarchy, CV_RETR_EXTERNAL, CV_CHAIN_APPROX_SIMPLE);
 *     boundingRect per top-level contour, kept if area() >= min_obj_size, in OpenCV's contour order;
 *     the rectangle spanning the kept boxes = what bm->setROI1() receives.
 * mask: CV_8UC1, any non-zero value is foreground.  rects: up to max_regions x (x, y, width, height).
 * *count = boxes kept; *ncontours = contours.size() (the reference skips the matcher when it is 0);
 * roi = spanning rectangle; with no box kept it is the reference's own (1000000, 1000000, -2000000, -2000000).
 * More than max_regions boxes: -EINVAL (nothing is truncated silently). */
typedef struct rtdm_regions rtdm_regions;
int rtdm_regions_create(rtdm_regions **out, int max_width, int max_height, int max_regions, int device);
void rtdm_regions_destroy(rtdm_regions *h);
/* HOST mask, synchronous */
int rtdm_regions_run(rtdm_regions *h, const uint8_t *mask, size_t mstep, int width, int height, int min_obj_size,
                     int *rects, int *count, int *ncontours, int *roi);
/* DEVICE mask (e.g. what rtdm_morph_run_device just wrote); the few result ints come back to the host pointers
 * before the call returns (they feed bm->setROI1 and rtdm_depth_run*'s rects, both host-side arguments) */
int rtdm_regions_run_device(rtdm_regions *h, const uint8_t *mask, size_t mstep, int width, int height, int min_obj_size,
                            int *rects, int *count, int *ncontours, int *roi, void *cuda_stream);
int rtdm_regions_last_launches(const rtdm_regions *h);

/* ---- measurement helper ------------------------------------------------------------------- */
/* Measures the integer-ALU issue peak of the device with dependent-free packed-integer loops
 * (the roofline denominator SURVEY.md 8(d) asks for).  Results in 1e12 lane-ops/s. */
int rtdm_measure_int_peak(int device, double *tiops_iadd3, double *tiops_vimnmx,
                          double *tiops_vabsdiff4, double *sm_mhz_est);

#ifdef __cplusplus
}
#endif
#endif /* RTDM_B200_H_ */
