// common.cuh -- shared declarations for the sm_100a stereo kernels and their launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>

#include "../../include/rtdm_b200.h"

namespace rtdm {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define RTDM_CUDA(expr)                                                        \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) return ::rtdm::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// image planes: n frames, `pitch` bytes (u8) or elements (i16) per row, `frame` per frame
// ---------------------------------------------------------------------------------------------
// Development switches: kernel selection for A/B timing runs and for the tests' kernel matrix.  They come from the
// environment and are read ONCE, by read_switches() in rtdm_*_create (and the stand-alone host stages) -- never on the
// compute path of a live handle.  None of them changes a result; the one that did (RTDM_BM_DEBUG, which skips whole
// stages for timing ablations) only exists in builds with -DRTDM_DEV.
struct Switches {
    int bm_kernel = 0;          // RTDM_BM_KERNEL: 1 = generic bm_sad.cu, 2 = bm_sad2.cu, 3 = bm_sad3.cu where bm_sad4.cu would run
    int bm3_shape = 0;          // RTDM_BM3_SHAPE: 1 = two 384-thread CTAs per SM
    int bm_chunk = 0;           // RTDM_BM_CHUNK: frames per chunk of the host batch pipeline (0 = automatic)
    int bm_fork_min = 16;       // RTDM_BM_FORK_MIN: smallest call whose row pass + speckle filter run as parts on several streams (0 = never)
    int bm_fork_parts = 2;      // RTDM_BM_FORK_PARTS: 2 .. 4 parts
    int bm_variant = 1;         // RTDM_BM_VARIANT (bm_sad2.cu CTA variants)
    int bm_occ3 = 0;            // RTDM_BM_OCC3
    int bm_debug = 0;           // RTDM_BM_DEBUG, RTDM_DEV builds only
    int bm_nofuse = 0;          // RTDM_BM_NOFUSE: separate texture kernel instead of the fused prefilter + texture kernel
    int speckle_scalar = 0;     // RTDM_SPECKLE_SCALAR: per-pixel speckle kernels
    int post_unfused = 0;       // RTDM_POST_UNFUSED: separate validate / row-run kernels
    int sgbm_oldcost = 0, sgbm_oldpath = 0, sgbm_nofuse = 0, sgbm_nosweep = 0;   // RTDM_SGBM_*
    int sgbm_sweep_rows = 0;    // RTDM_SGBM_SWEEP_ROWS: rows per sweep launch (0 = default)
    int sgbm_novpass = 0;       // RTDM_SGBM_NOVPASS: tiled row sweeps (sgbm_sweep_kernel) instead of the whole-height cluster pass
    int sgbm_vpass_min = 0;     // RTDM_SGBM_VPASS_MIN: smallest batch that takes the whole-height pass (0 = half the resident clusters)
    int sgbm_vpass_shape = 0;   // RTDM_SGBM_VPASS_SHAPE: 1 = 512 threads x 4 columns per thread instead of 1024 x 2
    int sgbm_vpass_maxcl = 0;   // RTDM_SGBM_VPASS_MAXCL: cap on the clusters of a pass (tests: forces several frames per cluster)
};
Switches read_switches();

struct PlaneU8 { const uint8_t *p; size_t pitch; size_t frame; };
struct PlaneU8W { uint8_t *p; size_t pitch; size_t frame; };
struct PlaneS16 { int16_t *p; size_t pitch; size_t frame; };   // pitch / frame in ELEMENTS

// Prefiltered planes in the layouts bm_sad4.cu's bulk copies read (all pitches / frame steps in ELEMENTS):
//   LE  u32 [frame][H][le_pitch]: LE[LPADL + x] = left pixel x times 0x01010101 for x < W, the last pixel again for W <= x < W + LPADR
//   RP  u8  [frame][H][rp_pitch]: RP[RPADL + x] = right pixel x; RPADL copies of pixel 0 before, RPADR copies of pixel W - 1 after
// Row starts are 16-byte aligned (le_pitch % 4 == 0, rp_pitch % 16 == 0); the planes end with >= 4 KB of slack (a copy may
// run past the last row).  The replicated columns are App. A.2's clamps.  LPADL = 1 makes the first left column of every
// stripe (x0 - HP + numDisparities - 1, x0 and HP multiples of 4) a 16-byte boundary.
struct BmStaged {
    static constexpr int LPADL = 1, LPADR = 8, RPADL = 16, RPADR = 16;
    uint32_t *LE; size_t le_pitch, le_frame;
    uint8_t *RP; size_t rp_pitch, rp_frame;
};

// ---- prefilter (prefilter.cu) ------------------------------------------------------------------
// type: RTDM_PREFILTER_*.  Both images of all n frames in one launch.  With `staged` both images are written in their
// staged layouts only (outL / outR unused).
int launch_prefilter(int type, int winsize, int cap, int n, int W, int H,
                     PlaneU8 left, PlaneU8 right, PlaneU8W outL, PlaneU8W outR,
                     cudaStream_t st, int *launches, const BmStaged *staged = nullptr);

// ---- block matching core (bm_sad.cu) -----------------------------------------------------------
struct BmGeom {
    int W, H, nd, minD, bs, cap, texThr, uniq;
    int lofs, rofs, W1;
    int row0, row1;          // rows to compute (valid rect rows)
    // minDisparity > 0 only: cv::StereoBM writes the computed columns lofs + x >= W of a row into the first minD pixels
    // of the NEXT row; all of them are overwritten or masked again except those of the last
    // computed row, which stay in row `row1`.  spill[frame * minD + k] receives them (nullptr: not wanted)
    int16_t *spill = nullptr;
    Switches sw;
};
size_t bm_sad_smem_bytes(const BmGeom &g, int TW, int BH);
int launch_bm_sad_wta(const BmGeom &g, int n, PlaneU8 Lp, PlaneU8 Rp, PlaneS16 disp, PlaneS16 cost,
                      cudaStream_t st, int *launches);

// fast path (bm_sad2.cu): minDisparity == 0, blockSize 5..15; tex = n * tex_frame uint16 scratch
bool bm_sad2_supported(const BmGeom &g, int n);
int launch_bm_sad2(const BmGeom &g, int n, PlaneU8 Lp, PlaneU8 Rp, PlaneS16 disp, PlaneS16 cost,
                   uint16_t *tex, size_t tex_pitch, size_t tex_frame, cudaStream_t st, int *launches, bool use3 = false);
// warp-specialised fast path (bm_sad3.cu): minDisparity == 0, blockSize 5 .. 15, numDisparities 32 / 48 / 64 / 96 / 128 / 192 / 256
bool bm_sad3_supported(const BmGeom &g, int n);
long long bm_sad3_cost(const BmGeom &g, int n);
int launch_bm_sad3_core(const BmGeom &g, int n, PlaneU8 Lp, PlaneU8 Rp, PlaneS16 disp, PlaneS16 cost,
                        const uint16_t *tex, size_t tex_pitch, size_t tex_frame, cudaStream_t st);

// TMA-staged warp-specialised fast path (bm_sad4.cu): same domain as bm_sad3.cu; reads the BmStaged planes
bool bm_sad4_supported(const BmGeom &g, int n);
long long bm_sad4_cost(const BmGeom &g, int n);
int launch_bm_sad4_core(const BmGeom &g, int n, const BmStaged &sp, PlaneS16 disp, PlaneS16 cost, cudaStream_t st);
// texture window sums of the prefiltered left image (bm_sad2.cu: bm_texture_kernel), blockSize 5 .. 15
int launch_bm_texture(const BmGeom &g, int n, PlaneU8 Lp, uint16_t *tex, size_t tex_pitch, size_t tex_frame, cudaStream_t st);

// ---- post-processing (postproc.cu) -------------------------------------------------------------
// validateDisparity (if d12 >= 0) + valid-rect mask; reads raw disp/cost, writes `out`
int launch_validate_mask(int n, int W, int H, int minD, int nd, int d12, int lofs, int W1,
                         int vx0, int vx1, int row0, int row1,
                         PlaneS16 raw, PlaneS16 cost, PlaneS16 out, cudaStream_t st, int *launches,
                         const int16_t *spill = nullptr);
// filterSpeckles on n frames in place; labels, sizes, runlen: n*W*H int32 scratch each
int launch_speckle(int n, int W, int H, PlaneS16 img, int newVal, int maxSize, int maxDiff,
                   int32_t *labels, int32_t *sizes, cudaStream_t st, int *launches, int32_t *runlen,
                   const Switches &sw = Switches());
int launch_validate_speckle(int n, int W, int H, int minD, int nd, int d12, int lofs, int W1,
                            int vx0, int vx1, int row0, int row1, PlaneS16 raw, PlaneS16 cost, PlaneS16 out,
                            bool speckle, int newVal, int maxSize, int maxDiff,
                            int32_t *labels, int32_t *sizes, int32_t *runlen, cudaStream_t st, int *launches,
                            void (*after_rows)(void *) = nullptr, void *ctx = nullptr,    // hook between the row pass and the rest (stage timing)
                            const int16_t *spill = nullptr,                               // BmGeom::spill (minDisparity > 0)
                            const Switches &sw = Switches());
int launch_median3(int n, int W, int H, PlaneS16 src, PlaneS16 dst, cudaStream_t st, int *launches);

// ---- morphology (morph.cu) ---------------------------------------------------------------------
struct MorphSE {
    int kw, kh, ax, ay; int j1[32], j2[32];
    // distinct horizontal runs of the rows (an ellipse has few): run u = [rj1[u], rj1[u] + rL[u]), row k uses run rowrun[k] (-1: empty row)
    int nrun, rj1[32], rL[32], rowrun[32];
};
void make_ellipse(int kw, int kh, MorphSE *se);
// need: optional per-frame flags; frames whose flag is 0 are skipped (nullptr = process all)
int launch_morph(int n, int W, int H, PlaneU8 src, PlaneU8W dst, const MorphSE &se, int op,
                 cudaStream_t st, int *launches, const int *need = nullptr);
// erode, dilate, dilate, erode.  ta / tb / fast: n-frame scratch planes, flags: n ints
int launch_morph_openclose(int n, int W, int H, PlaneU8 src, PlaneU8W dst, PlaneU8W ta, PlaneU8W tb, PlaneU8W fast,
                           int *flags, const MorphSE &se, cudaStream_t st, int *launches);

// ---- SGBM (sgbm.cu) ----------------------------------------------------------------------------
struct SgbmGeom {
    int W, H, D, minD, bs, P1, P2, uniq, d12, ftzero, mode;
    int minX1, maxX1, W1;
    Switches sw;
};
struct SgbmWork {
    uint8_t *planes;    // per frame: BT planes (see sgbm.cu)
    int16_t *pix;       // per-pixel cost rows / horizontal sums
    int16_t *C;         // cost volume      [H][W1][D]
    int16_t *S;         // aggregated volume [H][W1][D]
    int16_t *disp2;     // per-row scratch
    int *err;           // device pointer of a host-mapped flag the whole-height pass raises when its neighbour exchange breaks
    size_t frame_planes, frame_vol;
};
size_t sgbm_work_bytes(const SgbmGeom &g, size_t *planes, size_t *vol);
int sgbm_sweep_ctas_per_frame(const SgbmGeom &g);
int sgbm_vpass_frames_in_flight(const SgbmGeom &g, int n);
int launch_sgbm(const SgbmGeom &g, int n, PlaneU8 left, PlaneU8 right, PlaneS16 out, SgbmWork w,
                cudaStream_t st, int *launches);

// ---- depth epilogue (depth.cu): /16, reprojectImageTo3D, masked mean Z per rectangle ---------------
int launch_depth(const int16_t *disp, size_t dpitch, int W, int H, const double *Q, const uint8_t *mask, size_t mpitch,
                 int nregions, const int *rects_dev, int *minval, double *sums, int *counts, float *xyz, size_t xpitch,
                 cudaStream_t st, int *launches);

// ---- rectification front-end (rectify.cu): RGB -> gray -> remap(INTER_LINEAR, fixed-point maps) -> ROI crop ------
int launch_rectify(int n, const uint8_t *rgb, size_t pitch, size_t frame, int W, int H, const int16_t *map1, const uint16_t *map2,
                   int rw, int rh, uint8_t *out, size_t opitch, size_t oframe, cudaStream_t st, int *launches);

// ---- mask front-end / back-end (mask.cu): colour threshold before the filter, object boxes after it ---------------
int launch_colormask(int n, const uint8_t *rgb, size_t pitch, size_t frame, int W, int H, const int16_t *map1, const uint16_t *map2,
                     int rw, int rh, const int *lo, const int *hi, uint8_t *mask, size_t mpitch, size_t mframe,
                     uint8_t *bgr, size_t bpitch, size_t bframe, cudaStream_t st, int *launches);
int launch_regions(const uint8_t *mask, size_t pitch, int W, int H, int minSize, int maxR,
                   int *labels, void *bb, int *ext, int *keys, void *boxes, int *out, cudaStream_t st, int *launches);

// ---- int peak microbenchmark (intpeak.cu) -------------------------------------------------------
int measure_int_peak(int device, double *iadd3, double *vimnmx, double *vabsdiff4, double *mhz);

}  // namespace rtdm
