// api.cu -- the extern "C" boundary (include/rtdm_b200.h): handles, workspaces, pipelines.
//
// Host side of the drop-in: what SWMatcherKonolige / SWSemiGlobalMatcher / SWMorphologicalFilter do
// on the CPU through OpenCV (reference stereo-matcher/bm-sw.cpp, sgbm-sw.cpp, filter/mf-sw.cpp) is
// issued here as a fixed sequence of sm_100a kernels on a CUDA stream.  No CPU fallback exists: every
// compute entry point fails with -RTDM_ENODEV when no device is usable.
#include "common.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace rtdm {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }

// the one place the library looks at the environment (see common.cuh: Switches); called by rtdm_*_create
Switches read_switches()
{
    Switches sw;
    auto geti = [](const char *name, int dflt) { const char *e = getenv(name); return e && *e ? atoi(e) : dflt; };
    auto isset = [](const char *name) { const char *e = getenv(name); return (e && *e) ? 1 : 0; };
    sw.bm_kernel = geti("RTDM_BM_KERNEL", 0);
    sw.bm3_shape = geti("RTDM_BM3_SHAPE", 0);
    sw.bm_chunk = geti("RTDM_BM_CHUNK", 0);
    sw.bm_fork_min = geti("RTDM_BM_FORK_MIN", 16);
    sw.bm_fork_parts = geti("RTDM_BM_FORK_PARTS", 2);
    sw.bm_variant = geti("RTDM_BM_VARIANT", 1);
    sw.bm_occ3 = isset("RTDM_BM_OCC3");
    sw.bm_nofuse = isset("RTDM_BM_NOFUSE");
#ifdef RTDM_DEV
    sw.bm_debug = geti("RTDM_BM_DEBUG", 0);
#endif
    sw.speckle_scalar = isset("RTDM_SPECKLE_SCALAR");
    sw.post_unfused = isset("RTDM_POST_UNFUSED");
    sw.sgbm_oldcost = isset("RTDM_SGBM_OLDCOST");
    sw.sgbm_oldpath = isset("RTDM_SGBM_OLDPATH");
    sw.sgbm_nofuse = isset("RTDM_SGBM_NOFUSE");
    sw.sgbm_nosweep = isset("RTDM_SGBM_NOSWEEP");
    sw.sgbm_sweep_rows = geti("RTDM_SGBM_SWEEP_ROWS", 0);
    sw.sgbm_novpass = isset("RTDM_SGBM_NOVPASS");
    sw.sgbm_vpass_min = geti("RTDM_SGBM_VPASS_MIN", 0);
    sw.sgbm_vpass_maxcl = geti("RTDM_SGBM_VPASS_MAXCL", 0);
    sw.sgbm_vpass_shape = geti("RTDM_SGBM_VPASS_SHAPE", 0);
    return sw;
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_err = buf;
    cudaGetLastError();   // clear sticky-less errors
    if (e == cudaErrorMemoryAllocation) return -RTDM_ENOMEM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice ||
        e == cudaErrorNoKernelImageForDevice || e == cudaErrorInitializationError)
        return -RTDM_ENODEV;
    return -RTDM_EIO;
}

static int check_device(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no CUDA device available (this library has no CPU fallback)");
        return -RTDM_ENODEV;
    }
    if (device < 0 || device >= n) { set_error("invalid device index"); return -RTDM_ENODEV; }
    return 0;
}

template <typename T>
static int dev_alloc(T **p, size_t count)
{
    *p = nullptr;
    if (count == 0) return 0;
    RTDM_CUDA(cudaMalloc((void **)p, count * sizeof(T)));
    return 0;
}

struct ValidRect { int x, y, w, h; };

// getValidDisparityROI (oracle: orc_valid_roi)
static ValidRect valid_rect(const int roi1[4], const int roi2[4], int W, int H, int minD, int nd, int bs)
{
    int r1[4] = {0, 0, W, H}, r2[4] = {0, 0, W, H};
    if (roi1[2] > 0 && roi1[3] > 0) memcpy(r1, roi1, sizeof r1);
    if (roi2[2] > 0 && roi2[3] > 0) memcpy(r2, roi2, sizeof r2);
    int h = bs / 2, maxD = minD + nd - 1;
    int xmin = std::max(r1[0], r2[0] + maxD) + h;
    int xmax = std::min(r1[0] + r1[2], r2[0] + r2[2]) - h;
    int ymin = std::max(r1[1], r2[1]) + h;
    int ymax = std::min(r1[1] + r1[3], r2[1] + r2[3]) - h;
    ValidRect v = {0, 0, 0, 0};
    if (xmax - xmin > 0 && ymax - ymin > 0) { v.x = xmin; v.y = ymin; v.w = xmax - xmin; v.h = ymax - ymin; }
    return v;
}

}  // namespace rtdm

using namespace rtdm;

// =================================================================================================
// BM
// =================================================================================================
struct rtdm_bm {
    rtdm_params p;
    Switches sw;                 // development switches as they were when the handle was created
    int maxW, maxH, maxB, dev;
    cudaStream_t st;
    cudaStream_t lane[3];        // streams of the chunked host-batch pipeline (copy/compute overlap)
    cudaStream_t aux[3];         // extra streams of the post-processing fork (bm_pipeline)
    cudaEvent_t ev_fork[4];      // [0]: SAD stage done (main -> aux streams), [1 + i]: aux stream i done (-> main)
    // per-batch device workspace (maxB frames)
    uint8_t *Lp, *Rp;            size_t ppitch, pframe;      // prefiltered planes (bytes)
    uint32_t *LE; uint8_t *RPs;  size_t lepitch, leframe, rppitch, rpframe;   // staged layouts for bm_sad4.cu's TMA boxes (common.cuh: BmStaged)
    int16_t *raw, *cost;         size_t rpitch, rframe;      // raw WTA disparity + cost (elements)
    int32_t *labels, *sizes, *runlen;
    uint16_t *tex;                                           // texture window sums (rpitch / rframe)
    int16_t *spill;                                          // minDisparity > 0: [maxB][minD], see BmGeom::spill
    // staging for the host entry points
    uint8_t *dL, *dR;            size_t spitch, sframe;      // device copies of the inputs  (staging set 0)
    int16_t *dD;                 size_t dpitch, dframe;      // device copy of the output (elements)
    uint8_t *dL2, *dR2; int16_t *dD2;                        // staging set 1 (rtdm_bm_submit_batch alternates)
    cudaEvent_t done[2]; int busy[2]; unsigned seq;          // completion of the call that last used each set
    int launches;
    int lastW, lastH;
    int last_kernel;             // 1 = generic bm_sad.cu kernel, 2 = fast bm_sad2.cu kernel, 3 = warp-specialised bm_sad3.cu kernel, 4 = TMA-staged bm_sad4.cu kernel
    // optional per-stage CUDA-event timing (rtdm_bm_set_profiling)
    int prof;
    std::vector<cudaEvent_t> *ev;     // 5 events per profiled call: before prefilter, after each stage
    std::vector<cudaEvent_t> *pev;    // events of the host-batch pipeline (2 per chunk)
};

static const int BM_STAGES = 4;      // prefilter, sad+wta, validate+mask, speckle

static int bm_check_params(const rtdm_params *p)
{
    if (p->preFilterType != RTDM_PREFILTER_NORMALIZED_RESPONSE && p->preFilterType != RTDM_PREFILTER_XSOBEL) { set_error("bm: preFilterType must be 0 or 1"); return -RTDM_EINVAL; }
    if (p->preFilterSize < 5 || p->preFilterSize > 255 || p->preFilterSize % 2 == 0) { set_error("bm: preFilterSize must be odd and within 5..255"); return -RTDM_EINVAL; }
    if (p->preFilterCap < 1 || p->preFilterCap > 63) { set_error("bm: preFilterCap must be within 1..63"); return -RTDM_EINVAL; }
    if (p->blockSize < 5 || p->blockSize > 255 || p->blockSize % 2 == 0) { set_error("bm: blockSize must be odd and within 5..255"); return -RTDM_EINVAL; }
    if (p->numDisparities <= 0 || p->numDisparities % 16 != 0) { set_error("bm: numDisparities must be positive and divisible by 16"); return -RTDM_EINVAL; }
    if (p->textureThreshold < 0) { set_error("bm: textureThreshold must be non-negative"); return -RTDM_EINVAL; }
    if (p->uniquenessRatio < 0) { set_error("bm: uniquenessRatio must be non-negative"); return -RTDM_EINVAL; }
    // kernel domain (documented in DESIGN.md): 16-bit window sums, at most 256 disparities
    if (p->numDisparities > 256) { set_error("bm: numDisparities > 256 is not supported"); return -RTDM_EINVAL; }
    if (2 * p->preFilterCap * p->blockSize * p->blockSize > 65535) { set_error("bm: 2*cap*blockSize^2 exceeds the 16-bit SAD domain"); return -RTDM_EINVAL; }
    if (p->disp12MaxDiff >= 0 && (p->preFilterCap > 31 || p->blockSize > 21)) {
        set_error("bm: disp12MaxDiff >= 0 with preFilterCap > 31 or blockSize > 21 is outside the bit-exact domain (SURVEY.md App. B.2)");
        return -RTDM_EINVAL;
    }
    return 0;
}

extern "C" int rtdm_abi_version(void) { return RTDM_ABI_VERSION; }

extern "C" int rtdm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char *rtdm_last_error(void) { return g_err.c_str(); }

extern "C" void rtdm_params_default_bm(rtdm_params *p)
{
    memset(p, 0, sizeof *p);
    p->preFilterType = RTDM_PREFILTER_XSOBEL; p->preFilterSize = 9; p->preFilterCap = 31;
    p->blockSize = 13; p->minDisparity = 0; p->numDisparities = 128; p->textureThreshold = 10;
    p->uniquenessRatio = 10; p->speckleWindowSize = 100; p->speckleRange = 32; p->disp12MaxDiff = 1;
}

extern "C" void rtdm_params_default_sgbm(rtdm_params *p)
{
    memset(p, 0, sizeof *p);
    p->preFilterType = RTDM_PREFILTER_XSOBEL; p->preFilterSize = 9; p->preFilterCap = 0;
    p->blockSize = 5; p->minDisparity = 0; p->numDisparities = 128; p->uniquenessRatio = 10;
    p->speckleWindowSize = 100; p->speckleRange = 32; p->disp12MaxDiff = 1;
    p->mode = RTDM_SGBM_MODE_SGBM; p->P1 = 8 * 3 * 5 * 5; p->P2 = 32 * 3 * 5 * 5;
}

extern "C" void rtdm_bm_destroy(rtdm_bm *h)
{
    if (!h) return;
    cudaSetDevice(h->dev);
    cudaFree(h->Lp); cudaFree(h->Rp); cudaFree(h->raw); cudaFree(h->cost);
    cudaFree(h->labels); cudaFree(h->sizes); cudaFree(h->runlen); cudaFree(h->dL); cudaFree(h->dR); cudaFree(h->dD); cudaFree(h->tex);
    cudaFree(h->spill); cudaFree(h->LE); cudaFree(h->RPs);
    cudaFree(h->dL2); cudaFree(h->dR2); cudaFree(h->dD2);
    for (int i = 0; i < 2; i++) if (h->done[i]) cudaEventDestroy(h->done[i]);
    if (h->ev) { for (cudaEvent_t e : *h->ev) cudaEventDestroy(e); delete h->ev; }
    if (h->pev) { for (cudaEvent_t e : *h->pev) cudaEventDestroy(e); delete h->pev; }
    if (h->st) cudaStreamDestroy(h->st);
    for (int i = 0; i < 3; i++) if (h->lane[i]) cudaStreamDestroy(h->lane[i]);
    for (int i = 0; i < 3; i++) if (h->aux[i]) cudaStreamDestroy(h->aux[i]);
    for (int i = 0; i < 4; i++) if (h->ev_fork[i]) cudaEventDestroy(h->ev_fork[i]);
    delete h;
}

extern "C" int rtdm_bm_create(rtdm_bm **out, const rtdm_params *p, int max_width, int max_height,
                              int max_batch, int device)
{
    if (!out || !p) { set_error("bm_create: null argument"); return -RTDM_EINVAL; }
    *out = nullptr;
    int rc = bm_check_params(p);
    if (rc) return rc;
    if (max_width < 1 || max_height < 1 || max_batch < 1 || max_width > 8000 || max_height > 65535) {
        set_error("bm_create: bad maximum geometry (width <= 8000, height <= 65535, batch >= 1)");
        return -RTDM_EINVAL;
    }
    rc = check_device(device);
    if (rc) return rc;
    RTDM_CUDA(cudaSetDevice(device));
    rtdm_bm *h = new (std::nothrow) rtdm_bm();      // value-initialised: every member zero, switches at their defaults
    if (!h) return -RTDM_ENOMEM;
    h->p = *p; h->maxW = max_width; h->maxH = max_height; h->maxB = max_batch; h->dev = device;
    h->sw = read_switches();
    const size_t B = (size_t)max_batch;
    h->ppitch = align_up((size_t)max_width + 160, 64);          // over-read slack for the band loader
    h->pframe = h->ppitch * max_height;
    h->rpitch = align_up((size_t)max_width, 8); h->rframe = h->rpitch * max_height;
    h->spitch = align_up((size_t)max_width, 64); h->sframe = h->spitch * max_height;
    h->dpitch = h->rpitch; h->dframe = h->rframe;
    rc = (int)cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) == cudaSuccess ? 0 : -RTDM_EIO;
    for (int i = 0; i < 3 && !rc; i++)
        rc = (int)cudaStreamCreateWithFlags(&h->lane[i], cudaStreamNonBlocking) == cudaSuccess ? 0 : -RTDM_EIO;
    for (int i = 0; i < 3 && !rc; i++)
        rc = cudaStreamCreateWithFlags(&h->aux[i], cudaStreamNonBlocking) == cudaSuccess ? 0 : -RTDM_EIO;
    for (int i = 0; i < 4 && !rc; i++)
        rc = cudaEventCreateWithFlags(&h->ev_fork[i], cudaEventDisableTiming) == cudaSuccess ? 0 : -RTDM_EIO;
    if (!rc) rc = dev_alloc(&h->Lp, h->pframe * B + 4096);
    if (!rc) rc = dev_alloc(&h->Rp, h->pframe * B + 4096);
    // staged planes only where bm_sad4.cu can run (minDisparity 0, blockSize 5 .. 15)
    h->lepitch = align_up((size_t)max_width + BmStaged::LPADL + BmStaged::LPADR, 4); h->leframe = h->lepitch * max_height;
    h->rppitch = align_up((size_t)max_width + BmStaged::RPADL + BmStaged::RPADR, 16); h->rpframe = h->rppitch * max_height;
    if (!rc && p->minDisparity == 0 && p->blockSize <= 15) {
        rc = dev_alloc(&h->LE, h->leframe * B + 1024);          // + 4 KB: a bulk copy may run past the last row
        if (!rc) rc = dev_alloc(&h->RPs, h->rpframe * B + 4096);
    }
    if (!rc) rc = dev_alloc(&h->raw, h->rframe * B);
    if (!rc) rc = dev_alloc(&h->cost, h->rframe * B);
    if (!rc) rc = dev_alloc(&h->tex, h->rframe * B);
    if (!rc && p->minDisparity > 0) rc = dev_alloc(&h->spill, (size_t)p->minDisparity * B);
    if (!rc) rc = dev_alloc(&h->labels, (size_t)max_width * max_height * B);
    if (!rc) rc = dev_alloc(&h->sizes, (size_t)max_width * max_height * B);
    if (!rc) rc = dev_alloc(&h->runlen, (size_t)max_width * max_height * B);
    if (!rc) rc = dev_alloc(&h->dL, h->sframe * B);
    if (!rc) rc = dev_alloc(&h->dR, h->sframe * B);
    if (!rc) rc = dev_alloc(&h->dD, h->dframe * B);
    if (!rc) rc = dev_alloc(&h->dL2, h->sframe * B);
    if (!rc) rc = dev_alloc(&h->dR2, h->sframe * B);
    if (!rc) rc = dev_alloc(&h->dD2, h->dframe * B);
    for (int i = 0; i < 2 && !rc; i++)
        rc = cudaEventCreateWithFlags(&h->done[i], cudaEventDisableTiming) == cudaSuccess ? 0 : -RTDM_EIO;
    if (rc) { rtdm_bm_destroy(h); return rc; }
    *out = h;
    return 0;
}

extern "C" int rtdm_bm_set_roi1(rtdm_bm *h, int x, int y, int w, int hgt)
{
    if (!h) return -RTDM_EINVAL;
    h->p.roi1[0] = x; h->p.roi1[1] = y; h->p.roi1[2] = w; h->p.roi1[3] = hgt;
    return 0;
}

extern "C" int rtdm_bm_set_roi2(rtdm_bm *h, int x, int y, int w, int hgt)
{
    if (!h) return -RTDM_EINVAL;
    h->p.roi2[0] = x; h->p.roi2[1] = y; h->p.roi2[2] = w; h->p.roi2[3] = hgt;
    return 0;
}

static BmGeom bm_geom(const rtdm_bm *h, int W, int H, ValidRect *vrout = nullptr)
{
    const rtdm_params &p = h->p;
    const int nd = p.numDisparities, minD = p.minDisparity;
    BmGeom g;
    g.sw = h->sw;
    g.W = W; g.H = H; g.nd = nd; g.minD = minD; g.bs = p.blockSize; g.cap = p.preFilterCap;
    g.texThr = p.textureThreshold; g.uniq = p.uniquenessRatio;
    g.lofs = std::max(nd - 1 + minD, 0); g.rofs = -std::min(nd - 1 + minD, 0);
    g.W1 = W - g.rofs - nd + 1;
    ValidRect vr = valid_rect(p.roi1, p.roi2, W, H, minD, nd, p.blockSize);
    int row0 = std::min(std::max(vr.y, 0), H), row1 = std::min(std::max(vr.y + vr.h, 0), H);
    if (vr.w == 0 || vr.h == 0 || g.lofs >= W || g.rofs >= W || g.W1 < 1) row0 = row1 = 0;
    g.row0 = row0; g.row1 = row1;
    if (vrout) *vrout = vr;
    return g;
}

// chunk size of the host batch pipeline: at least two chunks per call from 32 frames on (copy / compute overlap inside
// the call), and among those the split whose SAD/WTA launches leave the fewest partly filled waves
// (bm_sad3_cost: waves x (band height + start-up) of the launch; 720p x 128, 63 frames -> 42 + 21)
static int bm_chunk_for(const rtdm_bm *h, int n, int W, int H)
{
    int chunk = n >= 64 ? 32 : (n >= 32 ? 16 : (n >= 8 ? (n + 3) / 4 : n));
    if (n < 16 || h->p.blockSize >= W || h->p.blockSize >= H) return chunk;
    const BmGeom g = bm_geom(h, W, H);
    if (h->sw.bm_kernel == 1 || h->sw.bm_kernel == 2) return chunk;
    const bool k4 = h->sw.bm_kernel != 3 && h->LE && bm_sad4_supported(g, n);
    long long best = -1;
    for (int m = std::max(8, n / 4); m < n; m++) {
        const int k = (n + m - 1) / m, last = n - (k - 1) * m;
        const long long c1 = k4 ? bm_sad4_cost(g, m) : bm_sad3_cost(g, m), c2 = k4 ? bm_sad4_cost(g, last) : bm_sad3_cost(g, last);
        if (c1 < 0 || c2 < 0) return chunk;
        const long long score = ((k - 1) * c1 + c2) * 16 + k;       // fewer row steps first, then fewer chunks
        if (best < 0 || score < best) { best = score; chunk = m; }
    }
    return chunk;
}

// the kernel pipeline on device-resident frames
static int bm_pipeline(rtdm_bm *h, int n, PlaneU8 L, PlaneU8 R, int W, int H, PlaneS16 out, cudaStream_t st, int f0 = 0)
{
    const rtdm_params &p = h->p;
    if (W > h->maxW || H > h->maxH || f0 + n > h->maxB || W < 1 || H < 1 || n < 1) {
        set_error("bm: frame geometry or batch exceeds what the handle was created for");
        return -RTDM_EINVAL;
    }
    if (p.blockSize >= W || p.blockSize >= H) {
        set_error("bm: blockSize must be smaller than the image"); return -RTDM_EINVAL;
    }
    h->lastW = W; h->lastH = H;
    const int nd = p.numDisparities, minD = p.minDisparity;
    const int FILT = (minD - 1) * 16;
    ValidRect vr;
    BmGeom g = bm_geom(h, W, H, &vr);
    const int row0 = g.row0, row1 = g.row1;
    // minDisparity > 0: what the last computed row writes past its end lands in row `row1` (App. B.3)
    if (h->spill && row1 > row0 && row1 < H) g.spill = h->spill + (size_t)f0 * minD;
    int rc = 0;
    auto mark = [&]() {
        if (!h->prof) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, st);
        h->ev->push_back(e);
    };
    // workspace slices of frames [f0, f0 + n)
    uint8_t *wLp = h->Lp + (size_t)f0 * h->pframe, *wRp = h->Rp + (size_t)f0 * h->pframe;
    PlaneS16 raw = {h->raw + (size_t)f0 * h->rframe, h->rpitch, h->rframe};
    PlaneS16 cost = {h->cost + (size_t)f0 * h->rframe, h->rpitch, h->rframe};
    int32_t *wlab = h->labels + (size_t)f0 * W * H, *wsiz = h->sizes + (size_t)f0 * W * H, *wrun = h->runlen + (size_t)f0 * W * H;
    mark();
    bool fast4 = false;
    if (row1 > row0) {
        PlaneU8W oL = {wLp, h->ppitch, h->pframe}, oR = {wRp, h->ppitch, h->pframe};
        // Switches::bm_kernel = 1 forces the generic kernel (A/B runs and tests); default: fast path when it applies
        const bool fast = h->sw.bm_kernel != 1 && bm_sad2_supported(g, n);
        // bm_kernel = 2 keeps the bm_sad2.cu kernel where the warp-specialised kernels would apply, 3 the bm_sad3.cu kernel
        const bool fast3 = fast && h->sw.bm_kernel != 2 && bm_sad3_supported(g, n);
        fast4 = fast3 && h->sw.bm_kernel != 3 && h->LE && bm_sad4_supported(g, n);
        const BmStaged sp = {h->LE + (size_t)f0 * h->leframe, h->lepitch, h->leframe, h->RPs + (size_t)f0 * h->rpframe, h->rppitch, h->rpframe};
        rc = launch_prefilter(p.preFilterType, p.preFilterSize, p.preFilterCap, n, W, H, L, R, oL, oR, st, &h->launches, fast4 ? &sp : nullptr);
        if (rc) return rc;
        mark();
        PlaneU8 iL = {wLp, h->ppitch, h->pframe}, iR = {wRp, h->ppitch, h->pframe};
        if (fast4) {
            rc = launch_bm_sad4_core(g, n, sp, raw, cost, st);
            h->launches += 1;
        } else if (fast)
            rc = launch_bm_sad2(g, n, iL, iR, raw, cost, h->tex + (size_t)f0 * h->rframe, h->rpitch, h->rframe, st, &h->launches, fast3);
        else
            rc = launch_bm_sad_wta(g, n, iL, iR, raw, cost, st, &h->launches);
        if (fast4) h->last_kernel = 4; else
        h->last_kernel = fast3 ? 3 : (fast ? 2 : 1);
        if (rc) return rc;
        mark();
    } else { mark(); mark(); }
    // stage timing: "validate_mask" ends after the row kernel (which, fused, already holds the speckle filter's row-run pass)
    struct Hook { decltype(mark) *m; } hook = {&mark};
    const bool speckle = p.speckleRange >= 0 && p.speckleWindowSize > 0;
    // The row pass and the speckle filter are chains of small latency-bound kernels (160 / 128 / 256 threads, 32 registers): for
    // a large call the two halves of the batch go through them on two streams at once and fill each other's idle issue slots
    const int nparts = (n >= h->sw.bm_fork_min && h->sw.bm_fork_min > 0 && !g.spill) ? std::max(2, std::min(4, h->sw.bm_fork_parts)) : 1;
    if (nparts > 1) {
        RTDM_CUDA(cudaEventRecord(h->ev_fork[0], st));
        for (int part = 0; part < nparts; part++) {
            const int fa = (int)((long long)n * part / nparts), fb = (int)((long long)n * (part + 1) / nparts), m = fb - fa;
            const bool last = part == nparts - 1;                       // the last part stays on the call's stream (and carries the stage marks)
            cudaStream_t ps = last ? st : h->aux[part];
            if (!last) RTDM_CUDA(cudaStreamWaitEvent(ps, h->ev_fork[0], 0));
            const size_t px = (size_t)fa * W * H;
            const PlaneS16 rawP = {raw.p + (size_t)fa * raw.frame, raw.pitch, raw.frame}, costP = {cost.p + (size_t)fa * cost.frame, cost.pitch, cost.frame};
            const PlaneS16 outP = {out.p + (size_t)fa * out.frame, out.pitch, out.frame};
            rc = launch_validate_speckle(m, W, H, minD, nd, p.disp12MaxDiff, g.lofs, g.W1, std::max(vr.x, 0), std::max(vr.x + vr.w, 0),
                                         row0, row1, rawP, costP, outP, speckle, FILT, p.speckleWindowSize, p.speckleRange,
                                         wlab + px, wsiz + px, wrun + px, ps, &h->launches,
                                         last ? [](void *c) { (*static_cast<Hook *>(c)->m)(); } : (void (*)(void *))nullptr, last ? &hook : nullptr, nullptr, h->sw);
            if (rc) return rc;
            if (!last) RTDM_CUDA(cudaEventRecord(h->ev_fork[1 + part], ps));
        }
        for (int part = 0; part + 1 < nparts; part++) RTDM_CUDA(cudaStreamWaitEvent(st, h->ev_fork[1 + part], 0));
        mark();
        return rc;
    }
    rc = launch_validate_speckle(n, W, H, minD, nd, p.disp12MaxDiff, g.lofs, g.W1, std::max(vr.x, 0), std::max(vr.x + vr.w, 0),
                                 row0, row1, raw, cost, out, speckle, FILT,
                                 p.speckleWindowSize, p.speckleRange, wlab, wsiz, wrun, st, &h->launches,
                                 [](void *c) { (*static_cast<Hook *>(c)->m)(); }, &hook, g.spill, h->sw);
    mark();
    return rc;
}

extern "C" int rtdm_bm_set_profiling(rtdm_bm *h, int on)
{
    if (!h) return -RTDM_EINVAL;
    if (!h->ev) h->ev = new std::vector<cudaEvent_t>();
    h->prof = on ? 1 : 0;
    return 0;
}

extern "C" int rtdm_bm_stage_times(rtdm_bm *h, double *ms_sum, int *calls)
{
    if (!h || !ms_sum || !calls) return -RTDM_EINVAL;
    for (int i = 0; i < BM_STAGES; i++) ms_sum[i] = 0.0;
    *calls = 0;
    if (!h->ev) return 0;
    RTDM_CUDA(cudaSetDevice(h->dev));
    std::vector<cudaEvent_t> &ev = *h->ev;
    const size_t per = BM_STAGES + 1;
    for (size_t c = 0; c + per <= ev.size(); c += per) {
        RTDM_CUDA(cudaEventSynchronize(ev[c + per - 1]));
        for (int i = 0; i < BM_STAGES; i++) {
            float ms = 0.f;
            RTDM_CUDA(cudaEventElapsedTime(&ms, ev[c + i], ev[c + i + 1]));
            ms_sum[i] += ms;
        }
        (*calls)++;
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    ev.clear();
    return 0;
}

extern "C" int rtdm_bm_compute_device(rtdm_bm *h, int n, const uint8_t *left, size_t lstep, size_t lframe,
                                      const uint8_t *right, size_t rstep, size_t rframe, int width, int height,
                                      int16_t *disp, size_t dstep, size_t dframe, void *cuda_stream)
{
    if (!h || !left || !right || !disp) { set_error("bm_compute_device: null argument"); return -RTDM_EINVAL; }
    if (dstep % 2 || dframe % 2) { set_error("bm: output steps must be multiples of 2 bytes"); return -RTDM_EINVAL; }
    RTDM_CUDA(cudaSetDevice(h->dev));
    h->launches = 0;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->st;
    PlaneU8 L = {left, lstep, lframe}, R = {right, rstep, rframe};
    PlaneS16 out = {disp, dstep / 2, dframe / 2};
    return bm_pipeline(h, n, L, R, width, height, out, st);
}

extern "C" int rtdm_bm_wait(rtdm_bm *h)
{
    if (!h) return -RTDM_EINVAL;
    RTDM_CUDA(cudaSetDevice(h->dev));
    cudaError_t e0 = cudaStreamSynchronize(h->lane[1]);
    { cudaError_t e = cudaStreamSynchronize(h->st); if (e0 == cudaSuccess) e0 = e; }
    { cudaError_t e = cudaStreamSynchronize(h->lane[0]); if (e0 == cudaSuccess) e0 = e; }
    h->busy[0] = h->busy[1] = 0;
    RTDM_CUDA(e0);
    return 0;
}

extern "C" int rtdm_bm_wait_oldest(rtdm_bm *h)
{
    if (!h) return -RTDM_EINVAL;
    RTDM_CUDA(cudaSetDevice(h->dev));
    const int newest = (int)((h->seq - 1u) & 1u), oldest = newest ^ 1;
    const int set = h->busy[oldest] ? oldest : newest;       // only one in flight: that one
    if (h->busy[set]) { RTDM_CUDA(cudaEventSynchronize(h->done[set])); h->busy[set] = 0; }
    return 0;
}

// enqueues one host batch on staging set `set`; any failure returns at once -- the caller drains the streams
static int bm_enqueue_batch(rtdm_bm *h, int set, int n, const uint8_t *left, size_t lstep, size_t lframe,
                            const uint8_t *right, size_t rstep, size_t rframe, int width, int height,
                            int16_t *disp, size_t dstep, size_t dframe)
{
    // three-stage pipeline over chunks of frames: H2D on lane[0], kernels on the handle's stream, D2H on lane[1],
    // chained by events.  The kernels always see whole chunks in order (no concurrent kernels from different
    // chunks fighting for the SMs); the copies of chunk c+1 / c-1 overlap the kernels of chunk c.
    // the SAD kernel is ~12 % more efficient on 32-frame launches than on 16-frame ones (fuller waves)
    int chunk = bm_chunk_for(h, n, width, height);
    if (h->sw.bm_chunk > 0) chunk = std::max(1, std::min(n, h->sw.bm_chunk));
    const int nchunks = (n + chunk - 1) / chunk;
    if (!h->pev) h->pev = new std::vector<cudaEvent_t>();
    cudaStream_t s_in = h->lane[0], s_out = h->lane[1], s_cmp = h->st;
    uint8_t *sL = set ? h->dL2 : h->dL, *sR = set ? h->dR2 : h->dR;
    int16_t *sD = set ? h->dD2 : h->dD;
    while ((int)h->pev->size() < 4 * nchunks) {
        cudaEvent_t e;
        RTDM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->pev->push_back(e);
    }
    int rc = 0;
    for (int c = 0; c < nchunks && !rc; c++) {
        const int f0 = c * chunk, m = std::min(chunk, n - f0);
        cudaEvent_t ev_in = (*h->pev)[2 * (set * nchunks + c)], ev_done = (*h->pev)[2 * (set * nchunks + c) + 1];
        // one contiguous copy when the caller's frames are laid out like the staging planes (a single frame has no frame step)
        const bool lpacked = lstep == (size_t)width && h->spitch == (size_t)width && (n == 1 || lframe == h->sframe);
        const bool rpacked = rstep == (size_t)width && h->spitch == (size_t)width && (n == 1 || rframe == h->sframe);
        if (lpacked) RTDM_CUDA(cudaMemcpyAsync(sL + f0 * h->sframe, left + f0 * lframe, (size_t)m * h->sframe, cudaMemcpyHostToDevice, s_in));
        if (rpacked) RTDM_CUDA(cudaMemcpyAsync(sR + f0 * h->sframe, right + f0 * rframe, (size_t)m * h->sframe, cudaMemcpyHostToDevice, s_in));
        for (int k = f0; k < f0 + m; k++) {
            if (!lpacked) RTDM_CUDA(cudaMemcpy2DAsync(sL + k * h->sframe, h->spitch, left + k * lframe, lstep, width, height, cudaMemcpyHostToDevice, s_in));
            if (!rpacked) RTDM_CUDA(cudaMemcpy2DAsync(sR + k * h->sframe, h->spitch, right + k * rframe, rstep, width, height, cudaMemcpyHostToDevice, s_in));
        }
        RTDM_CUDA(cudaEventRecord(ev_in, s_in));
        RTDM_CUDA(cudaStreamWaitEvent(s_cmp, ev_in, 0));
        PlaneU8 L = {sL + (size_t)f0 * h->sframe, h->spitch, h->sframe}, R = {sR + (size_t)f0 * h->sframe, h->spitch, h->sframe};
        PlaneS16 out = {sD + (size_t)f0 * h->dframe, h->dpitch, h->dframe};
        const int prof = h->prof; h->prof = 0;          // stage events are for the device entry point only
        rc = bm_pipeline(h, m, L, R, width, height, out, s_cmp, f0);
        h->prof = prof;
        if (rc) break;
        RTDM_CUDA(cudaEventRecord(ev_done, s_cmp));
        RTDM_CUDA(cudaStreamWaitEvent(s_out, ev_done, 0));
        const bool dpacked = dstep == (size_t)width * 2 && h->dpitch == (size_t)width && (n == 1 || dframe == h->dframe * 2);
        if (dpacked) RTDM_CUDA(cudaMemcpyAsync((uint8_t *)disp + f0 * dframe, sD + f0 * h->dframe, (size_t)m * h->dframe * 2, cudaMemcpyDeviceToHost, s_out));
        else for (int k = f0; k < f0 + m; k++)
            RTDM_CUDA(cudaMemcpy2DAsync((uint8_t *)disp + k * dframe, dstep, sD + k * h->dframe, h->dpitch * 2,
                                        (size_t)width * 2, height, cudaMemcpyDeviceToHost, s_out));
    }
    if (rc) return rc;
    RTDM_CUDA(cudaEventRecord(h->done[set], s_out));          // after the last D2H of this call
    h->busy[set] = 1;
    return 0;
}

extern "C" int rtdm_bm_submit_batch(rtdm_bm *h, int n, const uint8_t *left, size_t lstep, size_t lframe,
                                    const uint8_t *right, size_t rstep, size_t rframe, int width, int height,
                                    int16_t *disp, size_t dstep, size_t dframe)
{
    if (!h || !left || !right || !disp) { set_error("bm_compute: null argument"); return -RTDM_EINVAL; }
    if (n < 1 || n > h->maxB || width > h->maxW || height > h->maxH || width < 1 || height < 1) {
        set_error("bm: frame geometry or batch exceeds what the handle was created for");
        return -RTDM_EINVAL;
    }
    RTDM_CUDA(cudaSetDevice(h->dev));
    h->launches = 0;
    // staging set of this call; wait (host side) for the call that used it two submissions ago
    const int set = (int)(h->seq++ & 1u);
    if (h->busy[set]) { RTDM_CUDA(cudaEventSynchronize(h->done[set])); h->busy[set] = 0; }
    const int rc = bm_enqueue_batch(h, set, n, left, lstep, lframe, right, rstep, rframe, width, height, disp, dstep, dframe);
    if (rc) {
        // copies to or from the caller's buffers may already be in flight: never return before they have drained
        const std::string why = rtdm_last_error();
        rtdm_bm_wait(h);
        set_error(why);
    }
    return rc;
}

extern "C" int rtdm_bm_compute_batch(rtdm_bm *h, int n, const uint8_t *left, size_t lstep, size_t lframe,
                                     const uint8_t *right, size_t rstep, size_t rframe, int width, int height,
                                     int16_t *disp, size_t dstep, size_t dframe)
{
    int rc = rtdm_bm_submit_batch(h, n, left, lstep, lframe, right, rstep, rframe, width, height, disp, dstep, dframe);
    if (rc) return rc;
    return rtdm_bm_wait(h);
}

extern "C" int rtdm_bm_compute(rtdm_bm *h, const uint8_t *left, size_t lstep, const uint8_t *right,
                               size_t rstep, int width, int height, int16_t *disp, size_t dstep)
{
    return rtdm_bm_compute_batch(h, 1, left, lstep, 0, right, rstep, 0, width, height, disp, dstep, 0);
}

extern "C" int rtdm_bm_speckle_device(rtdm_bm *h, int n, int16_t *disp, size_t dstep, size_t dframe, int width, int height,
                                      void *cuda_stream)
{
    if (!h || !disp) { set_error("bm_speckle: null argument"); return -RTDM_EINVAL; }
    if (n < 1 || n > h->maxB || width < 1 || height < 1 || (size_t)width * height > (size_t)h->maxW * h->maxH || dstep % 2 || dframe % 2) {
        set_error("bm_speckle: geometry or batch exceeds what the handle was created for");
        return -RTDM_EINVAL;
    }
    RTDM_CUDA(cudaSetDevice(h->dev));
    h->launches = 0;
    if (!(h->p.speckleRange >= 0 && h->p.speckleWindowSize > 0)) return 0;
    const int FILT = (h->p.minDisparity - 1) * 16;
    return launch_speckle(n, width, height, PlaneS16{disp, dstep / 2, dframe / 2}, FILT, h->p.speckleWindowSize, h->p.speckleRange,
                          h->labels, h->sizes, cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->st, &h->launches, h->runlen, h->sw);
}

extern "C" int rtdm_bm_last_launches(const rtdm_bm *h) { return h ? h->launches : 0; }
extern "C" int rtdm_bm_last_kernel(const rtdm_bm *h) { return h ? h->last_kernel : 0; }

extern "C" int rtdm_bm_debug_fetch(rtdm_bm *h, int what, void *dst, size_t dst_bytes)
{
    if (!h || !dst || h->lastW <= 0) { set_error("debug_fetch: nothing computed yet"); return -RTDM_EINVAL; }
    RTDM_CUDA(cudaSetDevice(h->dev));
    RTDM_CUDA(cudaStreamSynchronize(h->st));
    const int W = h->lastW, H = h->lastH;
    if (what == 0 || what == 1) {
        if (dst_bytes < (size_t)W * H) return -RTDM_EINVAL;
        if (what == 1 && h->last_kernel == 4)       // the images only exist in their staged layouts
            RTDM_CUDA(cudaMemcpy2D(dst, W, h->RPs + BmStaged::RPADL, h->rppitch, W, H, cudaMemcpyDeviceToHost));
        else if (what == 0 && h->last_kernel == 4) {
            std::vector<uint32_t> tmp((size_t)W * H);
            RTDM_CUDA(cudaMemcpy2D(tmp.data(), (size_t)W * 4, h->LE + BmStaged::LPADL, h->lepitch * 4, (size_t)W * 4, H, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < tmp.size(); i++) static_cast<uint8_t *>(dst)[i] = (uint8_t)(tmp[i] & 0xFFu);
        } else
            RTDM_CUDA(cudaMemcpy2D(dst, W, what ? h->Rp : h->Lp, h->ppitch, W, H, cudaMemcpyDeviceToHost));
    } else if (what == 2 || what == 3) {
        if (dst_bytes < (size_t)W * H * 2) return -RTDM_EINVAL;
        RTDM_CUDA(cudaMemcpy2D(dst, (size_t)W * 2, what == 3 ? h->cost : h->raw, h->rpitch * 2, (size_t)W * 2, H, cudaMemcpyDeviceToHost));
    } else return -RTDM_EINVAL;
    return 0;
}

// =================================================================================================
// BM, one frame split into row bands over several GPUs of this process (rtdm_b200.h: rtdm_bm_rowband_*)
// =================================================================================================
struct rtdm_bm_rowband {
    rtdm_params p;
    int n, maxW, maxH;
    std::vector<int> dev;
    std::vector<rtdm_bm *> band;      // one matcher per band, speckle filter off, sized for the tallest band + halo
    rtdm_bm *full;                    // devices[0]: the speckle stage of the stitched frame
    int16_t *gather;                  // devices[0]: stitched frame, [maxH][full->dpitch]
    std::vector<cudaEvent_t> done;    // per band: its rows have arrived in `gather`
    int launches;
};

// output rows [y0, y1) of band i and the input rows [i0, i1) it needs; i0 is even so that the x-Sobel's "odd last row"
// rule (OpenCV pairs rows) lands on the same absolute row as in the whole frame
static void rowband_rows(int H, int n, int i, int halo, int *y0, int *y1, int *i0, int *i1)
{
    *y0 = (int)((long long)H * i / n); *y1 = (int)((long long)H * (i + 1) / n);
    *i0 = std::max(0, *y0 - halo) & ~1;
    *i1 = std::min(H, *y1 + halo);
}
static int rowband_halo(const rtdm_params &p)
{
    return p.blockSize / 2 + (p.preFilterType == RTDM_PREFILTER_XSOBEL ? 1 : p.preFilterSize / 2 + 1);
}

extern "C" void rtdm_bm_rowband_destroy(rtdm_bm_rowband *h)
{
    if (!h) return;
    for (rtdm_bm *b : h->band) rtdm_bm_destroy(b);
    for (size_t i = 0; i < h->done.size(); i++) if (h->done[i]) { cudaSetDevice(h->dev[i]); cudaEventDestroy(h->done[i]); }
    if (!h->dev.empty()) cudaSetDevice(h->dev[0]);
    cudaFree(h->gather);
    rtdm_bm_destroy(h->full);
    delete h;
}

extern "C" int rtdm_bm_rowband_create(rtdm_bm_rowband **out, const rtdm_params *p, int max_width, int max_height,
                                      int n_gpus, const int *devices)
{
    if (!out || !p || !devices) { set_error("bm_rowband_create: null argument"); return -RTDM_EINVAL; }
    *out = nullptr;
    if (n_gpus < 1 || n_gpus > 64) { set_error("bm_rowband_create: 1 .. 64 bands"); return -RTDM_EINVAL; }
    if (p->minDisparity > 0) { set_error("bm_rowband: minDisparity > 0 is not supported (row spill crosses bands)"); return -RTDM_EINVAL; }
    int rc = bm_check_params(p);
    if (rc) return rc;
    rtdm_bm_rowband *h = new (std::nothrow) rtdm_bm_rowband();
    if (!h) return -RTDM_ENOMEM;
    h->p = *p; h->n = n_gpus; h->maxW = max_width; h->maxH = max_height; h->full = nullptr; h->gather = nullptr;
    h->dev.assign(devices, devices + n_gpus);
    const int halo = rowband_halo(*p);
    int hb = 1;
    for (int i = 0; i < n_gpus; i++) {
        int y0, y1, i0, i1;
        rowband_rows(max_height, n_gpus, i, halo, &y0, &y1, &i0, &i1);
        hb = std::max(hb, i1 - i0 + 2);            // + 2: smaller frames shift the band edges by a row
    }
    hb = std::min(hb, max_height);
    rtdm_params pb = *p;
    pb.speckleWindowSize = 0;                       // the bands run without the speckle filter
    for (int i = 0; i < n_gpus && !rc; i++) {
        rtdm_bm *b = nullptr;
        rc = rtdm_bm_create(&b, &pb, max_width, hb, 1, devices[i]);
        if (!rc) h->band.push_back(b);
    }
    if (!rc) rc = rtdm_bm_create(&h->full, p, max_width, max_height, 1, devices[0]);
    if (!rc) {
        cudaSetDevice(devices[0]);
        rc = dev_alloc(&h->gather, h->full->dframe);
        for (int i = 0; i < n_gpus && !rc; i++) {          // an event is recorded on a stream of ITS device
            cudaEvent_t e = nullptr;
            cudaSetDevice(devices[i]);
            rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess ? 0 : -RTDM_EIO;
            h->done.push_back(e);
        }
        cudaSetDevice(devices[0]);
        // peer access in both directions between devices[0] and every other device (bands in, input rows out); devices that
        // are not peers still work: the copies are then staged through the host by the driver
        for (int i = 1; i < n_gpus && !rc; i++) {
            if (devices[i] == devices[0]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[0], devices[i]) == cudaSuccess && can) {
                cudaSetDevice(devices[0]);
                if (cudaDeviceEnablePeerAccess(devices[i], 0) != cudaSuccess) cudaGetLastError();   // already enabled is fine
                cudaSetDevice(devices[i]);
                if (cudaDeviceEnablePeerAccess(devices[0], 0) != cudaSuccess) cudaGetLastError();
            }
        }
    }
    if (rc) { rtdm_bm_rowband_destroy(h); return rc; }
    *out = h;
    return 0;
}

extern "C" int rtdm_bm_rowband_set_roi1(rtdm_bm_rowband *h, int x, int y, int w, int hgt)
{
    if (!h) return -RTDM_EINVAL;
    h->p.roi1[0] = x; h->p.roi1[1] = y; h->p.roi1[2] = w; h->p.roi1[3] = hgt;
    return 0;
}
extern "C" int rtdm_bm_rowband_set_roi2(rtdm_bm_rowband *h, int x, int y, int w, int hgt)
{
    if (!h) return -RTDM_EINVAL;
    h->p.roi2[0] = x; h->p.roi2[1] = y; h->p.roi2[2] = w; h->p.roi2[3] = hgt;
    return 0;
}
extern "C" int rtdm_bm_rowband_last_launches(const rtdm_bm_rowband *h) { return h ? h->launches : 0; }

// on_device: left / right / disp live on devices[0]; else in host memory
static int rowband_run(rtdm_bm_rowband *h, bool on_device, const uint8_t *left, size_t lstep, const uint8_t *right, size_t rstep,
                       int W, int H, int16_t *disp, size_t dstep)
{
    if (!h || !left || !right || !disp) { set_error("bm_rowband: null argument"); return -RTDM_EINVAL; }
    if (W < 1 || H < 1 || W > h->maxW || H > h->maxH || dstep % 2) { set_error("bm_rowband: frame geometry exceeds what the handle was created for"); return -RTDM_EINVAL; }
    if (h->p.blockSize >= W || h->p.blockSize >= H) { set_error("bm: blockSize must be smaller than the image"); return -RTDM_EINVAL; }
    const int halo = rowband_halo(h->p), n = h->n;
    rtdm_bm *full = h->full;
    const size_t gp = full->dpitch;                 // elements per row of the stitched frame on devices[0]
    h->launches = 0;
    int rc = 0;
    for (int i = 0; i < n && !rc; i++) {
        int y0, y1, i0, i1;
        rowband_rows(H, n, i, halo, &y0, &y1, &i0, &i1);
        if (y1 <= y0) continue;
        rtdm_bm *b = h->band[i];
        const int hb = i1 - i0;
        if (hb > b->maxH) { set_error("bm_rowband: band taller than the handle's workspace"); rc = -RTDM_EINVAL; break; }
        RTDM_CUDA(cudaSetDevice(h->dev[i]));
        // a ROI-less matcher derives its valid rows from its image height; inside a band that must be the FULL image's rows,
        // so the whole-image rectangle is passed explicitly, shifted into band coordinates
        const bool r1 = h->p.roi1[2] > 0 && h->p.roi1[3] > 0, r2 = h->p.roi2[2] > 0 && h->p.roi2[3] > 0;
        b->p.roi1[0] = r1 ? h->p.roi1[0] : 0; b->p.roi1[1] = (r1 ? h->p.roi1[1] : 0) - i0; b->p.roi1[2] = r1 ? h->p.roi1[2] : W; b->p.roi1[3] = r1 ? h->p.roi1[3] : H;
        b->p.roi2[0] = r2 ? h->p.roi2[0] : 0; b->p.roi2[1] = (r2 ? h->p.roi2[1] : 0) - i0; b->p.roi2[2] = r2 ? h->p.roi2[2] : W; b->p.roi2[3] = r2 ? h->p.roi2[3] : H;
        // input rows [i0, i1) -> this device's staging planes (H2D, or a peer copy from devices[0])
        const cudaMemcpyKind kin = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        if (on_device && h->dev[i] == h->dev[0]) {
            // same device: no copy, the pipeline reads the caller's rows in place
        } else {
            RTDM_CUDA(cudaMemcpy2DAsync(b->dL, b->spitch, left + (size_t)i0 * lstep, lstep, W, hb, kin, b->st));
            RTDM_CUDA(cudaMemcpy2DAsync(b->dR, b->spitch, right + (size_t)i0 * rstep, rstep, W, hb, kin, b->st));
        }
        const bool inplace = on_device && h->dev[i] == h->dev[0];
        PlaneU8 L = {inplace ? left + (size_t)i0 * lstep : b->dL, inplace ? lstep : b->spitch, 0};
        PlaneU8 R = {inplace ? right + (size_t)i0 * rstep : b->dR, inplace ? rstep : b->spitch, 0};
        PlaneS16 out = {b->dD, b->dpitch, 0};
        b->launches = 0;
        rc = bm_pipeline(b, 1, L, R, W, hb, out, b->st);
        if (rc) break;
        h->launches += b->launches;
        // rows [y0, y1) of the band -> the stitched frame on devices[0]
        const int16_t *src = b->dD + (size_t)(y0 - i0) * b->dpitch;
        int16_t *dst = h->gather + (size_t)y0 * gp;
        if (b->dpitch == gp)
            RTDM_CUDA(cudaMemcpyPeerAsync(dst, h->dev[0], src, h->dev[i], (size_t)(y1 - y0) * gp * 2, b->st));
        else
            RTDM_CUDA(cudaMemcpy2DAsync(dst, gp * 2, src, b->dpitch * 2, (size_t)W * 2, y1 - y0, cudaMemcpyDeviceToDevice, b->st));
        RTDM_CUDA(cudaEventRecord(h->done[i], b->st));
    }
    // devices[0]: wait for the bands, speckle filter on the stitched frame, result to the caller
    cudaSetDevice(h->dev[0]);
    if (!rc) {
        for (int i = 0; i < n; i++) {
            int y0, y1, i0, i1;
            rowband_rows(H, n, i, halo, &y0, &y1, &i0, &i1);
            if (y1 > y0) RTDM_CUDA(cudaStreamWaitEvent(full->st, h->done[i], 0));
        }
        if (h->p.speckleRange >= 0 && h->p.speckleWindowSize > 0) {
            full->launches = 0;
            rc = launch_speckle(1, W, H, PlaneS16{h->gather, gp, 0}, (h->p.minDisparity - 1) * 16, h->p.speckleWindowSize, h->p.speckleRange,
                                full->labels, full->sizes, full->st, &full->launches, full->runlen, full->sw);
            h->launches += full->launches;
        }
        if (!rc)
            RTDM_CUDA(cudaMemcpy2DAsync(disp, dstep, h->gather, gp * 2, (size_t)W * 2, H,
                                        on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, full->st));
    }
    // drain every stream that was touched (also after an error: copies from the caller's buffers may be in flight)
    cudaError_t e0 = cudaSuccess;
    for (int i = 0; i < n; i++) {
        cudaSetDevice(h->dev[i]);
        const cudaError_t e = cudaStreamSynchronize(h->band[i]->st);
        if (e0 == cudaSuccess) e0 = e;
    }
    cudaSetDevice(h->dev[0]);
    { const cudaError_t e = cudaStreamSynchronize(full->st); if (e0 == cudaSuccess) e0 = e; }
    if (rc) return rc;
    RTDM_CUDA(e0);
    return 0;
}

extern "C" int rtdm_bm_rowband_compute(rtdm_bm_rowband *h, const uint8_t *left, size_t lstep, const uint8_t *right, size_t rstep,
                                       int width, int height, int16_t *disp, size_t dstep)
{
    return rowband_run(h, false, left, lstep, right, rstep, width, height, disp, dstep);
}

extern "C" int rtdm_bm_rowband_compute_device(rtdm_bm_rowband *h, const uint8_t *left, size_t lstep, const uint8_t *right, size_t rstep,
                                              int width, int height, int16_t *disp, size_t dstep)
{
    return rowband_run(h, true, left, lstep, right, rstep, width, height, disp, dstep);
}

// =================================================================================================
// SGBM
// =================================================================================================
struct rtdm_sgbm {
    rtdm_params p;
    Switches sw;                 // development switches as they were when the handle was created
    int maxW, maxH, maxB, dev;
    int volB;                    // frames the cost volumes are sized for (sub-batches of the call)
    cudaStream_t st;
    uint8_t *planes;  size_t frame_planes;
    uint16_t *C, *S;  size_t frame_vol;          // elements per frame
    int16_t *raw;     size_t rpitch, rframe;     // WTA output before median (elements)
    int32_t *labels, *sizes, *runlen;
    uint8_t *dL, *dR; size_t spitch, sframe;
    int16_t *dD;      size_t dpitch, dframe;
    uint8_t *dL2, *dR2; int16_t *dD2;            // staging set 1 (rtdm_sgbm_submit_batch alternates)
    cudaStream_t lane[2];                        // H2D / D2H streams of the streaming host API
    cudaEvent_t ev_in[2], ev_cmp[2], done[2];    // per staging set: inputs arrived, kernels finished, outputs delivered
    int busy[2];
    unsigned seq;
    int launches;
    int prof;
    std::vector<cudaEvent_t> *ev;     // 3 events per profiled sub-batch: start, after matching, after post-filters
    int *err_host, *err_dev;          // host-mapped flag raised by sgbm_vpass_kernel when a neighbour's data never arrives
};

// the whole-height pass gives up waiting instead of hanging; whoever next calls or waits on the handle learns about it
static int sgbm_check_exchange(rtdm_sgbm *h)
{
    if (h->err_host && *reinterpret_cast<volatile int *>(h->err_host)) {
        *reinterpret_cast<volatile int *>(h->err_host) = 0;
        set_error("sgbm: the cluster pass's neighbour exchange timed out; the maps of the last call are invalid");
        return -RTDM_EIO;
    }
    return 0;
}

static int sgbm_check_params(const rtdm_params *p)
{
    if (p->numDisparities <= 0 || p->numDisparities % 16 != 0) { set_error("sgbm: numDisparities must be positive and divisible by 16"); return -RTDM_EINVAL; }
    if (p->blockSize < 1 || p->blockSize % 2 == 0) { set_error("sgbm: blockSize must be odd and >= 1"); return -RTDM_EINVAL; }
    if (p->numDisparities > 256) { set_error("sgbm: numDisparities > 256 is not supported"); return -RTDM_EINVAL; }
    if (p->blockSize > 11) { set_error("sgbm: blockSize > 11 is not supported (16-bit cost domain, SURVEY.md App. B.5)"); return -RTDM_EINVAL; }
    if (p->P1 > 16000 || p->P2 > 16000) { set_error("sgbm: P1/P2 above 16000 leave the 16-bit cost domain"); return -RTDM_EINVAL; }
    if (p->mode != RTDM_SGBM_MODE_SGBM && p->mode != RTDM_SGBM_MODE_HH) { set_error("sgbm: mode must be MODE_SGBM (0) or MODE_HH (1)"); return -RTDM_EINVAL; }
    return 0;
}

static SgbmGeom sgbm_geom(const rtdm_params &p, const Switches &sw, int W, int H)
{
    SgbmGeom g;
    g.sw = sw;
    g.W = W; g.H = H; g.D = p.numDisparities; g.minD = p.minDisparity; g.bs = p.blockSize;
    g.P1 = p.P1 > 0 ? p.P1 : 2; g.P2 = std::max(p.P2 > 0 ? p.P2 : 5, g.P1 + 1);
    g.uniq = p.uniquenessRatio >= 0 ? p.uniquenessRatio : 10;
    g.d12 = p.disp12MaxDiff > 0 ? p.disp12MaxDiff : 1;
    g.ftzero = std::max(p.preFilterCap, 15) | 1;
    g.mode = p.mode;
    const int maxD = g.minD + g.D;
    g.minX1 = std::max(maxD, 0); g.maxX1 = W + std::min(g.minD, 0); g.W1 = g.maxX1 - g.minX1;
    return g;
}

extern "C" void rtdm_sgbm_destroy(rtdm_sgbm *h)
{
    if (!h) return;
    cudaSetDevice(h->dev);
    cudaFree(h->planes); cudaFree(h->C); cudaFree(h->S); cudaFree(h->raw); cudaFree(h->labels); cudaFree(h->sizes); cudaFree(h->runlen);
    cudaFree(h->dL); cudaFree(h->dR); cudaFree(h->dD);
    cudaFree(h->dL2); cudaFree(h->dR2); cudaFree(h->dD2);
    for (int i = 0; i < 2; i++) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_cmp[i]) cudaEventDestroy(h->ev_cmp[i]);
        if (h->done[i]) cudaEventDestroy(h->done[i]);
        if (h->lane[i]) cudaStreamDestroy(h->lane[i]);
    }
    if (h->ev) { for (cudaEvent_t e : *h->ev) cudaEventDestroy(e); delete h->ev; }
    if (h->st) cudaStreamDestroy(h->st);
    if (h->err_host) cudaFreeHost(h->err_host);
    delete h;
}

extern "C" int rtdm_sgbm_create(rtdm_sgbm **out, const rtdm_params *p, int max_width, int max_height,
                                int max_batch, int device)
{
    if (!out || !p) { set_error("sgbm_create: null argument"); return -RTDM_EINVAL; }
    *out = nullptr;
    int rc = sgbm_check_params(p);
    if (rc) return rc;
    if (max_width < 1 || max_height < 1 || max_batch < 1 || max_width > 8000 || max_height > 65535) {
        set_error("sgbm_create: bad maximum geometry (width <= 8000, height <= 65535, batch >= 1)");
        return -RTDM_EINVAL;
    }
    rc = check_device(device);
    if (rc) return rc;
    RTDM_CUDA(cudaSetDevice(device));
    rtdm_sgbm *h = new (std::nothrow) rtdm_sgbm();      // value-initialised: every member zero, switches at their defaults
    if (!h) return -RTDM_ENOMEM;
    h->p = *p; h->maxW = max_width; h->maxH = max_height; h->maxB = max_batch; h->dev = device;
    h->sw = read_switches();
    SgbmGeom g = sgbm_geom(*p, h->sw, max_width, max_height);
    size_t pl = 0, vol = 0;
    sgbm_work_bytes(g, &pl, &vol);
    h->frame_planes = pl; h->frame_vol = vol;
    // the two cost volumes are the big consumers (2 x 2 bytes x H x W1 x D per frame): keep at most
    // ~40 GB of them resident (of 180 GB) and run larger calls as sub-batches
    const size_t per_frame = std::max<size_t>(1, vol * 4);
    h->volB = (int)std::max<size_t>(1, std::min<size_t>((size_t)max_batch, (size_t)40e9 / per_frame));
    const size_t B = (size_t)max_batch, VB = (size_t)h->volB;
    h->rpitch = align_up((size_t)max_width, 8); h->rframe = h->rpitch * max_height;
    h->spitch = align_up((size_t)max_width, 64); h->sframe = h->spitch * max_height;
    h->dpitch = h->rpitch; h->dframe = h->rframe;
    rc = (int)cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) == cudaSuccess ? 0 : -RTDM_EIO;
    if (!rc && cudaHostAlloc((void **)&h->err_host, sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
        *h->err_host = 0;
        if (cudaHostGetDevicePointer((void **)&h->err_dev, h->err_host, 0) != cudaSuccess) h->err_dev = nullptr;   // no flag: tiled sweeps only
    } else cudaGetLastError();
    if (!rc) rc = dev_alloc(&h->planes, pl * VB);
    if (!rc) rc = dev_alloc(&h->C, vol * VB + 64);
    if (!rc) rc = dev_alloc(&h->S, vol * VB + 64);
    if (!rc) rc = dev_alloc(&h->raw, h->rframe * VB);
    if (!rc) rc = dev_alloc(&h->labels, (size_t)max_width * max_height * VB);
    if (!rc) rc = dev_alloc(&h->sizes, (size_t)max_width * max_height * VB);
    if (!rc) rc = dev_alloc(&h->runlen, (size_t)max_width * max_height * VB);
    if (!rc) rc = dev_alloc(&h->dL, h->sframe * B);
    if (!rc) rc = dev_alloc(&h->dR, h->sframe * B);
    if (!rc) rc = dev_alloc(&h->dD, h->dframe * B);
    if (!rc) rc = dev_alloc(&h->dL2, h->sframe * B);
    if (!rc) rc = dev_alloc(&h->dR2, h->sframe * B);
    if (!rc) rc = dev_alloc(&h->dD2, h->dframe * B);
    for (int i = 0; i < 2 && !rc; i++) {
        if (cudaStreamCreateWithFlags(&h->lane[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_cmp[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->done[i], cudaEventDisableTiming) != cudaSuccess) rc = -RTDM_EIO;
    }
    if (rc) { rtdm_sgbm_destroy(h); return rc; }
    *out = h;
    return 0;
}

static int sgbm_pipeline(rtdm_sgbm *h, int n, PlaneU8 L, PlaneU8 R, int W, int H, PlaneS16 out, cudaStream_t st)
{
    if (W > h->maxW || H > h->maxH || n > h->maxB || W < 1 || H < 1 || n < 1) {
        set_error("sgbm: frame geometry or batch exceeds what the handle was created for");
        return -RTDM_EINVAL;
    }
    { const int rc = sgbm_check_exchange(h); if (rc) return rc; }
    SgbmGeom g = sgbm_geom(h->p, h->sw, W, H);
    const int INVS = (g.minD - 1) * 16;
    size_t pl = 0, vol = 0;
    sgbm_work_bytes(g, &pl, &vol);
    // sub-batch size.  Calls that take the tiled row sweeps (one 1024-thread CTA per SM, frames x column tiles CTAs per launch):
    // among the sizes the volumes allow, the one that fills whole waves best (720p x 128: 37 frames x 24 tiles = 6.0 waves
    // of 148 SMs; 32 frames would be 5.2 -> 6 waves).  Calls that take the whole-height passes: see below.
    int chunk = h->volB;
    // whole-height pass (batches): one cluster per frame, `q` frames at a time -> sub-batches are multiples of q
    const int q = n >= 2 ? sgbm_vpass_frames_in_flight(g, std::min(n, chunk)) : 0;      // (a single frame never takes the pass: skip the occupancy query)
    if (q > 0) {
        if (n <= chunk) chunk = n;                               // one sub-batch; the pass itself loops over rounds of q frames
        else {
            if (chunk > q) chunk -= chunk % q;
            const int rounds = cdiv(n, chunk);                   // spread the frames evenly over the sub-batches ...
            const int even = cdiv(cdiv(n, rounds), q) * q;       // ... in multiples of q
            if (even <= chunk) chunk = even;
        }
    } else if (n > 1) {
        const int ctas = sgbm_sweep_ctas_per_frame(g);
        int nsm = 0;
        if (ctas > 0 && cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->dev) == cudaSuccess && nsm > 0) {
            double best = 0.0;
            for (int m = 1; m <= std::min(h->volB, n); m++) {
                const long long c = (long long)m * ctas, waves = (c + nsm - 1) / nsm;
                const double eff = (double)c / (double)(waves * nsm);
                if (eff >= best - 1e-9 || (m >= 8 && eff >= best - 0.02)) { if (eff > best) best = eff; chunk = m; }
            }
        }
    }
    for (int f0 = 0; f0 < n; f0 += chunk) {
        const int m = std::min(chunk, n - f0);
        PlaneU8 l = {L.p + (size_t)f0 * L.frame, L.pitch, L.frame}, r = {R.p + (size_t)f0 * R.frame, R.pitch, R.frame};
        PlaneS16 o = {out.p + (size_t)f0 * out.frame, out.pitch, out.frame};
        PlaneS16 raw = {h->raw, h->rpitch, h->rframe};
        auto mark = [&]() {
            if (!h->prof) return;
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            cudaEventRecord(e, st);
            h->ev->push_back(e);
        };
        mark();
        if (g.W1 > 0) {
            SgbmWork w;
            memset(&w, 0, sizeof w);
            w.planes = h->planes; w.C = (int16_t *)h->C; w.S = (int16_t *)h->S; w.frame_planes = pl; w.frame_vol = vol; w.err = h->err_dev;
            int rc = launch_sgbm(g, m, l, r, raw, w, st, &h->launches);
            if (rc) return rc;
        } else {
            // no computable column: the whole map is invalid (validate_mask with an empty row range fills it)
            int rc = launch_validate_mask(m, W, H, g.minD, g.D, -1, 0, 0, 0, 0, 0, 0, raw, raw, raw, st, &h->launches);
            if (rc) return rc;
        }
        mark();
        int rc = launch_median3(m, W, H, raw, o, st, &h->launches);
        if (rc) return rc;
        if (h->p.speckleWindowSize > 0) {
            rc = launch_speckle(m, W, H, o, INVS, h->p.speckleWindowSize, 16 * h->p.speckleRange, h->labels, h->sizes, st, &h->launches, h->runlen, h->sw);
            if (rc) return rc;
        }
        mark();
    }
    return 0;
}

extern "C" int rtdm_sgbm_set_profiling(rtdm_sgbm *h, int on)
{
    if (!h) return -RTDM_EINVAL;
    if (!h->ev) h->ev = new std::vector<cudaEvent_t>();
    h->prof = on ? 1 : 0;
    return 0;
}

extern "C" int rtdm_sgbm_stage_times(rtdm_sgbm *h, double *ms_sum, int *calls)
{
    if (!h || !ms_sum || !calls) return -RTDM_EINVAL;
    ms_sum[0] = ms_sum[1] = 0.0;
    *calls = 0;
    if (!h->ev) return 0;
    RTDM_CUDA(cudaSetDevice(h->dev));
    std::vector<cudaEvent_t> &ev = *h->ev;
    for (size_t c = 0; c + 3 <= ev.size(); c += 3) {
        RTDM_CUDA(cudaEventSynchronize(ev[c + 2]));
        for (int i = 0; i < 2; i++) {
            float ms = 0.f;
            RTDM_CUDA(cudaEventElapsedTime(&ms, ev[c + i], ev[c + i + 1]));
            ms_sum[i] += ms;
        }
        (*calls)++;
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    ev.clear();
    return 0;
}

extern "C" int rtdm_sgbm_compute_device(rtdm_sgbm *h, int n, const uint8_t *left, size_t lstep, size_t lframe,
                                        const uint8_t *right, size_t rstep, size_t rframe, int width, int height,
                                        int16_t *disp, size_t dstep, size_t dframe, void *cuda_stream)
{
    if (!h || !left || !right || !disp) { set_error("sgbm_compute_device: null argument"); return -RTDM_EINVAL; }
    if (dstep % 2 || dframe % 2) { set_error("sgbm: output steps must be multiples of 2 bytes"); return -RTDM_EINVAL; }
    RTDM_CUDA(cudaSetDevice(h->dev));
    h->launches = 0;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->st;
    return sgbm_pipeline(h, n, PlaneU8{left, lstep, lframe}, PlaneU8{right, rstep, rframe}, width, height,
                         PlaneS16{disp, dstep / 2, dframe / 2}, st);
}

extern "C" int rtdm_sgbm_wait(rtdm_sgbm *h)
{
    if (!h) return -RTDM_EINVAL;
    RTDM_CUDA(cudaSetDevice(h->dev));
    cudaError_t e0 = cudaStreamSynchronize(h->lane[1]);
    { cudaError_t e = cudaStreamSynchronize(h->st); if (e0 == cudaSuccess) e0 = e; }
    { cudaError_t e = cudaStreamSynchronize(h->lane[0]); if (e0 == cudaSuccess) e0 = e; }
    h->busy[0] = h->busy[1] = 0;
    RTDM_CUDA(e0);
    return sgbm_check_exchange(h);
}

extern "C" int rtdm_sgbm_wait_oldest(rtdm_sgbm *h)
{
    if (!h) return -RTDM_EINVAL;
    RTDM_CUDA(cudaSetDevice(h->dev));
    const int newest = (int)((h->seq - 1u) & 1u), oldest = newest ^ 1;
    const int set = h->busy[oldest] ? oldest : newest;       // only one in flight: that one
    if (h->busy[set]) { RTDM_CUDA(cudaEventSynchronize(h->done[set])); h->busy[set] = 0; }
    return sgbm_check_exchange(h);
}

// H2D on lane[0], kernels on the handle's stream, D2H on lane[1], chained by events; two staging sets, so the copies of
// batch k+1 / k-1 run under the kernels of batch k (the cost volumes and scratch are used by one batch at a time: the
// kernels of successive batches are ordered on the one compute stream)
static int sgbm_enqueue_batch(rtdm_sgbm *h, int set, int n, const uint8_t *left, size_t lstep, size_t lframe,
                              const uint8_t *right, size_t rstep, size_t rframe, int width, int height,
                              int16_t *disp, size_t dstep, size_t dframe)
{
    uint8_t *sL = set ? h->dL2 : h->dL, *sR = set ? h->dR2 : h->dR;
    int16_t *sD = set ? h->dD2 : h->dD;
    cudaStream_t s_in = h->lane[0], s_out = h->lane[1], st = h->st;
    for (int k = 0; k < n; k++) {
        RTDM_CUDA(cudaMemcpy2DAsync(sL + k * h->sframe, h->spitch, left + k * lframe, lstep, width, height, cudaMemcpyHostToDevice, s_in));
        RTDM_CUDA(cudaMemcpy2DAsync(sR + k * h->sframe, h->spitch, right + k * rframe, rstep, width, height, cudaMemcpyHostToDevice, s_in));
    }
    RTDM_CUDA(cudaEventRecord(h->ev_in[set], s_in));
    RTDM_CUDA(cudaStreamWaitEvent(st, h->ev_in[set], 0));
    int rc = sgbm_pipeline(h, n, PlaneU8{sL, h->spitch, h->sframe}, PlaneU8{sR, h->spitch, h->sframe}, width, height,
                           PlaneS16{sD, h->dpitch, h->dframe}, st);
    if (rc) return rc;
    RTDM_CUDA(cudaEventRecord(h->ev_cmp[set], st));
    RTDM_CUDA(cudaStreamWaitEvent(s_out, h->ev_cmp[set], 0));
    for (int k = 0; k < n; k++)
        RTDM_CUDA(cudaMemcpy2DAsync((uint8_t *)disp + k * dframe, dstep, sD + k * h->dframe, h->dpitch * 2,
                                    (size_t)width * 2, height, cudaMemcpyDeviceToHost, s_out));
    RTDM_CUDA(cudaEventRecord(h->done[set], s_out));
    h->busy[set] = 1;
    return 0;
}

extern "C" int rtdm_sgbm_submit_batch(rtdm_sgbm *h, int n, const uint8_t *left, size_t lstep, size_t lframe,
                                      const uint8_t *right, size_t rstep, size_t rframe, int width, int height,
                                      int16_t *disp, size_t dstep, size_t dframe)
{
    if (!h || !left || !right || !disp) { set_error("sgbm_compute: null argument"); return -RTDM_EINVAL; }
    if (n < 1 || n > h->maxB || width > h->maxW || height > h->maxH || width < 1 || height < 1) {
        set_error("sgbm: frame geometry or batch exceeds what the handle was created for");
        return -RTDM_EINVAL;
    }
    RTDM_CUDA(cudaSetDevice(h->dev));
    h->launches = 0;
    const int set = (int)(h->seq++ & 1u);
    if (h->busy[set]) { RTDM_CUDA(cudaEventSynchronize(h->done[set])); h->busy[set] = 0; }
    const int rc = sgbm_enqueue_batch(h, set, n, left, lstep, lframe, right, rstep, rframe, width, height, disp, dstep, dframe);
    if (rc) {
        const std::string why = rtdm_last_error();
        rtdm_sgbm_wait(h);
        set_error(why);
    }
    return rc;
}

extern "C" int rtdm_sgbm_compute_batch(rtdm_sgbm *h, int n, const uint8_t *left, size_t lstep, size_t lframe,
                                       const uint8_t *right, size_t rstep, size_t rframe, int width, int height,
                                       int16_t *disp, size_t dstep, size_t dframe)
{
    int rc = rtdm_sgbm_submit_batch(h, n, left, lstep, lframe, right, rstep, rframe, width, height, disp, dstep, dframe);
    if (rc) return rc;
    return rtdm_sgbm_wait(h);
}

extern "C" int rtdm_sgbm_compute(rtdm_sgbm *h, const uint8_t *left, size_t lstep, const uint8_t *right,
                                 size_t rstep, int width, int height, int16_t *disp, size_t dstep)
{
    return rtdm_sgbm_compute_batch(h, 1, left, lstep, 0, right, rstep, 0, width, height, disp, dstep, 0);
}

extern "C" int rtdm_sgbm_last_launches(const rtdm_sgbm *h) { return h ? h->launches : 0; }

extern "C" int rtdm_sgbm_batch_quantum(rtdm_sgbm *h, int width, int height)
{
    if (!h || width < 1 || height < 1 || width > h->maxW || height > h->maxH) { set_error("sgbm_batch_quantum: bad argument"); return -RTDM_EINVAL; }
    RTDM_CUDA(cudaSetDevice(h->dev));
    const int q = sgbm_vpass_frames_in_flight(sgbm_geom(h->p, h->sw, width, height), h->volB);
    return q > 0 ? q : 1;
}

// =================================================================================================
// morphological filter
// =================================================================================================
struct rtdm_morph {
    int W, H, maxB, dev;
    cudaStream_t st;
    uint8_t *hin, *hout;         // pinned host frame buffers (video_in / video_out)
    uint8_t *d0, *d1, *d2, *d3;  // device scratch planes, maxB frames each, tightly packed
    int *flags;                  // per-frame "not a binary mask" flags
    MorphSE se;
    int launches;
};

extern "C" void rtdm_morph_destroy(rtdm_morph *h)
{
    if (!h) return;
    cudaSetDevice(h->dev);
    if (h->hin) cudaFreeHost(h->hin);
    if (h->hout) cudaFreeHost(h->hout);
    cudaFree(h->d0); cudaFree(h->d1); cudaFree(h->d2); cudaFree(h->d3); cudaFree(h->flags);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

extern "C" int rtdm_morph_create(rtdm_morph **out, int width, int height, int bpp, int max_batch, int device)
{
    if (!out) return -RTDM_EINVAL;
    *out = nullptr;
    if (bpp != 8 || width < 1 || height < 1 || max_batch < 1) { set_error("morph_create: bpp must be 8 and geometry positive"); return -RTDM_EINVAL; }
    int rc = check_device(device);
    if (rc) return rc;
    RTDM_CUDA(cudaSetDevice(device));
    rtdm_morph *h = new (std::nothrow) rtdm_morph();
    if (!h) return -RTDM_ENOMEM;
    memset(h, 0, sizeof *h);
    h->W = width; h->H = height; h->maxB = max_batch; h->dev = device;
    make_ellipse(10, 10, &h->se);      // MORPH_FILTER_DX x MORPH_FILTER_DY (include/filter/mf-sw.h:11-12)
    const size_t fb = (size_t)width * height;
    {
        // the filter's kernels are tiny next to a matcher's: on a high-priority stream their CTAs are placed as soon as an
        // SM has room instead of queueing behind thousands of SAD CTAs when both plugins stream side by side (measured:
        // the end-to-end rate of matcher + filter no longer flips between 11 and 13.7 kfps from run to run)
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        rc = cudaStreamCreateWithPriority(&h->st, cudaStreamNonBlocking, hi) == cudaSuccess ? 0 : -RTDM_EIO;
    }
    if (!rc && cudaHostAlloc((void **)&h->hin, fb, cudaHostAllocDefault) != cudaSuccess) rc = -RTDM_ENOMEM;
    if (!rc && cudaHostAlloc((void **)&h->hout, fb, cudaHostAllocDefault) != cudaSuccess) rc = -RTDM_ENOMEM;
    if (!rc) rc = dev_alloc(&h->d0, fb * max_batch);
    if (!rc) rc = dev_alloc(&h->d1, fb * max_batch);
    if (!rc) rc = dev_alloc(&h->d2, fb * max_batch);
    if (!rc) rc = dev_alloc(&h->d3, fb * max_batch);
    if (!rc) rc = dev_alloc(&h->flags, (size_t)max_batch);
    if (rc) { rtdm_morph_destroy(h); return rc; }
    *out = h;
    return 0;
}

extern "C" uint8_t *rtdm_morph_in_buffer(rtdm_morph *h) { return h ? h->hin : nullptr; }
extern "C" uint8_t *rtdm_morph_out_buffer(rtdm_morph *h) { return h ? h->hout : nullptr; }
extern "C" int rtdm_morph_last_launches(const rtdm_morph *h) { return h ? h->launches : 0; }

// erode, dilate, dilate, erode (open then close); src and dst may alias
static int morph_pipeline(rtdm_morph *h, int n, const uint8_t *src, uint8_t *dst, cudaStream_t st)
{
    const size_t W = h->W, fb = (size_t)h->W * h->H;
    return launch_morph_openclose(n, h->W, h->H, PlaneU8{src, W, fb}, PlaneU8W{dst, W, fb}, PlaneU8W{h->d0, W, fb},
                                  PlaneU8W{h->d1, W, fb}, PlaneU8W{h->d2, W, fb}, h->flags, h->se, st, &h->launches);
}

extern "C" int rtdm_morph_run(rtdm_morph *h, const uint8_t *in, uint8_t *out)
{
    if (!h || !in || !out) { set_error("morph_run: null argument"); return -RTDM_EINVAL; }
    RTDM_CUDA(cudaSetDevice(h->dev));
    h->launches = 0;
    const size_t fb = (size_t)h->W * h->H;
    RTDM_CUDA(cudaMemcpyAsync(h->d3, in, fb, cudaMemcpyHostToDevice, h->st));
    int rc = morph_pipeline(h, 1, h->d3, h->d2, h->st);
    if (rc) return rc;
    RTDM_CUDA(cudaMemcpyAsync(out, h->d2, fb, cudaMemcpyDeviceToHost, h->st));
    RTDM_CUDA(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int rtdm_morph_sync(rtdm_morph *h)
{
    if (!h) return -RTDM_EINVAL;
    RTDM_CUDA(cudaSetDevice(h->dev));
    RTDM_CUDA(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int rtdm_morph_run_batch_async(rtdm_morph *h, int n, const uint8_t *in, uint8_t *out)
{
    if (!h || !in || !out) { set_error("morph_run_batch_async: null argument"); return -RTDM_EINVAL; }
    if (n < 1 || n > h->maxB) { set_error("morph: batch exceeds what the handle was created for"); return -RTDM_EINVAL; }
    RTDM_CUDA(cudaSetDevice(h->dev));
    h->launches = 0;
    const size_t fb = (size_t)h->W * h->H;
    RTDM_CUDA(cudaMemcpyAsync(h->d3, in, fb * n, cudaMemcpyHostToDevice, h->st));
    int rc = morph_pipeline(h, n, h->d3, h->d2, h->st);
    if (rc) return rc;
    RTDM_CUDA(cudaMemcpyAsync(out, h->d2, fb * n, cudaMemcpyDeviceToHost, h->st));
    return 0;
}

extern "C" int rtdm_morph_run_batch(rtdm_morph *h, int n, const uint8_t *in, uint8_t *out)
{
    if (!h || !in || !out) { set_error("morph_run_batch: null argument"); return -RTDM_EINVAL; }
    if (n < 1 || n > h->maxB) { set_error("morph: batch exceeds what the handle was created for"); return -RTDM_EINVAL; }
    RTDM_CUDA(cudaSetDevice(h->dev));
    h->launches = 0;
    const size_t fb = (size_t)h->W * h->H;
    RTDM_CUDA(cudaMemcpyAsync(h->d3, in, fb * n, cudaMemcpyHostToDevice, h->st));
    int rc = morph_pipeline(h, n, h->d3, h->d2, h->st);
    if (rc) return rc;
    RTDM_CUDA(cudaMemcpyAsync(out, h->d2, fb * n, cudaMemcpyDeviceToHost, h->st));
    RTDM_CUDA(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int rtdm_morph_run_device(rtdm_morph *h, int n, const uint8_t *in, uint8_t *out, void *cuda_stream)
{
    if (!h || !in || !out) { set_error("morph_run_device: null argument"); return -RTDM_EINVAL; }
    if (n < 1 || n > h->maxB) { set_error("morph: batch exceeds what the handle was created for"); return -RTDM_EINVAL; }
    RTDM_CUDA(cudaSetDevice(h->dev));
    h->launches = 0;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->st;
    return morph_pipeline(h, n, in, out, st);
}

// =================================================================================================
// stand-alone stages (host pointers)
// =================================================================================================
extern "C" int rtdm_filter_speckles(int16_t *img, size_t step, int width, int height, int newVal,
                                    int maxSpeckleSize, int maxDiff, int device)
{
    if (!img || width < 1 || height < 1 || step % 2) return -RTDM_EINVAL;
    int rc = check_device(device);
    if (rc) return rc;
    RTDM_CUDA(cudaSetDevice(device));
    int16_t *d = nullptr; int32_t *lab = nullptr, *siz = nullptr, *rl = nullptr;
    const size_t N = (size_t)width * height;
    rc = dev_alloc(&d, N);
    if (!rc) rc = dev_alloc(&lab, N);
    if (!rc) rc = dev_alloc(&siz, N);
    if (!rc) rc = dev_alloc(&rl, N);
    if (!rc) {
        cudaError_t e = cudaMemcpy2D(d, (size_t)width * 2, img, step, (size_t)width * 2, height, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = cuda_fail(e, "memcpy2d", __FILE__, __LINE__);
    }
    if (!rc) rc = launch_speckle(1, width, height, PlaneS16{d, (size_t)width, N}, newVal, maxSpeckleSize, maxDiff, lab, siz, 0, nullptr, rl, read_switches());
    if (!rc) {
        cudaError_t e = cudaMemcpy2D(img, step, d, (size_t)width * 2, (size_t)width * 2, height, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "memcpy2d", __FILE__, __LINE__);
    }
    cudaFree(d); cudaFree(lab); cudaFree(siz); cudaFree(rl);
    return rc;
}

extern "C" int rtdm_median3_s16(const int16_t *src, size_t sstep, int16_t *dst, size_t dstep, int width,
                                int height, int device)
{
    if (!src || !dst || width < 1 || height < 1) return -RTDM_EINVAL;
    int rc = check_device(device);
    if (rc) return rc;
    RTDM_CUDA(cudaSetDevice(device));
    int16_t *a = nullptr, *b = nullptr;
    const size_t N = (size_t)width * height;
    rc = dev_alloc(&a, N);
    if (!rc) rc = dev_alloc(&b, N);
    if (!rc) { cudaError_t e = cudaMemcpy2D(a, (size_t)width * 2, src, sstep, (size_t)width * 2, height, cudaMemcpyHostToDevice); if (e != cudaSuccess) rc = cuda_fail(e, "memcpy2d", __FILE__, __LINE__); }
    if (!rc) rc = launch_median3(1, width, height, PlaneS16{a, (size_t)width, N}, PlaneS16{b, (size_t)width, N}, 0, nullptr);
    if (!rc) { cudaError_t e = cudaMemcpy2D(dst, dstep, b, (size_t)width * 2, (size_t)width * 2, height, cudaMemcpyDeviceToHost); if (e != cudaSuccess) rc = cuda_fail(e, "memcpy2d", __FILE__, __LINE__); }
    cudaFree(a); cudaFree(b);
    return rc;
}

extern "C" int rtdm_morph_op(const uint8_t *src, size_t sstep, uint8_t *dst, size_t dstep, int width,
                             int height, int kw, int kh, int op, int device)
{
    if (!src || !dst || width < 1 || height < 1 || kw < 1 || kh < 1 || kw > 31 || kh > 31 || (op != 0 && op != 1)) return -RTDM_EINVAL;
    int rc = check_device(device);
    if (rc) return rc;
    RTDM_CUDA(cudaSetDevice(device));
    uint8_t *a = nullptr, *b = nullptr;
    const size_t N = (size_t)width * height;
    rc = dev_alloc(&a, N);
    if (!rc) rc = dev_alloc(&b, N);
    MorphSE se; make_ellipse(kw, kh, &se);
    if (!rc) { cudaError_t e = cudaMemcpy2D(a, width, src, sstep, width, height, cudaMemcpyHostToDevice); if (e != cudaSuccess) rc = cuda_fail(e, "memcpy2d", __FILE__, __LINE__); }
    if (!rc) rc = launch_morph(1, width, height, PlaneU8{a, (size_t)width, N}, PlaneU8W{b, (size_t)width, N}, se, op, 0, nullptr);
    if (!rc) { cudaError_t e = cudaMemcpy2D(dst, dstep, b, width, width, height, cudaMemcpyDeviceToHost); if (e != cudaSuccess) rc = cuda_fail(e, "memcpy2d", __FILE__, __LINE__); }
    cudaFree(a); cudaFree(b);
    return rc;
}

extern "C" int rtdm_validate_disparity(int16_t *disp, size_t dstep, const int16_t *cost, size_t cstep,
                                       int width, int height, int minDisparity, int numDisparities,
                                       int disp12MaxDiff, int device)
{
    if (!disp || !cost || width < 1 || height < 1 || width > 8000 || dstep % 2 || cstep % 2) return -RTDM_EINVAL;
    int rc = check_device(device);
    if (rc) return rc;
    RTDM_CUDA(cudaSetDevice(device));
    int16_t *a = nullptr, *c = nullptr, *o = nullptr;
    const size_t N = (size_t)width * height;
    rc = dev_alloc(&a, N);
    if (!rc) rc = dev_alloc(&c, N);
    if (!rc) rc = dev_alloc(&o, N);
    if (!rc) { cudaError_t e = cudaMemcpy2D(a, (size_t)width * 2, disp, dstep, (size_t)width * 2, height, cudaMemcpyHostToDevice); if (e != cudaSuccess) rc = cuda_fail(e, "memcpy2d", __FILE__, __LINE__); }
    if (!rc) { cudaError_t e = cudaMemcpy2D(c, (size_t)width * 2, cost, cstep, (size_t)width * 2, height, cudaMemcpyHostToDevice); if (e != cudaSuccess) rc = cuda_fail(e, "memcpy2d", __FILE__, __LINE__); }
    if (!rc) rc = launch_validate_mask(1, width, height, minDisparity, numDisparities, disp12MaxDiff, 0, width, 0, width, 0, height,
                                       PlaneS16{a, (size_t)width, N}, PlaneS16{c, (size_t)width, N}, PlaneS16{o, (size_t)width, N}, 0, nullptr);
    if (!rc) { cudaError_t e = cudaMemcpy2D(disp, dstep, o, (size_t)width * 2, (size_t)width * 2, height, cudaMemcpyDeviceToHost); if (e != cudaSuccess) rc = cuda_fail(e, "memcpy2d", __FILE__, __LINE__); }
    cudaFree(a); cudaFree(c); cudaFree(o);
    return rc;
}

// =================================================================================================
// depth epilogue
// =================================================================================================
struct rtdm_depth {
    int maxW, maxH, maxR, dev;
    cudaStream_t st;
    int16_t *dD; uint8_t *dM; float *dX;       // staging for the host entry point (dX allocated on first use)
    int *rects, *minval, *counts; double *sums;
    double *hsums; int *hcounts;               // pinned result staging
    int launches;
};

extern "C" void rtdm_depth_destroy(rtdm_depth *h)
{
    if (!h) return;
    cudaSetDevice(h->dev);
    cudaFree(h->dD); cudaFree(h->dM); cudaFree(h->dX); cudaFree(h->rects); cudaFree(h->minval); cudaFree(h->counts); cudaFree(h->sums);
    cudaFreeHost(h->hsums); cudaFreeHost(h->hcounts);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

extern "C" int rtdm_depth_create(rtdm_depth **out, int max_width, int max_height, int max_regions, int device)
{
    if (!out) { set_error("depth_create: null argument"); return -RTDM_EINVAL; }
    *out = nullptr;
    if (max_width < 1 || max_height < 1 || max_regions < 0 || max_regions > 65535) { set_error("depth_create: bad geometry"); return -RTDM_EINVAL; }
    int rc = check_device(device);
    if (rc) return rc;
    RTDM_CUDA(cudaSetDevice(device));
    rtdm_depth *h = new (std::nothrow) rtdm_depth();
    if (!h) return -RTDM_ENOMEM;
    memset(h, 0, sizeof *h);
    h->maxW = max_width; h->maxH = max_height; h->maxR = max_regions; h->dev = device;
    const size_t N = (size_t)max_width * max_height, R = (size_t)std::max(max_regions, 1);
    rc = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) == cudaSuccess ? 0 : -RTDM_EIO;
    if (!rc) rc = dev_alloc(&h->dD, N);
    if (!rc) rc = dev_alloc(&h->dM, N);
    if (!rc) rc = dev_alloc(&h->rects, 4 * R);
    if (!rc) rc = dev_alloc(&h->minval, 1);
    if (!rc) rc = dev_alloc(&h->counts, R);
    if (!rc) rc = dev_alloc(&h->sums, R);
    if (!rc) rc = cudaMallocHost((void **)&h->hsums, R * sizeof(double)) == cudaSuccess ? 0 : -RTDM_ENOMEM;
    if (!rc) rc = cudaMallocHost((void **)&h->hcounts, R * sizeof(int)) == cudaSuccess ? 0 : -RTDM_ENOMEM;
    if (rc) { rtdm_depth_destroy(h); return rc; }
    *out = h;
    return 0;
}

static int depth_run(rtdm_depth *h, const int16_t *disp, size_t dpitch, int W, int H, const double *Q, const uint8_t *mask, size_t mpitch,
                     int nregions, const int *rects, double *mean_z, int *count, float *xyz, size_t xpitch, cudaStream_t st)
{
    if (nregions > 0) RTDM_CUDA(cudaMemcpyAsync(h->rects, rects, sizeof(int) * 4 * nregions, cudaMemcpyHostToDevice, st));
    h->launches = 0;
    int rc = launch_depth(disp, dpitch, W, H, Q, mask, mpitch, nregions, h->rects, h->minval, h->sums, h->counts, xyz, xpitch, st, &h->launches);
    if (rc) return rc;
    if (nregions > 0) {
        RTDM_CUDA(cudaMemcpyAsync(h->hsums, h->sums, sizeof(double) * nregions, cudaMemcpyDeviceToHost, st));
        RTDM_CUDA(cudaMemcpyAsync(h->hcounts, h->counts, sizeof(int) * nregions, cudaMemcpyDeviceToHost, st));
    }
    RTDM_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < nregions; i++) {
        count[i] = h->hcounts[i];
        mean_z[i] = h->hcounts[i] > 0 ? h->hsums[i] / h->hcounts[i] : 0.0;
    }
    return 0;
}

static int depth_check(rtdm_depth *h, const void *disp, int W, int H, const double *Q, int nregions, const int *rects,
                       const double *mean_z, const int *count)
{
    if (!h || !disp || !Q || (nregions > 0 && (!rects || !mean_z || !count))) { set_error("depth: null argument"); return -RTDM_EINVAL; }
    if (W < 1 || H < 1 || W > h->maxW || H > h->maxH || nregions < 0 || nregions > h->maxR) {
        set_error("depth: geometry or region count exceeds what the handle was created for");
        return -RTDM_EINVAL;
    }
    for (int i = 0; i < nregions; i++) {
        const int *r = rects + 4 * i;
        if (r[0] < 0 || r[1] < 0 || r[2] < 0 || r[3] < 0 || r[0] + r[2] > W || r[1] + r[3] > H) {
            set_error("depth: rectangle outside the image");
            return -RTDM_EINVAL;
        }
    }
    return 0;
}

extern "C" int rtdm_depth_run_device(rtdm_depth *h, const int16_t *disp, size_t dstep, int width, int height, const double *Q,
                                     const uint8_t *mask, size_t mstep, int nregions, const int *rects,
                                     double *mean_z, int *count, float *xyz, size_t xstep, void *cuda_stream)
{
    int rc = depth_check(h, disp, width, height, Q, nregions, rects, mean_z, count);
    if (rc) return rc;
    if (dstep % 2 || xstep % 4) { set_error("depth: steps must be multiples of the element size"); return -RTDM_EINVAL; }
    RTDM_CUDA(cudaSetDevice(h->dev));
    return depth_run(h, disp, dstep / 2, width, height, Q, mask, mstep, nregions, rects, mean_z, count, xyz, xstep / 4,
                     cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->st);
}

extern "C" int rtdm_depth_run(rtdm_depth *h, const int16_t *disp, size_t dstep, int width, int height, const double *Q,
                              const uint8_t *mask, size_t mstep, int nregions, const int *rects,
                              double *mean_z, int *count, float *xyz, size_t xstep)
{
    int rc = depth_check(h, disp, width, height, Q, nregions, rects, mean_z, count);
    if (rc) return rc;
    RTDM_CUDA(cudaSetDevice(h->dev));
    const size_t W = (size_t)width;
    RTDM_CUDA(cudaMemcpy2DAsync(h->dD, W * 2, disp, dstep, W * 2, height, cudaMemcpyHostToDevice, h->st));
    if (mask) RTDM_CUDA(cudaMemcpy2DAsync(h->dM, W, mask, mstep, W, height, cudaMemcpyHostToDevice, h->st));
    if (xyz && !h->dX) { rc = dev_alloc(&h->dX, (size_t)h->maxW * h->maxH * 3); if (rc) return rc; }
    rc = depth_run(h, h->dD, W, width, height, Q, mask ? h->dM : nullptr, W, nregions, rects, mean_z, count, xyz ? h->dX : nullptr, W * 3, h->st);
    if (rc) return rc;
    if (xyz) RTDM_CUDA(cudaMemcpy2D(xyz, xstep, h->dX, W * 12, W * 12, height, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int rtdm_depth_last_launches(const rtdm_depth *h) { return h ? h->launches : 0; }

// =================================================================================================
// rectification front-end
// =================================================================================================
struct rtdm_rectify {
    int W, H, rx, ry, rw, rh, maxB, dev;
    cudaStream_t st;
    int16_t *map1; uint16_t *map2;          // ROI part of the maps, tightly packed
    uint8_t *dIn, *dOut;                    // staging for the host entry point
    int launches;
};

extern "C" void rtdm_rectify_destroy(rtdm_rectify *h)
{
    if (!h) return;
    cudaSetDevice(h->dev);
    cudaFree(h->map1); cudaFree(h->map2); cudaFree(h->dIn); cudaFree(h->dOut);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

extern "C" int rtdm_rectify_create(rtdm_rectify **out, int src_width, int src_height, const int16_t *map1, size_t map1_step,
                                   const uint16_t *map2, size_t map2_step, int roi_x, int roi_y, int roi_width, int roi_height,
                                   int max_batch, int device)
{
    if (!out || !map1 || !map2) { set_error("rectify_create: null argument"); return -RTDM_EINVAL; }
    *out = nullptr;
    if (src_width < 1 || src_height < 1 || max_batch < 1 || roi_x < 0 || roi_y < 0 || roi_width < 1 || roi_height < 1 ||
        roi_x + roi_width > src_width || roi_y + roi_height > src_height) {
        set_error("rectify_create: bad geometry (the ROI must lie inside the image)");
        return -RTDM_EINVAL;
    }
    int rc = check_device(device);
    if (rc) return rc;
    RTDM_CUDA(cudaSetDevice(device));
    rtdm_rectify *h = new (std::nothrow) rtdm_rectify();
    if (!h) return -RTDM_ENOMEM;
    memset(h, 0, sizeof *h);
    h->W = src_width; h->H = src_height; h->rx = roi_x; h->ry = roi_y; h->rw = roi_width; h->rh = roi_height;
    h->maxB = max_batch; h->dev = device;
    const size_t RN = (size_t)roi_width * roi_height;
    rc = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) == cudaSuccess ? 0 : -RTDM_EIO;
    if (!rc) rc = dev_alloc(&h->map1, 2 * RN);
    if (!rc) rc = dev_alloc(&h->map2, RN);
    if (!rc) rc = dev_alloc(&h->dIn, (size_t)src_width * src_height * 3 * max_batch);
    if (!rc) rc = dev_alloc(&h->dOut, RN * max_batch);
    if (!rc) {
        cudaError_t e = cudaMemcpy2D(h->map1, (size_t)roi_width * 4, (const uint8_t *)map1 + (size_t)roi_y * map1_step + (size_t)roi_x * 4,
                                     map1_step, (size_t)roi_width * 4, roi_height, cudaMemcpyHostToDevice);
        if (e == cudaSuccess)
            e = cudaMemcpy2D(h->map2, (size_t)roi_width * 2, (const uint8_t *)map2 + (size_t)roi_y * map2_step + (size_t)roi_x * 2,
                             map2_step, (size_t)roi_width * 2, roi_height, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = cuda_fail(e, "rectify maps upload", __FILE__, __LINE__);
    }
    if (rc) { rtdm_rectify_destroy(h); return rc; }
    *out = h;
    return 0;
}

extern "C" int rtdm_rectify_run_device(rtdm_rectify *h, int n, const uint8_t *rgb, size_t step, size_t frame,
                                       uint8_t *out, size_t ostep, size_t oframe, void *cuda_stream)
{
    if (!h || !rgb || !out) { set_error("rectify: null argument"); return -RTDM_EINVAL; }
    if (n < 1 || step < (size_t)h->W * 3 || ostep < (size_t)h->rw) { set_error("rectify: bad batch or steps"); return -RTDM_EINVAL; }
    RTDM_CUDA(cudaSetDevice(h->dev));
    h->launches = 0;
    return launch_rectify(n, rgb, step, frame, h->W, h->H, h->map1, h->map2, h->rw, h->rh, out, ostep, oframe,
                          cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->st, &h->launches);
}

extern "C" int rtdm_rectify_run(rtdm_rectify *h, int n, const uint8_t *rgb, size_t step, size_t frame,
                                uint8_t *out, size_t ostep, size_t oframe)
{
    if (!h || !rgb || !out) { set_error("rectify: null argument"); return -RTDM_EINVAL; }
    if (n < 1 || n > h->maxB || step < (size_t)h->W * 3 || ostep < (size_t)h->rw) {
        set_error("rectify: batch exceeds what the handle was created for, or bad steps");
        return -RTDM_EINVAL;
    }
    RTDM_CUDA(cudaSetDevice(h->dev));
    const size_t row = (size_t)h->W * 3, fin = row * h->H, fout = (size_t)h->rw * h->rh;
    for (int k = 0; k < n; k++)
        RTDM_CUDA(cudaMemcpy2DAsync(h->dIn + k * fin, row, rgb + k * frame, step, row, h->H, cudaMemcpyHostToDevice, h->st));
    h->launches = 0;
    int rc = launch_rectify(n, h->dIn, row, fin, h->W, h->H, h->map1, h->map2, h->rw, h->rh, h->dOut, h->rw, fout, h->st, &h->launches);
    if (rc) return rc;
    for (int k = 0; k < n; k++)
        RTDM_CUDA(cudaMemcpy2DAsync(out + k * oframe, ostep, h->dOut + k * fout, h->rw, h->rw, h->rh, cudaMemcpyDeviceToHost, h->st));
    RTDM_CUDA(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int rtdm_rectify_last_launches(const rtdm_rectify *h) { return h ? h->launches : 0; }

// ===================================================================================================
// mask front-end (estimator.cpp:38-43) and back-end (estimator.cpp:46-53, :164-204)
// ===================================================================================================
struct rtdm_colormask {
    int W, H, rx, ry, rw, rh, maxB, dev;
    cudaStream_t st;
    int16_t *map1; uint16_t *map2;          // ROI part of the maps, tightly packed
    uint8_t *dIn, *dMask, *dBgr;            // staging for the host entry point
    int launches;
};

extern "C" void rtdm_colormask_destroy(rtdm_colormask *h)
{
    if (!h) return;
    cudaSetDevice(h->dev);
    cudaFree(h->map1); cudaFree(h->map2); cudaFree(h->dIn); cudaFree(h->dMask); cudaFree(h->dBgr);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

extern "C" int rtdm_colormask_create(rtdm_colormask **out, int src_width, int src_height, const int16_t *map1, size_t map1_step,
                                     const uint16_t *map2, size_t map2_step, int roi_x, int roi_y, int roi_width, int roi_height,
                                     int max_batch, int device)
{
    if (!out || !map1 || !map2) { set_error("colormask_create: null argument"); return -RTDM_EINVAL; }
    *out = nullptr;
    if (src_width < 1 || src_height < 1 || max_batch < 1 || roi_x < 0 || roi_y < 0 || roi_width < 1 || roi_height < 1 ||
        roi_x + roi_width > src_width || roi_y + roi_height > src_height) {
        set_error("colormask_create: bad geometry (the ROI must lie inside the image)");
        return -RTDM_EINVAL;
    }
    int rc = check_device(device);
    if (rc) return rc;
    RTDM_CUDA(cudaSetDevice(device));
    rtdm_colormask *h = new (std::nothrow) rtdm_colormask();
    if (!h) return -RTDM_ENOMEM;
    memset(h, 0, sizeof *h);
    h->W = src_width; h->H = src_height; h->rx = roi_x; h->ry = roi_y; h->rw = roi_width; h->rh = roi_height;
    h->maxB = max_batch; h->dev = device;
    const size_t RN = (size_t)roi_width * roi_height;
    rc = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) == cudaSuccess ? 0 : -RTDM_EIO;
    if (!rc) rc = dev_alloc(&h->map1, 2 * RN);
    if (!rc) rc = dev_alloc(&h->map2, RN);
    if (!rc) rc = dev_alloc(&h->dIn, (size_t)src_width * src_height * 3 * max_batch);
    if (!rc) rc = dev_alloc(&h->dMask, RN * max_batch);
    if (!rc) rc = dev_alloc(&h->dBgr, RN * 3 * max_batch);
    if (!rc) {
        cudaError_t e = cudaMemcpy2D(h->map1, (size_t)roi_width * 4, (const uint8_t *)map1 + (size_t)roi_y * map1_step + (size_t)roi_x * 4,
                                     map1_step, (size_t)roi_width * 4, roi_height, cudaMemcpyHostToDevice);
        if (e == cudaSuccess)
            e = cudaMemcpy2D(h->map2, (size_t)roi_width * 2, (const uint8_t *)map2 + (size_t)roi_y * map2_step + (size_t)roi_x * 2,
                             map2_step, (size_t)roi_width * 2, roi_height, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = cuda_fail(e, "colormask maps upload", __FILE__, __LINE__);
    }
    if (rc) { rtdm_colormask_destroy(h); return rc; }
    *out = h;
    return 0;
}

static int colormask_check_range(const int *low, const int *high)
{
    if (!low || !high) { set_error("colormask: null range"); return -RTDM_EINVAL; }
    return 0;
}

extern "C" int rtdm_colormask_run_device(rtdm_colormask *h, int n, const uint8_t *rgb, size_t step, size_t frame, const int *low,
                                         const int *high, uint8_t *mask, size_t mstep, size_t mframe, uint8_t *bgr, size_t bstep,
                                         size_t bframe, void *cuda_stream)
{
    if (!h || !rgb || !mask) { set_error("colormask: null argument"); return -RTDM_EINVAL; }
    int rc = colormask_check_range(low, high);
    if (rc) return rc;
    if (n < 1 || step < (size_t)h->W * 3 || mstep < (size_t)h->rw || (bgr && bstep < (size_t)h->rw * 3)) {
        set_error("colormask: bad batch or steps");
        return -RTDM_EINVAL;
    }
    RTDM_CUDA(cudaSetDevice(h->dev));
    h->launches = 0;
    return launch_colormask(n, rgb, step, frame, h->W, h->H, h->map1, h->map2, h->rw, h->rh, low, high, mask, mstep, mframe,
                            bgr, bstep, bframe, cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->st, &h->launches);
}

extern "C" int rtdm_colormask_run(rtdm_colormask *h, int n, const uint8_t *rgb, size_t step, size_t frame, const int *low, const int *high,
                                  uint8_t *mask, size_t mstep, size_t mframe, uint8_t *bgr, size_t bstep, size_t bframe)
{
    if (!h || !rgb || !mask) { set_error("colormask: null argument"); return -RTDM_EINVAL; }
    int rc = colormask_check_range(low, high);
    if (rc) return rc;
    if (n < 1 || n > h->maxB || step < (size_t)h->W * 3 || mstep < (size_t)h->rw || (bgr && bstep < (size_t)h->rw * 3)) {
        set_error("colormask: batch exceeds what the handle was created for, or bad steps");
        return -RTDM_EINVAL;
    }
    RTDM_CUDA(cudaSetDevice(h->dev));
    const size_t row = (size_t)h->W * 3, fin = row * h->H, fm = (size_t)h->rw * h->rh;
    for (int k = 0; k < n; k++)
        RTDM_CUDA(cudaMemcpy2DAsync(h->dIn + k * fin, row, rgb + k * frame, step, row, h->H, cudaMemcpyHostToDevice, h->st));
    h->launches = 0;
    rc = launch_colormask(n, h->dIn, row, fin, h->W, h->H, h->map1, h->map2, h->rw, h->rh, low, high, h->dMask, h->rw, fm,
                          bgr ? h->dBgr : nullptr, (size_t)h->rw * 3, fm * 3, h->st, &h->launches);
    if (rc) return rc;
    for (int k = 0; k < n; k++) {
        RTDM_CUDA(cudaMemcpy2DAsync(mask + k * mframe, mstep, h->dMask + k * fm, h->rw, h->rw, h->rh, cudaMemcpyDeviceToHost, h->st));
        if (bgr)
            RTDM_CUDA(cudaMemcpy2DAsync(bgr + k * bframe, bstep, h->dBgr + k * fm * 3, (size_t)h->rw * 3, (size_t)h->rw * 3, h->rh,
                                        cudaMemcpyDeviceToHost, h->st));
    }
    RTDM_CUDA(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int rtdm_colormask_last_launches(const rtdm_colormask *h) { return h ? h->launches : 0; }

struct rtdm_regions {
    int maxW, maxH, maxR, dev;
    cudaStream_t st;
    uint8_t *dMask;                           // staging for the host entry point
    int *labels, *ext, *keys, *out;
    int4 *bb, *boxes;
    int *hOut;                                // pinned result: 6 + 4 * maxR ints
    int launches;
};

extern "C" void rtdm_regions_destroy(rtdm_regions *h)
{
    if (!h) return;
    cudaSetDevice(h->dev);
    cudaFree(h->dMask); cudaFree(h->labels); cudaFree(h->ext); cudaFree(h->keys); cudaFree(h->out);
    cudaFree(h->bb); cudaFree(h->boxes);
    if (h->hOut) cudaFreeHost(h->hOut);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

extern "C" int rtdm_regions_create(rtdm_regions **out, int max_width, int max_height, int max_regions, int device)
{
    if (!out) { set_error("regions_create: null argument"); return -RTDM_EINVAL; }
    *out = nullptr;
    if (max_width < 1 || max_height < 1 || max_regions < 1 || (long long)max_width * max_height > (1LL << 30)) {
        set_error("regions_create: bad geometry");
        return -RTDM_EINVAL;
    }
    int rc = check_device(device);
    if (rc) return rc;
    RTDM_CUDA(cudaSetDevice(device));
    rtdm_regions *h = new (std::nothrow) rtdm_regions();
    if (!h) return -RTDM_ENOMEM;
    memset(h, 0, sizeof *h);
    h->maxW = max_width; h->maxH = max_height; h->maxR = max_regions; h->dev = device;
    const size_t N = (size_t)max_width * max_height;
    rc = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) == cudaSuccess ? 0 : -RTDM_EIO;
    if (!rc) rc = dev_alloc(&h->dMask, N);
    if (!rc) rc = dev_alloc(&h->labels, N + 1);
    if (!rc) rc = dev_alloc(&h->ext, N);
    if (!rc) rc = dev_alloc(&h->bb, N);
    if (!rc) rc = dev_alloc(&h->keys, (size_t)max_regions);
    if (!rc) rc = dev_alloc(&h->boxes, (size_t)max_regions);
    if (!rc) rc = dev_alloc(&h->out, 6 + 4 * (size_t)max_regions);
    if (!rc && cudaMallocHost(&h->hOut, (6 + 4 * (size_t)max_regions) * sizeof(int)) != cudaSuccess) rc = -RTDM_ENOMEM;
    if (rc) { rtdm_regions_destroy(h); return rc; }
    *out = h;
    return 0;
}

static int regions_finish(rtdm_regions *h, cudaStream_t st, int *rects, int *count, int *ncontours, int *roi)
{
    RTDM_CUDA(cudaMemcpyAsync(h->hOut, h->out, (6 + 4 * (size_t)h->maxR) * sizeof(int), cudaMemcpyDeviceToHost, st));
    RTDM_CUDA(cudaStreamSynchronize(st));
    const int n = h->hOut[0];
    if (n > h->maxR) { set_error("regions: more boxes than max_regions"); return -RTDM_EINVAL; }
    if (count) *count = n;
    if (ncontours) *ncontours = h->hOut[1];
    if (roi) { roi[0] = h->hOut[2]; roi[1] = h->hOut[3]; roi[2] = h->hOut[4] - h->hOut[2]; roi[3] = h->hOut[5] - h->hOut[3]; }
    if (rects) memcpy(rects, h->hOut + 6, (size_t)n * 4 * sizeof(int));
    return 0;
}

extern "C" int rtdm_regions_run_device(rtdm_regions *h, const uint8_t *mask, size_t mstep, int width, int height, int min_obj_size,
                                       int *rects, int *count, int *ncontours, int *roi, void *cuda_stream)
{
    if (!h || !mask) { set_error("regions: null argument"); return -RTDM_EINVAL; }
    if (width < 1 || height < 1 || width > h->maxW || height > h->maxH || mstep < (size_t)width) {
        set_error("regions: size exceeds what the handle was created for, or bad step");
        return -RTDM_EINVAL;
    }
    RTDM_CUDA(cudaSetDevice(h->dev));
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->st;
    h->launches = 0;
    int rc = launch_regions(mask, mstep, width, height, min_obj_size, h->maxR, h->labels, h->bb, h->ext, h->keys, h->boxes,
                            h->out, st, &h->launches);
    if (rc) return rc;
    return regions_finish(h, st, rects, count, ncontours, roi);
}

extern "C" int rtdm_regions_run(rtdm_regions *h, const uint8_t *mask, size_t mstep, int width, int height, int min_obj_size,
                                int *rects, int *count, int *ncontours, int *roi)
{
    if (!h || !mask) { set_error("regions: null argument"); return -RTDM_EINVAL; }
    if (width < 1 || height < 1 || width > h->maxW || height > h->maxH || mstep < (size_t)width) {
        set_error("regions: size exceeds what the handle was created for, or bad step");
        return -RTDM_EINVAL;
    }
    RTDM_CUDA(cudaSetDevice(h->dev));
    RTDM_CUDA(cudaMemcpy2DAsync(h->dMask, width, mask, mstep, width, height, cudaMemcpyHostToDevice, h->st));
    h->launches = 0;
    int rc = launch_regions(h->dMask, width, width, height, min_obj_size, h->maxR, h->labels, h->bb, h->ext, h->keys, h->boxes,
                            h->out, h->st, &h->launches);
    if (rc) return rc;
    return regions_finish(h, h->st, rects, count, ncontours, roi);
}

extern "C" int rtdm_regions_last_launches(const rtdm_regions *h) { return h ? h->launches : 0; }

extern "C" int rtdm_measure_int_peak(int device, double *tiops_iadd3, double *tiops_vimnmx,
                                     double *tiops_vabsdiff4, double *sm_mhz_est)
{
    int rc = check_device(device);
    if (rc) return rc;
    return measure_int_peak(device, tiops_iadd3, tiops_vimnmx, tiops_vabsdiff4, sm_mhz_est);
}

namespace rtdm { int measure_op_rates(int device, double *out, int n); }
// development aid (not part of the plugin boundary): statement rates of 15 integer idioms, in 1e12
// statements/s over the chip (a statement may be 1 or 2 SASS instructions, see intpeak.cu)
extern "C" int rtdm_dev_op_rates(int device, double *out, int n)
{
    int rc = check_device(device);
    if (rc) return rc;
    return rtdm::measure_op_rates(device, out, n);
}
