// postproc.cu -- disparity post-processing kernels: left-right validation + valid-rect mask,
// speckle filter (connected components), 3x3 median.
//
// Replaces, inside cv::StereoBM::compute / cv::StereoSGBM::compute as reached from
// SWMatcherKonolige::compute (reference stereo-matcher/bm-sw.cpp:33-38) and
// SWSemiGlobalMatcher::compute (stereo-matcher/sgbm-sw.cpp:32-37):
//   cv::validateDisparity  (oracle: orc_validate_disparity)
//   valid-rect masking     (getValidDisparityROI; oracle: orc_valid_roi / orc_bm_compute)
//   cv::filterSpeckles     (App. A.4; oracle: orc_filter_speckles)
//   cv::medianBlur(.., 3)  (App. A.6 tail; oracle: orc_median3_s16)
#include "common.cuh"
#include <cstdlib>

namespace rtdm {

// ------------------------------------------------------------------------------------------------
// validateDisparity + mask.  One CTA per (row, frame).  OpenCV's per-row serial scan
//   for x ascending: x2 = x - round(d/16); if (cost2[x2] > c) { cost2[x2] = c; disp2[x2] = d; }
// keeps, per x2, the candidate with the smallest cost and among equal costs the smallest x:
// exactly an atomicMin over the key (cost << 16 | x).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
validate_mask_kernel(int W, int H, int minD, int nd, int d12, int lofs, int W1,
                     int vx0, int vx1, int row0, int row1,
                     PlaneS16 raw, PlaneS16 cost, PlaneS16 out, const int16_t *spill)
{
    extern __shared__ uint32_t vm_smem[];
    uint32_t *key = vm_smem;                                 // [W]
    int16_t *sd = reinterpret_cast<int16_t *>(key + W);      // [W] raw disparity row
    const int y = blockIdx.x, f = blockIdx.y;
    const int INV = (minD - 1) * 16;
    int16_t *orow = out.p + (size_t)f * out.frame + (size_t)y * out.pitch;
    if (y < row0 || y >= row1) {
        // row1 keeps what the last computed row wrote beyond its end (minDisparity > 0, BmGeom::spill)
        const bool sp = spill && y == row1 && row1 > row0;
        for (int x = threadIdx.x; x < W; x += blockDim.x) orow[x] = (sp && x < minD) ? spill[(size_t)f * minD + x] : (int16_t)INV;
        return;
    }
    const int16_t *drow = raw.p + (size_t)f * raw.frame + (size_t)y * raw.pitch;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        sd[x] = (x >= lofs && x < lofs + W1) ? drow[x] : (int16_t)INV;
        key[x] = 0xFFFFFFFFu;
    }
    __syncthreads();
    if (d12 >= 0) {
        const int16_t *crow = cost.p + (size_t)f * cost.frame + (size_t)y * cost.pitch;
        const int minX1 = max(minD + nd, 0), maxX1 = W + min(minD, 0);
        for (int x = minX1 + threadIdx.x; x < maxX1; x += blockDim.x) {
            int d = sd[x];
            if (d == INV) continue;
            int x2 = x - ((d + 8) >> 4);
            if (x2 < 0 || x2 >= W) continue;
            uint32_t c = (uint32_t)(uint16_t)crow[x];
            atomicMin(&key[x2], (c << 16) | (uint32_t)x);
        }
        __syncthreads();
        const int lim = d12 * 16;
        for (int x = threadIdx.x; x < W; x += blockDim.x) {
            int d = sd[x];
            int o = d;
            if (x >= minX1 && x < maxX1 && d != INV) {
                int d0 = d >> 4, d1 = (d + 15) >> 4;
                int xa = x - d0, xb = x - d1;
                bool bad = true;
                if (0 <= xa && xa < W) {
                    uint32_t k = key[xa];
                    int d2 = (k == 0xFFFFFFFFu) ? INV : (int)sd[k & 0xFFFFu];
                    bad = bad && (d2 > INV) && (abs(d2 - d) > lim);
                } else bad = false;
                if (0 <= xb && xb < W) {
                    uint32_t k = key[xb];
                    int d2 = (k == 0xFFFFFFFFu) ? INV : (int)sd[k & 0xFFFFu];
                    bad = bad && (d2 > INV) && (abs(d2 - d) > lim);
                } else bad = false;
                if (bad) o = INV;
            }
            if (x < vx0 || x >= vx1) o = INV;
            orow[x] = (int16_t)o;
        }
    } else {
        for (int x = threadIdx.x; x < W; x += blockDim.x) {
            int o = sd[x];
            if (x < vx0 || x >= vx1) o = INV;
            orow[x] = (int16_t)o;
        }
    }
}

int launch_validate_mask(int n, int W, int H, int minD, int nd, int d12, int lofs, int W1,
                         int vx0, int vx1, int row0, int row1,
                         PlaneS16 raw, PlaneS16 cost, PlaneS16 out, cudaStream_t st, int *launches, const int16_t *spill)
{
    if (n <= 0) return 0;
    dim3 grid(H, n);
    size_t smem = (size_t)W * 6 + 8;
    validate_mask_kernel<<<grid, 256, smem, st>>>(W, H, minD, nd, d12, lofs, W1, vx0, vx1, row0, row1,
                                                 raw, cost, out, spill);
    if (launches) (*launches)++;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// filterSpeckles == delete 4-connected components (edges where both pixels != newVal and
// |a - b| <= maxDiff) of at most maxSize pixels.  Union-find over HORIZONTAL RUNS:
//   1. row runs     : per row, a block-wide max-scan finds every pixel's run start.  Run-start pixels are the
//                     union-find nodes (label = own index, run length stored, size accumulator cleared); all
//                     other valid pixels just point at their run start (label = start | RUN_FLAG)
//   2. vertical merge: union(run of (x,y), run of (x,y-1)) where the two pixels are connected; a link is skipped
//                     when the pixel pair to the left already joins the same two runs
//   3. count        : each run start finds its root, flattens, and adds its run length to sizes[root]
//   4. apply        : pixel -> run start -> root; components with size <= maxSize become newVal
// The component partition is unique, so the result equals OpenCV's flood fill for any scan order.
// ------------------------------------------------------------------------------------------------
constexpr int32_t RUN_FLAG = 0x40000000;       // W*H < 2^30 for every supported geometry

__device__ __forceinline__ int uf_find(const int32_t *lab, int i)
{
    int p = lab[i];
    while (p != i) { i = p; p = lab[i]; }
    return i;
}

// find with path halving: every visited node is re-pointed at its grandparent.  The plain stores race benignly with
// concurrent unions (a node only ever moves to an ancestor, and ancestors have smaller indices).
__device__ __forceinline__ int uf_find_halve(int32_t *lab, int i)
{
    int p = lab[i];
    while (p != i) {
        const int gp = lab[p];
        if (gp != p) lab[i] = gp;
        i = p; p = gp;
    }
    return i;
}

__device__ __forceinline__ void uf_union(int32_t *lab, int a, int b)
{
    while (true) {
        a = uf_find_halve(lab, a);
        b = uf_find_halve(lab, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }          // a > b : hook a under b
        int old = atomicMin(&lab[a], b);
        if (old == a) return;
        a = old;                                         // someone else hooked a; retry with it
    }
}

// one CTA per (row, frame)
__global__ void __launch_bounds__(256)
speckle_rowruns_kernel(int W, int H, PlaneS16 img, int newVal, int maxDiff, int32_t *labels, int32_t *sizes, int32_t *runlen)
{
    extern __shared__ int cnt[];                         // [W] pixels per run, indexed by run-start column
    __shared__ int warp_last[8];
    const int y = blockIdx.x, f = blockIdx.y;
    const int16_t *row = img.p + (size_t)f * img.frame + (size_t)y * img.pitch;
    const size_t rowbase = ((size_t)f * H + y) * W;
    int32_t *lab = labels + rowbase, *siz = sizes + rowbase, *rlen = runlen + rowbase;
    const int chunk = (W + 255) / 256;
    const int xa = threadIdx.x * chunk, xb = min(xa + chunk, W);
    for (int x = threadIdx.x; x < W; x += 256) cnt[x] = 0;
    // marker(x) = x where a run starts (or the pixel is invalid), -1 where the run continues
    int run = -1;
    int prev = (xa > 0 && xa < W) ? (int)row[xa - 1] : newVal;
    for (int x = xa; x < xb; x++) {
        int v = row[x];
        bool cont = (v != newVal) && (x > 0) && (prev != newVal) && (abs(prev - v) <= maxDiff);
        if (!cont) run = x;
        prev = v;
    }
    // exclusive max-scan of `run` over the threads of the block
    int incl = run;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl = max(incl, t);
    }
    if (lane == 31) warp_last[wid] = incl;
    int carry = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if (lane == 0) carry = -1;
    __syncthreads();
    for (int w = 0; w < wid; w++) carry = max(carry, warp_last[w]);
    run = carry;
    prev = (xa > 0 && xa < W) ? (int)row[xa - 1] : newVal;
    const int rowofs = y * W;                            // frame-local linear index of column 0
    int pending = 0;                                     // pixels of the current run seen by this thread
    for (int x = xa; x < xb; x++) {
        int v = row[x];
        bool cont = (v != newVal) && (x > 0) && (prev != newVal) && (abs(prev - v) <= maxDiff);
        if (!cont) {
            if (pending) atomicAdd(&cnt[run], pending);
            pending = 0;
            run = x;
        }
        prev = v;
        if (v == newVal) lab[x] = -1;
        else { lab[x] = (x == run) ? rowofs + x : ((rowofs + run) | RUN_FLAG); pending++; }
    }
    if (pending) atomicAdd(&cnt[run], pending);
    __syncthreads();
    for (int x = xa; x < xb; x++) {
        const int c = cnt[x];
        if (c > 0) { rlen[x] = c; siz[x] = 0; }          // only run starts own a counter
    }
}

__global__ void __launch_bounds__(256)
speckle_vmerge_kernel(int W, int H, PlaneS16 img, int newVal, int maxDiff, int32_t *labels)
{
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y + 1;
    if (x >= W || y >= H) return;
    const int16_t *im = img.p + (size_t)f * img.frame;
    int32_t *lab = labels + (size_t)f * H * W;
    int v = im[(size_t)y * img.pitch + x], u = im[(size_t)(y - 1) * img.pitch + x];
    if (v == newVal || u == newVal || abs(u - v) > maxDiff) return;
    // skip if the left neighbours form the same vertical link already (same two runs)
    if (x > 0) {
        int vl = im[(size_t)y * img.pitch + x - 1], ul = im[(size_t)(y - 1) * img.pitch + x - 1];
        if (vl != newVal && ul != newVal && abs(vl - v) <= maxDiff && abs(ul - u) <= maxDiff &&
            abs(ul - vl) <= maxDiff) return;
    }
    int a = lab[y * W + x], b = lab[(y - 1) * W + x];     // pixel -> node (its run start)
    a = (a & RUN_FLAG) ? (a & ~RUN_FLAG) : y * W + x;
    b = (b & RUN_FLAG) ? (b & ~RUN_FLAG) : (y - 1) * W + x;
    uf_union(lab, a, b);
}

__global__ void __launch_bounds__(256)
speckle_count_kernel(int W, int H, int32_t *labels, int32_t *sizes, const int32_t *runlen)
{
    const int f = blockIdx.y;
    const size_t N = (size_t)W * H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int32_t *lab = labels + (size_t)f * N;
    const int l = lab[i];
    if (l < 0 || (l & RUN_FLAG)) return;                  // invalid pixel or not a run start
    const int root = uf_find(lab, (int)i);
    lab[i] = root;                                       // flatten: only shortens the path to the final root
    atomicAdd(&sizes[(size_t)f * N + root], runlen[(size_t)f * N + i]);
}

__global__ void __launch_bounds__(256)
speckle_apply_kernel(int W, int H, PlaneS16 img, int newVal, int maxSize, const int32_t *labels, const int32_t *sizes)
{
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const size_t N = (size_t)W * H;
    const int32_t *lab = labels + (size_t)f * N;
    int l = lab[(size_t)y * W + x];
    if (l < 0) return;
    const int start = (l & RUN_FLAG) ? (l & ~RUN_FLAG) : y * W + x;
    int root = lab[start];                               // flattened by the count kernel
    root = (root & RUN_FLAG) ? start : root;             // (cannot happen: starts never carry the flag)
    if (sizes[(size_t)f * N + root] <= maxSize)
        img.p[(size_t)f * img.frame + (size_t)y * img.pitch + x] = (int16_t)newVal;
}

// ---- 8 pixels per thread (W % 8 == 0, 16-byte aligned rows): 128-bit loads, one thread block per row ----------
__device__ __forceinline__ void load8_s16(const int16_t *p, int v[8])
{
    const uint4 q = *reinterpret_cast<const uint4 *>(p);
    v[0] = (int16_t)(q.x & 0xFFFFu); v[1] = (int16_t)(q.x >> 16); v[2] = (int16_t)(q.y & 0xFFFFu); v[3] = (int16_t)(q.y >> 16);
    v[4] = (int16_t)(q.z & 0xFFFFu); v[5] = (int16_t)(q.z >> 16); v[6] = (int16_t)(q.w & 0xFFFFu); v[7] = (int16_t)(q.w >> 16);
}

// Row kernel, one thread block per (row, frame), 8 pixels per thread:
//   VALIDATE: validateDisparity + valid-rectangle mask of the raw row (same arithmetic as validate_mask_kernel),
//             written to `img`; then, if `runs`, the row-run pass of the speckle filter on the values still in registers
//   else    : the row-run pass alone on `img`.
struct ValArgs {
    int minD, nd, d12, lofs, W1, vx0, vx1, row0, row1;
    PlaneS16 raw, cost;
    const int16_t *spill;       // BmGeom::spill (minDisparity > 0), else nullptr
};

template <bool VALIDATE>
__global__ void __launch_bounds__(1024)
post_row8_kernel(int W, int H, PlaneS16 img, ValArgs va, int runs, int newVal, int maxDiff,
                 int32_t *labels, int32_t *sizes, int32_t *runlen)
{
    extern __shared__ __align__(16) int row_smem[];
    int *cnt = row_smem;                                 // [W] pixels per run, indexed by run-start column
    __shared__ int warp_last[32], warp_tail[32];
    const int y = blockIdx.x, f = blockIdx.y;
    int16_t *row = img.p + (size_t)f * img.frame + (size_t)y * img.pitch;
    const int x0 = threadIdx.x * 8;
    const bool active = x0 < W;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int v[8];
    if (VALIDATE) {
        uint32_t *key = reinterpret_cast<uint32_t *>(row_smem + W);            // [W]
        int16_t *sd = reinterpret_cast<int16_t *>(key + W);                    // [W] raw disparity row
        const int INV = (va.minD - 1) * 16;
        const bool inrows = y >= va.row0 && y < va.row1;                       // block-uniform
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = INV;
        if (inrows) {
            if (active) {
                load8_s16(va.raw.p + (size_t)f * va.raw.frame + (size_t)y * va.raw.pitch + x0, v);
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (x0 + k < va.lofs || x0 + k >= va.lofs + va.W1) v[k] = INV;
                if (va.d12 >= 0) {
                    uint32_t pk[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) pk[k] = (uint32_t)(uint16_t)v[2 * k] | ((uint32_t)(uint16_t)v[2 * k + 1] << 16);
                    *reinterpret_cast<uint4 *>(sd + x0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    reinterpret_cast<uint4 *>(key + x0)[0] = make_uint4(~0u, ~0u, ~0u, ~0u);
                    reinterpret_cast<uint4 *>(key + x0)[1] = make_uint4(~0u, ~0u, ~0u, ~0u);
                }
            }
            if (va.d12 >= 0) {
                __syncthreads();
                const int minX1 = max(va.minD + va.nd, 0), maxX1 = W + min(va.minD, 0);
                if (active) {
                    int c[8];
                    load8_s16(va.cost.p + (size_t)f * va.cost.frame + (size_t)y * va.cost.pitch + x0, c);
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const int x = x0 + k, d = v[k];
                        if (x < minX1 || x >= maxX1 || d == INV) continue;
                        const int x2 = x - ((d + 8) >> 4);
                        if (x2 < 0 || x2 >= W) continue;
                        atomicMin(&key[x2], ((uint32_t)(uint16_t)c[k] << 16) | (uint32_t)x);
                    }
                }
                __syncthreads();
                if (active) {
                    const int lim = va.d12 * 16;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const int x = x0 + k, d = v[k];
                        if (x < minX1 || x >= maxX1 || d == INV) continue;
                        const int xa = x - (d >> 4), xb = x - ((d + 15) >> 4);
                        bool bad = true;
                        if (0 <= xa && xa < W) {
                            const uint32_t kk = key[xa];
                            const int d2 = (kk == 0xFFFFFFFFu) ? INV : (int)sd[kk & 0xFFFFu];
                            bad = bad && (d2 > INV) && (abs(d2 - d) > lim);
                        } else bad = false;
                        if (0 <= xb && xb < W) {
                            const uint32_t kk = key[xb];
                            const int d2 = (kk == 0xFFFFFFFFu) ? INV : (int)sd[kk & 0xFFFFu];
                            bad = bad && (d2 > INV) && (abs(d2 - d) > lim);
                        } else bad = false;
                        if (bad) v[k] = INV;
                    }
                }
            }
        }
        if (active) {
            uint32_t pk[4];
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (x0 + k < va.vx0 || x0 + k >= va.vx1) v[k] = INV;
            if (va.spill && y == va.row1 && va.row1 > va.row0) {                 // block-uniform; minDisparity > 0 only
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (x0 + k < va.minD) v[k] = va.spill[(size_t)f * va.minD + x0 + k];
            }
#pragma unroll
            for (int k = 0; k < 4; k++) pk[k] = (uint32_t)(uint16_t)v[2 * k] | ((uint32_t)(uint16_t)v[2 * k + 1] << 16);
            *reinterpret_cast<uint4 *>(row + x0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        if (!runs) return;                                                    // block-uniform
    } else {
        if (active) load8_s16(row + x0, v);
    }
    if (!active) {
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = newVal;
    }
    // ---- row runs ------------------------------------------------------------------------------------------
    const size_t rowbase = ((size_t)f * H + y) * W;
    int32_t *lab = labels + rowbase, *siz = sizes + rowbase, *rlen = runlen + rowbase;
    for (int x = threadIdx.x; x < W; x += blockDim.x) cnt[x] = 0;
    if (lane == 31) warp_tail[wid] = v[7];
    __syncthreads();
    int prev = __shfl_up_sync(0xFFFFFFFFu, v[7], 1);
    if (lane == 0) prev = wid > 0 ? warp_tail[wid - 1] : newVal;
    // bit k: pixel x0 + k continues the run of its left neighbour
    uint32_t cont = 0u;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int p = k ? v[k - 1] : prev;
        if (v[k] != newVal && (x0 + k) > 0 && p != newVal && abs(p - v[k]) <= maxDiff) cont |= 1u << k;
    }
    int run = -1;                                        // last run start (or invalid pixel) inside this thread's pixels
    if (active && cont != 0xFFu) run = x0 + 31 - __clz((~cont) & 0xFFu);
    // exclusive max-scan over the block
    int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl = max(incl, t);
    }
    if (lane == 31) warp_last[wid] = incl;
    int carry = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if (lane == 0) carry = -1;
    __syncthreads();
    for (int w = 0; w < wid && w < nwarp; w++) carry = max(carry, warp_last[w]);
    if (active) {
        const int rowofs = y * W;
        int cur = carry, pending = 0;
        int o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (!((cont >> k) & 1u)) {
                if (pending) atomicAdd(&cnt[cur], pending);
                pending = 0;
                cur = x0 + k;
            }
            if (v[k] == newVal) o[k] = -1;
            else { o[k] = (x0 + k == cur) ? rowofs + x0 + k : ((rowofs + cur) | RUN_FLAG); pending++; }
        }
        if (pending) atomicAdd(&cnt[cur], pending);
        reinterpret_cast<int4 *>(lab + x0)[0] = make_int4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<int4 *>(lab + x0)[1] = make_int4(o[4], o[5], o[6], o[7]);
    }
    __syncthreads();
    if (active) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int c = cnt[x0 + k];
            if (c > 0) { rlen[x0 + k] = c; siz[x0 + k] = 0; }      // only run starts own a counter
        }
    }
}

__global__ void __launch_bounds__(128)
speckle_vmerge8_kernel(int W, int H, PlaneS16 img, int newVal, int maxDiff, int32_t *labels)
{
    const int f = blockIdx.z;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const int y = blockIdx.y + 1;
    if (x0 >= W || y >= H) return;
    const int16_t *r1 = img.p + (size_t)f * img.frame + (size_t)y * img.pitch, *r0 = r1 - img.pitch;
    int v[8], u[8];
    load8_s16(r1 + x0, v);
    load8_s16(r0 + x0, u);
    // cheap reject: no vertical link at all in these 8 columns
    uint32_t link = 0u;
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (v[k] != newVal && u[k] != newVal && abs(u[k] - v[k]) <= maxDiff) link |= 1u << k;
    if (!link) return;
    const int vl = x0 > 0 ? (int)r1[x0 - 1] : newVal, ul = x0 > 0 ? (int)r0[x0 - 1] : newVal;
    int32_t *lab = labels + (size_t)f * H * W;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (!((link >> k) & 1u)) continue;
        const int pv = k ? v[k - 1] : vl, pu = k ? u[k - 1] : ul;
        // skip if the left neighbours form the same vertical link already (same two runs)
        if (pv != newVal && pu != newVal && abs(pv - v[k]) <= maxDiff && abs(pu - u[k]) <= maxDiff && abs(pu - pv) <= maxDiff) continue;
        const int x = x0 + k;
        int a = lab[y * W + x], b = lab[(y - 1) * W + x];     // pixel -> node (its run start)
        a = (a & RUN_FLAG) ? (a & ~RUN_FLAG) : y * W + x;
        b = (b & RUN_FLAG) ? (b & ~RUN_FLAG) : (y - 1) * W + x;
        uf_union(lab, a, b);
    }
}

// MODE 0: count (run start -> root, flatten, sizes[root] += run length); MODE 1: apply (runs of small components
// are overwritten with newVal).  Both scan the labels 8 at a time; only run starts do any work.
template <int MODE>
__global__ void __launch_bounds__(256)
speckle_runs8_kernel(int W, int H, PlaneS16 img, int newVal, int maxSize, int32_t *labels, int32_t *sizes, const int32_t *runlen)
{
    const int f = blockIdx.y;
    const size_t N = (size_t)W * H;
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i0 >= N) return;
    int32_t *lab = labels + (size_t)f * N;
    const int4 q0 = reinterpret_cast<const int4 *>(lab + i0)[0], q1 = reinterpret_cast<const int4 *>(lab + i0)[1];
    const int l[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (l[k] < 0 || (l[k] & RUN_FLAG)) continue;          // invalid pixel or not a run start
        const int i = (int)i0 + k;
        if (MODE == 0) {
            const int root = uf_find(lab, i);
            lab[i] = root;                                   // flatten: only shortens the path to the final root
            atomicAdd(&sizes[(size_t)f * N + root], runlen[(size_t)f * N + i]);
        } else {
            if (sizes[(size_t)f * N + l[k]] > maxSize) continue;      // l[k] is the root (flattened by MODE 0)
            const int y = i / W, x = i - y * W, len = runlen[(size_t)f * N + i];
            int16_t *p = img.p + (size_t)f * img.frame + (size_t)y * img.pitch + x;
            for (int j = 0; j < len; j++) p[j] = (int16_t)newVal;
        }
    }
}

// can launch_speckle take the 8-pixel kernels for this geometry?  (launch_validate_speckle relies on it)
static bool speckle_vec_ok(int W, PlaneS16 img, const int32_t *labels, const Switches &sw)
{
    return W % 8 == 0 && W <= 8192 && ((reinterpret_cast<uintptr_t>(img.p) | (img.pitch * 2) | (img.frame * 2)) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(labels) & 15) == 0 && !sw.speckle_scalar;
}

static int launch_speckle_impl(int n, int W, int H, PlaneS16 img, int newVal, int maxSize, int maxDiff,
                               int32_t *labels, int32_t *sizes, cudaStream_t st, int *launches, int32_t *runlen, bool rowruns_done,
                               const Switches &sw)
{
    if (n <= 0) return 0;
    const bool vec = speckle_vec_ok(W, img, labels, sw);
    if (vec) {
        const size_t N8 = ((size_t)W * H + 7) / 8;
        const int nt = ((W / 8) + 31) & ~31;
        if (!rowruns_done)
            post_row8_kernel<false><<<dim3(H, n), nt, (size_t)W * sizeof(int), st>>>(W, H, img, ValArgs(), 1, newVal, maxDiff, labels, sizes, runlen);
        if (H > 1)
            speckle_vmerge8_kernel<<<dim3(cdiv(W / 8, 128), H - 1, n), 128, 0, st>>>(W, H, img, newVal, maxDiff, labels);
        speckle_runs8_kernel<0><<<dim3((unsigned)((N8 + 255) / 256), n), 256, 0, st>>>(W, H, img, newVal, maxSize, labels, sizes, runlen);
        speckle_runs8_kernel<1><<<dim3((unsigned)((N8 + 255) / 256), n), 256, 0, st>>>(W, H, img, newVal, maxSize, labels, sizes, runlen);
        if (launches) (*launches) += ((H > 1) ? 4 : 3) - (rowruns_done ? 1 : 0);
        RTDM_CUDA(cudaGetLastError());
        return 0;
    }
    if (rowruns_done) { set_error("speckle: fused row pass without the vector path"); return -RTDM_EINVAL; }
    speckle_rowruns_kernel<<<dim3(H, n), 256, (size_t)W * sizeof(int), st>>>(W, H, img, newVal, maxDiff, labels, sizes, runlen);
    if (H > 1)
        speckle_vmerge_kernel<<<dim3(cdiv(W, 256), H - 1, n), 256, 0, st>>>(W, H, img, newVal, maxDiff, labels);
    const size_t N = (size_t)W * H;
    speckle_count_kernel<<<dim3((unsigned)((N + 255) / 256), n), 256, 0, st>>>(W, H, labels, sizes, runlen);
    speckle_apply_kernel<<<dim3(cdiv(W, 256), H, n), 256, 0, st>>>(W, H, img, newVal, maxSize, labels, sizes);
    if (launches) (*launches) += (H > 1) ? 4 : 3;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

int launch_speckle(int n, int W, int H, PlaneS16 img, int newVal, int maxSize, int maxDiff,
                   int32_t *labels, int32_t *sizes, cudaStream_t st, int *launches, int32_t *runlen, const Switches &sw)
{
    return launch_speckle_impl(n, W, H, img, newVal, maxSize, maxDiff, labels, sizes, st, launches, runlen, false, sw);
}

// validateDisparity + valid-rectangle mask (raw, cost -> out), then filterSpeckles on `out` when speckle is set.
// With 16-byte aligned rows and W % 8 == 0 the two row passes are one kernel.
int launch_validate_speckle(int n, int W, int H, int minD, int nd, int d12, int lofs, int W1,
                            int vx0, int vx1, int row0, int row1, PlaneS16 raw, PlaneS16 cost, PlaneS16 out,
                            bool speckle, int newVal, int maxSize, int maxDiff,
                            int32_t *labels, int32_t *sizes, int32_t *runlen, cudaStream_t st, int *launches,
                            void (*after_rows)(void *), void *ctx, const int16_t *spill, const Switches &sw)
{
    if (n <= 0) { if (after_rows) after_rows(ctx); return 0; }
    const auto al16 = [](PlaneS16 p) { return ((reinterpret_cast<uintptr_t>(p.p) | (p.pitch * 2) | (p.frame * 2)) & 15) == 0; };
    if (speckle_vec_ok(W, out, labels, sw) && al16(raw) && al16(cost) && !sw.post_unfused) {
        ValArgs va;
        va.minD = minD; va.nd = nd; va.d12 = d12; va.lofs = lofs; va.W1 = W1; va.vx0 = vx0; va.vx1 = vx1; va.row0 = row0; va.row1 = row1;
        va.raw = raw; va.cost = cost; va.spill = spill;
        const int nt = ((W / 8) + 31) & ~31;
        // rows wider than ~4900 pixels need more than the default 48 KB of dynamic shared memory (W <= 8192: 82 KB)
        if ((size_t)W * 10 + 16 > 48 * 1024)
            RTDM_CUDA(cudaFuncSetAttribute(post_row8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)W * 10 + 16)));
        post_row8_kernel<true><<<dim3(H, n), nt, (size_t)W * 10 + 16, st>>>(W, H, out, va, speckle ? 1 : 0, newVal, maxDiff, labels, sizes, runlen);
        if (launches) (*launches)++;
        RTDM_CUDA(cudaGetLastError());
        if (after_rows) after_rows(ctx);
        if (!speckle) return 0;
        return launch_speckle_impl(n, W, H, out, newVal, maxSize, maxDiff, labels, sizes, st, launches, runlen, true, sw);
    }
    int rc = launch_validate_mask(n, W, H, minD, nd, d12, lofs, W1, vx0, vx1, row0, row1, raw, cost, out, st, launches, spill);
    if (after_rows) after_rows(ctx);
    if (rc || !speckle) return rc;
    return launch_speckle_impl(n, W, H, out, newVal, maxSize, maxDiff, labels, sizes, st, launches, runlen, false, sw);
}

// ------------------------------------------------------------------------------------------------
// 3x3 median on int16 with replicate border (19-exchange network on 9 values)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mm(int &a, int &b) { int t = min(a, b); b = max(a, b); a = t; }

__global__ void __launch_bounds__(256)
median3_kernel(int W, int H, PlaneS16 src, PlaneS16 dst)
{
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int16_t *s = src.p + (size_t)f * src.frame;
    int xm = max(x - 1, 0), xp = min(x + 1, W - 1), ym = max(y - 1, 0), yp = min(y + 1, H - 1);
    const int16_t *r0 = s + (size_t)ym * src.pitch, *r1 = s + (size_t)y * src.pitch, *r2 = s + (size_t)yp * src.pitch;
    int p0 = r0[xm], p1 = r0[x], p2 = r0[xp], p3 = r1[xm], p4 = r1[x], p5 = r1[xp], p6 = r2[xm], p7 = r2[x], p8 = r2[xp];
    mm(p1, p2); mm(p4, p5); mm(p7, p8); mm(p0, p1); mm(p3, p4); mm(p6, p7);
    mm(p1, p2); mm(p4, p5); mm(p7, p8); mm(p0, p3); mm(p5, p8); mm(p4, p7);
    mm(p3, p6); mm(p1, p4); mm(p2, p5); mm(p4, p7); mm(p4, p2); mm(p6, p4); mm(p4, p2);
    dst.p[(size_t)f * dst.frame + (size_t)y * dst.pitch + x] = (int16_t)p4;
}

int launch_median3(int n, int W, int H, PlaneS16 src, PlaneS16 dst, cudaStream_t st, int *launches)
{
    if (n <= 0) return 0;
    median3_kernel<<<dim3(cdiv(W, 256), H, n), 256, 0, st>>>(W, H, src, dst);
    if (launches) (*launches)++;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rtdm
