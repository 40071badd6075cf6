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
                     PlaneS16 raw, PlaneS16 cost, PlaneS16 out)
{
    extern __shared__ uint32_t vm_smem[];
    uint32_t *key = vm_smem;                                 // [W]
    int16_t *sd = reinterpret_cast<int16_t *>(key + W);      // [W] raw disparity row
    const int y = blockIdx.x, f = blockIdx.y;
    const int INV = (minD - 1) * 16;
    int16_t *orow = out.p + (size_t)f * out.frame + (size_t)y * out.pitch;
    if (y < row0 || y >= row1) {
        for (int x = threadIdx.x; x < W; x += blockDim.x) orow[x] = (int16_t)INV;
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
                         PlaneS16 raw, PlaneS16 cost, PlaneS16 out, cudaStream_t st, int *launches)
{
    if (n <= 0) return 0;
    dim3 grid(H, n);
    size_t smem = (size_t)W * 6 + 8;
    validate_mask_kernel<<<grid, 256, smem, st>>>(W, H, minD, nd, d12, lofs, W1, vx0, vx1, row0, row1,
                                                 raw, cost, out);
    if (launches) (*launches)++;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// filterSpeckles == delete 4-connected components (edges where both pixels != newVal and
// |a - b| <= maxDiff) of at most maxSize pixels.  Union-find over the pixel grid:
//   1. init + row runs : label = index of the first pixel of the pixel's horizontal run
//   2. vertical merge  : union(run root of (x,y), run root of (x,y-1)) where connected
//   3. flatten         : label = root
//   4. count           : sizes[root] += 1 (warp-aggregated)
//   5. apply           : pixels whose component size <= maxSize become newVal
// The component partition is unique, so the result equals OpenCV's flood fill for any scan order.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int32_t *lab, int i)
{
    int p = lab[i];
    while (p != i) { i = p; p = lab[i]; }
    return i;
}

__device__ __forceinline__ void uf_union(int32_t *lab, int a, int b)
{
    while (true) {
        a = uf_find(lab, a);
        b = uf_find(lab, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }          // a > b : hook a under b
        int old = atomicMin(&lab[a], b);
        if (old == a) return;
        a = old;                                         // someone else hooked a; retry with it
    }
}

// one CTA per (row, frame): horizontal runs via a block-wide max-scan of run-start columns
__global__ void __launch_bounds__(256)
speckle_rowruns_kernel(int W, int H, PlaneS16 img, int newVal, int maxDiff, int32_t *labels, int32_t *sizes)
{
    __shared__ int warp_last[8];
    const int y = blockIdx.x, f = blockIdx.y;
    const int16_t *row = img.p + (size_t)f * img.frame + (size_t)y * img.pitch;
    int32_t *lab = labels + ((size_t)f * H + y) * W;
    int32_t *siz = sizes + ((size_t)f * H + y) * W;
    const int chunk = (W + 255) / 256;
    const int xa = threadIdx.x * chunk, xb = min(xa + chunk, W);
    // marker(x) = x where a run starts (or the pixel is invalid), -1 where the run continues
    int run = -1;
    int prev = (xa > 0 && xa < W) ? (int)row[xa - 1] : newVal;
    for (int x = xa; x < xb; x++) {
        int v = row[x];
        bool cont = (v != newVal) && (x > 0) && (prev != newVal) && (abs(prev - v) <= maxDiff);
        if (!cont) run = x;
        prev = v;
        siz[x] = 0;
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
    for (int x = xa; x < xb; x++) {
        int v = row[x];
        bool cont = (v != newVal) && (x > 0) && (prev != newVal) && (abs(prev - v) <= maxDiff);
        if (!cont) run = x;
        prev = v;
        lab[x] = (v == newVal) ? -1 : (int)((size_t)y * W + run);
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
    uf_union(lab, y * W + x, (y - 1) * W + x);
}

__global__ void __launch_bounds__(256)
speckle_count_kernel(int W, int H, int32_t *labels, int32_t *sizes)
{
    const int f = blockIdx.y;
    const size_t N = (size_t)W * H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int32_t *lab = labels + (size_t)f * N;
    int32_t *siz = sizes + (size_t)f * N;
    int root = -1;
    if (i < N && lab[i] >= 0) {
        root = uf_find(lab, (int)i);
        lab[i] = root;      // safe: only shortens paths towards the (final) root
    }
    // warp-aggregated histogram: one atomic per distinct root per warp
    unsigned active = __ballot_sync(0xFFFFFFFFu, root >= 0);
    if (root >= 0) {
        unsigned peers = __match_any_sync(active, root);
        int leader = __ffs(peers) - 1;
        if ((int)(threadIdx.x & 31) == leader) atomicAdd(&siz[root], __popc(peers));
    }
}

__global__ void __launch_bounds__(256)
speckle_apply_kernel(int W, int H, PlaneS16 img, int newVal, int maxSize, const int32_t *labels, const int32_t *sizes)
{
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const size_t N = (size_t)W * H;
    int root = labels[(size_t)f * N + (size_t)y * W + x];
    if (root < 0) return;
    if (sizes[(size_t)f * N + root] <= maxSize)
        img.p[(size_t)f * img.frame + (size_t)y * img.pitch + x] = (int16_t)newVal;
}

int launch_speckle(int n, int W, int H, PlaneS16 img, int newVal, int maxSize, int maxDiff,
                   int32_t *labels, int32_t *sizes, cudaStream_t st, int *launches)
{
    if (n <= 0) return 0;
    speckle_rowruns_kernel<<<dim3(H, n), 256, 0, st>>>(W, H, img, newVal, maxDiff, labels, sizes);
    if (H > 1)
        speckle_vmerge_kernel<<<dim3(cdiv(W, 256), H - 1, n), 256, 0, st>>>(W, H, img, newVal, maxDiff, labels);
    const size_t N = (size_t)W * H;
    speckle_count_kernel<<<dim3((unsigned)((N + 255) / 256), n), 256, 0, st>>>(W, H, labels, sizes);
    speckle_apply_kernel<<<dim3(cdiv(W, 256), H, n), 256, 0, st>>>(W, H, img, newVal, maxSize, labels, sizes);
    if (launches) (*launches) += (H > 1) ? 4 : 3;
    RTDM_CUDA(cudaGetLastError());
    return 0;
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
