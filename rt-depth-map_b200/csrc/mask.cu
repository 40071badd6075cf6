// mask.cu -- mask front-end and back-end: the two steps either side of the morphological filter.
//
// front-end, estimator.cpp:38-43, fused into one kernel per frame:
//     remap(img[0], img_rectified, map1, map2, INTER_LINEAR); img_rectified = img_rectified(roif);
//     cvtColor(RGB2BGR); cvtColor(BGR2HSV); inRange(low, high) -> filter_in
//   one thread per ROI pixel: map entry -> the (up to) four RGB source pixels -> OpenCV's fixed-point bilinear blend per
//   channel -> integer HSV (RGB2HSV_b: 12-bit division tables, H in [0, 180)) -> range test -> 0 / 255.
//   The rectified BGR image (the reference displays it) is an optional second output.
//
// back-end, estimator.cpp:47-53 with :164-204:
//     findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) -> boundingRect per top-level contour -> area >= minObjSize
//     -> the rectangle spanning them all (bm->setROI1)
//   No contour is ever traced.  A top-level contour is the outer border of an 8-connected component of non-zero pixels
//   whose surrounding background (4-connected; the image sits in a virtual zero frame) is the frame's own; its
//   boundingRect is the component's bounding box; OpenCV lists them in reverse raster order of their first pixel.
//   So: one union-find labelling pass over both colours (foreground: W, NW, N, NE links; background: W, N links and
//   border pixels to a frame node; roots = smallest pixel index = first pixel in raster order), bounding boxes by
//   atomics from the components' boundary pixels only, an "adjacent to the frame's background" flag, a compaction and
//   a rank sort of the few surviving boxes.  Only (count, boxes, spanning rectangle) leave the GPU.
// Integer work throughout, bit-exact (oracle/oracle.py: color_mask, contour_boxes, object_regions; pinned to cv2 4.13.0).
#include "common.cuh"

namespace rtdm {
namespace {

// ---------------------------------------------------------------------------------------------------------------
// front-end
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int3 rgb_at(const uint8_t *img, size_t pitch, int W, int H, int x, int y)
{
    if ((unsigned)x >= (unsigned)W || (unsigned)y >= (unsigned)H) return make_int3(0, 0, 0);     // BORDER_CONSTANT, value 0
    const uint8_t *p = img + (size_t)y * pitch + 3 * (size_t)x;
    return make_int3(p[0], p[1], p[2]);
}

struct HsvRange { int lo[3], hi[3]; };

__global__ void __launch_bounds__(256)
colormask_kernel(const uint8_t *rgb, size_t pitch, size_t frame, int W, int H, const short2 *map1, const uint16_t *map2,
                 int rw, int rh, HsvRange rg, uint8_t *mask, size_t mpitch, size_t mframe, uint8_t *bgr, size_t bpitch, size_t bframe)
{
    // OpenCV's tables: sdiv[i] = cvRound((255 << 12) / i), hdiv[i] = cvRound((180 << 12) / (6 i)) (round half to even)
    __shared__ int sdiv[256], hdiv[256];
    {
        const int i = threadIdx.x;
        sdiv[i] = i ? __double2int_rn((double)(255 << 12) / (double)i) : 0;
        hdiv[i] = i ? __double2int_rn((double)(180 << 12) / (6.0 * (double)i)) : 0;
    }
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= rw) return;
    const short2 s = map1[(size_t)y * rw + x];
    const int m = map2[(size_t)y * rw + x];
    const int fx = m & 31, fy = (m >> 5) & 31;
    const uint8_t *img = rgb + (size_t)f * frame;
    const int3 p00 = rgb_at(img, pitch, W, H, s.x, s.y), p01 = rgb_at(img, pitch, W, H, s.x + 1, s.y);
    const int3 p10 = rgb_at(img, pitch, W, H, s.x, s.y + 1), p11 = rgb_at(img, pitch, W, H, s.x + 1, s.y + 1);
    const int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
    const int r = (w00 * p00.x + w01 * p01.x + w10 * p10.x + w11 * p11.x + (1 << 14)) >> 15;
    const int g = (w00 * p00.y + w01 * p01.y + w10 * p10.y + w11 * p11.y + (1 << 14)) >> 15;
    const int b = (w00 * p00.z + w01 * p01.z + w10 * p10.z + w11 * p11.z + (1 << 14)) >> 15;
    if (bgr) {
        uint8_t *o = bgr + (size_t)f * bframe + (size_t)y * bpitch + 3 * (size_t)x;
        o[0] = (uint8_t)b; o[1] = (uint8_t)g; o[2] = (uint8_t)r;
    }
    const int v = max(max(b, g), r), diff = v - min(min(b, g), r);
    const int sat = (diff * sdiv[v] + (1 << 11)) >> 12;
    int hh = v == r ? g - b : (v == g ? b - r + 2 * diff : r - g + 4 * diff);
    hh = (hh * hdiv[diff] + (1 << 11)) >> 12;
    if (hh < 0) hh += 180;
    const bool in = hh >= rg.lo[0] && hh <= rg.hi[0] && sat >= rg.lo[1] && sat <= rg.hi[1] && v >= rg.lo[2] && v <= rg.hi[2];
    mask[(size_t)f * mframe + (size_t)y * mpitch + x] = in ? 255 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// back-end: union-find labelling (label = smallest pixel index of the component; node W * H = the frame)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cc_find(const int *L, int a)
{
    int p = L[a];
    while (p != a) { a = p; p = L[a]; }
    return a;
}
__device__ __forceinline__ void cc_union(int *L, int a, int b)
{
    bool done;
    do {
        a = cc_find(L, a); b = cc_find(L, b);
        if (a < b) { const int old = atomicMin(&L[b], a); done = old == b; b = old; }
        else if (b < a) { const int old = atomicMin(&L[a], b); done = old == a; a = old; }
        else done = true;
    } while (!done);
}

__global__ void cc_init_kernel(int *L, int4 *bb, int *ext, int N, int W, int H, int *counters)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= N) L[i] = i;
    if (i < N) { bb[i] = make_int4(W, H, -1, -1); ext[i] = 0; }
    if (i < 2) counters[i] = 0;                                      // boxes kept, top-level contours
    if (i >= 2 && i < 6) counters[i] = i < 4 ? 1000000 : -1000000;    // spanning rectangle: min x, min y, max x, max y
}

__global__ void cc_merge_kernel(const uint8_t *mask, size_t pitch, int W, int H, int *L)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const uint8_t *row = mask + (size_t)y * pitch, *up = row - pitch;
    const int p = y * W + x;
    const bool v = row[x] != 0;
    if (v) {
        if (x > 0 && row[x - 1]) cc_union(L, p, p - 1);
        if (y > 0) {
            if (up[x]) cc_union(L, p, p - W);
            else {                                   // NW and NE are already linked through N when N is set
                if (x > 0 && up[x - 1]) cc_union(L, p, p - W - 1);
                if (x + 1 < W && up[x + 1]) cc_union(L, p, p - W + 1);
            }
        }
    } else {
        if (x > 0 && !row[x - 1]) cc_union(L, p, p - 1);
        if (y > 0 && !up[x]) cc_union(L, p, p - W);
        if (x == 0 || y == 0 || x == W - 1 || y == H - 1) cc_union(L, p, W * H);
    }
}

// bounding boxes from boundary pixels, "touches the frame's background" flag; flattens the foreground labels
__global__ void cc_boxes_kernel(const uint8_t *mask, size_t pitch, int W, int H, int *L, int4 *bb, int *ext)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const uint8_t *row = mask + (size_t)y * pitch;
    if (!row[x]) return;
    const int p = y * W + x;
    const int r = cc_find(L, p);
    L[p] = r;
    const bool l = x > 0 && row[x - 1], rr = x + 1 < W && row[x + 1];
    const bool u = y > 0 && (row - pitch)[x], d = y + 1 < H && (row + pitch)[x];
    if (!l) atomicMin(&bb[r].x, x);
    if (!rr) atomicMax(&bb[r].z, x);
    if (!u) atomicMin(&bb[r].y, y);
    if (!d) atomicMax(&bb[r].w, y);
    if (l && rr && u && d) return;
    // a 4-neighbour outside the image is the frame itself; a background 4-neighbour counts if it hangs on the frame node
    bool e = x == 0 || y == 0 || x == W - 1 || y == H - 1;
    if (!e) {
        const int fr = cc_find(L, W * H);
        e = (!l && cc_find(L, p - 1) == fr) || (!rr && cc_find(L, p + 1) == fr) || (!u && cc_find(L, p - W) == fr) || (!d && cc_find(L, p + W) == fr);
    }
    if (e) ext[r] = 1;
}

// roots of top-level components whose box area reaches minSize -> unordered list of (root index, box)
__global__ void cc_compact_kernel(const uint8_t *mask, size_t pitch, int W, int H, const int *L, const int4 *bb, const int *ext,
                                  int minSize, int maxR, int *counters, int *keys, int4 *boxes)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    if (!mask[(size_t)y * pitch + x]) return;
    const int p = y * W + x;
    if (L[p] != p || !ext[p]) return;
    atomicAdd(&counters[1], 1);                                      // contours.size()
    const int4 b = bb[p];
    const int w = b.z - b.x + 1, h = b.w - b.y + 1;
    if (w * h < minSize) return;
    const int slot = atomicAdd(&counters[0], 1);
    if (slot < maxR) { keys[slot] = p; boxes[slot] = make_int4(b.x, b.y, w, h); }
}

// reverse raster order of the first pixels (OpenCV's contour order) by rank counting, and the spanning rectangle
// out: [0] = boxes kept, [1] = top-level contours, [2..5] = min x, min y, max x, max y of the kept boxes, then the boxes
__global__ void cc_sort_kernel(const int *keys, const int4 *boxes, int maxR, int *out)
{
    const int n = min(out[0], maxR);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = keys[i];
    int rank = 0;
    for (int j = 0; j < n; j++) rank += keys[j] > k;
    const int4 b = boxes[i];
    out[6 + 4 * rank] = b.x; out[7 + 4 * rank] = b.y; out[8 + 4 * rank] = b.z; out[9 + 4 * rank] = b.w;
    atomicMin(&out[2], b.x); atomicMin(&out[3], b.y); atomicMax(&out[4], b.x + b.z); atomicMax(&out[5], b.y + b.w);
}

}  // namespace

int launch_colormask(int n, const uint8_t *rgb, size_t pitch, size_t frame, int W, int H, const int16_t *map1, const uint16_t *map2,
                     int rw, int rh, const int *lo, const int *hi, uint8_t *mask, size_t mpitch, size_t mframe,
                     uint8_t *bgr, size_t bpitch, size_t bframe, cudaStream_t st, int *launches)
{
    if (n <= 0 || rw <= 0 || rh <= 0) return 0;
    HsvRange rg;
    for (int c = 0; c < 3; c++) { rg.lo[c] = lo[c]; rg.hi[c] = hi[c]; }
    colormask_kernel<<<dim3(cdiv(rw, 256), rh, n), 256, 0, st>>>(rgb, pitch, frame, W, H, reinterpret_cast<const short2 *>(map1), map2,
                                                                 rw, rh, rg, mask, mpitch, mframe, bgr, bpitch, bframe);
    if (launches) (*launches)++;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

// work: labels (W*H + 1 ints), bb (W*H int4), ext (W*H ints), keys (maxR ints), boxes (maxR int4);
// out: 6 + 4 * maxR ints on the device (see cc_sort_kernel)
int launch_regions(const uint8_t *mask, size_t pitch, int W, int H, int minSize, int maxR,
                   int *labels, void *bb, int *ext, int *keys, void *boxes, int *out, cudaStream_t st, int *launches)
{
    int *counters = out;
    const int N = W * H;
    const dim3 grid(cdiv(W, 128), H);
    cc_init_kernel<<<cdiv(N + 1, 256), 256, 0, st>>>(labels, static_cast<int4 *>(bb), ext, N, W, H, counters);
    cc_merge_kernel<<<grid, 128, 0, st>>>(mask, pitch, W, H, labels);
    cc_boxes_kernel<<<grid, 128, 0, st>>>(mask, pitch, W, H, labels, static_cast<int4 *>(bb), ext);
    cc_compact_kernel<<<grid, 128, 0, st>>>(mask, pitch, W, H, labels, static_cast<const int4 *>(bb), ext, minSize, maxR, counters, keys,
                                            static_cast<int4 *>(boxes));
    cc_sort_kernel<<<cdiv(maxR, 256), 256, 0, st>>>(keys, static_cast<const int4 *>(boxes), maxR, out);
    if (launches) (*launches) += 5;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rtdm
