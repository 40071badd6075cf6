// rectify.cu -- rectification front-end, the step right before the matcher:
//   cvtColor(img, gray, CV_RGB2GRAY)                              estimator.cpp:29-30
//   remap(gray, rect, map1, map2, INTER_LINEAR); rect = rect(roif) estimator.cpp:32-36 (maps: main.cpp:95-96)
// fused: one thread per pixel of the ROI reads its map entry, converts the (up to) four source pixels it needs
// from RGB to gray on the fly and blends them with OpenCV's fixed-point bilinear weights.  Integer arithmetic,
// bit-exact (restated in oracle/oracle.py: rgb2gray, remap_linear_fixed).
#include "common.cuh"

namespace rtdm {
namespace {

__device__ __forceinline__ int gray_at(const uint8_t *img, size_t pitch, int W, int H, int x, int y)
{
    if ((unsigned)x >= (unsigned)W || (unsigned)y >= (unsigned)H) return 0;           // BORDER_CONSTANT, value 0
    const uint8_t *p = img + (size_t)y * pitch + 3 * (size_t)x;
    return ((int)p[0] * 9798 + (int)p[1] * 19235 + (int)p[2] * 3735 + (1 << 14)) >> 15;
}

// maps: the ROI part only, [roi_h][roi_w] of (sx, sy) int16 pairs and fy << 5 | fx uint16
__global__ void __launch_bounds__(256)
rectify_kernel(const uint8_t *rgb, size_t pitch, size_t frame, int W, int H, const short2 *map1, const uint16_t *map2,
               int rw, int rh, uint8_t *out, size_t opitch, size_t oframe)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= rw) return;
    const short2 s = map1[(size_t)y * rw + x];
    const int m = map2[(size_t)y * rw + x];
    const int fx = m & 31, fy = (m >> 5) & 31;
    const uint8_t *img = rgb + (size_t)f * frame;
    const int g00 = gray_at(img, pitch, W, H, s.x, s.y), g01 = gray_at(img, pitch, W, H, s.x + 1, s.y);
    const int g10 = gray_at(img, pitch, W, H, s.x, s.y + 1), g11 = gray_at(img, pitch, W, H, s.x + 1, s.y + 1);
    const int v = ((32 - fy) * (32 - fx) * 32) * g00 + ((32 - fy) * fx * 32) * g01 + (fy * (32 - fx) * 32) * g10 + (fy * fx * 32) * g11;
    out[(size_t)f * oframe + (size_t)y * opitch + x] = (uint8_t)((v + (1 << 14)) >> 15);
}

}  // namespace

int launch_rectify(int n, const uint8_t *rgb, size_t pitch, size_t frame, int W, int H, const int16_t *map1, const uint16_t *map2,
                   int rw, int rh, uint8_t *out, size_t opitch, size_t oframe, cudaStream_t st, int *launches)
{
    if (n <= 0 || rw <= 0 || rh <= 0) return 0;
    rectify_kernel<<<dim3(cdiv(rw, 256), rh, n), 256, 0, st>>>(rgb, pitch, frame, W, H, reinterpret_cast<const short2 *>(map1), map2,
                                                               rw, rh, out, opitch, oframe);
    if (launches) (*launches)++;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rtdm
