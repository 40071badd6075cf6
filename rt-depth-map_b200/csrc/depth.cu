// depth.cu -- depth epilogue on the matcher's output:
//   left_disp /= 16.                                       estimator.cpp:75  (CV_16S, round half to even)
//   reprojectImageTo3D(left_disp, xyz, Q, true, CV_32F)    estimator.cpp:76  (Z = 10000 at the minimum disparity)
//   calc_depth(xyz, ., filter_out, ., obj_boundings, .)    estimator.cpp:77, 206-263 (masked mean Z per rectangle)
// Fused so that only (mean Z, count) per rectangle leaves the GPU; the xyz image is written only on request.
// Arithmetic follows OpenCV's (restated in oracle/oracle.py: reproject_to_3d): double 4x4 product accumulated left
// to right without contraction, narrowed to float, times the double reciprocal of w, narrowed again.
#include "common.cuh"

namespace rtdm {
namespace {

__device__ __forceinline__ int div16_rne(int d)
{
    const int q = d >> 4, r = d & 15;
    return q + (r > 8 ? 1 : 0) + ((r == 8 && (q & 1)) ? 1 : 0);
}

struct DepthQ { double q[16]; };

// row i of Q * (x, y, d, 1): ((Qi0*x + Qi1*y) + Qi2*d) + Qi3, every operation rounded on its own
__device__ __forceinline__ double qrow(const DepthQ &Q, int i, double x, double y, double d)
{
    double s = __dmul_rn(Q.q[4 * i], x);
    s = __dadd_rn(s, __dmul_rn(Q.q[4 * i + 1], y));
    s = __dadd_rn(s, __dmul_rn(Q.q[4 * i + 2], d));
    return __dadd_rn(s, Q.q[4 * i + 3]);
}

__device__ __forceinline__ float reproject_component(const DepthQ &Q, int i, double x, double y, double d, double iw)
{
    return (float)__dmul_rn((double)(float)qrow(Q, i, x, y, d), iw);
}

__global__ void __launch_bounds__(256)
depth_min_kernel(const int16_t *disp, size_t pitch, int W, int H, int *minval)
{
    int m = 0x7FFFFFFF;
    for (int y = blockIdx.x; y < H; y += gridDim.x)
        for (int x = threadIdx.x; x < W; x += blockDim.x) m = min(m, div16_rne(disp[(size_t)y * pitch + x]));
    m = __reduce_min_sync(0xFFFFFFFFu, m);
    if ((threadIdx.x & 31) == 0) atomicMin(minval, m);
}

// grid = (chunks of rows, regions); every block adds its partial (sum Z, count) to the region's accumulators
__global__ void __launch_bounds__(256)
depth_regions_kernel(const int16_t *disp, size_t dpitch, const uint8_t *mask, size_t mpitch, DepthQ Q,
                     const int *rects, const int *minval, double *sums, int *counts)
{
    const int reg = blockIdx.y;
    const int rx = rects[4 * reg], ry = rects[4 * reg + 1], rw = rects[4 * reg + 2], rh = rects[4 * reg + 3];
    const double mind = (double)*minval;
    double acc = 0.0;
    int cnt = 0;
    for (int yy = blockIdx.x; yy < rh; yy += gridDim.x) {
        const int y = ry + yy;
        for (int xx = threadIdx.x; xx < rw; xx += blockDim.x) {
            const int x = rx + xx;
            if (mask && mask[(size_t)y * mpitch + x] == 0) continue;
            const double d = (double)div16_rne(disp[(size_t)y * dpitch + x]);
            if (fabs(d - mind) <= 1.1920928955078125e-7) continue;            // Z = 10000: skipped by calc_depth
            const double iw = __ddiv_rn(1.0, qrow(Q, 3, (double)x, (double)y, d));
            const float z = reproject_component(Q, 2, (double)x, (double)y, d, iw);
            if (fabs((double)z - 1.0e4) < 1.1920928955078125e-7 || fabs((double)z) > 1.0e4) continue;
            acc += (double)z;
            cnt++;
        }
    }
    __shared__ double ssum[8];
    __shared__ int scnt[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
        cnt += __shfl_down_sync(0xFFFFFFFFu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = acc; scnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0; int c = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { s += ssum[w]; c += scnt[w]; }
        if (c) { atomicAdd(&sums[reg], s); atomicAdd(&counts[reg], c); }
    }
}

__global__ void __launch_bounds__(256)
depth_xyz_kernel(const int16_t *disp, size_t dpitch, int W, int H, DepthQ Q, const int *minval, float *xyz, size_t xpitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const double d = (double)div16_rne(disp[(size_t)y * dpitch + x]);
    const double iw = __ddiv_rn(1.0, qrow(Q, 3, (double)x, (double)y, d));
    float *o = xyz + (size_t)y * xpitch + 3 * (size_t)x;
    o[0] = reproject_component(Q, 0, (double)x, (double)y, d, iw);
    o[1] = reproject_component(Q, 1, (double)x, (double)y, d, iw);
    float z = reproject_component(Q, 2, (double)x, (double)y, d, iw);
    if (fabs(d - (double)*minval) <= 1.1920928955078125e-7) z = 10000.f;
    o[2] = z;
}

}  // namespace

// disp, mask, xyz: DEVICE pointers (pitches in elements: int16 / bytes / floats); rects, minval, sums, counts: device scratch
int launch_depth(const int16_t *disp, size_t dpitch, int W, int H, const double *Q, const uint8_t *mask, size_t mpitch,
                 int nregions, const int *rects_dev, int *minval, double *sums, int *counts, float *xyz, size_t xpitch,
                 cudaStream_t st, int *launches)
{
    DepthQ q;
    for (int i = 0; i < 16; i++) q.q[i] = Q[i];
    RTDM_CUDA(cudaMemsetAsync(minval, 0x7F, sizeof(int), st));                 // 0x7F7F7F7F: larger than any int16
    if (nregions > 0) {
        RTDM_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * nregions, st));
        RTDM_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * nregions, st));
    }
    depth_min_kernel<<<std::min(H, 592), 256, 0, st>>>(disp, dpitch, W, H, minval);
    if (nregions > 0)
        depth_regions_kernel<<<dim3(64, nregions), 256, 0, st>>>(disp, dpitch, mask, mpitch, q, rects_dev, minval, sums, counts);
    if (xyz) depth_xyz_kernel<<<dim3(cdiv(W, 256), H), 256, 0, st>>>(disp, dpitch, W, H, q, minval, xyz, xpitch);
    if (launches) (*launches) += 1 + (nregions > 0) + (xyz != nullptr);
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rtdm
