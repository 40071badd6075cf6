// intpeak.cu -- integer-ALU issue-rate microbenchmark (roofline denominator).
// Dependent-free chains (8 independent accumulators per thread) of IADD3, VIMNMX.U16x2 and
// VABSDIFF4.U8; result = lane-operations per second over the whole chip.
#include "common.cuh"

namespace rtdm {
namespace {

template <int OP>
__global__ void __launch_bounds__(256)
intpeak_kernel(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + 1) + i * 0x01010101u;
    uint32_t b = seed ^ 0x00ff00ffu, c = seed + 0x10203040u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) a[i] = a[i] + b + c;                        // IADD3
                else if (OP == 1) a[i] = __vminu2(a[i] + 0u, b) ^ c;     // VIMNMX.U16x2 (+ LOP3 to keep it live)
                else if (OP == 2) a[i] = __vabsdiffu4(a[i], b) + c;      // VABSDIFF4 (+ IADD)
                else if (OP == 3) a[i] = __byte_perm(a[i], b, 0x5140);   // PRMT
                else if (OP == 4) a[i] = __funnelshift_r(a[i], b, 8);    // SHF
                else if (OP == 5) a[i] = __vadd2(a[i], b);               // VIADD.16x2
                else if (OP == 6) a[i] = a[i] * 3u + b;                  // IMAD
                else if (OP == 7) a[i] = __vimin3_u16x2(a[i], b, c);     // VIMNMX3.U16x2
                else if (OP == 8) a[i] = min(a[i], b);                   // VIMNMX.U32
                else if (OP == 9) a[i] = (a[i] & b) ^ c;                 // LOP3
                else if (OP == 10) a[i] = __vimin3_u32(a[i], b, c);      // VIMNMX3.U32
                else if (OP == 11) a[i] = __vminu2(a[i], b);             // VIMNMX.U16x2 alone
                else if (OP == 12) a[i] = __vabsdiffu4(a[i], b);         // VABSDIFF4 alone
                else if (OP == 13) a[i] = __vsub2(a[i], b);              // VIADD.16x2 (sub)
                else a[i] = __viaddmin_s16x2(a[i], b, c);                // VIADDMNMX.S16x2
            }
            b += 0x00010001u;
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
int run_one(uint32_t *buf, int blocks, int iters, double ops_per_inner, double *tiops)
{
    cudaEvent_t e0, e1;
    RTDM_CUDA(cudaEventCreate(&e0));
    RTDM_CUDA(cudaEventCreate(&e1));
    intpeak_kernel<OP><<<blocks, 256>>>(buf, iters / 8, 12345u);        // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        RTDM_CUDA(cudaEventRecord(e0));
        intpeak_kernel<OP><<<blocks, 256>>>(buf, iters, 12345u + rep);
        RTDM_CUDA(cudaEventRecord(e1));
        RTDM_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        RTDM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    double lane_ops = (double)blocks * 256.0 * (double)iters * 64.0 * ops_per_inner;
    *tiops = lane_ops / (best * 1e-3) / 1e12;
    return 0;
}
}  // namespace

int measure_op_rates(int device, double *out, int n)
{
    RTDM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RTDM_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8;
    uint32_t *buf = nullptr;
    RTDM_CUDA(cudaMalloc(&buf, (size_t)blocks * 256 * sizeof(uint32_t)));
    const int iters = 2048;
    double v[15] = {0};
    int rc = 0;
    if (!rc) rc = run_one<0>(buf, blocks, iters, 1.0, &v[0]);
    if (!rc) rc = run_one<1>(buf, blocks, iters, 1.0, &v[1]);
    if (!rc) rc = run_one<2>(buf, blocks, iters, 1.0, &v[2]);
    if (!rc) rc = run_one<3>(buf, blocks, iters, 1.0, &v[3]);
    if (!rc) rc = run_one<4>(buf, blocks, iters, 1.0, &v[4]);
    if (!rc) rc = run_one<5>(buf, blocks, iters, 1.0, &v[5]);
    if (!rc) rc = run_one<6>(buf, blocks, iters, 1.0, &v[6]);
    if (!rc) rc = run_one<7>(buf, blocks, iters, 1.0, &v[7]);
    if (!rc) rc = run_one<8>(buf, blocks, iters, 1.0, &v[8]);
    if (!rc) rc = run_one<9>(buf, blocks, iters, 1.0, &v[9]);
    if (!rc) rc = run_one<10>(buf, blocks, iters, 1.0, &v[10]);
    if (!rc) rc = run_one<11>(buf, blocks, iters, 1.0, &v[11]);
    if (!rc) rc = run_one<12>(buf, blocks, iters, 1.0, &v[12]);
    if (!rc) rc = run_one<13>(buf, blocks, iters, 1.0, &v[13]);
    if (!rc) rc = run_one<14>(buf, blocks, iters, 1.0, &v[14]);
    cudaFree(buf);
    for (int i = 0; i < n && i < 15; i++) out[i] = v[i];
    return rc;
}

int measure_int_peak(int device, double *iadd3, double *vimnmx, double *vabsdiff4, double *mhz)
{
    RTDM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RTDM_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8;
    uint32_t *buf = nullptr;
    RTDM_CUDA(cudaMalloc(&buf, (size_t)blocks * 256 * sizeof(uint32_t)));
    const int iters = 4096;
    int rc = 0;
    double a = 0, b = 0, c = 0;
    // ops_per_inner counts every integer instruction of the inner statement (IADD3 = 1;
    // VIMNMX + LOP3 = 2; VABSDIFF4 + IADD = 2)
    if (!rc) rc = run_one<0>(buf, blocks, iters, 1.0, &a);
    if (!rc) rc = run_one<1>(buf, blocks, iters, 2.0, &b);
    if (!rc) rc = run_one<2>(buf, blocks, iters, 2.0, &c);
    cudaFree(buf);
    if (rc) return rc;
    if (iadd3) *iadd3 = a;
    if (vimnmx) *vimnmx = b;
    if (vabsdiff4) *vabsdiff4 = c;
    if (mhz) *mhz = a * 1e12 / ((double)prop.multiProcessorCount * 64.0) / 1e6;   // clock if IADD3 ran 64 lanes/clk/SM
    return 0;
}

}  // namespace rtdm
