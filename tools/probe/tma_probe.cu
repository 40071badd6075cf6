// Development probe: one CTA stages image rows with cp.async.bulk.tensor exactly as bm_sad4.cu does (same helpers, same
// tensor-map encoding); argv[1] selects how far it goes.  nvcc -gencode arch=compute_100a,code=sm_100a tma_probe.cu -o tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}
__global__ void probe2(const __grid_constant__ CUtensorMap tm2, const CUtensorMap *tmg, const uint32_t *src, int step, int boxL, int x, int y, uint32_t *out)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const uint32_t bar = smem_u32(smem + 4096), dst = smem_u32(smem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (step == 10) mbar_expect_tx(bar, 0);
        if (step == 5) { mbar_expect_tx(bar, 4 * boxL); tma_load_2d(dst, &tm2, x, y, bar); }
        if (step == 6) { mbar_expect_tx(bar, 4 * boxL); tma_load_3d(dst, tmg, x, y, 0, bar); }
        if (step == 8) { mbar_expect_tx(bar, 512); bulk_load(dst, src + 1312 / 4 * y, 512, bar); }
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = reinterpret_cast<uint32_t *>(smem)[i];
}

__global__ void probe(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR, int step, int boxL, int x, int y, uint32_t *out)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const uint32_t bar = smem_u32(smem + 4096), dst = smem_u32(smem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (step >= 2 && threadIdx.x == 64) {
        uint32_t bytes = 0;
        if (step == 2 || step >= 4) bytes += 4 * boxL;
        if (step == 3 || step >= 4) bytes += 128;
        mbar_expect_tx(bar, bytes);
        if (step == 2 || step >= 4) tma_load_3d(dst, &tmL, x, y, 0, bar);
        if (step == 3 || step >= 4) tma_load_3d(dst + 2048, &tmR, x, y, 0, bar);
    }
    if (step >= 2) mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = reinterpret_cast<uint32_t *>(smem)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv)
{
    const int step = argc > 1 ? atoi(argv[1]) : 4;
    const int W = 320, H = 240, boxL = 144;
    const size_t lep = 328, rpp = 352;
    void *p = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    printf("entry point: err %d qr %d ptr %p\n", (int)e, (int)qr, p);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    uint32_t *LE; uint8_t *RP; uint32_t *out;
    cudaMalloc(&LE, lep * H * 4); cudaMalloc(&RP, rpp * H); cudaMalloc(&out, 4096);
    std::vector<uint32_t> hl(lep * H); std::vector<uint8_t> hr(rpp * H);
    for (size_t i = 0; i < hl.size(); i++) hl[i] = (uint32_t)i;
    for (size_t i = 0; i < hr.size(); i++) hr[i] = (uint8_t)(i * 7);
    cudaMemcpy(LE, hl.data(), hl.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(RP, hr.data(), hr.size(), cudaMemcpyHostToDevice);
    CUtensorMap tmL, tmR;
    {
        const cuuint64_t dims[3] = {lep, (cuuint64_t)H, 1}; const cuuint64_t strides[2] = {lep * 4, lep * H * 4};
        const cuuint32_t box[3] = {(cuuint32_t)boxL, 1, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&tmL, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, LE, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode L: %d\n", (int)r);
    }
    {
        const cuuint64_t dims[3] = {rpp, (cuuint64_t)H, 1}; const cuuint64_t strides[2] = {rpp, rpp * H};
        const cuuint32_t box[3] = {128, 1, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&tmR, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, RP, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode R: %d\n", (int)r);
    }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int x = argc > 2 ? atoi(argv[2]) : 57, y = 3;
    if (step >= 5) {
        CUtensorMap tm2, *tmg;
        const cuuint64_t dims[2] = {lep, (cuuint64_t)H}; const cuuint64_t strides[1] = {lep * 4};
        const cuuint32_t box[2] = {(cuuint32_t)boxL, 1}, es[2] = {1, 1};
        CUresult r = enc(&tm2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, LE, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode 2D: %d\n", (int)r);
        cudaMalloc(&tmg, sizeof(CUtensorMap)); cudaMemcpy(tmg, &tmL, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
        probe2<<<1, 128, 64 * 1024>>>(tm2, tmg, LE, step, boxL, x, y, out);
    } else
    probe<<<1, 128, 64 * 1024>>>(tmL, tmR, step, boxL, x, y, out);
    e = cudaDeviceSynchronize();
    printf("step %d: sync err %d (%s)\n", step, (int)e, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<uint32_t> ho(1024);
        cudaMemcpy(ho.data(), out, 4096, cudaMemcpyDeviceToHost);
        printf("L[0..3] = %u %u %u %u (expect %u..)  R word0 = %08x (expect bytes from %u)\n", ho[0], ho[1], ho[2], ho[3], (unsigned)(y * lep + x), ho[512], (unsigned)((y * rpp + x) * 7 & 255));
    }
    return 0;
}
