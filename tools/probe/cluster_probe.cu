// Development probe for the SGBM pass kernel's neighbour exchange: a chain of CTAs in one thread-block cluster hands a
// boundary column (LPC x 16 bytes + a minimum) to both neighbours EVERY row through distributed shared memory
// (st.shared::cluster + remote mbarrier.arrive.release.cluster, consumer mbarrier.try_wait.parity.acquire.cluster), with
// one CTA barrier per row as in the sweep.  Prints, per cluster size: how many clusters fit the GPU, the cycles per row
// (exchange latency on the critical path) and the number of wrong values seen.  Every wait is bounded.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 cluster_probe.cu -o cluster_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ void st_cluster_v4(uint32_t a, uint4 v) { asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }

constexpr int LPC = 8;

struct Sm {
    uint4 padL[2][LPC], padR[2][LPC];       // written by the left / right neighbour
    unsigned long long fullL[2], fullR[2];
};

__global__ void __launch_bounds__(1024, 1) chain_kernel(int rows, int work, unsigned long long *cycles, int *errors, int *timeouts)
{
    extern __shared__ __align__(16) uint8_t raw[];
    Sm *s = reinterpret_cast<Sm *>(raw);
    const uint32_t rank = cluster_rank(), cs = cluster_size();
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) {
        for (int b = 0; b < 2; b++) { mbar_init(smem_u32(&s->fullL[b]), LPC); mbar_init(smem_u32(&s->fullR[b]), LPC); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_sync();
    const bool isL = tid < LPC;                       // first column: receives from the left, sends to the left
    const bool isR = tid >= nthr - LPC;               // last column: receives from the right, sends to the right
    const int sl = isL ? tid : tid - (nthr - LPC);
    const bool hasL = rank > 0, hasR = rank + 1 < cs;
    // my sends to the right neighbour land in ITS padL, arrive on ITS fullL; to the left neighbour: padR / fullR
    uint32_t dstR[2], barR[2], dstL[2], barL[2];
    for (int b = 0; b < 2; b++) {
        dstR[b] = hasR ? mapa(smem_u32(&s->padL[b][sl]), rank + 1) : 0u; barR[b] = hasR ? mapa(smem_u32(&s->fullL[b]), rank + 1) : 0u;
        dstL[b] = hasL ? mapa(smem_u32(&s->padR[b][sl]), rank - 1) : 0u; barL[b] = hasL ? mapa(smem_u32(&s->fullR[b]), rank - 1) : 0u;
    }
    int err = 0, tmo = 0;
    uint32_t acc = tid;
    const long long t0 = clock64();
    for (int t = 0; t < rows; t++) {
        const int b = t & 1;
        if (t > 0) {
            // row t consumes what the neighbours sent at their row t - 1: phase number (t - 1) / 2 ... of barrier b
            const uint32_t ph = (uint32_t)(((t - (b ? 1 : 2)) >> 1) & 1);
            if (isL && hasL) {
                int spin = 0;
                while (!mbar_try(smem_u32(&s->fullL[b]), ph)) if (++spin > (1 << 20)) { tmo++; break; }
                const uint4 v = s->padL[b][sl];
                if (v.x != (rank - 1) * 1000003u + (uint32_t)(t - 1) * 17u + (uint32_t)sl || v.w != ~v.x) err++;
            }
            if (isR && hasR) {
                int spin = 0;
                while (!mbar_try(smem_u32(&s->fullR[b]), ph)) if (++spin > (1 << 20)) { tmo++; break; }
                const uint4 v = s->padR[b][sl];
                if (v.x != (rank + 1) * 1000003u + (uint32_t)(t - 1) * 17u + (uint32_t)sl + 7u || v.w != ~v.x) err++;
            }
        }
        for (int i = 0; i < work; i++) acc = acc * 1664525u + 1013904223u;       // stands in for the row's arithmetic
        if (t + 1 < rows) {
            if (isR && hasR) { const uint32_t x = rank * 1000003u + (uint32_t)t * 17u + (uint32_t)sl; st_cluster_v4(dstR[b ^ 1], make_uint4(x, acc, 0u, ~x)); mbar_arrive_cluster(barR[b ^ 1]); }
            if (isL && hasL) { const uint32_t x = rank * 1000003u + (uint32_t)t * 17u + (uint32_t)sl + 7u; st_cluster_v4(dstL[b ^ 1], make_uint4(x, acc, 0u, ~x)); mbar_arrive_cluster(barL[b ^ 1]); }
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    cluster_sync();
    if (err) atomicAdd(errors, err);
    if (tmo) atomicAdd(timeouts, tmo);
    if (tid == 0 && blockIdx.x == 0) *cycles = (unsigned long long)(t1 - t0);
    if (acc == 0xDEADBEEFu) *errors = -1;
}

int main(int argc, char **argv)
{
    const int rows = argc > 1 ? atoi(argv[1]) : 720 * 4;
    unsigned long long *cyc; int *err, *tmo;
    cudaMalloc(&cyc, 8); cudaMalloc(&err, 4); cudaMalloc(&tmo, 4);
    const size_t smem = 150 * 1024;                     // one CTA per SM, like the pass kernel
    cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
    printf("cluster,threads,work,max_active_clusters,grid,cycles_per_row,errors,timeouts,status\n");
    const int sizes[] = {1, 2, 4, 6, 8, 9, 10, 12, 16};
    for (int nthr : {1024, 576}) for (int work : {0, 400}) for (int cs : sizes) {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.blockDim = dim3(nthr); cfg.dynamicSmemBytes = smem; cfg.attrs = at; cfg.numAttrs = 1;
        cfg.gridDim = dim3(cs);
        int ncl = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, chain_kernel, &cfg);
        if (e != cudaSuccess) { printf("%d,%d,%d,0,0,0,0,0,occupancy: %s\n", cs, nthr, work, cudaGetErrorString(e)); cudaGetLastError(); continue; }
        if (ncl < 1) { printf("%d,%d,%d,0,0,0,0,0,no cluster fits\n", cs, nthr, work); continue; }
        cfg.gridDim = dim3(cs * ncl);
        cudaMemset(cyc, 0, 8); cudaMemset(err, 0, 4); cudaMemset(tmo, 0, 4);
        e = cudaLaunchKernelEx(&cfg, chain_kernel, rows, work, cyc, err, tmo);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        unsigned long long c = 0; int he = 0, ht = 0;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&he, err, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&ht, tmo, 4, cudaMemcpyDeviceToHost);
        printf("%d,%d,%d,%d,%d,%.1f,%d,%d,%s\n", cs, nthr, work, ncl, cs * ncl, (double)c / rows, he, ht, cudaGetErrorString(e));
        cudaGetLastError();
    }
    return 0;
}
