// morph.cu -- erode / dilate with an elliptical structuring element (row-run decomposition).
//
// Replaces the four cv::erode / cv::dilate calls of SWMorphologicalFilter::run (reference
// filter/mf-sw.cpp:19-28, SE = getStructuringElement(MORPH_ELLIPSE, 10x10), include/filter/mf-sw.h:11-12).
// Semantics per SURVEY.md App. A.5 (oracle: orc_morph): anchor (kw/2, kh/2), out-of-image pixels never
// win (erode: +inf, dilate: -inf).
//
// Every row of an ellipse is ONE horizontal run [j1, j2).  Per tile we build power-of-two horizontal
// running minima (m1, m2, m4, m8, m16) in shared memory; a run of length L is then
// min(m_p[s], m_p[s + L - p]) with p the largest power of two <= L, i.e. two shared loads per SE row
// instead of L.  Dilate runs the same code on complemented bytes (max(a,b) = ~min(~a,~b)).
#include "common.cuh"
#include <math.h>

namespace rtdm {

void make_ellipse(int kw, int kh, MorphSE *se)
{
    se->kw = kw; se->kh = kh; se->ax = kw / 2; se->ay = kh / 2;
    int r = kh / 2, c = kw / 2;
    double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < 32; i++) { se->j1[i] = 0; se->j2[i] = 0; }
    for (int i = 0; i < kh; i++) {
        int dy = i - r;
        if (abs(dy) <= r) {
            int dx = (int)nearbyint(c * sqrt((r * r - dy * dy) * inv_r2));   // cvRound
            se->j1[i] = dx > c ? 0 : c - dx;
            se->j2[i] = c + dx + 1 < kw ? c + dx + 1 : kw;
        }
    }
}

namespace {
constexpr int TOW = 64, TOH = 32;          // output tile (TOW must stay 64: lane mapping below)
constexpr int SLACK = 16;                  // columns of identity padding for the doubling reads

// per-byte unsigned minimum of two packed words
__device__ __forceinline__ uint32_t min4(uint32_t a, uint32_t b) { return __vminu4(a, b); }

// m_{2p}[c] = min(m_p[c], m_p[c + p]) for one tile level, word-wise (p in bytes: 1, 2, 4 or 8).
// The last words of a row read the identity instead of running into the next row.
template <int P>
__device__ __forceinline__ void morph_level(const uint8_t *in, uint8_t *out, int TR, int TCP, int tx, int ty)
{
    const int TCW = TCP / 4;
    for (int r = ty; r < TR; r += 8) {
        const uint32_t *a = reinterpret_cast<const uint32_t *>(in + (size_t)r * TCP);
        uint32_t *o = reinterpret_cast<uint32_t *>(out + (size_t)r * TCP);
        for (int w = tx; w < TCW; w += 32) {
            const uint32_t x = a[w];
            uint32_t y;
            if (P < 4) {
                const uint32_t n = (w + 1 < TCW) ? a[w + 1] : 0xFFFFFFFFu;
                y = __funnelshift_r(x, n, 8 * P);
            } else {
                y = (w + P / 4 < TCW) ? a[w + P / 4] : 0xFFFFFFFFu;
            }
            o[w] = min4(x, y);
        }
    }
}

__global__ void __launch_bounds__(256)
morph_kernel(int W, int H, PlaneU8 src, PlaneU8W dst, MorphSE se, int op)
{
    extern __shared__ __align__(16) uint8_t ms[];
    const int TR = TOH + se.kh - 1, TC = TOW + se.kw - 1;
    const int TCP = (TC + SLACK + 3) & ~3;            // row pitch (bytes), multiple of 4
    const int TCW = TCP / 4;                          // words per row
    const size_t LV = (size_t)TR * TCP;
    uint8_t *m1 = ms, *m2 = m1 + LV, *m4 = m2 + LV, *m8 = m4 + LV, *m16 = m8 + LV;
    const int f = blockIdx.z;
    const int ox = blockIdx.x * TOW, oy = blockIdx.y * TOH;
    const uint8_t *s = src.p + (size_t)f * src.frame;
    const uint32_t flip = op ? 0xFFu : 0u;            // dilate: work on complemented values
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 word-columns x 8 rows per pass
    // level 0: tile with identity (255 after flip) outside the image, one word per thread
    for (int r = ty; r < TR; r += 8) {
        const int gy = oy + r - se.ay;
        const bool rowok = gy >= 0 && gy < H;
        const uint8_t *srow = s + (size_t)(rowok ? gy : 0) * src.pitch;
        for (int w = tx; w < TCW; w += 32) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int c = 4 * w + b, gx = ox + c - se.ax;
                uint32_t px = 0xFFu;
                if (rowok && c < TC && gx >= 0 && gx < W) px = (uint32_t)srow[gx] ^ flip;
                v |= px << (8 * b);
            }
            reinterpret_cast<uint32_t *>(m1 + (size_t)r * TCP)[w] = v;
        }
    }
    __syncthreads();
    morph_level<1>(m1, m2, TR, TCP, tx, ty);
    __syncthreads();
    morph_level<2>(m2, m4, TR, TCP, tx, ty);
    __syncthreads();
    morph_level<4>(m4, m8, TR, TCP, tx, ty);
    __syncthreads();
    morph_level<8>(m8, m16, TR, TCP, tx, ty);
    __syncthreads();
    uint8_t *d = dst.p + (size_t)f * dst.frame;
    // output: thread (tx, ty) -> pixels lx = tx and tx + 32 of rows ty, ty + 8, ...
    for (int ly = ty; ly < TOH; ly += 8) {
        const int gy = oy + ly;
        if (gy >= H) break;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int lx = tx + 32 * half, gx = ox + lx;
            if (gx >= W) continue;
            int acc = 255;
            for (int k = 0; k < se.kh; k++) {
                const int L = se.j2[k] - se.j1[k];
                if (L <= 0) continue;
                const size_t base = (size_t)(ly + k) * TCP + lx + se.j1[k];
                int v;
                if (L >= 16)      v = min(m16[base], m16[base + L - 16]);
                else if (L >= 8)  v = min(m8[base], m8[base + L - 8]);
                else if (L >= 4)  v = min(m4[base], m4[base + L - 4]);
                else if (L >= 2)  v = min(m2[base], m2[base + L - 2]);
                else              v = m1[base];
                acc = min(acc, v);
            }
            d[(size_t)gy * dst.pitch + gx] = (uint8_t)((uint32_t)acc ^ flip);
        }
    }
}
}  // namespace

int launch_morph(int n, int W, int H, PlaneU8 src, PlaneU8W dst, const MorphSE &se, int op,
                 cudaStream_t st, int *launches)
{
    if (n <= 0) return 0;
    const int TR = TOH + se.kh - 1, TC = TOW + se.kw - 1;
    const int TCP = (TC + SLACK + 3) & ~3;
    size_t smem = (size_t)5 * TR * TCP;
    if (smem > 48 * 1024)
        RTDM_CUDA(cudaFuncSetAttribute(morph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(cdiv(W, TOW), cdiv(H, TOH), n);
    morph_kernel<<<grid, 256, smem, st>>>(W, H, src, dst, se, op);
    if (launches) (*launches)++;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rtdm
