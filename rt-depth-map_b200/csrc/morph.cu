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
#include <algorithm>

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
morph_kernel(int W, int H, PlaneU8 src, PlaneU8W dst, MorphSE se, int op, const int *need)
{
    extern __shared__ __align__(16) uint8_t ms[];
    if (need && need[blockIdx.z] == 0) return;      // the bit-packed fast path already produced this frame
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

// ------------------------------------------------------------------------------------------------
// Binary fast path for SWMorphologicalFilter::run: erode, dilate, dilate, erode on {0,255} masks (what
// inRange() feeds the filter, estimator.cpp:43).  One CTA packs a tile (+20 px halo) to 1 bit/pixel with
// warp ballots, runs all four passes on 32-pixel words in shared memory (a run of the structuring element is
// an AND of two shifted power-of-two run words; dilate = erode of the complement) and unpacks the centre.
// A frame that contains any other byte value raises its flag and is recomputed by the generic kernels.
// ------------------------------------------------------------------------------------------------
constexpr int BT_W = 7, BT_H = 32;               // output tile: 7 words (224 px) x 32 rows
constexpr int BT_RW = BT_W + 2, BT_HALO = 20, BT_RR = BT_H + 2 * BT_HALO;

__device__ __forceinline__ uint32_t bit_erode_word(const uint32_t *buf, int r, int w, const MorphSE &se)
{
    uint32_t acc = 0xFFFFFFFFu;
    for (int k = 0; k < se.kh; k++) {
        const int L = se.j2[k] - se.j1[k];
        if (L <= 0) continue;
        const int rr = r + k - se.ay;
        if (rr < 0 || rr >= BT_RR) continue;                       // outside the region: garbage zone anyway
        const uint32_t *row = buf + rr * BT_RW;
        const uint32_t prev = w > 0 ? row[w - 1] : 0xFFFFFFFFu, cur = row[w], next = w + 1 < BT_RW ? row[w + 1] : 0xFFFFFFFFu;
        // R bit j = pixel (32*w + j - 16): 16 pixels of left context
        const unsigned long long R = (unsigned long long)(prev >> 16) | ((unsigned long long)cur << 16) | ((unsigned long long)next << 48);
        int p = 1;
        unsigned long long A = R;
        while (2 * p <= L) { A &= A >> p; p *= 2; }               // A bit j = AND of R bits j .. j+p-1
        const int s0 = se.j1[k] - se.ax + 16;                      // run start relative to R's origin
        acc &= (uint32_t)(A >> s0) & (uint32_t)(A >> (s0 + L - p));
    }
    return acc;
}

__global__ void __launch_bounds__(256)
morph_binary_openclose_kernel(int W, int H, PlaneU8 src, PlaneU8W dst, MorphSE se, int *nonbinary)
{
    __shared__ uint32_t bufA[BT_RR * BT_RW], bufB[BT_RR * BT_RW], inside[BT_RR * BT_RW];
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * BT_W * 32 - 32, y0 = blockIdx.y * BT_H - BT_HALO;   // region origin
    const uint8_t *s = src.p + (size_t)f * src.frame;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool bad = false;
    // 8 independent byte loads in flight per lane before the ballots (the loop is latency-bound otherwise)
    for (int i0 = warp; i0 < BT_RR * BT_RW; i0 += 64) {
        int v[8];
        bool in[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int i = i0 + 8 * u;
            const int r = i / BT_RW, w = i - r * BT_RW;
            const int gx = x0 + 32 * w + lane, gy = y0 + r;
            in[u] = i < BT_RR * BT_RW && gx >= 0 && gx < W && gy >= 0 && gy < H;
            v[u] = in[u] ? s[(size_t)gy * src.pitch + gx] : 0;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int i = i0 + 8 * u;
            const uint32_t bits = __ballot_sync(0xFFFFFFFFu, v[u] == 255);
            const uint32_t ins = __ballot_sync(0xFFFFFFFFu, in[u]);
            bad = bad || (v[u] != 0 && v[u] != 255);
            if (lane == 0 && i < BT_RR * BT_RW) { bufA[i] = bits; inside[i] = ins; }
        }
    }
    if (__any_sync(0xFFFFFFFFu, bad) && lane == 0) atomicOr(&nonbinary[f], 1);
    __syncthreads();
    // pass 1 erode (direct polarity), passes 2+3 dilate (complement polarity), pass 4 erode (direct).
    // In every pass pixels outside the image are the identity of an erode in the current polarity: 1.
    uint32_t *a = bufA, *b = bufB;
    for (int pass = 0; pass < 4; pass++) {
        const bool flip = (pass == 1 || pass == 3);                // polarity changes before passes 2 and 4
        for (int i = threadIdx.x; i < BT_RR * BT_RW; i += 256) a[i] = (flip ? ~a[i] : a[i]) | ~inside[i];
        __syncthreads();
        for (int i = threadIdx.x; i < BT_RR * BT_RW; i += 256) {
            const int r = i / BT_RW, w = i - r * BT_RW;
            b[i] = bit_erode_word(a, r, w, se);
        }
        __syncthreads();
        uint32_t *t = a; a = b; b = t;
    }
    // unpack the centre of the tile: words 1..BT_W, rows BT_HALO..BT_HALO+BT_H
    uint8_t *d = dst.p + (size_t)f * dst.frame;
    for (int i = warp; i < BT_H * BT_W; i += 8) {
        const int r = BT_HALO + i / BT_W, w = 1 + i % BT_W;
        const int gx = x0 + 32 * w + lane, gy = y0 + r;
        if (gx < W && gy < H) d[(size_t)gy * dst.pitch + gx] = ((a[r * BT_RW + w] >> lane) & 1u) ? 255 : 0;
    }
}

// dst = flag ? dst : fast   (per frame)
__global__ void __launch_bounds__(256)
morph_select_kernel(int W, int H, PlaneU8 fast, PlaneU8W dst, const int *nonbinary)
{
    const int f = blockIdx.z;
    if (nonbinary[f]) return;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x >= W) return;
    const uint8_t *sp = fast.p + (size_t)f * fast.frame + (size_t)y * fast.pitch + x;
    uint8_t *dp = dst.p + (size_t)f * dst.frame + (size_t)y * dst.pitch + x;
    if (x + 3 < W && (((uintptr_t)sp | (uintptr_t)dp) & 3) == 0) *reinterpret_cast<uint32_t *>(dp) = *reinterpret_cast<const uint32_t *>(sp);
    else for (int i = 0; i < 4 && x + i < W; i++) dp[i] = sp[i];
}
}  // namespace

int launch_morph(int n, int W, int H, PlaneU8 src, PlaneU8W dst, const MorphSE &se, int op,
                 cudaStream_t st, int *launches, const int *need)
{
    if (n <= 0) return 0;
    const int TR = TOH + se.kh - 1, TC = TOW + se.kw - 1;
    const int TCP = (TC + SLACK + 3) & ~3;
    size_t smem = (size_t)5 * TR * TCP;
    if (smem > 48 * 1024)
        RTDM_CUDA(cudaFuncSetAttribute(morph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(cdiv(W, TOW), cdiv(H, TOH), n);
    morph_kernel<<<grid, 256, smem, st>>>(W, H, src, dst, se, op, need);
    if (launches) (*launches)++;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

// Open+close of n frames: bit-packed fast path into `fast`, generic chain (only for frames whose flag was
// raised) src -> ta -> tb -> ta -> dst, then dst = fast where the flag stayed clear.
int launch_morph_openclose(int n, int W, int H, PlaneU8 src, PlaneU8W dst, PlaneU8W ta, PlaneU8W tb, PlaneU8W fast,
                           int *flags, const MorphSE &se, cudaStream_t st, int *launches)
{
    if (n <= 0) return 0;
    const bool bitpath = se.kw <= 16 && se.kh <= 16;            // 16 px of context per word side, halo 20 >= 4*ax
    if (bitpath && 4 * std::max(std::max(se.ax, se.kw - 1 - se.ax), std::max(se.ay, se.kh - 1 - se.ay)) <= BT_HALO) {
        RTDM_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * n, st));
        dim3 grid(cdiv(W, BT_W * 32), cdiv(H, BT_H), n);
        morph_binary_openclose_kernel<<<grid, 256, 0, st>>>(W, H, src, fast, se, flags);
        if (launches) (*launches)++;
    } else {
        flags = nullptr;                                           // generic chain for every frame
    }
    int rc;
    rc = launch_morph(n, W, H, src, ta, se, 0, st, launches, flags); if (rc) return rc;
    rc = launch_morph(n, W, H, PlaneU8{ta.p, ta.pitch, ta.frame}, tb, se, 1, st, launches, flags); if (rc) return rc;
    rc = launch_morph(n, W, H, PlaneU8{tb.p, tb.pitch, tb.frame}, ta, se, 1, st, launches, flags); if (rc) return rc;
    rc = launch_morph(n, W, H, PlaneU8{ta.p, ta.pitch, ta.frame}, dst, se, 0, st, launches, flags); if (rc) return rc;
    if (flags) {
        morph_select_kernel<<<dim3(cdiv(cdiv(W, 4), 256), H, n), 256, 0, st>>>(W, H, PlaneU8{fast.p, fast.pitch, fast.frame}, dst, flags);
        if (launches) (*launches)++;
    }
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rtdm
