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
    se->nrun = 0;
    for (int i = 0; i < 32; i++) { se->rowrun[i] = -1; se->rj1[i] = 0; se->rL[i] = 0; }
    for (int i = 0; i < kh; i++) {
        const int L = se->j2[i] - se->j1[i];
        if (L <= 0) continue;
        int u = 0;
        while (u < se->nrun && !(se->rj1[u] == se->j1[i] && se->rL[u] == L)) u++;
        if (u == se->nrun) { se->rj1[u] = se->j1[i]; se->rL[u] = L; se->nrun++; }
        se->rowrun[i] = u;
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
    const uint8_t *s = src.p + (size_t)f * src.frame;
    const uint32_t flip = op ? 0xFFu : 0u;            // dilate: work on complemented values
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 word-columns x 8 rows per pass
    const int tiles_x = (W + TOW - 1) / TOW, tiles = tiles_x * ((H + TOH - 1) / TOH);
    // bounded grid: each CTA walks over tiles (frames that took the fast path cost only a few empty CTAs)
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int ox = (tile % tiles_x) * TOW, oy = (tile / tiles_x) * TOH;
    __syncthreads();                                   // the previous tile's readers are done with the smem
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
    }   // tile loop
}

// ------------------------------------------------------------------------------------------------
// Binary fast path for SWMorphologicalFilter::run: erode, dilate, dilate, erode on {0,255} masks (what
// inRange() feeds the filter, estimator.cpp:43).  One CTA packs a tile (+20 px halo) to 1 bit/pixel (one thread
// per 32-pixel word, 128-bit loads), runs all four passes on 32-pixel words in shared memory (a run of the structuring element is
// an AND of two shifted power-of-two run words; dilate = erode of the complement) and unpacks the centre.
// A frame that contains any other byte value raises its flag and is recomputed by the generic kernels.
// ------------------------------------------------------------------------------------------------
constexpr int BT_W = 7;                          // output tile: 7 words (224 px) x BT_H rows (32: few frames, 90: batches)
constexpr int BT_RW = BT_W + 2, BT_HALO = 20;

constexpr int BT_MAXRUN = 6;                     // distinct runs the fast path keeps in shared memory

// structuring-element tables: either the kernel argument (generic) or compile-time constants for the
// reference's 10x10 ellipse (rows 1,7,9,10,10,10,10,10,9,7 wide), which lets the
// compiler fold every shift amount and unroll both loops
struct SeView {
    int kh, ay, ax, nrun;
    const int *rj1, *rL, *rowrun;
};

template <bool FIXED10, int BT_H>
__global__ void __launch_bounds__(256)
morph_binary_openclose_kernel(int W, int H, PlaneU8 src, PlaneU8W dst, MorphSE se, int *nonbinary, int vec)
{
    constexpr int BT_RR = BT_H + 2 * BT_HALO, NWORD = BT_RR * BT_RW;
    __shared__ uint32_t buf[NWORD], inside[NWORD], runs[FIXED10 ? 4 : BT_MAXRUN][NWORD];
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * BT_W * 32 - 32, y0 = blockIdx.y * BT_H - BT_HALO;   // region origin (x0 % 32 == 0)
    const uint8_t *s = src.p + (size_t)f * src.frame;
    const int lane = threadIdx.x & 31;
    bool bad = false;
    // pack: one thread per 32-pixel word.  16-byte aligned planes: two 128-bit loads, the low bit of every byte
    // gathered by a multiply ((t & 0x01010101) * 0x01020408 >> 24 = 4 pixels -> 4 bits); otherwise byte loads.
    for (int i = threadIdx.x; i < NWORD; i += 256) {
        const int r = i / BT_RW, w = i - r * BT_RW;
        const int gx0 = x0 + 32 * w, gy = y0 + r;
        uint32_t bits = 0u, ins = 0u;
        if (gy >= 0 && gy < H && gx0 >= 0 && gx0 < W) {
            const int nvalid = min(32, W - gx0);
            ins = nvalid == 32 ? 0xFFFFFFFFu : (1u << nvalid) - 1u;
            const uint8_t *p = s + (size_t)gy * src.pitch + gx0;
            if (vec && nvalid == 32) {
                const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p)), b = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
                const uint32_t q[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t t = q[k] & 0x01010101u;
                    bad = bad || (q[k] != t * 255u);
                    bits |= ((t * 0x01020408u) >> 24) << (4 * k);
                }
            } else {
                for (int k = 0; k < nvalid; k++) {
                    const int v = p[k];
                    bits |= (uint32_t)(v == 255) << k;
                    bad = bad || (v != 0 && v != 255);
                }
            }
        }
        buf[i] = bits; inside[i] = ins;
    }
    if (__any_sync(0xFFFFFFFFu, bad) && lane == 0) atomicOr(&nonbinary[f], 1);
    __syncthreads();
    constexpr int F_KH = 10, F_A = 5, F_NRUN = 4;
    constexpr int f_rj1[4] = {5, 2, 1, 0}, f_rL[4] = {1, 7, 9, 10}, f_rowrun[10] = {0, 1, 2, 3, 3, 3, 3, 3, 2, 1};
    const int kh = FIXED10 ? F_KH : se.kh, ay = FIXED10 ? F_A : se.ay, ax = FIXED10 ? F_A : se.ax;
    const int nrun = FIXED10 ? F_NRUN : se.nrun;
    const int reach = max(max(ay, kh - 1 - ay), max(ax, (FIXED10 ? 10 : se.kw) - 1 - ax));   // <= BT_HALO / 4
    // pass 1 erode (direct polarity), passes 2+3 dilate (complement polarity), pass 4 erode (direct).
    // In every pass pixels outside the image are the identity of an erode in the current polarity: 1.
    // Pass p only has to be right within (3 - p) * reach rows of the output tile.
    for (int pass = 0; pass < 4; pass++) {
        const bool flip = (pass == 1 || pass == 3);                // polarity changes before passes 2 and 4
        const int mo = (3 - pass) * reach;                         // margin of the rows this pass must produce
        const int ro0 = BT_HALO - mo, ro1 = BT_HALO + BT_H + mo;   // output rows of this pass
        const int ri0 = max(ro0 - reach, 0), ri1 = min(ro1 + reach, BT_RR);   // input rows it reads
        for (int i = ri0 * BT_RW + threadIdx.x; i < ri1 * BT_RW; i += 256) buf[i] = (flip ? ~buf[i] : buf[i]) | ~inside[i];
        __syncthreads();
        // phase A: per (row, word) the AND over every distinct horizontal run, already aligned to the output x
        for (int i = ri0 * BT_RW + threadIdx.x; i < ri1 * BT_RW; i += 256) {
            const int r = i / BT_RW, w = i - r * BT_RW;
            const uint32_t *row = buf + r * BT_RW;
            const uint32_t prev = w > 0 ? row[w - 1] : 0xFFFFFFFFu, cur = row[w], next = w + 1 < BT_RW ? row[w + 1] : 0xFFFFFFFFu;
            // A1 bit j = pixel (32*w + j - 16): 16 pixels of left context
            const unsigned long long A1 = (unsigned long long)(prev >> 16) | ((unsigned long long)cur << 16) | ((unsigned long long)next << 48);
            const unsigned long long A2 = A1 & (A1 >> 1), A4 = A2 & (A2 >> 2), A8 = A4 & (A4 >> 4), A16 = A8 & (A8 >> 8);
#pragma unroll
            for (int u = 0; u < (FIXED10 ? F_NRUN : BT_MAXRUN); u++) {
                if (u < nrun) {
                    const int L = FIXED10 ? f_rL[u] : se.rL[u], s0 = (FIXED10 ? f_rj1[u] : se.rj1[u]) - ax + 16;
                    const int p = L >= 16 ? 16 : (L >= 8 ? 8 : (L >= 4 ? 4 : (L >= 2 ? 2 : 1)));
                    const unsigned long long A = p == 16 ? A16 : (p == 8 ? A8 : (p == 4 ? A4 : (p == 2 ? A2 : A1)));
                    runs[u][i] = (uint32_t)(A >> s0) & (uint32_t)(A >> (s0 + L - p));
                }
            }
        }
        __syncthreads();
        // phase B: AND over the rows of the structuring element
        for (int i = ro0 * BT_RW + threadIdx.x; i < ro1 * BT_RW; i += 256) {
            const int r = i / BT_RW, w = i - r * BT_RW;
            uint32_t acc = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < (FIXED10 ? F_KH : 16); k++) {
                if (k < kh) {
                    const int u = FIXED10 ? f_rowrun[k] : se.rowrun[k], rr = r + k - ay;
                    if (u >= 0 && rr >= 0 && rr < BT_RR) acc &= runs[u][rr * BT_RW + w];
                }
            }
            buf[i] = acc;
        }
        __syncthreads();
    }
    // unpack the centre of the tile: words 1..BT_W, rows BT_HALO..BT_HALO+BT_H; one thread per word
    uint8_t *d = dst.p + (size_t)f * dst.frame;
    for (int i = threadIdx.x; i < BT_H * BT_W; i += 256) {
        const int r = BT_HALO + i / BT_W, w = 1 + i % BT_W;
        const int gx0 = x0 + 32 * w, gy = y0 + r;
        if (gy >= H || gx0 >= W) continue;
        const uint32_t bits = buf[r * BT_RW + w];
        uint8_t *o = d + (size_t)gy * dst.pitch + gx0;
        if (vec && gx0 + 32 <= W) {
            uint32_t q[8];
#pragma unroll
            for (int k = 0; k < 8; k++) q[k] = ((((bits >> (4 * k)) & 0xFu) * 0x00204081u) & 0x01010101u) * 255u;   // 4 bits -> 4 bytes
            reinterpret_cast<uint4 *>(o)[0] = make_uint4(q[0], q[1], q[2], q[3]);
            reinterpret_cast<uint4 *>(o)[1] = make_uint4(q[4], q[5], q[6], q[7]);
        } else {
            for (int k = 0; k < 32 && gx0 + k < W; k++) o[k] = ((bits >> k) & 1u) ? 255 : 0;
        }
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
    const int tiles = cdiv(W, TOW) * cdiv(H, TOH);
    dim3 grid(std::min(tiles, std::max(32, 1184 / n)), 1, n);
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
    const uint8_t *s0 = src.p, *s1 = src.p + (size_t)n * src.frame, *d0 = dst.p, *d1 = dst.p + (size_t)n * dst.frame;
    const bool alias = !(s1 <= d0 || d1 <= s0);
    const bool bitpath = se.kw <= 16 && se.kh <= 16 && se.nrun <= BT_MAXRUN;            // 16 px of context per word side, halo 20 >= 4*ax
    if (bitpath && 4 * std::max(std::max(se.ax, se.kw - 1 - se.ax), std::max(se.ay, se.kh - 1 - se.ay)) <= BT_HALO) {
        RTDM_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * n, st));
        // without aliasing the fast path writes dst directly (flagged frames are overwritten by the generic chain)
        if (!alias) fast = dst;
        // the reference's 10x10 ellipse gets the compile-time specialisation
        static const int f_rj1[4] = {5, 2, 1, 0}, f_rL[4] = {1, 7, 9, 10}, f_rowrun[10] = {0, 1, 2, 3, 3, 3, 3, 3, 2, 1};
        bool fixed10 = se.kw == 10 && se.kh == 10 && se.ax == 5 && se.ay == 5 && se.nrun == 4;
        for (int u = 0; fixed10 && u < 4; u++) fixed10 = se.rj1[u] == f_rj1[u] && se.rL[u] == f_rL[u];
        for (int k = 0; fixed10 && k < 10; k++) fixed10 = se.rowrun[k] == f_rowrun[k];
        const auto al16 = [](const void *p, size_t pitch, size_t frame) { return ((reinterpret_cast<uintptr_t>(p) | pitch | frame) & 15) == 0; };
        const int vec = al16(src.p, src.pitch, src.frame) && al16(fast.p, fast.pitch, fast.frame);
        // tall tiles (less halo work) once there are enough frames to fill the machine, short ones for latency
        if (n >= 4) {
            const dim3 grid(cdiv(W, BT_W * 32), cdiv(H, 90), n);
            if (fixed10) morph_binary_openclose_kernel<true, 90><<<grid, 256, 0, st>>>(W, H, src, fast, se, flags, vec);
            else morph_binary_openclose_kernel<false, 72><<<dim3(grid.x, cdiv(H, 72), n), 256, 0, st>>>(W, H, src, fast, se, flags, vec);
        } else {
            const dim3 grid(cdiv(W, BT_W * 32), cdiv(H, 32), n);
            if (fixed10) morph_binary_openclose_kernel<true, 32><<<grid, 256, 0, st>>>(W, H, src, fast, se, flags, vec);
            else morph_binary_openclose_kernel<false, 32><<<grid, 256, 0, st>>>(W, H, src, fast, se, flags, vec);
        }
        if (launches) (*launches)++;
    } else {
        flags = nullptr;                                           // generic chain for every frame
    }
    int rc;
    rc = launch_morph(n, W, H, src, ta, se, 0, st, launches, flags); if (rc) return rc;
    rc = launch_morph(n, W, H, PlaneU8{ta.p, ta.pitch, ta.frame}, tb, se, 1, st, launches, flags); if (rc) return rc;
    rc = launch_morph(n, W, H, PlaneU8{tb.p, tb.pitch, tb.frame}, ta, se, 1, st, launches, flags); if (rc) return rc;
    rc = launch_morph(n, W, H, PlaneU8{ta.p, ta.pitch, ta.frame}, dst, se, 0, st, launches, flags); if (rc) return rc;
    if (flags && alias) {
        morph_select_kernel<<<dim3(cdiv(cdiv(W, 4), 256), H, n), 256, 0, st>>>(W, H, PlaneU8{fast.p, fast.pitch, fast.frame}, dst, flags);
        if (launches) (*launches)++;
    }
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rtdm
