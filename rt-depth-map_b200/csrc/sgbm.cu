// sgbm.cu -- semi-global matching: Birchfield-Tomasi cost, box aggregation, 5/8-path min-plus
// aggregation over 16-bit costs, winner-take-all with uniqueness / sub-pixel / left-right check.
//
// Replaces computeDisparitySGBM inside cv::StereoSGBM::compute as reached from
// SWSemiGlobalMatcher::compute (reference stereo-matcher/sgbm-sw.cpp:32-37).  Arithmetic per
// SURVEY.md App. A.6 (restated and pinned in oracle/sgbm_oracle.c).
//
// Pipeline per batch of frames (volumes are [frame][y][x1][d] uint16, x1 in [0, W1)).  D = 48 / 64 / 96 / 128 / 192 (the fast path):
//   sgbm_planes2_kernel     : per pixel (value, lo, hi) of the x-Sobel plane and the raw plane in the cost kernel's staging format
//   sgbm_cost_fused_kernel  : BT pixel cost -> horizontal window -> vertical window + P2 -> C, rows staged by cp.async
//   sgbm_path4_kernel<.,.,0>: left-to-right path, S = L (8 / 16 / 32 lanes per chain, 8 or 6 disparities per lane, L in registers)
//   sgbm_vpass_kernel       : the three paths that come from the previous row, all H rows in one persistent launch, one
//                             thread-block cluster per frame, boundary columns exchanged through distributed shared memory
//                             (batches); sgbm_sweep_kernel: the same three paths in 8-row tiles with halo columns and a
//                             frontier buffer (single frames, small batches at D = 64 / 128; per-direction chains at the other
//                             D).  Once per direction (MODE_HH: down and up)
//   sgbm_path4_kernel<.,.,2>: right-to-left path with S + L kept in registers and the winner-take-all taken right there
//   sgbm_lr_kernel          : disp2 by atomicMin (OpenCV's right-to-left strict '>' scan == min over (cost, -x)), sub-pixel, LR check
// Other D (16 .. 256, % 16): sgbm_planes / sgbm_cost_hsum / sgbm_vsum (D = 256: the fused cost kernel), one sgbm_path_kernel
// launch per direction (one warp per chain), sgbm_wta_kernel.  Then launch_median3 and launch_speckle (postproc.cu).
// Every pixel belongs to exactly one chain per direction, so the S read-modify-write needs no atomics.
#include "common.cuh"
#include <stdlib.h>
#include <algorithm>
#include <type_traits>

namespace rtdm {
namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ------------------------------------------------------------------------------------------------
// planes: out[img][plane][comp][y][x], comp 0 = value, 1 = lo, 2 = hi (half-pixel interval)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int bt_plane_value(const uint8_t *img, size_t pitch, int W, int H, int x, int y, int plane, int ftzero)
{
    if (x <= 0 || x >= W - 1) return ftzero;
    const uint8_t *r = img + (size_t)y * pitch;
    if (plane) return r[x];
    const uint8_t *rn = y > 0 ? r - pitch : r, *rs = y < H - 1 ? r + pitch : r;
    int g = 2 * ((int)r[x + 1] - (int)r[x - 1]) + ((int)rn[x + 1] - (int)rn[x - 1]) + ((int)rs[x + 1] - (int)rs[x - 1]);
    return clampi(g, -ftzero, ftzero) + ftzero;
}

__global__ void __launch_bounds__(256)
sgbm_planes_kernel(PlaneU8 left, PlaneU8 right, uint8_t *planes, size_t frame_planes, int W, int H, int Wp, int ftzero)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int f = blockIdx.z >> 1, img = blockIdx.z & 1;
    if (x >= W) return;
    const uint8_t *src = img ? right.p + (size_t)f * right.frame : left.p + (size_t)f * left.frame;
    const size_t sp = img ? right.pitch : left.pitch;
    uint8_t *out = planes + (size_t)f * frame_planes + (size_t)img * 6 * H * Wp;
#pragma unroll
    for (int pl = 0; pl < 2; pl++) {
        int v = bt_plane_value(src, sp, W, H, x, y, pl, ftzero);
        int a = x > 0 ? (v + bt_plane_value(src, sp, W, H, x - 1, y, pl, ftzero)) / 2 : v;
        int b = x < W - 1 ? (v + bt_plane_value(src, sp, W, H, x + 1, y, pl, ftzero)) / 2 : v;
        uint8_t *o = out + (size_t)(pl * 3) * H * Wp + (size_t)y * Wp + x;
        o[0] = (uint8_t)v;
        o[(size_t)H * Wp] = (uint8_t)min(min(a, b), v);
        o[(size_t)2 * H * Wp] = (uint8_t)max(max(a, b), v);
    }
}

// ------------------------------------------------------------------------------------------------
// BT cost + horizontal window.  CTA = (tile of TX cost columns, row y, frame).
// ------------------------------------------------------------------------------------------------
constexpr int TX = 32;

struct CostArgs {
    const uint8_t *planes; size_t frame_planes;
    uint16_t *Hs; size_t frame_vol;
    int W, H, Wp, D, minD, h, minX1, W1;
};

// Packed version: two disparities per 32-bit word (s16x2 / u16x2 SIMD: VIADD.16x2, VIMNMX3.S16x2).  The right
// image's (value, lo, hi) arrays are staged REVERSED as u16, so that increasing d is increasing address, in two
// copies (offset by one element) so that every (d, d+1) pair is one aligned 32-bit load; the left image's six
// per-column values are staged pre-splatted (value * 0x00010001).  A thread keeps its disparity pair fixed and walks
// over the tile's columns, so all indices advance incrementally (no divisions in the loops).
__global__ void __launch_bounds__(256)
sgbm_cost_hsum_kernel(CostArgs a)
{
    extern __shared__ __align__(16) uint8_t cs[];
    const int y = blockIdx.y, f = blockIdx.z;
    const int x0 = blockIdx.x * TX;                         // first cost column of the tile
    const int D = a.D, h = a.h, D2 = D / 2;
    const int NXC = TX + 2 * h;                             // cost columns incl. halo (clamped)
    const int maxD = a.minD + D;
    const int cl0 = clampi(x0 - h, 0, a.W1 - 1), cl1 = clampi(x0 + TX - 1 + h, 0, a.W1 - 1);
    const int xl_lo = cl0 + a.minX1, xl_hi = cl1 + a.minX1;
    const int xr_lo = xl_lo - (maxD - 1), xr_hi = xl_hi - a.minD;
    const int NR = xr_hi - xr_lo + 1;
    const int NRP = (NR + 3) & ~1;                          // elements per reversed array copy (even, +slack)
    uint32_t *pix = reinterpret_cast<uint32_t *>(cs);                   // [NXC][D2] packed u16x2
    uint32_t *lsw = pix + (size_t)NXC * D2;                             // [NXC][6] splatted left values per halo column
    uint16_t *rv = reinterpret_cast<uint16_t *>(lsw + (size_t)NXC * 6); // [6 arrays][2 copies][NRP]
    const uint8_t *pf = a.planes + (size_t)f * a.frame_planes;
    const size_t comp = (size_t)a.H * a.Wp;
    const uint8_t *prow = pf + (size_t)y * a.Wp;
    // left: one (column, component) per thread-iteration; the clamp of the cost column is applied here
    for (int i = threadIdx.x; i < NXC * 6; i += blockDim.x) {
        const int c = i / 6, k = i - c * 6;
        const int xl = clampi(x0 - h + c, 0, a.W1 - 1) + a.minX1;
        lsw[i] = (uint32_t)prow[(size_t)k * comp + xl] * 0x00010001u;
    }
    // right, reversed: element e of copy 0 = value at xr_hi - e; copy 1 element e = copy 0 element e + 1
    for (int arr = 0; arr < 6; arr++) {
        const uint8_t *src = prow + (size_t)(6 + arr) * comp;
        uint16_t *d0 = rv + (size_t)(arr * 2) * NRP, *d1 = d0 + NRP;
        for (int e = threadIdx.x; e <= NRP; e += blockDim.x) {
            const uint16_t v = src[clampi(xr_hi - e, 0, a.W - 1)];
            if (e < NRP) d0[e] = v;
            if (e >= 1) d1[e - 1] = v;
        }
    }
    __syncthreads();
    // pixel cost: thread = disparity pair dp (fixed), walking over the halo columns with stride blockDim / D2
    {
        const int dp = threadIdx.x % D2, cstep = blockDim.x / D2;      // D2 divides 256 for D = 16 .. 256 (powers of two) ...
        const bool regular = (blockDim.x % D2) == 0;                   // ... otherwise fall back to the generic index
        if (regular) {
            for (int c = threadIdx.x / D2; c < NXC; c += cstep) {
                const int xc = clampi(x0 - h + c, 0, a.W1 - 1);
                // d = 2*dp: right pixel xr = xc + minX1 - (d + minD) -> reversed index e = xr_hi - xr
                const int e = xr_hi - (xc + a.minX1 - a.minD) + 2 * dp;
                const int cp = e & 1, w = (e - cp) >> 1;
                const uint32_t *lw = lsw + c * 6;
                const uint32_t *rw = reinterpret_cast<const uint32_t *>(rv + (size_t)cp * NRP) + w;
                const size_t astep = (size_t)NRP;                      // words between consecutive arrays (2 copies x NRP u16)
                uint32_t cost = 0;
#pragma unroll
                for (int pl = 0; pl < 2; pl++) {
                    const uint32_t u = lw[pl * 3 + 0], u0 = lw[pl * 3 + 1], u1 = lw[pl * 3 + 2];
                    const uint32_t v = rw[(pl * 3 + 0) * astep], v0 = rw[(pl * 3 + 1) * astep], v1 = rw[(pl * 3 + 2) * astep];
                    const uint32_t c0 = __vimax3_s16x2(0u, __vsub2(u, v1), __vsub2(v0, u));
                    const uint32_t c1 = __vimax3_s16x2(0u, __vsub2(v, u1), __vsub2(u0, v));
                    uint32_t m = __vmins2(c0, c1);
                    if (pl) m = (m >> 2) & 0x3FFF3FFFu;
                    cost += m;
                }
                pix[c * D2 + dp] = cost;
            }
        } else {
            for (int i = threadIdx.x; i < NXC * D2; i += blockDim.x) {
                const int c = i / D2, dq = i - c * D2;
                const int xc = clampi(x0 - h + c, 0, a.W1 - 1);
                const int e = xr_hi - (xc + a.minX1 - a.minD) + 2 * dq;
                const int cp = e & 1, w = (e - cp) >> 1;
                const uint32_t *lw = lsw + c * 6;
                const uint32_t *rw = reinterpret_cast<const uint32_t *>(rv + (size_t)cp * NRP) + w;
                const size_t astep = (size_t)NRP;
                uint32_t cost = 0;
#pragma unroll
                for (int pl = 0; pl < 2; pl++) {
                    const uint32_t u = lw[pl * 3 + 0], u0 = lw[pl * 3 + 1], u1 = lw[pl * 3 + 2];
                    const uint32_t v = rw[(pl * 3 + 0) * astep], v0 = rw[(pl * 3 + 1) * astep], v1 = rw[(pl * 3 + 2) * astep];
                    const uint32_t c0 = __vimax3_s16x2(0u, __vsub2(u, v1), __vsub2(v0, u));
                    const uint32_t c1 = __vimax3_s16x2(0u, __vsub2(v, u1), __vsub2(u0, v));
                    uint32_t m = __vmins2(c0, c1);
                    if (pl) m = (m >> 2) & 0x3FFF3FFFu;
                    cost += m;
                }
                pix[i] = cost;
            }
        }
    }
    __syncthreads();
    // horizontal window: thread = pair dp, sliding over the tile's columns
    uint32_t *out = reinterpret_cast<uint32_t *>(a.Hs + (size_t)f * a.frame_vol + ((size_t)y * a.W1 + x0) * D);
    const int ncol = min(TX, a.W1 - x0);
    for (int i = threadIdx.x; i < ncol * D2; i += blockDim.x) {
        const int c = i / D2, dp = i - c * D2;
        uint32_t sum = 0;
        for (int k = 0; k <= 2 * h; k++) sum += pix[(c + k) * D2 + dp];
        out[i] = sum;
    }
}

// ------------------------------------------------------------------------------------------------
// Fused cost path (D = 48, 64, 96, 128, 192, 256; blockSize <= 7): BT pixel cost -> horizontal window -> vertical window + P2
// -> C, without the Hs volume.
//
// sgbm_planes2_kernel writes the per-pixel (value, lo, hi) of both BT planes in the form the cost kernel consumes,
// so that staging a row is nothing but 16-byte asynchronous copies:
//   left  GL[y][x1 + 8][8 words]: (value, -value, lo, -hi) of plane 0 then plane 1, each splatted to s16x2, per COST column
//         x1 (replicated 8 columns beyond both ends: the window's column clamp becomes a plain read);
//   right GR[y][6 arrays: value, lo, -hi per plane][2 copies][WR] s16, REVERSED (index eg = W - 1 + 16 - x, so a larger disparity is a larger
//         address) and stored twice, copy 1 shifted by one element, so that any (d, d + 1) pair is one aligned word.
// sgbm_cost_fused_kernel: CTA = (tile of TXk = (256 / D2) * CSEG cost columns, band of rows, frame; (256 / D2) * D2 threads); a thread owns
// one disparity pair and CSEG adjacent columns.  Per row it computes the CSEG + 2h pixel costs of its columns
// (sliding horizontal sum in registers) and pushes the CSEG horizontal sums into per-column vertical rings that
// also live in registers (BS x CSEG words); a row of C leaves once the ring is full.  Rows are replicate-clamped
// by index; the two tiles at the image's left / right end take the EDGE variant that clamps the cost column.
// Rows are staged with cp.async two rows ahead, double buffered: one barrier per row, no staging arithmetic.
// ------------------------------------------------------------------------------------------------
constexpr int CSEG = 8;
constexpr int RSTRIDE = 432;            // u16 elements per staged right array copy (>= TXk + 2h + D + 16)
constexpr int PADL = 8, PADR = 16;

struct Cost2Args {
    const uint8_t *planes; size_t frame_planes;   // per frame: GL block then GR block
    uint16_t *C; size_t frame_vol;
    int W, H, D, minD, minX1, W1, LW, WR, BY;
    uint32_t P2x2;
};

__global__ void __launch_bounds__(256)
sgbm_planes2_kernel(PlaneU8 left, PlaneU8 right, uint8_t *planes, size_t frame_planes, int W, int H, int ftzero,
                    int minX1, int W1, int LW, int WR)
{
    // a warp covers 30 pixels: lanes 1..30 own one each, lanes 0 and 31 only supply the neighbours' plane values
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int x = (blockIdx.x * (blockDim.x >> 5) + wid) * 30 + lane - 1, y = blockIdx.y;
    const int f = blockIdx.z >> 1, img = blockIdx.z & 1;
    const uint8_t *src = img ? right.p + (size_t)f * right.frame : left.p + (size_t)f * left.frame;
    const size_t sp = img ? right.pitch : left.pitch;
    uint32_t val[6];
#pragma unroll
    for (int pl = 0; pl < 2; pl++) {
        const int v = (x >= 0 && x < W) ? bt_plane_value(src, sp, W, H, x, y, pl, ftzero) : 0;
        const int vm = __shfl_up_sync(0xFFFFFFFFu, v, 1), vp = __shfl_down_sync(0xFFFFFFFFu, v, 1);
        const int a = x > 0 ? (v + vm) / 2 : v;
        const int b = x < W - 1 ? (v + vp) / 2 : v;
        val[pl * 3 + 0] = (uint32_t)v; val[pl * 3 + 1] = (uint32_t)min(min(a, b), v); val[pl * 3 + 2] = (uint32_t)max(max(a, b), v);
    }
    if (lane == 0 || lane == 31 || x >= W) return;
    const auto neg16 = [](uint32_t v) { return (0u - v) & 0xFFFFu; };
    uint8_t *pf = planes + (size_t)f * frame_planes;
    if (img == 0) {
        const int x1 = x - minX1;
        if (x1 < 0 || x1 >= W1) return;
        uint4 *row = reinterpret_cast<uint4 *>(pf) + (size_t)y * LW * 2;
        const uint4 q0 = make_uint4(val[0] * 0x00010001u, neg16(val[0]) * 0x00010001u, val[1] * 0x00010001u, neg16(val[2]) * 0x00010001u);
        const uint4 q1 = make_uint4(val[3] * 0x00010001u, neg16(val[3]) * 0x00010001u, val[4] * 0x00010001u, neg16(val[5]) * 0x00010001u);
        const int k0 = x1 == 0 ? 0 : x1 + PADL, k1 = x1 == W1 - 1 ? LW - 1 : x1 + PADL;
        for (int k = k0; k <= k1; k++) { row[k * 2] = q0; row[k * 2 + 1] = q1; }
    } else {
        uint16_t *row = reinterpret_cast<uint16_t *>(pf + (size_t)H * LW * 32) + (size_t)y * 12 * WR;
        const int eg = W - 1 + PADR - x;
        const int k0 = x == W - 1 ? 0 : eg, k1 = x == 0 ? WR : eg;        // copy 1 holds index k - 1, so k may reach WR
        val[2] = neg16(val[2]); val[5] = neg16(val[5]);                    // the hi arrays are stored negated
#pragma unroll
        for (int arr = 0; arr < 6; arr++) {
            uint16_t *c0 = row + (size_t)(arr * 2) * WR, *c1 = c0 + WR;
            for (int k = k0; k <= k1; k++) {
                if (k < WR) c0[k] = (uint16_t)val[arr];
                if (k >= 1) c1[k - 1] = (uint16_t)val[arr];
            }
        }
    }
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// l = (u, -u, lo(u), -hi(u)) of the left pixel (splatted), rw -> (v, lo(v), -hi(v)) arrays of the right pixel pair:
// cost = min(max(0, u - hi(v), lo(v) - u), max(0, v - hi(u), lo(u) - v)), 5 packed ops + the negation of v
__device__ __forceinline__ uint32_t bt_plane2(const uint4 l, const uint32_t v, const uint32_t v0, const uint32_t nv1)
{
    const uint32_t nv = __vadd2(~v, 0x00010001u);
    const uint32_t c0 = __viaddmax_s16x2_relu(v0, l.y, __vadd2(l.x, nv1));
    const uint32_t c1 = __viaddmax_s16x2_relu(l.z, nv, __vadd2(v, l.w));
    return __vmins2(c0, c1);
}

__device__ __forceinline__ uint32_t bt_cost2(const uint4 l0, const uint4 l1, const uint32_t *rw)
{
    const uint32_t m0 = bt_plane2(l0, rw[0], rw[RSTRIDE], rw[2 * RSTRIDE]);
    const uint32_t m1 = bt_plane2(l1, rw[3 * RSTRIDE], rw[4 * RSTRIDE], rw[5 * RSTRIDE]);
    return m0 + ((m1 >> 2) & 0x3FFF3FFFu);
}

// the same with the six right-image words already in registers
__device__ __forceinline__ uint32_t bt_cost2r(const uint4 l0, const uint4 l1, const uint32_t (&r)[6])
{
    const uint32_t m0 = bt_plane2(l0, r[0], r[1], r[2]);
    const uint32_t m1 = bt_plane2(l1, r[3], r[4], r[5]);
    return m0 + ((m1 >> 2) & 0x3FFF3FFFu);
}

template <int BS>
__global__ void __launch_bounds__(256, 2)
sgbm_cost_fused_kernel(Cost2Args a)
{
    constexpr int h = BS / 2, NC = CSEG + 2 * h;
    extern __shared__ __align__(16) uint8_t cs[];
    const int D2 = a.D / 2, NT = blockDim.x, nseg = NT / D2, TXk = nseg * CSEG, NXC = TXk + 2 * h;   // NT = (256 / D2) * D2 threads
    const int dp = threadIdx.x % D2, seg = threadIdx.x / D2;
    const int x0 = blockIdx.x * TXk, y0 = blockIdx.y * a.BY, y1 = min(y0 + a.BY, a.H), f = blockIdx.z;
    const bool edge = x0 - h < 0 || x0 + TXk - 1 + h > a.W1 - 1;
    const int cl1 = clampi(x0 + TXk - 1 + h, 0, a.W1 - 1);
    const int xr_hi = cl1 + a.minX1 - a.minD;                          // right pixel of reversed index e = 0
    // e = E0 - xc: reversed index of the right pixel of cost column xc at this thread's disparity pair
    const int E0 = xr_hi - (a.minX1 - a.minD) + 2 * dp;
    const int xs = x0 - h + seg * CSEG;                                // this thread's first cost column (unclamped)
    const int p0 = (E0 - xs) & 1;                                      // CTA-uniform: parity of e at even c
    // staged copy A serves even c (pairs start at parity p0), copy B odd c.  Global sources, 8-element aligned:
    const int eg0 = a.W - 1 + PADR - xr_hi;
    const int sA = eg0 + p0, sB = eg0 + 1 - p0;
    const int srcA = sA & 1, srcB = sB & 1;                            // which global copy
    // first global element staged; copy A starts at least two elements below the first pair it serves: interior tiles take
    // the odd columns' pairs from copy A as well (high half of one word, low half of the next), which reaches one word lower
    const int iA = (sA - srcA - 2) & ~7, iB = (sB - srcB) & ~7;
    const int dA = sA - srcA - iA, dB = sB - srcB - iB;                // staged index of e = p0 / e = 1 - p0 (2 <= dA <= 9, dB <= 7)
    const int nelem = (cl1 - clampi(x0 - h, 0, a.W1 - 1)) + a.D + 2;   // elements needed from e = 0
    const int NRC = (max(dA, dB) + nelem + 7) / 8;                     // chunks per array copy
    const int NL = NXC * 2;
    const int lbytes = NXC * 32, rbytes = 12 * RSTRIDE * 2, bbytes = lbytes + rbytes;
    const uint8_t *pf = a.planes + (size_t)f * a.frame_planes;
    const uint8_t *gl = pf + (size_t)(x0 - h + PADL) * 32;
    const uint8_t *gr = pf + (size_t)a.H * a.LW * 32;
    const size_t lrow = (size_t)a.LW * 32, rrow = (size_t)12 * a.WR * 2;
    // this thread's (at most two) 16-byte chunks of a row: global pointer for row 0 + row stride, shared offset
    const uint8_t *gsrc[2]; uint32_t gstr[2], soff[2];
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(cs);
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int t = threadIdx.x + q * NT;
        gsrc[q] = nullptr; gstr[q] = 0u; soff[q] = 0u;
        if (t < NL) { gsrc[q] = gl + t * 16; gstr[q] = (uint32_t)lrow; soff[q] = sbase + t * 16; }
        else if (t < NL + 12 * NRC) {
            const int r = t - NL, ac = r / NRC, ch = r - ac * NRC, arr = ac >> 1, cpy = ac & 1;
            const int gcopy = cpy ? srcB : srcA, gi = (cpy ? iB : iA) + ch * 8;
            if (edge || cpy == 0) {                                    // interior tiles never read copy B
                gsrc[q] = gr + (size_t)((arr * 2 + gcopy) * a.WR + gi) * 2; gstr[q] = (uint32_t)rrow;
                soff[q] = sbase + lbytes + (ac * RSTRIDE + ch * 8) * 2;
            }
        }
    }
    auto stage = [&](int row, int b) {
#pragma unroll
        for (int q = 0; q < 2; q++)
            if (gsrc[q])
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(soff[q] + b * bbytes), "l"(gsrc[q] + (size_t)row * gstr[q]) : "memory");
        cp_async_commit();
    };

    uint32_t ring[BS][CSEG], V[CSEG];
#pragma unroll
    for (int j = 0; j < CSEG; j++) {
        V[j] = a.P2x2;
#pragma unroll
        for (int u = 0; u < BS; u++) ring[u][j] = 0u;
    }
    const int nrows = (y1 - y0) + 2 * h;
    // byte offset inside a buffer: even c reads copy A at ta - 2c (edge tiles compute their offsets per column)
    const int ta = lbytes + 2 * (E0 - xs - p0 + dA);
    const int tl = seg * CSEG * 32;
    uint32_t *cptr = reinterpret_cast<uint32_t *>(a.C + (size_t)f * a.frame_vol) + ((size_t)y0 * a.W1 + x0 + seg * CSEG) * D2 + dp;
    const size_t crow = (size_t)a.W1 * D2;
    const int ncols = a.W1 - (x0 + seg * CSEG);                       // columns of this thread inside the image (may be <= 0)
    stage(clampi(y0 - h, 0, a.H - 1), 0);
    if (nrows > 1) stage(clampi(y0 - h + 1, 0, a.H - 1), 1);
    cp_async_wait_all();
    __syncthreads();
    for (int base = 0; base < nrows; base += BS) {
#pragma unroll
        for (int u = 0; u < BS; u++) {
            const int i = base + u;
            if (i < nrows) {                                        // CTA-uniform
                const uint8_t *xb = cs + (i & 1) * bbytes;
                const uint4 *lp = reinterpret_cast<const uint4 *>(xb + tl);
                uint32_t hc[BS], hs = 0u;
                if (!edge) {
                    // columns in pairs: the even column's six words come from copy A (one word per array, stepping down),
                    // the odd column's pairs straddle two of those words -- 42 instead of 72 shared-memory loads per row
                    // (the kernel is bound by shared-memory bandwidth), the byte permutes go to the half-idle ALU pipe
                    const uint32_t *pa = reinterpret_cast<const uint32_t *>(xb + ta);
                    uint32_t cur[6], nxt[6], odd[6];
#pragma unroll
                    for (int q = 0; q < 6; q++) cur[q] = pa[q * RSTRIDE];
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        uint32_t cost;
                        if ((c & 1) == 0) {
#pragma unroll
                            for (int q = 0; q < 6; q++) nxt[q] = pa[q * RSTRIDE - (c / 2 + 1)];
                            cost = bt_cost2r(lp[c * 2], lp[c * 2 + 1], cur);
                        } else {
#pragma unroll
                            for (int q = 0; q < 6; q++) { odd[q] = __byte_perm(nxt[q], cur[q], 0x5432); cur[q] = nxt[q]; }
                            cost = bt_cost2r(lp[c * 2], lp[c * 2 + 1], odd);
                        }
                        if (c >= BS) hs -= hc[c % BS];
                        hs += cost;
                        hc[c % BS] = cost;
                        if (c >= 2 * h) {
                            const int j = c - 2 * h;
                            V[j] += hs - ring[u][j];
                            ring[u][j] = hs;
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        const int xc = clampi(xs + c, 0, a.W1 - 1);
                        const int e = E0 - xc;
                        const int off = ((e - p0) & 1) ? lbytes + RSTRIDE * 2 + 2 * (e - 1 + p0 + dB) : lbytes + 2 * (e - p0 + dA);
                        const uint32_t *rw = reinterpret_cast<const uint32_t *>(xb + off);
                        const uint32_t cost = bt_cost2(lp[c * 2], lp[c * 2 + 1], rw);
                        if (c >= BS) hs -= hc[c % BS];
                        hs += cost;
                        hc[c % BS] = cost;
                        if (c >= 2 * h) {
                            const int j = c - 2 * h;
                            V[j] += hs - ring[u][j];
                            ring[u][j] = hs;
                        }
                    }
                }
                if (i >= 2 * h) {
                    if (ncols >= CSEG) {
#pragma unroll
                        for (int j = 0; j < CSEG; j++) cptr[j * D2] = V[j];
                    } else {
#pragma unroll
                        for (int j = 0; j < CSEG; j++)
                            if (j < ncols) cptr[j * D2] = V[j];
                    }
                    cptr += crow;
                }
                cp_async_wait_all();                                // row i + 1 (in flight during this row) has landed
                __syncthreads();                                    // ... for everyone, and buffer i & 1 is free
                if (i + 2 < nrows) stage(clampi(y0 - h + i + 2, 0, a.H - 1), i & 1);
            }
        }
    }
}

// C = P2 + sum over clamped rows y-h..y+h of Hs.  One thread per u16x2 word column and band of rows: the
// 2h+1 window values live in a register ring (compile-time size), so every Hs word is read once.
template <int BS>
__global__ void __launch_bounds__(256)
sgbm_vsum_kernel(const uint32_t *Hs, uint32_t *C, size_t frame_words, size_t row_words, int H, int band, uint32_t P2x2)
{
    constexpr int h = BS / 2;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = blockIdx.y * band, y1 = min(y0 + band, H), f = blockIdx.z;
    if (i >= row_words || y0 >= y1) return;
    const uint32_t *src = Hs + (size_t)f * frame_words + i;
    uint32_t *dst = C + (size_t)f * frame_words + i;
    uint32_t ring[BS];
    uint32_t s = P2x2;
#pragma unroll
    for (int j = 0; j < BS; j++) { ring[j] = src[(size_t)clampi(y0 - h + j, 0, H - 1) * row_words]; s += ring[j]; }
    // ring[(y - y0 + j) % BS] holds row clamp(y - h + j); slot (y - y0) % BS is the oldest
    for (int yb = y0; yb < y1; yb += BS) {
#pragma unroll
        for (int u = 0; u < BS; u++) {
            const int y = yb + u;
            if (y < y1) {
                dst[(size_t)y * row_words] = s;
                const uint32_t in = src[(size_t)clampi(y + 1 + h, 0, H - 1) * row_words];
                s += in - ring[u];
                ring[u] = in;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// path aggregation: one warp per chain.  K2 = u16x2 words per lane; lane l owns d in [2*K2*l, 2*K2*(l+1)).
// ------------------------------------------------------------------------------------------------
struct PathArgs {
    const uint32_t *C; uint32_t *S; size_t frame_words;
    int W1, H, D, P1, P2;
    int px, py;          // predecessor offset
    int first;           // 1: S = L (no read), 0: S += L
    int nchains;
    uint2 *rec; size_t frame_rec;    // fused last path: per-pixel WTA records
    int mul;                         // 100 - uniquenessRatio
    uint32_t one;                    // 1: a multiplier the compiler cannot fold (see vcore)
};

__device__ __forceinline__ uint32_t min2(uint32_t a, uint32_t b) { return __vminu2(a, b); }

// LPC = lanes per chain (32, or 16 / 8 when D/2 words divide evenly: then one warp carries 32/LPC chains, which
// doubles / quadruples the loads in flight of this latency-bound kernel).  Lanes >= D/2/K2 of a chain are idle.
template <int K2, int LPC>
__global__ void __launch_bounds__(128)
sgbm_path_kernel(PathArgs a)
{
    constexpr int CPW = 32 / LPC;                          // chains per warp
    const int lane = threadIdx.x & 31, sl = lane % LPC;
    const int chain = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * CPW + lane / LPC;
    const int f = blockIdx.y;
    const int sx = -a.px, sy = -a.py;
    bool alive = chain < a.nchains;
    if (CPW == 1 && !alive) return;                         // whole-warp chains: nothing to keep in step with
    // chain start: a pixel whose predecessor lies outside the image
    int x = 0, y = 0;
    if (alive) {
        const int nrow = sy != 0 ? a.W1 : 0;
        if (chain < nrow) { x = chain; y = sy > 0 ? 0 : a.H - 1; }
        else {
            const int j = chain - nrow;
            x = sx > 0 ? 0 : a.W1 - 1;
            y = sy > 0 ? j + 1 : j;                        // skip the corner owned by the row set
        }
    }
    const int wordsD = a.D / 2;
    const int nlanes = wordsD / K2;                        // active lanes of a chain
    const bool act = sl < nlanes;
    const uint32_t P1x2 = (uint32_t)a.P1 * 0x00010001u;
    const uint32_t *Cf = a.C + (size_t)f * a.frame_words;
    uint32_t *Sf = a.S + (size_t)f * a.frame_words;
    uint32_t Lp[K2];
#pragma unroll
    for (int k = 0; k < K2; k++) Lp[k] = 0u;               // out-of-image predecessor: L = 0
    uint32_t minLp = 0u;
    const long long stepw = ((long long)sy * a.W1 + sx) * wordsD;
    size_t off = ((size_t)y * a.W1 + x) * wordsD + (size_t)sl * K2;
    uint32_t c[K2], s[K2];
#pragma unroll
    for (int k = 0; k < K2; k++) { c[k] = (alive && act) ? Cf[off + k] : 0u; s[k] = (alive && act && !a.first) ? Sf[off + k] : 0u; }
    while (CPW == 1 || __any_sync(0xFFFFFFFFu, alive)) {
        const int xn = x + sx, yn = y + sy;
        const bool more = alive && (xn >= 0 && xn < a.W1 && yn >= 0 && yn < a.H);
        // prefetch the next pixel's C and S
        uint32_t cn[K2], sn[K2];
        const size_t offn = (size_t)((long long)off + stepw);
#pragma unroll
        for (int k = 0; k < K2; k++) {
            cn[k] = (more && act) ? Cf[offn + k] : 0u;
            sn[k] = (more && act && !a.first) ? Sf[offn + k] : 0u;
        }
        // neighbours across lanes: last half of the left lane, first half of the right lane
        uint32_t left = __shfl_up_sync(0xFFFFFFFFu, Lp[K2 - 1], 1, LPC);
        uint32_t right = __shfl_down_sync(0xFFFFFFFFu, Lp[0], 1, LPC);
        if (sl == 0) left = 0x7FFF0000u;                   // L[-1] = MAX_COST (upper half is used)
        if (sl >= nlanes - 1) right = 0x00007FFFu;         // L[D]  = MAX_COST (lower half is used)
        const uint32_t delta = (uint32_t)a.P2 + minLp;     // < 65536
        const uint32_t dx2 = delta * 0x00010001u;
        uint32_t Ln[K2];
        uint32_t m = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < K2; k++) {
            const uint32_t prevw = k == 0 ? left : Lp[k - 1];
            const uint32_t nextw = k == K2 - 1 ? right : Lp[k + 1];
            const uint32_t lm1 = __byte_perm(prevw, Lp[k], 0x5432);     // (L[d-1], L[d])   for the pair (d, d+1)
            const uint32_t lp1 = __byte_perm(Lp[k], nextw, 0x5432);     // (L[d+1], L[d+2])
            uint32_t t = __vimin3_u16x2(Lp[k], __vadd2(lm1, P1x2), __vadd2(lp1, P1x2));
            t = min2(t, dx2);
            t = __vsub2(__vadd2(t, c[k]), dx2);
            Ln[k] = t;
            m = min2(m, t);
            if (alive && act) {
                uint32_t sv = a.first ? t : min2(__vadd2(s[k], t), 0x7FFF7FFFu);
                Sf[off + k] = sv;
            }
        }
        // min over the chain's D values: inactive lanes must not contribute
        uint32_t mm = min(m & 0xFFFFu, m >> 16);
        if (!act) mm = 0xFFFFu;
        // hardware warp reduction (REDUX.MIN) over the lanes of this chain
        minLp = __reduce_min_sync(LPC == 32 ? 0xFFFFFFFFu : (((1u << LPC) - 1u) << (lane & ~(LPC - 1))), mm);
#pragma unroll
        for (int k = 0; k < K2; k++) { Lp[k] = Ln[k]; c[k] = cn[k]; s[k] = sn[k]; }
        if (CPW == 1 && !more) break;
        alive = more;
        x = xn; y = yn; off = offn;
    }
}

// ------------------------------------------------------------------------------------------------
// path aggregation, specialised: D = 8 * LPC (LPC = 16: D 128, LPC = 8: D 64), every lane owns one uint4 = 4 words
// = 8 disparities of its chain; 32 / LPC chains per warp.  Counted loop over the chain length, pointer stepping,
// 128-bit loads / stores, the next pixel's C and S in flight while the current one is computed (two steps per
// loop iteration, so the prefetch registers swap roles without moves).
// MODE 0: first path (S = L), 1: S += L, 2: last path -- S + L stays in registers and the winner-take-all of the
// pixel (argmin, uniqueness, the two neighbours for the sub-pixel fit) is taken right there: the finished S volume
// is never written nor read back; an 8-byte record per pixel goes to sgbm_lr_kernel instead.
// ------------------------------------------------------------------------------------------------
template <int LPC>
__device__ __forceinline__ uint32_t group_min_u32(uint32_t v, int grp)
{
    // min over the LPC lanes of each chain: one full-warp REDUX per group, the other groups contribute ~0
    uint32_t r = 0u;
#pragma unroll
    for (int g = 0; g < 32 / LPC; g++) {
        const uint32_t m = __reduce_min_sync(0xFFFFFFFFu, grp == g ? v : 0xFFFFFFFFu);
        if (grp == g) r = m;
    }
    return r;
}

// NW packed words (2 NW disparities) of one lane: the specialised kernels below keep a lane's share of a pixel's D values in
// registers.  D = 2 * NW * LPC with LPC lanes per pixel: NW = 4 gives 64 / 128 (LPC 8 / 16), NW = 3 gives 48 / 96 / 192 (LPC 8 /
// 16 / 32) -- the reference's default `-nd 192` scaled to 320 / 640 / 1280 pixel frames (utils/cmdline-parser.h:85-89).
template <int NW> struct PV { uint32_t w[NW]; };

template <int NW>
__device__ __forceinline__ PV<NW> pv_zero()
{
    PV<NW> r;
#pragma unroll
    for (int k = 0; k < NW; k++) r.w[k] = 0u;
    return r;
}

// loads / stores at word pointers: one 128-bit access for NW = 4 (16-byte aligned by construction), NW 32-bit accesses else
// (a lane's 12 bytes are only 4-byte aligned; the lanes of a pixel still cover one contiguous run)
template <int NW>
__device__ __forceinline__ PV<NW> pv_ldg(const uint32_t *p)
{
    PV<NW> r;
    if constexpr (NW == 4) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p));
        r.w[0] = q.x; r.w[1] = q.y; r.w[2] = q.z; r.w[3] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < NW; k++) r.w[k] = __ldg(p + k);
    }
    return r;
}

template <int NW>
__device__ __forceinline__ PV<NW> pv_ld(const uint32_t *p)
{
    PV<NW> r;
    if constexpr (NW == 4) {
        const uint4 q = *reinterpret_cast<const uint4 *>(p);
        r.w[0] = q.x; r.w[1] = q.y; r.w[2] = q.z; r.w[3] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < NW; k++) r.w[k] = p[k];
    }
    return r;
}

template <int NW>
__device__ __forceinline__ void pv_st(uint32_t *p, const PV<NW> &v)
{
    if constexpr (NW == 4) *reinterpret_cast<uint4 *>(p) = make_uint4(v.w[0], v.w[1], v.w[2], v.w[3]);
    else {
#pragma unroll
        for (int k = 0; k < NW; k++) p[k] = v.w[k];
    }
}

// One step of one path for a lane's 2 NW disparities (horizontal chains, row sweeps, whole-height passes).  `minLp` and the
// returned `m` are the minimum over d SPLATTED into both halves, so the 32-bit minimum over a pixel's lanes is the splatted
// minimum and P2 + min is one add.  The packed 16-bit minima only issue on the ALU pipe (64 lanes per clock and SM), which
// is what bounds the passes, so the plain 32-bit arithmetic is written as multiply-add with `one`, a run-time 1 the compiler
// cannot see through and therefore keeps as IMAD -- the other integer pipe.  v + C - (P2 + min): every 16-bit half of the
// result is a path cost (0 <= L < 65536, v >= min), so no carry or borrow crosses the halves and 32-bit arithmetic is exact.
template <int LPC, int NW>
__device__ __forceinline__ void vcoreN(const PV<NW> &Lp, const uint32_t minLp, const PV<NW> &c, int sl, uint32_t one, uint32_t P1x2, uint32_t P2x2,
                                       uint32_t (&t)[NW], uint32_t &m)
{
    uint32_t left = __shfl_up_sync(0xFFFFFFFFu, Lp.w[NW - 1], 1, LPC);
    uint32_t right = __shfl_down_sync(0xFFFFFFFFu, Lp.w[0], 1, LPC);
    if (sl == 0) left = 0x7FFF0000u;                       // L[-1] = MAX_COST (upper half is used)
    if (sl == LPC - 1) right = 0x00007FFFu;                // L[D]  = MAX_COST (lower half is used)
    const uint32_t dx2 = minLp * one + P2x2, ndx2 = 0u - dx2;      // P2 + min_k L_r(p - r, k) < 65536
    uint32_t w[NW + 2];
    w[0] = left; w[NW + 1] = right;
#pragma unroll
    for (int i = 0; i < NW; i++) w[i + 1] = Lp.w[i];
    uint32_t mm = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        const uint32_t lm1 = __byte_perm(w[i], w[i + 1], 0x5432);      // (L[d-1], L[d])   for the pair (d, d+1)
        const uint32_t lp1 = __byte_perm(w[i + 1], w[i + 2], 0x5432);  // (L[d+1], L[d+2])
        // min(a + P1, b + P1) = min(a, b) + P1, and MAX_COST + P1 < 65536: the add is plain (IMAD), the three-way minimum one ALU op
        const uint32_t v = __vimin3_u16x2(w[i + 1], min2(lm1, lp1) * one + P1x2, dx2);
        t[i] = (v * one + c.w[i]) * one + ndx2;
        mm = i == 0 ? t[0] : min2(mm, t[i]);
    }
    m = min2(mm, __byte_perm(mm, mm, 0x1032));
}

// record of one pixel: x = minS | code << 16 (code = best d, | 0x8000 when the uniqueness test failed, 0xFFFF when
// the pixel is degenerate), y = S[d-1] | S[d+1] << 16
template <int LPC, int NW>
__device__ __forceinline__ uint2 wta_record(const uint32_t (&sv)[NW], int sl, int grp, int mul)
{
    const uint32_t dbase = 2u * NW * (uint32_t)sl;
    uint32_t key = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < NW; k++) {
        const uint32_t d0 = dbase + 2u * k;
        key = min(key, min((sv[k] << 16) | d0, (sv[k] & 0xFFFF0000u) | (d0 + 1u)));      // first minimum wins
    }
    key = group_min_u32<LPC>(key, grp);
    const uint32_t minS = key >> 16, bd = key & 0xFFFFu;
    // smallest S outside [bd - 1, bd + 1]
    uint32_t m2 = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < NW; k++) {
        const uint32_t d0 = dbase + 2u * k;
        uint32_t w = sv[k];
        if (d0 - bd + 1u <= 2u) w |= 0x0000FFFFu;
        if (d0 - bd + 2u <= 2u) w |= 0xFFFF0000u;
        m2 = min2(m2, w);
    }
    const uint32_t other = group_min_u32<LPC>(min(m2 & 0xFFFFu, m2 >> 16), grp);
    const bool viol = other * (uint32_t)mul < minS * 100u;            // exists d: S[d] * (100 - uniq) < minS * 100
    // neighbours of the winner, fetched from their owner lanes (clamped at the ends; unused there)
    uint32_t nb[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const uint32_t dq = q ? min(bd + 1u, 2u * NW * LPC - 1u) : (bd > 0u ? bd - 1u : 0u);
        const uint32_t owner = dq / (2u * NW), wi = (dq >> 1) - owner * NW;
        uint32_t v = sv[NW - 1];
#pragma unroll
        for (int k = NW - 2; k >= 0; k--) v = wi == (uint32_t)k ? sv[k] : v;
        v = __shfl_sync(0xFFFFFFFFu, v, grp * LPC + (int)owner);
        nb[q] = (dq & 1u) ? v >> 16 : v & 0xFFFFu;
    }
    uint32_t code = bd | (viol ? 0x8000u : 0u);
    if (minS >= 32767u) code = 0xFFFFu;
    return make_uint2(minS | (code << 16), nb[0] | (nb[1] << 16));
}

template <int LPC, int NW, int MODE>
__device__ __forceinline__ void path_step(PV<NW> &Lp, uint32_t &minLp, const PV<NW> &c, const PV<NW> &s, uint32_t *sp, uint2 *rp,
                                          bool live, int sl, int grp, uint32_t P1x2, uint32_t P2, int mul, uint32_t one)
{
    uint32_t t[NW], m;
    vcoreN<LPC, NW>(Lp, minLp, c, sl, one, P1x2, P2, t, m);
    PV<NW> tv;
#pragma unroll
    for (int k = 0; k < NW; k++) tv.w[k] = t[k];
    if (MODE == 0) {
        if (live) pv_st<NW>(sp, tv);
    } else {
        // S <= 32767 and L < 32768 per half: the sum cannot carry into the other half
        uint32_t sv[NW];
        PV<NW> so;
#pragma unroll
        for (int k = 0; k < NW; k++) { sv[k] = min2(t[k] * one + s.w[k], 0x7FFF7FFFu); so.w[k] = sv[k]; }
        if (MODE == 1) {
            if (live) pv_st<NW>(sp, so);
        } else {
            const uint2 rec = wta_record<LPC, NW>(sv, sl, grp, mul);
            if (live && sl == 0) *rp = rec;
        }
    }
    Lp = tv;
    minLp = group_min_u32<LPC>(m, grp);
}

template <int LPC, int NW, int MODE>
__global__ void __launch_bounds__(128)
sgbm_path4_kernel(PathArgs a)
{
    constexpr int CPW = 32 / LPC;
    const int lane = threadIdx.x & 31, sl = lane % LPC, grp = lane / LPC;
    const int chain = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * CPW + grp;
    const int f = blockIdx.y;
    const int sx = -a.px, sy = -a.py;
    // chain start (a pixel whose predecessor lies outside the image) and length
    int x = 0, y = 0, len = 0;
    if (chain < a.nchains) {
        const int nrow = sy != 0 ? a.W1 : 0;
        if (chain < nrow) { x = chain; y = sy > 0 ? 0 : a.H - 1; }
        else {
            const int j = chain - nrow;
            x = sx > 0 ? 0 : a.W1 - 1;
            y = sy > 0 ? j + 1 : j;                        // skip the corner owned by the row set
        }
        const int nx = sx > 0 ? a.W1 - x : (sx < 0 ? x + 1 : 0x7FFFFFFF);
        const int ny = sy > 0 ? a.H - y : (sy < 0 ? y + 1 : 0x7FFFFFFF);
        len = min(nx, ny);
    }
    const int maxlen = __reduce_max_sync(0xFFFFFFFFu, len);
    const uint32_t P1x2 = (uint32_t)a.P1 * 0x00010001u, P2 = (uint32_t)a.P2 * 0x00010001u;
    constexpr int wordsD = NW * LPC;
    const long long stepp = (long long)sy * a.W1 + sx;                        // pixels
    const long long stepw = stepp * wordsD;                                   // words
    const size_t pix0 = (size_t)y * a.W1 + x;
    const size_t off = (size_t)f * a.frame_words + pix0 * wordsD + (size_t)sl * NW;
    const uint32_t *cp = a.C + off;
    uint32_t *sp = a.S + off;
    uint2 *rp = MODE == 2 ? a.rec + (size_t)f * a.frame_rec + pix0 : nullptr;
    const PV<NW> z = pv_zero<NW>();
    PV<NW> Lp = z;                                          // out-of-image predecessor: L = 0
    uint32_t minLp = 0u;
    PV<NW> c0 = len > 0 ? pv_ldg<NW>(cp) : z, s0 = (MODE != 0 && len > 0) ? pv_ld<NW>(sp) : z, c1, s1;
    for (int i = 0; i < maxlen; i += 2) {
        const bool m1 = i + 1 < len;
        c1 = m1 ? pv_ldg<NW>(cp + stepw) : z;
        s1 = (MODE != 0 && m1) ? pv_ld<NW>(sp + stepw) : z;
        path_step<LPC, NW, MODE>(Lp, minLp, c0, s0, sp, rp, i < len, sl, grp, P1x2, P2, a.mul, a.one);
        const bool m2 = i + 2 < len;
        c0 = m2 ? pv_ldg<NW>(cp + 2 * stepw) : z;
        s0 = (MODE != 0 && m2) ? pv_ld<NW>(sp + 2 * stepw) : z;
        path_step<LPC, NW, MODE>(Lp, minLp, c1, s1, sp + stepw, rp + stepp, m1, sl, grp, P1x2, P2, a.mul, a.one);
        cp += 2 * stepw; sp += 2 * stepw;
        if (MODE == 2) rp += 2 * stepp;
    }
}

// ------------------------------------------------------------------------------------------------
// Row sweep: the three paths that come from the previous row (r = (-1,dy'), (0,dy'), (+1,dy')) advance together, so
// C is read once and S is read-modified-written once for three directions instead of three times each.
// One launch handles a tile of SW_R image rows for every frame; a CTA owns X = NS - 2*SW_R adjacent columns and
// recomputes SW_R halo columns on both sides (a diagonal chain crosses at most one column per row, so everything
// an owned pixel needs within the tile starts inside the halo).  LPC lanes per column as in sgbm_path4_kernel:
// the vertical path's L stays in registers, the two diagonal paths hand their L to the neighbouring column through
// shared memory (double buffered, one barrier per row).  Between tiles the last row's L of all three paths and
// their minima travel through a small global "frontier" buffer (ping-pong).
// ------------------------------------------------------------------------------------------------
constexpr int SW_R = 8;

struct SweepArgs {
    const uint32_t *C; uint32_t *S; size_t frame_words;
    const uint32_t *Fin; uint32_t *Fout;     // frontier [3][W1][wordsD] per frame (0: diagonal from x-1, 1: vertical, 2: diagonal from x+1)
    const uint32_t *Min; uint32_t *Mout;     // minima   [3][W1] per frame
    size_t frame_front;                      // words between consecutive frames of the four arrays above
    int W1, H, P1, P2;
    int ystart, ystep, nrows;                // rows of this tile: ystart, ystart + ystep, ...
    int first;                               // first tile of the sweep: predecessors are outside the image
    uint32_t one;                            // 1: a multiplier the compiler cannot fold (see vcore)
};

// the tiled sweep keeps its uint4 form (8 disparities per lane)
template <int LPC>
__device__ __forceinline__ void vcore(const uint4 Lp, const uint32_t minLp, const uint4 c, int sl, uint32_t one, uint32_t P1x2, uint32_t P2x2,
                                      uint32_t (&t)[4], uint32_t &m)
{
    const PV<4> l = {{Lp.x, Lp.y, Lp.z, Lp.w}}, cc = {{c.x, c.y, c.z, c.w}};
    vcoreN<LPC, 4>(l, minLp, cc, sl, one, P1x2, P2x2, t, m);
}

template <int LPC, bool SAFE3>
__global__ void __launch_bounds__(1024, 1)
sgbm_sweep_kernel(SweepArgs a)
{
    constexpr int NS = 1024 / LPC, X = NS - 2 * SW_R, WQ = LPC;      // slots, owned columns, uint4 per column
    extern __shared__ __align__(16) uint8_t sw[];
    uint4 *exL = reinterpret_cast<uint4 *>(sw);                       // [2][NS + 2][WQ]  L of the path from x-1, by producing slot + 1
    uint4 *exR = exL + 2 * (NS + 2) * WQ;                             // [2][NS + 2][WQ]  L of the path from x+1
    uint32_t *mnL = reinterpret_cast<uint32_t *>(exR + 2 * (NS + 2) * WQ);   // [2][NS + 2]
    uint32_t *mnR = mnL + 2 * (NS + 2);
    const int tid = threadIdx.x, j = tid / LPC, sl = tid % LPC, grp = (tid & 31) / LPC;
    const int f = blockIdx.y;
    const int x = blockIdx.x * X - SW_R + j;
    const bool inimg = x >= 0 && x < a.W1;
    const bool own = inimg && j >= SW_R && j < SW_R + X;
    const int wordsD = 4 * LPC;
    // zero both exchange buffers (pads and out-of-image slots stay zero: L = 0, min = 0)
    for (int i = tid; i < 2 * 2 * (NS + 2) * WQ; i += 1024) exL[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 2 * 2 * (NS + 2); i += 1024) mnL[i] = 0u;
    __syncthreads();
    const size_t fw = (size_t)a.W1 * wordsD;                          // words per frontier direction
    uint4 Lv = make_uint4(0u, 0u, 0u, 0u);
    uint32_t minV = 0u;
    if (!a.first && inimg) {
        const uint32_t *F = a.Fin + (size_t)f * a.frame_front + (size_t)x * wordsD + sl * 4;
        const uint32_t *M = a.Min + (size_t)f * a.frame_front + x;
        exL[(0 * (NS + 2) + j + 1) * WQ + sl] = *reinterpret_cast<const uint4 *>(F);
        Lv = *reinterpret_cast<const uint4 *>(F + fw);
        exR[(0 * (NS + 2) + j + 1) * WQ + sl] = *reinterpret_cast<const uint4 *>(F + 2 * fw);
        minV = M[a.W1];
        if (sl == 0) { mnL[j + 1] = M[0]; mnR[j + 1] = M[2 * a.W1]; }
    }
    const uint32_t P1x2 = (uint32_t)a.P1 * 0x00010001u, P2 = (uint32_t)a.P2 * 0x00010001u;
    const long long rowq = (long long)a.ystep * a.W1 * (wordsD / 4);  // uint4 units per row step
    const size_t off = (((size_t)f * a.frame_words) + ((size_t)a.ystart * a.W1 + (inimg ? x : 0)) * wordsD) / 4 + sl;
    const uint4 *cp = reinterpret_cast<const uint4 *>(a.C) + off;
    uint4 *sp = reinterpret_cast<uint4 *>(a.S) + off;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    uint4 c = inimg ? __ldg(cp) : z, s = own ? *sp : z;
    uint4 tL = z, tR = z;
    uint32_t mL = 0u, mR = 0u;
    __syncthreads();
    for (int r = 0; r < a.nrows; r++) {
        const int b = r & 1;
        const bool more = r + 1 < a.nrows;
        const uint4 cn = (inimg && more) ? __ldg(cp + rowq) : z;
        const uint4 sn = (own && more) ? sp[rowq] : z;
        const uint4 pL = exL[(b * (NS + 2) + j) * WQ + sl];           // path from (x-1, previous row): slot j-1
        const uint4 pR = exR[(b * (NS + 2) + j + 2) * WQ + sl];       // path from (x+1, previous row): slot j+1
        const uint32_t pmL = mnL[b * (NS + 2) + j], pmR = mnR[b * (NS + 2) + j + 2];
        uint32_t t0[4], t1[4], t2[4], m0, m1, m2;
        vcore<LPC>(pL, pmL, c, sl, a.one, P1x2, P2, t0, m0);
        vcore<LPC>(Lv, minV, c, sl, a.one, P1x2, P2, t1, m1);
        vcore<LPC>(pR, pmR, c, sl, a.one, P1x2, P2, t2, m2);
        mL = group_min_u32<LPC>(m0, grp);
        minV = group_min_u32<LPC>(m1, grp);
        mR = group_min_u32<LPC>(m2, grp);
        tL = make_uint4(t0[0], t0[1], t0[2], t0[3]);
        Lv = make_uint4(t1[0], t1[1], t1[2], t1[3]);
        tR = make_uint4(t2[0], t2[1], t2[2], t2[3]);
        if (inimg) {
            exL[((b ^ 1) * (NS + 2) + j + 1) * WQ + sl] = tL;
            exR[((b ^ 1) * (NS + 2) + j + 1) * WQ + sl] = tR;
            if (sl == 0) { mnL[(b ^ 1) * (NS + 2) + j + 1] = mL; mnR[(b ^ 1) * (NS + 2) + j + 1] = mR; }
        }
        if (own) {
            const uint32_t sv[4] = {s.x, s.y, s.z, s.w};
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (SAFE3) o[k] = min2(sv[k] + t0[k] + t1[k] + t2[k], 0x7FFF7FFFu);       // no 16-bit overflow possible
                else o[k] = min2(__vadd2(min2(__vadd2(min2(__vadd2(sv[k], t0[k]), 0x7FFF7FFFu), t1[k]), 0x7FFF7FFFu), t2[k]), 0x7FFF7FFFu);
            }
            *sp = make_uint4(o[0], o[1], o[2], o[3]);
        }
        c = cn; s = sn;
        cp += rowq; sp += rowq;
        __syncthreads();
    }
    if (own) {
        uint32_t *F = a.Fout + (size_t)f * a.frame_front + (size_t)x * wordsD + sl * 4;
        *reinterpret_cast<uint4 *>(F) = tL;
        *reinterpret_cast<uint4 *>(F + fw) = Lv;
        *reinterpret_cast<uint4 *>(F + 2 * fw) = tR;
        if (sl == 0) {
            uint32_t *M = a.Mout + (size_t)f * a.frame_front + x;
            M[0] = mL; M[a.W1] = minV; M[2 * a.W1] = mR;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Whole-height pass: the same three paths as sgbm_sweep_kernel, but ONE persistent launch per direction.  The CTAs of a
// thread-block cluster lie side by side across the image (XC = 2 * NS columns each, no halo, no recomputation) and walk
// down (or up) all H rows of a frame together; a cluster takes frames cid, cid + nclusters, ...  What a diagonal path needs
// from the neighbouring CTA -- one column's L (LPC x 16 bytes) and its minimum per row and side -- goes straight into the
// neighbour's exchange buffer through distributed shared memory (st.shared::cluster) and is announced on the neighbour's
// mbarrier (remote arrive.release.cluster / try_wait.acquire.cluster); there is no frontier in global memory and no
// relaunch every 8 rows.  Neighbours stay within one row of each other by construction (each needs the other's previous
// row), which is also the flow control of the double-buffered pads.  A thread owns two columns, NS apart ("halves"); even
// CTAs do the left half first, odd CTAs the right half, so every boundary column is produced half a row before the
// neighbour consumes it and the ~0.2 us exchange latency never sits on the critical path (tools/probe/cluster_probe.cu).
// The last row of a frame hands zeros on (the state of a path that enters the image), so the next frame starts without a
// special case and the row counter T that selects buffers and barrier phases simply runs on.
// ------------------------------------------------------------------------------------------------
struct VPassArgs {
    const uint32_t *C; uint32_t *S; size_t frame_words;
    int W1, H, P1, P2;
    int ystart, ystep;                       // first row and row step of the pass
    int nframes, ncta;                       // frames of this launch; CTAs per cluster (= per frame)
    uint32_t one;                            // 1: a multiplier the compiler cannot fold (see vcore)
    int *err;                                // host-mapped flag: raised when a neighbour's data did not arrive (see `broken`)
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
// remote store that reports its bytes to the remote mbarrier when it has landed: no fence on the sender's side
__device__ __forceinline__ void st_async_v4(uint32_t a, const uint4 v, uint32_t bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar) : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t a, const uint32_t v, uint32_t bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(a), "r"(v), "r"(bar) : "memory");
}
template <int NW>
__device__ __forceinline__ void st_async_pv(uint32_t a, const PV<NW> &v, uint32_t bar)
{
    if constexpr (NW == 4) st_async_v4(a, make_uint4(v.w[0], v.w[1], v.w[2], v.w[3]), bar);
    else {
#pragma unroll
        for (int k = 0; k < NW; k++) st_async_b32(a + 4u * k, v.w[k], bar);
    }
}
__device__ __forceinline__ bool mbar_try_cta(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void l2_prefetch_bulk(const void *p, uint32_t bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory"); }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

constexpr int VP_AHEAD = 3;                  // rows between the L2 prefetch and their use

// NT threads per CTA, a thread owns NQ columns NS = NT / LPC apart ("quarters"; XC = NQ * NS columns per CTA).  1024 x 2 is the
// default; 512 x 4 (RTDM_SGBM_VPASS_SHAPE=1: 128 registers per thread, nothing recomputed) has too few warps to hide the
// shuffle / reduction latencies: 823 against 656 us per MODE_HH frame when both were measured.
template <int LPC, int NW, int NT, int NQ, bool SAFE3>
__global__ void __launch_bounds__(NT, 1)
sgbm_vpass_kernel(VPassArgs a)
{
    constexpr int NS = NT / LPC, XC = NQ * NS, NP = XC + 2;               // slots per quarter, columns per CTA, columns + pads
    constexpr uint32_t CB = LPC * NW * 4u;                                // bytes of one column's L in the exchange buffers
    extern __shared__ __align__(16) uint8_t sw[];
    // exL [2][NP][CB bytes]: L of the path from x-1, stored at producing column + 1; exR the same for the path from x+1;
    // mnL, mnR [2][NP] their minima; then four mbarriers, fullL[2] and fullR[2]
    constexpr uint32_t BUFB = NP * CB, BUFM = NP * 4u;
    constexpr uint32_t OFF_EXR = 2u * BUFB, OFF_MNL = 2u * OFF_EXR, OFF_MNR = OFF_MNL + 2u * BUFM, OFF_BAR = OFF_MNR + 2u * BUFM;
    constexpr uint32_t TXB = CB + 4u;                                 // bytes a neighbour sends per row and side: one column's L + its minimum
    const int tid = threadIdx.x, j = tid / LPC, sl = tid % LPC, grp = (tid & 31) / LPC;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int ncta = a.ncta, ncl = gridDim.x / ncta, cid = blockIdx.x / ncta;
    constexpr int wordsD = NW * LPC;
    for (uint32_t i = tid * 16u; i < OFF_BAR; i += NT * 16u) *reinterpret_cast<uint4 *>(sw + i) = make_uint4(0u, 0u, 0u, 0u);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sw);
    const bool hasL = rank > 0u, hasR = (int)rank + 1 < ncta;
    if (tid == 0) {
        // one arrival (the receiving group's own arrive.expect_tx) + TXB bytes of st.async per phase; phase 0 is armed here
        for (int b = 0; b < 4; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sbase + OFF_BAR + 8u * b), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int b = 0; b < 4; b++)
            if (b < 2 ? hasL : hasR)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sbase + OFF_BAR + 8u * b), "r"(TXB) : "memory");
    }
    __syncthreads();
    cluster_sync_all();
    // q = position in the row's schedule (compile time); even CTAs go through their quarters left to right, odd CTAs right to
    // left, so column 0 / XC - 1 of a CTA is produced NQ - 1 steps before the neighbour consumes it
    const bool odd = (rank & 1u) != 0u;
    bool inq[NQ];
    int bq[NQ];                                  // 1: this thread's column at position q is column 0 (talks to the left neighbour), 2: column XC - 1
    uint32_t aq[NQ], mq[NQ];                     // byte offsets of this thread's exL / mnL READ entries (buffer 0) in sw
    // C / S addresses: 32-bit WORD offsets from the volume base (the volumes of a call stay below 2^32 words: launch_sgbm),
    // formed by one IMAD.WIDE; loads are unconditional -- a column outside the image reads column 0 of its CTA and only its
    // stores are predicated
    uint32_t off[NQ];
    const uint32_t framew = (uint32_t)a.frame_words;
    const uint32_t roww = (uint32_t)(a.ystep * a.W1 * wordsD);                             // two's complement for the upward pass
    const uint32_t framestep = (uint32_t)ncl * framew - (uint32_t)(a.H - 1) * roww;         // last row of a frame -> first row of this cluster's next
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int quarter = odd ? NQ - 1 - q : q, col = quarter * NS + j;
        inq[q] = (int)rank * XC + col < a.W1;
        bq[q] = (quarter == 0 && hasL && j == 0) ? 1 : ((quarter == NQ - 1 && hasR && j == NS - 1) ? 2 : 0);
        aq[q] = (uint32_t)(col * LPC + sl) * (NW * 4u);
        mq[q] = OFF_MNL + (uint32_t)col * 4u;
        off[q] = (uint32_t)cid * framew + (uint32_t)(a.ystart * a.W1 + (int)rank * XC + (inq[q] ? col : 0)) * (uint32_t)wordsD + (uint32_t)(sl * NW);
    }
    const uint32_t remR = hasR ? mapa_u32(sbase, rank + 1u) : 0u, remL = hasL ? mapa_u32(sbase, rank - 1u) : 0u;
    const uint32_t P1x2 = (uint32_t)a.P1 * 0x00010001u, P2x2 = (uint32_t)a.P2 * 0x00010001u;
    const uint32_t one = a.one;                                       // 1, opaque to the compiler (see vcore)
    const PV<NW> z = pv_zero<NW>();
    PV<NW> Lv[NQ];
    uint32_t minV[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) { Lv[q] = z; minV[q] = 0u; }
    uint32_t T = 0u;
    // a wait ran out (the neighbour is at most a row behind: 2^18 failed polls mean the exchange is broken): stop waiting, so
    // that the launch ends, and raise the host-mapped flag -- the next call or wait on the handle fails with -EIO
    bool broken = false;
    PV<NW> c = z, s = z;
    if (cid < a.nframes) { c = pv_ldg<NW>(a.C + off[0]); s = pv_ld<NW>(a.S + off[0]); }

    // one image row; LAST: the frame's last row hands zeros on (the state of a path that enters the image), so the next frame
    // starts without a special case and the row counter T that selects buffers and barrier phases simply runs on
    auto row = [&](auto last_tag, int f, int r, bool lastframe) {
        constexpr bool LAST = decltype(last_tag)::value;
        const uint32_t b = T & 1u;
        const uint32_t ph = ((T - (b ? 1u : 2u)) >> 1) & 1u;          // phase of buffer b's barriers that row T consumes (T >= 1)
        const uint32_t rb = b * BUFB, wb = BUFB - rb, rm = b * BUFM, wm = BUFM - rm;       // read / write buffer offsets
        if (tid == 0) {
            // this CTA's segment of row r + VP_AHEAD (C and S: contiguous [column][d]) is pulled into L2 by the copy engine;
            // the register prefetch one step ahead then only pays L2 latency
            int pr = r + VP_AHEAD, pf = f;
            while (pr >= a.H) { pr -= a.H; pf += ncl; }                   // (frames of fewer than VP_AHEAD rows)
            if (pf < a.nframes) {
                const size_t o = ((size_t)pf * a.frame_words) + ((size_t)(a.ystart + pr * a.ystep) * a.W1 + (size_t)((int)rank * XC)) * wordsD;
                const uint32_t bytes = (uint32_t)min(XC, a.W1 - (int)rank * XC) * wordsD * 4u;
                l2_prefetch_bulk(a.C + o, bytes);
                l2_prefetch_bulk(a.S + o, bytes);
            }
        }
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            // next step's C and S: the next column of this row, or the first column of the next row / next frame
            const uint32_t on = q + 1 < NQ ? off[q + 1] : (!LAST ? off[0] + roww : (lastframe ? off[0] : off[0] + framestep));
            const PV<NW> cn = pv_ldg<NW>(a.C + on), sn = pv_ld<NW>(a.S + on);
            if (bq[q] != 0 && T > 0u && !broken) {
                const uint32_t bar = sbase + OFF_BAR + 8u * ((bq[q] == 1 ? 0u : 2u) + b);
                int spin = 0;
                while (!mbar_try_cta(bar, ph)) if (++spin > (1 << 18)) { broken = true; *reinterpret_cast<volatile int *>(a.err) = 1; break; }
                // armed again for the row after next (the neighbour's next send into this buffer)
                if (sl == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TXB) : "memory");
            }
            const PV<NW> pL = pv_ld<NW>(reinterpret_cast<const uint32_t *>(sw + aq[q] + rb));                      // column - 1's entry
            const PV<NW> pR = pv_ld<NW>(reinterpret_cast<const uint32_t *>(sw + aq[q] + rb + OFF_EXR + 2u * CB));   // column + 1's entry
            const uint32_t pmL = *reinterpret_cast<const uint32_t *>(sw + mq[q] + rm);
            const uint32_t pmR = *reinterpret_cast<const uint32_t *>(sw + mq[q] + rm + 2u * BUFM + 8u);
            uint32_t t0[NW], t1[NW], t2[NW], m0, m1, m2;
            vcoreN<LPC, NW>(pL, pmL, c, sl, one, P1x2, P2x2, t0, m0);
            vcoreN<LPC, NW>(Lv[q], minV[q], c, sl, one, P1x2, P2x2, t1, m1);
            vcoreN<LPC, NW>(pR, pmR, c, sl, one, P1x2, P2x2, t2, m2);
            if (inq[q]) {
                PV<NW> o;
#pragma unroll
                for (int i = 0; i < NW; i++) {
                    if (SAFE3) o.w[i] = min2(t2[i] * one + (t1[i] * one + (t0[i] * one + s.w[i])), 0x7FFF7FFFu);   // no 16-bit overflow possible
                    else o.w[i] = min2(__vadd2(min2(__vadd2(min2(__vadd2(s.w[i], t0[i]), 0x7FFF7FFFu), t1[i]), 0x7FFF7FFFu), t2[i]), 0x7FFF7FFFu);
                }
                pv_st<NW>(a.S + off[q], o);
            }
            PV<NW> tL, tR;
            uint32_t mL, mR;
            if (LAST) { tL = z; tR = z; Lv[q] = z; mL = 0u; mR = 0u; minV[q] = 0u; }
            else {
                mL = group_min_u32<LPC>(m0, grp);
                minV[q] = group_min_u32<LPC>(m1, grp);
                mR = group_min_u32<LPC>(m2, grp);
#pragma unroll
                for (int i = 0; i < NW; i++) { tL.w[i] = t0[i]; Lv[q].w[i] = t1[i]; tR.w[i] = t2[i]; }
            }
            if (inq[q]) {
                pv_st<NW>(reinterpret_cast<uint32_t *>(sw + aq[q] + wb + CB), tL);                         // own entry: slot column + 1
                pv_st<NW>(reinterpret_cast<uint32_t *>(sw + aq[q] + wb + OFF_EXR + CB), tR);
                if (sl == 0) {
                    *reinterpret_cast<uint32_t *>(sw + mq[q] + wm + 4u) = mL;
                    *reinterpret_cast<uint32_t *>(sw + mq[q] + wm + 2u * BUFM + 4u) = mR;
                }
            }
            if (bq[q] != 0 && !(LAST && lastframe)) {
                if (bq[q] == 2) {                // column XC - 1: its rightward diagonal is the right neighbour's left pad (slot 0)
                    const uint32_t bar = remR + OFF_BAR + 8u * (b ^ 1u);
                    st_async_pv<NW>(remR + wb + (uint32_t)sl * (NW * 4u), tL, bar);
                    if (sl == 0) st_async_b32(remR + OFF_MNL + wm, mL, bar);
                } else {                         // column 0: its leftward diagonal is the left neighbour's right pad (slot XC + 1)
                    const uint32_t bar = remL + OFF_BAR + 8u * (2u + (b ^ 1u));
                    st_async_pv<NW>(remL + OFF_EXR + wb + (uint32_t)((XC + 1) * LPC + sl) * (NW * 4u), tR, bar);
                    if (sl == 0) st_async_b32(remL + OFF_MNR + wm + (uint32_t)(XC + 1) * 4u, mR, bar);
                }
            }
            c = cn; s = sn;
        }
        T++;
        __syncthreads();
    };
    for (int f = cid; f < a.nframes; f += ncl) {
        const bool lastframe = f + ncl >= a.nframes;
        for (int r = 0; r < a.H - 1; r++) {
            row(std::false_type(), f, r, lastframe);
#pragma unroll
            for (int q = 0; q < NQ; q++) off[q] += roww;
        }
        row(std::true_type(), f, a.H - 1, lastframe);
#pragma unroll
        for (int q = 0; q < NQ; q++) off[q] += framestep;
    }
    cluster_sync_all();
}

// ------------------------------------------------------------------------------------------------
// left-right check and sub-pixel fit from the per-pixel records of the fused last path; one CTA per (row, frame).
// Same arithmetic as the second half of sgbm_wta_kernel.
// ------------------------------------------------------------------------------------------------
struct LrArgs {
    const uint2 *rec; size_t frame_rec;
    PlaneS16 out;
    int W, H, D, minD, minX1, maxX1, W1, d12;
};

__global__ void __launch_bounds__(256)
sgbm_lr_kernel(LrArgs a)
{
    extern __shared__ __align__(16) uint8_t ws[];
    uint32_t *key2 = reinterpret_cast<uint32_t *>(ws);                 // [W]  (minS << 16 | 0xFFFF - x1)
    int16_t *dval = reinterpret_cast<int16_t *>(key2 + a.W);           // [W]  sub-pixel disparity or INV
    int16_t *best = dval + a.W;                                        // [W1] integer disparity index
    const int y = blockIdx.x, f = blockIdx.y;
    const int INV = a.minD - 1, INVS = INV * 16;
    for (int x = threadIdx.x; x < a.W; x += blockDim.x) { key2[x] = 0xFFFFFFFFu; dval[x] = (int16_t)INVS; }
    __syncthreads();
    const uint2 *rrow = a.rec + (size_t)f * a.frame_rec + (size_t)y * a.W1;
    for (int x = threadIdx.x; x < a.W1; x += blockDim.x) {
        const uint2 r = rrow[x];
        const int minS = (int)(r.x & 0xFFFFu), code = (int)(r.x >> 16);
        if (code == 0xFFFF) { best[x] = -1; continue; }
        const int bd = code & 0x7FFF;
        best[x] = (int16_t)bd;
        if (code & 0x8000) continue;
        const int x2 = x + a.minX1 - bd - a.minD;
        if (x2 >= 0 && x2 < a.W) atomicMin(&key2[x2], ((uint32_t)minS << 16) | (uint32_t)(0xFFFF - x));
        int d = bd;
        if (0 < d && d < a.D - 1) {
            const int sm = (int)(r.y & 0xFFFFu), sp = (int)(r.y >> 16);
            const int den = max(sm + sp - 2 * minS, 1);
            d = d * 16 + ((sm - sp) * 16 + den) / (den * 2);
        } else d *= 16;
        dval[x + a.minX1] = (int16_t)(d + a.minD * 16);
    }
    __syncthreads();
    int16_t *orow = a.out.p + (size_t)f * a.out.frame + (size_t)y * a.out.pitch;
    for (int x = threadIdx.x; x < a.W; x += blockDim.x) {
        int d1 = dval[x];
        if (x >= a.minX1 && x < a.maxX1 && d1 != INVS) {
            const int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
            const int _x = x - _d, x_ = x - d_;
            bool bad = true;
            if (0 <= _x && _x < a.W) {
                const uint32_t k = key2[_x];
                const int d2 = (k == 0xFFFFFFFFu) ? INVS : (int)best[0xFFFF - (int)(k & 0xFFFFu)] + a.minD;
                bad = bad && d2 >= a.minD && abs(d2 - _d) > a.d12;
            } else bad = false;
            if (0 <= x_ && x_ < a.W) {
                const uint32_t k = key2[x_];
                const int d2 = (k == 0xFFFFFFFFu) ? INVS : (int)best[0xFFFF - (int)(k & 0xFFFFu)] + a.minD;
                bad = bad && d2 >= a.minD && abs(d2 - d_) > a.d12;
            } else bad = false;
            if (bad) d1 = INVS;
        }
        orow[x] = (int16_t)d1;
    }
}

// ------------------------------------------------------------------------------------------------
// WTA + uniqueness + sub-pixel + LR check, one CTA per (row, frame)
// ------------------------------------------------------------------------------------------------
struct WtaArgs {
    const uint16_t *S; size_t frame_vol;
    PlaneS16 out;
    int W, H, D, minD, minX1, maxX1, W1, uniq, d12;
};

__global__ void __launch_bounds__(256)
sgbm_wta_kernel(WtaArgs a)
{
    extern __shared__ __align__(16) uint8_t ws[];
    uint32_t *key2 = reinterpret_cast<uint32_t *>(ws);                 // [W]  (minS << 16 | 0xFFFF - x1)
    int16_t *dval = reinterpret_cast<int16_t *>(key2 + a.W);           // [W]  sub-pixel disparity or INV
    int16_t *best = dval + a.W;                                        // [W1] integer disparity index
    const int y = blockIdx.x, f = blockIdx.y;
    const int INV = a.minD - 1, INVS = INV * 16;
    const int D = a.D;
    for (int x = threadIdx.x; x < a.W; x += blockDim.x) { key2[x] = 0xFFFFFFFFu; dval[x] = (int16_t)INVS; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint16_t *Srow = a.S + (size_t)f * a.frame_vol + (size_t)y * a.W1 * D;
    const int wordsD = D / 2;
    // two pixels per warp iteration, S read as u16x2 words (lane l reads words l, l+32, l+64, l+96)
    for (int xb = warp; xb < a.W1; xb += 2 * nw) {
        uint32_t v[2][4];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int x = xb + q * nw;
            const uint32_t *Sw = reinterpret_cast<const uint32_t *>(Srow + (size_t)min(x, a.W1 - 1) * D);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int wi = lane + 32 * k;
                v[q][k] = wi < wordsD ? Sw[wi] : 0xFFFFFFFFu;
            }
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int x = xb + q * nw;
            if (x >= a.W1) break;                                   // warp-uniform
            const uint16_t *Sp = Srow + (size_t)x * D;
            uint32_t kmin = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t d0 = 2u * (uint32_t)(lane + 32 * k);
                kmin = min(kmin, min(((v[q][k] & 0xFFFFu) << 16) | d0, (v[q][k] & 0xFFFF0000u) | (d0 + 1u)));
            }
            kmin = __reduce_min_sync(0xFFFFFFFFu, kmin);
            const int minS = (int)(kmin >> 16), bd = (int)(kmin & 0xFFFFu);
            if (minS >= 32767) { if (lane == 0) best[x] = -1; continue; }      // degenerate (outside the domain)
            // S[d] * (100 - uniq) < minS * 100   <=>   S[d] < T  with  T = ceil(minS * 100 / (100 - uniq))   (uniq < 100);
            // lanes beyond D hold 0xFFFF and never qualify
            bool viol = false;
            const int mul = 100 - a.uniq;
            if (mul > 0) {
                const uint32_t T = (uint32_t)((minS * 100 + mul - 1) / mul);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int d0 = 2 * (lane + 32 * k);
                    const uint32_t lo = v[q][k] & 0xFFFFu, hi = v[q][k] >> 16;
                    viol = viol || (lo < T && (unsigned)(bd - d0 + 1) > 2u);          // |bd - d0| > 1
                    viol = viol || (hi < T && (unsigned)(bd - d0) > 2u);              // |bd - (d0 + 1)| > 1
                }
            } else {
                const int lim = minS * 100;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int d0 = 2 * (lane + 32 * k);
                    if (d0 < D) {
                        viol = viol || ((int)(v[q][k] & 0xFFFFu) * mul < lim && abs(bd - d0) > 1);
                        viol = viol || ((int)(v[q][k] >> 16) * mul < lim && abs(bd - d0 - 1) > 1);
                    }
                }
            }
            viol = __any_sync(0xFFFFFFFFu, viol);
            if (lane == 0) {
                best[x] = (int16_t)bd;
                if (!viol) {
                    const int x2 = x + a.minX1 - bd - a.minD;
                    if (x2 >= 0 && x2 < a.W) atomicMin(&key2[x2], ((uint32_t)minS << 16) | (uint32_t)(0xFFFF - x));
                    int d = bd;
                    if (0 < d && d < D - 1) {
                        const int sm = Sp[d - 1], sp = Sp[d + 1], s0 = Sp[d];
                        const int den = max(sm + sp - 2 * s0, 1);
                        d = d * 16 + ((sm - sp) * 16 + den) / (den * 2);
                    } else d *= 16;
                    dval[x + a.minX1] = (int16_t)(d + a.minD * 16);
                }
            }
        }
    }
    __syncthreads();
    int16_t *orow = a.out.p + (size_t)f * a.out.frame + (size_t)y * a.out.pitch;
    for (int x = threadIdx.x; x < a.W; x += blockDim.x) {
        int d1 = dval[x];
        if (x >= a.minX1 && x < a.maxX1 && d1 != INVS) {
            const int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
            const int _x = x - _d, x_ = x - d_;
            bool bad = true;
            if (0 <= _x && _x < a.W) {
                const uint32_t k = key2[_x];
                const int d2 = (k == 0xFFFFFFFFu) ? INVS : (int)best[0xFFFF - (int)(k & 0xFFFFu)] + a.minD;
                bad = bad && d2 >= a.minD && abs(d2 - _d) > a.d12;
            } else bad = false;
            if (0 <= x_ && x_ < a.W) {
                const uint32_t k = key2[x_];
                const int d2 = (k == 0xFFFFFFFFu) ? INVS : (int)best[0xFFFF - (int)(k & 0xFFFFu)] + a.minD;
                bad = bad && d2 >= a.minD && abs(d2 - d_) > a.d12;
            } else bad = false;
            if (bad) d1 = INVS;
        }
        orow[x] = (int16_t)d1;
    }
}

}  // namespace

static bool sgbm_fused_cost(const SgbmGeom &g) { return (g.D == 48 || g.D == 64 || g.D == 96 || g.D == 128 || g.D == 192 || g.D == 256) && g.bs <= 7 && g.W1 > 0; }

// disparity counts of the specialised aggregation kernels: D = 2 * NW * LPC (PV<NW> per lane, LPC lanes per pixel)
static bool sgbm_fast_d(int D, int *lpc = nullptr, int *nw = nullptr)
{
    int l = 0, k = 0;
    switch (D) {
        case 48: l = 8; k = 3; break;
        case 64: l = 8; k = 4; break;
        case 96: l = 16; k = 3; break;
        case 128: l = 16; k = 4; break;
        case 192: l = 32; k = 3; break;
        default: return false;                  // (256 = 32 lanes x 8: the pass's exchange buffers would need 266 KB)
    }
    if (lpc) *lpc = l;
    if (nw) *nw = k;
    return true;
}

size_t sgbm_work_bytes(const SgbmGeom &g, size_t *planes, size_t *vol)
{
    const size_t Wp = align_up((size_t)g.W, 16);
    size_t pl = (size_t)12 * g.H * Wp;
    if (sgbm_fused_cost(g))                                            // GL + GR blocks of sgbm_planes2_kernel
        pl = std::max(pl, (size_t)g.H * ((size_t)(g.W1 + 2 * PADL) * 32 + (size_t)12 * align_up((size_t)g.W + 2 * PADR, 8) * 2));
    if (g.W1 > 0 && sgbm_fast_d(g.D))                                   // WTA records + two sweep frontiers reuse the block
        pl = std::max(pl, align_up((size_t)8 * g.H * g.W1, 256) + 2 * align_up(((size_t)3 * g.W1 * (g.D / 2) + (size_t)3 * g.W1) * 4, 256));
    pl = align_up(pl, 256);
    const size_t v = (size_t)g.H * (g.W1 > 0 ? g.W1 : 0) * g.D;        // elements
    if (planes) *planes = pl;
    if (vol) *vol = v;
    return pl + 2 * v * sizeof(uint16_t);
}

// CTAs per frame of one row-sweep launch (0 when the sweep kernels do not apply): the host layer sizes its sub-batches so
// that frames x this fills whole waves of SMs (one 1024-thread CTA per SM)
int sgbm_sweep_ctas_per_frame(const SgbmGeom &g)
{
    if (!(g.D == 64 || g.D == 128) || g.bs > 7) return 0;
    const int LPC = g.D / 8, NS = 1024 / LPC, X = NS - 2 * SW_R;
    return cdiv(g.W1, X);
}

// Whole-height pass (sgbm_vpass_kernel): CTAs per frame = cluster size, dynamic shared memory, and how many clusters the
// device keeps resident at once (0: the pass does not apply -- other D, more than 16 CTAs per row, RTDM_SGBM_NOVPASS).
namespace {
struct VPassPlan { int ncta, nclusters; size_t smem; };

template <int LPC, int NW, int NT, int NQ, bool SAFE3>
int vpass_config(int ncta, size_t smem, int *nclusters)
{
    RTDM_CUDA(cudaFuncSetAttribute(sgbm_vpass_kernel<LPC, NW, NT, NQ, SAFE3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (ncta > 8) RTDM_CUDA(cudaFuncSetAttribute(sgbm_vpass_kernel<LPC, NW, NT, NQ, SAFE3>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)ncta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)ncta); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, sgbm_vpass_kernel<LPC, NW, NT, NQ, SAFE3>, &cfg) != cudaSuccess) { cudaGetLastError(); ncl = 0; }
    *nclusters = ncl;
    return 0;
}

template <int LPC, int NW, int NT, int NQ, bool SAFE3>
int vpass_launch(const VPassArgs &a, int nclusters, size_t smem, cudaStream_t st)
{
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)a.ncta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)(a.ncta * nclusters)); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cfg.attrs = at; cfg.numAttrs = 1;
    RTDM_CUDA(cudaLaunchKernelEx(&cfg, sgbm_vpass_kernel<LPC, NW, NT, NQ, SAFE3>, a));
    return 0;
}

bool sgbm_safe3(const SgbmGeom &g)
{
    const int pixmax = 2 * g.ftzero + 63, Lmax = 2 * g.P2 + g.bs * g.bs * pixmax;
    return 32767 + 3 * Lmax <= 65535;
}

// one switch over (D, shape, clamp variant) for the plan and the launch
#define RTDM_VPASS_DISPATCH(FN, ...)                                                                                                      \
    ([&]() -> int {                                                                                                                       \
        const bool alt = g.sw.sgbm_vpass_shape == 1;                                                                                      \
        switch (g.D) {                                                                                                                    \
            case 128: return alt ? (safe3 ? FN<16, 4, 512, 4, true>(__VA_ARGS__) : FN<16, 4, 512, 4, false>(__VA_ARGS__))                 \
                                 : (safe3 ? FN<16, 4, 1024, 2, true>(__VA_ARGS__) : FN<16, 4, 1024, 2, false>(__VA_ARGS__));              \
            case 64: return safe3 ? FN<8, 4, 1024, 2, true>(__VA_ARGS__) : FN<8, 4, 1024, 2, false>(__VA_ARGS__);                          \
            case 48: return safe3 ? FN<8, 3, 1024, 2, true>(__VA_ARGS__) : FN<8, 3, 1024, 2, false>(__VA_ARGS__);                          \
            case 96: return safe3 ? FN<16, 3, 1024, 2, true>(__VA_ARGS__) : FN<16, 3, 1024, 2, false>(__VA_ARGS__);                        \
            default: return safe3 ? FN<32, 3, 1024, 4, true>(__VA_ARGS__) : FN<32, 3, 1024, 4, false>(__VA_ARGS__);    /* 192 */          \
        }                                                                                                                                 \
    })()

int vpass_plan(const SgbmGeom &g, VPassPlan *p)
{
    p->ncta = 0; p->nclusters = 0; p->smem = 0;
    int LPC = 0, NW = 0;
    if (!sgbm_fast_d(g.D, &LPC, &NW) || g.W1 <= 0 || g.sw.sgbm_oldpath || g.sw.sgbm_nosweep || g.sw.sgbm_novpass) return 0;
    if (g.P2 + g.bs * g.bs * (2 * g.ftzero + 63) >= 32768) return 0;               // see `fast` in launch_sgbm
    // columns per CTA: 1024 threads x 2 columns per thread (D = 192, one pixel per warp: x 4); 512 x 4 gives the same
    const int XC = (LPC == 32 ? 4 : 2) * 1024 / LPC, NP = XC + 2;
    const int ncta = cdiv(g.W1, XC);
    if (ncta > 16) return 0;
    const size_t smem = (size_t)2 * 2 * NP * LPC * NW * 4 + (size_t)2 * 2 * NP * 4 + 32;   // exchange, minima, 4 mbarriers
    int dev = 0, optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || smem > (size_t)optin) {
        cudaGetLastError();
        return 0;
    }
    const bool safe3 = sgbm_safe3(g);
    int ncl = 0;
    const int rc = RTDM_VPASS_DISPATCH(vpass_config, ncta, smem, &ncl);
    if (rc) return rc;
    if (ncl < 1) return 0;
    if (g.sw.sgbm_vpass_maxcl > 0) ncl = std::min(ncl, g.sw.sgbm_vpass_maxcl);
    p->ncta = ncta; p->nclusters = ncl; p->smem = smem;
    return 0;
}
}  // namespace

// Does a batch of n frames take the whole-height pass?  One cluster per frame: small batches leave most SMs idle and are
// quicker as tiled sweeps (720p x 128, 15 clusters: 8 frames 990 against 1198 us/frame, 4 frames 1692 against 1193), so
// the pass starts at half the resident clusters unless RTDM_SGBM_VPASS_MIN says otherwise.
static bool vpass_wanted(const SgbmGeom &g, const VPassPlan &p, int n)
{
    if (p.nclusters < 1) return false;
    const int least = g.sw.sgbm_vpass_min > 0 ? g.sw.sgbm_vpass_min : std::max(2, (p.nclusters + 1) / 2);
    return n >= least;
}

// frames one whole-height pass of an n-frame batch works on at the same time (0 when such a batch takes the tiled sweeps):
// the host layer makes its sub-batches multiples of it
int sgbm_vpass_frames_in_flight(const SgbmGeom &g, int n)
{
    VPassPlan p;
    if (vpass_plan(g, &p)) { cudaGetLastError(); return 0; }
    return vpass_wanted(g, p, n) ? p.nclusters : 0;
}

int launch_sgbm(const SgbmGeom &g, int n, PlaneU8 left, PlaneU8 right, PlaneS16 out, SgbmWork w,
                cudaStream_t st, int *launches)
{
    if (n <= 0 || g.W1 <= 0) return 0;
    const int Wp = (int)align_up((size_t)g.W, 16);
    const int h = g.bs / 2;
    const size_t row_words = (size_t)g.W1 * g.D / 2, frame_words = w.frame_vol / 2;
    // D = 48 / 64 / 96 / 128 / 192 / 256 and windows up to 7 (larger ones would spill the register rings)
    const bool fusedcost = sgbm_fused_cost(g) && !g.sw.sgbm_oldcost;
    if (fusedcost) {
        // 1-3. planes in staging format, then fused BT cost + horizontal and vertical windows + P2 -> C
        const int LW = g.W1 + 2 * PADL, WR = (int)align_up((size_t)g.W + 2 * PADR, 8);
        sgbm_planes2_kernel<<<dim3(cdiv(g.W, 8 * 30), g.H, 2 * n), 256, 0, st>>>(left, right, w.planes, w.frame_planes, g.W, g.H,
                                                                            g.ftzero, g.minX1, g.W1, LW, WR);
        Cost2Args a;
        a.planes = w.planes; a.frame_planes = w.frame_planes;
        a.C = reinterpret_cast<uint16_t *>(w.C); a.frame_vol = w.frame_vol;
        a.W = g.W; a.H = g.H; a.D = g.D; a.minD = g.minD; a.minX1 = g.minX1; a.W1 = g.W1; a.LW = LW; a.WR = WR; a.BY = 48;
        a.P2x2 = (uint32_t)g.P2 * 0x00010001u;
        const int D2 = g.D / 2, NT = (256 / D2) * D2, TXk = (256 / D2) * CSEG, NXC = TXk + 2 * h;
        const size_t smem = 2 * ((size_t)NXC * 32 + (size_t)12 * RSTRIDE * 2);
        // bands of 120 rows where that still leaves several waves of CTAs (the 2h rows above a band are recomputed: 3 % instead of
        // 8 %; 597 -> 591 us per 720p MODE_HH frame), 48 rows for single frames and small batches
        if ((long long)cdiv(g.W1, TXk) * cdiv(g.H, 120) * n >= 1200) a.BY = 120;
        const dim3 grid(cdiv(g.W1, TXk), cdiv(g.H, a.BY), n);
#define RTDM_COST_CASE(BS_)                                                                                             \
        case BS_:                                                                                                       \
            RTDM_CUDA(cudaFuncSetAttribute(sgbm_cost_fused_kernel<BS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            sgbm_cost_fused_kernel<BS_><<<grid, NT, smem, st>>>(a);                                                     \
            break;
        switch (g.bs) {
            RTDM_COST_CASE(1) RTDM_COST_CASE(3) RTDM_COST_CASE(5)
            default: RTDM_CUDA(cudaFuncSetAttribute(sgbm_cost_fused_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                     sgbm_cost_fused_kernel<7><<<grid, NT, smem, st>>>(a); break;
        }
#undef RTDM_COST_CASE
        if (launches) (*launches) += 2;
    } else {
    // 1. planes
    sgbm_planes_kernel<<<dim3(cdiv(g.W, 256), g.H, 2 * n), 256, 0, st>>>(left, right, w.planes, w.frame_planes, g.W, g.H, Wp, g.ftzero);
    // 2. BT cost + horizontal window -> Hs (stored in the S volume, which is rewritten by the first path)
    {
        CostArgs a;
        a.planes = w.planes; a.frame_planes = w.frame_planes;
        a.Hs = reinterpret_cast<uint16_t *>(w.S); a.frame_vol = w.frame_vol;
        a.W = g.W; a.H = g.H; a.Wp = Wp; a.D = g.D; a.minD = g.minD; a.h = h; a.minX1 = g.minX1; a.W1 = g.W1;
        const int NXC = TX + 2 * h;
        size_t smem = (size_t)NXC * g.D * 2 + (size_t)NXC * 6 * 4 + (size_t)12 * (NXC + g.D + 8) * 2 + 16;
        if (smem > 48 * 1024)
            RTDM_CUDA(cudaFuncSetAttribute(sgbm_cost_hsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sgbm_cost_hsum_kernel<<<dim3(cdiv(g.W1, TX), g.H, n), 256, smem, st>>>(a);
    }
    // 3. vertical window + P2 -> C
    {
        const int band = 90;
        dim3 grid((unsigned)((row_words + 255) / 256), cdiv(g.H, band), n);
        const uint32_t *hs = reinterpret_cast<const uint32_t *>(w.S);
        uint32_t *cc = reinterpret_cast<uint32_t *>(w.C);
        const uint32_t p2 = (uint32_t)g.P2 * 0x00010001u;
        switch (g.bs) {
            case 1: sgbm_vsum_kernel<1><<<grid, 256, 0, st>>>(hs, cc, frame_words, row_words, g.H, band, p2); break;
            case 3: sgbm_vsum_kernel<3><<<grid, 256, 0, st>>>(hs, cc, frame_words, row_words, g.H, band, p2); break;
            case 5: sgbm_vsum_kernel<5><<<grid, 256, 0, st>>>(hs, cc, frame_words, row_words, g.H, band, p2); break;
            case 7: sgbm_vsum_kernel<7><<<grid, 256, 0, st>>>(hs, cc, frame_words, row_words, g.H, band, p2); break;
            case 9: sgbm_vsum_kernel<9><<<grid, 256, 0, st>>>(hs, cc, frame_words, row_words, g.H, band, p2); break;
            default: sgbm_vsum_kernel<11><<<grid, 256, 0, st>>>(hs, cc, frame_words, row_words, g.H, band, p2); break;
        }
    }
    if (launches) (*launches) += 3;
    }
    RTDM_CUDA(cudaGetLastError());
    // 4. paths.  D = 64 / 128: the 4-words-per-lane kernel, horizontal right-to-left path last and fused with the
    // winner-take-all (S is complete there); other D: generic kernel + separate WTA
    static const int dirs[8][2] = {{-1, 0}, {-1, -1}, {0, -1}, {1, -1}, {-1, 1}, {0, 1}, {1, 1}, {1, 0}};
    const bool hh = g.mode == RTDM_SGBM_MODE_HH;
    const int ndirs = hh ? 8 : 5;
    const int K2 = g.D <= 64 ? 1 : (g.D <= 128 ? 2 : 4);   // u16x2 words per lane (divides D/2 for any D % 16 == 0)
    // the 4-word kernels add S and L as whole words: needs L <= C <= P2 + bs^2 * (2 * ftzero + 63) < 32768 next to S <= 32767
    // (any sane setting; the defaults give 4725), otherwise the generic chain kernel runs
    int LPC = 0, NW = 0;
    const bool fast = sgbm_fast_d(g.D, &LPC, &NW) && !g.sw.sgbm_oldpath && g.P2 + g.bs * g.bs * (2 * g.ftzero + 63) < 32768;
    const bool fused = fast && g.uniq < 100 && !g.sw.sgbm_nofuse;
    const size_t frame_rec = w.frame_planes / 8;            // the BT planes are dead by now: their buffer takes the records
    // the two vertical triplets as row sweeps (one C read and one S update for three paths)
    const size_t rec_bytes = align_up((size_t)8 * g.H * g.W1, 256);
    const size_t front_words = (size_t)3 * g.W1 * (g.D / 2), fmin_words = (size_t)3 * g.W1;
    const size_t front_bytes = align_up((front_words + fmin_words) * 4, 256);
    // (tiled sweeps: 8 disparities per lane only; the 6-disparity forms run per-direction chains where the whole-height pass does not apply)
    const bool sweep = fast && NW == 4 && !g.sw.sgbm_nosweep && rec_bytes + 2 * front_bytes <= w.frame_planes;
    const bool safe3 = sgbm_safe3(g);
    // whole-height cluster pass for batches that fill at least half of the resident clusters
    VPassPlan vp = {0, 0, 0};
    if (fast && n >= 2 && w.err) {
        const int rc = vpass_plan(g, &vp);
        if (rc) return rc;
        if (!vpass_wanted(g, vp, n) || (unsigned long long)n * frame_words >= (1ull << 32)) vp.nclusters = 0;         // 32-bit word offsets
    }
    auto launch_vpass = [&](int dy) -> int {
        VPassArgs a;
        a.C = reinterpret_cast<const uint32_t *>(w.C); a.S = reinterpret_cast<uint32_t *>(w.S); a.frame_words = frame_words;
        a.W1 = g.W1; a.H = g.H; a.P1 = g.P1; a.P2 = g.P2;
        a.ystart = dy > 0 ? 0 : g.H - 1; a.ystep = dy; a.nframes = n; a.ncta = vp.ncta; a.one = 1u; a.err = w.err;
        const int ncl = std::min(vp.nclusters, n);
        const int rc = RTDM_VPASS_DISPATCH(vpass_launch, a, ncl, vp.smem, st);
        if (launches) (*launches)++;
        return rc;
    };
    auto launch_sweep = [&](int dy) -> int {
        if (vp.nclusters > 0) return launch_vpass(dy);
        const int ntiles = cdiv(g.H, SW_R);
        const int LPC = g.D / 8, NS = 1024 / LPC, X = NS - 2 * SW_R;
        const size_t smem = (size_t)2 * 2 * (NS + 2) * LPC * 16 + (size_t)2 * 2 * (NS + 2) * 4;
        for (int t = 0; t < ntiles; t++) {
            SweepArgs a;
            a.C = reinterpret_cast<const uint32_t *>(w.C); a.S = reinterpret_cast<uint32_t *>(w.S); a.frame_words = frame_words;
            // frontier buffers live behind the WTA records inside every frame's (dead) planes block
            uint8_t *fb = w.planes + rec_bytes;
            uint32_t *F0 = reinterpret_cast<uint32_t *>(fb), *F1 = reinterpret_cast<uint32_t *>(fb + front_bytes);
            a.Fin = (t & 1) ? F0 : F1; a.Fout = (t & 1) ? F1 : F0;
            a.Min = a.Fin + front_words; a.Mout = a.Fout + front_words;
            a.frame_front = w.frame_planes / 4;
            a.W1 = g.W1; a.H = g.H; a.P1 = g.P1; a.P2 = g.P2;
            a.ystart = dy > 0 ? t * SW_R : g.H - 1 - t * SW_R; a.ystep = dy; a.nrows = std::min(SW_R, g.H - t * SW_R);
            a.first = t == 0; a.one = 1u;
            const dim3 grid(cdiv(g.W1, X), n);
#define RTDM_SWEEP(LPC_, SAFE_)                                                                                              \
            do {                                                                                                             \
                if (t == 0) RTDM_CUDA(cudaFuncSetAttribute(sgbm_sweep_kernel<LPC_, SAFE_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                sgbm_sweep_kernel<LPC_, SAFE_><<<grid, 1024, smem, st>>>(a);                                                 \
            } while (0)
            if (g.D == 128) { if (safe3) RTDM_SWEEP(16, true); else RTDM_SWEEP(16, false); }
            else { if (safe3) RTDM_SWEEP(8, true); else RTDM_SWEEP(8, false); }
#undef RTDM_SWEEP
        }
        if (launches) (*launches) += ntiles;
        return 0;
    };
    for (int k = 0; k < ndirs; k++) {
        const int di = (!hh && k == 4) ? 7 : k;
        if ((sweep || vp.nclusters > 0) && di >= 1 && di <= 6) {
            if (di == 1 || di == 4) { const int rc = launch_sweep(di == 1 ? 1 : -1); if (rc) return rc; }
            continue;                                       // di = 2, 3 / 5, 6 ride along
        }
        PathArgs a;
        a.C = reinterpret_cast<const uint32_t *>(w.C); a.S = reinterpret_cast<uint32_t *>(w.S); a.frame_words = frame_words;
        a.W1 = g.W1; a.H = g.H; a.D = g.D; a.P1 = g.P1; a.P2 = g.P2; a.px = dirs[di][0]; a.py = dirs[di][1];
        a.first = (k == 0);
        a.rec = reinterpret_cast<uint2 *>(w.planes); a.frame_rec = frame_rec; a.mul = 100 - g.uniq; a.one = 1u;
        const int sx = -a.px, sy = -a.py;
        a.nchains = (sy != 0 ? g.W1 : 0) + (sx != 0 ? (sy != 0 ? g.H - 1 : g.H) : 0);
        if (fast) {
            const int mode = k == 0 ? 0 : ((fused && k == ndirs - 1) ? 2 : 1);
            const int cpc = 4 * (32 / LPC);                   // chains per 128-thread CTA
            const dim3 grid(cdiv(a.nchains, cpc), n);
#define RTDM_PATH(LPC_, NW_)                                                                   \
            do {                                                                               \
                if (mode == 0) sgbm_path4_kernel<LPC_, NW_, 0><<<grid, 128, 0, st>>>(a);       \
                else if (mode == 1) sgbm_path4_kernel<LPC_, NW_, 1><<<grid, 128, 0, st>>>(a);  \
                else sgbm_path4_kernel<LPC_, NW_, 2><<<grid, 128, 0, st>>>(a);                 \
            } while (0)
            switch (g.D) {
                case 128: RTDM_PATH(16, 4); break;
                case 64: RTDM_PATH(8, 4); break;
                case 48: RTDM_PATH(8, 3); break;
                case 96: RTDM_PATH(16, 3); break;
                default: RTDM_PATH(32, 3); break;           // 192
            }
#undef RTDM_PATH
        } else {
            // sub-warp chains where D/2 words split evenly into 8 or 16 lanes of 4 words (D = 64, 128)
            // (only when there are enough chains to still fill the machine: ~9.5k resident warps)
            const bool many = (long long)a.nchains * n >= 2 * 9472;
            if (g.D == 128 && many) { sgbm_path_kernel<4, 16><<<dim3(cdiv(a.nchains, 8), n), 128, 0, st>>>(a); }
            else if (g.D == 64 && many) { sgbm_path_kernel<4, 8><<<dim3(cdiv(a.nchains, 16), n), 128, 0, st>>>(a); }
            else {
                dim3 grid(cdiv(a.nchains, 4), n);
                switch (K2) {
                    case 1: sgbm_path_kernel<1, 32><<<grid, 128, 0, st>>>(a); break;
                    case 2: sgbm_path_kernel<2, 32><<<grid, 128, 0, st>>>(a); break;
                    default: sgbm_path_kernel<4, 32><<<grid, 128, 0, st>>>(a); break;
                }
            }
        }
        if (launches) (*launches)++;
    }
    RTDM_CUDA(cudaGetLastError());
    // 5. WTA (or, after the fused last path, only the left-right check)
    const size_t smem = (size_t)g.W * 4 + (size_t)g.W * 2 + (size_t)g.W1 * 2 + 16;
    if (fused) {
        LrArgs a;
        a.rec = reinterpret_cast<const uint2 *>(w.planes); a.frame_rec = frame_rec; a.out = out;
        a.W = g.W; a.H = g.H; a.D = g.D; a.minD = g.minD; a.minX1 = g.minX1; a.maxX1 = g.maxX1; a.W1 = g.W1; a.d12 = g.d12;
        sgbm_lr_kernel<<<dim3(g.H, n), 256, smem, st>>>(a);
    } else {
        WtaArgs a;
        a.S = reinterpret_cast<const uint16_t *>(w.S); a.frame_vol = w.frame_vol; a.out = out;
        a.W = g.W; a.H = g.H; a.D = g.D; a.minD = g.minD; a.minX1 = g.minX1; a.maxX1 = g.maxX1; a.W1 = g.W1;
        a.uniq = g.uniq; a.d12 = g.d12;
        sgbm_wta_kernel<<<dim3(g.H, n), 256, smem, st>>>(a);
    }
    if (launches) (*launches)++;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rtdm
