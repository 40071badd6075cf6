// bm_sad2.cu -- fast path of the Konolige block-matching core (minDisparity == 0, blockSize 5..15).
//
// Same arithmetic as bm_sad.cu (oracle: orc_bm_core), reorganised around the
// measured sm_100a pipe rates (PRMT / SHF / VABSDIFF4 / VIMNMX all issue at 64 lanes/clk/SM, plain adds
// can also go to the FMA pipe):
//
//   CTA = stripe of TW computed columns x band of BH rows of one frame, NT2 threads, sweeping down rows.
//   Per row:
//   loader   : the next prefiltered row enters a 16-row shared-memory ring as "virtual" rows:
//              Lv[c] = L'[lcol(c)] (clamps applied here), Rv[k] = R'[x0 - h + k]; a second copy shifted by
//              2 bytes keeps every thread's word loads 4-byte aligned with compile-time funnel shifts.
//   producer : thread (column group, disparity octet) owns CT = KT*2h ADJACENT virtual columns x 8
//              disparities.  Vertical window sums V (packed u16x2) slide down in registers:
//              V += |L-R|(row y+h) - |L-R|(row y-h-1), 4 disparities per VABSDIFF4.  Per group of G = 2h
//              columns it emits the in-group prefix sums Pre[c] and suffix sums Suf[c].
//              A (2h+1)-column window always spans exactly two groups, so
//                  SAD(x, d) = Suf[x][d] + Pre[x + 2h][d]         (no sliding start-up, no SAD buffer)
//   consumer : one thread per pixel: per disparity octet 2 x LDS.128 + 4 adds, packed u16x2 minima,
//              octet-level (min << 16 | octet) keys -> argmin octet, exact position inside the octet,
//              texture / uniqueness (octet minima + exact check of the <= 2 octets touching the argmin
//              neighbourhood) / sub-pixel, then disparity and cost are written.
//   The texture sums come from a small separable kernel (bm_texture_kernel).
// The cost volume never leaves the SM; HBM traffic is the two prefiltered images in and disparity + cost out.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace rtdm {
namespace {

// rows in the shared-memory ring: rows y-h-1 .. y+h are live while row y+h+1 is prefetched -> 2h + 3
__host__ __device__ constexpr int ring_rows(int h) { return 2 * h + 3 <= 16 ? 16 : 32; }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

struct Bm2Args {
    PlaneU8 Lp, Rp;
    PlaneS16 disp, cost;
    const uint16_t *tex; size_t tex_pitch, tex_frame;      // texture window sums, [frame][y][x1]
    int W, H, nd, cap, texThr, uniq;
    int W1, row0, row1;
    int TW, BH, NO, NGT;         // stripe width, band height, octets, thread groups
    int LVP, RVP, PP;            // smem pitches (bytes): Lv row copy, Rv row copy, Pre/Suf column
    int dbg;                     // timing experiments only: 1 = skip consumer, 2 = skip producer, 4 = skip loader
};

// ------------------------------------------------------------------------------------------------
// texture: T(x, y) = sum over the (2h+1)^2 window of |L'[y+j][lcol(x+i)] - cap|, x in [0, W1)
// ------------------------------------------------------------------------------------------------
// No shared memory and no barriers: a thread owns 4 adjacent output columns of a band of TXH rows.  The horizontal
// window sums of one image row come from the 4 + 2h bytes around them (6 aligned words, one uniform funnel shift,
// VABSDIFF4 against cap, IDP.4A byte sums); the vertical window slides by adding the entering row's sums and
// subtracting the leaving row's, both recomputed (the rows are L1 / L2 resident).  |L' - cap| <= 63, so the byte
// sums fit the signed dot product and T <= 15 * 15 * 63 fits 16 bits.
constexpr int TXH = 32, TXT = 128;
template <int H_>
__device__ __forceinline__ void tex_row_sums(const uint8_t *row, int col0, int W, uint32_t capx4, int s[4])
{
    constexpr int N = 2 * H_ + 1;                       // window taps; bytes needed: N + 3 <= 18
    uint32_t w[5];
    if (col0 + N + 2 <= W - 1) {
        const uint32_t *p = reinterpret_cast<const uint32_t *>(row + (col0 & ~3));
        const int sh = (col0 & 3) * 8;
        uint32_t q[6];
#pragma unroll
        for (int k = 0; k < 6; k++) q[k] = p[k];
#pragma unroll
        for (int k = 0; k < 5; k++) w[k] = __funnelshift_r(q[k], q[k + 1], sh);
    } else {
        // right image border: columns beyond W - 1 replicate it
#pragma unroll
        for (int k = 0; k < 5; k++) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) v |= (uint32_t)row[min(col0 + 4 * k + b, W - 1)] << (8 * b);
            w[k] = v;
        }
    }
#pragma unroll
    for (int k = 0; k < 5; k++) w[k] = __vabsdiffu4(w[k], capx4);
    // s[0] = bytes 0 .. N-1
    int acc = 0;
#pragma unroll
    for (int k = 0; k < (N + 3) / 4; k++) {
        const int nb = N - 4 * k >= 4 ? 4 : N - 4 * k;
        const int m = nb == 4 ? 0x01010101 : (nb == 3 ? 0x00010101 : (nb == 2 ? 0x00000101 : 0x00000001));
        acc = __dp4a((int)w[k], m, acc);
    }
    s[0] = acc;
#pragma unroll
    for (int i = 1; i < 4; i++) {
        // + byte (N - 1 + i), - byte (i - 1)
        const int bi = N - 1 + i, bo = i - 1;
        acc = __dp4a((int)w[bi / 4], 1 << (8 * (bi % 4)), acc);
        acc = __dp4a((int)w[bo / 4], (int)(0xFFu << (8 * (bo % 4))), acc);    // multiplier byte -1
        s[i] = acc;
    }
}

template <int H_>
__global__ void __launch_bounds__(TXT)
bm_texture_kernel(PlaneU8 Lp, uint16_t *tex, size_t tex_pitch, size_t tex_frame, int W, int H, int nd,
                  int cap, int W1, int row0, int row1)
{
    const int f = blockIdx.z;
    const int x1 = (blockIdx.x * TXT + threadIdx.x) * 4;
    const int y0 = row0 + blockIdx.y * TXH, y1 = min(y0 + TXH, row1);
    if (x1 >= W1 || y0 >= y1) return;
    const int col0 = x1 + (nd - 1) - H_;                 // real column of window tap 0 (>= 0: nd - 1 >= 15 > H_)
    const uint8_t *L = Lp.p + (size_t)f * Lp.frame;
    const uint32_t capx4 = (uint32_t)cap * 0x01010101u;
    int T[4] = {0, 0, 0, 0}, s[4];
    for (int r = y0 - H_; r <= y0 + H_; r++) {
        tex_row_sums<H_>(L + (size_t)clampi(r, 0, H - 1) * Lp.pitch, col0, W, capx4, s);
#pragma unroll
        for (int i = 0; i < 4; i++) T[i] += s[i];
    }
    uint16_t *out = tex + (size_t)f * tex_frame + (size_t)y0 * tex_pitch + x1;
    const bool vec = ((tex_pitch & 3) == 0) && x1 + 3 < W1;
    for (int y = y0; y < y1; y++) {
        if (vec) *reinterpret_cast<uint2 *>(out) = make_uint2((uint32_t)T[0] | ((uint32_t)T[1] << 16), (uint32_t)T[2] | ((uint32_t)T[3] << 16));
        else
            for (int i = 0; i < 4 && x1 + i < W1; i++) out[i] = (uint16_t)T[i];
        out += tex_pitch;
        if (y + 1 < y1) {
            tex_row_sums<H_>(L + (size_t)clampi(y + 1 + H_, 0, H - 1) * Lp.pitch, col0, W, capx4, s);
#pragma unroll
            for (int i = 0; i < 4; i++) T[i] += s[i];
            tex_row_sums<H_>(L + (size_t)clampi(y - H_, 0, H - 1) * Lp.pitch, col0, W, capx4, s);
#pragma unroll
            for (int i = 0; i < 4; i++) T[i] -= s[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// one ring word of the loader: where it comes from in the prefiltered image row
// ------------------------------------------------------------------------------------------------
struct RingWord {
    int dst;        // byte offset inside a ring slot (Lv area followed by Rv area), -1 = no word
    int src;        // aligned byte offset inside the image row (left or right plane), -1 = clamped gather
    int sh;         // funnel shift (bits) for the aligned pair, or first virtual column for a gather
    int right;      // 1: right image, 0: left image
};

// NT2 = threads per CTA; SUF_ = also store in-group suffix sums (2 x LDS.128 per octet in the consumer
// instead of 3, at the price of twice the producer stores)
template <int H_, int KT_, int NT2, bool SUF_, int MINB = 2>
__global__ void __launch_bounds__(NT2, MINB)
bm_sad2_kernel(Bm2Args a)
{
    constexpr int G = 2 * H_, CT = G * KT_, RING = ring_rows(H_);
    constexpr int NLW = (CT + 3) / 4, NRW = ((CT - 1) >> 2) + 3;
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, f = blockIdx.z;
    const int nd = a.nd, NO = a.NO;
    const int x0 = blockIdx.x * a.TW;
    const int TWc = min(a.TW, a.W1 - x0);
    const int y0 = a.row0 + blockIdx.y * a.BH, y1 = min(y0 + a.BH, a.row1);
    if (TWc <= 0 || y0 >= y1) return;
    const int NCT = a.NGT * CT;                        // virtual columns held by the producers

    // ---- shared memory carve-up ----------------------------------------------------------------
    // Pre[buf][c][d]: in-group prefix sums of the vertical sums; column NCT is all zero
    const size_t PBUF = (size_t)(NCT + 1) * a.PP;
    uint8_t *Pre = smem;
    uint8_t *Suf = smem + PBUF;                                                // only with SUF_
    uint16_t *Smin = reinterpret_cast<uint16_t *>(smem + (SUF_ ? 2 : 1) * PBUF);   // [NO][NT2]
    uint8_t *Ring = reinterpret_cast<uint8_t *>(Smin + (size_t)NO * NT2);      // [RING][2*LVP + 2*RVP]
    const int SLOT = 2 * a.LVP + 2 * a.RVP;

    const uint8_t *Lg = a.Lp.p + (size_t)f * a.Lp.frame;
    const uint8_t *Rg = a.Rp.p + (size_t)f * a.Rp.frame;
    const int lofs = nd - 1;

    // ---- loader descriptors: each thread owns up to 2 words of every ring row ---------------------
    RingWord rw_desc[2];
    {
        const int lwords = a.LVP / 4, rwords = a.RVP / 4;
        const int total = 2 * lwords + 2 * rwords;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int i = tid + q * NT2;
            RingWord d; d.dst = -1; d.src = -1; d.sh = 0; d.right = 0;
            if (i < total) {
                d.dst = 4 * i;
                if (i < 2 * lwords) {
                    const int cpy = i >= lwords, w = cpy ? i - lwords : i;
                    const int c = 4 * w + 2 * cpy;                     // first virtual column of the word
                    const int xa = x0 - H_ + c;
                    if (xa >= -lofs && xa + 3 <= a.W - lofs - 1) { const int g = xa + lofs; d.src = g & ~3; d.sh = (g & 3) * 8; }
                    else d.sh = c;
                } else {
                    const int k = i - 2 * lwords;
                    const int cpy = k >= rwords, w = cpy ? k - rwords : k;
                    const int c = 4 * w + 2 * cpy;
                    const int xa = x0 - H_ + c;
                    d.right = 1;
                    if (xa >= 0 && xa + 3 <= a.W - 1) { d.src = xa & ~3; d.sh = (xa & 3) * 8; }
                    else d.sh = c;
                }
            }
            rw_desc[q] = d;
        }
    }
    // fetch (global -> registers) and commit (registers -> ring slot) of one image row
    // fetch keeps the RAW aligned word pairs in registers (no use of the loaded values, so the loads stay in
    // flight behind the producer work); commit funnel-shifts them into the ring slot
    auto fetch_row = [&](int gy, uint32_t (&v)[4]) {
        const int gyc = clampi(gy, 0, a.H - 1);
        const uint8_t *lsrc = Lg + (size_t)gyc * a.Lp.pitch, *rsrc = Rg + (size_t)gyc * a.Rp.pitch;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const RingWord d = rw_desc[q];
            uint32_t w0 = 0, w1 = 0;
            if (d.dst >= 0) {
                const uint8_t *src = d.right ? rsrc : lsrc;
                if (d.src >= 0) {
                    const uint32_t *pw = reinterpret_cast<const uint32_t *>(src + d.src);
                    w0 = pw[0]; w1 = pw[1];
                } else {
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int xa = x0 - H_ + d.sh + b;
                        const int col = d.right ? clampi(xa, 0, a.W - 1) : clampi(xa, -lofs, a.W - lofs - 1) + lofs;
                        w0 |= (uint32_t)src[col] << (8 * b);
                    }
                }
            }
            v[2 * q] = w0; v[2 * q + 1] = w1;
        }
    };
    auto commit_row = [&](int gy, const uint32_t (&v)[4]) {
        uint8_t *slot = Ring + (size_t)(gy & (RING - 1)) * SLOT;
#pragma unroll
        for (int q = 0; q < 2; q++)
            if (rw_desc[q].dst >= 0)
                *reinterpret_cast<uint32_t *>(slot + rw_desc[q].dst) =
                    rw_desc[q].src >= 0 ? __funnelshift_r(v[2 * q], v[2 * q + 1], rw_desc[q].sh) : v[2 * q];
    };

    // ---- producer task of this thread -------------------------------------------------------------
    const int tg = tid / NO, j = tid - tg * NO;
    const bool prod = tg < a.NGT;
    const int c0 = tg * CT;                             // first virtual column of the thread
    const int cpy = (c0 & 2) ? 1 : 0;                   // which shifted copy keeps the word loads aligned
    const int lbo = cpy * a.LVP + (c0 - 2 * cpy);                       // byte offset of column c0 in the slot
    const int rbo = 2 * a.LVP + cpy * a.RVP + (c0 - 2 * cpy) + 8 * j;   // byte offset of R byte (c0 + 8j)
    // R clamp (App. A.2, minD = 0): rbase(xc) = clip(xc, 0, W - nd); in virtual columns c = xc - x0 + h
    const int cmin = H_ - x0, cmax = (a.W - nd) - x0 + H_;
    const bool border = (c0 < cmin) || (c0 + CT - 1 > cmax);

    uint32_t V[CT][4];
#pragma unroll
    for (int kk = 0; kk < CT; kk++) V[kk][0] = V[kk][1] = V[kk][2] = V[kk][3] = 0u;

    // |L - R| of column kk (8 disparities) of ring row gy
    auto ad_col = [&](const uint32_t (&lw)[NLW], const uint32_t (&rw)[NRW], const uint8_t *slot, int kk, uint32_t &lo, uint32_t &hi) {
        if (!border) {
            const int w = kk >> 2, sft = kk & 3;
            const uint32_t l4 = __byte_perm(lw[w], 0, sft == 0 ? 0x0000 : (sft == 1 ? 0x1111 : (sft == 2 ? 0x2222 : 0x3333)));
            uint32_t r0, r1;
            if (sft == 0) { r0 = rw[w]; r1 = rw[w + 1]; }
            else { r0 = __funnelshift_r(rw[w], rw[w + 1], 8 * sft); r1 = __funnelshift_r(rw[w + 1], rw[w + 2], 8 * sft); }
            lo = __vabsdiffu4(l4, r0);
            hi = __vabsdiffu4(l4, r1);
        } else {
            // clamped columns read the R window of the nearest unclamped column (unshifted copy)
            const int c = c0 + kk;
            const int bo = clampi(c, cmin, cmax) + 8 * j;
            const uint32_t *pw = reinterpret_cast<const uint32_t *>(slot + 2 * a.LVP) + (bo >> 2);
            const int sh = (bo & 3) * 8;
            const uint32_t l4 = (uint32_t)slot[c] * 0x01010101u;
            lo = __vabsdiffu4(l4, __funnelshift_r(pw[0], pw[1], sh));
            hi = __vabsdiffu4(l4, __funnelshift_r(pw[1], pw[2], sh));
        }
    };
    auto load_words = [&](const uint8_t *slot, uint32_t (&lw)[NLW], uint32_t (&rw)[NRW]) {
        const uint32_t *lp = reinterpret_cast<const uint32_t *>(slot + lbo);
        const uint32_t *rp = reinterpret_cast<const uint32_t *>(slot + rbo);
#pragma unroll
        for (int i = 0; i < NLW; i++) lw[i] = lp[i];
#pragma unroll
        for (int i = 0; i < NRW; i++) rw[i] = rp[i];
    };

    // ---- prologue: ring rows y0-h .. y0+h, zero column, vertical sums over rows y0-h .. y0+h-1 ----
    for (int r = y0 - H_; r <= y0 + H_; r++) { uint32_t v[4]; fetch_row(r, v); commit_row(r, v); }
    for (int i = tid; i < a.PP / 4; i += NT2) reinterpret_cast<uint32_t *>(Pre + (size_t)NCT * a.PP)[i] = 0u;
    __syncthreads();
    if (prod) {
        for (int r = y0 - H_; r < y0 + H_; r++) {
            const uint8_t *slot = Ring + (size_t)(r & (RING - 1)) * SLOT;
            uint32_t lw[NLW], rw[NRW];
            if (!border) load_words(slot, lw, rw);
#pragma unroll
            for (int kk = 0; kk < CT; kk++) {
                uint32_t lo, hi;
                ad_col(lw, rw, slot, kk, lo, hi);
                V[kk][0] += __byte_perm(lo, 0, 0x4140);
                V[kk][1] += __byte_perm(lo, 0, 0x4342);
                V[kk][2] += __byte_perm(hi, 0, 0x4140);
                V[kk][3] += __byte_perm(hi, 0, 0x4342);
            }
        }
    }

    int16_t *dispf = a.disp.p + (size_t)f * a.disp.frame;
    int16_t *costf = a.cost.p ? a.cost.p + (size_t)f * a.cost.frame : nullptr;
    const uint16_t *texf = a.tex + (size_t)f * a.tex_frame;
    const int16_t FILT = (int16_t)(-16);                // (minD - 1) * 16 with minD = 0

    // Consumer pixels are spread evenly over all warps (lanes 0..PW-1 of each warp; a second round if the
    // stripe is wider than the CTA), so that every warp has the same serial path per row.  (Measured: packing
    // the pixels into fewer, fuller warps is 40 % slower -- the kernel is bound by per-warp latency.)
    int cx = -1;
    {
        const int nw = NT2 / 32, w = tid >> 5, l = tid & 31;
        const int PW = (TWc + nw - 1) / nw;                 // pixels per warp, <= 32 because TW <= NT2
        if (l < PW && w * PW + l < TWc) cx = w * PW + l;
    }

    for (int y = y0; y < y1; y++) {
        // start fetching the next row and this row's texture sum; both are consumed after the producer work
        uint32_t nextv[4];
        const bool have_next = y + 1 < y1;
        if (have_next && !(a.dbg & 4)) fetch_row(y + 1 + H_, nextv);
        int tsum = 0;
        if (cx >= 0) tsum = texf[(size_t)y * a.tex_pitch + x0 + cx];
        // ---------------- producer ---------------------------------------------------------------
        if (prod && !(a.dbg & 2)) {
            const uint8_t *sin = Ring + (size_t)((y + H_) & (RING - 1)) * SLOT;
            const uint8_t *sout = Ring + (size_t)((y - H_ - 1) & (RING - 1)) * SLOT;
            uint32_t lwi[NLW], rwi[NRW], lwo[NLW], rwo[NRW];
            const bool has_out = y > y0;
            if (!border) { load_words(sin, lwi, rwi); if (has_out) load_words(sout, lwo, rwo); }
            uint8_t *pdst = Pre + (size_t)c0 * a.PP + 16 * j;
#pragma unroll
            for (int g = 0; g < KT_; g++) {
                uint4 p = make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int i = 0; i < G; i++) {
                    const int kk = g * G + i;
                    uint32_t lo, hi;
                    ad_col(lwi, rwi, sin, kk, lo, hi);
                    if (has_out) {
                        uint32_t olo, ohi;
                        ad_col(lwo, rwo, sout, kk, olo, ohi);
                        lo = lo + 0x80808080u - olo;                 // per byte: in + 128 - out (no borrow)
                        hi = hi + 0x80808080u - ohi;
                        V[kk][0] += __byte_perm(lo, 0, 0x4140) - 0x00800080u;
                        V[kk][1] += __byte_perm(lo, 0, 0x4342) - 0x00800080u;
                        V[kk][2] += __byte_perm(hi, 0, 0x4140) - 0x00800080u;
                        V[kk][3] += __byte_perm(hi, 0, 0x4342) - 0x00800080u;
                    } else {
                        V[kk][0] += __byte_perm(lo, 0, 0x4140);
                        V[kk][1] += __byte_perm(lo, 0, 0x4342);
                        V[kk][2] += __byte_perm(hi, 0, 0x4140);
                        V[kk][3] += __byte_perm(hi, 0, 0x4342);
                    }
                    p.x += V[kk][0]; p.y += V[kk][1]; p.z += V[kk][2]; p.w += V[kk][3];
                    *reinterpret_cast<uint4 *>(pdst + (size_t)kk * a.PP) = p;
                }
                if (SUF_) {
                    uint8_t *sdst = Suf + (size_t)c0 * a.PP + 16 * j;
                    uint4 q = make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int i = G - 1; i >= 0; i--) {
                        const int kk = g * G + i;
                        q.x += V[kk][0]; q.y += V[kk][1]; q.z += V[kk][2]; q.w += V[kk][3];
                        *reinterpret_cast<uint4 *>(sdst + (size_t)kk * a.PP) = q;
                    }
                }
            }
        }
        if (have_next && !(a.dbg & 4)) commit_row(y + 1 + H_, nextv);
        __syncthreads();

        // ---------------- consumer: one thread per pixel -----------------------------------------------
        if (cx >= 0 && !(a.dbg & 1)) {
            const int x = cx;
            // SAD(x, d) = T_g - Pre[x-1] + Pre[x+G]  (T_g = Pre of the last column of x's group; Pre[-1] = 0)
            const int gi = x % G;
            const uint8_t *tp = SUF_ ? Suf + (size_t)x * a.PP : Pre + (size_t)(x - gi + G - 1) * a.PP;
            const uint8_t *mp = Pre + (size_t)(gi ? x - 1 : NCT) * a.PP;
            const uint8_t *pp = Pre + (size_t)(x + G) * a.PP;
            // SAD of one octet as 4 packed u16x2 words
            auto sad4 = [&](int o, uint32_t (&sv)[4]) {
                const uint4 t = *reinterpret_cast<const uint4 *>(tp + 16 * o);
                const uint4 w = *reinterpret_cast<const uint4 *>(pp + 16 * o);
                if (SUF_) { sv[0] = t.x + w.x; sv[1] = t.y + w.y; sv[2] = t.z + w.z; sv[3] = t.w + w.w; }
                else {
                    const uint4 u = *reinterpret_cast<const uint4 *>(mp + 16 * o);
                    sv[0] = t.x - u.x + w.x; sv[1] = t.y - u.y + w.y; sv[2] = t.z - u.z + w.z; sv[3] = t.w - u.w + w.w;
                }
            };
            int16_t dout = FILT;
            if (tsum >= a.texThr) {
                // pass 1: octet minima and the argmin octet
                uint32_t best = 0xFFFFFFFFu;
                for (int o = 0; o < NO; o++) {
                    uint32_t sv[4];
                    sad4(o, sv);
                    uint32_t m = __vimin3_u16x2(sv[0], sv[1], sv[2]);
                    m = __vminu2(m, sv[3]);
                    const uint32_t mm = min(m & 0xFFFFu, m >> 16);
                    Smin[o * NT2 + tid] = (uint16_t)mm;
                    best = min(best, mm * 65536u + (uint32_t)o);
                }
                const int minsad = (int)(best >> 16), oc = (int)(best & 0xFFFFu);
                const uint16_t *t16 = reinterpret_cast<const uint16_t *>(tp);
                const uint16_t *m16 = reinterpret_cast<const uint16_t *>(mp);
                const uint16_t *p16 = reinterpret_cast<const uint16_t *>(pp);
                // exact position inside the argmin octet (first minimum) via (value << 3 | index) keys
                int mind;
                {
                    uint32_t sv[4];
                    sad4(oc, sv);
                    const uint32_t s0 = sv[0], s1 = sv[1], s2 = sv[2], s3 = sv[3];
                    uint32_t k = __vimin3_u32((s0 & 0xFFFFu) * 8u, (s0 >> 16) * 8u + 1u, (s1 & 0xFFFFu) * 8u + 2u);
                    k = __vimin3_u32(k, (s1 >> 16) * 8u + 3u, (s2 & 0xFFFFu) * 8u + 4u);
                    k = __vimin3_u32(k, (s2 >> 16) * 8u + 5u, (s3 & 0xFFFFu) * 8u + 6u);
                    k = min(k, (s3 >> 16) * 8u + 7u);
                    mind = 8 * oc + (int)(k & 7u);
                }
                const int dp = mind + 1 < nd ? mind + 1 : nd - 2, dn = mind > 0 ? mind - 1 : 1;
                const int p = (int)t16[dp] - (SUF_ ? 0 : (int)m16[dp]) + (int)p16[dp];
                const int n = (int)t16[dn] - (SUF_ ? 0 : (int)m16[dn]) + (int)p16[dn];
                bool ok = true;
                if (a.uniq > 0) {
                    const int thresh = minsad + (minsad * a.uniq / 100);
                    const int zlo = max(mind - 1, 0), zhi = min(mind + 1, nd - 1);
                    const int olo = zlo >> 3, ohi = zhi >> 3;
                    // octets that do not touch [mind-1, mind+1]: their minimum decides
                    Smin[olo * NT2 + tid] = 0xFFFFu;
                    Smin[ohi * NT2 + tid] = 0xFFFFu;
                    uint32_t m2 = 0xFFFFu;
                    for (int o = 0; o < NO; o++) m2 = min(m2, (uint32_t)Smin[o * NT2 + tid]);
                    ok = (int)m2 > thresh;
                    // the (at most two) touching octets: exact check with the neighbourhood masked out
                    for (int oo = olo; ok && oo <= ohi; oo++) {
                        uint32_t sv[4];
                        sad4(oo, sv);
                        const int rel = mind - 8 * oo;                        // -1 .. 8
                        const uint32_t Z = (7u << (rel + 1)) >> 2;            // bit p set: position p is excluded
                        uint32_t mz = 0xFFFFFFFFu;
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const uint32_t msk = (((Z >> (2 * r)) & 1u) ? 0x0000FFFFu : 0u) | (((Z >> (2 * r + 1)) & 1u) ? 0xFFFF0000u : 0u);
                            mz = __vminu2(mz, sv[r] | msk);
                        }
                        ok = (int)min(mz & 0xFFFFu, mz >> 16) > thresh;
                    }
                }
                if (ok) {
                    const int q = p + n - 2 * minsad + abs(p - n);
                    const int v = (nd - mind - 1) * 256 + (q != 0 ? ((p - n) * 256) / q : 0) + 15;
                    dout = (int16_t)(v >> 4);
                    if (costf) costf[(size_t)y * a.cost.pitch + lofs + x0 + x] = (int16_t)minsad;
                }
            }
            dispf[(size_t)y * a.disp.pitch + lofs + x0 + x] = dout;
        }
        __syncthreads();        // Pre is rewritten by the next row's producers
    }
}

struct Tiling2 { int KT, CT, NO, NGT, TW, BH, nstripes, nbands, LVP, RVP, PP, NT, SUF; size_t smem; };

bool pick_tiling2(const BmGeom &g, int n, Tiling2 *t)
{
    const int h = g.bs / 2;
    if (g.minD != 0 || h < 2 || h > 7 || g.nd > 256 || g.nd < 16) return false;
    t->KT = h == 2 ? 3 : (h == 3 ? 2 : 1);
    t->CT = 2 * h * t->KT;
    t->NO = g.nd / 8;
    // variant: RTDM_BM_VARIANT = 0: 192 threads + suffix sums, 1: 192 prefix-only, 2: 256 + suffix, 3: 256 prefix-only
    const int variant = g.sw.bm_variant;
    t->NT = (variant & 2) ? 256 : 192;
    t->SUF = (variant & 1) ? 0 : 1;
    const int NT2 = t->NT;
    t->NGT = std::min(NT2 / t->NO, std::max(1, 192 / t->CT));
    int twmax = std::min(t->NGT * t->CT - 2 * h, NT2);
    if (twmax < 8) return false;
    t->nstripes = cdiv(g.W1, twmax);
    t->TW = cdiv(g.W1, t->nstripes);
    const int rows = g.row1 - g.row0;
    // bands: amortise the 2h-row start-up, but keep enough CTAs in flight for small batches
    int bhmax = 128;
    while (bhmax > 32 && (long long)n * t->nstripes * cdiv(rows, bhmax) < 2 * 148 * 2) bhmax /= 2;
    t->nbands = cdiv(rows, bhmax);
    t->BH = cdiv(rows, t->nbands);
    const int NCT = t->NGT * t->CT;
    t->LVP = (int)align_up(NCT + 8, 4);
    t->RVP = (int)align_up(NCT + g.nd + 16, 4);
    t->PP = g.nd * 2 + 16;
    t->smem = (size_t)(t->SUF ? 2 : 1) * (NCT + 1) * t->PP + (size_t)t->NO * NT2 * 2 + (size_t)ring_rows(h) * 2 * (t->LVP + t->RVP);
    if (2 * (t->LVP / 4) + 2 * (t->RVP / 4) > 2 * NT2) return false;     // loader: <= 2 ring words per thread
    return t->smem <= ((t->NT == 256 && t->SUF) ? 200 : 112) * 1024;
}

template <int H_, int KT_, int NT2, bool SUF_, int MINB = 2>
int launch2v(const Bm2Args &a, const Tiling2 &t, int n, cudaStream_t st)
{
    RTDM_CUDA(cudaFuncSetAttribute(bm_sad2_kernel<H_, KT_, NT2, SUF_, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(t.smem, 48 * 1024)));
    bm_sad2_kernel<H_, KT_, NT2, SUF_, MINB><<<dim3(t.nstripes, t.nbands, n), NT2, t.smem, st>>>(a);
    return 0;
}

template <int H_, int KT_>
int launch2(const Bm2Args &a, const Tiling2 &t, int n, cudaStream_t st, bool occ3)
{
    if (t.NT == 192 && !t.SUF && t.smem <= 74 * 1024 && occ3) return launch2v<H_, KT_, 192, false, 3>(a, t, n, st);
    if (t.NT == 192) return t.SUF ? launch2v<H_, KT_, 192, true>(a, t, n, st) : launch2v<H_, KT_, 192, false>(a, t, n, st);
    return t.SUF ? launch2v<H_, KT_, 256, true>(a, t, n, st) : launch2v<H_, KT_, 256, false>(a, t, n, st);
}

}  // namespace

bool bm_sad2_supported(const BmGeom &g, int n)
{
    Tiling2 t;
    return g.W1 >= 1 && g.row1 > g.row0 && pick_tiling2(g, n, &t);
}

int launch_bm_texture(const BmGeom &g, int n, PlaneU8 Lp, uint16_t *tex, size_t tex_pitch, size_t tex_frame, cudaStream_t st)
{
    const int h = g.bs / 2;
    dim3 grid(cdiv(cdiv(g.W1, 4), TXT), cdiv(g.row1 - g.row0, TXH), n);
#define RTDM_TEX_CASE(H_) bm_texture_kernel<H_><<<grid, TXT, 0, st>>>(Lp, tex, tex_pitch, tex_frame, g.W, g.H, g.nd, g.cap, g.W1, g.row0, g.row1)
    switch (h) {
        case 2: RTDM_TEX_CASE(2); break;
        case 3: RTDM_TEX_CASE(3); break;
        case 4: RTDM_TEX_CASE(4); break;
        case 5: RTDM_TEX_CASE(5); break;
        case 6: RTDM_TEX_CASE(6); break;
        default: RTDM_TEX_CASE(7); break;
    }
#undef RTDM_TEX_CASE
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

// tex: scratch of n * H * tex_pitch uint16
int launch_bm_sad2(const BmGeom &g, int n, PlaneU8 Lp, PlaneU8 Rp, PlaneS16 disp, PlaneS16 cost,
                   uint16_t *tex, size_t tex_pitch, size_t tex_frame, cudaStream_t st, int *launches, bool use3)
{
    Tiling2 t;
    if (!pick_tiling2(g, n, &t)) { set_error("bm_sad2: unsupported geometry"); return -RTDM_EINVAL; }
    const int h = g.bs / 2;
    {
        const int rct = launch_bm_texture(g, n, Lp, tex, tex_pitch, tex_frame, st);
        if (rct) return rct;
    }
    if (use3) {
        const int rc3 = launch_bm_sad3_core(g, n, Lp, Rp, disp, cost, tex, tex_pitch, tex_frame, st);
        if (rc3) return rc3;
        if (launches) (*launches) += 2;
        return 0;
    }
    Bm2Args a;
    a.Lp = Lp; a.Rp = Rp; a.disp = disp; a.cost = cost;
    a.tex = tex; a.tex_pitch = tex_pitch; a.tex_frame = tex_frame;
    a.W = g.W; a.H = g.H; a.nd = g.nd; a.cap = g.cap; a.texThr = g.texThr; a.uniq = g.uniq;
    a.W1 = g.W1; a.row0 = g.row0; a.row1 = g.row1;
    a.dbg = g.sw.bm_debug;
    a.TW = t.TW; a.BH = t.BH; a.NO = t.NO; a.NGT = t.NGT; a.LVP = t.LVP; a.RVP = t.RVP; a.PP = t.PP;
    int rc = 0;
    switch (h) {
        case 2: rc = launch2<2, 3>(a, t, n, st, g.sw.bm_occ3 != 0); break;
        case 3: rc = launch2<3, 2>(a, t, n, st, g.sw.bm_occ3 != 0); break;
        case 4: rc = launch2<4, 1>(a, t, n, st, g.sw.bm_occ3 != 0); break;
        case 5: rc = launch2<5, 1>(a, t, n, st, g.sw.bm_occ3 != 0); break;
        case 6: rc = launch2<6, 1>(a, t, n, st, g.sw.bm_occ3 != 0); break;
        default: rc = launch2<7, 1>(a, t, n, st, g.sw.bm_occ3 != 0); break;
    }
    if (rc) return rc;
    if (launches) (*launches) += 2;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rtdm
