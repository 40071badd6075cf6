// bm_sad3.cu -- warp-specialised Konolige block-matching core (minDisparity == 0, blockSize 5 .. 15,
// numDisparities 32 / 48 / 64 / 96 / 128 / 192 / 256; the reference's default is 192 scaled by the frame width).  Same arithmetic as bm_sad.cu / bm_sad2.cu (oracle:
// orc_bm_core); replaces findStereoCorrespondenceBM as reached from SWMatcherKonolige::compute
// (reference stereo-matcher/bm-sw.cpp:33-38).
//
// bm_sad2's profile (profiles/r01_prof_bm_r1d_summary.csv) shows a latency-bound kernel: 12 warps per SM (158
// registers for 12 columns of window sums per thread), half of the issue slots empty, producer and consumer
// phases that never overlap.  This kernel splits the roles instead of the time:
//
//   one CTA per SM = stripe of TW <= 180 computed columns x band of BH rows of one frame,
//   NG producer warps (one per group of G = 2h virtual columns) + 8 consumer warps, <= 80 registers each.
//
//   producer thread = (half group of h ADJACENT columns, disparity octet): vertical window sums V of its h columns
//       slide down in 4h registers (VABSDIFF4, per-byte in + 128 - out deltas).  The first half of a group
//       ("A") emits in-half PREFIX sums left to right, the second half ("B") emits SUFFIX sums right to left,
//       both with the same straight-line code: B threads read byte-reversed (mirrored) copies of the rows, so
//       their columns and their eight disparities simply come out in reverse order.  The two halves swap their
//       totals with four shuffles and store the group total T in their own order.
//       A (2h+1)-column window always spans two groups, so with i = x mod G
//           i <  h :  SAD(x) = T[g]   - PreA[x - 1]  + PreA'[i]          (all in A order)
//           i >= h :  SAD(x) = SufB[x] + T[g + 1]     - SufB'[i + 1]      (all in B order)
//       i.e. always  a + b - c  with three 16-byte loads per octet, and no halo.
//   consumer thread = one pixel: octet minima (packed u16x2, two octets per word), argmin octet by
//       (min << 16 | octet) keys, exact position, texture / uniqueness / sub-pixel as in bm_sad2.
//   The consumers also run the loader: the next prefiltered row enters a 16-row shared-memory ring as four
//   "virtual" rows (left / right, forward / mirrored; the clamps of App. A.2 are applied here).
//   Producers work on row y + 1 while consumers work on row y: the sums are double buffered and handed over
//   through named barriers (bar.arrive / bar.sync), the only synchronisation per row.
// The cost volume never leaves the SM; HBM traffic is the two prefiltered images in and disparity + cost out.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace rtdm {
namespace {

// CTA shapes: WIDE = one CTA of <= 768 threads per SM (6 winner-take-all + 2 loader warps next to the producers),
// PAIR = two CTAs of <= 384 threads per SM (3 + 1): narrower stripes (more halo), but two independent row pipelines
// fill each other's barrier bubbles
struct ShapeWide { static constexpr int NCW = 8, NLD = 2, MAXT = 768, MINB = 1, MAXW = 5; };    // MAXW: loader items per lane
struct ShapePair { static constexpr int NCW = 4, NLD = 1, MAXT = 384, MINB = 2, MAXW = 6; };

__host__ __device__ constexpr int ring_rows3(int h) { return 2 * h + 4; }     // rows y-h-1 .. y+h+2 are live

__device__ __forceinline__ int clampi3(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct Bm3Args {
    PlaneU8 Lp, Rp;
    PlaneS16 disp, cost;
    const uint16_t *tex; size_t tex_pitch, tex_frame;      // texture window sums, [frame][y][x1]
    int W, H, nd, cap, texThr, uniq;
    int W1, row0, row1;
    int TW, BH, NG;              // stripe width, band height, column groups (= producer warps x groups per warp)
    int dbg;                     // timing experiments only (RTDM_BM_DEBUG): 1 = skip the winner-take-all maths
};

// masks of the uniqueness test: entry rel + 1 (rel = mind - 8 * octet in -1 .. 8) has 0xFFFF in the 16-bit lanes
// of the positions rel - 1 .. rel + 1 that fall inside the octet
__constant__ uint4 c_zmask[10];

// shared-memory geometry shared by host and device
// bytes per pixel of the octet-key rows: nd / 8 keys of 4 bytes, rounded up to whole 128-bit words (the spare keys stay
// 0xFFFFFFFF), padded to an ODD number of 16-byte units so that 128-bit rows of 8 neighbouring pixels hit 8 bank groups
__host__ __device__ constexpr int mnp_bytes(int nd) { return (((nd / 2 + 15) / 16) | 1) * 16; }
struct Geo3 {
    // one buffer: X[NCT] | HA[NG] | HB[NG] | TA[NG] | TB[NG] | zero row (pitch PP each) | Mn[NCT] (pitch MNP)
    int NCT, NCTP, PP, BUFSZ, XOFF, HAOFF, HBOFF, TAOFF, TBOFF, ZOFF, MNOFF;
    int LF, LM, RF, RM, RMX, SLOT;                   // ring slot: byte offsets of the four virtual rows
    int RINGOFF, DESCOFF, NITEM, total;
};
__host__ __device__ inline Geo3 make_geo3(int h, int nd, int NG)
{
    Geo3 q;
    const int G = 2 * h;
    q.NCT = NG * G;
    q.NCTP = (q.NCT + 3) & ~3;                 // virtual-row length in the ring (odd h: NG * 2h need not be a multiple of 4)
    q.PP = nd * 2 + 16;
    q.XOFF = 0;
    q.HAOFF = q.NCT * q.PP;
    q.HBOFF = q.HAOFF + NG * q.PP;
    q.TAOFF = q.HBOFF + NG * q.PP;
    q.TBOFF = q.TAOFF + NG * q.PP;
    q.ZOFF = q.TBOFF + NG * q.PP;
    q.MNOFF = q.ZOFF + q.PP;
    q.BUFSZ = q.MNOFF + q.NCT * mnp_bytes(nd);
    q.LF = 0;                                  // left rows are stored EXPANDED: one word = one pixel x 0x01010101
    q.LM = 4 * q.NCTP;
    q.RF = 8 * q.NCTP;
    const int RFP = q.NCTP + nd + 8;
    q.RM = q.RF + RFP;
    q.RMX = q.NCTP + nd + 2;                   // mirrored right row: RMv[n] = Rv[RMX - n];  RMX == 2 (mod 4)
    const int RMP = q.NCTP + nd + 16;
    q.SLOT = (q.RM + RMP + 15) & ~15;         // 16-byte aligned slots: the expanded left rows move as 128-bit words
    q.RINGOFF = 2 * q.BUFSZ;
    q.DESCOFF = q.RINGOFF + ring_rows3(h) * q.SLOT;
    q.DESCOFF = (q.DESCOFF + 15) & ~15;
    q.NITEM = q.NCTP / 2 + RFP / 4 + RMP / 4;   // loader items: one source word each (left: 4 expanded words out)
    q.total = q.DESCOFF + q.NITEM * 8;
    return q;
}

// one loader item = four consecutive virtual-row bytes:
//   item index i -> (right?, first forward virtual index cs, byte-reversed?, destination byte offset in the slot);
//   left items are written as four expanded words (16 bytes), right items as one word
__device__ __forceinline__ void ring_item_src(const Geo3 &q, int i, int &right, int &cs, int &rev, int &dst)
{
    const int nl = q.NCTP / 4;
    if (i < nl) { right = 0; rev = 0; cs = 4 * i; dst = q.LF + 16 * i; }
    else if (i < 2 * nl) { const int m = i - nl; right = 0; rev = 1; cs = q.NCTP - 4 - 4 * m; dst = q.LM + 16 * m; }
    else {
        const int b = 4 * (i - 2 * nl);               // byte offset from RF
        right = 1; dst = q.RF + b;
        if (q.RF + b < q.RM) { rev = 0; cs = b; }
        else { rev = 1; cs = q.RMX - 3 - (q.RF + b - q.RM); }
    }
}
__device__ __forceinline__ void ring_store(uint8_t *slot, int dst, bool left, uint32_t v)
{
    if (left)
        *reinterpret_cast<uint4 *>(slot + dst) = make_uint4(__byte_perm(v, 0, 0x0000), __byte_perm(v, 0, 0x1111), __byte_perm(v, 0, 0x2222), __byte_perm(v, 0, 0x3333));
    else
        *reinterpret_cast<uint32_t *>(slot + dst) = v;
}

template <int H_, int NO_, class SH>
__global__ void __launch_bounds__(SH::MAXT, SH::MINB)
bm_sad3_kernel(Bm3Args a)
{
    constexpr int NCW = SH::NCW, NLD = SH::NLD;
    constexpr int G = 2 * H_, RING = ring_rows3(H_);
    constexpr int NLW = H_, NRW = ((H_ % 2 ? 2 : 0) + H_ + 7 + 3) / 4;   // left: one expanded word per column; right: 8 + h - 1 bytes from a byte offset of 0 (or 2: odd h)
    constexpr int ND = NO_ * 8, PP = ND * 2 + 16, MNP = mnp_bytes(ND);
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, f = blockIdx.z;
    const int x0 = blockIdx.x * a.TW;
    const int TWc = min(a.TW, a.W1 - x0);
    const int y0 = a.row0 + blockIdx.y * a.BH, y1 = min(y0 + a.BH, a.row1);
    if (TWc <= 0 || y0 >= y1) return;
    const Geo3 q = make_geo3(H_, ND, a.NG);
    const int NPT = 128 * (((a.NG + 1) / 2 + 32 / NO_ - 1) / (32 / NO_));   // producer threads: (A, B) x (even, odd groups) x warps
    const int NT = NPT + NCW * 32;
    uint8_t *Ring = smem + q.RINGOFF;
    int2 *Desc = reinterpret_cast<int2 *>(smem + q.DESCOFF);
    const uint8_t *Lg = a.Lp.p + (size_t)f * a.Lp.frame;
    const uint8_t *Rg = a.Rp.p + (size_t)f * a.Rp.frame;
    const int lofs = ND - 1;

    // real image column of forward virtual index v (left: Lv[v], right: Rv[v]); App. A.2 clamps
    auto src_col = [&](int right, int v) {
        const int xa = x0 - H_ + v;
        return right ? clampi3(xa, 0, a.W - 1) : clampi3(xa, -lofs, a.W - lofs - 1) + lofs;
    };

    // ---- prologue: ring rows y0-h-1 .. y0+h+1 by all threads (byte gathers), zero rows ------------------
    {
        const int nw = q.NITEM, nrows = 2 * H_ + 3;
        for (int i = tid; i < nw * nrows; i += NT) {
            const int r = i / nw, w = i - r * nw;
            const int gy = y0 - H_ - 1 + r;
            const int gyc = clampi3(gy, 0, a.H - 1);
            int right, cs, rev, dst;
            ring_item_src(q, w, right, cs, rev, dst);
            const uint8_t *src = (right ? Rg + (size_t)gyc * a.Rp.pitch : Lg + (size_t)gyc * a.Lp.pitch);
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) v |= (uint32_t)src[src_col(right, cs + b)] << (8 * b);
            if (rev) v = __byte_perm(v, 0, 0x0123);
            ring_store(Ring + (size_t)((gy + RING) % RING) * q.SLOT, dst, !right, v);
        }
        // loader descriptors: x = aligned byte offset of the word pair in the image row (or, for a clamped gather, the
        // first virtual index), y = funnel shift | right image << 8 | byte-reversed << 9 | gather << 10 | dst << 12
        for (int i = tid; i < nw; i += NT) {
            int right, cs, rev, dst;
            ring_item_src(q, i, right, cs, rev, dst);
            const int xa = x0 - H_ + cs;
            const bool inr = right ? (xa >= 0 && xa + 3 <= a.W - 1) : (xa >= -lofs && xa + 3 <= a.W - lofs - 1);
            const int gcol = right ? xa : xa + lofs;
            const int fl = (right << 8) | (rev << 9) | (dst << 12);
            Desc[i] = inr ? make_int2(gcol & ~3, ((gcol & 3) * 8) | fl) : make_int2(cs, fl | 0x400);
        }
        for (int i = tid; i < PP / 4; i += NT) {
            reinterpret_cast<uint32_t *>(smem + q.ZOFF)[i] = 0u;
            reinterpret_cast<uint32_t *>(smem + q.BUFSZ + q.ZOFF)[i] = 0u;
        }
        if (NO_ % 4)                                            // spare keys of the last 128-bit word of every key row
            for (int i = tid; i < q.NCT * (MNP / 4); i += NT) {
                reinterpret_cast<uint32_t *>(smem + q.MNOFF)[i] = 0xFFFFFFFFu;
                reinterpret_cast<uint32_t *>(smem + q.BUFSZ + q.MNOFF)[i] = 0xFFFFFFFFu;
            }
    }
    __syncthreads();

    if (tid < NPT) {
        // =========================================================================================
        // producer
        // =========================================================================================
        // even producer warps hold A halves, odd warps B halves (type-uniform warps: phase 2 differs per type); within a
        // type, warps alternate between even and odd groups, so that the byte alignment of a thread's streams (below)
        // is the same for the whole warp
        constexpr int SPW = 32 / NO_;                             // groups per warp
        const int pw = tid >> 5, isB = pw & 1, widx = pw >> 1, par = widx & 1;
        const int sub = (tid & 31) / NO_, j = (tid & 31) - sub * NO_;
        const int g = 2 * ((widx >> 1) * SPW + sub) + par;
        const int hg = 2 * g + isB;
        const bool live = g < a.NG && sub < SPW;                  // trailing groups may not exist; 32 % NO_ lanes of a warp stay idle
        // byte offsets of the thread's L and R streams inside a ring slot.  The left stream is one word per column; the
        // right stream starts `off` bytes into a word: 0 when 2h is a multiple of 4, else 0 or 2 by group parity
        const int lbo = isB ? q.LM + 4 * (q.NCTP - (g + 1) * G) : q.LF + 4 * g * G;
        const int off = (isB ? q.NCTP - (g + 1) * G : g * G) & 3;                 // warp-uniform
        const int rbo = isB ? q.RM + (q.RMX - (g + 1) * G - 6 - 8 * j) : q.RF + g * G + 8 * j;
        // R clamp (App. A.2, minD = 0): rbase(xc) = clip(xc, 0, W - nd); in virtual columns c = xc - x0 + h
        const int cmin = H_ - x0, cmax = (a.W - ND) - x0 + H_;
        uint32_t clmask = 0;
        int crc = 0;
#pragma unroll
        for (int k = 0; k < H_; k++) {
            const int c = isB ? g * G + G - 1 - k : g * G + k;
            if (c < cmin) { clmask |= 1u << k; crc = cmin; }
            if (c > cmax) { clmask |= 1u << k; crc = cmax; }
        }
        const bool wborder = __any_sync(0xFFFFFFFFu, live && clmask != 0);
        // clamped columns read the R window of the nearest unclamped column: its stream offset in this thread's copy
        const int cbo = isB ? q.RM + (q.RMX - crc - 8 * j - 7) : q.RF + crc + 8 * j;

        uint32_t V[H_][4];
#pragma unroll
        for (int k = 0; k < H_; k++) V[k][0] = V[k][1] = V[k][2] = V[k][3] = 0u;

        auto load_words = [&](const uint8_t *slot, uint32_t (&lw)[NLW], uint32_t (&rw)[NRW]) {
            const uint32_t *rp = reinterpret_cast<const uint32_t *>(slot + (rbo & ~3));
            constexpr int W4 = (H_ % 2 == 0 && H_ >= 4) ? 4 : 0;      // even h: the left stream is 16-byte aligned
            if (W4) {
                const uint4 t = *reinterpret_cast<const uint4 *>(slot + lbo);
                lw[0] = t.x; lw[1] = t.y; lw[2] = t.z; lw[3] = t.w;
            }
#pragma unroll
            for (int i = W4; i + 1 < H_; i += 2) {                    // any h: 8-byte aligned
                const uint2 t = *reinterpret_cast<const uint2 *>(slot + lbo + 4 * i);
                lw[i] = t.x; lw[i + 1] = t.y;
            }
            if ((H_ - W4) & 1) lw[H_ - 1] = *reinterpret_cast<const uint32_t *>(slot + lbo + 4 * (H_ - 1));
#pragma unroll
            for (int i = 0; i < NRW; i++) rw[i] = rp[i];
            if (H_ % 2 == 1 && off) {                                 // odd h, stream starts 2 bytes into its first word (warp-uniform)
#pragma unroll
                for (int i = 0; i < NRW; i++) rw[i] = __funnelshift_r(rw[i], i + 1 < NRW ? rw[i + 1] : 0u, 16);
            }
        };
        auto clamped_window = [&](const uint8_t *slot, uint32_t &c0w, uint32_t &c1w) {
            const uint32_t *pw = reinterpret_cast<const uint32_t *>(slot + (cbo & ~3));
            const int sh = (cbo & 3) * 8;
            c0w = __funnelshift_r(pw[0], pw[1], sh);
            c1w = __funnelshift_r(pw[1], pw[2], sh);
        };
        // |L - R| of column k (8 disparities in stream order)
        auto ad_col = [&](const uint32_t (&lw)[NLW], const uint32_t (&rw)[NRW], int k, bool border, uint32_t c0w, uint32_t c1w,
                          uint32_t &lo, uint32_t &hi) {
            const int w = k >> 2, sft = k & 3;
            const uint32_t l4 = lw[k];
            uint32_t r0, r1;
            if (sft == 0) { r0 = rw[w]; r1 = rw[w + 1]; }
            else { r0 = __funnelshift_r(rw[w], rw[w + 1], 8 * sft); r1 = __funnelshift_r(rw[w + 1], rw[w + 2], 8 * sft); }
            if (border && ((clmask >> k) & 1u)) { r0 = c0w; r1 = c1w; }
            lo = __vabsdiffu4(l4, r0);
            hi = __vabsdiffu4(l4, r1);
        };

        // vertical sums over rows y0-h-1 .. y0+h-1 (the first loop iteration removes row y0-h-1 again)
        for (int r = y0 - H_ - 1; live && r < y0 + H_; r++) {
            const uint8_t *slot = Ring + (size_t)((r + RING) % RING) * q.SLOT;
            uint32_t lw[NLW], rw[NRW], c0w = 0, c1w = 0;
            load_words(slot, lw, rw);
            if (wborder) clamped_window(slot, c0w, c1w);
#pragma unroll
            for (int k = 0; k < H_; k++) {
                uint32_t lo, hi;
                ad_col(lw, rw, k, wborder, c0w, c1w, lo, hi);
                V[k][0] += __byte_perm(lo, 0, 0x4140);
                V[k][1] += __byte_perm(lo, 0, 0x4342);
                V[k][2] += __byte_perm(hi, 0, 0x4140);
                V[k][3] += __byte_perm(hi, 0, 0x4342);
            }
        }

        const int xst = q.XOFF + (hg * H_) * PP + 16 * j;                    // store base of the thread's h slots
        const int hst = (isB ? q.HBOFF : q.HAOFF) + g * PP + 16 * j;          // own half total
        // phase 2: A half of group g makes the pixels gG + k, B half the pixels gG + G - 1 - k; both need group g + 1
        const bool ph2 = live && g + 1 < a.NG;
        const int nxt = q.XOFF + ((g + 1) * G + (isB ? H_ : 0)) * PP + 16 * j; // next group's half of the same type
        const int mst = q.MNOFF + (isB ? g * G + G - 1 : g * G) * MNP + 4 * j;
        const int mstep = isB ? -MNP : MNP;
        auto rev4 = [](uint4 v) {
            return make_uint4(__byte_perm(v.w, 0, 0x1032), __byte_perm(v.z, 0, 0x1032), __byte_perm(v.y, 0, 0x1032), __byte_perm(v.x, 0, 0x1032));
        };
        auto emit_min = [&](uint8_t *buf, int k, uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3) {
            uint32_t m = __vminu2(__vimin3_u16x2(s0, s1, s2), s3);
            m = __vminu2(m, m >> 16);
            *reinterpret_cast<uint32_t *>(buf + mst + k * mstep) = m * 65536u + (uint32_t)j;      // (octet minimum << 16) | octet
        };

        // Even rows of the band add the byte deltas (in + 128 - out), odd rows subtract the mirrored deltas
        // (out + 128 - in): V += d + 128, then V += d - 128.  No per-lane bias correction is needed: after an even
        // row every column sum carries +128, so every (2h+1)-column window sum carries the same (2h+1) * 128, which
        // the consumers subtract from the minimum and the two neighbours (BIASC).
        auto row = [&](int y, auto border_tag, auto odd_tag) {
            constexpr bool BORDER = decltype(border_tag)::value;
            constexpr bool ODD = decltype(odd_tag)::value;
            uint8_t *buf = smem + (y & 1) * q.BUFSZ;
            uint4 p = make_uint4(0, 0, 0, 0);
            if (live) {
                const uint8_t *sin = Ring + (size_t)((y + H_ + RING) % RING) * q.SLOT;
                const uint8_t *sout = Ring + (size_t)((y - H_ - 1 + RING) % RING) * q.SLOT;
                uint32_t lwi[NLW], rwi[NRW], lwo[NLW], rwo[NRW];
                uint32_t ci0 = 0, ci1 = 0, co0 = 0, co1 = 0;
                load_words(sin, lwi, rwi);
                load_words(sout, lwo, rwo);
                if (BORDER) { clamped_window(sin, ci0, ci1); clamped_window(sout, co0, co1); }
                uint8_t *pdst = buf + xst;
#pragma unroll
                for (int k = 0; k < H_; k++) {
                    uint32_t lo, hi, olo, ohi;
                    ad_col(lwi, rwi, k, BORDER, ci0, ci1, lo, hi);
                    ad_col(lwo, rwo, k, BORDER, co0, co1, olo, ohi);
                    if (!ODD) {
                        lo = lo + 0x80808080u - olo;                 // per byte: in + 128 - out (no borrow)
                        hi = hi + 0x80808080u - ohi;
                        V[k][0] += __byte_perm(lo, 0, 0x4140);
                        V[k][1] += __byte_perm(lo, 0, 0x4342);
                        V[k][2] += __byte_perm(hi, 0, 0x4140);
                        V[k][3] += __byte_perm(hi, 0, 0x4342);
                    } else {
                        lo = olo + 0x80808080u - lo;                 // per byte: out + 128 - in
                        hi = ohi + 0x80808080u - hi;
                        V[k][0] -= __byte_perm(lo, 0, 0x4140);
                        V[k][1] -= __byte_perm(lo, 0, 0x4342);
                        V[k][2] -= __byte_perm(hi, 0, 0x4140);
                        V[k][3] -= __byte_perm(hi, 0, 0x4342);
                    }
                    p.x += V[k][0]; p.y += V[k][1]; p.z += V[k][2]; p.w += V[k][3];
                    *reinterpret_cast<uint4 *>(pdst + k * PP) = p;
                }
                *reinterpret_cast<uint4 *>(buf + hst) = p;       // half total, in this half's order
            }
            bar_sync(5, NPT);                                    // all prefix / suffix sums of row y are in shared memory
            if (ph2) {
                if (!isB) {
                    // A pixels: SAD(gG + k) = T[g] - PreA[k - 1] + PreA'[k]
                    const uint4 hb = rev4(*reinterpret_cast<const uint4 *>(buf + q.HBOFF + g * PP + 16 * j));
                    uint4 run = make_uint4(p.x + hb.x, p.y + hb.y, p.z + hb.z, p.w + hb.w);
                    *reinterpret_cast<uint4 *>(buf + q.TAOFF + g * PP + 16 * j) = run;
#pragma unroll
                    for (int k = 0; k < H_; k++) {
                        const uint4 l = *reinterpret_cast<const uint4 *>(buf + nxt + k * PP);
                        emit_min(buf, k, run.x + l.x, run.y + l.y, run.z + l.z, run.w + l.w);
                        run.x -= V[k][0]; run.y -= V[k][1]; run.z -= V[k][2]; run.w -= V[k][3];
                    }
                } else {
                    // B pixels: SAD(gG + G - 1 - k) = SufB[k] + T[g + 1] - SufB'[k - 1]
                    const uint4 ha = rev4(*reinterpret_cast<const uint4 *>(buf + q.HAOFF + (g + 1) * PP + 16 * j));
                    const uint4 hb = *reinterpret_cast<const uint4 *>(buf + q.HBOFF + (g + 1) * PP + 16 * j);
                    uint4 run = make_uint4(ha.x + hb.x, ha.y + hb.y, ha.z + hb.z, ha.w + hb.w);
                    *reinterpret_cast<uint4 *>(buf + q.TBOFF + (g + 1) * PP + 16 * j) = run;
#pragma unroll
                    for (int k = 0; k < H_; k++) {
                        run.x += V[k][0]; run.y += V[k][1]; run.z += V[k][2]; run.w += V[k][3];
                        uint4 l = make_uint4(0, 0, 0, 0);
                        if (k > 0) l = *reinterpret_cast<const uint4 *>(buf + nxt + (k - 1) * PP);
                        emit_min(buf, k, run.x - l.x, run.y - l.y, run.z - l.z, run.w - l.w);
                    }
                }
            }
        };

        for (int y = y0; y < y1; y++) {
            if (y - y0 >= 2) bar_sync(3 + (y & 1), NT);            // consumers are done with this buffer (row y - 2)
            if ((y - y0) & 1) {
                if (wborder) row(y, std::true_type(), std::true_type());
                else row(y, std::false_type(), std::true_type());
            } else {
                if (wborder) row(y, std::true_type(), std::false_type());
                else row(y, std::false_type(), std::false_type());
            }
            bar_arrive(1 + (y & 1), NT);                           // sums and octet minima of row y are complete
        }
    } else {
        // =========================================================================================
        // consumer (and loader)
        // =========================================================================================
        const int ct = tid - NPT, cw = ct >> 5, lane = ct & 31;
        if (cw >= NCW - NLD) {
            // ---- loader warps: row y + h + 2 enters the ring while the consumers work on row y ---------------
            // (rows up to y + h + 1 are there; the slot it overwrites was last read by the producers in row y - 1)
            constexpr int MAXW = SH::MAXW;                      // loader items per lane (NITEM <= 32 * NLD * MAXW)
            const int nw = q.NITEM;
            const int ll = (cw - (NCW - NLD)) * 32 + lane;
            int off[MAXW];
            uint32_t meta[MAXW];                                // Desc.y | valid << 11
#pragma unroll
            for (int s = 0; s < MAXW; s++) {
                const int i = ll + 32 * NLD * s;
                off[s] = 0; meta[s] = 0;
                if (i < nw) { const int2 d = Desc[i]; off[s] = d.x; meta[s] = (uint32_t)d.y | 0x800u; }
            }
            for (int y = y0; y < y1; y++) {
                const bool have_next = y + 1 < y1;
                uint32_t w0[MAXW], w1[MAXW];
                const int gy = y + H_ + 2, gyc = clampi3(gy, 0, a.H - 1);
                const uint8_t *lrow = Lg + (size_t)gyc * a.Lp.pitch, *rrow = Rg + (size_t)gyc * a.Rp.pitch;
                if (have_next) {
#pragma unroll
                    for (int s = 0; s < MAXW; s++) {
                        w0[s] = 0; w1[s] = 0;
                        if (meta[s] & 0x800u) {
                            const uint8_t *src = (meta[s] & 0x100u) ? rrow : lrow;
                            if (!(meta[s] & 0x400u)) {
                                const uint32_t *pw = reinterpret_cast<const uint32_t *>(src + off[s]);
                                w0[s] = pw[0]; w1[s] = pw[1];
                            } else {
#pragma unroll
                                for (int b = 0; b < 4; b++) w0[s] |= (uint32_t)src[src_col((meta[s] >> 8) & 1, off[s] + b)] << (8 * b);
                            }
                        }
                    }
                }
                bar_sync(1 + (y & 1), NT);                      // the producers have finished row y (and its ring reads)
                if (have_next) {
                    uint8_t *slot = Ring + (size_t)((gy + RING) % RING) * q.SLOT;
#pragma unroll
                    for (int s = 0; s < MAXW; s++) {
                        if (meta[s] & 0x800u) {
                            uint32_t v = !(meta[s] & 0x400u) ? __funnelshift_r(w0[s], w1[s], meta[s] & 31u) : w0[s];
                            v = __byte_perm(v, 0, (meta[s] & 0x200u) ? 0x0123u : 0x3210u);
                            ring_store(slot, (int)(meta[s] >> 12), !(meta[s] & 0x100u), v);
                        }
                    }
                }
                if (y + 2 < y1) bar_arrive(3 + (y & 1), NT);
            }
            return;
        }
        // ---- the pixel of this thread --------------------------------------------------------------------
        int x = -1;
        {
            const int PW = (TWc + NCW - NLD - 1) / (NCW - NLD);  // pixels per WTA warp (<= 32)
            if (lane < PW && cw * PW + lane < TWc) x = cw * PW + lane;
        }
        const int xx = max(x, 0);
        const int gq = xx / G, gi = xx - gq * G;
        const bool flip = gi >= H_;
        int oa, ob, oc_;                                        // SAD = [oa] + [ob] - [oc_], byte offsets inside a buffer
        if (!flip) {
            oa = q.TAOFF + gq * PP;
            ob = q.XOFF + ((gq + 1) * G + gi) * PP;
            oc_ = gi > 0 ? q.XOFF + (xx - 1) * PP : q.ZOFF;
        } else {
            oa = q.XOFF + (gq * G + H_ + (G - 1 - gi)) * PP;
            ob = q.TBOFF + (gq + 1) * PP;
            oc_ = gi < G - 1 ? q.XOFF + ((gq + 1) * G + H_ + (G - 2 - gi)) * PP : q.ZOFF;
        }
        const uint32_t usel = flip ? 0x5476u : 0x3210u;         // un-reverse selector: (sv[r], sv[3 - r]) -> forward word r
        const int xflip = flip ? 7 : 0;
        const int omn = q.MNOFF + xx * MNP;                     // octet minima of this pixel (made by the producers)
        int16_t *dptr = a.disp.p + (size_t)f * a.disp.frame + (size_t)y0 * a.disp.pitch + lofs + x0 + xx;
        int16_t *cptr = a.cost.p ? a.cost.p + (size_t)f * a.cost.frame + (size_t)y0 * a.cost.pitch + lofs + x0 + xx : nullptr;
        const uint16_t *tptr = a.tex + (size_t)f * a.tex_frame + (size_t)y0 * a.tex_pitch + x0 + xx;
        const int16_t FILT = (int16_t)(-16);                    // (minD - 1) * 16 with minD = 0
        int tsum = x >= 0 ? (int)*tptr : 0;

        for (int y = y0; y < y1; y++) {
            // texture sum of the next row
            const bool have_next = y + 1 < y1;
            int tnext = 0;
            if (have_next && x >= 0) tnext = (int)tptr[a.tex_pitch];
            tptr += a.tex_pitch;

            bar_sync(1 + (y & 1), NT);                          // the sums of row y are complete
            const uint8_t *buf = smem + (y & 1) * q.BUFSZ;
            int16_t dout = FILT;
            int costv = 0;
            bool okc = false;
            if (x >= 0 && tsum >= a.texThr && !(a.dbg & 1)) {
                const uint8_t *pa = buf + oa, *pb = buf + ob, *pc = buf + oc_;
                auto sad4 = [&](int o, uint32_t (&sv)[4]) {
                    const uint4 t = *reinterpret_cast<const uint4 *>(pa + 16 * o);
                    const uint4 w = *reinterpret_cast<const uint4 *>(pb + 16 * o);
                    const uint4 u = *reinterpret_cast<const uint4 *>(pc + 16 * o);
                    sv[0] = t.x + w.x - u.x; sv[1] = t.y + w.y - u.y; sv[2] = t.z + w.z - u.z; sv[3] = t.w + w.w - u.w;
                };
                // exact octet in forward order (position p = disparity index 8 * o + p)
                auto sad4f = [&](int o, uint32_t (&sv)[4]) {
                    uint32_t r[4];
                    sad4(o, r);
                    sv[0] = __byte_perm(r[0], r[3], usel);
                    sv[1] = __byte_perm(r[1], r[2], usel);
                    sv[2] = __byte_perm(r[2], r[1], usel);
                    sv[3] = __byte_perm(r[3], r[0], usel);
                };
                // pass 1: (octet minimum << 16 | octet) keys made by the producers -> argmin octet (smallest octet on ties)
                uint4 *s4 = reinterpret_cast<uint4 *>(const_cast<uint8_t *>(buf) + omn);
                uint32_t *s32 = reinterpret_cast<uint32_t *>(s4);
                uint32_t best = 0xFFFFFFFFu;
#pragma unroll
                for (int k = 0; k < (NO_ + 3) / 4; k++) {
                    const uint4 v = s4[k];
                    best = __vimin3_u32(best, v.x, v.y);
                    best = __vimin3_u32(best, v.z, v.w);
                }
                const int BIASC = ((y - y0) & 1) ? 0 : (2 * H_ + 1) * 128;     // see the producers' row()
                const int minsad = (int)(best >> 16) - BIASC, oc = (int)(best & 0xFFFFu);
                // exact position inside the argmin octet (first minimum) via (value << 3 | index) keys
                int mind;
                {
                    uint32_t sv[4];
                    sad4f(oc, sv);
                    const uint32_t s0 = sv[0], s1 = sv[1], s2 = sv[2], s3 = sv[3];
                    uint32_t k = __vimin3_u32(__umul24(s0 & 0xFFFFu, 8u), __umul24(s0 >> 16, 8u) + 1u, __umul24(s1 & 0xFFFFu, 8u) + 2u);
                    k = __vimin3_u32(k, __umul24(s1 >> 16, 8u) + 3u, __umul24(s2 & 0xFFFFu, 8u) + 4u);
                    k = __vimin3_u32(k, __umul24(s2 >> 16, 8u) + 5u, __umul24(s3 & 0xFFFFu, 8u) + 6u);
                    k = min(k, __umul24(s3 >> 16, 8u) + 7u);
                    mind = 8 * oc + (int)(k & 7u);
                }
                const int dp = mind + 1 < ND ? mind + 1 : ND - 2, dn = mind > 0 ? mind - 1 : 1;
                const uint16_t *a16 = reinterpret_cast<const uint16_t *>(pa);
                const uint16_t *b16 = reinterpret_cast<const uint16_t *>(pb);
                const uint16_t *c16 = reinterpret_cast<const uint16_t *>(pc);
                const int ip = dp ^ xflip, in = dn ^ xflip;
                const int p = (int)a16[ip] + (int)b16[ip] - (int)c16[ip] - BIASC;
                const int n = (int)a16[in] + (int)b16[in] - (int)c16[in] - BIASC;
                bool ok = true;
                if (a.uniq > 0) {
                    const int thresh = minsad + (minsad * a.uniq / 100) + BIASC;     // compared with biased sums
                    const int zlo = max(mind - 1, 0), zhi = min(mind + 1, ND - 1);
                    const int olo = zlo >> 3, ohi = zhi >> 3;
                    // octets that do not touch [mind-1, mind+1]: their minimum decides
                    s32[olo] = 0xFFFFFFFFu;
                    s32[ohi] = 0xFFFFFFFFu;
                    uint32_t m2 = 0xFFFFFFFFu;
#pragma unroll
                    for (int k = 0; k < (NO_ + 3) / 4; k++) {
                        const uint4 v = s4[k];
                        m2 = __vimin3_u32(m2, v.x, v.y);
                        m2 = __vimin3_u32(m2, v.z, v.w);
                    }
                    ok = (int)(m2 >> 16) > thresh;
                    // the (at most two) touching octets: exact check with the neighbourhood masked out
                    for (int oo = olo; ok && oo <= ohi; oo++) {
                        uint32_t sv[4];
                        sad4f(oo, sv);
                        const uint4 Z = c_zmask[mind - 8 * oo + 1];
                        const uint32_t mz = __vminu2(__vminu2(sv[0] | Z.x, sv[1] | Z.y), __vminu2(sv[2] | Z.z, sv[3] | Z.w));
                        ok = (int)min(mz & 0xFFFFu, mz >> 16) > thresh;
                    }
                }
                if (ok) {
                    // v = (nd - mind - 1) * 256 + (q ? (p - n) * 256 / q : 0) + 15, C division (truncating);
                    // q = p + n - 2 * minsad + |p - n| >= 2 |p - n| -> |quotient| <= 128, exact through one fp32 reciprocal
                    const int dpn = p - n, adpn = abs(dpn);
                    const int qd = p + n - 2 * minsad + adpn;
                    int quo = 0;
                    if (qd != 0) {
                        const int num = adpn * 256;                                 // < 2^24
                        int t = (int)__fdividef((float)num, (float)qd);
                        const int rem = num - t * qd;
                        t += rem >= qd ? 1 : 0;
                        t -= rem < 0 ? 1 : 0;
                        quo = dpn < 0 ? -t : t;
                    }
                    const int v = (ND - mind - 1) * 256 + quo + 15;
                    dout = (int16_t)(v >> 4);
                    costv = minsad;
                    okc = true;
                }
            }
            if (y + 2 < y1) bar_arrive(3 + (y & 1), NT);        // the buffer of row y may be overwritten (row y + 2)
            if (x >= 0) {
                *dptr = dout;
                if (cptr && okc) *cptr = (int16_t)costv;
            }
            dptr += a.disp.pitch;
            if (cptr) cptr += a.cost.pitch;
            tsum = tnext;
        }
    }
}

struct Tiling3 { int NG, TW, BH, nstripes, nbands, NT, pair; size_t smem; long long cost; };

// SMs of the current device (cached per device)
int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return cached[dev];
}

bool pick_tiling3(const BmGeom &g, int n, Tiling3 *t)
{
    const int h = g.bs / 2;
    if (g.minD != 0 || h < 2 || h > 7) return false;
    if (!(g.nd == 256 || g.nd == 192 || g.nd == 128 || g.nd == 96 || g.nd == 64 || g.nd == 48 || g.nd == 32)) return false;
    const int NO = g.nd / 8, G = 2 * h;
    // RTDM_BM3_SHAPE = 0: one wide CTA per SM, 1: two narrower CTAs per SM
    const int pair = (g.sw.bm3_shape && (g.nd == 64 || g.nd == 128)) ? 1 : 0;
    t->pair = pair;
    const int NCW = pair ? ShapePair::NCW : ShapeWide::NCW, NLD = pair ? ShapePair::NLD : ShapeWide::NLD;
    const int MAXT = pair ? ShapePair::MAXT : ShapeWide::MAXT;
    const size_t smem_max = pair ? 110 * 1024 : 220 * 1024;
    const int item_max = 32 * (pair ? ShapePair::MAXW : ShapeWide::MAXW);
    // producer warps = (A, B) x (even, odd groups) x wpp, with (32 / NO) groups per warp
    const int spw = 32 / NO, wpp_max = (MAXT - NCW * 32) / 128;
    int ngmax = 2 * wpp_max * spw;
    while (ngmax > 2 && ((size_t)make_geo3(h, g.nd, ngmax).total > smem_max || make_geo3(h, g.nd, ngmax).NITEM > item_max * NLD)) ngmax--;   // 320: loader warp, 10 words per lane
    int twmax = std::min(ngmax * G - 2 * h, (NCW - NLD) * 32);
    if (twmax < 16) return false;
    t->nstripes = cdiv(g.W1, twmax);
    t->TW = cdiv(g.W1, t->nstripes);
    t->NG = cdiv(t->TW + 2 * h, G);
    if (t->NG > ngmax) return false;
    t->NT = 128 * (((t->NG + 1) / 2 + spw - 1) / spw) + NCW * 32;
    // bands: a band pays ~START rows of start-up (ring prologue, 2h+1 rows of window sums) and the launch runs in waves
    // of (SMs x CTAs per SM) CTAs -> take the band count with the smallest (waves + 1/2) x (band height + START); the
    // half wave stands for the tail (stripes differ a little), and bands stay <= 128 rows: measured, a few long CTAs
    // per SM (706-row bands, 3 waves) lose more to that tail than they save in start-up
    const int rows = g.row1 - g.row0;
    const int slots = sm_count() * (pair ? 2 : 1), START = 2 * h + 8;
    long long best = -1;
    for (int nb = 1; nb <= std::max(1, rows / 8) && nb <= 64; nb++) {
        const int bh = cdiv(rows, nb);
        if (cdiv(rows, bh) != nb || (bh > 128 && nb < std::max(1, rows / 8))) continue;
        const long long waves = ((long long)n * t->nstripes * nb + slots - 1) / slots;
        const long long cost = (2 * waves + 1) * (bh + START);
        if (best < 0 || cost < best) { best = cost; t->nbands = nb; t->BH = bh; }
    }
    t->cost = best;
    t->smem = (size_t)make_geo3(h, g.nd, t->NG).total;
    return t->smem <= smem_max;
}

bool g_zmask_ready[64] = {false};

int upload_zmask()
{
    int dev = 0;
    RTDM_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && g_zmask_ready[dev]) return 0;
    uint32_t zm[10][4];
    for (int e = 0; e < 10; e++) {
        const int rel = e - 1;
        for (int r = 0; r < 4; r++) {
            uint32_t m = 0;
            for (int s = 0; s < 2; s++) {
                const int pos = 2 * r + s;
                if (pos >= rel - 1 && pos <= rel + 1) m |= 0xFFFFu << (16 * s);
            }
            zm[e][r] = m;
        }
    }
    RTDM_CUDA(cudaMemcpyToSymbol(c_zmask, zm, sizeof(zm)));
    if (dev >= 0 && dev < 64) g_zmask_ready[dev] = true;
    return 0;
}

template <int H_, int NO_, class SH>
int launch3s(const Bm3Args &a, const Tiling3 &t, int n, cudaStream_t st)
{
    RTDM_CUDA(cudaFuncSetAttribute(bm_sad3_kernel<H_, NO_, SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem));
    bm_sad3_kernel<H_, NO_, SH><<<dim3(t.nstripes, t.nbands, n), t.NT, t.smem, st>>>(a);
    return 0;
}
template <int H_, int NO_>
int launch3(const Bm3Args &a, const Tiling3 &t, int n, cudaStream_t st)
{
    if constexpr (NO_ == 8 || NO_ == 16) { if (t.pair) return launch3s<H_, NO_, ShapePair>(a, t, n, st); }
    return launch3s<H_, NO_, ShapeWide>(a, t, n, st);
}

}  // namespace

bool bm_sad3_supported(const BmGeom &g, int n)
{
    Tiling3 t;
    return g.W1 >= 1 && g.row1 > g.row0 && pick_tiling3(g, n, &t);
}

// estimated cost of one launch of n frames in row steps (waves x (band height + start-up)), -1 when the kernel does not
// apply: the host layer sizes its chunks with it
long long bm_sad3_cost(const BmGeom &g, int n)
{
    Tiling3 t;
    if (!(g.W1 >= 1 && g.row1 > g.row0 && pick_tiling3(g, n, &t))) return -1;
    return t.cost;
}

// SAD + WTA kernel only; the texture sums `tex` must have been produced already (bm_sad2.cu: bm_texture_kernel)
int launch_bm_sad3_core(const BmGeom &g, int n, PlaneU8 Lp, PlaneU8 Rp, PlaneS16 disp, PlaneS16 cost,
                        const uint16_t *tex, size_t tex_pitch, size_t tex_frame, cudaStream_t st)
{
    Tiling3 t;
    if (!pick_tiling3(g, n, &t)) { set_error("bm_sad3: unsupported geometry"); return -RTDM_EINVAL; }
    int rc = upload_zmask();
    if (rc) return rc;
    Bm3Args a;
    a.Lp = Lp; a.Rp = Rp; a.disp = disp; a.cost = cost;
    a.tex = tex; a.tex_pitch = tex_pitch; a.tex_frame = tex_frame;
    a.W = g.W; a.H = g.H; a.nd = g.nd; a.cap = g.cap; a.texThr = g.texThr; a.uniq = g.uniq;
    a.W1 = g.W1; a.row0 = g.row0; a.row1 = g.row1;
    a.TW = t.TW; a.BH = t.BH; a.NG = t.NG;
    a.dbg = g.sw.bm_debug;
    const int h = g.bs / 2;
#define RTDM_SAD3_ND(H_) (g.nd == 128 ? launch3<H_, 16>(a, t, n, st) : g.nd == 64 ? launch3<H_, 8>(a, t, n, st) : \
                          g.nd == 192 ? launch3<H_, 24>(a, t, n, st) : g.nd == 96 ? launch3<H_, 12>(a, t, n, st) : \
                          g.nd == 48 ? launch3<H_, 6>(a, t, n, st) : g.nd == 256 ? launch3<H_, 32>(a, t, n, st) : launch3<H_, 4>(a, t, n, st))
    switch (h) {
        case 2: rc = RTDM_SAD3_ND(2); break;
        case 3: rc = RTDM_SAD3_ND(3); break;
        case 4: rc = RTDM_SAD3_ND(4); break;
        case 5: rc = RTDM_SAD3_ND(5); break;
        case 6: rc = RTDM_SAD3_ND(6); break;
        default: rc = RTDM_SAD3_ND(7); break;
    }
#undef RTDM_SAD3_ND
    if (rc) return rc;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rtdm
