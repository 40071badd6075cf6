// bm_sad4.cu -- TMA-staged, warp-specialised Konolige block-matching core (minDisparity == 0, blockSize 5 .. 15,
// numDisparities 32 / 48 / 64 / 96 / 128 / 192 / 256).  Same arithmetic as bm_sad.cu / bm_sad2.cu / bm_sad3.cu
// (oracle: orc_bm_core); replaces findStereoCorrespondenceBM as reached from
// SWMatcherKonolige::compute (reference stereo-matcher/bm-sw.cpp:33-38).
//
// What changed against bm_sad3.cu (profiles/r01_prof_bm3_h_summary.csv: issue slots 62 % busy, 20 % of the shared-memory
// wavefronts bank conflicts, two loader warps spending 212 instructions per row on funnel shifts and byte permutes):
//
//   * The rectified row bands are staged by the TMA unit.  The prefilter (prefilter.cu) writes the two layouts the
//     producers read -- the left image EXPANDED (one 32-bit word = pixel x 0x01010101, 8 replicated columns past the
//     right border) and the right image with 16 replicated columns on both sides -- so a ring row is two bulk
//     asynchronous copies (cp.async.bulk.shared.global, SASS UBLKCP: one left, one right) issued by ONE elected lane per
//     row and signalled on an mbarrier per ring slot (SYNCS); the border clamps of App. A.2 are the replicated columns,
//     what a copy reads beyond them only feeds pixels that are never written.  No loader warps, no descriptors, no byte
//     gathers in the band start-up.  Bulk copies need 16-byte aligned sources: stripe s starts at pixel s * TW - (HP - h)
//     with TW a multiple of 4 and HP = h rounded up to 4, so that its first virtual column x0 - h is a multiple of 4 (the
//     HP - h pixels left of the image in stripe 0 are computed and dropped), the expanded plane stores pixel x at element
//     x + 1 (numDisparities - 1 = 3 mod 4), and the right row is copied from the 16-byte boundary below its first byte.  (The tensor-map form, cp.async.bulk.tensor / UTMALDG, raises "illegal instruction"
//     on this pool's B200s in every variant tried -- tools/probe/tma_probe.cu -- while the bulk form runs.)
//   * No mirrored copies.  The second half ("B") of a column group emits its in-half SUFFIX sums from the same forward
//     rows with its own compile-time shifts (warps are type-uniform), so every sum in shared memory is in forward
//     disparity order: the byte reversals of the half totals (8 PRMT per producer thread and row) and the un-reversal in
//     the winner-take-all warps are gone, and a ring row is 5 instead of 10 bytes per column.
//   * The texture window sums (textureThreshold test) are made in the kernel, by the otherwise idle warp that feeds the
//     ring, from the expanded left rows already in shared memory: the separate texture kernel and the plain prefiltered
//     planes are gone.
//   * Even half-widths keep the groups of a warp ADJACENT (bm_sad3 interleaved them by parity for the odd half-widths),
//     which takes the two-way bank conflicts off the right-row loads and the octet-key stores.
//
//   one CTA per SM = stripe of TW <= 180 computed columns x band of BH rows of one frame,
//   producer warps (thread = (half group of h adjacent columns, disparity octet)) + 6 winner-take-all warps (thread = pixel).
//   A (2h+1)-column window always spans two groups, so with i = x mod 2h
//           i <  h :  SAD(x) = T[g]   - PreA[x - 1]  + PreA'[i]
//           i >= h :  SAD(x) = SufB[x] + T[g + 1]     - SufB'[i + 1]
//   Producers work on row y + 1 while the winner-take-all warps work on row y (sums double buffered, named barriers).
// The cost volume never leaves the SM; HBM traffic is the staged rows in and disparity + cost out.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace rtdm {
namespace {

constexpr int NCW4 = 6;                    // winner-take-all warps (<= 192 pixels per stripe)
constexpr int NFW4 = 1;                    // feeder warp: one lane issues the bulk copies of the ring
constexpr int MAXT4 = 736;                 // 16 producer warps + NCW4 + NFW4
constexpr int RING_EXTRA = 2;              // ring rows in flight beyond the 2h + 3 live ones

__host__ __device__ constexpr int ring_rows4(int h) { return 2 * h + 3 + RING_EXTRA; }

__device__ __forceinline__ int clampi4(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ void bar_sync4(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive4(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// ---- mbarrier / TMA (PTX ISA 8.x; sm_100a) --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// bulk asynchronous copy global -> shared (TMA unit, SASS UBLKCP): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}
__host__ __device__ constexpr int hp_of(int h) { return (h + 3) & ~3; }     // stripe s starts at pixel s * TW - (hp_of(h) - h)

struct Bm4Args {
    PlaneS16 disp, cost;
    int W, H, nd, cap, texThr, uniq;
    int W1, row0, row1;
    int TW, BH, NG;              // stripe width (a multiple of 4), band height, column groups
    BmStaged sp;                 // the staged planes (prefilter.cu)
};

// masks of the uniqueness test: entry rel + 1 (rel = mind - 8 * octet in -1 .. 8) has 0xFFFF in the 16-bit lanes
// of the positions rel - 1 .. rel + 1 that fall inside the octet
__constant__ uint4 c_zmask4[10];

// bytes per pixel of the octet-key rows: nd / 8 keys of 4 bytes, rounded up to whole 128-bit words (the spare keys stay
// 0xFFFFFFFF), padded to an ODD number of 16-byte units so that 128-bit rows of 8 neighbouring pixels hit 8 bank groups
__host__ __device__ constexpr int mnp_bytes4(int nd) { return (((nd / 2 + 15) / 16) | 1) * 16; }

// shared-memory geometry shared by host and device
struct Geo4 {
    // one sums buffer: X[NCT] | HA[NG] | HB[NG] | TA[NG] | TB[NG] | zero row (pitch PP each) | Mn[NCT] (pitch MNP)
    int NCT, NCTP, PP, BUFSZ, XOFF, HAOFF, HBOFF, TAOFF, TBOFF, ZOFF, MNOFF;
    int LFB, RBY, SLOT;                              // ring slot: expanded left row (LFB bytes), then RBY bytes of the right row
    int RINGOFF, BAROFF, TEXOFF, total;              // TEX: 2 x 256 u16 texture sums (row parity)
};
__host__ __device__ inline Geo4 make_geo4(int h, int nd, int NG)
{
    Geo4 q;
    const int G = 2 * h;
    q.NCT = NG * G;
    q.NCTP = (q.NCT + 3) & ~3;
    q.PP = nd * 2 + 16;
    q.XOFF = 0;
    q.HAOFF = q.NCT * q.PP;
    q.HBOFF = q.HAOFF + NG * q.PP;
    q.TAOFF = q.HBOFF + NG * q.PP;
    q.TBOFF = q.TAOFF + NG * q.PP;
    q.ZOFF = q.TBOFF + NG * q.PP;
    q.MNOFF = q.ZOFF + q.PP;
    q.BUFSZ = (q.MNOFF + q.NCT * mnp_bytes4(nd) + 127) & ~127;
    q.LFB = (4 * q.NCTP + 127) & ~127;
    // right-row bytes a producer may touch (stream + word rounding + pre-shift) + the <= 12 bytes between the 16-byte
    // boundary the copy starts at and the stripe's first byte
    q.RBY = (q.NCTP + nd + 16 + 12 + 15) & ~15;
    q.SLOT = q.LFB + ((q.RBY + 127) & ~127);
    q.RINGOFF = 2 * q.BUFSZ;
    q.BAROFF = q.RINGOFF + ring_rows4(h) * q.SLOT;
    q.TEXOFF = (q.BAROFF + 8 * ring_rows4(h) + 15) & ~15;
    q.total = q.TEXOFF + 2 * 256 * 2 + 128;          // + slack to align the carve-up to 128 bytes
    return q;
}

__host__ __device__ constexpr int producer_threads4(int h, int no, int NG)
{
    const int spw = 32 / no;
    return (h % 2 == 0) ? 64 * ((NG + spw - 1) / spw)                       // (A, B) warp pairs over adjacent groups
                        : 128 * (((NG + 1) / 2 + spw - 1) / spw);           // (A, B) x (even, odd groups)
}

// h words of an expanded left row starting at byte offset `bo` whose alignment (16 / 8 / 4) is known at compile time
template <int H_, int ALIGN>
__device__ __forceinline__ void load_left4(const uint8_t *slot, int bo, uint32_t (&lw)[H_])
{
    if constexpr (ALIGN == 4) {
        // 4-byte aligned stream: start one word early (8-byte aligned), h + 1 words (h is odd here)
        static_assert(H_ % 2 == 1, "ALIGN 4 only occurs for odd half-widths");
        const uint8_t *p = slot + bo - 4;
        uint32_t t[H_ + 1];
#pragma unroll
        for (int i = 0; i < H_ + 1; i += 2) {
            const uint2 v = *reinterpret_cast<const uint2 *>(p + 4 * i);
            t[i] = v.x; t[i + 1] = v.y;
        }
#pragma unroll
        for (int i = 0; i < H_; i++) lw[i] = t[i + 1];
    } else {
        const uint8_t *p = slot + bo;
        constexpr int W4 = (ALIGN == 16) ? (H_ & ~3) : 0;
#pragma unroll
        for (int i = 0; i < W4; i += 4) {
            const uint4 v = *reinterpret_cast<const uint4 *>(p + 4 * i);
            lw[i] = v.x; lw[i + 1] = v.y; lw[i + 2] = v.z; lw[i + 3] = v.w;
        }
#pragma unroll
        for (int i = W4; i + 1 < H_; i += 2) {
            const uint2 v = *reinterpret_cast<const uint2 *>(p + 4 * i);
            lw[i] = v.x; lw[i + 1] = v.y;
        }
        if ((H_ - W4) & 1) lw[H_ - 1] = *reinterpret_cast<const uint32_t *>(p + 4 * (H_ - 1));
    }
}

// the producer side of the kernel for one (type, right-stream offset, left-row alignment) variant.
// OFFC >= 0: the right stream's byte offset inside its first word, known at compile time; -1: warp-uniform run-time value
// (pre-shifted to 0 when the words are loaded)
template <int H_, int NO_, bool ISB, int OFFC, int LAL>
__device__ __forceinline__ void producer4(uint8_t *smem, const Geo4 &q, const Bm4Args &a, int x0, int y0, int y1, int NPT, int NT, uint32_t bar_s)
{
    constexpr int G = 2 * H_, RING = ring_rows4(H_);
    constexpr bool EVENH = (H_ % 2 == 0);
    constexpr int ND = NO_ * 8, PP = ND * 2 + 16, MNP = mnp_bytes4(ND);
    constexpr int SPW = 32 / NO_;                             // groups per producer warp
    constexpr int isB = ISB ? 1 : 0;
    const int tid = threadIdx.x;
    uint8_t *Ring = smem + q.RINGOFF;
        // even producer warps hold A halves, odd warps B halves (type-uniform warps).  Even h: a warp holds SPW adjacent
        // groups.  Odd h: 2h is only 2 (mod 4), so within a type warps alternate between even and odd groups, which keeps
        // the byte alignment of a thread's right-row stream warp-uniform.
        const int pw = tid >> 5, widx = pw >> 1;
        const int sub = (tid & 31) / NO_, j = (tid & 31) - sub * NO_;
        const int g = EVENH ? widx * SPW + sub : 2 * ((widx >> 1) * SPW + sub) + (widx & 1);
        const bool live = g < a.NG && sub < SPW;                  // trailing groups may not exist; 32 % NO_ lanes of a warp stay idle
        const int cb = g * G + (isB ? H_ : 0);                    // first virtual column of this thread (forward order, both types)
        const int lbo = 4 * cb;                                   // expanded left row: one word per column
        const int rsh = (x0 - H_ + BmStaged::RPADL) & 12;          // the right row was copied from the 16-byte boundary below its first byte
        const int rbo = q.LFB + rsh + cb + 8 * j;                 // right row: byte stream
        const int roff = cb & 3;                                  // even h: 0 (A) / h & 3 (B); odd h: warp-uniform
        // R clamp (App. A.2, minD = 0): rbase(xc) = clip(xc, 0, W - nd); in virtual columns c = xc - x0 + h
        const int cmin = H_ - x0, cmax = (a.W - ND) - x0 + H_;
        uint32_t clmask = 0;
        int crc = 0;
#pragma unroll
        for (int k = 0; k < H_; k++) {
            const int c = cb + k;
            if (c < cmin) { clmask |= 1u << k; crc = cmin; }
            if (c > cmax) { clmask |= 1u << k; crc = cmax; }
        }
        const bool wborder = __any_sync(0xFFFFFFFFu, live && clmask != 0);
        // clamped columns read the R window of the nearest unclamped column
        const int cbo = q.LFB + rsh + crc + 8 * j;

        uint32_t V[H_][4];
#pragma unroll
        for (int k = 0; k < H_; k++) V[k][0] = V[k][1] = V[k][2] = V[k][3] = 0u;

        // OFFC >= 0: the stream's byte offset inside its first word, known at compile time; -1: warp-uniform run-time value
        // (pre-shifted to 0 when the words are loaded)
        constexpr int NRWX = (3 + H_ + 6) / 4 + 2;                // words held for any variant

        auto load_right = [&](const uint8_t *slot, uint32_t (&rw)[NRWX]) {
            const uint32_t *rp = reinterpret_cast<const uint32_t *>(slot + (rbo & ~3));
            constexpr int NRW = ((OFFC < 0 ? 0 : OFFC) + H_ + 6) / 4 + 1;
            if constexpr (OFFC >= 0) {
#pragma unroll
                for (int i = 0; i < NRW; i++) rw[i] = rp[i];
            } else {
                uint32_t t[NRW + 1];
#pragma unroll
                for (int i = 0; i < NRW + 1; i++) t[i] = rp[i];
                const int sh = 8 * roff;
#pragma unroll
                for (int i = 0; i < NRW; i++) rw[i] = __funnelshift_r(t[i], t[i + 1], sh);     // sh == 0: t[i]
            }
        };
        auto clamped_window = [&](const uint8_t *slot, uint32_t &c0w, uint32_t &c1w) {
            const uint32_t *pwd = reinterpret_cast<const uint32_t *>(slot + (cbo & ~3));
            const int sh = (cbo & 3) * 8;
            c0w = __funnelshift_r(pwd[0], pwd[1], sh);
            c1w = __funnelshift_r(pwd[1], pwd[2], sh);
        };
        // |L - R| of column k (8 disparities in stream order)
        auto ad_col = [&](const uint32_t (&lw)[H_], const uint32_t (&rw)[NRWX], int k, int offc, bool border, uint32_t c0w, uint32_t c1w,
                          uint32_t &lo, uint32_t &hi) {
            const int pos = (offc < 0 ? 0 : offc) + k;
            const int w = pos >> 2, sft = pos & 3;
            const uint32_t l4 = lw[k];
            uint32_t r0, r1;
            if (sft == 0) { r0 = rw[w]; r1 = rw[w + 1]; }
            else { r0 = __funnelshift_r(rw[w], rw[w + 1], 8 * sft); r1 = __funnelshift_r(rw[w + 1], rw[w + 2], 8 * sft); }
            if (border && ((clmask >> k) & 1u)) { r0 = c0w; r1 = c1w; }
            lo = __vabsdiffu4(l4, r0);
            hi = __vabsdiffu4(l4, r1);
        };

        const int xst = q.XOFF + ((2 * g + isB) * H_) * PP + 16 * j;           // store base of the thread's h prefix / suffix slots
        const int hst = (isB ? q.HBOFF : q.HAOFF) + g * PP + 16 * j;            // own half total
        // phase 2: A half of group g makes the pixels gG + k, B half the pixels gG + G - 1 - k; both need group g + 1
        const bool ph2 = live && g + 1 < a.NG;
        const int nxt = q.XOFF + ((g + 1) * G + (isB ? H_ : 0)) * PP + 16 * j;  // next group's half of the same type
        const int mst = q.MNOFF + (isB ? g * G + G - 1 : g * G) * MNP + 4 * j;
        const int mstep = isB ? -MNP : MNP;
        auto emit_min = [&](uint8_t *buf, int k, uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3) {
            uint32_t m = __vminu2(__vimin3_u16x2(s0, s1, s2), s3);
            m = __vminu2(m, m >> 16);
            *reinterpret_cast<uint32_t *>(buf + mst + k * mstep) = m * 65536u + (uint32_t)j;      // (octet minimum << 16) | octet
        };


            // vertical sums over ring rows 0 .. 2h, i.e. image rows y0-h-1 .. y0+h-1 (the first row removes row y0-h-1 again)
            if (live) {
#pragma unroll 1
                for (int t = 0; t <= 2 * H_; t++) {
                    mbar_wait(bar_s + 8 * t, 0);
                    const uint8_t *slot = Ring + t * q.SLOT;
                    uint32_t lw[H_], rw[NRWX], c0w = 0, c1w = 0;
                    load_left4<H_, LAL>(slot, lbo, lw);
                    load_right(slot, rw);
                    if (wborder) clamped_window(slot, c0w, c1w);
#pragma unroll
                    for (int k = 0; k < H_; k++) {
                        uint32_t lo, hi;
                        ad_col(lw, rw, k, OFFC, wborder, c0w, c1w, lo, hi);
                        V[k][0] += __byte_perm(lo, 0, 0x4140);
                        V[k][1] += __byte_perm(lo, 0, 0x4342);
                        V[k][2] += __byte_perm(hi, 0, 0x4140);
                        V[k][3] += __byte_perm(hi, 0, 0x4342);
                    }
                }
            }

            // Even rows of the band add the byte deltas (in + 128 - out), odd rows subtract the mirrored deltas
            // (out + 128 - in): V += d + 128, then V += d - 128.  No per-lane bias correction is needed: after an even
            // row every column sum carries +128, so every (2h+1)-column window sum carries the same (2h+1) * 128, which
            // the winner-take-all warps subtract from the minimum and the two neighbours (BIASC).
            auto row = [&](int y, int sin_i, int sout_i, uint32_t pin, auto border_tag, auto odd_tag) {
                constexpr bool BORDER = decltype(border_tag)::value;
                constexpr bool ODD = decltype(odd_tag)::value;
                uint8_t *buf = smem + (y & 1) * q.BUFSZ;
                uint4 p = make_uint4(0, 0, 0, 0);
                if (live) {
                    mbar_wait(bar_s + 8 * sin_i, pin);               // row y + h has landed
                    const uint8_t *sin = Ring + sin_i * q.SLOT;
                    const uint8_t *sout = Ring + sout_i * q.SLOT;
                    uint32_t lwi[H_], rwi[NRWX], lwo[H_], rwo[NRWX];
                    uint32_t ci0 = 0, ci1 = 0, co0 = 0, co1 = 0;
                    load_left4<H_, LAL>(sin, lbo, lwi);
                    load_right(sin, rwi);
                    load_left4<H_, LAL>(sout, lbo, lwo);
                    load_right(sout, rwo);
                    if (BORDER) { clamped_window(sin, ci0, ci1); clamped_window(sout, co0, co1); }
                    uint8_t *pdst = buf + xst;
#pragma unroll
                    for (int m = 0; m < H_; m++) {
                        const int k = ISB ? H_ - 1 - m : m;          // A: prefix sums left to right, B: suffix sums right to left
                        uint32_t lo, hi, olo, ohi;
                        ad_col(lwi, rwi, k, OFFC, BORDER, ci0, ci1, lo, hi);
                        ad_col(lwo, rwo, k, OFFC, BORDER, co0, co1, olo, ohi);
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
                        *reinterpret_cast<uint4 *>(pdst + m * PP) = p;
                    }
                    *reinterpret_cast<uint4 *>(buf + hst) = p;       // half total
                }
                bar_sync4(5, NPT);                                   // all prefix / suffix sums of row y are in shared memory
                if (ph2) {
                    if (!ISB) {
                        // A pixels: SAD(gG + k) = T[g] - PreA[k - 1] + PreA'[k]
                        const uint4 hb = *reinterpret_cast<const uint4 *>(buf + q.HBOFF + g * PP + 16 * j);
                        uint4 run = make_uint4(p.x + hb.x, p.y + hb.y, p.z + hb.z, p.w + hb.w);
                        *reinterpret_cast<uint4 *>(buf + q.TAOFF + g * PP + 16 * j) = run;
#pragma unroll
                        for (int k = 0; k < H_; k++) {
                            const uint4 l = *reinterpret_cast<const uint4 *>(buf + nxt + k * PP);
                            emit_min(buf, k, run.x + l.x, run.y + l.y, run.z + l.z, run.w + l.w);
                            run.x -= V[k][0]; run.y -= V[k][1]; run.z -= V[k][2]; run.w -= V[k][3];
                        }
                    } else {
                        // B pixels: SAD(gG + G - 1 - m) = SufB[m] + T[g + 1] - SufB'[m - 1]
                        const uint4 ha = *reinterpret_cast<const uint4 *>(buf + q.HAOFF + (g + 1) * PP + 16 * j);
                        const uint4 hb = *reinterpret_cast<const uint4 *>(buf + q.HBOFF + (g + 1) * PP + 16 * j);
                        uint4 run = make_uint4(ha.x + hb.x, ha.y + hb.y, ha.z + hb.z, ha.w + hb.w);
                        *reinterpret_cast<uint4 *>(buf + q.TBOFF + (g + 1) * PP + 16 * j) = run;
#pragma unroll
                        for (int m = 0; m < H_; m++) {
                            const int k = H_ - 1 - m;
                            run.x += V[k][0]; run.y += V[k][1]; run.z += V[k][2]; run.w += V[k][3];
                            uint4 l = make_uint4(0, 0, 0, 0);
                            if (m > 0) l = *reinterpret_cast<const uint4 *>(buf + nxt + (m - 1) * PP);
                            emit_min(buf, m, run.x - l.x, run.y - l.y, run.z - l.z, run.w - l.w);
                        }
                    }
                }
            };

            int sin_i = 2 * H_ + 1, sout_i = 0;
            uint32_t pin = 0;
#pragma unroll 1
            for (int y = y0; y < y1; y++) {
                if (y - y0 >= 2) bar_sync4(3 + (y & 1), NT);           // the winner-take-all warps are done with this buffer (row y - 2)
                if ((y - y0) & 1) {
                    if (wborder) row(y, sin_i, sout_i, pin, std::true_type(), std::true_type());
                    else row(y, sin_i, sout_i, pin, std::false_type(), std::true_type());
                } else {
                    if (wborder) row(y, sin_i, sout_i, pin, std::true_type(), std::false_type());
                    else row(y, sin_i, sout_i, pin, std::false_type(), std::false_type());
                }
                bar_arrive4(1 + (y & 1), NT);                          // sums and octet minima of row y are complete
                if (++sin_i == RING) { sin_i = 0; pin ^= 1u; }
                if (++sout_i == RING) sout_i = 0;
            }
}

template <int H_, int NO_>
__global__ void __launch_bounds__(MAXT4, 1)
bm_sad4_kernel(Bm4Args a)
{
    constexpr int G = 2 * H_, RING = ring_rows4(H_), HP = hp_of(H_);
    constexpr bool EVENH = (H_ % 2 == 0);
    constexpr int ND = NO_ * 8, PP = ND * 2 + 16, MNP = mnp_bytes4(ND);
    constexpr int SPW = 32 / NO_;                             // groups per producer warp
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int tid = threadIdx.x, f = blockIdx.z;
    const int x0 = blockIdx.x * a.TW - (HP - H_);             // first pixel of the stripe (negative in stripe 0 when h % 4 != 0)
    const int TWc = min(a.TW, a.W1 - x0);
    const int y0 = a.row0 + blockIdx.y * a.BH, y1 = min(y0 + a.BH, a.row1);
    if (TWc <= 0 || y0 >= y1) return;
    uint8_t *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const Geo4 q = make_geo4(H_, ND, a.NG);
    const int NPT = producer_threads4(H_, NO_, a.NG);
    const int NT = NPT + (NCW4 + NFW4) * 32;
    uint8_t *Ring = smem + q.RINGOFF;
    const uint32_t ring_s = smem_u32(Ring), bar_s = smem_u32(smem + q.BAROFF);
    const int lofs = ND - 1;
    const int nrows = y1 - y0;
    const int t_last = nrows - 1 + 2 * H_ + 1;                // last ring row (band-relative) any producer reads

    // ---- prologue: mbarriers, zero rows, spare keys --------------------------------------------------------------
    if (tid == 0) {
#pragma unroll 1
        for (int s = 0; s < RING; s++) mbar_init(bar_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int i = tid; i < PP / 4; i += NT) {
        reinterpret_cast<uint32_t *>(smem + q.ZOFF)[i] = 0u;
        reinterpret_cast<uint32_t *>(smem + q.BUFSZ + q.ZOFF)[i] = 0u;
    }
    if (NO_ % 4)                                                // spare keys of the last 128-bit word of every key row
        for (int i = tid; i < q.NCT * (MNP / 4); i += NT) {
            reinterpret_cast<uint32_t *>(smem + q.MNOFF)[i] = 0xFFFFFFFFu;
            reinterpret_cast<uint32_t *>(smem + q.BUFSZ + q.MNOFF)[i] = 0xFFFFFFFFu;
        }
    __syncthreads();

    if (tid < NPT) {
        // =========================================================================================
        // producer: even warps hold A halves, odd warps B halves (type-uniform warps)
        // =========================================================================================
        const bool isB = (tid >> 5) & 1;
        if constexpr (EVENH) {
            if (!isB) producer4<H_, NO_, false, 0, 16>(smem, q, a, x0, y0, y1, NPT, NT, bar_s);
            else producer4<H_, NO_, true, H_ & 3, (H_ % 4 == 0) ? 16 : 8>(smem, q, a, x0, y0, y1, NPT, NT, bar_s);
        } else {
            if (!isB) producer4<H_, NO_, false, -1, 8>(smem, q, a, x0, y0, y1, NPT, NT, bar_s);
            else producer4<H_, NO_, true, -1, 4>(smem, q, a, x0, y0, y1, NPT, NT, bar_s);
        }
    } else {
        // =========================================================================================
        // winner-take-all warps and the feeder warp of the TMA ring
        // =========================================================================================
        const int ct = tid - NPT, cw = ct >> 5, lane = ct & 31;
        if (cw == NCW4) {
            // ---- feeder warp: lane 0 keeps the ring RING_EXTRA rows ahead of the producers -----------------------
            const uint32_t lbytes = (uint32_t)(4 * q.NCTP), rbytes = (uint32_t)q.RBY;
            // first virtual column = image column x0 - h (right) / x0 - h + lofs (left); both copies start 16-byte aligned
            const uint32_t *srcL = a.sp.LE + (size_t)f * a.sp.le_frame + (x0 - H_ + lofs + BmStaged::LPADL);
            const uint8_t *srcR = a.sp.RP + (size_t)f * a.sp.rp_frame + ((x0 - H_ + BmStaged::RPADL) & ~15);
            int slot = 0;
            auto feed = [&](int t) {                              // band-relative ring row t -> the next slot
                const int gy = clampi4(y0 - H_ - 1 + t, 0, a.H - 1);
                const uint32_t bar = bar_s + 8 * slot, dst = ring_s + slot * q.SLOT;
                mbar_expect_tx(bar, lbytes + rbytes);
                bulk_load(dst, srcL + (size_t)gy * a.sp.le_pitch, lbytes, bar);
                bulk_load(dst + q.LFB, srcR + (size_t)gy * a.sp.rp_pitch, rbytes, bar);
                if (++slot == RING) slot = 0;
            };
            if (lane == 0)
                for (int t = 0; t < RING && t <= t_last; t++) feed(t);

            // texture: T(x) = sum over the (2h+1)^2 window of |L' - cap| from the expanded left rows of the
            // ring.  Lane l owns the pixels [8l, 8l + 8): per ring row the 2h + 8 bytes |L' - cap| they touch are compacted
            // from 6 expanded 128-bit words (24 columns), the 8 horizontal window sums come from IDP.4A byte sums (first
            // window, then + entering - leaving byte), and the vertical window slides by adding the entering row's sums and
            // subtracting the leaving row's.  No cross-lane traffic.
            constexpr int NTAP = 2 * H_ + 1;
            const uint32_t capx4 = (uint32_t)a.cap * 0x01010101u;
            const bool tlane = 8 * lane < TWc;
            int T[8];
#pragma unroll
            for (int i = 0; i < 8; i++) T[i] = 0;
            auto row_sums = [&](const uint8_t *slot_p, int (&R)[8]) {
                const uint4 *p4 = reinterpret_cast<const uint4 *>(slot_p + 32 * lane);
                uint32_t w[6];
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    const uint4 v = p4[k];                                        // 4 expanded columns
                    const uint32_t lo = __byte_perm(v.x, v.y, 0x0040), hi = __byte_perm(v.z, v.w, 0x0040);
                    w[k] = __vabsdiffu4(__byte_perm(lo, hi, 0x5410), capx4);      // |L' - cap| <= 63 per byte
                }
                int acc = 0;
#pragma unroll
                for (int k = 0; k < (NTAP + 3) / 4; k++) {
                    const int nb = NTAP - 4 * k >= 4 ? 4 : NTAP - 4 * k;
                    const int m = nb == 4 ? 0x01010101 : (nb == 3 ? 0x00010101 : (nb == 2 ? 0x00000101 : 0x00000001));
                    acc = __dp4a((int)w[k], m, acc);
                }
                R[0] = acc;
#pragma unroll
                for (int i = 1; i < 8; i++) {
                    const int bi = NTAP - 1 + i, bo = i - 1;                      // + byte bi, - byte bo
                    acc = __dp4a((int)w[bi / 4], 1 << (8 * (bi % 4)), acc);
                    acc = __dp4a((int)w[bo / 4], (int)(0xFFu << (8 * (bo % 4))), acc);   // multiplier byte -1
                    R[i] = acc;
                }
            };
            uint16_t *TEX = reinterpret_cast<uint16_t *>(smem + q.TEXOFF);
            for (int t = 0; t <= 2 * H_; t++) {                   // ring rows 0 .. 2h (the first row step removes row 0 again)
                mbar_wait(bar_s + 8 * t, 0);
                if (tlane) {
                    int R[8];
                    row_sums(Ring + t * q.SLOT, R);
#pragma unroll
                    for (int i = 0; i < 8; i++) T[i] += R[i];
                }
            }
            auto tex_row = [&](int r) {                           // texture sums of band row r -> TEX[r & 1]
                const int tin = r + 2 * H_ + 1;
                mbar_wait(bar_s + 8 * (tin % RING), (uint32_t)(tin / RING) & 1u);
                if (tlane) {
                    int Ri[8], Ro[8];
                    row_sums(Ring + (tin % RING) * q.SLOT, Ri);
                    row_sums(Ring + (r % RING) * q.SLOT, Ro);
#pragma unroll
                    for (int i = 0; i < 8; i++) T[i] += Ri[i] - Ro[i];
                    *reinterpret_cast<uint4 *>(TEX + (r & 1) * 256 + 8 * lane) =
                        make_uint4((uint32_t)T[0] | ((uint32_t)T[1] << 16), (uint32_t)T[2] | ((uint32_t)T[3] << 16),
                                   (uint32_t)T[4] | ((uint32_t)T[5] << 16), (uint32_t)T[6] | ((uint32_t)T[7] << 16));
                }
                __syncwarp();                                     // lane 0 refills a ring slot next
            };
            tex_row(0);
#pragma unroll 1
            for (int y = y0; y < y1; y++) {
                // the producers have finished row y: ring row y - y0 (its "out" row) is dead, its slot takes row y - y0 + RING;
                // the winner-take-all warps have finished row y - 1: TEX[(y + 1) & 1] is free
                bar_sync4(1 + (y & 1), NT);
                if (y + 2 < y1) bar_arrive4(3 + (y & 1), NT);   // (this warp reads no sums: it only has to be counted)
                const int t = (y - y0) + RING;
                if (lane == 0 && t <= t_last) feed(t);
                if (y + 1 < y1) tex_row(y - y0 + 1);
            }
            return;
        }

        // ---- the pixel of this thread --------------------------------------------------------------------
        int x = -1;
        {
            const int PW = (TWc + NCW4 - 1) / NCW4;                // pixels per warp (<= 32)
            if (lane < PW && cw * PW + lane < TWc && x0 + cw * PW + lane >= 0) x = cw * PW + lane;
        }
        const int xx = max(x, 0);
        const int xv = xx;                                      // pixel xx's window starts at virtual column xx
        const int gq = xv / G, gi = xv - gq * G;
        const bool second = gi >= H_;
        int oa, ob, oc_;                                        // SAD = [oa] + [ob] - [oc_], byte offsets inside a buffer
        if (!second) {
            oa = q.TAOFF + gq * PP;
            ob = q.XOFF + ((gq + 1) * G + gi) * PP;
            oc_ = gi > 0 ? q.XOFF + (xv - 1) * PP : q.ZOFF;
        } else {
            oa = q.XOFF + (gq * G + H_ + (G - 1 - gi)) * PP;
            ob = q.TBOFF + (gq + 1) * PP;
            oc_ = gi < G - 1 ? q.XOFF + ((gq + 1) * G + H_ + (G - 2 - gi)) * PP : q.ZOFF;
        }
        const int omn = q.MNOFF + xv * MNP;                     // octet minima of this pixel (made by the producers)
        int16_t *dptr = a.disp.p + (size_t)f * a.disp.frame + (size_t)y0 * a.disp.pitch + lofs + x0 + xx;
        int16_t *cptr = a.cost.p ? a.cost.p + (size_t)f * a.cost.frame + (size_t)y0 * a.cost.pitch + lofs + x0 + xx : nullptr;
        const uint16_t *tex16 = reinterpret_cast<const uint16_t *>(smem + q.TEXOFF) + xx;
        const int16_t FILT = (int16_t)(-16);                    // (minD - 1) * 16 with minD = 0

#pragma unroll 1
        for (int y = y0; y < y1; y++) {
            bar_sync4(1 + (y & 1), NT);                         // the sums of row y are complete (and its texture sums)
            const int tsum = (int)tex16[((y - y0) & 1) * 256];
            const uint8_t *buf = smem + (y & 1) * q.BUFSZ;
            int16_t dout = FILT;
            int costv = 0;
            bool okc = false;
            if (x >= 0 && tsum >= a.texThr) {
                const uint8_t *pa = buf + oa, *pb = buf + ob, *pc = buf + oc_;
                auto sad4 = [&](int o, uint32_t (&sv)[4]) {
                    const uint4 t = *reinterpret_cast<const uint4 *>(pa + 16 * o);
                    const uint4 w = *reinterpret_cast<const uint4 *>(pb + 16 * o);
                    const uint4 u = *reinterpret_cast<const uint4 *>(pc + 16 * o);
                    sv[0] = t.x + w.x - u.x; sv[1] = t.y + w.y - u.y; sv[2] = t.z + w.z - u.z; sv[3] = t.w + w.w - u.w;
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
                    sad4(oc, sv);
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
                const int p = (int)a16[dp] + (int)b16[dp] - (int)c16[dp] - BIASC;
                const int n = (int)a16[dn] + (int)b16[dn] - (int)c16[dn] - BIASC;
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
                        sad4(oo, sv);
                        const uint4 Z = c_zmask4[mind - 8 * oo + 1];
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
            if (y + 2 < y1) bar_arrive4(3 + (y & 1), NT);        // the buffer of row y may be overwritten (row y + 2)
            if (x >= 0) {
                *dptr = dout;
                if (cptr && okc) *cptr = (int16_t)costv;
            }
            dptr += a.disp.pitch;
            if (cptr) cptr += a.cost.pitch;
        }
    }
}

struct Tiling4 { int NG, TW, BH, nstripes, nbands, NT; size_t smem; long long cost; };

int sm_count4()
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

bool pick_tiling4(const BmGeom &g, int n, Tiling4 *t)
{
    const int h = g.bs / 2;
    if (g.minD != 0 || h < 2 || h > 7) return false;
    if (!(g.nd == 256 || g.nd == 192 || g.nd == 128 || g.nd == 96 || g.nd == 64 || g.nd == 48 || g.nd == 32)) return false;
    const int NO = g.nd / 8, G = 2 * h;
    const size_t smem_max = 224 * 1024;
    const int E = hp_of(h) - h;                      // pixels stripe 0 starts left of the image
    int ngmax = 2;
    while (producer_threads4(h, NO, ngmax + 1) + (NCW4 + NFW4) * 32 <= MAXT4 && (size_t)make_geo4(h, g.nd, ngmax + 1).total <= smem_max)
        ngmax++;
    if (producer_threads4(h, NO, ngmax) + (NCW4 + NFW4) * 32 > MAXT4 || (size_t)make_geo4(h, g.nd, ngmax).total > smem_max) return false;
    const int twmax = std::min(ngmax * G - 2 * h, NCW4 * 32) & ~3;
    if (twmax < 16) return false;
    // stripe s covers the pixels [s * TW - E, (s + 1) * TW - E): its first virtual column is a multiple of 4 (16-byte aligned bulk copies)
    t->nstripes = cdiv(g.W1 + E, twmax);
    t->TW = (cdiv(g.W1 + E, t->nstripes) + 3) & ~3;
    t->nstripes = cdiv(g.W1 + E, t->TW);
    t->NG = cdiv(t->TW + 2 * h, G);
    if (t->NG > ngmax) return false;
    t->NT = producer_threads4(h, NO, t->NG) + (NCW4 + NFW4) * 32;
    // bands: a band pays ~START rows of start-up (2h+1 rows of window sums) and the launch runs in waves of one CTA per
    // SM -> take the band count with the smallest (waves + 1/2) x (band height + START); the half wave stands for the
    // tail (stripes differ a little), and bands stay <= 128 rows (bm_sad3.cu: taller bands lose more to that tail)
    const int rows = g.row1 - g.row0;
    const int slots = sm_count4(), START = h + 4;
    long long best = -1;
    for (int nb = 1; nb <= std::max(1, rows / 8) && nb <= 64; nb++) {
        const int bh = cdiv(rows, nb);
        if (cdiv(rows, bh) != nb || (bh > 128 && nb < std::max(1, rows / 8))) continue;
        const long long waves = ((long long)n * t->nstripes * nb + slots - 1) / slots;
        const long long cost = (2 * waves + 1) * (bh + START);
        if (best < 0 || cost < best) { best = cost; t->nbands = nb; t->BH = bh; }
    }
    t->cost = best;
    t->smem = (size_t)make_geo4(h, g.nd, t->NG).total;
    return t->smem <= smem_max;
}

bool g_zmask4_ready[64] = {false};

int upload_zmask4()
{
    int dev = 0;
    RTDM_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && g_zmask4_ready[dev]) return 0;
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
    RTDM_CUDA(cudaMemcpyToSymbol(c_zmask4, zm, sizeof(zm)));
    if (dev >= 0 && dev < 64) g_zmask4_ready[dev] = true;
    return 0;
}

template <int H_, int NO_>
int launch4(const Bm4Args &a, const Tiling4 &t, int n, cudaStream_t st)
{
    RTDM_CUDA(cudaFuncSetAttribute(bm_sad4_kernel<H_, NO_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem));
    bm_sad4_kernel<H_, NO_><<<dim3(t.nstripes, t.nbands, n), t.NT, t.smem, st>>>(a);
    return 0;
}

}  // namespace

bool bm_sad4_supported(const BmGeom &g, int n)
{
    Tiling4 t;
    return g.W1 >= 1 && g.row1 > g.row0 && pick_tiling4(g, n, &t);
}

// estimated cost of one launch of n frames in row steps (waves x (band height + start-up)), -1 when the kernel does not
// apply: the host layer sizes its chunks with it
long long bm_sad4_cost(const BmGeom &g, int n)
{
    Tiling4 t;
    if (!(g.W1 >= 1 && g.row1 > g.row0 && pick_tiling4(g, n, &t))) return -1;
    return t.cost;
}

// texture sums + SAD + WTA; the staged planes come from the prefilter (prefilter.cu: BmStaged)
int launch_bm_sad4_core(const BmGeom &g, int n, const BmStaged &sp, PlaneS16 disp, PlaneS16 cost, cudaStream_t st)
{
    Tiling4 t;
    if (!pick_tiling4(g, n, &t)) { set_error("bm_sad4: unsupported geometry"); return -RTDM_EINVAL; }
    int rc = upload_zmask4();
    if (rc) return rc;
    const int h = g.bs / 2;
    Bm4Args a;
    a.disp = disp; a.cost = cost;
    a.W = g.W; a.H = g.H; a.nd = g.nd; a.cap = g.cap; a.texThr = g.texThr; a.uniq = g.uniq;
    a.W1 = g.W1; a.row0 = g.row0; a.row1 = g.row1;
    a.TW = t.TW; a.BH = t.BH; a.NG = t.NG;
    a.sp = sp;
#define RTDM_SAD4_ND(H_) (g.nd == 128 ? launch4<H_, 16>(a, t, n, st) : g.nd == 64 ? launch4<H_, 8>(a, t, n, st) : \
                          g.nd == 192 ? launch4<H_, 24>(a, t, n, st) : g.nd == 96 ? launch4<H_, 12>(a, t, n, st) : \
                          g.nd == 48 ? launch4<H_, 6>(a, t, n, st) : g.nd == 256 ? launch4<H_, 32>(a, t, n, st) : \
                          launch4<H_, 4>(a, t, n, st))
    switch (h) {
        case 2: rc = RTDM_SAD4_ND(2); break;
        case 3: rc = RTDM_SAD4_ND(3); break;
        case 4: rc = RTDM_SAD4_ND(4); break;
        case 5: rc = RTDM_SAD4_ND(5); break;
        case 6: rc = RTDM_SAD4_ND(6); break;
        default: rc = RTDM_SAD4_ND(7); break;
    }
#undef RTDM_SAD4_ND
    if (rc) return rc;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rtdm
