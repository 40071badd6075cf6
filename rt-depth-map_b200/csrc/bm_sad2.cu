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

namespace rtdm {
namespace {

constexpr int NT2 = 192;          // threads per CTA
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
};

// ------------------------------------------------------------------------------------------------
// texture: T(x, y) = sum over the (2h+1)^2 window of |L'[y+j][lcol(x+i)] - cap|, x in [0, W1)
// ------------------------------------------------------------------------------------------------
constexpr int TXW = 64, TXH = 32;
__global__ void __launch_bounds__(256)
bm_texture_kernel(PlaneU8 Lp, uint16_t *tex, size_t tex_pitch, size_t tex_frame, int W, int H, int nd,
                  int cap, int h, int W1, int row0, int row1)
{
    extern __shared__ uint16_t tx_s[];
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * TXW, y0 = row0 + blockIdx.y * TXH;
    const int NCc = TXW + 2 * h, NRr = TXH + 2 * h;
    uint16_t *a = tx_s;                     // [NRr][NCc] |L' - cap|
    uint16_t *v = a + NRr * NCc;            // [TXH][NCc] vertical sums
    const uint8_t *L = Lp.p + (size_t)f * Lp.frame;
    const int lofs = nd - 1;
    for (int i = threadIdx.x; i < NRr * NCc; i += blockDim.x) {
        int r = i / NCc, c = i - r * NCc;
        int gy = clampi(y0 - h + r, 0, H - 1);
        int lc = clampi(x0 - h + c, -lofs, W - lofs - 1) + lofs;
        a[i] = (uint16_t)abs((int)L[(size_t)gy * Lp.pitch + lc] - cap);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TXH * NCc; i += blockDim.x) {
        int r = i / NCc, c = i - r * NCc;
        int s = 0;
        for (int j = 0; j <= 2 * h; j++) s += a[(r + j) * NCc + c];
        v[i] = (uint16_t)s;
    }
    __syncthreads();
    uint16_t *out = tex + (size_t)f * tex_frame;
    for (int i = threadIdx.x; i < TXH * TXW; i += blockDim.x) {
        int r = i / TXW, c = i - r * TXW;
        int gx = x0 + c, gy = y0 + r;
        if (gx >= W1 || gy >= row1) continue;
        int s = 0;
        for (int k = 0; k <= 2 * h; k++) s += v[r * NCc + c + k];
        out[(size_t)gy * tex_pitch + gx] = (uint16_t)s;
    }
}

// ------------------------------------------------------------------------------------------------
// producer: |L - R| of one row for CT adjacent columns x 8 disparities
// ------------------------------------------------------------------------------------------------
// Fast form: the thread's column range is interior, so column kk reads R bytes at (base + kk .. +7):
// shared words + compile-time funnel shifts.
template <int CT>
__device__ __forceinline__ void ad_row_fast(const uint32_t *lw_p, const uint32_t *rw_p, uint32_t (&lo)[CT], uint32_t (&hi)[CT])
{
    constexpr int NLW = (CT + 3) / 4, NRW = (CT + 7 + 3) / 4 + 1;
    uint32_t lw[NLW], rw[NRW];
#pragma unroll
    for (int i = 0; i < NLW; i++) lw[i] = lw_p[i];
#pragma unroll
    for (int i = 0; i < NRW; i++) rw[i] = rw_p[i];
#pragma unroll
    for (int kk = 0; kk < CT; kk++) {
        const int w = kk >> 2, s = kk & 3;
        const uint32_t l4 = __byte_perm(lw[w], 0, s == 0 ? 0x0000 : (s == 1 ? 0x1111 : (s == 2 ? 0x2222 : 0x3333)));
        uint32_t r0, r1;
        if (s == 0) { r0 = rw[w]; r1 = rw[w + 1]; }
        else { r0 = __funnelshift_r(rw[w], rw[w + 1], 8 * s); r1 = __funnelshift_r(rw[w + 1], rw[w + 2], 8 * s); }
        lo[kk] = __vabsdiffu4(l4, r0);
        hi[kk] = __vabsdiffu4(l4, r1);
    }
}

// Border form: per-column R offset (clamped columns read the R window of the nearest unclamped column)
template <int CT>
__device__ __forceinline__ void ad_row_border(const uint8_t *lrow, const uint8_t *rrow0, int c0, int j8, int cmin, int cmax,
                                              uint32_t (&lo)[CT], uint32_t (&hi)[CT])
{
#pragma unroll
    for (int kk = 0; kk < CT; kk++) {
        const int c = c0 + kk;
        const int b = clampi(c, cmin, cmax) + j8;               // byte offset in the unshifted Rv copy
        const uint32_t *rw = reinterpret_cast<const uint32_t *>(rrow0) + (b >> 2);
        const int sh = (b & 3) * 8;
        const uint32_t r0 = __funnelshift_r(rw[0], rw[1], sh), r1 = __funnelshift_r(rw[1], rw[2], sh);
        const uint32_t l4 = (uint32_t)lrow[c] * 0x01010101u;
        lo[kk] = __vabsdiffu4(l4, r0);
        hi[kk] = __vabsdiffu4(l4, r1);
    }
}

template <int H_, int KT_>
__global__ void __launch_bounds__(NT2, 2)
bm_sad2_kernel(Bm2Args a)
{
    constexpr int G = 2 * H_, CT = G * KT_, RING = ring_rows(H_);
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, f = blockIdx.z;
    const int nd = a.nd, NO = a.NO;
    const int x0 = blockIdx.x * a.TW;
    const int TWc = min(a.TW, a.W1 - x0);
    const int y0 = a.row0 + blockIdx.y * a.BH, y1 = min(y0 + a.BH, a.row1);
    if (TWc <= 0 || y0 >= y1) return;
    const int NCT = a.NGT * CT;                        // virtual columns held by the producers

    // ---- shared memory carve-up ----------------------------------------------------------------
    uint8_t *Pre = smem;                                          // [NCT][PP]
    uint8_t *Suf = Pre + (size_t)NCT * a.PP;                      // [NCT][PP]
    uint16_t *Smin = reinterpret_cast<uint16_t *>(Suf + (size_t)NCT * a.PP);   // [NO][NT2]
    uint8_t *Lv = reinterpret_cast<uint8_t *>(Smin + (size_t)NO * NT2);        // [RING][2][LVP]
    uint8_t *Rv = Lv + (size_t)RING * 2 * a.LVP;                               // [RING][2][RVP]

    const uint8_t *Lg = a.Lp.p + (size_t)f * a.Lp.frame;
    const uint8_t *Rg = a.Rp.p + (size_t)f * a.Rp.frame;
    const int lofs = nd - 1;

    // loader: image row gy -> ring slot (gy & (RING-1)); copy 1 is copy 0 shifted left by 2 bytes
    auto load_row = [&](int gy) {
        const int slot = gy & (RING - 1);
        const int gyc = clampi(gy, 0, a.H - 1);
        uint8_t *l0 = Lv + (size_t)slot * 2 * a.LVP, *r0 = Rv + (size_t)slot * 2 * a.RVP;
        const uint8_t *lsrc = Lg + (size_t)gyc * a.Lp.pitch, *rsrc = Rg + (size_t)gyc * a.Rp.pitch;
        const int lwords = a.LVP / 4, rwords = a.RVP / 4;
        for (int i = tid; i < 2 * lwords; i += NT2) {
            const int cp = i >= lwords, w = cp ? i - lwords : i;
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int c = 4 * w + b + 2 * cp;                                     // virtual column
                const int lc = clampi(x0 - H_ + c, -lofs, a.W - lofs - 1) + lofs;
                v |= (uint32_t)lsrc[lc] << (8 * b);
            }
            reinterpret_cast<uint32_t *>(l0 + (size_t)cp * a.LVP)[w] = v;
        }
        for (int i = tid; i < 2 * rwords; i += NT2) {
            const int cp = i >= rwords, w = cp ? i - rwords : i;
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int k = 4 * w + b + 2 * cp;
                v |= (uint32_t)rsrc[clampi(x0 - H_ + k, 0, a.W - 1)] << (8 * b);
            }
            reinterpret_cast<uint32_t *>(r0 + (size_t)cp * a.RVP)[w] = v;
        }
    };

    // ---- producer task of this thread -------------------------------------------------------------
    const int tg = tid / NO, j = tid - tg * NO;
    const bool prod = tg < a.NGT;
    const int c0 = tg * CT;                             // first virtual column of the thread
    const int cp = (c0 & 2) ? 1 : 0;                    // which shifted copy keeps the word loads aligned
    const int lwo = (c0 - 2 * cp) >> 2;                 // word offset of column c0 in copy cp
    const int rwo = (c0 - 2 * cp + 8 * j) >> 2;         // word offset of R byte (c0 + 8j) in copy cp
    // R clamp (App. A.2, minD = 0): rbase(xc) = clip(xc, 0, W - nd); in virtual columns c = xc - x0 + h
    const int cmin = H_ - x0, cmax = (a.W - nd) - x0 + H_;
    const bool border = (c0 < cmin) || (c0 + CT - 1 > cmax);    // thread-uniform within a column group

    uint32_t V[CT][4];
#pragma unroll
    for (int kk = 0; kk < CT; kk++) V[kk][0] = V[kk][1] = V[kk][2] = V[kk][3] = 0u;

    auto ad_row = [&](int gy, uint32_t (&lo)[CT], uint32_t (&hi)[CT]) {
        const int slot = gy & (RING - 1);
        const uint8_t *lrow = Lv + (size_t)slot * 2 * a.LVP;
        const uint8_t *rrow = Rv + (size_t)slot * 2 * a.RVP;
        if (!border)
            ad_row_fast<CT>(reinterpret_cast<const uint32_t *>(lrow + (size_t)cp * a.LVP) + lwo,
                            reinterpret_cast<const uint32_t *>(rrow + (size_t)cp * a.RVP) + rwo, lo, hi);
        else
            ad_row_border<CT>(lrow, rrow, c0, 8 * j, cmin, cmax, lo, hi);
    };

    // ---- prologue: ring rows y0-h .. y0+h, vertical sums over rows y0-h .. y0+h-1 -----------------
    for (int r = y0 - H_; r <= y0 + H_; r++) load_row(r);
    __syncthreads();
    if (prod) {
        for (int r = y0 - H_; r < y0 + H_; r++) {
            uint32_t lo[CT], hi[CT];
            ad_row(r, lo, hi);
#pragma unroll
            for (int kk = 0; kk < CT; kk++) {
                V[kk][0] += __byte_perm(lo[kk], 0, 0x4140);
                V[kk][1] += __byte_perm(lo[kk], 0, 0x4342);
                V[kk][2] += __byte_perm(hi[kk], 0, 0x4140);
                V[kk][3] += __byte_perm(hi[kk], 0, 0x4342);
            }
        }
    }

    int16_t *dispf = a.disp.p + (size_t)f * a.disp.frame;
    int16_t *costf = a.cost.p ? a.cost.p + (size_t)f * a.cost.frame : nullptr;
    const uint16_t *texf = a.tex + (size_t)f * a.tex_frame;
    const int16_t FILT = (int16_t)(-16);                // (minD - 1) * 16 with minD = 0

    for (int y = y0; y < y1; y++) {
        // prefetch the next row into the ring (consumed after the next two barriers)
        if (y + 1 < y1) load_row(y + 1 + H_);
        // ---------------- producer ---------------------------------------------------------------
        if (prod) {
            uint32_t lo[CT], hi[CT];
            ad_row(y + H_, lo, hi);
            if (y > y0) {
                uint32_t olo[CT], ohi[CT];
                ad_row(y - H_ - 1, olo, ohi);
#pragma unroll
                for (int kk = 0; kk < CT; kk++) {
                    const uint32_t dl = lo[kk] + 0x80808080u - olo[kk];      // per byte: in + 128 - out
                    const uint32_t dh = hi[kk] + 0x80808080u - ohi[kk];
                    V[kk][0] += __byte_perm(dl, 0, 0x4140) - 0x00800080u;
                    V[kk][1] += __byte_perm(dl, 0, 0x4342) - 0x00800080u;
                    V[kk][2] += __byte_perm(dh, 0, 0x4140) - 0x00800080u;
                    V[kk][3] += __byte_perm(dh, 0, 0x4342) - 0x00800080u;
                }
            } else {
#pragma unroll
                for (int kk = 0; kk < CT; kk++) {
                    V[kk][0] += __byte_perm(lo[kk], 0, 0x4140);
                    V[kk][1] += __byte_perm(lo[kk], 0, 0x4342);
                    V[kk][2] += __byte_perm(hi[kk], 0, 0x4140);
                    V[kk][3] += __byte_perm(hi[kk], 0, 0x4342);
                }
            }
            // in-group prefix and suffix sums
#pragma unroll
            for (int g = 0; g < KT_; g++) {
                uint4 p = make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int i = 0; i < G; i++) {
                    const int kk = g * G + i;
                    p.x += V[kk][0]; p.y += V[kk][1]; p.z += V[kk][2]; p.w += V[kk][3];
                    *reinterpret_cast<uint4 *>(Pre + (size_t)(c0 + kk) * a.PP + 16 * j) = p;
                }
                uint4 s = make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int i = G - 1; i >= 0; i--) {
                    const int kk = g * G + i;
                    s.x += V[kk][0]; s.y += V[kk][1]; s.z += V[kk][2]; s.w += V[kk][3];
                    *reinterpret_cast<uint4 *>(Suf + (size_t)(c0 + kk) * a.PP + 16 * j) = s;
                }
            }
        }
        __syncthreads();

        // ---------------- consumer: one thread per pixel -----------------------------------------------
        for (int x = tid; x < TWc; x += NT2) {
            const uint8_t *sufp = Suf + (size_t)x * a.PP;
            const uint8_t *prep = Pre + (size_t)(x + G) * a.PP;
            const int tsum = texf[(size_t)y * a.tex_pitch + x0 + x];
            int16_t dout = FILT;
            if (tsum >= a.texThr) {
                // pass 1: octet minima and the argmin octet
                uint32_t best = 0xFFFFFFFFu;
                for (int o = 0; o < NO; o++) {
                    const uint4 u = *reinterpret_cast<const uint4 *>(sufp + 16 * o);
                    const uint4 w = *reinterpret_cast<const uint4 *>(prep + 16 * o);
                    uint32_t m = __vimin3_u16x2(u.x + w.x, u.y + w.y, u.z + w.z);
                    m = __vminu2(m, u.w + w.w);
                    const uint32_t mm = min(m & 0xFFFFu, m >> 16);
                    Smin[o * NT2 + tid] = (uint16_t)mm;
                    best = min(best, mm * 65536u + (uint32_t)o);
                }
                const int minsad = (int)(best >> 16), oc = (int)(best & 0xFFFFu);
                // exact position inside the argmin octet (first minimum)
                int mind;
                {
                    const uint4 u = *reinterpret_cast<const uint4 *>(sufp + 16 * oc);
                    const uint4 w = *reinterpret_cast<const uint4 *>(prep + 16 * oc);
                    const uint32_t s0 = u.x + w.x, s1 = u.y + w.y, s2 = u.z + w.z, s3 = u.w + w.w;
                    int idx = 7;
                    if ((int)(s3 & 0xFFFFu) == minsad) idx = 6;
                    if ((int)(s2 >> 16) == minsad) idx = 5;
                    if ((int)(s2 & 0xFFFFu) == minsad) idx = 4;
                    if ((int)(s1 >> 16) == minsad) idx = 3;
                    if ((int)(s1 & 0xFFFFu) == minsad) idx = 2;
                    if ((int)(s0 >> 16) == minsad) idx = 1;
                    if ((int)(s0 & 0xFFFFu) == minsad) idx = 0;
                    mind = 8 * oc + idx;
                }
                const uint16_t *suf16 = reinterpret_cast<const uint16_t *>(sufp);
                const uint16_t *pre16 = reinterpret_cast<const uint16_t *>(prep);
                const int dp = mind + 1 < nd ? mind + 1 : nd - 2, dn = mind > 0 ? mind - 1 : 1;
                const int p = (int)suf16[dp] + (int)pre16[dp], n = (int)suf16[dn] + (int)pre16[dn];
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
                        const uint4 u = *reinterpret_cast<const uint4 *>(sufp + 16 * oo);
                        const uint4 w = *reinterpret_cast<const uint4 *>(prep + 16 * oo);
                        const int rel = mind - 8 * oo;                        // -1 .. 8
                        const uint32_t Z = (7u << (rel + 1)) >> 2;            // bit p set: position p is excluded
                        uint32_t s[4] = {u.x + w.x, u.y + w.y, u.z + w.z, u.w + w.w};
                        uint32_t mz = 0xFFFFFFFFu;
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            uint32_t msk = ((Z >> (2 * r)) & 1u ? 0x0000FFFFu : 0u) | ((Z >> (2 * r + 1)) & 1u ? 0xFFFF0000u : 0u);
                            mz = __vminu2(mz, s[r] | msk);
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
        __syncthreads();
    }
}

struct Tiling2 { int KT, CT, NO, NGT, TW, BH, nstripes, nbands, LVP, RVP, PP; size_t smem; };

bool pick_tiling2(const BmGeom &g, int n, Tiling2 *t)
{
    const int h = g.bs / 2;
    if (g.minD != 0 || h < 2 || h > 7 || g.nd > 256 || g.nd < 16) return false;
    t->KT = h == 2 ? 3 : (h == 3 ? 2 : 1);
    t->CT = 2 * h * t->KT;
    t->NO = g.nd / 8;
    t->NGT = std::min(NT2 / t->NO, std::max(1, 192 / t->CT));
    int twmax = t->NGT * t->CT - 2 * h;
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
    t->smem = (size_t)2 * NCT * t->PP + (size_t)t->NO * NT2 * 2 + (size_t)ring_rows(h) * 2 * (t->LVP + t->RVP);
    return t->smem <= 112 * 1024;
}

template <int H_, int KT_>
int launch2(const Bm2Args &a, const Tiling2 &t, int n, cudaStream_t st)
{
    RTDM_CUDA(cudaFuncSetAttribute(bm_sad2_kernel<H_, KT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    bm_sad2_kernel<H_, KT_><<<dim3(t.nstripes, t.nbands, n), NT2, t.smem, st>>>(a);
    return 0;
}

}  // namespace

bool bm_sad2_supported(const BmGeom &g, int n)
{
    Tiling2 t;
    return g.W1 >= 1 && g.row1 > g.row0 && pick_tiling2(g, n, &t);
}

// tex: scratch of n * H * tex_pitch uint16
int launch_bm_sad2(const BmGeom &g, int n, PlaneU8 Lp, PlaneU8 Rp, PlaneS16 disp, PlaneS16 cost,
                   uint16_t *tex, size_t tex_pitch, size_t tex_frame, cudaStream_t st, int *launches)
{
    Tiling2 t;
    if (!pick_tiling2(g, n, &t)) { set_error("bm_sad2: unsupported geometry"); return -RTDM_EINVAL; }
    const int h = g.bs / 2;
    {
        dim3 grid(cdiv(g.W1, TXW), cdiv(g.row1 - g.row0, TXH), n);
        size_t smem = ((size_t)(TXH + 2 * h) * (TXW + 2 * h) + (size_t)TXH * (TXW + 2 * h)) * sizeof(uint16_t);
        bm_texture_kernel<<<grid, 256, smem, st>>>(Lp, tex, tex_pitch, tex_frame, g.W, g.H, g.nd, g.cap, h, g.W1, g.row0, g.row1);
    }
    Bm2Args a;
    a.Lp = Lp; a.Rp = Rp; a.disp = disp; a.cost = cost;
    a.tex = tex; a.tex_pitch = tex_pitch; a.tex_frame = tex_frame;
    a.W = g.W; a.H = g.H; a.nd = g.nd; a.cap = g.cap; a.texThr = g.texThr; a.uniq = g.uniq;
    a.W1 = g.W1; a.row0 = g.row0; a.row1 = g.row1;
    a.TW = t.TW; a.BH = t.BH; a.NO = t.NO; a.NGT = t.NGT; a.LVP = t.LVP; a.RVP = t.RVP; a.PP = t.PP;
    int rc = 0;
    switch (h) {
        case 2: rc = launch2<2, 3>(a, t, n, st); break;
        case 3: rc = launch2<3, 2>(a, t, n, st); break;
        case 4: rc = launch2<4, 1>(a, t, n, st); break;
        case 5: rc = launch2<5, 1>(a, t, n, st); break;
        case 6: rc = launch2<6, 1>(a, t, n, st); break;
        default: rc = launch2<7, 1>(a, t, n, st); break;
    }
    if (rc) return rc;
    if (launches) (*launches) += 2;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rtdm
