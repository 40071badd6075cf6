// bm_sad.cu -- Konolige block matching core: SAD cost over all disparities, winner-take-all with
// texture / uniqueness tests and sub-pixel refinement.  The cost volume never leaves the SM.
//
// Replaces findStereoCorrespondenceBM inside cv::StereoBM::compute as reached from
// SWMatcherKonolige::compute (reference stereo-matcher/bm-sw.cpp:33-38).  Arithmetic per
// SURVEY.md App. A.2 (restated and pinned in oracle/stereo_oracle.c: orc_bm_core).
//
// Decomposition (one CTA = a stripe of TW computed columns x a band of BH rows of one frame):
//   prologue : the prefiltered row band (BH + 2h rows; left stripe and right stripe + nd columns)
//              is staged once in shared memory.
//   stage 1  : thread (virtual column c, disparity octet j) keeps the VERTICAL window sum
//              V(c, y, 8j..8j+7) in 4 registers as packed u16x2 and slides it down one row:
//              V += |L - R|(row y+h) - |L - R|(row y-h-1) with VABSDIFF4 on 4 disparities at a time.
//   stage 2  : thread (column segment, octet) slides the HORIZONTAL window over V in shared
//              memory (LDS.128 in / out, packed adds) -> SAD(x, y, 0..nd) in shared memory.
//   stage 3  : thread per pixel: 32-bit (SAD << 16 | d) keys -> first-minimum argmin, texture test,
//              uniqueness test (packed u16x2 minimum with the argmin neighbourhood masked out),
//              sub-pixel interpolation; writes disparity (x16) and cost.
// All sums are exact (max 2*cap*bs^2 <= 65535), so the order of summation is free.
#include "common.cuh"

namespace rtdm {

namespace {

constexpr int NT = 256;     // threads per CTA
constexpr int KT = 5;       // stage-1 tasks (column, octet) per thread
constexpr int SEG = 16;     // stage-2 segment length (columns)

struct BmKArgs {
    PlaneU8 Lp, Rp;
    PlaneS16 disp, cost;
    int W, H, nd, minD, h, cap, texThr, uniq;
    int lofs, rofs, W1, row0, row1;
    int TW, BH;              // stripe width (computed columns), band height (rows)
    int NO;                  // nd / 8
    int NC;                  // TW + 2h virtual columns
    int LP, RP;              // smem row pitches (bytes) of the left / right band
    int VP, SP;              // smem pitches (bytes) of the V rows and SAD rows
    int16_t *spill;          // BmGeom::spill
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ uint4 add4(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ uint4 addsub4(uint4 a, uint4 b, uint4 c)
{
    return make_uint4(a.x + b.x - c.x, a.y + b.y - c.y, a.z + b.z - c.z, a.w + b.w - c.w);
}

// 8 absolute differences |l - R[a..a+7]| from an unaligned position in a shared-memory row
__device__ __forceinline__ void ad8(const uint32_t *rw, int wa, int sh, uint32_t l4, uint32_t &lo, uint32_t &hi)
{
    uint32_t w0 = rw[wa], w1 = rw[wa + 1], w2 = rw[wa + 2];
    uint32_t r0 = __funnelshift_r(w0, w1, sh), r1 = __funnelshift_r(w1, w2, sh);
    lo = __vabsdiffu4(l4, r0);
    hi = __vabsdiffu4(l4, r1);
}

__global__ void __launch_bounds__(NT, 3)
bm_sad_wta_kernel(BmKArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x;
    const int f = blockIdx.z;
    const int h = a.h, nd = a.nd, NO = a.NO;
    const int x0 = blockIdx.x * a.TW;                       // first computed column of the stripe
    const int TWc = min(a.TW, a.W1 - x0);                   // columns actually present
    const int y0 = a.row0 + blockIdx.y * a.BH;
    const int y1 = min(y0 + a.BH, a.row1);
    if (TWc <= 0 || y0 >= y1) return;
    const int NC = a.NC;
    const int nrows = (y1 - y0) + 2 * h;                    // band rows y0-h .. y1+h-1

    // ---- shared memory carve-up -----------------------------------------------------------
    uint8_t *Ls = smem;
    uint8_t *Rs = Ls + (size_t)(a.BH + 2 * h) * a.LP;
    uint8_t *Vs = Rs + (size_t)(a.BH + 2 * h) * a.RP;
    uint8_t *Ss = Vs + (size_t)NC * a.VP;
    int *Ts = reinterpret_cast<int *>(Ss + (size_t)a.TW * a.SP);

    // virtual column xc = x0 - h + c ; clamped source columns (App. A.2)
    const int lc_first = clampi(x0 - h, -a.lofs, a.W - a.lofs - 1) + a.lofs;
    const int rb_first = clampi(x0 - h, -a.rofs, a.W - a.rofs - nd) + a.rofs;
    const int lc_al = lc_first & ~3, rb_al = rb_first & ~3;

    // ---- prologue: stage the prefiltered band ---------------------------------------------
    {
        const uint8_t *Lg = a.Lp.p + (size_t)f * a.Lp.frame;
        const uint8_t *Rg = a.Rp.p + (size_t)f * a.Rp.frame;
        const int lw = a.LP / 4, rw = a.RP / 4;
        for (int i = tid; i < nrows * lw; i += NT) {
            int r = i / lw, w = i - r * lw;
            int gy = clampi(y0 - h + r, 0, a.H - 1);
            reinterpret_cast<uint32_t *>(Ls + (size_t)r * a.LP)[w] =
                *reinterpret_cast<const uint32_t *>(Lg + (size_t)gy * a.Lp.pitch + lc_al + 4 * w);
        }
        for (int i = tid; i < nrows * rw; i += NT) {
            int r = i / rw, w = i - r * rw;
            int gy = clampi(y0 - h + r, 0, a.H - 1);
            reinterpret_cast<uint32_t *>(Rs + (size_t)r * a.RP)[w] =
                *reinterpret_cast<const uint32_t *>(Rg + (size_t)gy * a.Rp.pitch + rb_al + 4 * w);
        }
    }

    // ---- per-task constants ------------------------------------------------------------------
    int loff[KT], rwa[KT], rsh[KT], vofs[KT];
    uint32_t V[KT][4];
    const int ntask = NC * NO;
#pragma unroll
    for (int k = 0; k < KT; k++) {
        int t = tid + k * NT;
        int c = t / NO, j = t - c * NO;
        if (t >= ntask) { c = 0; j = 0; vofs[k] = -1; } else vofs[k] = c * a.VP + j * 16;
        int xc = x0 - h + c;
        int lc = clampi(xc, -a.lofs, a.W - a.lofs - 1) + a.lofs;
        int rb = clampi(xc, -a.rofs, a.W - a.rofs - nd) + a.rofs;
        loff[k] = lc - lc_al;
        int ra = rb - rb_al + 8 * j;
        rwa[k] = ra >> 2;
        rsh[k] = (ra & 3) * 8;
        V[k][0] = V[k][1] = V[k][2] = V[k][3] = 0u;
    }
    // texture column sums T(c) = sum over the window rows of |L' - cap|: thread c < NC owns column c
    int Tc = 0, tloff = 0;
    if (tid < NC) tloff = clampi(x0 - h + tid, -a.lofs, a.W - a.lofs - 1) + a.lofs - lc_al;
    __syncthreads();

    // ---- vertical window prologue: rows y0-h .. y0+h-1 (band rows 0 .. 2h-1) ------------------
    for (int r = 0; r < 2 * h; r++) {
        const uint8_t *Lr = Ls + (size_t)r * a.LP;
        const uint32_t *Rr = reinterpret_cast<const uint32_t *>(Rs + (size_t)r * a.RP);
#pragma unroll
        for (int k = 0; k < KT; k++) {
            if (vofs[k] < 0) continue;
            uint32_t l = Lr[loff[k]];
            uint32_t lo, hi;
            ad8(Rr, rwa[k], rsh[k], l * 0x01010101u, lo, hi);
            V[k][0] += __byte_perm(lo, 0, 0x4140);
            V[k][1] += __byte_perm(lo, 0, 0x4342);
            V[k][2] += __byte_perm(hi, 0, 0x4140);
            V[k][3] += __byte_perm(hi, 0, 0x4342);
        }
        if (tid < NC) Tc += abs((int)Lr[tloff] - a.cap);
    }

    int16_t *dispf = a.disp.p + (size_t)f * a.disp.frame;
    int16_t *costf = a.cost.p ? a.cost.p + (size_t)f * a.cost.frame : nullptr;
    const int16_t FILT = (int16_t)((a.minD - 1) * 16);

    for (int y = y0; y < y1; y++) {
        // ---------------- stage 1: slide the vertical sums to rows y-h .. y+h -------------------
        {
            const int rin = (y - y0) + 2 * h;                 // band row of image row y+h
            const uint8_t *Lr = Ls + (size_t)rin * a.LP;
            const uint32_t *Rr = reinterpret_cast<const uint32_t *>(Rs + (size_t)rin * a.RP);
            const bool has_out = (y > y0);
            const int rout = (y - y0) - 1;                    // band row of image row y-h-1
            const uint8_t *Lo = Ls + (size_t)max(rout, 0) * a.LP;
            const uint32_t *Ro = reinterpret_cast<const uint32_t *>(Rs + (size_t)max(rout, 0) * a.RP);
#pragma unroll
            for (int k = 0; k < KT; k++) {
                if (vofs[k] < 0) continue;
                uint32_t l = Lr[loff[k]];
                uint32_t lo, hi;
                ad8(Rr, rwa[k], rsh[k], l * 0x01010101u, lo, hi);
                if (has_out) {
                    uint32_t lp = Lo[loff[k]];
                    uint32_t olo, ohi;
                    ad8(Ro, rwa[k], rsh[k], lp * 0x01010101u, olo, ohi);
                    // per-byte (in + 128 - out): no borrow can cross a byte
                    lo = lo + 0x80808080u - olo;
                    hi = hi + 0x80808080u - ohi;
                    V[k][0] += __byte_perm(lo, 0, 0x4140) - 0x00800080u;
                    V[k][1] += __byte_perm(lo, 0, 0x4342) - 0x00800080u;
                    V[k][2] += __byte_perm(hi, 0, 0x4140) - 0x00800080u;
                    V[k][3] += __byte_perm(hi, 0, 0x4342) - 0x00800080u;
                } else {
                    V[k][0] += __byte_perm(lo, 0, 0x4140);
                    V[k][1] += __byte_perm(lo, 0, 0x4342);
                    V[k][2] += __byte_perm(hi, 0, 0x4140);
                    V[k][3] += __byte_perm(hi, 0, 0x4342);
                }
                *reinterpret_cast<uint4 *>(Vs + vofs[k]) = make_uint4(V[k][0], V[k][1], V[k][2], V[k][3]);
            }
            if (tid < NC) {
                Tc += abs((int)Lr[tloff] - a.cap);
                if (has_out) Tc -= abs((int)Lo[tloff] - a.cap);
                Ts[tid] = Tc;
            }
        }
        __syncthreads();

        // ---------------- stage 2: horizontal window over the V columns -------------------------
        {
            const int nseg = (TWc + SEG - 1) / SEG;
            for (int t = tid; t < nseg * NO; t += NT) {
                int s = t / NO, j = t - s * NO;
                int xs = s * SEG, xe = min(xs + SEG, TWc);
                const uint8_t *vp = Vs + (size_t)xs * a.VP + j * 16;
                uint4 acc = *reinterpret_cast<const uint4 *>(vp);
                for (int k = 1; k <= 2 * h; k++)
                    acc = add4(acc, *reinterpret_cast<const uint4 *>(vp + (size_t)k * a.VP));
                uint8_t *sp = Ss + (size_t)xs * a.SP + j * 16;
                *reinterpret_cast<uint4 *>(sp) = acc;
                for (int x = xs + 1; x < xe; x++) {
                    vp += a.VP; sp += a.SP;
                    acc = addsub4(acc, *reinterpret_cast<const uint4 *>(vp + (size_t)(2 * h) * a.VP),
                                  *reinterpret_cast<const uint4 *>(vp - a.VP));
                    *reinterpret_cast<uint4 *>(sp) = acc;
                }
            }
        }
        __syncthreads();

        // ---------------- stage 3: winner-take-all per pixel -------------------------------------
        for (int x = tid; x < TWc; x += NT) {
            uint8_t *srow = Ss + (size_t)x * a.SP;
            int tsum = 0;
            for (int k = 0; k <= 2 * h; k++) tsum += Ts[x + k];
            int16_t dout = FILT;
            if (tsum >= a.texThr) {
                // pass 1: first minimum over d via (SAD << 16 | d) keys
                uint32_t best = 0xFFFFFFFFu;
                uint32_t dd = 0;
                for (int o = 0; o < NO; o++, dd += 8) {
                    uint4 v = *reinterpret_cast<const uint4 *>(srow + o * 16);
                    uint32_t k0 = __byte_perm(v.x, dd + 0, 0x1054), k1 = __byte_perm(v.x, dd + 1, 0x3254);
                    uint32_t k2 = __byte_perm(v.y, dd + 2, 0x1054), k3 = __byte_perm(v.y, dd + 3, 0x3254);
                    uint32_t k4 = __byte_perm(v.z, dd + 4, 0x1054), k5 = __byte_perm(v.z, dd + 5, 0x3254);
                    uint32_t k6 = __byte_perm(v.w, dd + 6, 0x1054), k7 = __byte_perm(v.w, dd + 7, 0x3254);
                    best = __vimin3_u32(best, k0, k1);
                    best = __vimin3_u32(best, k2, k3);
                    best = __vimin3_u32(best, k4, k5);
                    best = __vimin3_u32(best, k6, k7);
                }
                const int minsad = (int)(best >> 16), mind = (int)(best & 0xFFFFu);
                const uint16_t *s16 = reinterpret_cast<const uint16_t *>(srow);
                const int p = s16[mind + 1 < nd ? mind + 1 : nd - 2];
                const int n = s16[mind > 0 ? mind - 1 : 1];
                bool ok = true;
                if (a.uniq > 0) {
                    const int thresh = minsad + (minsad * a.uniq / 100);
                    uint16_t *w16 = reinterpret_cast<uint16_t *>(srow);
                    if (mind > 0) w16[mind - 1] = 0xFFFFu;
                    w16[mind] = 0xFFFFu;
                    if (mind + 1 < nd) w16[mind + 1] = 0xFFFFu;
                    uint32_t m2 = 0xFFFFFFFFu;
                    for (int o = 0; o < NO; o++) {
                        uint4 v = *reinterpret_cast<const uint4 *>(srow + o * 16);
                        m2 = __vimin3_u16x2(m2, v.x, v.y);
                        m2 = __vimin3_u16x2(m2, v.z, v.w);
                    }
                    const int mm = (int)min(m2 & 0xFFFFu, m2 >> 16);
                    ok = !(mm <= thresh);
                }
                if (ok) {
                    const int q = p + n - 2 * minsad + abs(p - n);
                    const int v = (nd - mind - 1 + a.minD) * 256 + (q != 0 ? ((p - n) * 256) / q : 0) + 15;
                    dout = (int16_t)(v >> 4);
                    if (costf && a.lofs + x0 + x < a.W) costf[(size_t)y * a.cost.pitch + a.lofs + x0 + x] = (int16_t)minsad;
                }
            }
            // minDisparity > 0: the last minD computed columns lie beyond the row (cv2 lets them run into the next row)
            const int xo = a.lofs + x0 + x;
            if (xo < a.W) dispf[(size_t)y * a.disp.pitch + xo] = dout;
            else if (a.spill && y == a.row1 - 1) a.spill[(size_t)blockIdx.z * a.minD + (xo - a.W)] = dout;
        }
        // the next iteration's stage-1 writes touch only Vs/Ts (read in stage 2, already fenced by
        // the second barrier); Ss is rewritten only after the next first barrier.
    }
}

struct Tiling { int TW, BH, nstripes, nbands, NC, LP, RP, VP, SP; size_t smem; };

bool pick_tiling(const BmGeom &g, Tiling *t)
{
    const int h = g.bs / 2, NO = g.nd / 8;
    int twmax = (NT * KT) / NO - 2 * h;
    if (twmax > 64) twmax = 64;
    if (twmax < 4) return false;
    t->nstripes = cdiv(g.W1, twmax);
    t->TW = cdiv(g.W1, t->nstripes);
    const int rows = g.row1 - g.row0;
    const int bhmax = 48;
    t->nbands = cdiv(rows, bhmax);
    t->BH = cdiv(rows, t->nbands);
    t->NC = t->TW + 2 * h;
    t->LP = (int)align_up(t->NC + 4, 4);
    t->RP = (int)align_up(t->NC + g.nd + 16, 4);
    t->VP = g.nd * 2;
    t->SP = g.nd * 2 + 16;
    size_t s = (size_t)(t->BH + 2 * h) * (t->LP + t->RP);
    s = align_up(s, 16);
    // keep Vs 16-byte aligned: the L/R band sizes are multiples of 4 only
    t->smem = s + (size_t)t->NC * t->VP + (size_t)t->TW * t->SP + (size_t)t->NC * sizeof(int) + 16;
    return true;
}

}  // namespace

size_t bm_sad_smem_bytes(const BmGeom &g, int, int)
{
    Tiling t;
    if (!pick_tiling(g, &t)) return 0;
    return t.smem;
}

int launch_bm_sad_wta(const BmGeom &g, int n, PlaneU8 Lp, PlaneU8 Rp, PlaneS16 disp, PlaneS16 cost,
                      cudaStream_t st, int *launches)
{
    if (n <= 0 || g.row1 <= g.row0 || g.W1 < 1) return 0;
    Tiling t;
    if (!pick_tiling(g, &t)) {
        set_error("bm: blockSize / numDisparities combination not supported by the kernel tiling");
        return -RTDM_EINVAL;
    }
    BmKArgs a;
    a.Lp = Lp; a.Rp = Rp; a.disp = disp; a.cost = cost;
    a.W = g.W; a.H = g.H; a.nd = g.nd; a.minD = g.minD; a.h = g.bs / 2; a.cap = g.cap;
    a.texThr = g.texThr; a.uniq = g.uniq; a.lofs = g.lofs; a.rofs = g.rofs; a.W1 = g.W1;
    a.row0 = g.row0; a.row1 = g.row1; a.spill = g.spill;
    a.TW = t.TW; a.BH = t.BH; a.NO = g.nd / 8; a.NC = t.NC;
    a.LP = t.LP; a.RP = t.RP; a.VP = t.VP; a.SP = t.SP;
    // the L/R bands must end on a 16-byte boundary so that Vs (uint4 accesses) is aligned
    size_t band = (size_t)(t.BH + 2 * a.h) * (t.LP + t.RP);
    if (band % 16) {
        // grow RP so that the band size is a multiple of 16
        int rowsb = t.BH + 2 * a.h;
        while (((size_t)rowsb * (t.LP + a.RP)) % 16) a.RP += 4;
        band = (size_t)rowsb * (t.LP + a.RP);
    }
    size_t smem = band + (size_t)t.NC * t.VP + (size_t)t.TW * t.SP + (size_t)t.NC * sizeof(int);
    RTDM_CUDA(cudaFuncSetAttribute(bm_sad_wta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (smem > 200 * 1024) {
        set_error("bm: shared-memory tile too large for these parameters");
        return -RTDM_EINVAL;
    }
    dim3 grid(t.nstripes, t.nbands, n);
    bm_sad_wta_kernel<<<grid, NT, smem, st>>>(a);
    if (launches) (*launches)++;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rtdm
