// prefilter.cu -- StereoBM prefilters (x-Sobel and normalized response) for both images of a batch.
//
// Replaces the prefilter stage inside cv::StereoBM::compute as reached from
// SWMatcherKonolige::compute (reference stereo-matcher/bm-sw.cpp:33-38); algorithm per
// SURVEY.md App. A.1.  HBM-bound elementwise work: one thread per 4 output pixels, both images and
// all frames folded into one launch (blockIdx.z = 2*frame + {0: left, 1: right}).
#include "common.cuh"

namespace rtdm {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ---- the layouts bm_sad4.cu stages with TMA (common.cuh: BmStaged); LE == nullptr: not wanted ------------------
__device__ __forceinline__ uint4 splat4(uint32_t v) { return make_uint4(v, v, v, v); }
// one aligned word o = pixels x0 .. x0+3 of row y (all inside the image; x0 % 4 == 0)
__device__ __forceinline__ void staged_store_word(const BmStaged &s, int img, int f, int y, int x0, uint32_t o, bool first, bool last)
{
    if (img == 0) {
        uint32_t *e = s.LE + (size_t)f * s.le_frame + (size_t)y * s.le_pitch + BmStaged::LPADL + x0;     // 4 (mod 16) bytes
        e[0] = __byte_perm(o, 0, 0x0000); e[1] = __byte_perm(o, 0, 0x1111); e[2] = __byte_perm(o, 0, 0x2222); e[3] = __byte_perm(o, 0, 0x3333);
        if (first) e[-1] = e[0];
        if (last) {
            const uint32_t v = __byte_perm(o, 0, 0x3333);
#pragma unroll
            for (int i = 0; i < BmStaged::LPADR; i++) e[4 + i] = v;
        }
    } else {
        uint8_t *r = s.RP + (size_t)f * s.rp_frame + (size_t)y * s.rp_pitch;
        *reinterpret_cast<uint32_t *>(r + BmStaged::RPADL + x0) = o;
        if (first) *reinterpret_cast<uint4 *>(r) = splat4(__byte_perm(o, 0, 0x0000));                // BmStaged::RPADL = 16
        if (last) {
            const uint32_t v = __byte_perm(o, 0, 0x3333);
            uint32_t *t = reinterpret_cast<uint32_t *>(r + BmStaged::RPADL + x0 + 4);
            t[0] = v; t[1] = v; t[2] = v; t[3] = v;                                                   // BmStaged::RPADR = 16
        }
    }
}
// one pixel (any width / alignment)
__device__ __forceinline__ void staged_store_px(const BmStaged &s, int img, int f, int y, int x, int W, uint32_t v)
{
    if (img == 0) {
        uint32_t *e = s.LE + (size_t)f * s.le_frame + (size_t)y * s.le_pitch + BmStaged::LPADL;
        e[x] = v * 0x01010101u;
        if (x == 0) e[-1] = v * 0x01010101u;
        if (x == W - 1)
            for (int i = 0; i < BmStaged::LPADR; i++) e[W + i] = v * 0x01010101u;
    } else {
        uint8_t *r = s.RP + (size_t)f * s.rp_frame + (size_t)y * s.rp_pitch;
        r[BmStaged::RPADL + x] = (uint8_t)v;
        if (x == 0)
            for (int i = 0; i < BmStaged::RPADL; i++) r[i] = (uint8_t)v;
        if (x == W - 1)
            for (int i = 0; i < BmStaged::RPADR; i++) r[BmStaged::RPADL + W + i] = (uint8_t)v;
    }
}

// x-Sobel: dst = clip(d(y-1) + 2 d(y) + d(y+1), -cap, cap) + cap with d(r) = r[x+1]-r[x-1];
// rows reflect-101, first/last column = cap, and an odd last row = cap (OpenCV pairs rows).
__global__ void __launch_bounds__(256)
prefilter_xsobel_kernel(PlaneU8 left, PlaneU8 right, PlaneU8W outL, PlaneU8W outR,
                        int W, int H, int cap, BmStaged sg)
{
    const int img = blockIdx.z & 1, f = blockIdx.z >> 1;
    const uint8_t *src = (img ? right.p + (size_t)f * right.frame : left.p + (size_t)f * left.frame);
    const size_t sp = img ? right.pitch : left.pitch;
    uint8_t *dst = (img ? outR.p + (size_t)f * outR.frame : outL.p + (size_t)f * outL.frame);
    const size_t dp = img ? outR.pitch : outL.pitch;
    const int y = blockIdx.y;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x0 >= W) return;
    const int paired = (H > 1) ? (H & ~1) : 0;
    uint8_t o[4];
    if (y >= paired) {
        o[0] = o[1] = o[2] = o[3] = (uint8_t)cap;
    } else {
        const int ya = y > 0 ? y - 1 : 1;
        const int yb = y < H - 1 ? y + 1 : H - 2;
        const uint8_t *r0 = src + (size_t)ya * sp, *r1 = src + (size_t)y * sp, *r2 = src + (size_t)yb * sp;
        // column sums c[x] = r0[x] + 2 r1[x] + r2[x] for x0-1 .. x0+4
        int c[6];
#pragma unroll
        for (int i = 0; i < 6; i++) {
            int x = clampi(x0 - 1 + i, 0, W - 1);
            c[i] = (int)r0[x] + 2 * (int)r1[x] + (int)r2[x];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int x = x0 + i;
            int v = c[i + 2] - c[i];
            v = clampi(v, -cap, cap) + cap;
            if (x == 0 || x >= W - 1) v = cap;
            o[i] = (uint8_t)v;
        }
    }
    if (sg.LE) {
        for (int i = 0; i < 4 && x0 + i < W; i++) staged_store_px(sg, img, f, y, x0 + i, W, o[i]);
        return;
    }
    uint8_t *d = dst + (size_t)y * dp + x0;
    if (x0 + 3 < W && ((reinterpret_cast<uintptr_t>(d) & 3) == 0)) {
        *reinterpret_cast<uchar4 *>(d) = make_uchar4(o[0], o[1], o[2], o[3]);
    } else {
        for (int i = 0; i < 4 && x0 + i < W; i++) d[i] = o[i];
    }
}

// Word-wise x-Sobel for 4-byte aligned planes with W % 4 == 0: a thread owns one 32-bit word (4 pixels) of a band
// of PFB rows.  Per row the horizontal differences d = r[x+1] - r[x-1] of its 4 pixels are formed from three
// aligned words (two funnel shifts) as biased u16x2 pairs; the vertical 1-2-1 slides through registers, so every
// source word is loaded once per band (+2 halo rows) instead of 18 byte loads per 4 pixels.
constexpr int PFB = 16;
__device__ __forceinline__ void sobel_row_diff(const uint8_t *row, int wi, int nw, uint32_t &lo, uint32_t &hi)
{
    const uint32_t *p = reinterpret_cast<const uint32_t *>(row) + wi;
    const uint32_t w0 = p[0], wl = wi > 0 ? p[-1] : 0u, wr = wi + 1 < nw ? p[1] : 0u;
    const uint32_t R = __funnelshift_r(w0, wr, 8);      // bytes x+1
    const uint32_t L = __funnelshift_l(wl, w0, 8);      // bytes x-1
    // per pixel R - L + 256 in [1, 511]: no borrow between the halves
    lo = (__byte_perm(R, 0u, 0x4140) | 0x01000100u) - __byte_perm(L, 0u, 0x4140);
    hi = (__byte_perm(R, 0u, 0x4342) | 0x01000100u) - __byte_perm(L, 0u, 0x4342);
}

__global__ void __launch_bounds__(128)
prefilter_xsobel4_kernel(PlaneU8 left, PlaneU8 right, PlaneU8W outL, PlaneU8W outR, int W, int H, int cap, BmStaged sg)
{
    const int img = blockIdx.z & 1, f = blockIdx.z >> 1;
    const uint8_t *src = (img ? right.p + (size_t)f * right.frame : left.p + (size_t)f * left.frame);
    const size_t sp = img ? right.pitch : left.pitch;
    uint8_t *dst = (img ? outR.p + (size_t)f * outR.frame : outL.p + (size_t)f * outL.frame);
    const size_t dp = img ? outR.pitch : outL.pitch;
    const int nw = W >> 2, wi = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = blockIdx.y * PFB, y1 = min(y0 + PFB, H);
    if (wi >= nw || y0 >= y1) return;
    const int paired = (H > 1) ? (H & ~1) : 0;
    const uint32_t capx4 = (uint32_t)cap * 0x01010101u;
    const uint32_t lob = (uint32_t)(1024 - cap) * 0x00010001u, hib = (uint32_t)(1024 + cap) * 0x00010001u;
    uint32_t pl = 0, ph = 0, cl = 0, ch = 0, nl, nh;
    if (y0 < paired) {
        sobel_row_diff(src + (size_t)(y0 > 0 ? y0 - 1 : 1) * sp, wi, nw, pl, ph);
        sobel_row_diff(src + (size_t)y0 * sp, wi, nw, cl, ch);
    }
    uint32_t *d = reinterpret_cast<uint32_t *>(dst + (size_t)y0 * dp) + wi;
    for (int y = y0; y < y1; y++) {
        uint32_t o = capx4;
        if (y < paired) {
            sobel_row_diff(src + (size_t)(y < H - 1 ? y + 1 : H - 2) * sp, wi, nw, nl, nh);
            // biased sum = sum + 1024; clip(sum, -cap, cap) + cap = clamp(biased, 1024 - cap, 1024 + cap) - (1024 - cap)
            const uint32_t sl = __vminu2(__vmaxu2(pl + 2u * cl + nl, lob), hib) - lob;
            const uint32_t sh = __vminu2(__vmaxu2(ph + 2u * ch + nh, lob), hib) - lob;
            o = __byte_perm(sl, sh, 0x6420);
            if (wi == 0) o = (o & 0xFFFFFF00u) | (uint32_t)cap;
            if (wi == nw - 1) o = (o & 0x00FFFFFFu) | ((uint32_t)cap << 24);
            pl = cl; ph = ch; cl = nl; ch = nh;
        }
        if (sg.LE) staged_store_word(sg, img, f, y, 4 * wi, o, wi == 0, wi == nw - 1);
        else *d = o;
        d = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(d) + dp);
    }
}

// Normalized response: val = ((4c + l + r + u + d) * sg - boxsum * ss) >> 10, clipped to +-cap.
// boxsum = ws x ws box with replicate-clamped coordinates (equivalent to OpenCV's sliding sums,
// which never wrap 16 bits for ws <= 255).
__global__ void __launch_bounds__(256)
prefilter_norm_kernel(PlaneU8 left, PlaneU8 right, PlaneU8W outL, PlaneU8W outR,
                      int W, int H, int ws, int cap, BmStaged stg)
{
    const int img = blockIdx.z & 1, f = blockIdx.z >> 1;
    const uint8_t *src = (img ? right.p + (size_t)f * right.frame : left.p + (size_t)f * left.frame);
    const size_t sp = img ? right.pitch : left.pitch;
    uint8_t *dst = (img ? outR.p + (size_t)f * outR.frame : outL.p + (size_t)f * outL.frame);
    const size_t dp = img ? outR.pitch : outL.pitch;
    const int y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const int h = ws / 2;
    int sg = ws * ws / 8;
    const int ss = (1024 + sg) / (sg * 2);
    sg *= ss;
    int sum = 0;
    for (int j = -h; j <= h; j++) {
        const uint8_t *r = src + (size_t)clampi(y + j, 0, H - 1) * sp;
        for (int i = -h; i <= h; i++) sum += r[clampi(x + i, 0, W - 1)];
    }
    const uint8_t *cur = src + (size_t)y * sp;
    const int c = cur[x], l = cur[max(x - 1, 0)], r = cur[min(x + 1, W - 1)];
    const int u = src[(size_t)max(y - 1, 0) * sp + x], d = src[(size_t)min(y + 1, H - 1) * sp + x];
    int val = ((4 * c + l + r + u + d) * sg - sum * ss) >> 10;
    const uint32_t o = (uint32_t)(clampi(val, -cap, cap) + cap);
    if (stg.LE) {
        staged_store_px(stg, img, f, y, x, W, o);
        return;
    }
    dst[(size_t)y * dp + x] = (uint8_t)o;
}

int launch_prefilter(int type, int winsize, int cap, int n, int W, int H,
                     PlaneU8 left, PlaneU8 right, PlaneU8W outL, PlaneU8W outR,
                     cudaStream_t st, int *launches, const BmStaged *staged)
{
    if (n <= 0) return 0;
    BmStaged sg = {nullptr, 0, 0, nullptr, 0, 0};
    if (staged) sg = *staged;
    if (type == RTDM_PREFILTER_XSOBEL) {
        const auto al4 = [](const void *p, size_t pitch, size_t frame) { return ((reinterpret_cast<uintptr_t>(p) | pitch | frame) & 3) == 0; };
        if (W % 4 == 0 && W >= 8 && al4(left.p, left.pitch, left.frame) && al4(right.p, right.pitch, right.frame) &&
            (staged || (al4(outL.p, outL.pitch, outL.frame) && al4(outR.p, outR.pitch, outR.frame)))) {
            dim3 grid(cdiv(W / 4, 128), cdiv(H, PFB), 2 * n);
            prefilter_xsobel4_kernel<<<grid, 128, 0, st>>>(left, right, outL, outR, W, H, cap, sg);
        } else {
            dim3 grid(cdiv(cdiv(W, 4), 256), H, 2 * n);
            prefilter_xsobel_kernel<<<grid, 256, 0, st>>>(left, right, outL, outR, W, H, cap, sg);
        }
    } else {
        dim3 grid(cdiv(W, 256), H, 2 * n);
        prefilter_norm_kernel<<<grid, 256, 0, st>>>(left, right, outL, outR, W, H, winsize, cap, sg);
    }
    if (launches) (*launches)++;
    RTDM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace rtdm
