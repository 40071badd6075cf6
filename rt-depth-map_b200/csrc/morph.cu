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
constexpr int TOW = 64, TOH = 16;          // output tile
constexpr int SLACK = 16;                  // columns of identity padding for the doubling reads

__global__ void __launch_bounds__(256)
morph_kernel(int W, int H, PlaneU8 src, PlaneU8W dst, MorphSE se, int op)
{
    extern __shared__ uint8_t ms[];
    const int TR = TOH + se.kh - 1, TC = TOW + se.kw - 1;
    const int TCP = (TC + SLACK + 3) & ~3;
    const size_t LV = (size_t)TR * TCP;
    uint8_t *m1 = ms, *m2 = m1 + LV, *m4 = m2 + LV, *m8 = m4 + LV, *m16 = m8 + LV;
    const int f = blockIdx.z;
    const int ox = blockIdx.x * TOW, oy = blockIdx.y * TOH;
    const uint8_t *s = src.p + (size_t)f * src.frame;
    const uint8_t flip = op ? 0xFF : 0x00;       // dilate: work on complemented values
    // level 0: tile with identity (255 after flip) outside the image
    for (int i = threadIdx.x; i < TR * TCP; i += blockDim.x) {
        int r = i / TCP, c = i - r * TCP;
        int gy = oy + r - se.ay, gx = ox + c - se.ax;
        uint8_t v = 0xFF;
        if (c < TC && gy >= 0 && gy < H && gx >= 0 && gx < W) v = s[(size_t)gy * src.pitch + gx] ^ flip;
        m1[i] = v;
    }
    __syncthreads();
    // doubling levels (reads beyond the row end land in the slack / next row start: guard with c)
    for (int i = threadIdx.x; i < TR * TCP; i += blockDim.x) {
        int c = i % TCP;
        m2[i] = (c + 1 < TCP) ? min(m1[i], m1[i + 1]) : m1[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TR * TCP; i += blockDim.x) {
        int c = i % TCP;
        m4[i] = (c + 2 < TCP) ? min(m2[i], m2[i + 2]) : m2[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TR * TCP; i += blockDim.x) {
        int c = i % TCP;
        m8[i] = (c + 4 < TCP) ? min(m4[i], m4[i + 4]) : m4[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TR * TCP; i += blockDim.x) {
        int c = i % TCP;
        m16[i] = (c + 8 < TCP) ? min(m8[i], m8[i + 8]) : m8[i];
    }
    __syncthreads();
    uint8_t *d = dst.p + (size_t)f * dst.frame;
    for (int i = threadIdx.x; i < TOW * TOH; i += blockDim.x) {
        int ly = i / TOW, lx = i - ly * TOW;
        int gx = ox + lx, gy = oy + ly;
        if (gx >= W || gy >= H) continue;
        int acc = 255;
        for (int k = 0; k < se.kh; k++) {
            int L = se.j2[k] - se.j1[k];
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
        d[(size_t)gy * dst.pitch + gx] = (uint8_t)acc ^ flip;
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
