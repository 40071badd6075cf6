/*
 * stereo_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic that rt-depth-map's software plugins
 * delegate to OpenCV:
 *   - SWMatcherKonolige::compute      /root/reference/stereo-matcher/bm-sw.cpp:33-38  -> cv::StereoBM::compute
 *   - SWSemiGlobalMatcher::compute    /root/reference/stereo-matcher/sgbm-sw.cpp:32-37 -> cv::StereoSGBM::compute
 *   - SWMorphologicalFilter::run      /root/reference/filter/mf-sw.cpp:19-28           -> cv::erode / cv::dilate
 *
 * The algorithm itself lives in the un-vendored, un-pinned third-party OpenCV
 * (calib3d: StereoBM, StereoSGBM, filterSpeckles, validateDisparity; imgproc:
 * erode, dilate, getStructuringElement, medianBlur).  It is restated here from its
 * published behaviour as written down in SURVEY.md Appendix A, and PINNED against
 * the in-image build `cv2` 4.13.0 (tests/golden/make_golden.py generates the
 * committed fixtures; tests/test_oracle_vs_cv2.py re-checks live when cv2 imports).
 * The reference repo itself holds no tests or golden vectors for this path
 * (SURVEY.md section 4), so the pin is cv2 4.13.0 and nothing else.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.  The product path (rt-depth-map_b200/csrc)
 * never links or calls it.
 *
 * Organisation differs from OpenCV on purpose (row-major sweeps with explicit
 * column sums instead of OpenCV's column-major sliding buffers); all sums are exact
 * integers so the order is irrelevant.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

typedef struct orc_params {
    int preFilterType;      /* 0 = NORMALIZED_RESPONSE, 1 = XSOBEL (StereoBM default) */
    int preFilterSize;      /* 9 by default (only used by NORMALIZED_RESPONSE)        */
    int preFilterCap;
    int blockSize;
    int minDisparity;
    int numDisparities;
    int textureThreshold;
    int uniquenessRatio;
    int speckleWindowSize;
    int speckleRange;
    int disp12MaxDiff;
    int mode;               /* SGBM: 0 = MODE_SGBM (5 paths), 1 = MODE_HH (8 paths)   */
    int P1, P2;
    int roi1[4];            /* x, y, w, h ; w == 0 || h == 0 means "empty" (whole image) */
    int roi2[4];
} orc_params;

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int mini(int a, int b) { return a < b ? a : b; }
static inline int maxi(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------- */
/* A.1  prefilters                                                           */
/* ------------------------------------------------------------------------- */

/* x-Sobel prefilter (StereoBM default, preFilterType = 1).  SURVEY App. A.1.
 * Rows are processed in pairs; an odd last row is filled with `cap`. */
void orc_prefilter_xsobel(const uint8_t *src, int sstep, uint8_t *dst, int dstep,
                          int W, int H, int cap)
{
    int paired = (H - 1 > 0) ? ((H) & ~1) : 0;   /* rows [0, paired) get real values */
    if (H - 1 <= 0) paired = 0;
    for (int y = 0; y < H; y++) {
        uint8_t *d = dst + (size_t)y * dstep;
        if (y >= paired) { for (int x = 0; x < W; x++) d[x] = (uint8_t)cap; continue; }
        int ya = (y > 0) ? y - 1 : (H > 1 ? 1 : 0);            /* reflect-101 */
        int yb = (y < H - 1) ? y + 1 : (H > 1 ? H - 2 : 0);
        const uint8_t *r0 = src + (size_t)ya * sstep;
        const uint8_t *r1 = src + (size_t)y * sstep;
        const uint8_t *r2 = src + (size_t)yb * sstep;
        d[0] = (uint8_t)cap;
        if (W > 1) d[W - 1] = (uint8_t)cap;
        for (int x = 1; x < W - 1; x++) {
            int v = (r0[x + 1] - r0[x - 1]) + 2 * (r1[x + 1] - r1[x - 1]) + (r2[x + 1] - r2[x - 1]);
            d[x] = (uint8_t)(clampi(v, -cap, cap) + cap);
        }
    }
}

/* Normalized-response prefilter (preFilterType = 0).  SURVEY App. A.1. */
void orc_prefilter_norm(const uint8_t *src, int sstep, uint8_t *dst, int dstep,
                        int W, int H, int winsize, int cap)
{
    int wsz2 = winsize / 2;
    int scale_g = winsize * winsize / 8, scale_s = (1024 + scale_g) / (scale_g * 2);
    scale_g *= scale_s;
    uint16_t *vsum = (uint16_t *)calloc((size_t)W + 2 * (wsz2 + 1) + 2, sizeof(uint16_t));
    uint16_t *vs = vsum + wsz2 + 1;           /* vs[-wsz2-1 .. W+wsz2] */
    /* initial vertical sums for "row -1": top row weighted wsz2+2, plus rows 1..wsz2-1 */
    for (int x = 0; x < W; x++) vs[x] = (uint16_t)(src[x] * (wsz2 + 2));
    for (int y = 1; y < wsz2; y++) {
        const uint8_t *r = src + (size_t)mini(y, H - 1) * sstep;
        for (int x = 0; x < W; x++) vs[x] = (uint16_t)(vs[x] + r[x]);
    }
    for (int y = 0; y < H; y++) {
        const uint8_t *top = src + (size_t)maxi(y - wsz2 - 1, 0) * sstep;
        const uint8_t *bot = src + (size_t)mini(y + wsz2, H - 1) * sstep;
        const uint8_t *prev = src + (size_t)maxi(y - 1, 0) * sstep;
        const uint8_t *curr = src + (size_t)y * sstep;
        const uint8_t *next = src + (size_t)mini(y + 1, H - 1) * sstep;
        uint8_t *d = dst + (size_t)y * dstep;
        for (int x = 0; x < W; x++) vs[x] = (uint16_t)(vs[x] + bot[x] - top[x]);
        for (int x = 0; x <= wsz2; x++) { vs[-x - 1] = vs[0]; vs[W + x] = vs[W - 1]; }
        int sum = vs[0] * (wsz2 + 1);
        for (int x = 1; x <= wsz2; x++) sum += vs[x];
        int val = ((curr[0] * 5 + curr[W > 1 ? 1 : 0] + prev[0] + next[0]) * scale_g - sum * scale_s) >> 10;
        d[0] = (uint8_t)(clampi(val, -cap, cap) + cap);
        int x;
        for (x = 1; x < W - 1; x++) {
            sum += vs[x + wsz2] - vs[x - wsz2 - 1];
            val = ((curr[x] * 4 + curr[x - 1] + curr[x + 1] + prev[x] + next[x]) * scale_g - sum * scale_s) >> 10;
            d[x] = (uint8_t)(clampi(val, -cap, cap) + cap);
        }
        if (W > 1) {
            sum += vs[x + wsz2] - vs[x - wsz2 - 1];
            val = ((curr[x] * 5 + curr[x - 1] + prev[x] + next[x]) * scale_g - sum * scale_s) >> 10;
            d[x] = (uint8_t)(clampi(val, -cap, cap) + cap);
        }
    }
    free(vsum);
}

/* ------------------------------------------------------------------------- */
/* getValidDisparityROI                                                      */
/* ------------------------------------------------------------------------- */
void orc_valid_roi(const int roi1[4], const int roi2[4], int W, int H,
                   int minD, int nd, int bs, int out[4])
{
    int r1[4] = {0, 0, W, H}, r2[4] = {0, 0, W, H};
    if (roi1[2] > 0 && roi1[3] > 0) memcpy(r1, roi1, sizeof r1);
    if (roi2[2] > 0 && roi2[3] > 0) memcpy(r2, roi2, sizeof r2);
    int h = bs / 2, maxD = minD + nd - 1;
    int xmin = maxi(r1[0], r2[0] + maxD) + h;
    int xmax = mini(r1[0] + r1[2], r2[0] + r2[2]) - h;
    int ymin = maxi(r1[1], r2[1]) + h;
    int ymax = mini(r1[1] + r1[3], r2[1] + r2[3]) - h;
    if (xmax - xmin > 0 && ymax - ymin > 0) {
        out[0] = xmin; out[1] = ymin; out[2] = xmax - xmin; out[3] = ymax - ymin;
    } else {
        out[0] = out[1] = out[2] = out[3] = 0;
    }
}

/* ------------------------------------------------------------------------- */
/* A.2  block-matching core: SAD volume + WTA + texture + uniqueness + subpixel */
/* ------------------------------------------------------------------------- */
/* Lp / Rp are the PREFILTERED images (step `step`).  Rows [row0,row1) are computed
 * (disp and cost rows outside are untouched).  Window rows are clamped to the image
 * (never hit inside the valid rect).  disp/cost are full-width rows. */
void orc_bm_core(const uint8_t *Lp, const uint8_t *Rp, int step, int W, int H,
                 int row0, int row1, int cap, int bs, int minD, int nd,
                 int texThr, int uniq,
                 int16_t *disp, int dstep, int16_t *cost, int cstep)
{
    const int h = bs / 2;
    const int lofs = maxi(nd - 1 + minD, 0), rofs = -mini(nd - 1 + minD, 0);
    const int W1 = W - rofs - nd + 1;
    const int16_t FILT = (int16_t)((minD - 1) * 16);
    if (lofs >= W || rofs >= W || W1 < 1) {
        for (int y = row0; y < row1; y++)
            for (int x = 0; x < W; x++) disp[(size_t)y * dstep + x] = FILT;
        return;
    }
    const int NV = W1 + 2 * h;                       /* virtual columns xc = xv - h */
    int *lcol = (int *)malloc(sizeof(int) * NV), *rbase = (int *)malloc(sizeof(int) * NV);
    for (int xv = 0; xv < NV; xv++) {
        int xc = xv - h;
        lcol[xv] = clampi(xc, -lofs, W - lofs - 1) + lofs;
        rbase[xv] = clampi(xc, -rofs, W - rofs - nd) + rofs;
    }
    int32_t *V = (int32_t *)calloc((size_t)NV * nd, sizeof(int32_t));   /* vertical window sums of AD */
    int32_t *T = (int32_t *)calloc((size_t)NV, sizeof(int32_t));        /* vertical window sums of |L-cap| */
    int32_t *sad = (int32_t *)malloc(sizeof(int32_t) * (nd + 2));
    int32_t *S = sad + 1;

    for (int y = row0; y < row1; y++) {
        /* (re)build or slide the vertical sums so that they cover rows y-h .. y+h */
        int first = (y == row0);
        for (int k = first ? -h : h; k <= h; k++) {
            int ra = clampi(y + k, 0, H - 1);
            const uint8_t *l = Lp + (size_t)ra * step, *r = Rp + (size_t)ra * step;
            const uint8_t *lo = NULL, *ro = NULL;
            if (!first) {
                int rs = clampi(y - h - 1, 0, H - 1);
                lo = Lp + (size_t)rs * step; ro = Rp + (size_t)rs * step;
            }
            for (int xv = 0; xv < NV; xv++) {
                int lv = l[lcol[xv]];
                const uint8_t *rp = r + rbase[xv];
                int32_t *v = V + (size_t)xv * nd;
                if (first) {
                    for (int d = 0; d < nd; d++) v[d] += abs(lv - rp[d]);
                    T[xv] += abs(lv - cap);
                } else {
                    int lvo = lo[lcol[xv]];
                    const uint8_t *rpo = ro + rbase[xv];
                    for (int d = 0; d < nd; d++) v[d] += abs(lv - rp[d]) - abs(lvo - rpo[d]);
                    T[xv] += abs(lv - cap) - abs(lvo - cap);
                }
            }
        }
        int16_t *drow = disp + (size_t)y * dstep;
        int16_t *crow = cost ? cost + (size_t)y * cstep : NULL;
        for (int x = 0; x < lofs; x++) drow[x] = FILT;
        for (int x = lofs + W1; x < W; x++) drow[x] = FILT;
        /* horizontal window */
        int tsum = 0;
        for (int d = 0; d < nd; d++) S[d] = 0;
        for (int xv = 0; xv < 2 * h; xv++) {
            const int32_t *v = V + (size_t)xv * nd;
            for (int d = 0; d < nd; d++) S[d] += v[d];
            tsum += T[xv];
        }
        for (int x = 0; x < W1; x++) {
            const int32_t *vin = V + (size_t)(x + 2 * h) * nd;
            for (int d = 0; d < nd; d++) S[d] += vin[d];
            tsum += T[x + 2 * h];
            /* WTA: first minimum in d order (d = nd-1 is true disparity minD) */
            int minsad = INT_MAX, mind = -1;
            for (int d = 0; d < nd; d++) if (S[d] < minsad) { minsad = S[d]; mind = d; }
            int16_t out;
            int ok = 1;
            if (tsum < texThr) ok = 0;
            if (ok && uniq > 0) {
                int thresh = minsad + (minsad * uniq / 100);
                for (int d = 0; d < nd; d++)
                    if ((d < mind - 1 || d > mind + 1) && S[d] <= thresh) { ok = 0; break; }
            }
            if (ok) {
                S[-1] = S[1]; S[nd] = S[nd - 2];
                int p = S[mind + 1], n = S[mind - 1];
                int q = p + n - 2 * S[mind] + abs(p - n);
                int v = (nd - mind - 1 + minD) * 256 + (q != 0 ? ((p - n) * 256) / q : 0) + 15;
                out = (int16_t)(v >> 4);
                if (crow) crow[lofs + x] = (int16_t)S[mind];
            } else {
                out = FILT;
            }
            drow[lofs + x] = out;
            const int32_t *vout = V + (size_t)x * nd;
            for (int d = 0; d < nd; d++) S[d] -= vout[d];
            tsum -= T[x];
        }
    }
    free(sad); free(T); free(V); free(lcol); free(rbase);
}

/* Raw SAD volume for a single row (debug / stage tests): sad[x*nd + d], x in [0,W1). */
void orc_bm_sad_row(const uint8_t *Lp, const uint8_t *Rp, int step, int W, int H,
                    int y, int bs, int minD, int nd, int32_t *sadout)
{
    const int h = bs / 2;
    const int lofs = maxi(nd - 1 + minD, 0), rofs = -mini(nd - 1 + minD, 0);
    const int W1 = W - rofs - nd + 1;
    for (int x = 0; x < W1; x++)
        for (int d = 0; d < nd; d++) {
            int s = 0;
            for (int j = -h; j <= h; j++) {
                int r = clampi(y + j, 0, H - 1);
                for (int i = -h; i <= h; i++) {
                    int lc = clampi(x + i, -lofs, W - lofs - 1) + lofs;
                    int rb = clampi(x + i, -rofs, W - rofs - nd) + rofs;
                    s += abs((int)Lp[(size_t)r * step + lc] - (int)Rp[(size_t)r * step + rb + d]);
                }
            }
            sadout[(size_t)x * nd + d] = s;
        }
}

/* ------------------------------------------------------------------------- */
/* A.3  validateDisparity (left-right consistency from the left cost only)    */
/* ------------------------------------------------------------------------- */
void orc_validate_disparity(int16_t *disp, int dstep, const int16_t *cost, int cstep,
                            int W, int rows, int minD, int nd, int d12)
{
    const int maxD = minD + nd;
    const int minX1 = maxi(maxD, 0), maxX1 = W + mini(minD, 0);
    const int INV = (minD - 1) * 16;
    int *disp2 = (int *)malloc(sizeof(int) * 2 * (size_t)W), *cost2 = disp2 + W;
    d12 *= 16;
    for (int y = 0; y < rows; y++) {
        int16_t *dp = disp + (size_t)y * dstep;
        const int16_t *cp = cost + (size_t)y * cstep;
        for (int x = 0; x < W; x++) { disp2[x] = INV; cost2[x] = INT_MAX; }
        for (int x = minX1; x < maxX1; x++) {
            int d = dp[x], c = cp[x];
            if (d == INV) continue;
            int x2 = x - ((d + 8) >> 4);
            if (x2 < 0 || x2 >= W) continue;          /* cannot happen for valid d */
            if (cost2[x2] > c) { cost2[x2] = c; disp2[x2] = d; }
        }
        for (int x = minX1; x < maxX1; x++) {
            int d = dp[x];
            if (d == INV) continue;
            int d0 = d >> 4, d1 = (d + 15) >> 4;
            int x0 = x - d0, x1 = x - d1;
            if ((0 <= x0 && x0 < W && disp2[x0] > INV && abs(disp2[x0] - d) > d12) &&
                (0 <= x1 && x1 < W && disp2[x1] > INV && abs(disp2[x1] - d) > d12))
                dp[x] = (int16_t)INV;
        }
    }
    free(disp2);
}

/* ------------------------------------------------------------------------- */
/* A.4  filterSpeckles == remove small 4-connected components                 */
/* ------------------------------------------------------------------------- */
void orc_filter_speckles(int16_t *img, int step, int W, int H,
                         int newVal, int maxSize, int maxDiff)
{
    size_t N = (size_t)W * H;
    int32_t *label = (int32_t *)calloc(N, sizeof(int32_t));
    int32_t *stack = (int32_t *)malloc(N * sizeof(int32_t));
    uint8_t *small = (uint8_t *)malloc(N + 1);      /* per label: is it a speckle */
    int cur = 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            size_t idx = (size_t)y * W + x;
            int16_t v = img[(size_t)y * step + x];
            if (v == newVal) continue;
            if (label[idx]) { if (small[label[idx]]) img[(size_t)y * step + x] = (int16_t)newVal; continue; }
            /* flood */
            cur++;
            int sp = 0, count = 0;
            stack[sp++] = (int32_t)idx; label[idx] = cur;
            while (sp) {
                int32_t p = stack[--sp];
                count++;
                int py = p / W, px = p - py * W;
                int pv = img[(size_t)py * step + px];
                static const int dx[4] = {1, -1, 0, 0}, dy[4] = {0, 0, 1, -1};
                for (int k = 0; k < 4; k++) {
                    int qx = px + dx[k], qy = py + dy[k];
                    if (qx < 0 || qx >= W || qy < 0 || qy >= H) continue;
                    size_t q = (size_t)qy * W + qx;
                    int qv = img[(size_t)qy * step + qx];
                    if (label[q] || qv == newVal || abs(pv - qv) > maxDiff) continue;
                    label[q] = cur; stack[sp++] = (int32_t)q;
                }
            }
            small[cur] = (uint8_t)(count <= maxSize);
            if (small[cur]) img[(size_t)y * step + x] = (int16_t)newVal;
        }
    free(small); free(stack); free(label);
}

/* ------------------------------------------------------------------------- */
/* cv::StereoBM::compute restated                                             */
/* ------------------------------------------------------------------------- */
/* returns 0 ok, -22 (EINVAL) on parameter errors OpenCV would assert on. */
int orc_bm_check_params(const orc_params *p)
{
    if (p->preFilterType != 0 && p->preFilterType != 1) return -22;
    if (p->preFilterSize < 5 || p->preFilterSize > 255 || p->preFilterSize % 2 == 0) return -22;
    if (p->preFilterCap < 1 || p->preFilterCap > 63) return -22;
    if (p->blockSize < 5 || p->blockSize > 255 || p->blockSize % 2 == 0) return -22;
    if (p->numDisparities <= 0 || p->numDisparities % 16 != 0) return -22;
    if (p->textureThreshold < 0) return -22;
    if (p->uniquenessRatio < 0) return -22;
    return 0;
}

int orc_bm_compute(const uint8_t *left, int lstep, const uint8_t *right, int rstep,
                   int W, int H, const orc_params *p, int16_t *disp, int dstep)
{
    int rc = orc_bm_check_params(p);
    if (rc) return rc;
    if (p->blockSize >= W || p->blockSize >= H) return -22;
    const int nd = p->numDisparities, minD = p->minDisparity, bs = p->blockSize;
    const int16_t FILT = (int16_t)((minD - 1) * 16);
    const int lofs = maxi(nd - 1 + minD, 0), rofs = -mini(nd - 1 + minD, 0);
    const int W1 = W - rofs - nd + 1;
    if (lofs >= W || rofs >= W || W1 < 1) {
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) disp[(size_t)y * dstep + x] = FILT;
        return 0;
    }
    uint8_t *Lp = (uint8_t *)malloc((size_t)W * H), *Rp = (uint8_t *)malloc((size_t)W * H);
    if (p->preFilterType == 1) {
        orc_prefilter_xsobel(left, lstep, Lp, W, W, H, p->preFilterCap);
        orc_prefilter_xsobel(right, rstep, Rp, W, W, H, p->preFilterCap);
    } else {
        orc_prefilter_norm(left, lstep, Lp, W, W, H, p->preFilterSize, p->preFilterCap);
        orc_prefilter_norm(right, rstep, Rp, W, W, H, p->preFilterSize, p->preFilterCap);
    }
    int vr[4];
    orc_valid_roi(p->roi1, p->roi2, W, H, minD, nd, bs, vr);
    /* intersect with image rows */
    int row0 = clampi(vr[1], 0, H), row1 = clampi(vr[1] + vr[3], 0, H);
    if (vr[2] == 0 || vr[3] == 0) row0 = row1 = 0;
    for (int y = 0; y < H; y++) {
        if (y >= row0 && y < row1) continue;
        for (int x = 0; x < W; x++) disp[(size_t)y * dstep + x] = FILT;
    }
    if (row1 > row0) {
        int16_t *cost = (int16_t *)calloc((size_t)W * H, sizeof(int16_t));
        orc_bm_core(Lp, Rp, W, W, H, row0, row1, p->preFilterCap, bs, minD, nd,
                    p->textureThreshold, p->uniquenessRatio, disp, dstep, cost, W);
        if (p->disp12MaxDiff >= 0)
            orc_validate_disparity(disp + (size_t)row0 * dstep, dstep, cost + (size_t)row0 * W, W,
                                   W, row1 - row0, minD, nd, p->disp12MaxDiff);
        for (int y = row0; y < row1; y++) {
            int16_t *d = disp + (size_t)y * dstep;
            for (int x = 0; x < mini(maxi(vr[0], 0), W); x++) d[x] = FILT;
            for (int x = maxi(vr[0] + vr[2], 0); x < W; x++) d[x] = FILT;
        }
        free(cost);
    }
    if (p->speckleRange >= 0 && p->speckleWindowSize > 0)
        orc_filter_speckles(disp, dstep, W, H, FILT, p->speckleWindowSize, p->speckleRange);
    free(Lp); free(Rp);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* A.5  morphology: erode / dilate with MORPH_ELLIPSE, open+close sequence   */
/* ------------------------------------------------------------------------- */
/* cv::getStructuringElement(MORPH_ELLIPSE, Size(kw,kh)) row extents [j1,j2). */
void orc_ellipse_rows(int kw, int kh, int *j1, int *j2)
{
    int r = kh / 2, c = kw / 2;
    double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < kh; i++) {
        int dy = i - r;
        if (abs(dy) <= r) {
            /* cvRound(c*sqrt((r*r - dy*dy)*inv_r2)) */
            double v = c * __builtin_sqrt((r * r - dy * dy) * inv_r2);
            int dx = (int)__builtin_nearbyint(v);
            j1[i] = maxi(c - dx, 0);
            j2[i] = mini(c + dx + 1, kw);
        } else { j1[i] = 0; j2[i] = 0; }
    }
}

/* op = 0 erode (min, border +inf), op = 1 dilate (max, border -inf); anchor = (kw/2,kh/2) */
void orc_morph(const uint8_t *src, int sstep, uint8_t *dst, int dstep, int W, int H,
               int kw, int kh, int op)
{
    int *j1 = (int *)malloc(sizeof(int) * 2 * kh), *j2 = j1 + kh;
    orc_ellipse_rows(kw, kh, j1, j2);
    int ax = kw / 2, ay = kh / 2;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int acc = op ? 0 : 255;
            for (int i = 0; i < kh; i++) {
                int yy = y + i - ay;
                if (yy < 0 || yy >= H) continue;           /* border never wins */
                const uint8_t *s = src + (size_t)yy * sstep;
                for (int j = j1[i]; j < j2[i]; j++) {
                    int xx = x + j - ax;
                    if (xx < 0 || xx >= W) continue;
                    int v = s[xx];
                    acc = op ? maxi(acc, v) : mini(acc, v);
                }
            }
            dst[(size_t)y * dstep + x] = (uint8_t)acc;
        }
    free(j1);
}

/* SWMorphologicalFilter::run (mf-sw.cpp:22-27): erode, dilate, dilate, erode, 10x10 ellipse */
void orc_morph_open_close(const uint8_t *src, int sstep, uint8_t *dst, int dstep,
                          int W, int H, int kw, int kh)
{
    uint8_t *t0 = (uint8_t *)malloc((size_t)W * H), *t1 = (uint8_t *)malloc((size_t)W * H);
    orc_morph(src, sstep, t0, W, W, H, kw, kh, 0);
    orc_morph(t0, W, t1, W, W, H, kw, kh, 1);
    orc_morph(t1, W, t0, W, W, H, kw, kh, 1);
    orc_morph(t0, W, dst, dstep, W, H, kw, kh, 0);
    free(t0); free(t1);
}

/* 3x3 median on int16, replicate border (cv::medianBlur(disp, 3) as used by StereoSGBM) */
static inline void cswap(int *a, int *b) { if (*a > *b) { int t = *a; *a = *b; *b = t; } }
void orc_median3_s16(const int16_t *src, int sstep, int16_t *dst, int dstep, int W, int H)
{
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int v[9], k = 0;
            for (int j = -1; j <= 1; j++)
                for (int i = -1; i <= 1; i++)
                    v[k++] = src[(size_t)clampi(y + j, 0, H - 1) * sstep + clampi(x + i, 0, W - 1)];
            for (int a = 0; a < 9; a++) for (int b = a + 1; b < 9; b++) cswap(&v[a], &v[b]);
            dst[(size_t)y * dstep + x] = (int16_t)v[4];
        }
}
