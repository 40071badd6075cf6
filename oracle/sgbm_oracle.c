/*
 * sgbm_oracle.c -- CPU ORACLE for semi-global matching (test infrastructure, NOT product code).
 *
 * Restates what SWSemiGlobalMatcher::compute (/root/reference/stereo-matcher/sgbm-sw.cpp:32-37)
 * obtains from the un-vendored cv::StereoSGBM::compute: Birchfield-Tomasi pixel cost on an
 * x-Sobel plane and the raw plane, box aggregation, 5-path (MODE_SGBM) or 8-path (MODE_HH)
 * min-plus path aggregation, winner-take-all with uniqueness / sub-pixel / left-right check,
 * then medianBlur(3) and filterSpeckles.  Algorithm per SURVEY.md App. A.6; pinned against
 * cv2 4.13.0 by tests/golden (sgbm_*.npz) and tests/test_oracle_golden.py.
 *
 * Organised by whole volumes (pixel cost -> C -> one sweep per path direction -> S) instead of
 * OpenCV's rolling row buffers; every L_r is a deterministic function of C along its path and all
 * L_r >= 0, so the order in which the saturating sum S is accumulated is irrelevant.
 * Domain: max S < 32767; the function returns 1 (and still produces an
 * output) when that is violated so tests can assert they stay inside the domain.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

typedef struct orc_params {
    int preFilterType, preFilterSize, preFilterCap, blockSize, minDisparity, numDisparities;
    int textureThreshold, uniquenessRatio, speckleWindowSize, speckleRange, disp12MaxDiff;
    int mode, P1, P2;
    int roi1[4], roi2[4];
} orc_params;

void orc_median3_s16(const int16_t *src, int sstep, int16_t *dst, int dstep, int W, int H);
void orc_filter_speckles(int16_t *img, int step, int W, int H, int newVal, int maxSize, int maxDiff);

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int mini(int a, int b) { return a < b ? a : b; }
static inline int maxi(int a, int b) { return a > b ? a : b; }
static inline int16_t sat16(int v) { return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }

/* planes of one image row: plane 0 = clipped x-Sobel, plane 1 = raw; columns 0 and W-1 = ftzero */
static void bt_planes(const uint8_t *img, int step, int W, int H, int y, int ftzero, uint8_t *p0, uint8_t *p1)
{
    const uint8_t *r = img + (size_t)y * step;
    const uint8_t *rn = y > 0 ? r - step : r, *rs = y < H - 1 ? r + step : r;
    for (int x = 0; x < W; x++) { p0[x] = (uint8_t)ftzero; p1[x] = (uint8_t)ftzero; }
    for (int x = 1; x < W - 1; x++) {
        int g = 2 * (r[x + 1] - r[x - 1]) + (rn[x + 1] - rn[x - 1]) + (rs[x + 1] - rs[x - 1]);
        p0[x] = (uint8_t)(clampi(g, -ftzero, ftzero) + ftzero);
        p1[x] = r[x];
    }
}

/* half-pixel interval [lo, hi] of a plane at x */
static inline void bt_interval(const uint8_t *p, int W, int x, int *lo, int *hi)
{
    int v = p[x];
    int a = x > 0 ? (v + p[x - 1]) / 2 : v;
    int b = x < W - 1 ? (v + p[x + 1]) / 2 : v;
    *lo = mini(mini(a, b), v);
    *hi = maxi(maxi(a, b), v);
}

/* Birchfield-Tomasi cost of one row: pix[x1*D + d], x1 in [0,W1), d in [0,D) */
void orc_sgbm_pixel_cost_row(const uint8_t *left, int lstep, const uint8_t *right, int rstep,
                             int W, int H, int y, int minD, int D, int ftzero, int16_t *pix)
{
    const int maxD = minD + D;
    const int minX1 = maxi(maxD, 0), maxX1 = W + mini(minD, 0), W1 = maxX1 - minX1;
    uint8_t *buf = (uint8_t *)malloc((size_t)W * 4);
    uint8_t *l0 = buf, *l1 = buf + W, *r0 = buf + 2 * W, *r1 = buf + 3 * W;
    int *iv = (int *)malloc(sizeof(int) * 4 * (size_t)W);      /* v0, v1 per plane for the right image */
    bt_planes(left, lstep, W, H, y, ftzero, l0, l1);
    bt_planes(right, rstep, W, H, y, ftzero, r0, r1);
    memset(pix, 0, sizeof(int16_t) * (size_t)W1 * D);
    for (int pl = 0; pl < 2; pl++) {
        const uint8_t *u_ = pl ? l1 : l0, *v_ = pl ? r1 : r0;
        int *v0 = iv + 2 * W * pl, *v1 = v0 + W;
        const int shift = pl ? 2 : 0;
        for (int x = 0; x < W; x++) bt_interval(v_, W, x, &v0[x], &v1[x]);
        for (int x = minX1; x < maxX1; x++) {
            int u = u_[x], u0, u1;
            bt_interval(u_, W, x, &u0, &u1);
            int16_t *c = pix + (size_t)(x - minX1) * D;
            for (int d = minD; d < maxD; d++) {
                int xr = x - d;
                if (xr < 0 || xr >= W) continue;           /* cannot happen inside [minX1,maxX1) */
                int v = v_[xr];
                int c0 = maxi(maxi(0, u - v1[xr]), v0[xr] - u);
                int c1 = maxi(maxi(0, v - u1), u0 - v);
                c[d - minD] = (int16_t)(c[d - minD] + (mini(c0, c1) >> shift));
            }
        }
    }
    free(iv); free(buf);
}

/* L_r for one pixel given the predecessor's L vector (with slots [-1] and [D] available for the
 * sentinels) and its minimum; accumulates into S (saturating).  Returns min_d L_r. */
static inline int path_step(const int16_t *Cp, int16_t *Sp, int16_t *Lp, int minLp, int16_t *Lc,
                            int D, int P1, int P2)
{
    Lp[-1] = 32767; Lp[D] = 32767;          /* set on READ, like OpenCV (zero border vectors too) */
    const int delta = P2 + minLp;
    int m = INT_MAX;
    for (int d = 0; d < D; d++) {
        int L = Cp[d] + mini(mini((int)Lp[d], Lp[d - 1] + P1), mini(Lp[d + 1] + P1, delta)) - delta;
        Lc[d] = (int16_t)L;
        if (L < m) m = L;
        Sp[d] = sat16(Sp[d] + L);
    }
    return m;
}

/* one aggregation path; (px,py) = offset of the predecessor pixel.  S += L_r (saturating).
 * Predecessors outside the image read as L = 0, min L = 0. */
static void aggregate_path(const int16_t *C, int16_t *S, int W1, int H, int D, int P1, int P2, int px, int py)
{
    const int DP = D + 2;                    /* slot layout: [sentinel, L[0..D), sentinel] */
    if (py == 0) {
        /* horizontal: rows are independent chains */
#pragma omp parallel
        {
            int16_t *row = (int16_t *)malloc((size_t)(W1 + 2) * DP * sizeof(int16_t));
            int *mn = (int *)malloc((size_t)(W1 + 2) * sizeof(int));
#pragma omp for schedule(static)
            for (int y = 0; y < H; y++) {
                memset(row, 0, (size_t)(W1 + 2) * DP * sizeof(int16_t));
                memset(mn, 0, (size_t)(W1 + 2) * sizeof(int));
                int x0 = px < 0 ? 0 : W1 - 1, xstep = px < 0 ? 1 : -1;
                for (int xi = 0, x = x0; xi < W1; xi++, x += xstep) {
                    int xp = x + px;
                    mn[x + 1] = path_step(C + ((size_t)y * W1 + x) * D, S + ((size_t)y * W1 + x) * D,
                                          row + (size_t)(xp + 1) * DP + 1, mn[xp + 1],
                                          row + (size_t)(x + 1) * DP + 1, D, P1, P2);
                }
            }
            free(row); free(mn);
        }
        return;
    }
    /* vertical / diagonal: rows are serial, the pixels of a row are independent */
    int16_t *rowA = (int16_t *)calloc((size_t)(W1 + 2) * DP, sizeof(int16_t));
    int16_t *rowB = (int16_t *)calloc((size_t)(W1 + 2) * DP, sizeof(int16_t));
    int *minA = (int *)calloc((size_t)(W1 + 2), sizeof(int)), *minB = (int *)calloc((size_t)(W1 + 2), sizeof(int));
    int y0 = py < 0 ? 0 : H - 1, ystep = py < 0 ? 1 : -1;
    for (int yi = 0, y = y0; yi < H; yi++, y += ystep) {
        /* the sentinel writes touch only slots [-1] and [D] of the PREVIOUS row: do them up front so the
         * parallel loop below only reads rowA */
        for (int x = 0; x < W1 + 2; x++) { rowA[(size_t)x * DP] = 32767; rowA[(size_t)x * DP + D + 1] = 32767; }
#pragma omp parallel for schedule(static)
        for (int x = 0; x < W1; x++) {
            int xp = x + px;
            const int16_t *Lp = rowA + (size_t)(xp + 1) * DP + 1;
            const int delta = P2 + minA[xp + 1];
            const int16_t *Cp = C + ((size_t)y * W1 + x) * D;
            int16_t *Sp = S + ((size_t)y * W1 + x) * D;
            int16_t *Lc = rowB + (size_t)(x + 1) * DP + 1;
            int m = INT_MAX;
            for (int d = 0; d < D; d++) {
                int L = Cp[d] + mini(mini((int)Lp[d], Lp[d - 1] + P1), mini(Lp[d + 1] + P1, delta)) - delta;
                Lc[d] = (int16_t)L;
                if (L < m) m = L;
                Sp[d] = sat16(Sp[d] + L);
            }
            minB[x + 1] = m;
        }
        int16_t *t = rowA; rowA = rowB; rowB = t;
        int *tm = minA; minA = minB; minB = tm;
        /* border slots (x = -1, W1) of both buffers are never written and stay zero */
    }
    free(rowA); free(rowB); free(minA); free(minB);
}

/* returns 0 ok, 1 ok-but-outside-domain (S saturated), negative errno on bad parameters */
int orc_sgbm_compute(const uint8_t *left, int lstep, const uint8_t *right, int rstep,
                     int W, int H, const orc_params *p, int16_t *disp, int dstep)
{
    const int D = p->numDisparities, minD = p->minDisparity, maxD = minD + D;
    if (D <= 0 || D % 16 != 0) return -22;
    if (p->blockSize < 1 || p->blockSize % 2 == 0) return -22;
    const int bs = p->blockSize > 0 ? p->blockSize : 5;
    const int h = bs / 2;
    const int uniq = p->uniquenessRatio >= 0 ? p->uniquenessRatio : 10;
    const int d12 = p->disp12MaxDiff > 0 ? p->disp12MaxDiff : 1;
    const int P1 = p->P1 > 0 ? p->P1 : 2, P2 = maxi(p->P2 > 0 ? p->P2 : 5, P1 + 1);
    const int ftzero = maxi(p->preFilterCap, 15) | 1;
    const int minX1 = maxi(maxD, 0), maxX1 = W + mini(minD, 0), W1 = maxX1 - minX1;
    const int INV = minD - 1, INVS = INV * 16;
    int outside = 0;
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) disp[(size_t)y * dstep + x] = (int16_t)INVS;
    if (W1 > 0) {
        const size_t rowsz = (size_t)W1 * D, vol = rowsz * H;
        int16_t *Hs = (int16_t *)malloc(vol * sizeof(int16_t));
        int16_t *C = (int16_t *)malloc(vol * sizeof(int16_t));
        int16_t *S = (int16_t *)calloc(vol, sizeof(int16_t));
        /* pixel cost + horizontal window (cost columns clamped to [0, W1-1]) */
#pragma omp parallel
        {
            int16_t *pix = (int16_t *)malloc(rowsz * sizeof(int16_t));
#pragma omp for schedule(dynamic, 4)
            for (int y = 0; y < H; y++) {
                orc_sgbm_pixel_cost_row(left, lstep, right, rstep, W, H, y, minD, D, ftzero, pix);
                int16_t *hs = Hs + (size_t)y * rowsz;
                for (int x = 0; x < W1; x++)
                    for (int d = 0; d < D; d++) {
                        int s = 0;
                        for (int i = -h; i <= h; i++) s += pix[(size_t)clampi(x + i, 0, W1 - 1) * D + d];
                        hs[(size_t)x * D + d] = (int16_t)s;
                    }
            }
            free(pix);
        }
        /* vertical window (rows clamped to [0, H-1]) + P2 */
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; y++) {
            int16_t *c = C + (size_t)y * rowsz;
            for (size_t i = 0; i < rowsz; i++) {
                int s = P2;
                for (int j = -h; j <= h; j++) s += Hs[(size_t)clampi(y + j, 0, H - 1) * rowsz + i];
                c[i] = (int16_t)s;
            }
        }
        free(Hs);
        /* path aggregation: pass 1 (always) */
        static const int dirs1[4][2] = {{-1, 0}, {-1, -1}, {0, -1}, {1, -1}};
        static const int dirs2[4][2] = {{1, 0}, {-1, 1}, {0, 1}, {1, 1}};
        const int ndir2 = p->mode == 1 ? 4 : 1;        /* MODE_HH: 4 more; MODE_SGBM: the in-row R->L path */
        for (int k = 0; k < 4; k++) aggregate_path(C, S, W1, H, D, P1, P2, dirs1[k][0], dirs1[k][1]);
        for (int k = 0; k < ndir2; k++) aggregate_path(C, S, W1, H, D, P1, P2, dirs2[k][0], dirs2[k][1]);
        free(C);
        /* winner-take-all, uniqueness, sub-pixel, left-right check: per row */
#pragma omp parallel
        {
            int *disp2 = (int *)malloc(sizeof(int) * 2 * (size_t)W), *cost2 = disp2 + W;
#pragma omp for schedule(static) reduction(| : outside)
            for (int y = 0; y < H; y++) {
                int16_t *drow = disp + (size_t)y * dstep;
                /* cv2 initialises disp2 with the x16-SCALED invalid value (minD - 1) * 16, which for minD >= 2 passes the
                 * `disp2 >= minD` test of the LR check below: unset entries then count as a (mismatching) disparity */
                for (int x = 0; x < W; x++) { disp2[x] = INVS; cost2[x] = 32767; }
                for (int x = W1 - 1; x >= 0; x--) {
                    const int16_t *Sp = S + ((size_t)y * W1 + x) * D;
                    int minS = 32767, best = -1;
                    for (int d = 0; d < D; d++) {
                        if (Sp[d] >= 32767) outside = 1;
                        if (Sp[d] < minS) { minS = Sp[d]; best = d; }
                    }
                    if (best < 0) continue;
                    int d;
                    for (d = 0; d < D; d++)
                        if (Sp[d] * (100 - uniq) < minS * 100 && abs(best - d) > 1) break;
                    if (d < D) continue;
                    d = best;
                    int x2 = x + minX1 - d - minD;
                    if (x2 >= 0 && x2 < W && cost2[x2] > minS) { cost2[x2] = minS; disp2[x2] = d + minD; }
                    if (0 < d && d < D - 1) {
                        int den = maxi(Sp[d - 1] + Sp[d + 1] - 2 * Sp[d], 1);
                        d = d * 16 + ((Sp[d - 1] - Sp[d + 1]) * 16 + den) / (den * 2);
                    } else d *= 16;
                    drow[x + minX1] = (int16_t)(d + minD * 16);
                }
                for (int x = minX1; x < maxX1; x++) {
                    int d1 = drow[x];
                    if (d1 == INVS) continue;
                    int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
                    int _x = x - _d, x_ = x - d_;
                    if (0 <= _x && _x < W && disp2[_x] >= minD && abs(disp2[_x] - _d) > d12 &&
                        0 <= x_ && x_ < W && disp2[x_] >= minD && abs(disp2[x_] - d_) > d12)
                        drow[x] = (int16_t)INVS;
                }
            }
            free(disp2);
        }
        free(S);
    }
    /* medianBlur(disp, 3) over the whole map, then speckles with range x16 */
    int16_t *tmp = (int16_t *)malloc(sizeof(int16_t) * (size_t)W * H);
    for (int y = 0; y < H; y++) memcpy(tmp + (size_t)y * W, disp + (size_t)y * dstep, sizeof(int16_t) * W);
    orc_median3_s16(tmp, W, disp, dstep, W, H);
    free(tmp);
    if (p->speckleWindowSize > 0)
        orc_filter_speckles(disp, dstep, W, H, INVS, p->speckleWindowSize, 16 * p->speckleRange);
    return outside;
}
