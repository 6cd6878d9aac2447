/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * CPU restatement of the pyramidal Lucas-Kanade tracker the reference calls at
 *   /root/reference/src/vo/features/klt.py:233-239   cv2.calcOpticalFlowPyrLK(prev, next, pts, None,
 *        winSize=(17,17), maxLevel=2, criteria=(EPS|COUNT, 10, 0.03))          (klt.py:29-33)
 *   /root/reference/src/vo/features/klt.py:244-249   status / error filter
 *
 * The algorithm lives in a third-party dependency that is not part of /root/reference:
 * OpenCV (opencv-python==4.8.1.78 pinned in environment.yml; this image has 4.13.0),
 * video/src/lkpyramid.cpp + imgproc pyrDown.  Published algorithm restated here:
 *   - pyramid: level l+1 = pyrDown(level l): separable [1 4 6 4 1]/16 Gaussian, BORDER_REFLECT_101,
 *     integer arithmetic, (sum + 128) >> 8, size ((w+1)/2, (h+1)/2); the pyramid stops before a
 *     level that is not larger than the window.
 *   - derivatives: Scharr ([3 10 3] x [-1 0 1]) on each level, int16, reflect-101 at the image
 *     edge, zero outside the image; image samples outside the image are reflect-101.
 *   - per point, coarse to fine: fixed-point (14-bit weights) bilinear patch of I (5 fractional
 *     bits) and of its derivatives, 2x2 Gram matrix, min-eigenvalue test, then <= maxCount
 *     Newton steps  delta = G^-1 b  with the EPS / oscillation stopping rules; level-0 failures
 *     clear `status`; `err` is the mean absolute patch difference / 32.
 * Integer sums (Gram matrix, mismatch vector, error) are accumulated exactly (int64) and converted
 * to float32 once; OpenCV accumulates them in float32 SIMD lanes, which differs in the last bits.
 * All float32 operations keep OpenCV's order.  Parity with cv2 is pinned by tests/golden/klt.npz.
 *
 * Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define W_BITS 14
#define KLT_MAX_LEVELS 8

static inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}

/* cv2.pyrDown for 8-bit single channel */
void oracle_pyr_down(const uint8_t *src, int H, int W, int pitch, uint8_t *dst, int dpitch)
{
    const int dh = (H + 1) / 2, dw = (W + 1) / 2;
    int *row = malloc(sizeof(int) * 5 * dw);
    for (int dy = 0; dy < dh; dy++) {
        for (int k = 0; k < 5; k++) {
            const int sy = reflect101(2 * dy + k - 2, H);
            const uint8_t *s = src + (size_t)sy * pitch;
            int *r = row + k * dw;
            for (int dx = 0; dx < dw; dx++) {
                const int x = 2 * dx;
                r[dx] = s[reflect101(x - 2, W)] + s[reflect101(x + 2, W)] +
                        4 * (s[reflect101(x - 1, W)] + s[reflect101(x + 1, W)]) + 6 * s[x];
            }
        }
        for (int dx = 0; dx < dw; dx++) {
            const int v = row[dx] + row[4 * dw + dx] + 4 * (row[dw + dx] + row[3 * dw + dx]) + 6 * row[2 * dw + dx];
            dst[(size_t)dy * dpitch + dx] = (uint8_t)((v + 128) >> 8);
        }
    }
    free(row);
}

/* number of usable levels and their sizes (buildOpticalFlowPyramid's stopping rule) */
int oracle_klt_levels(int H, int W, int max_level, int win, int *lh, int *lw)
{
    int h = H, w = W, level = 0;
    for (;;) {
        lh[level] = h; lw[level] = w;
        if (level == max_level) break;
        const int nh = (h + 1) / 2, nw = (w + 1) / 2;
        if (nw <= win || nh <= win) break;
        h = nh; w = nw; level++;
    }
    return level + 1;
}

typedef struct {
    int h, w;
    uint8_t *img;      /* [h][w] */
    int16_t *deriv;    /* [h][w][2] Scharr (Ix, Iy) */
} Level;

static void scharr(const uint8_t *img, int h, int w, int16_t *d)
{
    for (int y = 0; y < h; y++) {
        const uint8_t *r0 = img + (size_t)reflect101(y - 1, h) * w;
        const uint8_t *r1 = img + (size_t)y * w;
        const uint8_t *r2 = img + (size_t)reflect101(y + 1, h) * w;
        for (int x = 0; x < w; x++) {
            const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            const int t0m = (r0[xm] + r2[xm]) * 3 + r1[xm] * 10;
            const int t0p = (r0[xp] + r2[xp]) * 3 + r1[xp] * 10;
            const int t1m = r2[xm] - r0[xm], t1c = r2[x] - r0[x], t1p = r2[xp] - r0[xp];
            d[((size_t)y * w + x) * 2] = (int16_t)(t0p - t0m);
            d[((size_t)y * w + x) * 2 + 1] = (int16_t)((t1p + t1m) * 3 + t1c * 10);
        }
    }
}

static inline int img_at(const Level *L, int x, int y)
{
    return L->img[(size_t)reflect101(y, L->h) * L->w + reflect101(x, L->w)];
}
static inline int deriv_at(const Level *L, int x, int y, int c)
{
    if (x < 0 || x >= L->w || y < 0 || y >= L->h) return 0;
    return L->deriv[((size_t)y * L->w + x) * 2 + c];
}

static inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

static void build_levels(const uint8_t *img, int H, int W, int pitch, int n_levels, Level *lv)
{
    lv[0].h = H; lv[0].w = W;
    lv[0].img = malloc((size_t)H * W);
    for (int y = 0; y < H; y++) memcpy(lv[0].img + (size_t)y * W, img + (size_t)y * pitch, W);
    for (int l = 1; l < n_levels; l++) {
        lv[l].h = (lv[l - 1].h + 1) / 2; lv[l].w = (lv[l - 1].w + 1) / 2;
        lv[l].img = malloc((size_t)lv[l].h * lv[l].w);
        oracle_pyr_down(lv[l - 1].img, lv[l - 1].h, lv[l - 1].w, lv[l - 1].w, lv[l].img, lv[l].w);
    }
}

/*
 * cv2.calcOpticalFlowPyrLK(prev, next, prev_pts, None, winSize=(win,win), maxLevel=max_level,
 *     criteria=(EPS|COUNT, max_iters, epsilon), flags=0, minEigThreshold=min_eig)
 * prev_pts/next_pts: float32 [n][2]; status uint8 [n]; err float32 [n].
 */
int oracle_klt_track(const uint8_t *prev, const uint8_t *next, int H, int W, int pitch,
                     int max_level, int win, int max_iters, double epsilon, double min_eig,
                     const float *prev_pts, int n_pts, float *next_pts, uint8_t *status, float *err)
{
    if (win < 3 || win > 31 || max_level < 0 || max_level >= KLT_MAX_LEVELS) return -1;
    int lh[KLT_MAX_LEVELS], lw[KLT_MAX_LEVELS];
    const int n_levels = oracle_klt_levels(H, W, max_level, win, lh, lw);
    Level I[KLT_MAX_LEVELS], J[KLT_MAX_LEVELS];
    build_levels(prev, H, W, pitch, n_levels, I);
    build_levels(next, H, W, pitch, n_levels, J);
    for (int l = 0; l < n_levels; l++) {
        I[l].deriv = malloc(sizeof(int16_t) * 2 * (size_t)I[l].h * I[l].w);
        scharr(I[l].img, I[l].h, I[l].w, I[l].deriv);
        J[l].deriv = NULL;
    }
    /* criteria clamps of calcOpticalFlowPyrLK */
    if (max_iters < 0) max_iters = 0;
    if (max_iters > 100) max_iters = 100;
    if (epsilon < 0.) epsilon = 0.;
    if (epsilon > 10.) epsilon = 10.;
    const double eps2 = epsilon * epsilon;
    const float half = (float)(win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const int top = n_levels - 1;
    int16_t *Iw = malloc(sizeof(int16_t) * win * win);
    int16_t *dIw = malloc(sizeof(int16_t) * win * win * 2);

    for (int p = 0; p < n_pts; p++) { status[p] = 1; err[p] = 0.f; }

    for (int level = top; level >= 0; level--) {
        const Level *Li = &I[level], *Lj = &J[level];
        const int cols = Li->w, rows = Li->h;
        for (int p = 0; p < n_pts; p++) {
            const float sc = (float)(1. / (1 << level));
            float prx = prev_pts[2 * p] * sc, pry = prev_pts[2 * p + 1] * sc;
            float nx, ny;
            if (level == top) { nx = prx; ny = pry; }
            else { nx = next_pts[2 * p] * 2.f; ny = next_pts[2 * p + 1] * 2.f; }
            next_pts[2 * p] = nx; next_pts[2 * p + 1] = ny;

            prx -= half; pry -= half;
            const int ipx = (int)floorf(prx), ipy = (int)floorf(pry);
            if (ipx < -win || ipx >= cols || ipy < -win || ipy >= rows) {
                if (level == 0) { status[p] = 0; err[p] = 0.f; }
                continue;
            }
            float a = prx - ipx, b = pry - ipy;
            int iw00 = (int)lrintf((1.f - a) * (1.f - b) * (1 << W_BITS));
            int iw01 = (int)lrintf(a * (1.f - b) * (1 << W_BITS));
            int iw10 = (int)lrintf((1.f - a) * b * (1 << W_BITS));
            int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            int64_t iA11 = 0, iA12 = 0, iA22 = 0;
            for (int y = 0; y < win; y++) {
                for (int x = 0; x < win; x++) {
                    const int gx = ipx + x, gy = ipy + y;
                    const int ival = descale(img_at(Li, gx, gy) * iw00 + img_at(Li, gx + 1, gy) * iw01 +
                                             img_at(Li, gx, gy + 1) * iw10 + img_at(Li, gx + 1, gy + 1) * iw11,
                                             W_BITS - 5);
                    const int ixval = descale(deriv_at(Li, gx, gy, 0) * iw00 + deriv_at(Li, gx + 1, gy, 0) * iw01 +
                                              deriv_at(Li, gx, gy + 1, 0) * iw10 + deriv_at(Li, gx + 1, gy + 1, 0) * iw11,
                                              W_BITS);
                    const int iyval = descale(deriv_at(Li, gx, gy, 1) * iw00 + deriv_at(Li, gx + 1, gy, 1) * iw01 +
                                              deriv_at(Li, gx, gy + 1, 1) * iw10 + deriv_at(Li, gx + 1, gy + 1, 1) * iw11,
                                              W_BITS);
                    Iw[y * win + x] = (int16_t)ival;
                    dIw[(y * win + x) * 2] = (int16_t)ixval;
                    dIw[(y * win + x) * 2 + 1] = (int16_t)iyval;
                    iA11 += (int64_t)ixval * ixval;
                    iA12 += (int64_t)ixval * iyval;
                    iA22 += (int64_t)iyval * iyval;
                }
            }
            const float A11 = (float)iA11 * FLT_SCALE, A12 = (float)iA12 * FLT_SCALE, A22 = (float)iA22 * FLT_SCALE;
            float D = A11 * A22 - A12 * A12;
            const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
            if ((double)minEig < min_eig || D < 1.1920928955078125e-07f) {
                if (level == 0) status[p] = 0;
                continue;
            }
            D = 1.f / D;
            nx -= half; ny -= half;
            float pdx = 0.f, pdy = 0.f;
            for (int j = 0; j < max_iters; j++) {
                const int inx = (int)floorf(nx), iny = (int)floorf(ny);
                if (inx < -win || inx >= cols || iny < -win || iny >= rows) {
                    if (level == 0) status[p] = 0;
                    break;
                }
                a = nx - inx; b = ny - iny;
                iw00 = (int)lrintf((1.f - a) * (1.f - b) * (1 << W_BITS));
                iw01 = (int)lrintf(a * (1.f - b) * (1 << W_BITS));
                iw10 = (int)lrintf((1.f - a) * b * (1 << W_BITS));
                iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                int64_t ib1 = 0, ib2 = 0;
                for (int y = 0; y < win; y++) {
                    for (int x = 0; x < win; x++) {
                        const int gx = inx + x, gy = iny + y;
                        const int diff = descale(img_at(Lj, gx, gy) * iw00 + img_at(Lj, gx + 1, gy) * iw01 +
                                                 img_at(Lj, gx, gy + 1) * iw10 + img_at(Lj, gx + 1, gy + 1) * iw11,
                                                 W_BITS - 5) - Iw[y * win + x];
                        ib1 += (int64_t)diff * dIw[(y * win + x) * 2];
                        ib2 += (int64_t)diff * dIw[(y * win + x) * 2 + 1];
                    }
                }
                const float b1 = (float)ib1 * FLT_SCALE, b2 = (float)ib2 * FLT_SCALE;
                const float dx = (float)((A12 * b2 - A22 * b1) * D);
                const float dy = (float)((A12 * b1 - A11 * b2) * D);
                nx += dx; ny += dy;
                next_pts[2 * p] = nx + half; next_pts[2 * p + 1] = ny + half;
                if ((double)dx * dx + (double)dy * dy <= eps2) break;
                if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
                    next_pts[2 * p] -= dx * 0.5f;
                    next_pts[2 * p + 1] -= dy * 0.5f;
                    break;
                }
                pdx = dx; pdy = dy;
            }
            if (status[p] && level == 0) {
                const float fx = next_pts[2 * p] - half, fy = next_pts[2 * p + 1] - half;
                const int inx = (int)floorf(fx), iny = (int)floorf(fy);
                if (inx < -win || inx >= cols || iny < -win || iny >= rows) {
                    status[p] = 0;
                    continue;
                }
                const float aa = fx - inx, bb = fy - iny;
                iw00 = (int)lrintf((1.f - aa) * (1.f - bb) * (1 << W_BITS));
                iw01 = (int)lrintf(aa * (1.f - bb) * (1 << W_BITS));
                iw10 = (int)lrintf((1.f - aa) * bb * (1 << W_BITS));
                iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                int64_t e = 0;
                for (int y = 0; y < win; y++)
                    for (int x = 0; x < win; x++) {
                        const int gx = inx + x, gy = iny + y;
                        const int diff = descale(img_at(Lj, gx, gy) * iw00 + img_at(Lj, gx + 1, gy) * iw01 +
                                                 img_at(Lj, gx, gy + 1) * iw10 + img_at(Lj, gx + 1, gy + 1) * iw11,
                                                 W_BITS - 5) - Iw[y * win + x];
                        e += diff < 0 ? -diff : diff;
                    }
                err[p] = (float)e * 1.f / (float)(32 * win * win);
            }
        }
    }
    for (int l = 0; l < n_levels; l++) { free(I[l].img); free(J[l].img); free(I[l].deriv); }
    free(Iw); free(dIw);
    return n_levels;
}
