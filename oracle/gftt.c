/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path.
 *
 * CPU restatement of cv2.goodFeaturesToTrack as the reference calls it,
 *   /root/reference/src/vo/features/klt.py:24-26   maxCorners=500, qualityLevel=0.01, minDistance=8, blockSize=7
 *   /root/reference/src/vo/features/klt.py:87-115  find_corners (mask = all 255: no effect)
 * OpenCV is a third-party dependency that is not vendored in /root/reference (environment.yml pins
 * opencv-python 4.8.1; this image has 4.13.0): the algorithm restated here is the one of
 * modules/imgproc/src/featureselect.cpp + corner.cpp (cornerMinEigenVal):
 *   1. Dx, Dy = Sobel(src, CV_32F, aperture 3, scale = 1 / (4 * blockSize * 255), BORDER_REFLECT_101)
 *   2. cov = (Dx*Dx, Dx*Dy, Dy*Dy) in float32;  boxFilter(blockSize x blockSize, not normalised, REFLECT_101) with
 *      float64 running sums (rows: s += new - old; columns: SUM += row, out = (float)(SUM), SUM -= oldest row)
 *   3. eig = (a + c) - sqrt((a - c)^2 + b^2) with a = 0.5 xx, b = xy, c = 0.5 yy, float32, no FMA
 *   4. threshold at qualityLevel * max(eig); a pixel is a candidate if it is non-zero and equals the 3x3 maximum,
 *      image border excluded; candidates sorted by (value descending, address descending)
 *   5. greedy selection: a candidate is taken unless an already taken one is closer than minDistance (Euclidean);
 *      stop at maxCorners.
 * Parity pinned (tests/test_oracle_golden.py): the eigenvalue map is BIT-EQUAL to cv2.cornerMinEigenVal and the corner
 * list equal to cv2.goodFeaturesToTrack on the KITTI frames and on synthetic images, on hosts where OpenCV dispatches
 * its AVX2 + FMA3 filter code (every x86 CPU since 2013).  Float details that matter for that:
 *   - the Sobel kernels are scaled on their smoothing side: k = float32([1, 2, 1] * scale);
 *   - Dx: rows [-1 0 1] (exact integers), columns fma(S0 + S2, k1, S1 * k0);
 *   - Dy: rows are filtered 32 columns at a time as fma(c, k1, fma(b, k0, a * k1)); the last W mod 32 columns by the
 *     scalar loop (a * k1 + b * k0) + c * k1; columns [-1 0 1] are an exact difference.
 * Build: gcc -ffp-contract=off; fused operations are written as fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * n - 2 - i; }
    return i;
}

/* cv2.cornerMinEigenVal(img, blockSize, ksize=3): eig float32 (H, W) */
int oracle_min_eigen_val(const uint8_t *img, int H, int W, int pitch, int block, float *eig)
{
    if (H < 1 || W < 1 || block < 1 || !(block & 1)) return -1;
    const double scale = 1.0 / ((double)(1 << 2) * block) / 255.0;     /* corner.cpp: 1/((1<<(aperture-1))*block), /255 for 8U */
    const float k1 = (float)((double)1.0f * scale), k0 = (float)((double)2.0f * scale);
    const size_t n = (size_t)H * W;
    float *dx = malloc(n * 4), *dy = malloc(n * 4), *rowx = malloc((size_t)(H + 2) * W * 4), *rowy = malloc((size_t)(H + 2) * W * 4);
    float *cov = malloc(n * 12);
    if (!dx || !dy || !rowx || !rowy || !cov) return -2;
    const int nb = (W / 32) * 32;
    /* row pass on rows -1 .. H (reflected), both derivative images */
    for (int y = -1; y <= H; y++) {
        const uint8_t *r = img + (size_t)reflect101(y, H) * pitch;
        float *ox = rowx + (size_t)(y + 1) * W, *oy = rowy + (size_t)(y + 1) * W;
        for (int x = 0; x < W; x++) {
            const float a = (float)r[reflect101(x - 1, W)], b = (float)r[x], c = (float)r[reflect101(x + 1, W)];
            ox[x] = c - a;                                             /* [-1 0 1]: exact */
            if (x < nb) oy[x] = fmaf(c, k1, fmaf(b, k0, a * k1));     /* vector body */
            else oy[x] = (a * k1 + b * k0) + c * k1;                  /* scalar tail */
        }
    }
    for (int y = 0; y < H; y++) {
        const float *x0 = rowx + (size_t)y * W, *x1 = x0 + W, *x2 = x1 + W;
        const float *y0 = rowy + (size_t)y * W, *y2 = y0 + 2 * (size_t)W;
        for (int x = 0; x < W; x++) {
            const float gx = fmaf(x0[x] + x2[x], k1, x1[x] * k0);     /* symmetric column filter, fused */
            const float gy = y2[x] - y0[x];
            dx[(size_t)y * W + x] = gx; dy[(size_t)y * W + x] = gy;
            float *cv = cov + ((size_t)y * W + x) * 3;
            cv[0] = gx * gx; cv[1] = gx * gy; cv[2] = gy * gy;
        }
    }
    /* box filter: row sums (float64, sliding) of every row, then sliding column sums */
    const int r = block / 2;
    double *rs = malloc(n * 24);
    double *SUM = calloc((size_t)W * 3, 8);
    if (!rs || !SUM) return -2;
    for (int y = 0; y < H; y++) {
        const float *cv = cov + (size_t)y * W * 3;
        double s[3] = {0, 0, 0};
        for (int k = -r; k <= r; k++) {
            const float *p = cv + (size_t)reflect101(k, W) * 3;
            s[0] += (double)p[0]; s[1] += (double)p[1]; s[2] += (double)p[2];
        }
        double *o = rs + (size_t)y * W * 3;
        o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
        for (int x = 1; x < W; x++) {
            const float *pn = cv + (size_t)reflect101(x + r, W) * 3, *po = cv + (size_t)reflect101(x - 1 - r, W) * 3;
            for (int c = 0; c < 3; c++) { s[c] += (double)pn[c] - (double)po[c]; o[3 * x + c] = s[c]; }
        }
    }
    for (int k = -r; k < r; k++) {
        const double *p = rs + (size_t)reflect101(k, H) * W * 3;
        for (int i = 0; i < 3 * W; i++) SUM[i] += p[i];
    }
    for (int y = 0; y < H; y++) {
        const double *sp = rs + (size_t)reflect101(y + r, H) * W * 3, *sm = rs + (size_t)reflect101(y - r, H) * W * 3;
        for (int x = 0; x < W; x++) {
            float v[3];
            for (int c = 0; c < 3; c++) {
                const double s0 = SUM[3 * x + c] + sp[3 * x + c];
                v[c] = (float)s0;
                SUM[3 * x + c] = s0 - sm[3 * x + c];
            }
            const float a = v[0] * 0.5f, b = v[1], c = v[2] * 0.5f;
            const float t = a - c;
            eig[(size_t)y * W + x] = (a + c) - sqrtf(t * t + b * b);
        }
    }
    free(dx); free(dy); free(rowx); free(rowy); free(cov); free(rs); free(SUM);
    return 0;
}

typedef struct { float v; int idx; } Cand;
static int cand_cmp(const void *pa, const void *pb) {
    const Cand *a = pa, *b = pb;
    if (a->v > b->v) return -1;
    if (a->v < b->v) return 1;
    return a->idx > b->idx ? -1 : (a->idx < b->idx ? 1 : 0);          /* greaterThanPtr: higher address first */
}

/* cv2.goodFeaturesToTrack on a precomputed eigenvalue map (steps 4 and 5): returns the number of corners, xy float32 */
int oracle_gftt_select(const float *eig, int H, int W, int max_corners, double quality, double min_distance, float *xy,
                       int *n_candidates)
{
    float mx = -INFINITY;
    for (size_t i = 0; i < (size_t)H * W; i++) if (eig[i] > mx) mx = eig[i];
    const double thr_d = (double)mx * quality;
    /* threshold(eig, eig, maxVal*qualityLevel, 0, THRESH_TOZERO): keeps v > thresh (thresh converted to float) */
    const float thr = (float)thr_d;
    float *e = malloc((size_t)H * W * 4);
    if (!e) return -2;
    for (size_t i = 0; i < (size_t)H * W; i++) e[i] = eig[i] > thr ? eig[i] : 0.f;
    Cand *c = malloc(sizeof(Cand) * (size_t)H * W / 2 + 64);
    int nc = 0;
    for (int y = 1; y < H - 1; y++)
        for (int x = 1; x < W - 1; x++) {
            const float v = e[(size_t)y * W + x];
            if (v == 0.f) continue;
            float m = v;
            for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) { const float u = e[(size_t)(y + dy) * W + x + dx]; if (u > m) m = u; }
            if (v == m) { c[nc].v = v; c[nc].idx = y * W + x; nc++; }
        }
    qsort(c, nc, sizeof(Cand), cand_cmp);
    if (n_candidates) *n_candidates = nc;
    int n = 0;
    if (min_distance >= 1) {
        const double md2 = min_distance * min_distance;
        for (int i = 0; i < nc; i++) {
            const int y = c[i].idx / W, x = c[i].idx % W;
            int good = 1;
            for (int j = 0; j < n && good; j++) {
                const float dx = (float)x - xy[2 * j], dy = (float)y - xy[2 * j + 1];
                if ((double)(dx * dx + dy * dy) < md2) good = 0;
            }
            if (good) { xy[2 * n] = (float)x; xy[2 * n + 1] = (float)y; n++; if (max_corners > 0 && n == max_corners) break; }
        }
    } else {
        for (int i = 0; i < nc; i++) { xy[2 * n] = (float)(c[i].idx % W); xy[2 * n + 1] = (float)(c[i].idx / W); n++; if (max_corners > 0 && n == max_corners) break; }
    }
    free(e); free(c);
    return n;
}
