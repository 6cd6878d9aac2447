/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by
 * the product path (visual-odometry-project_b200/).  Only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * CPU restatement of the reference Harris detector,
 *   /root/reference/src/vo/features/harris.py:86-158  (extractKeypoints)
 *   /root/reference/src/vo/features/harris.py:160-194 (extractDescriptors)
 *
 * Parity pinned: tests/golden/harris_kitti_*.npz were produced by importing the
 * reference itself (tests/golden/make_golden.py) on its own KITTI test frames.
 *
 * Arithmetic notes (what makes a bit-exact restatement possible):
 *  - harris.py:108-109 call scipy.signal.convolve2d(sobel, img, mode="valid")
 *    on integer arrays -> exact int64 gradients of shape (H-2, W-2).  True
 *    convolution flips the kernel; the sign flip cancels in Ix*Ix, Iy*Iy, Ix*Iy.
 *  - harris.py:118-120 box sums with a float64 ones() patch: all partial sums
 *    are integers < 2^53, so float64 accumulation is exact in any order.
 *  - harris.py:123-127: trace, determinant and the score are separate numpy
 *    ufunc calls (no FMA contraction): every product / difference below is an
 *    individually rounded double op.  Build with -ffp-contract=off.
 *  - harris.py:129-137 pads the (H-2r-2, W-2r-2) score map back to (H, W) with
 *    patch_radius+1 zeros on every side.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* response map, float64, shape (H, W); img is uint8 with row pitch `pitch`. */
int oracle_harris_response(const uint8_t *img, int H, int W, int pitch,
                           int patch_size, double kappa, double *resp)
{
    const int pr = patch_size / 2;
    const int pad = pr + 1;
    if (H < 2 * pad + 1 || W < 2 * pad + 1 || patch_size < 1) return -1;
    const int gh = H - 2, gw = W - 2;            /* gradient image size */
    int64_t *ixx = malloc(sizeof(int64_t) * gh * gw);
    int64_t *iyy = malloc(sizeof(int64_t) * gh * gw);
    int64_t *ixy = malloc(sizeof(int64_t) * gh * gw);
    if (!ixx || !iyy || !ixy) return -2;
    for (int i = 0; i < gh; i++) {
        const uint8_t *r0 = img + (size_t)i * pitch;
        const uint8_t *r1 = r0 + pitch, *r2 = r1 + pitch;
        for (int j = 0; j < gw; j++) {
            /* harris.py:103-109, flipped-kernel convolution, centre (i+1, j+1) */
            int64_t gx = ((int)r0[j] + 2 * (int)r1[j] + (int)r2[j])
                       - ((int)r0[j + 2] + 2 * (int)r1[j + 2] + (int)r2[j + 2]);
            int64_t gy = ((int)r0[j] + 2 * (int)r0[j + 1] + (int)r0[j + 2])
                       - ((int)r2[j] + 2 * (int)r2[j + 1] + (int)r2[j + 2]);
            ixx[(size_t)i * gw + j] = gx * gx;          /* harris.py:111 */
            iyy[(size_t)i * gw + j] = gy * gy;          /* harris.py:112 */
            ixy[(size_t)i * gw + j] = gx * gy;          /* harris.py:113 */
        }
    }
    memset(resp, 0, sizeof(double) * (size_t)H * W);
    const int sh = gh - patch_size + 1, sw = gw - patch_size + 1;
    for (int i = 0; i < sh; i++) {
        for (int j = 0; j < sw; j++) {
            int64_t a = 0, b = 0, c = 0;                /* harris.py:118-120 */
            for (int u = 0; u < patch_size; u++) {
                const size_t base = (size_t)(i + u) * gw + j;
                for (int v = 0; v < patch_size; v++) {
                    a += ixx[base + v];
                    b += iyy[base + v];
                    c += ixy[base + v];
                }
            }
            const double sa = (double)a, sb = (double)b, sc = (double)c;
            const double trace = sa + sb;               /* harris.py:123 */
            const double p1 = sa * sb;
            const double p2 = sc * sc;
            const double det = p1 - p2;                 /* harris.py:124 */
            const double t2 = trace * trace;
            const double kt = kappa * t2;
            double s = det - kt;                        /* harris.py:126 */
            if (s < 0) s = 0;                           /* harris.py:127 */
            resp[(size_t)(i + pad) * W + (j + pad)] = s;   /* harris.py:129-137 */
        }
    }
    free(ixx); free(iyy); free(ixy);
    return 0;
}

/*
 * Greedy non-maximum suppression, harris.py:148-152, restated literally:
 * repeat num_keypoints times { argmax over the whole map (first occurrence on
 * ties, as numpy.argmax); zero the (2r+1)^2 box; record (x=w_max, y=h_max) }.
 * numpy slicing semantics are kept, including the negative-start case
 * (h_max - r < 0 wraps to H + h_max - r and normally gives an empty slice, so
 * nothing is zeroed and the same pixel is returned again).
 * `scores` is modified in place.  kp_xy: int32[num_keypoints][2] = (x, y).
 */
static void py_slice(int start, int stop, int n, int *lo, int *hi)
{
    if (start < 0) { start += n; if (start < 0) start = 0; }
    if (start > n) start = n;
    if (stop < 0) { stop += n; if (stop < 0) stop = 0; }
    if (stop > n) stop = n;
    *lo = start; *hi = stop;
}

int oracle_harris_nms(double *scores, int H, int W, int radius,
                      int num_keypoints, int32_t *kp_xy)
{
    const size_t n = (size_t)H * W;
    for (int k = 0; k < num_keypoints; k++) {
        size_t best = 0; double bv = scores[0];
        for (size_t i = 1; i < n; i++)
            if (scores[i] > bv) { bv = scores[i]; best = i; }
        const int hm = (int)(best / W), wm = (int)(best % W);
        int y0, y1, x0, x1;
        py_slice(hm - radius, hm + radius + 1, H, &y0, &y1);
        py_slice(wm - radius, wm + radius + 1, W, &x0, &x1);
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++)
                scores[(size_t)y * W + x] = 0.0;
        kp_xy[2 * k] = wm; kp_xy[2 * k + 1] = hm;
    }
    return 0;
}

/*
 * Patch descriptors, harris.py:160-194: (2r+1)^2 raw pixels around each
 * keypoint of a zero-padded image, row-major, as float64.
 */
int oracle_harris_descriptors(const uint8_t *img, int H, int W, int pitch,
                              const int32_t *kp_xy, int n_kp, int r, double *desc)
{
    const int d = 2 * r + 1;
    for (int k = 0; k < n_kp; k++) {
        const int cx = kp_xy[2 * k], cy = kp_xy[2 * k + 1];
        double *out = desc + (size_t)k * d * d;
        for (int dy = -r; dy <= r; dy++)
            for (int dx = -r; dx <= r; dx++) {
                const int y = cy + dy, x = cx + dx;
                double v = 0.0;
                if (y >= 0 && y < H && x >= 0 && x < W) v = img[(size_t)y * pitch + x];
                *out++ = v;
            }
    }
    return 0;
}
