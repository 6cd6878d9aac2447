// Device functions of the P3P minimal solver and the reprojection test, shared by p3p.cu (batched scoring) and
// pipeline.cu (the per-sequence RANSAC of the resident pipeline).  See p3p.cu for the algorithm notes; float64 with
// + - * / sqrt only and --fmad=false, in the same order as the CPU checker used by the tests.
#pragma once
#include "common.cuh"

namespace p3pdev {

constexpr int P3P_NEWTON_ITERS = 40;
constexpr int P3P_REFINE_ITERS = 5;

struct Intr { double fx, fy, cx, cy; };

static __device__ double cubic_real_root(double b, double c, double d) {
    double r0;
    const double disc = b * b - 3.0 * c;
    if (disc >= 0.0) {
        const double v = sqrt(disc);
        const double t1 = (-b - v) / 3.0;
        double k = ((t1 + b) * t1 + c) * t1 + d;
        if (k > 0.0) {
            r0 = t1 - sqrt(-k / (3.0 * t1 + b));
        } else {
            const double t2 = (-b + v) / 3.0;
            k = ((t2 + b) * t2 + c) * t2 + d;
            r0 = t2 + sqrt(-k / (3.0 * t2 + b));
        }
    } else {
        r0 = -b / 3.0;
        if (fabs((3.0 * r0 + 2.0 * b) * r0 + c) < 1e-4) r0 += 1.0;
    }
    if (!(r0 == r0)) r0 = -b / 3.0;
    for (int it = 0; it < P3P_NEWTON_ITERS; it++) {
        const double fx = ((r0 + b) * r0 + c) * r0 + d;
        if (it >= 7 && fabs(fx) < 1e-13) break;
        const double fpx = (3.0 * r0 + 2.0 * b) * r0 + c;
        if (fpx == 0.0) break;
        r0 -= fx / fpx;
    }
    return r0;
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double* a, const double* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

static __device__ bool null_vector(const double* m0, const double* m1, const double* m2, double* e) {
    double c01[3], c02[3], c12[3];
    cross3(m0, m1, c01); cross3(m0, m2, c02); cross3(m1, m2, c12);
    const double n01 = dot3(c01, c01), n02 = dot3(c02, c02), n12 = dot3(c12, c12);
    double bx = c01[0], by = c01[1], bz = c01[2], nb = n01;
    if (n02 > nb) { bx = c02[0]; by = c02[1]; bz = c02[2]; nb = n02; }
    if (n12 > nb) { bx = c12[0]; by = c12[1]; bz = c12[2]; nb = n12; }
    if (!(nb > 0.0)) return false;
    const double inv = 1.0 / sqrt(nb);
    e[0] = bx * inv; e[1] = by * inv; e[2] = bz * inv;
    return true;
}

static __device__ int p3p_depths(const double (*y)[3], const double (*X)[3], double (*Ls)[3]) {
    const double b12 = dot3(y[0], y[1]), b13 = dot3(y[0], y[2]), b23 = dot3(y[1], y[2]);
    double d12[3], d13[3], d23[3];
    for (int i = 0; i < 3; i++) { d12[i] = X[0][i] - X[1][i]; d13[i] = X[0][i] - X[2][i]; d23[i] = X[1][i] - X[2][i]; }
    const double a12 = dot3(d12, d12), a13 = dot3(d13, d13), a23 = dot3(d23, d23);
    if (!(a12 > 0.0) || !(a13 > 0.0) || !(a23 > 0.0)) return 0;

    const double p00 = a23, p01 = -a23 * b12, p02 = 0.0, p11 = a23 - a12, p12 = a12 * b23, p22 = -a12;
    const double q00 = a23, q01 = 0.0, q02 = -a23 * b13, q11 = -a13, q12 = a13 * b23, q22 = a23 - a13;

    const double A00 = p11 * p22 - p12 * p12, A01 = p02 * p12 - p01 * p22, A02 = p01 * p12 - p02 * p11;
    const double A11 = p00 * p22 - p02 * p02, A12 = p01 * p02 - p00 * p12, A22 = p00 * p11 - p01 * p01;
    const double B00 = q11 * q22 - q12 * q12, B01 = q02 * q12 - q01 * q22, B02 = q01 * q12 - q02 * q11;
    const double B11 = q00 * q22 - q02 * q02, B12 = q01 * q02 - q00 * q12, B22 = q00 * q11 - q01 * q01;
    const double c0 = p00 * A00 + p01 * A01 + p02 * A02;
    const double c3 = q00 * B00 + q01 * B01 + q02 * B02;
    const double c1 = A00 * q00 + A11 * q11 + A22 * q22 + 2.0 * (A01 * q01 + A02 * q02 + A12 * q12);
    const double c2 = B00 * p00 + B11 * p11 + B22 * p22 + 2.0 * (B01 * p01 + B02 * p02 + B12 * p12);
    if (c3 == 0.0) return 0;
    const double ic3 = 1.0 / c3;
    const double g = cubic_real_root(c2 * ic3, c1 * ic3, c0 * ic3);

    const double m00 = p00 + g * q00, m01 = p01 + g * q01, m02 = p02 + g * q02;
    const double m11 = p11 + g * q11, m12 = p12 + g * q12, m22 = p22 + g * q22;
    const double tr = m00 + m11 + m22;
    const double mm = (m00 * m11 - m01 * m01) + (m00 * m22 - m02 * m02) + (m11 * m22 - m12 * m12);
    const double dsc = tr * tr - 4.0 * mm;
    if (!(dsc >= 0.0)) return 0;
    const double sq = sqrt(dsc);
    const double s1 = 0.5 * (tr + sq), s2 = 0.5 * (tr - sq);
    if (!(s1 > 0.0) || !(s2 < 0.0)) return 0;
    double e1[3], e2[3];
    {
        const double r0[3] = {m00 - s1, m01, m02}, r1[3] = {m01, m11 - s1, m12}, r2[3] = {m02, m12, m22 - s1};
        if (!null_vector(r0, r1, r2, e1)) return 0;
    }
    {
        const double r0[3] = {m00 - s2, m01, m02}, r1[3] = {m01, m11 - s2, m12}, r2[3] = {m02, m12, m22 - s2};
        if (!null_vector(r0, r1, r2, e2)) return 0;
    }
    const double s = sqrt(-s2 / s1);
    const bool use_d1 = fabs(g) >= 1.0;
    int n = 0;
    for (int sign = 0; sign < 2; sign++) {
        const double sg = sign ? -s : s;
        const double l0 = e1[0] + sg * e2[0], l1c = e1[1] + sg * e2[1], l2c = e1[2] + sg * e2[2];
        if (fabs(l0) < 1e-300) continue;
        const double w0 = -l1c / l0, w1 = -l2c / l0;
        double qa, qb, qc;
        if (use_d1) {
            qa = a23 * w1 * w1 - a12;
            qb = a23 * (2.0 * w0 * w1 - 2.0 * b12 * w1) + 2.0 * a12 * b23;
            qc = a23 * (w0 * w0 + 1.0 - 2.0 * b12 * w0) - a12;
        } else {
            qa = a23 * (w1 * w1 + 1.0 - 2.0 * b13 * w1) - a13;
            qb = a23 * (2.0 * w0 * w1 - 2.0 * b13 * w0) + 2.0 * a13 * b23;
            qc = a23 * w0 * w0 - a13;
        }
        double taus[2];
        int nt = 0;
        if (qa == 0.0) {
            if (qb != 0.0) taus[nt++] = -qc / qb;
        } else {
            const double dq = qb * qb - 4.0 * qa * qc;
            if (dq >= 0.0) {
                const double sd = sqrt(dq);
                const double qq = -0.5 * (qb + (qb >= 0.0 ? sd : -sd));
                taus[nt++] = qq / qa;
                if (qq != 0.0) taus[nt++] = qc / qq;
            }
        }
        for (int k = 0; k < nt; k++) {
            const double tau = taus[k];
            if (!(tau > 0.0)) continue;
            const double den = tau * (tau - 2.0 * b23) + 1.0;
            if (!(den > 0.0)) continue;
            const double L2 = sqrt(a23 / den);
            const double L3 = tau * L2;
            const double L1 = L2 * (w0 + w1 * tau);
            if (!(L1 > 0.0)) continue;
            if (n < 4) { Ls[n][0] = L1; Ls[n][1] = L2; Ls[n][2] = L3; n++; }
        }
    }
    for (int k = 0; k < n; k++) {
        double L1 = Ls[k][0], L2 = Ls[k][1], L3 = Ls[k][2];
        for (int it = 0; it < P3P_REFINE_ITERS; it++) {
            const double r1 = L1 * L1 + L2 * L2 - 2.0 * b12 * L1 * L2 - a12;
            const double r2 = L1 * L1 + L3 * L3 - 2.0 * b13 * L1 * L3 - a13;
            const double r3 = L2 * L2 + L3 * L3 - 2.0 * b23 * L2 * L3 - a23;
            if (fabs(r1) + fabs(r2) + fabs(r3) < 1e-10) break;
            const double j00 = 2.0 * (L1 - b12 * L2), j01 = 2.0 * (L2 - b12 * L1);
            const double j10 = 2.0 * (L1 - b13 * L3), j12 = 2.0 * (L3 - b13 * L1);
            const double j21 = 2.0 * (L2 - b23 * L3), j22 = 2.0 * (L3 - b23 * L2);
            const double det = -j00 * j12 * j21 - j01 * j10 * j22;
            if (fabs(det) < 1e-300) break;
            const double idet = 1.0 / det;
            const double dl1 = (-j12 * j21 * r1 - j01 * j22 * r2 + j01 * j12 * r3) * idet;
            const double dl2 = (-j10 * j22 * r1 + j00 * j22 * r2 - j00 * j12 * r3) * idet;
            const double dl3 = (j10 * j21 * r1 - j00 * j21 * r2 - j01 * j10 * r3) * idet;
            L1 -= dl1; L2 -= dl2; L3 -= dl3;
        }
        Ls[k][0] = L1; Ls[k][1] = L2; Ls[k][2] = L3;
    }
    return n;
}

static __device__ bool pose_from_depths(const double (*y)[3], const double (*X)[3], const double* L, double* model) {
    double Y[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Y[i][j] = L[i] * y[i][j];
    double x1[3], x2[3], x3[3], y1[3], y2[3], y3[3];
    for (int i = 0; i < 3; i++) {
        x1[i] = X[0][i] - X[1][i]; x2[i] = X[1][i] - X[2][i];
        y1[i] = Y[0][i] - Y[1][i]; y2[i] = Y[1][i] - Y[2][i];
    }
    cross3(x1, x2, x3); cross3(y1, y2, y3);
    double i0[3], i1[3], i2[3];
    cross3(x2, x3, i0); cross3(x3, x1, i1); cross3(x1, x2, i2);
    const double det = dot3(x1, i0);
    if (!(fabs(det) > 0.0)) return false;
    const double id = 1.0 / det;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            model[3 * r + c] = (y1[r] * i0[c] + y2[r] * i1[c] + y3[r] * i2[c]) * id;
    for (int r = 0; r < 3; r++)
        model[9 + r] = Y[0][r] - (model[3 * r] * X[0][0] + model[3 * r + 1] * X[0][1] + model[3 * r + 2] * X[0][2]);
    for (int i = 0; i < 12; i++) if (!(model[i] == model[i])) return false;
    return true;
}

__device__ __forceinline__ double reproj_err2(const double* m, double X0, double X1, double X2, double ku, double kv,
                                              const Intr& K) {
    const double xc = m[0] * X0 + m[1] * X1 + m[2] * X2 + m[9];
    const double yc = m[3] * X0 + m[4] * X1 + m[5] * X2 + m[10];
    const double zc = m[6] * X0 + m[7] * X1 + m[8] * X2 + m[11];
    const double iz = zc != 0.0 ? 1.0 / zc : 1.0;
    const double u = (xc * iz) * K.fx + K.cx;
    const double v = (yc * iz) * K.fy + K.cy;
    const double du = ku - u, dv = kv - v;
    return du * du + dv * dv;
}

// The inlier test of one correspondence under the two rules the reference can run:
//   rule 0  ransac.py:104-106 with p3p.py:81-108: squared reprojection error in float64, err < threshold;
//   rule 1  cv2.solvePnPRansac (p3p.py:142-151): PnPRansacCallback::computeError + findInliers -- the projection is
//           rounded to float32, the squared distance to the (float32) image point is evaluated in float32 and compared
//           with err <= (float)threshold.  The caller passes landmarks already rounded to float32 (solvePnPRansac
//           converts its inputs) and threshold = reprojectionError^2.
__device__ __forceinline__ bool reproj_inlier(const double* m, double X0, double X1, double X2, double ku, double kv, const Intr& K,
                                              double thr, int rule) {
    if (rule == 0) return reproj_err2(m, X0, X1, X2, ku, kv, K) < thr;
    const double xc = m[0] * X0 + m[1] * X1 + m[2] * X2 + m[9];
    const double yc = m[3] * X0 + m[4] * X1 + m[5] * X2 + m[10];
    const double zc = m[6] * X0 + m[7] * X1 + m[8] * X2 + m[11];
    const double iz = zc != 0.0 ? 1.0 / zc : 1.0;
    const float u = (float)((xc * iz) * K.fx + K.cx), v = (float)((yc * iz) * K.fy + K.cy);
    const float dx = __fsub_rn((float)ku, u), dy = __fsub_rn((float)kv, v);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= (float)thr;
}

// model_fn of p3p.py:51-79 for one sample: X4 / uv4 = the four sampled correspondences (three solve, the fourth picks
// among the <= 4 solutions, as cv2.solvePnP(SOLVEPNP_P3P) does).  Returns false when no real solution exists.
static __device__ bool solve4(const double (*X4)[3], const double (*uv4)[2], const Intr& K, double* bm) {
    double y[3][3];
    for (int i = 0; i < 3; i++) {
        const double xn = (uv4[i][0] - K.cx) / K.fx, yn = (uv4[i][1] - K.cy) / K.fy;
        const double inv = 1.0 / sqrt(xn * xn + yn * yn + 1.0);
        y[i][0] = xn * inv; y[i][1] = yn * inv; y[i][2] = inv;
    }
    double Ls[4][3];
    const int n = p3p_depths(y, X4, Ls);
    bool found = false;
    double best = 0.0;
    for (int i = 0; i < 12; i++) bm[i] = 0.0;
    for (int k = 0; k < n; k++) {
        double m[12];
        if (!pose_from_depths(y, X4, Ls[k], m)) continue;
        const double e = reproj_err2(m, X4[3][0], X4[3][1], X4[3][2], uv4[3][0], uv4[3][1], K);
        if (!(e == e)) continue;
        if (!found || e < best) {
            best = e; found = true;
            for (int i = 0; i < 12; i++) bm[i] = m[i];
        }
    }
    return found;
}

}  // namespace p3pdev
