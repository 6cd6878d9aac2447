// Batched P3P + RANSAC scoring for sm_100a.
//
// Replaces  /root/reference/src/vo/pose_estimation/p3p.py:51-79   model_fn (cv2.solvePnP, SOLVEPNP_P3P, 4 points)
//           /root/reference/src/vo/pose_estimation/p3p.py:81-108  error_fn (squared reprojection error)
//           /root/reference/src/vo/algorithms/ransac.py:90-121    hypothesis loop with adaptive iteration count
//
// Layout: one thread per hypothesis solves Grunert's distance system (conic-pencil elimination:
// one cubic root by Newton, two quadratics, Gauss-Newton polish), aligns the triples to (R, t)
// and lets the 4th sample pick among the <=4 solutions; one warp per hypothesis counts
// reprojection inliers with ballot/popc; one CTA per frame replays the reference's sequential
// loop (adaptive stop via a host-computed table) and emits the inlier mask of the winner.
// Everything is float64 with + - * / sqrt only and --fmad=false, in the same order as the CPU
// oracle, so models, counts and masks are reproducible bit for bit.
#include "common.cuh"

namespace {

constexpr int P3P_NEWTON_ITERS = 40;
constexpr int P3P_REFINE_ITERS = 5;

struct Intr { double fx, fy, cx, cy; };

__device__ double cubic_real_root(double b, double c, double d) {
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

__device__ bool null_vector(const double* m0, const double* m1, const double* m2, double* e) {
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

__device__ int p3p_depths(const double (*y)[3], const double (*X)[3], double (*Ls)[3]) {
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

__device__ bool pose_from_depths(const double (*y)[3], const double (*X)[3], const double* L, double* model) {
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

// ransac.py:105 `errors < inlier_threshold`; with inclusive != 0 OpenCV's rule `err <= t` (findInliers of
// cv2.solvePnPRansac, the reference's use_opencv=True path, p3p.py:142-151)
__device__ __forceinline__ bool is_inlier(double e, double thr, int inclusive) {
    return inclusive ? (e <= thr) : (e < thr);
}

// one thread per (frame, hypothesis)
__global__ void __launch_bounds__(128)
p3p_solve_kernel(const double* __restrict__ landmarks, const double* __restrict__ keypoints, int N, Intr K,
                 const int* __restrict__ sample_idx, int n_hyp, double* __restrict__ models,
                 unsigned char* __restrict__ valid) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (h >= n_hyp) return;
    const double* L = landmarks + (size_t)f * N * 3;
    const double* P = keypoints + (size_t)f * N * 2;
    const int* S = sample_idx + ((size_t)f * n_hyp + h) * 4;
    double X4[4][3], uv4[4][2];
    bool idx_ok = true;
    for (int j = 0; j < 4; j++) {
        int id = S[j];
        if (id < 0 || id >= N) { idx_ok = false; id = 0; }
        X4[j][0] = L[3 * id]; X4[j][1] = L[3 * id + 1]; X4[j][2] = L[3 * id + 2];
        uv4[j][0] = P[2 * id]; uv4[j][1] = P[2 * id + 1];
    }
    double y[3][3];
    for (int i = 0; i < 3; i++) {
        const double xn = (uv4[i][0] - K.cx) / K.fx, yn = (uv4[i][1] - K.cy) / K.fy;
        const double inv = 1.0 / sqrt(xn * xn + yn * yn + 1.0);
        y[i][0] = xn * inv; y[i][1] = yn * inv; y[i][2] = inv;
    }
    double Ls[4][3];
    const int n = idx_ok ? p3p_depths(y, X4, Ls) : 0;
    bool found = false;
    double best = 0.0;
    double bm[12];
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
    double* out = models + ((size_t)f * n_hyp + h) * 12;
    for (int i = 0; i < 12; i++) out[i] = bm[i];
    valid[(size_t)f * n_hyp + h] = found ? 1 : 0;
}

// one warp per (frame, hypothesis): inlier count over all N correspondences.  The 16 hypotheses of a CTA belong to
// the same frame, so its correspondences are staged once per CTA in shared memory as five planes (X, Y, Z, u, v):
// the inner loop then reads conflict-free 8-byte words instead of waiting on 24-byte-strided global loads, and the
// FP64 pipe is what is left.
constexpr int COUNT_WARPS = 16;
constexpr int COUNT_CHUNK = 1024;       // correspondences staged at a time (40 KB)
__global__ void __launch_bounds__(COUNT_WARPS * 32)
p3p_count_kernel(const double* __restrict__ landmarks, const double* __restrict__ keypoints, int N, Intr K,
                 const double* __restrict__ models, const unsigned char* __restrict__ valid, int n_hyp,
                 double threshold, int inclusive, int* __restrict__ counts) {
    __shared__ double sX[COUNT_CHUNK], sY[COUNT_CHUNK], sZ[COUNT_CHUNK], sU[COUNT_CHUNK], sV[COUNT_CHUNK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x * COUNT_WARPS + warp;
    const int f = blockIdx.y;
    const size_t hi = (size_t)f * n_hyp + (h < n_hyp ? h : 0);
    const bool active = h < n_hyp && valid[hi] != 0;        // warp-uniform; inactive warps still help staging
    double m[12];
#pragma unroll
    for (int i = 0; i < 12; i++) m[i] = active ? models[hi * 12 + i] : 0.0;
    const double* L = landmarks + (size_t)f * N * 3;
    const double* P = keypoints + (size_t)f * N * 2;
    int c = 0;
    for (int c0 = 0; c0 < N; c0 += COUNT_CHUNK) {
        const int n = min(COUNT_CHUNK, N - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += COUNT_WARPS * 32) {
            const double* l = L + 3 * (size_t)(c0 + i);
            const double* q = P + 2 * (size_t)(c0 + i);
            sX[i] = l[0]; sY[i] = l[1]; sZ[i] = l[2]; sU[i] = q[0]; sV[i] = q[1];
        }
        __syncthreads();
        if (active) {
            for (int base = 0; base < n; base += 64) {     // two points per lane and step: independent FP64 chains
                const int i0 = base + lane, i1 = i0 + 32;
                bool in0 = false, in1 = false;
                if (i0 < n) in0 = is_inlier(reproj_err2(m, sX[i0], sY[i0], sZ[i0], sU[i0], sV[i0], K), threshold, inclusive);
                if (i1 < n) in1 = is_inlier(reproj_err2(m, sX[i1], sY[i1], sZ[i1], sU[i1], sV[i1], K), threshold, inclusive);
                c += __popc(__ballot_sync(0xFFFFFFFFu, in0)) + __popc(__ballot_sync(0xFFFFFFFFu, in1));
            }
        }
    }
    if (lane == 0 && h < n_hyp) counts[(size_t)f * n_hyp + h] = c;
}

// one CTA per frame: ransac.py:90-121 over the pre-scored hypotheses, then the winner's mask
__global__ void __launch_bounds__(256)
p3p_select_kernel(const double* __restrict__ landmarks, const double* __restrict__ keypoints, int N, Intr K,
                  const double* __restrict__ models, const unsigned char* __restrict__ valid,
                  const int* __restrict__ counts, int n_hyp, double threshold, int inclusive,
                  const int* __restrict__ iters_for_count, int initial_iters, int start_n, int start_best,
                  int* __restrict__ best_out, int* __restrict__ consumed_out, int* __restrict__ iters_out,
                  unsigned char* __restrict__ inliers, double* __restrict__ best_model) {
    __shared__ int s_best;
    const int f = blockIdx.x;
    if (threadIdx.x == 0) {
        int n = start_n, best = start_best, best_h = -1, n_iter = initial_iters, h = 0;
        const unsigned char* v = valid + (size_t)f * n_hyp;
        const int* c = counts + (size_t)f * n_hyp;
        while (n < n_iter && h < n_hyp) {
            if (v[h]) {
                const int ch = c[h];
                if (ch > best) { best = ch; best_h = h; n_iter = iters_for_count[ch]; }
                n++;
            }
            h++;
        }
        s_best = best_h;
        best_out[f * 4 + 0] = best_h;
        best_out[f * 4 + 1] = best;
        best_out[f * 4 + 2] = n;
        best_out[f * 4 + 3] = (n < n_iter) ? 1 : 0;  // exhausted: the caller must supply more hypotheses
        consumed_out[f] = h;
        iters_out[f] = n_iter;
    }
    __syncthreads();
    const int bh = s_best;
    unsigned char* in = inliers + (size_t)f * N;
    double* bm = best_model + (size_t)f * 12;
    if (bh < 0) {
        for (int i = threadIdx.x; i < N; i += blockDim.x) in[i] = 0;
        if (threadIdx.x < 12) bm[threadIdx.x] = 0.0;
        return;
    }
    double m[12];
#pragma unroll
    for (int i = 0; i < 12; i++) m[i] = models[((size_t)f * n_hyp + bh) * 12 + i];
    if (threadIdx.x < 12) bm[threadIdx.x] = m[threadIdx.x];
    const double* L = landmarks + (size_t)f * N * 3;
    const double* P = keypoints + (size_t)f * N * 2;
    for (int i = threadIdx.x; i < N; i += blockDim.x)
        in[i] = is_inlier(reproj_err2(m, L[3 * i], L[3 * i + 1], L[3 * i + 2], P[2 * i], P[2 * i + 1], K), threshold, inclusive) ? 1 : 0;
}

}  // namespace

static Intr make_intr(const double* K9) { return Intr{K9[0], K9[4], K9[2], K9[5]}; }

int vo_launch_p3p_score(vo_ctx* ctx, const double* d_landmarks, const double* d_keypoints, int n_frames, int N,
                        const double* K9, const int* d_sample_idx, int n_hyp, double threshold, int inclusive, double* d_models,
                        unsigned char* d_valid, int* d_counts, cudaStream_t stream) {
    VO_REQUIRE(n_frames >= 1 && N >= 4 && n_hyp >= 1, "p3p: need n_frames >= 1, N >= 4, n_hyp >= 1");
    VO_REQUIRE(K9[0] != 0.0 && K9[4] != 0.0, "p3p: singular intrinsic matrix");
    const Intr K = make_intr(K9);
    dim3 g1(vo_div_up(n_hyp, 128), n_frames);
    p3p_solve_kernel<<<g1, 128, 0, stream>>>(d_landmarks, d_keypoints, N, K, d_sample_idx, n_hyp, d_models, d_valid);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    dim3 g2(vo_div_up(n_hyp, COUNT_WARPS), n_frames);
    p3p_count_kernel<<<g2, COUNT_WARPS * 32, 0, stream>>>(d_landmarks, d_keypoints, N, K, d_models, d_valid, n_hyp,
                                                          threshold, inclusive, d_counts);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    return VO_OK;
}

int vo_launch_p3p_select(vo_ctx* ctx, const double* d_landmarks, const double* d_keypoints, int n_frames, int N,
                         const double* K9, const double* d_models, const unsigned char* d_valid, const int* d_counts,
                         int n_hyp, double threshold, int inclusive, const int* d_iters_for_count, int initial_iters, int start_n,
                         int start_best, int* d_best4, int* d_consumed, int* d_iters_out, unsigned char* d_inliers,
                         double* d_best_model, cudaStream_t stream) {
    VO_REQUIRE(n_frames >= 1 && N >= 4 && n_hyp >= 1, "p3p select: bad sizes");
    const Intr K = make_intr(K9);
    p3p_select_kernel<<<n_frames, 256, 0, stream>>>(d_landmarks, d_keypoints, N, K, d_models, d_valid, d_counts, n_hyp,
                                                    threshold, inclusive, d_iters_for_count, initial_iters, start_n, start_best,
                                                    d_best4, d_consumed, d_iters_out, d_inliers, d_best_model);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    return VO_OK;
}
