// Two-view bootstrap on the GPU: fundamental matrix by RANSAC over the 7-point solver, essential-matrix
// decomposition, cheirality vote and the landmarks of all matches -- one CTA per sequence.
//
// Replaces  /root/reference/src/vo/landmarks/triangulation.py:88-108   (triangulate_matches)
//           /root/reference/src/vo/landmarks/triangulation.py:110-134  (_find_fundamental_matrix_ransac, use_opencv=True:
//                                                                       cv2.findFundamentalMat(FM_RANSAC))
//           /root/reference/src/vo/landmarks/triangulation.py:224-350  (_find_essential_matrix, _decompose_essential_matrix,
//                                                                       _find_relative_pose)
// as src/main.py:185-222 configures them (use_ransac=True, use_opencv=True, threshold 0.25, confidence 0.999).
//
// cv2.findFundamentalMat is restated (OpenCV calib3d fundam.cpp / ptsetreg.cpp): float32 points, cv::RNG(2^64-1)
// subsets of 7 (redraw on repetition; the whole subset is redrawn when its last point is collinear with two earlier
// ones in either image), run7Point (normalised points, null space of the 7x9 system, cubic in lambda by solveCubic's
// closed forms, F33 = 1), float32 epipolar errors against (float)thr^2, a model accepted when its count exceeds
// max(best, 6), RANSACUpdateNumIters after every improvement.  The serial parts of
// OpenCV's loop (the generator and the accept / update rule) run on thread 0; a round of BS_ROUND samples is solved
// by 32 threads spread over the warps, and every (sample, root) model is scored by a whole warp with ballot-free
// shuffle sums.  Float64 throughout, no FMA contraction (build flag).
#include "common.cuh"
#include "launchers.cuh"
#include "tri_device.cuh"

namespace {

constexpr int BS_THREADS = 256, BS_WARPS = BS_THREADS / 32;
constexpr int BS_ROUND = 32;                  // samples per round (one solver thread each)
constexpr int BS_MAX_N = 8192;
constexpr float BS_FLT_EPS = 1.1920929e-07f;
constexpr double BS_DBL_EPS = 2.220446049250313e-16;

struct BootArgs {
    const double* p1;       // [S][N][2] float64 pixels of frame 1 (as the reference passes them)
    const double* p2;       // [S][N][2]
    const int* n_pts;       // [S] points in use per sequence, or null: N
    int N;                  // row stride
    double K[9];
    double thr, conf;
    int max_iters;
    double* F_out;          // [S][9]
    double* M_out;          // [S][12]  [R | t], frame 1 -> frame 2
    double* land;           // [S][N][3] landmarks in frame-1 coordinates (all points)
    unsigned char* mask;    // [S][N]   F inlier and in front of both cameras
    unsigned char* f_mask;  // [S][N]   F inlier (optional)
    int* info;              // [S][4]   1 = model found, RANSAC iterations, F inliers, cheirality-valid inliers of the winner
};

__device__ __forceinline__ unsigned int bs_rng_next(unsigned long long& st) {   // cv::RNG (multiply with carry)
    st = (unsigned long long)(unsigned int)st * 4164903690ull + (st >> 32);
    return (unsigned int)st;
}

__device__ __forceinline__ int bs_update_num_iters(double p, double ep, int model_points, int max_iters) {   // RANSACUpdateNumIters
    p = fmin(fmax(p, 0.0), 1.0); ep = fmin(fmax(ep, 0.0), 1.0);
    const double tiny = 2.2250738585072014e-308;
    double num = fmax(1.0 - p, tiny);
    double denom = 1.0 - pow(1.0 - ep, (double)model_points);
    if (denom < tiny) return 0;
    num = log(num); denom = log(denom);
    return (denom >= 0.0 || -num >= (double)max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

// haveCollinearPoints: the last of the 7 points against every pair of earlier ones
__device__ bool bs_collinear_last(const float2* m, const int* idx) {
    const float2 pi = m[idx[6]];
    for (int j = 0; j < 6; j++) {
        const float2 pj = m[idx[j]];
        const double dx1 = (double)pj.x - (double)pi.x, dy1 = (double)pj.y - (double)pi.y;
        for (int k = 0; k < j; k++) {
            const float2 pk = m[idx[k]];
            const double dx2 = (double)pk.x - (double)pi.x, dy2 = (double)pk.y - (double)pi.y;
            if (fabs(dx2 * dy1 - dy2 * dx1) <= (double)BS_FLT_EPS * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return true;
        }
    }
    return false;
}

// cv::solveCubic, real roots in OpenCV's order
__device__ int bs_solve_cubic(double a0, double a1, double a2, double a3, double* x) {
    const double PI = 3.14159265358979323846;
    if (a0 == 0.0) {
        if (a1 == 0.0) {
            if (a2 == 0.0) return 0;
            x[0] = -a3 / a2;
            return 1;
        }
        double d = a2 * a2 - 4 * a1 * a3;
        if (d < 0) return 0;
        d = sqrt(d);
        const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
        if (fabs(q1) > fabs(q2)) { x[0] = q1 / a1; x[1] = a3 / q1; }
        else { x[0] = q2 / a1; x[1] = a3 / q2; }
        return d > 0 ? 2 : 1;
    }
    a0 = 1.0 / a0; a1 *= a0; a2 *= a0; a3 *= a0;
    const double Q = (a1 * a1 - 3 * a2) * (1.0 / 9);
    const double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1.0 / 54);
    const double Qc = Q * Q * Q;
    double d = Qc - R * R;
    if (d > 0) {
        const double theta = acos(R / sqrt(Qc));
        const double t0 = -2 * sqrt(Q), t1 = theta * (1.0 / 3), t2 = a1 * (1.0 / 3);
        x[0] = t0 * cos(t1) - t2;
        x[1] = t0 * cos(t1 + (2.0 * PI / 3)) - t2;
        x[2] = t0 * cos(t1 + (4.0 * PI / 3)) - t2;
        return 3;
    }
    if (d == 0) {
        if (R >= 0) { x[0] = -2 * cbrt(R) - a1 / 3; x[1] = cbrt(R) - a1 / 3; }
        else { x[0] = 2 * cbrt(-R) - a1 / 3; x[1] = -cbrt(-R) - a1 / 3; }
        return x[0] == x[1] ? 1 : 2;
    }
    d = sqrt(-d);
    double e = cbrt(d + fabs(R));
    if (R > 0) e = -e;
    x[0] = (e + Q / e) - a1 * (1.0 / 3);
    return 1;
}

__device__ __forceinline__ double bs_det3(const double* r0, const double* r1, const double* r2) {
    return r0[0] * (r1[1] * r2[2] - r1[2] * r2[1]) - r0[1] * (r1[0] * r2[2] - r1[2] * r2[0]) + r0[2] * (r1[0] * r2[1] - r1[1] * r2[0]);
}

// run7Point: up to three fundamental matrices (row-major) for 7 correspondences; returns their number.
// The two-dimensional null space of the 7x9 system comes from Gauss-Jordan elimination with complete pivoting (the
// solutions do not depend on the basis: they are the singular members of the pencil it spans).
__device__ int bs_seven_point(const float2* m1, const float2* m2, const int* idx, double* Fout /*[3][9]*/) {
    double x1[7], y1[7], x2[7], y2[7];
    double c1x = 0, c1y = 0, c2x = 0, c2y = 0;
    for (int i = 0; i < 7; i++) {
        const float2 a = m1[idx[i]], b = m2[idx[i]];
        x1[i] = a.x; y1[i] = a.y; x2[i] = b.x; y2[i] = b.y;
        c1x += x1[i]; c1y += y1[i]; c2x += x2[i]; c2y += y2[i];
    }
    const double t = 1.0 / 7;
    c1x *= t; c1y *= t; c2x *= t; c2y *= t;
    double s1 = 0, s2 = 0;
    for (int i = 0; i < 7; i++) {
        s1 += sqrt((x1[i] - c1x) * (x1[i] - c1x) + (y1[i] - c1y) * (y1[i] - c1y));
        s2 += sqrt((x2[i] - c2x) * (x2[i] - c2x) + (y2[i] - c2y) * (y2[i] - c2y));
    }
    s1 *= t; s2 *= t;
    if (s1 < (double)BS_FLT_EPS || s2 < (double)BS_FLT_EPS) return 0;
    s1 = sqrt(2.0) / s1; s2 = sqrt(2.0) / s2;
    double A[7][9];
    for (int i = 0; i < 7; i++) {
        const double u0 = (x1[i] - c1x) * s1, v0 = (y1[i] - c1y) * s1, u1 = (x2[i] - c2x) * s2, v1 = (y2[i] - c2y) * s2;
        A[i][0] = u1 * u0; A[i][1] = u1 * v0; A[i][2] = u1; A[i][3] = v1 * u0; A[i][4] = v1 * v0; A[i][5] = v1;
        A[i][6] = u0; A[i][7] = v0; A[i][8] = 1.0;
    }
    int perm[9];
    for (int j = 0; j < 9; j++) perm[j] = j;
    for (int k = 0; k < 7; k++) {
        int pi = k, pj = k;
        double best = -1.0;
        for (int i = k; i < 7; i++)
            for (int j = k; j < 9; j++) {
                const double v = fabs(A[i][j]);
                if (v > best) { best = v; pi = i; pj = j; }
            }
        if (!(best > 0.0)) return 0;                       // rank below 7: no isolated solutions
        if (pi != k) for (int j = 0; j < 9; j++) { const double v = A[k][j]; A[k][j] = A[pi][j]; A[pi][j] = v; }
        if (pj != k) {
            for (int i = 0; i < 7; i++) { const double v = A[i][k]; A[i][k] = A[i][pj]; A[i][pj] = v; }
            const int v = perm[k]; perm[k] = perm[pj]; perm[pj] = v;
        }
        const double inv = 1.0 / A[k][k];
        for (int j = k; j < 9; j++) A[k][j] *= inv;
        for (int i = 0; i < 7; i++) {
            if (i == k) continue;
            const double f = A[i][k];
            if (f != 0.0) for (int j = k; j < 9; j++) A[i][j] -= f * A[k][j];
        }
    }
    double f1[9], f2[9];
    for (int k = 0; k < 7; k++) { f1[perm[k]] = -A[k][7]; f2[perm[k]] = -A[k][8]; }
    f1[perm[7]] = 1.0; f1[perm[8]] = 0.0; f2[perm[7]] = 0.0; f2[perm[8]] = 1.0;
    {   // orthonormal basis (keeps the cubic well scaled)
        double n = 0, d = 0;
        for (int j = 0; j < 9; j++) n += f1[j] * f1[j];
        n = 1.0 / sqrt(n);
        for (int j = 0; j < 9; j++) f1[j] *= n;
        for (int j = 0; j < 9; j++) d += f1[j] * f2[j];
        n = 0;
        for (int j = 0; j < 9; j++) { f2[j] -= d * f1[j]; n += f2[j] * f2[j]; }
        n = 1.0 / sqrt(n);
        for (int j = 0; j < 9; j++) f2[j] *= n;
    }
    for (int j = 0; j < 9; j++) f1[j] -= f2[j];            // F ~ lambda f1 + f2
    const double c0 = bs_det3(f1, f1 + 3, f1 + 6);
    const double c1 = bs_det3(f2, f1 + 3, f1 + 6) + bs_det3(f1, f2 + 3, f1 + 6) + bs_det3(f1, f1 + 3, f2 + 6);
    const double c2 = bs_det3(f1, f2 + 3, f2 + 6) + bs_det3(f2, f1 + 3, f2 + 6) + bs_det3(f2, f2 + 3, f1 + 6);
    const double c3 = bs_det3(f2, f2 + 3, f2 + 6);
    double roots[3];
    const int n = bs_solve_cubic(c0, c1, c2, c3, roots);
    if (n < 1 || n > 3) return 0;
    for (int k = 0; k < n; k++) {
        double lam = roots[k], mu = 1.0, G[9];
        const double s = f1[8] * lam + f2[8];
        if (fabs(s) > BS_DBL_EPS) { mu = 1.0 / s; lam *= mu; G[8] = 1.0; }
        else G[8] = 0.0;
        for (int j = 0; j < 8; j++) G[j] = f1[j] * lam + f2[j] * mu;
        // F = T2^T G T1 with T = [s 0 -s cx; 0 s -s cy; 0 0 1]
        double H[9];                                       // G T1
        for (int r = 0; r < 3; r++) {
            H[3 * r] = G[3 * r] * s1; H[3 * r + 1] = G[3 * r + 1] * s1;
            H[3 * r + 2] = G[3 * r] * (-s1 * c1x) + G[3 * r + 1] * (-s1 * c1y) + G[3 * r + 2];
        }
        double* F = Fout + 9 * k;
        for (int c = 0; c < 3; c++) {
            F[c] = s2 * H[c]; F[3 + c] = s2 * H[3 + c];
            F[6 + c] = (-s2 * c2x) * H[c] + (-s2 * c2y) * H[3 + c] + H[6 + c];
        }
        if (fabs(F[8]) > (double)BS_FLT_EPS) { const double inv = 1.0 / F[8]; for (int j = 0; j < 9; j++) F[j] *= inv; }
    }
    return n;
}

// FMEstimatorCallback::computeError for one pair: max of the two squared point-line distances, as float32
__device__ __forceinline__ float bs_fm_error(const double* F, float2 a, float2 b) {
    const double ax = a.x, ay = a.y, bx = b.x, by = b.y;
    double p = F[0] * ax + F[1] * ay + F[2], q = F[3] * ax + F[4] * ay + F[5], r = F[6] * ax + F[7] * ay + F[8];
    const double s2 = 1.0 / (p * p + q * q), d2 = bx * p + by * q + r;
    p = F[0] * bx + F[3] * by + F[6]; q = F[1] * bx + F[4] * by + F[7]; r = F[2] * bx + F[5] * by + F[8];
    const double s1 = 1.0 / (p * p + q * q), d1 = ax * p + ay * q + r;
    const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
    return (float)((e1 < e2) ? e2 : e1);   // std::max(e1, e2)
}

// symmetric 3x3 eigen decomposition by cyclic Jacobi: A = V diag(w) V^T
__device__ void bs_eig3(double A[3][3], double V[3][3], double* w) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; sweep++) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off <= 1e-300 || off <= 1e-17 * (fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]))) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (A[p][q] == 0.0) continue;
                const double zeta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < 3; k++) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
                for (int k = 0; k < 3; k++) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
                for (int k = 0; k < 3; k++) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
            }
    }
    for (int i = 0; i < 3; i++) w[i] = A[i][i];
}

// triangulation.py:245-277: the four [R | t] candidates of E (row-major 3x4 each), in the reference's order
// idx = 2 i + j -> [R_j | (-1)^i t].  U, V of the SVD come from the eigenvectors of E^T E (the set of candidates does not
// depend on the sign / rotation freedom of the SVD; their order matters only for exact ties in the vote).
__device__ void bs_decompose_essential(const double* E, double* M4) {
    double B[3][3], V[3][3], w[3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) B[i][j] = E[i] * E[j] + E[3 + i] * E[3 + j] + E[6 + i] * E[6 + j];
    bs_eig3(B, V, w);
    int o[3] = {0, 1, 2};
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2 - a; b++) if (w[o[b]] < w[o[b + 1]]) { const int v = o[b]; o[b] = o[b + 1]; o[b + 1] = v; }
    double v[3][3], u[3][3];                               // v[k] = k-th right singular vector, u[k] = left
    for (int k = 0; k < 3; k++) for (int i = 0; i < 3; i++) v[k][i] = V[i][o[k]];
    for (int k = 0; k < 2; k++) {
        double n = 0;
        for (int i = 0; i < 3; i++) { u[k][i] = E[3 * i] * v[k][0] + E[3 * i + 1] * v[k][1] + E[3 * i + 2] * v[k][2]; n += u[k][i] * u[k][i]; }
        n = 1.0 / sqrt(n);
        for (int i = 0; i < 3; i++) u[k][i] *= n;
    }
    {   // u1 orthogonal to u0 exactly, u2 = u0 x u1
        double d = 0, n = 0;
        for (int i = 0; i < 3; i++) d += u[0][i] * u[1][i];
        for (int i = 0; i < 3; i++) { u[1][i] -= d * u[0][i]; n += u[1][i] * u[1][i]; }
        n = 1.0 / sqrt(n);
        for (int i = 0; i < 3; i++) u[1][i] *= n;
        u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1];
        u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2];
        u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
    }
    // R0 = U W V^T = u1 v0^T - u0 v1^T + u2 v2^T,  R1 = U W^T V^T = -u1 v0^T + u0 v1^T + u2 v2^T
    double R[2][9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const double a = u[1][i] * v[0][j] - u[0][i] * v[1][j], b = u[2][i] * v[2][j];
            R[0][3 * i + j] = a + b; R[1][3 * i + j] = -a + b;
        }
    for (int k = 0; k < 2; k++)
        if (bs_det3(R[k], R[k] + 3, R[k] + 6) < 0) for (int j = 0; j < 9; j++) R[k][j] = -R[k][j];
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) {
            double* M = M4 + 12 * (2 * i + j);
            for (int r = 0; r < 3; r++) {
                M[4 * r] = R[j][3 * r]; M[4 * r + 1] = R[j][3 * r + 1]; M[4 * r + 2] = R[j][3 * r + 2];
                M[4 * r + 3] = i ? -u[2][r] : u[2][r];
            }
        }
}

struct BootShared {
    double models[BS_ROUND * 3 * 9];
    double bestF[9];
    double M4[48];          // the four candidates
    double P1[12], P2[48];  // K [I | 0], K M_m
    int idx[BS_ROUND][7];
    int nmod[BS_ROUND];
    int counts[BS_ROUND * 3];
    int votes[4];
    int n_gen, done, have, it, niters, best, gen_failed, winner, n_f_inl;
};

__global__ void __launch_bounds__(BS_THREADS)
bootstrap_kernel(BootArgs a) {
    extern __shared__ __align__(16) unsigned char bs_dyn[];
    __shared__ BootShared sh;
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.n_pts ? min(a.n_pts[s], a.N) : a.N;
    float2* m1 = reinterpret_cast<float2*>(bs_dyn);
    float2* m2 = m1 + a.N;
    unsigned char* finl = reinterpret_cast<unsigned char*>(m2 + a.N);
    const double* g1 = a.p1 + (size_t)s * a.N * 2;
    const double* g2 = a.p2 + (size_t)s * a.N * 2;
    for (int i = tid; i < N; i += BS_THREADS) {
        m1[i] = make_float2((float)g1[2 * i], (float)g1[2 * i + 1]);
        m2[i] = make_float2((float)g2[2 * i], (float)g2[2 * i + 1]);
    }
    double thr_d = a.thr, conf = a.conf;
    if (thr_d <= 0) thr_d = 3.0;
    if (conf < BS_DBL_EPS || conf > 1 - BS_DBL_EPS) conf = 0.99;
    const float thr = (float)(thr_d * thr_d);
    if (tid == 0) { sh.done = N < 15 ? 1 : 0; sh.have = 0; sh.it = 0; sh.niters = a.max_iters; sh.best = 0; sh.gen_failed = 0; sh.n_f_inl = 0; }
    unsigned long long rng = 0xFFFFFFFFFFFFFFFFull;      // thread 0 only
    __syncthreads();
    while (!sh.done) {
        if (tid == 0) {                                    // getSubset x BS_ROUND (serial: one generator)
            int g = 0;
            for (; g < BS_ROUND; g++) {
                int* id = sh.idx[g];
                int attempt = 0;
                for (; attempt < 10000; attempt++) {
                    for (int i = 0; i < 7; i++) {
                        for (;;) {
                            const int v = (int)(bs_rng_next(rng) % (unsigned int)N);
                            bool dup = false;
                            for (int j = 0; j < i; j++) dup |= (id[j] == v);
                            if (!dup) { id[i] = v; break; }
                        }
                    }
                    if (!bs_collinear_last(m1, id) && !bs_collinear_last(m2, id)) break;
                }
                if (attempt == 10000) { sh.gen_failed = 1; break; }
            }
            sh.n_gen = g;
        }
        __syncthreads();
        const int n_gen = sh.n_gen;
        if ((tid & 7) == 0 && (tid >> 3) < n_gen) {       // 32 solver threads, four per warp
            const int g = tid >> 3;
            sh.nmod[g] = bs_seven_point(m1, m2, sh.idx[g], sh.models + g * 27);
        }
        __syncthreads();
        for (int p = warp; p < n_gen * 3; p += BS_WARPS) {   // one warp per (sample, root)
            const int g = p / 3, m = p - 3 * g;
            if (m >= sh.nmod[g]) continue;
            double F[9];
#pragma unroll
            for (int j = 0; j < 9; j++) F[j] = sh.models[p * 9 + j];
            int cnt = 0;
            for (int i = lane; i < N; i += 32) cnt += (bs_fm_error(F, m1[i], m2[i]) <= thr) ? 1 : 0;
            cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
            if (lane == 0) sh.counts[p] = cnt;
        }
        __syncthreads();
        if (tid == 0) {                                    // OpenCV's loop body, sample by sample
            int it = sh.it, niters = sh.niters, best = sh.best, done = 0;
            for (int g = 0; g < n_gen; g++) {
                if (it >= niters) { done = 1; break; }
                it++;
                for (int m = 0; m < sh.nmod[g]; m++) {
                    const int good = sh.counts[3 * g + m];
                    if (good > max(best, 6)) {
                        best = good;
                        for (int j = 0; j < 9; j++) sh.bestF[j] = sh.models[(3 * g + m) * 9 + j];
                        sh.have = 1;
                        niters = bs_update_num_iters(conf, (double)(N - good) / (double)N, 7, niters);
                    }
                }
            }
            if (it >= niters || sh.gen_failed) done = 1;
            sh.it = it; sh.niters = niters; sh.best = best; sh.done = done;
        }
        __syncthreads();
    }
    int* info = a.info + 4 * s;
    double* land = a.land + (size_t)s * a.N * 3;
    unsigned char* mask = a.mask + (size_t)s * a.N;
    if (!sh.have) {                                        // no model: everything empty
        for (int i = tid; i < a.N; i += BS_THREADS) {
            mask[i] = 0;
            if (a.f_mask) a.f_mask[(size_t)s * a.N + i] = 0;
            land[3 * i] = land[3 * i + 1] = land[3 * i + 2] = __longlong_as_double(0x7ff8000000000000ll);
        }
        if (tid < 9) a.F_out[9 * s + tid] = 0.0;
        if (tid < 12) a.M_out[12 * s + tid] = 0.0;
        if (tid == 0) { info[0] = 0; info[1] = sh.it; info[2] = 0; info[3] = 0; }
        return;
    }
    {   // inliers of the winner
        double F[9];
#pragma unroll
        for (int j = 0; j < 9; j++) F[j] = sh.bestF[j];
        int cnt = 0;
        for (int i = tid; i < N; i += BS_THREADS) {
            const unsigned char f = (bs_fm_error(F, m1[i], m2[i]) <= thr) ? 1 : 0;
            finl[i] = f; cnt += f;
            if (a.f_mask) a.f_mask[(size_t)s * a.N + i] = f;
        }
        cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
        if (lane == 0 && cnt) atomicAdd(&sh.n_f_inl, cnt);
    }
    if (tid == 0) {                                        // E = K^T F K, the four candidates and their projections
        double E[9], T[9];
        const double* K = a.K;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T[3 * i + j] = sh.bestF[3 * i] * K[j] + sh.bestF[3 * i + 1] * K[3 + j] + sh.bestF[3 * i + 2] * K[6 + j];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) E[3 * i + j] = K[i] * T[j] + K[3 + i] * T[3 + j] + K[6 + i] * T[6 + j];
        bs_decompose_essential(E, sh.M4);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) sh.P1[4 * i + j] = j < 3 ? K[3 * i + j] : 0.0;
        for (int m = 0; m < 4; m++)
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 4; j++)
                    sh.P2[12 * m + 4 * i + j] = K[3 * i] * sh.M4[12 * m + j] + K[3 * i + 1] * sh.M4[12 * m + 4 + j] + K[3 * i + 2] * sh.M4[12 * m + 8 + j];
        for (int m = 0; m < 4; m++) sh.votes[m] = 0;
    }
    __syncthreads();
    // cheirality vote (triangulation.py:306-328): inliers in front of both cameras, per candidate
    for (int w = tid; w < 4 * N; w += BS_THREADS) {
        const int m = w / N, i = w - m * N;
        if (!finl[i]) continue;
        double x[4];
        tridev::triangulate_point(sh.P1, sh.P2 + 12 * m, g1[2 * i], g1[2 * i + 1], g2[2 * i], g2[2 * i + 1], 0, x);
        const double X = x[0] / x[3], Y = x[1] / x[3], Z = x[2] / x[3];
        const double* M = sh.M4 + 12 * m;
        const double Z2 = (M[8] * X + M[9] * Y + M[10] * Z) + M[11];
        if (Z >= 0.0 && Z2 >= 0.0) atomicAdd(&sh.votes[m], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int bv = -1, bm = 0;
        for (int m = 0; m < 4; m++) if (sh.votes[m] > bv) { bv = sh.votes[m]; bm = m; }
        sh.winner = bm;
        info[0] = 1; info[1] = sh.it; info[2] = sh.n_f_inl; info[3] = bv;
    }
    __syncthreads();
    const int wm = sh.winner;
    if (tid < 9) a.F_out[9 * s + tid] = sh.bestF[tid];
    if (tid < 12) a.M_out[12 * s + tid] = sh.M4[12 * wm + tid];
    for (int i = tid; i < a.N; i += BS_THREADS) {          // landmarks of every match (triangulation.py:331-336)
        if (i >= N) { mask[i] = 0; land[3 * i] = land[3 * i + 1] = land[3 * i + 2] = __longlong_as_double(0x7ff8000000000000ll); continue; }
        double x[4];
        tridev::triangulate_point(sh.P1, sh.P2 + 12 * wm, g1[2 * i], g1[2 * i + 1], g2[2 * i], g2[2 * i + 1], 0, x);
        const double X = x[0] / x[3], Y = x[1] / x[3], Z = x[2] / x[3];
        const double* M = sh.M4 + 12 * wm;
        const double Z2 = (M[8] * X + M[9] * Y + M[10] * Z) + M[11];
        land[3 * i] = X; land[3 * i + 1] = Y; land[3 * i + 2] = Z;
        mask[i] = (finl[i] && Z >= 0.0 && Z2 >= 0.0) ? 1 : 0;
    }
}

}  // namespace

int vo_launch_bootstrap(vo_ctx* ctx, const double* d_p1, const double* d_p2, int n_seq, int N, const int* d_n_pts,
                        const double* K9, double threshold, double confidence, int max_iters, double* d_F, double* d_M,
                        double* d_landmarks, unsigned char* d_mask, unsigned char* d_f_mask, int* d_info, cudaStream_t stream) {
    VO_REQUIRE(n_seq >= 1 && N >= 1 && N <= BS_MAX_N, "bootstrap: n_seq >= 1 and 1 <= N <= %d (got %d, %d)", BS_MAX_N, n_seq, N);
    VO_REQUIRE(max_iters >= 1, "bootstrap: max_iters must be positive");
    BootArgs a;
    a.p1 = d_p1; a.p2 = d_p2; a.n_pts = d_n_pts; a.N = N;
    for (int i = 0; i < 9; i++) a.K[i] = K9[i];
    a.thr = threshold; a.conf = confidence; a.max_iters = max_iters;
    a.F_out = d_F; a.M_out = d_M; a.land = d_landmarks; a.mask = d_mask; a.f_mask = d_f_mask; a.info = d_info;
    const size_t smem = (size_t)N * 17 + 16;
    if (vo_ctx_once(ctx, VO_ATTR_BOOT))
        VO_CUDA(cudaFuncSetAttribute(bootstrap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BS_MAX_N * 17 + 16));
    bootstrap_kernel<<<n_seq, BS_THREADS, smem, stream>>>(a);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    return VO_OK;
}
