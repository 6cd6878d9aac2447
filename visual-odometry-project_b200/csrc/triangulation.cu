// Linear (DLT) triangulation for sm_100a: one point per thread, one-sided Jacobi SVD in registers.
//
// Replaces  /root/reference/src/vo/landmarks/triangulation.py:352-389  _linear_triangulation
//           /root/reference/src/vo/landmarks/triangulation.py:38-86    triangulate_candidates
//
// mode 0 builds the reference's own 6x4 system  A = [[p1]x C1 ; [p2]x C2]  (triangulation.py:379-382);
// mode 1 builds the 4x4 system of cv2.triangulatePoints (x*P[2]-P[0], y*P[2]-P[1] per view), which
// triangulate_candidates uses when use_opencv=True (triangulation.py:59-74).  The reference takes
// the last right singular vector from LAPACK; here the columns of A are orthogonalised by Hestenes
// rotations (no A^T A, so the conditioning is that of A) and the column of V belonging to the
// smallest column norm is dehomogenised (helpers.py:19-29).  float64 throughout.
#include "common.cuh"

namespace {

constexpr int TRI_MAX_SWEEPS = 20;

template <int ROWS>
__device__ __forceinline__ void jacobi_null_vector(double (&a)[6][4], double* x) {
    double v[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) v[i][j] = (i == j) ? 1.0 : 0.0;

    for (int sweep = 0; sweep < TRI_MAX_SWEEPS; sweep++) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < 3; p++) {
#pragma unroll
            for (int q = p + 1; q < 4; q++) {
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
                for (int i = 0; i < ROWS; i++) {
                    alpha += a[i][p] * a[i][p];
                    beta += a[i][q] * a[i][q];
                    gamma += a[i][p] * a[i][q];
                }
                if (fabs(gamma) > 1e-15 * sqrt(alpha * beta) && gamma != 0.0) {
                    rotated = true;
                    const double zeta = (beta - alpha) / (2.0 * gamma);
                    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                    for (int i = 0; i < ROWS; i++) {
                        const double tp = a[i][p], tq = a[i][q];
                        a[i][p] = c * tp - s * tq;
                        a[i][q] = s * tp + c * tq;
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const double tp = v[i][p], tq = v[i][q];
                        v[i][p] = c * tp - s * tq;
                        v[i][q] = s * tp + c * tq;
                    }
                }
            }
        }
        if (!rotated) break;
    }
    double nmin = 0.0;
    int jmin = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        double n = 0.0;
#pragma unroll
        for (int i = 0; i < ROWS; i++) n += a[i][j] * a[i][j];
        if (j == 0 || n < nmin) { nmin = n; jmin = j; }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double xi = v[i][0];
        if (jmin == 1) xi = v[i][1];
        if (jmin == 2) xi = v[i][2];
        if (jmin == 3) xi = v[i][3];
        x[i] = xi;
    }
}

__global__ void __launch_bounds__(128)
triangulate_kernel(const double* __restrict__ p1, const double* __restrict__ p2, int n,
                   const double* __restrict__ proj1, int per_point, const double* __restrict__ proj2, int group,
                   int mode, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double c1[12], c2[12];
    const double* P1 = proj1 + (per_point ? (size_t)i * 12 : 0);
    const double* P2 = proj2 + (group > 0 ? (size_t)(i / group) * 12 : 0);  // one end pose per sequence
#pragma unroll
    for (int k = 0; k < 12; k++) { c1[k] = P1[k]; c2[k] = P2[k]; }
    const double x1 = p1[2 * i], y1 = p1[2 * i + 1], x2 = p2[2 * i], y2 = p2[2 * i + 1];
    double a[6][4];
    double x[4];
    if (mode == 0) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            a[0][c] = -c1[4 + c] + y1 * c1[8 + c];
            a[1][c] = c1[c] - x1 * c1[8 + c];
            a[2][c] = -y1 * c1[c] + x1 * c1[4 + c];
            a[3][c] = -c2[4 + c] + y2 * c2[8 + c];
            a[4][c] = c2[c] - x2 * c2[8 + c];
            a[5][c] = -y2 * c2[c] + x2 * c2[4 + c];
        }
        jacobi_null_vector<6>(a, x);
    } else {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            a[0][c] = x1 * c1[8 + c] - c1[c];
            a[1][c] = y1 * c1[8 + c] - c1[4 + c];
            a[2][c] = x2 * c2[8 + c] - c2[c];
            a[3][c] = y2 * c2[8 + c] - c2[4 + c];
            a[4][c] = 0.0;
            a[5][c] = 0.0;
        }
        jacobi_null_vector<4>(a, x);
    }
    out[3 * i] = x[0] / x[3];
    out[3 * i + 1] = x[1] / x[3];
    out[3 * i + 2] = x[2] / x[3];
}

}  // namespace

int vo_launch_triangulate(vo_ctx* ctx, const double* d_p1, const double* d_p2, int n, const double* d_proj1,
                          int proj1_per_point, const double* d_proj2, int proj2_group, int mode, double* d_out,
                          cudaStream_t stream) {
    VO_REQUIRE(n >= 0 && (mode == 0 || mode == 1), "triangulate: bad n / mode");
    if (n == 0) return VO_OK;
    triangulate_kernel<<<vo_div_up(n, 128), 128, 0, stream>>>(d_p1, d_p2, n, d_proj1, proj1_per_point, d_proj2, proj2_group,
                                                               mode, d_out);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    return VO_OK;
}
