// One-sided (Hestenes) Jacobi SVD of the 6x4 / 4x4 triangulation system in registers; shared by triangulation.cu and
// pipeline.cu.  See triangulation.cu for the notes.
#pragma once
#include "common.cuh"

namespace tridev {

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


// The 4x4 system of cv2.triangulatePoints (mode 1) or the reference's 6x4 cross-product system (mode 0) for one point
// pair with projections c1, c2 (row-major 3x4): homogeneous solution x[4].
__device__ __forceinline__ void triangulate_point(const double* c1, const double* c2, double x1, double y1, double x2,
                                                  double y2, int mode, double* x) {
    double a[6][4];
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
}

}  // namespace tridev
