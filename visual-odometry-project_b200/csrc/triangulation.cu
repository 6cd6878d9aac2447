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
#include "tri_device.cuh"

namespace {

using namespace tridev;

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
    double x[4];
    triangulate_point(c1, c2, x1, y1, x2, y2, mode, x);
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
