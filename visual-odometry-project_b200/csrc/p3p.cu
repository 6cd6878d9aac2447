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
#include "p3p_device.cuh"

namespace {
using namespace p3pdev;

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
    double bm[12];
    const bool found = idx_ok && solve4(X4, uv4, K, bm);
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
                if (i0 < n) in0 = reproj_inlier(m, sX[i0], sY[i0], sZ[i0], sU[i0], sV[i0], K, threshold, inclusive);
                if (i1 < n) in1 = reproj_inlier(m, sX[i1], sY[i1], sZ[i1], sU[i1], sV[i1], K, threshold, inclusive);
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
        in[i] = reproj_inlier(m, L[3 * i], L[3 * i + 1], L[3 * i + 2], P[2 * i], P[2 * i + 1], K, threshold, inclusive) ? 1 : 0;
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
