// The chained, device-resident VO pipeline: per-sequence feature tables live in HBM and one step advances every
// sequence by one frame with the reference's data flow; only the frame goes up, only the pose and a few counters
// come back.
//
// Replaces (for S independent sequences at once, KLT tracker mode)
//   /root/reference/src/main.py:248-287                  loop body
//   /root/reference/src/vo/features/klt.py:191-280       track_features: re-detection rule, LK, status / error filter
//   /root/reference/src/vo/features/klt.py:117-189       update_features (appending fresh corners)
//   /root/reference/src/vo/primitives/matches.py:10-212  Matches: stable regrouping [triangulated | matched | new]
//   /root/reference/src/vo/algorithms/ransac.py:69-129   adaptive RANSAC, numpy Generator(PCG64(2023)).choice stream
//   /root/reference/src/vo/pose_estimation/p3p.py:123-213 estimate_pose + nonlinear refinement
//   /root/reference/src/vo/primitives/state.py:39-229    world pose, reset_outliers, compute_candidates (bearing angle),
//                                                        update_with_world_landmarks, _check_landmarks
//   /root/reference/src/vo/landmarks/triangulation.py:38-86 triangulate_candidates
//
// Kernels (one CTA per sequence unless noted; all of them are a few microseconds of latency-bound work that runs under
// the tracker and the detector of other sequences):
//   pipe_append_kernel    conditional append of the detector's corners (n < 0.8 * num_features)
//   klt_track_packed      (klt.cu) one warp per table row, rows beyond the sequence's count exit at once
//   pipe_regroup_kernel   keep = status && err < 100; stable 3-way partition by the old state into the other table
//   pipe_pose_kernel      RANSAC: thread 0 draws 16 samples from the sequence's PCG64 stream (bit-exact numpy
//                         Generator.choice), 16 threads solve P3P, 16 warps count inliers (ballot/popc), thread 0
//                         replays ransac.py's loop and rewinds the stream to where the reference's would be; then the
//                         winner's inlier mask and a damped Gauss-Newton refinement on SE(3) (block reductions)
//   pipe_update_kernel    pose inverse, reset_outliers, bearing-angle candidates, per-row DLT (Jacobi SVD in registers),
//                         cheirality check, per-sequence summary
// float64 for geometry (--fmad=false, same operation order as the CPU restatement the tests check it against), float32 keypoints as cv2 returns them.
#include "../../include/vo_b200.h"

#include <initializer_list>
#include <math.h>
#include <vector>

#include "common.cuh"
#include "launchers.cuh"
#include "p3p_device.cuh"
#include "tri_device.cuh"

namespace {
using p3pdev::Intr;

// ---------------------------------------------------------------------------------------------
// numpy's PCG64 (pcg64 XSL-RR 128/64, "setseq") and Generator.choice(arange(N), size=4, replace=False)
// ---------------------------------------------------------------------------------------------
struct PipeRng { unsigned long long s_hi, s_lo, inc_hi, inc_lo; unsigned int has32, u32, pad0, pad1; };

__device__ __forceinline__ unsigned long long pcg64_next64(PipeRng& r) {
    const unsigned long long M_HI = 0x2360ED051FC65DA4ull, M_LO = 0x4385DF649FCCF645ull;
    // state = state * MULT + inc (mod 2^128)
    unsigned long long lo = r.s_lo * M_LO;
    unsigned long long hi = __umul64hi(r.s_lo, M_LO) + r.s_hi * M_LO + r.s_lo * M_HI;
    const unsigned long long lo2 = lo + r.inc_lo;
    hi = hi + r.inc_hi + (lo2 < lo ? 1ull : 0ull);
    r.s_hi = hi; r.s_lo = lo2;
    const unsigned long long x = hi ^ lo2;
    const unsigned int rot = (unsigned int)(hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}
__device__ __forceinline__ unsigned int pcg64_next32(PipeRng& r) {         // numpy buffers the upper half
    if (r.has32) { r.has32 = 0; return r.u32; }
    const unsigned long long v = pcg64_next64(r);
    r.has32 = 1; r.u32 = (unsigned int)(v >> 32);
    return (unsigned int)(v & 0xffffffffull);
}
// random_bounded_uint64(off = 0, rng, mask = 0, use_masked = 0) for rng < 2^32 - 1: Lemire's method on 32-bit draws
__device__ __forceinline__ unsigned int bounded32(PipeRng& r, unsigned int rng) {
    if (rng == 0) return 0;
    const unsigned int rng_excl = rng + 1u;
    unsigned long long m = (unsigned long long)pcg64_next32(r) * rng_excl;
    unsigned int leftover = (unsigned int)(m & 0xffffffffull);
    if (leftover < rng_excl) {
        const unsigned int threshold = (0xffffffffu - rng) % rng_excl;
        while (leftover < threshold) {
            m = (unsigned long long)pcg64_next32(r) * rng_excl;
            leftover = (unsigned int)(m & 0xffffffffull);
        }
    }
    return (unsigned int)(m >> 32);
}
// The same from pre-drawn 32-bit words (one per bounded draw): false when a draw would have been rejected (probability
// ~N / 2^32 per draw) -- the caller then falls back to the sequential sampler.
__device__ __forceinline__ bool bounded32_from(unsigned int word, unsigned int rng, unsigned int& out) {
    const unsigned int rng_excl = rng + 1u;
    const unsigned long long m = (unsigned long long)word * rng_excl;
    const unsigned int leftover = (unsigned int)(m & 0xffffffffull);
    out = (unsigned int)(m >> 32);
    if (leftover < rng_excl) {
        const unsigned int threshold = (0xffffffffu - rng) % rng_excl;
        if (leftover < threshold) return false;
    }
    return true;
}
__device__ __forceinline__ bool choice4_from(const unsigned int* w, int N, int* out) {      // needs N > 4 (7 words)
    bool ok = true;
    for (int k = 0; k < 4; k++) {
        const int j = N - 4 + k;
        unsigned int v;
        ok &= bounded32_from(w[k], (unsigned int)j, v);
        bool dup = false;
        for (int q = 0; q < k; q++) dup |= (out[q] == (int)v);
        out[k] = dup ? j : (int)v;
    }
    for (int i = 3; i >= 1; i--) {
        unsigned int j;
        ok &= bounded32_from(w[7 - i], (unsigned int)i, j);
        const int tmp = out[i]; out[i] = out[j]; out[j] = tmp;
    }
    return ok;
}

// Floyd's algorithm + the final shuffle, as numpy/random/_generator.pyx does for replace=False, p=None, small sizes
__device__ __forceinline__ void choice4(PipeRng& r, int N, int* out) {
    for (int k = 0; k < 4; k++) {
        const int j = N - 4 + k;
        const int val = (int)bounded32(r, (unsigned int)j);
        bool dup = false;
        for (int q = 0; q < k; q++) dup |= (out[q] == val);
        out[k] = dup ? j : val;
    }
    for (int i = 3; i >= 1; i--) {
        const int j = (int)bounded32(r, (unsigned int)i);
        const int tmp = out[i]; out[i] = out[j]; out[j] = tmp;
    }
}

// ---------------------------------------------------------------------------------------------
// resident tables
// ---------------------------------------------------------------------------------------------
struct PipeTable {          // every array is [n_seq][capacity] rows
    float2* kp;             // keypoint in the current frame (cv2's float32)
    float2* track;          // keypoint at the start of the track
    uint8_t* state;         // 0 unmatched, 1 matched, 2 triangulated (features.py:41-43)
    double* land;           // [3] world landmark (NaN when unknown)
    double* pose;           // [12] camera-to-world pose at the start of the track, row-major 3x4
    uint8_t* cand;          // candidate flag of this frame
};

struct PipeSeq {            // per-sequence scalars
    int* n_rows;            // rows in the current table (regroup writes it, the next step's append may grow it)
    int* n_keep;            // rows after the last regroup: what pose / update of that step work on (they may still be
                            // running when the next step's append grows n_rows)
    int* n_tri;             // leading rows that were triangulated before this frame (the P3P population)
    int* num_features;      // klt.py:49 / 114  _num_features
    int* n_iterations;      // RANSAC.n_iterations, carried between frames (ransac.py:56,120)
    int* appended;          // rows appended by the last append (0 = no re-detection)
    double* c2w;            // [12] current camera-to-world pose
    double* c2w_prev;       // [12]
    double* w2c;            // [12] current world-to-camera pose (the estimator's output)
    double* w2c_prev;       // [12]
    double* p3p_model;      // [12] RANSAC winner before refinement
    PipeRng* rng;
    int* counts;            // [VO_PIPE_NCOUNTS] summary of the last step
};

struct PipeParams {
    int C;                  // capacity
    Intr K;                 // promoted intrinsics
    double K9[9], Kinv9[9];
    double thr; int inclusive; double log1mconf; int max_iter; int refine;   // inclusive != 0: cv2.solvePnPRansac's semantics
    double conf;
    double bearing_thr; int tri_mode; float err_thr; double redetect_frac;
};

__device__ __forceinline__ double nan64() { return __longlong_as_double(0x7ff8000000000000ll); }

// ---- append: klt.py:207-230 ------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pipe_append_kernel(PipeTable T, PipeSeq Q, PipeParams P, const int* __restrict__ det_xy, const float* __restrict__ det_f,
                   const int* __restrict__ det_n, int det_cap) {
    const int s = blockIdx.x;
    const int n = Q.n_rows[s];
    const bool need = (double)n < (double)Q.num_features[s] * P.redetect_frac;
    if (!need) { if (threadIdx.x == 0) Q.appended[s] = 0; return; }
    int m = det_n[s];
    const int room = P.C - n;
    const bool overflow = m > room;
    if (overflow) m = room;
    const size_t base = (size_t)s * P.C;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        float x, y;
        if (det_f) { x = det_f[((size_t)s * det_cap + i) * 2]; y = det_f[((size_t)s * det_cap + i) * 2 + 1]; }
        else { x = (float)det_xy[((size_t)s * det_cap + i) * 2]; y = (float)det_xy[((size_t)s * det_cap + i) * 2 + 1]; }
        const size_t r = base + n + i;
        T.kp[r] = make_float2(x, y);
        T.track[r] = make_float2(x, y);
        T.state[r] = 0;
        T.cand[r] = 0;
        for (int k = 0; k < 3; k++) T.land[r * 3 + k] = nan64();
        double* p = T.pose + r * 12;                       // klt.py:161-166: np.eye(4)
        for (int k = 0; k < 12; k++) p[k] = (k == 0 || k == 5 || k == 10) ? 1.0 : 0.0;
    }
    if (threadIdx.x == 0) {
        Q.n_rows[s] = n + m;
        Q.num_features[s] = det_n[s];                      // klt.py:114 (find_corners sets it to what it found)
        Q.appended[s] = overflow ? -(m + 1) : m;
    }
}

// ---- regroup: klt.py:244-266 + matches.py with identity pairs -----------------------------------------
constexpr int RG_THREADS = 1024;
__device__ __forceinline__ int block_excl_scan_1bit(bool flag, int* warp_tot, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned b = __ballot_sync(0xffffffffu, flag);
    const int within = __popc(b & ((1u << lane) - 1u));
    __syncthreads();                                       // warp_tot is reused between calls
    if (lane == 0) warp_tot[warp] = __popc(b);
    __syncthreads();
    int before = 0, tot = 0;
    for (int w = 0; w < RG_THREADS / 32; w++) { const int c = warp_tot[w]; if (w < warp) before += c; tot += c; }
    total = tot;
    return before + within;
}

__global__ void __launch_bounds__(RG_THREADS)
pipe_regroup_kernel(PipeTable A, PipeTable B, PipeSeq Q, PipeParams P, const float2* __restrict__ nxt,
                    const uint8_t* __restrict__ status, const float* __restrict__ err) {
    __shared__ int warp_tot[RG_THREADS / 32];
    const int s = blockIdx.x;
    const int n = Q.n_rows[s];
    const size_t base = (size_t)s * P.C;
    // pass 1: group sizes
    int c2 = 0, c1 = 0, c0 = 0;
    for (int i0 = 0; i0 < n; i0 += RG_THREADS) {
        const int i = i0 + threadIdx.x;
        bool keep = false; int st = -1;
        if (i < n) { keep = status[base + i] != 0 && err[base + i] < P.err_thr; st = A.state[base + i]; }
        c2 += __syncthreads_count(keep && st == 2);
        c1 += __syncthreads_count(keep && st == 1);
        c0 += __syncthreads_count(keep && st == 0);
    }
    // pass 2: stable placement
    int o2 = 0, o1 = c2, o0 = c2 + c1;
    for (int i0 = 0; i0 < n; i0 += RG_THREADS) {
        const int i = i0 + threadIdx.x;
        bool keep = false; int st = -1;
        if (i < n) { keep = status[base + i] != 0 && err[base + i] < P.err_thr; st = A.state[base + i]; }
        int t2, t1, t0;
        const int r2 = block_excl_scan_1bit(keep && st == 2, warp_tot, t2);
        const int r1 = block_excl_scan_1bit(keep && st == 1, warp_tot, t1);
        const int r0 = block_excl_scan_1bit(keep && st == 0, warp_tot, t0);
        if (keep) {
            const size_t src = base + i;
            const size_t dst = base + (st == 2 ? o2 + r2 : (st == 1 ? o1 + r1 : o0 + r0));
            B.kp[dst] = nxt[src];
            B.cand[dst] = 0;                                          // fresh Features (klt.py:256)
            if (st == 2) {                                            // matches.py:152-206, first group
                B.state[dst] = 2;
                for (int k = 0; k < 3; k++) B.land[dst * 3 + k] = A.land[src * 3 + k];
                B.track[dst] = make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
                for (int k = 0; k < 12; k++) B.pose[dst * 12 + k] = nan64();
            } else {
                B.state[dst] = 1;
                for (int k = 0; k < 3; k++) B.land[dst * 3 + k] = nan64();
                B.track[dst] = (st == 1) ? A.track[src] : A.kp[src];  // matches.py:83-90: a new track starts at the old keypoint
                for (int k = 0; k < 12; k++) B.pose[dst * 12 + k] = A.pose[src * 12 + k];
            }
        }
        o2 += t2; o1 += t1; o0 += t0;
    }
    if (threadIdx.x == 0) {
        int* cnt = Q.counts + (size_t)s * VO_PIPE_NCOUNTS;
        cnt[1] = n;                         // rows tracked
        cnt[2] = c2 + c1 + c0;              // rows kept
        cnt[3] = c2;                        // P3P population
        const int ap = Q.appended[s];
        cnt[7] = (ap != 0 ? 1 : 0) | (ap < 0 ? 4 : 0);
        Q.n_rows[s] = c2 + c1 + c0;
        Q.n_keep[s] = c2 + c1 + c0;
        Q.n_tri[s] = c2;
    }
}

// ---- pose: RANSAC (ransac.py:69-129 + p3p.py:51-108) and refinement (p3p.py:188-213) ------------------
// 4 warps per sequence: the kernel is a chain of short dependent phases (one CTA per sequence, one per SM), so a small
// CTA costs no time and leaves three quarters of the SM's registers to the tracker CTAs it runs under
constexpr int PO_WARPS = 4, PO_THREADS = PO_WARPS * 32, PO_HYP = 16;
constexpr int GN_MAX_ITERS = 30;

__device__ __forceinline__ int ransac_iterations(int best, int N, double log1mconf, int max_iter) {
    // ransac.py:113-120 with the numpy expressions of compute_n_iterations (ransac.py:58-67)
    double ratio = 1.0 - (double)best / (double)N;
    ratio = fmin(fmax(ratio, 0.01), 0.99);
    const double w = 1.0 - ratio;
    const double k = ceil(log1mconf / log(1.0 - pow(w, 4.0)));
    if (!(k < 2147483647.0)) return max_iter;
    const int ki = (int)k;
    return ki < max_iter ? ki : max_iter;
}

// RANSACUpdateNumIters (calib3d/src/ptsetreg.cpp), as cv2.solvePnPRansac calls it after an improvement
__device__ __forceinline__ int cv_update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = fmin(fmax(p, 0.0), 1.0); ep = fmin(fmax(ep, 0.0), 1.0);
    const double tiny = 2.2250738585072014e-308;
    double num = fmax(1.0 - p, tiny);
    double denom = 1.0 - pow(1.0 - ep, (double)model_points);
    if (denom < tiny) return 0;
    num = log(num); denom = log(denom);
    return (denom >= 0.0 || -num >= (double)max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}
// cv::RNG: multiply-with-carry; uniform(0, n) = next() % n
__device__ __forceinline__ unsigned int cv_rng_next(unsigned long long& st) {
    st = (unsigned long long)(unsigned int)st * 4164903690ull + (st >> 32);
    return (unsigned int)st;
}
// RANSACPointSetRegistrator::getSubset for 4 model points (PnP has no degeneracy check)
__device__ __forceinline__ void cv_subset4(unsigned long long& st, int N, int* out) {
    for (int i = 0; i < 4; i++) {
        for (;;) {
            const int v = (int)(cv_rng_next(st) % (unsigned int)N);
            bool dup = false;
            for (int j = 0; j < i; j++) dup |= (out[j] == v);
            if (!dup) { out[i] = v; break; }
        }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// block reduction of NV values per thread; result valid in thread 0 (and written to red[0..NV))
template <int NV>
__device__ __forceinline__ void block_sum(double* v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        const double w = warp_sum(v[k]);
        if (lane == 0) red[warp * NV + k] = w;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double a = 0.0;
        for (int w = 0; w < PO_WARPS; w++) a += red[w * NV + threadIdx.x];
        red[PO_WARPS * NV + threadIdx.x] = a;
    }
    __syncthreads();
}

__device__ __forceinline__ void so3_exp(const double* w, double* E) {
    const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double a, b;
    if (th < 1e-12) { a = 1.0; b = 0.0; }
    else { a = sin(th) / th; b = (1.0 - cos(th)) / (th * th); }
    const double Wx[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double W2[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        W2[3 * i + j] = Wx[3 * i] * Wx[j] + Wx[3 * i + 1] * Wx[3 + j] + Wx[3 * i + 2] * Wx[6 + j];
    for (int i = 0; i < 9; i++) E[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * Wx[i] + b * W2[i];
}

// 6x6 solve, Gaussian elimination with partial pivoting (what numpy.linalg.solve's LU does); false if singular
__device__ bool solve6(double (*A)[7]) {
    for (int c = 0; c < 6; c++) {
        int piv = c; double best = fabs(A[c][c]);
        for (int r = c + 1; r < 6; r++) if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
        if (!(best > 0.0)) return false;
        if (piv != c) for (int k = 0; k < 7; k++) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
        for (int r = c + 1; r < 6; r++) {
            const double f = A[r][c] / A[c][c];
            for (int k = c; k < 7; k++) A[r][k] -= f * A[c][k];
        }
    }
    for (int r = 5; r >= 0; r--) {
        double v = A[r][6];
        for (int k = r + 1; k < 6; k++) v -= A[r][k] * A[k][6];
        A[r][6] = v / A[r][r];
    }
    return true;
}

// Damped Gauss-Newton on SE(3) over the flagged points' 2N reprojection residuals (same minimum as p3p.py:188-213's
// least_squares over per-point distances; the tests hold a numpy restatement of exactly these steps).  Called by all
// PO_THREADS threads of a CTA; s_pose holds the start (R row-major | t) and receives the result.  Returns the number of
// steps tried.
struct GnShared { double red[(PO_WARPS + 1) * 28]; double tr[12]; double delta[6]; int ok; };
__device__ int gn_refine(const double* sX, const double* sY, const double* sZ, const double* sU, const double* sV,
                         const uint8_t* sIn, int N, const Intr& K, double* s_pose, GnShared& G) {
    const int tid = threadIdx.x;
    double* s_red = G.red; double* s_try = G.tr; double* s_delta = G.delta;
    int gn_iters = 0;
    // Damped Gauss-Newton on SE(3) over the inliers' 2N reprojection residuals (same minimum as p3p.py:188-213's
    // least_squares over per-point distances; the tests hold a numpy restatement of exactly these steps).
    double lam = 0.0, cost = 0.0;
    bool have_system = false;
    double Hs[21], gs[6];
    for (int it = 0; it < GN_MAX_ITERS; it++) {
        if (!have_system) {
            double acc[28];
#pragma unroll
            for (int k = 0; k < 28; k++) acc[k] = 0.0;
            double m[12];
#pragma unroll
            for (int i = 0; i < 12; i++) m[i] = s_pose[i];
            for (int i = tid; i < N; i += PO_THREADS) {
                if (!sIn[i]) continue;
                const double xc = m[0] * sX[i] + m[1] * sY[i] + m[2] * sZ[i] + m[9];
                const double yc = m[3] * sX[i] + m[4] * sY[i] + m[5] * sZ[i] + m[10];
                const double zc = m[6] * sX[i] + m[7] * sY[i] + m[8] * sZ[i] + m[11];
                const double iz = 1.0 / zc, xn = xc * iz, yn = yc * iz;
                const double ru = sU[i] - (K.fx * xn + K.cx), rv = sV[i] - (K.fy * yn + K.cy);
                const double Ju[6] = {K.fx * iz, 0.0, -K.fx * xn * iz, -K.fx * xn * yn, K.fx * (1.0 + xn * xn), -K.fx * yn};
                const double Jv[6] = {0.0, K.fy * iz, -K.fy * yn * iz, -K.fy * (1.0 + yn * yn), K.fy * xn * yn, K.fy * xn};
                int q = 0;
#pragma unroll
                for (int a = 0; a < 6; a++)
#pragma unroll
                    for (int b = a; b < 6; b++) acc[q++] += Ju[a] * Ju[b] + Jv[a] * Jv[b];
#pragma unroll
                for (int a = 0; a < 6; a++) acc[21 + a] += Ju[a] * ru + Jv[a] * rv;
                acc[27] += ru * ru + rv * rv;
            }
            block_sum<28>(acc, s_red);
            for (int k = 0; k < 21; k++) Hs[k] = s_red[PO_WARPS * 28 + k];
            for (int k = 0; k < 6; k++) gs[k] = s_red[PO_WARPS * 28 + 21 + k];
            cost = s_red[PO_WARPS * 28 + 27];
            have_system = true;
            __syncthreads();
        }
        if (tid == 0) {
            double A[6][7];
            int q = 0;
            for (int a = 0; a < 6; a++) for (int b = a; b < 6; b++) { A[a][b] = Hs[q]; A[b][a] = Hs[q]; q++; }
            for (int a = 0; a < 6; a++) { A[a][a] += lam * A[a][a]; A[a][6] = gs[a]; }
            bool ok = solve6(A);
            for (int a = 0; a < 6; a++) ok = ok && isfinite(A[a][6]);
            if (ok) {
                double d[6], E[9];
                for (int a = 0; a < 6; a++) { d[a] = A[a][6]; s_delta[a] = d[a]; }
                so3_exp(d + 3, E);
                for (int i = 0; i < 3; i++) {
                    for (int j = 0; j < 3; j++)
                        s_try[3 * i + j] = E[3 * i] * s_pose[j] + E[3 * i + 1] * s_pose[3 + j] + E[3 * i + 2] * s_pose[6 + j];
                    s_try[9 + i] = E[3 * i] * s_pose[9] + E[3 * i + 1] * s_pose[10] + E[3 * i + 2] * s_pose[11] + d[i];
                }
            }
            G.ok = ok ? 1 : 0;
        }
        __syncthreads();
        if (!G.ok) break;
        double c2[1] = {0.0};
        {
            double m[12];
#pragma unroll
            for (int i = 0; i < 12; i++) m[i] = s_try[i];
            for (int i = tid; i < N; i += PO_THREADS) {
                if (!sIn[i]) continue;
                const double xc = m[0] * sX[i] + m[1] * sY[i] + m[2] * sZ[i] + m[9];
                const double yc = m[3] * sX[i] + m[4] * sY[i] + m[5] * sZ[i] + m[10];
                const double zc = m[6] * sX[i] + m[7] * sY[i] + m[8] * sZ[i] + m[11];
                const double iz = 1.0 / zc;
                const double ru = sU[i] - (K.fx * xc * iz + K.cx), rv = sV[i] - (K.fy * yc * iz + K.cy);
                c2[0] += ru * ru + rv * rv;
            }
        }
        block_sum<1>(c2, s_red);
        const double cost2 = s_red[PO_WARPS];
        double dmax = 0.0;
        for (int a = 0; a < 6; a++) dmax = fmax(dmax, fabs(s_delta[a]));
        __syncthreads();
        gn_iters++;
        if (cost2 <= cost) {
            if (tid < 12) s_pose[tid] = s_try[tid];
            lam = lam > 1e-9 ? lam * 0.1 : 0.0;
            have_system = false;
            __syncthreads();
            if (dmax < 1e-11) break;
        } else {
            lam = (lam == 0.0) ? 1e-4 : lam * 10.0;
            if (lam > 1e8) break;
        }
    }
    return gn_iters;
}

__global__ void __launch_bounds__(PO_THREADS, 4)   // 128 registers: a quarter less of the SM taken away from the tracker it runs under
pipe_pose_kernel(PipeTable T, PipeSeq Q, PipeParams P, uint8_t* __restrict__ inliers_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = P.C;
    double* sX = reinterpret_cast<double*>(smem_raw);
    double* sY = sX + C; double* sZ = sY + C; double* sU = sZ + C; double* sV = sU + C;
    uint8_t* sIn = reinterpret_cast<uint8_t*>(sV + C);                    // [C]
    __shared__ double s_models[PO_HYP][12];
    __shared__ double s_best[12];
    __shared__ GnShared s_gn;
    __shared__ double s_pose[12];
    __shared__ int s_idx[PO_HYP][4], s_valid[PO_HYP], s_cnt[PO_HYP];
    __shared__ PipeRng s_snap[PO_HYP + 1];
    __shared__ int s_ctl[8];      // 0 stop, 1 n, 2 best, 3 n_iter, 4 draws, 5 flags, 7 parallel sampler ok
    __shared__ unsigned int s_words[PO_HYP * 7];
    __shared__ unsigned long long s_cvrng;
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = Q.n_tri[s];
    const size_t base = (size_t)s * C;
    const Intr K = P.K;
    int* cnt = Q.counts + (size_t)s * VO_PIPE_NCOUNTS;
    const bool cvmode = P.inclusive != 0;       // cv2.solvePnPRansac: float32 points, its own generator, nothing carried over
    for (int i = tid; i < N; i += PO_THREADS) {
        double x = T.land[(base + i) * 3], y = T.land[(base + i) * 3 + 1], z = T.land[(base + i) * 3 + 2];
        if (cvmode) { x = (double)(float)x; y = (double)(float)y; z = (double)(float)z; }     // solvePnPRansac converts to CV_32F
        sX[i] = x; sY[i] = y; sZ[i] = z;
        const float2 k = T.kp[base + i];
        sU[i] = (double)k.x; sV[i] = (double)k.y;
    }
    if (tid == 0) {
        s_ctl[0] = 0; s_ctl[1] = 0; s_ctl[2] = cvmode ? 0 : -1; s_ctl[3] = cvmode ? P.max_iter : Q.n_iterations[s]; s_ctl[4] = 0; s_ctl[5] = 0;
        s_snap[0] = Q.rng[s];
        s_cvrng = 0xFFFFFFFFFFFFFFFFull;        // RANSACPointSetRegistrator::run: RNG rng((uint64)-1)
    }
    __syncthreads();
    if (N < 4 || (cvmode && N == 4)) {
        // the reference raises here (Generator.choice cannot take 4 of fewer than 4): report and keep the last pose
        for (int i = tid; i < N; i += PO_THREADS) inliers_out[base + i] = 0;
        if (tid == 0) { cnt[4] = 0; cnt[7] |= 2; cnt[8] = Q.n_iterations[s]; cnt[9] = 0; cnt[11] = 0; }
        if (tid < 12) { Q.p3p_model[(size_t)s * 12 + tid] = Q.w2c[(size_t)s * 12 + tid]; Q.w2c_prev[(size_t)s * 12 + tid] = Q.w2c[(size_t)s * 12 + tid]; }
        return;
    }
    const long long draw_cap = 10ll * P.max_iter + 4096;
    while (true) {
        // the sequence's sample stream, 16 samples ahead: thread 0 draws the 16 x 7 words a sample needs when nothing is
        // rejected (the only serial part: 56 steps of the 128-bit LCG), 16 threads turn them into index sets
        if (cvmode) {
            if (tid == 0) {
                unsigned long long st = s_cvrng;
                for (int j = 0; j < PO_HYP; j++) cv_subset4(st, N, s_idx[j]);
                s_cvrng = st;
                s_ctl[7] = 1;
            }
        } else if (tid == 0) {
            PipeRng r = s_snap[0];
            for (int k = 0; k < PO_HYP * 7; k++) {
                if (k % 7 == 0) s_snap[k / 7] = r;
                s_words[k] = pcg64_next32(r);
            }
            s_snap[PO_HYP] = r;
            s_ctl[7] = (N > 4) ? 1 : 0;                   // N == 4: the first bounded draw of a sample consumes nothing
        }
        __syncthreads();
        if (!cvmode && tid < PO_HYP && N > 4) {
            int idx[4];
            if (!choice4_from(s_words + tid * 7, N, idx)) s_ctl[7] = 0;
            for (int k = 0; k < 4; k++) s_idx[tid][k] = idx[k];
        }
        __syncthreads();
        if (!cvmode && tid == 0 && !s_ctl[7]) {           // a rejected draw shifts everything after it: redo in sequence
            PipeRng r = s_snap[0];
            for (int j = 0; j < PO_HYP; j++) {
                s_snap[j] = r;
                choice4(r, N, s_idx[j]);
            }
            s_snap[PO_HYP] = r;
        }
        __syncthreads();
        if (tid < PO_HYP) {                               // model_fn (p3p.py:51-79)
            double X4[4][3], uv4[4][2];
            for (int j = 0; j < 4; j++) {
                const int id = s_idx[tid][j];
                X4[j][0] = sX[id]; X4[j][1] = sY[id]; X4[j][2] = sZ[id];
                uv4[j][0] = sU[id]; uv4[j][1] = sV[id];
            }
            double bm[12];
            const bool ok = p3pdev::solve4(X4, uv4, K, bm);
            s_valid[tid] = ok ? 1 : 0;
            for (int i = 0; i < 12; i++) s_models[tid][i] = bm[i];
        }
        __syncthreads();
        for (int h = warp; h < PO_HYP; h += PO_WARPS) {    // error_fn + threshold (p3p.py:81-108, ransac.py:104-106)
            int c = 0;
            if (s_valid[h]) {
                double m[12];
#pragma unroll
                for (int i = 0; i < 12; i++) m[i] = s_models[h][i];
                for (int b0 = 0; b0 < N; b0 += 64) {
                    const int i0 = b0 + lane, i1 = i0 + 32;
                    bool in0 = false, in1 = false;
                    if (i0 < N) in0 = p3pdev::reproj_inlier(m, sX[i0], sY[i0], sZ[i0], sU[i0], sV[i0], K, P.thr, P.inclusive);
                    if (i1 < N) in1 = p3pdev::reproj_inlier(m, sX[i1], sY[i1], sZ[i1], sU[i1], sV[i1], K, P.thr, P.inclusive);
                    c += __popc(__ballot_sync(0xffffffffu, in0)) + __popc(__ballot_sync(0xffffffffu, in1));
                }
            }
            if (lane == 0) s_cnt[h] = c;
        }
        __syncthreads();
        if (tid == 0 && cvmode) {                         // RANSACPointSetRegistrator::run over the 16 pre-scored subsets
            int it = s_ctl[1], best = s_ctl[2], niters = s_ctl[3];
            int stop = 0;
            for (int j = 0; j < PO_HYP; j++) {
                if (!(it < niters)) { stop = 1; break; }
                it++;                                     // a subset without a model still is an iteration
                if (!s_valid[j]) continue;
                const int good = s_cnt[j];
                if (good > (best > 3 ? best : 3)) {
                    best = good;
                    for (int i = 0; i < 12; i++) s_best[i] = s_models[j][i];
                    niters = cv_update_num_iters(P.conf, (double)(N - good) / (double)N, 4, niters);
                }
            }
            if (!stop && !(it < niters)) stop = 1;
            s_ctl[0] = stop; s_ctl[1] = it; s_ctl[2] = best; s_ctl[3] = niters; s_ctl[4] = it;
        } else if (tid == 0) {                            // ransac.py:90-121 over the 16 pre-scored samples
            int n = s_ctl[1], best = s_ctl[2], n_iter = s_ctl[3], draws = s_ctl[4];
            int j = 0, stop = 0;
            for (; j < PO_HYP; j++) {
                if (!(n < n_iter)) { stop = 1; break; }
                draws++;
                if (!s_valid[j]) continue;
                if (s_cnt[j] > best) {
                    best = s_cnt[j];
                    for (int i = 0; i < 12; i++) s_best[i] = s_models[j][i];
                    n_iter = ransac_iterations(best, N, P.log1mconf, P.max_iter);
                }
                n++;
            }
            if (!stop && !(n < n_iter)) stop = 1;
            if (!stop && draws >= draw_cap) { stop = 1; s_ctl[5] |= 8; }
            s_snap[0] = s_snap[j];                        // the stream continues (or stays) right after the last consumed sample
            s_ctl[0] = stop; s_ctl[1] = n; s_ctl[2] = best; s_ctl[3] = n_iter; s_ctl[4] = draws;
        }
        __syncthreads();
        if (s_ctl[0]) break;
    }
    const int best = (cvmode && s_ctl[2] == 0) ? -1 : s_ctl[2];
    if (tid == 0) {
        if (!cvmode) { Q.rng[s] = s_snap[0]; Q.n_iterations[s] = s_ctl[3]; }   // ransac.py keeps both between calls; OpenCV nothing
        cnt[8] = s_ctl[3]; cnt[9] = s_ctl[4];
    }
    if (best < 0) {                                       // no sample produced a model
        for (int i = tid; i < N; i += PO_THREADS) inliers_out[base + i] = 0;
        if (tid == 0) { cnt[4] = 0; cnt[7] |= 2 | s_ctl[5]; cnt[11] = 0; }
        if (tid < 12) { Q.p3p_model[(size_t)s * 12 + tid] = Q.w2c[(size_t)s * 12 + tid]; Q.w2c_prev[(size_t)s * 12 + tid] = Q.w2c[(size_t)s * 12 + tid]; }
        return;
    }
    // the winner's inlier mask
    {
        double m[12];
#pragma unroll
        for (int i = 0; i < 12; i++) m[i] = s_best[i];
        for (int i = tid; i < N; i += PO_THREADS) {
            const uint8_t in = p3pdev::reproj_inlier(m, sX[i], sY[i], sZ[i], sU[i], sV[i], K, P.thr, P.inclusive) ? 1 : 0;
            sIn[i] = in;
            inliers_out[base + i] = in;
        }
        if (tid < 12) { s_pose[tid] = s_best[tid]; Q.p3p_model[(size_t)s * 12 + tid] = s_best[tid]; }
    }
    __syncthreads();
    if (cvmode) {       // the refinement runs on the reference's own float64 landmarks (p3p.py:158-163), not on OpenCV's float32 copies
        for (int i = tid; i < N; i += PO_THREADS) { sX[i] = T.land[(base + i) * 3]; sY[i] = T.land[(base + i) * 3 + 1]; sZ[i] = T.land[(base + i) * 3 + 2]; }
        __syncthreads();
    }
    int gn_iters = 0;
    if (P.refine) gn_iters = gn_refine(sX, sY, sZ, sU, sV, sIn, N, K, s_pose, s_gn);
    __syncthreads();
    if (tid < 12) {
        Q.w2c_prev[(size_t)s * 12 + tid] = Q.w2c[(size_t)s * 12 + tid];
    }
    __syncthreads();
    if (tid < 12) Q.w2c[(size_t)s * 12 + tid] = s_pose[tid];
    if (tid == 0) { cnt[4] = best; cnt[7] |= s_ctl[5]; cnt[11] = gn_iters; }
}

// ---- the refinement alone (p3p.py:188-213 for the drop-in P3PPoseEstimator): one CTA per problem ------------------
__global__ void __launch_bounds__(PO_THREADS)
refine_pose_kernel(const double* __restrict__ landmarks, const double* __restrict__ keypoints, const uint8_t* __restrict__ mask,
                   int N, Intr K, const double* __restrict__ pose_in, double* __restrict__ pose_out, int* __restrict__ iters_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sX = reinterpret_cast<double*>(smem_raw);
    double* sY = sX + N; double* sZ = sY + N; double* sU = sZ + N; double* sV = sU + N;
    uint8_t* sIn = reinterpret_cast<uint8_t*>(sV + N);
    __shared__ GnShared s_gn;
    __shared__ double s_pose[12];
    const int f = blockIdx.x, tid = threadIdx.x;
    const double* L = landmarks + (size_t)f * N * 3;
    const double* P2 = keypoints + (size_t)f * N * 2;
    for (int i = tid; i < N; i += PO_THREADS) {
        sX[i] = L[3 * i]; sY[i] = L[3 * i + 1]; sZ[i] = L[3 * i + 2]; sU[i] = P2[2 * i]; sV[i] = P2[2 * i + 1];
        sIn[i] = mask ? mask[(size_t)f * N + i] : 1;
    }
    if (tid < 12) s_pose[tid] = pose_in[(size_t)f * 12 + tid];
    __syncthreads();
    const int it = gn_refine(sX, sY, sZ, sU, sV, sIn, N, K, s_pose, s_gn);
    __syncthreads();
    if (tid < 12) pose_out[(size_t)f * 12 + tid] = s_pose[tid];
    if (tid == 0 && iters_out) iters_out[f] = it;
}

// ---- update: state.py + triangulation.py:38-86 ----------------------------------------------------
constexpr int UP_THREADS = 256;
__global__ void __launch_bounds__(UP_THREADS, 2)   // 128 registers (it runs under the tracker)

pipe_update_kernel(PipeTable T, PipeSeq Q, PipeParams P, const uint8_t* __restrict__ inliers, double* __restrict__ summary) {
    __shared__ double s_c2w[12], s_w2c[12], s_w2c_prev[12], s_proj2[12];
    __shared__ int s_ncand, s_nbehind, s_ntri;
    const int s = blockIdx.x, tid = threadIdx.x;
    const int C = P.C, n = Q.n_keep[s], N = Q.n_tri[s];
    const size_t base = (size_t)s * C;
    int* cnt = Q.counts + (size_t)s * VO_PIPE_NCOUNTS;
    const bool pose_ok = !(cnt[7] & 2);
    if (tid == 0) {
        // state.py:19-23, 39-51: prev <- curr; curr = inv([R t; 0 1]) (rigid inverse)
        const double* w = Q.w2c + (size_t)s * 12;
        double c[12];
        if (pose_ok) {
            for (int i = 0; i < 3; i++) {
                for (int j = 0; j < 3; j++) c[3 * i + j] = w[3 * j + i];
                c[9 + i] = -(w[i] * w[9] + w[3 + i] * w[10] + w[6 + i] * w[11]);
            }
        }
        // c holds R^T (row-major 3x3) followed by -R^T t; the table keeps poses as row-major 3x4
        double* cw = Q.c2w + (size_t)s * 12; double* cp = Q.c2w_prev + (size_t)s * 12;
        for (int k = 0; k < 12; k++) cp[k] = cw[k];
        if (pose_ok)
            for (int i = 0; i < 3; i++) { cw[4 * i] = c[3 * i]; cw[4 * i + 1] = c[3 * i + 1]; cw[4 * i + 2] = c[3 * i + 2]; cw[4 * i + 3] = c[9 + i]; }
        for (int k = 0; k < 12; k++) { s_c2w[k] = cw[k]; s_w2c[k] = w[k]; s_w2c_prev[k] = Q.w2c_prev[(size_t)s * 12 + k]; }
        // proj2 = K @ [R | t]  (triangulation.py:54-57 with inv(inv(.)))
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 4; j++) {
                const double m0 = j < 3 ? w[j] : w[9], m1 = j < 3 ? w[3 + j] : w[10], m2 = j < 3 ? w[6 + j] : w[11];
                s_proj2[4 * i + j] = P.K9[3 * i] * m0 + P.K9[3 * i + 1] * m1 + P.K9[3 * i + 2] * m2;
            }
        s_ncand = 0; s_nbehind = 0; s_ntri = 0;
    }
    __syncthreads();
    // main.py:264-268 + state.py:167-178 (P3P outliers), then state.py:139-165 (candidates)
    int my_cand = 0;
    for (int i = tid; i < n; i += UP_THREADS) {
        const size_t r = base + i;
        int st = T.state[r];
        const float2 kp = T.kp[r];
        if (pose_ok && i < N && !inliers[r]) {
            st = 0;
            T.state[r] = 0;
            T.track[r] = kp;
            for (int k = 0; k < 12; k++) T.pose[r * 12 + k] = s_c2w[k];
        }
        uint8_t cand = 0;
        if (st == 1 && pose_ok) {
            const float2 tr = T.track[r];
            const double* Ki = P.Kinv9;
            const double* ps = T.pose + r * 12;
            const double sx = (double)tr.x, sy = (double)tr.y, ex = (double)kp.x, ey = (double)kp.y;
            const double a0 = sx * Ki[0] + sy * Ki[1] + Ki[2], a1 = sx * Ki[3] + sy * Ki[4] + Ki[5], a2 = sx * Ki[6] + sy * Ki[7] + Ki[8];
            const double b0 = ex * Ki[0] + ey * Ki[1] + Ki[2], b1 = ex * Ki[3] + ey * Ki[4] + Ki[5], b2 = ex * Ki[6] + ey * Ki[7] + Ki[8];
            const double d10 = ps[0] * a0 + ps[1] * a1 + ps[2] * a2, d11 = ps[4] * a0 + ps[5] * a1 + ps[6] * a2, d12 = ps[8] * a0 + ps[9] * a1 + ps[10] * a2;
            const double d20 = s_c2w[0] * b0 + s_c2w[1] * b1 + s_c2w[2] * b2, d21 = s_c2w[4] * b0 + s_c2w[5] * b1 + s_c2w[6] * b2,
                         d22 = s_c2w[8] * b0 + s_c2w[9] * b1 + s_c2w[10] * b2;
            const double dot = d10 * d20 + d11 * d21 + d12 * d22;
            const double n1 = sqrt(d10 * d10 + d11 * d11 + d12 * d12), n2 = sqrt(d20 * d20 + d21 * d21 + d22 * d22);
            const double ang = acos(dot / (n1 * n2));
            cand = ang >= P.bearing_thr ? 1 : 0;
        }
        T.cand[r] = cand;
        my_cand += cand;
    }
    if (my_cand) atomicAdd(&s_ncand, my_cand);
    __syncthreads();
    const int ncand = s_ncand;
    if (ncand > 0) {
        // triangulation.py:38-86 for the candidates (each with its own start pose), state.py:83-84
        for (int i = tid; i < n; i += UP_THREADS) {
            const size_t r = base + i;
            if (!T.cand[r]) continue;
            const double* ps = T.pose + r * 12;          // camera-to-world [R | t], row-major 3x4
            double e[12];                                // inv -> world-to-camera
            for (int a = 0; a < 3; a++) {
                for (int b = 0; b < 3; b++) e[4 * a + b] = ps[4 * b + a];
                e[4 * a + 3] = -(ps[a] * ps[3] + ps[4 + a] * ps[7] + ps[8 + a] * ps[11]);
            }
            double c1[12];
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 4; b++)
                    c1[4 * a + b] = P.K9[3 * a] * e[b] + P.K9[3 * a + 1] * e[4 + b] + P.K9[3 * a + 2] * e[8 + b];
            double c2[12];
            for (int k = 0; k < 12; k++) c2[k] = s_proj2[k];
            const float2 tr = T.track[r], kp = T.kp[r];
            double x[4];
            tridev::triangulate_point(c1, c2, (double)tr.x, (double)tr.y, (double)kp.x, (double)kp.y, P.tri_mode, x);
            T.land[r * 3] = x[0] / x[3]; T.land[r * 3 + 1] = x[1] / x[3]; T.land[r * 3 + 2] = x[2] / x[3];
            T.state[r] = 2;
        }
        __syncthreads();
        // state.py:92-110: landmarks behind the current or the previous camera are dropped (all rows)
        int my_b = 0;
        for (int i = tid; i < n; i += UP_THREADS) {
            const size_t r = base + i;
            const double X = T.land[r * 3], Y = T.land[r * 3 + 1], Z = T.land[r * 3 + 2];
            const double zc = s_w2c[6] * X + s_w2c[7] * Y + s_w2c[8] * Z + s_w2c[11];
            const double zp = s_w2c_prev[6] * X + s_w2c_prev[7] * Y + s_w2c_prev[8] * Z + s_w2c_prev[11];
            if (zc < 0.0 || zp < 0.0) {
                for (int k = 0; k < 3; k++) T.land[r * 3 + k] = nan64();
                T.state[r] = 0;
                T.track[r] = T.kp[r];
                for (int k = 0; k < 12; k++) T.pose[r * 12 + k] = s_c2w[k];
                my_b++;
            }
        }
        if (my_b) atomicAdd(&s_nbehind, my_b);
    }
    __syncthreads();
    int my_t = 0;
    for (int i = tid; i < n; i += UP_THREADS) my_t += T.state[base + i] == 2;
    if (my_t) atomicAdd(&s_ntri, my_t);
    __syncthreads();
    if (tid == 0) {
        cnt[0] = n; cnt[5] = ncand; cnt[6] = s_ntri; cnt[10] = s_nbehind;
    }
    // summary row: pose (3x4 camera-to-world) followed by the counters (as int32 pairs in the same buffer)
    double* out = summary + (size_t)s * VO_PIPE_SUMMARY_DOUBLES;
    if (tid < 12) out[tid] = s_c2w[tid];
    __syncthreads();
    if (tid < VO_PIPE_NCOUNTS) reinterpret_cast<int*>(out + 12)[tid] = cnt[tid];
}

// ---- two-view bootstrap of a resident sequence (main.py:203-231) ---------------------------------------------
// After the regroup every kept row is matched (state 1) with its track start in the first frame: rows [n_tri, n_keep)
// are `matched_candidate_inliers` of both frames.  Their (track start, keypoint) pairs go to the bootstrap kernel
// (bootstrap.cu) as float64, as the reference hands cv2 its float keypoints.
__global__ void __launch_bounds__(256)
pipe_boot_gather_kernel(PipeTable T, PipeSeq Q, PipeParams P, double* __restrict__ p1, double* __restrict__ p2, int* __restrict__ n_pts) {
    const int s = blockIdx.x, C = P.C;
    const int n = Q.n_keep[s], n0 = Q.n_tri[s];
    const size_t base = (size_t)s * C;
    for (int j = threadIdx.x; j < n - n0; j += blockDim.x) {
        const float2 a = T.track[base + n0 + j], b = T.kp[base + n0 + j];
        p1[(base + j) * 2] = (double)a.x; p1[(base + j) * 2 + 1] = (double)a.y;
        p2[(base + j) * 2] = (double)b.x; p2[(base + j) * 2 + 1] = (double)b.y;
    }
    if (threadIdx.x == 0) n_pts[s] = n - n0;
}

// main.py:213-222: update_with_local_pose(M) (state.py:25-37), update_with_local_landmarks(landmarks[inliers], mask)
// (state.py:51-110, with the cheirality check of all rows), reset_outliers(~inliers) (state.py:167-178); summary row.
__global__ void __launch_bounds__(UP_THREADS)
pipe_boot_apply_kernel(PipeTable T, PipeSeq Q, PipeParams P, const double* __restrict__ M_all, const double* __restrict__ land,
                       const uint8_t* __restrict__ mask, const int* __restrict__ info, double* __restrict__ summary) {
    __shared__ double s_c2w[12], s_c2w_prev[12], s_w2c[12], s_w2c_prev[12];
    __shared__ int s_nbehind, s_ntri, s_ninl;
    const int s = blockIdx.x, tid = threadIdx.x, C = P.C;
    const int n = Q.n_keep[s], n0 = Q.n_tri[s];
    const size_t base = (size_t)s * C;
    int* cnt = Q.counts + (size_t)s * VO_PIPE_NCOUNTS;
    const bool found = info[4 * s] != 0;
    if (tid == 0) {
        double* cw = Q.c2w + (size_t)s * 12; double* cp = Q.c2w_prev + (size_t)s * 12;
        double* wc = Q.w2c + (size_t)s * 12; double* wp = Q.w2c_prev + (size_t)s * 12;
        for (int k = 0; k < 12; k++) { cp[k] = cw[k]; wp[k] = wc[k]; }         // state.py:19-23: prev <- curr
        if (found) {
            const double* M = M_all + (size_t)s * 12;                          // [R | t], frame 1 -> frame 2
            double inv[12];                                                    // inv(M) = [R^T | -R^T t]
            for (int i = 0; i < 3; i++) {
                for (int j = 0; j < 3; j++) inv[4 * i + j] = M[4 * j + i];
                inv[4 * i + 3] = -(M[i] * M[3] + M[4 + i] * M[7] + M[8 + i] * M[11]);
            }
            double c[12];                                                      // curr = prev @ inv(M)
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 4; j++)
                    c[4 * i + j] = cp[4 * i] * inv[j] + cp[4 * i + 1] * inv[4 + j] + cp[4 * i + 2] * inv[8 + j] + (j == 3 ? cp[4 * i + 3] : 0.0);
            for (int k = 0; k < 12; k++) cw[k] = c[k];
            // world-to-camera in the estimator's layout (R row-major, then t): the rigid inverse of curr
            for (int i = 0; i < 3; i++) {
                for (int j = 0; j < 3; j++) wc[3 * i + j] = c[4 * j + i];
                wc[9 + i] = -(c[i] * c[3] + c[4 + i] * c[7] + c[8 + i] * c[11]);
            }
        }
        for (int k = 0; k < 12; k++) { s_c2w[k] = cw[k]; s_c2w_prev[k] = cp[k]; s_w2c[k] = wc[k]; s_w2c_prev[k] = wp[k]; }
        s_nbehind = 0; s_ntri = 0; s_ninl = 0;
    }
    __syncthreads();
    if (found) {
        int my_inl = 0;
        for (int i = tid; i < n; i += UP_THREADS) {
            const size_t r = base + i;
            T.cand[r] = 0;
            if (i < n0 || !mask[base + i - n0]) continue;
            const double* X = land + (base + i - n0) * 3;                      // frame-1 coordinates -> world (state.py:59-63)
            for (int k = 0; k < 3; k++)
                T.land[r * 3 + k] = s_c2w_prev[4 * k] * X[0] + s_c2w_prev[4 * k + 1] * X[1] + s_c2w_prev[4 * k + 2] * X[2] + s_c2w_prev[4 * k + 3];
            T.state[r] = 2;
            my_inl++;
        }
        if (my_inl) atomicAdd(&s_ninl, my_inl);
        __syncthreads();
        int my_b = 0;
        for (int i = tid; i < n; i += UP_THREADS) {                            // state.py:92-110 (all rows)
            const size_t r = base + i;
            const double X = T.land[r * 3], Y = T.land[r * 3 + 1], Z = T.land[r * 3 + 2];
            const double zc = s_w2c[6] * X + s_w2c[7] * Y + s_w2c[8] * Z + s_w2c[11];
            const double zp = s_w2c_prev[6] * X + s_w2c_prev[7] * Y + s_w2c_prev[8] * Z + s_w2c_prev[11];
            bool reset = false;
            if (zc < 0.0 || zp < 0.0) {
                for (int k = 0; k < 3; k++) T.land[r * 3 + k] = nan64();
                reset = true;
                my_b++;
            }
            if (i >= n0 && !mask[base + i - n0]) reset = true;                 // main.py:211-212, 222: the bootstrap's outliers
            if (reset) {
                T.state[r] = 0;
                T.track[r] = T.kp[r];
                for (int k = 0; k < 12; k++) T.pose[r * 12 + k] = s_c2w[k];
            }
        }
        if (my_b) atomicAdd(&s_nbehind, my_b);
    }
    __syncthreads();
    int my_t = 0;
    for (int i = tid; i < n; i += UP_THREADS) my_t += T.state[base + i] == 2;
    if (my_t) atomicAdd(&s_ntri, my_t);
    __syncthreads();
    if (tid == 0) {
        cnt[0] = n; cnt[3] = n - n0; cnt[4] = s_ninl; cnt[5] = 0; cnt[6] = s_ntri; cnt[8] = info[4 * s + 1]; cnt[9] = info[4 * s + 1];
        cnt[10] = s_nbehind; cnt[11] = 0;
        if (!found) cnt[7] |= 2;
    }
    double* out = summary + (size_t)s * VO_PIPE_SUMMARY_DOUBLES;
    if (tid < 12) out[tid] = s_c2w[tid];
    __syncthreads();
    if (tid < VO_PIPE_NCOUNTS) reinterpret_cast<int*>(out + 12)[tid] = cnt[tid];
}

// ---- detector outputs -> float corners (Harris keypoints are integers) ------------------------------------
__global__ void pipe_fill_int_kernel(int* p, int v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- table initialisation from the detector (KLTTracker.__init__, klt.py:40-50) -----------------------------
__global__ void __launch_bounds__(256)
pipe_init_from_det_kernel(PipeTable T, PipeSeq Q, PipeParams P, const int* __restrict__ det_xy, const float* __restrict__ det_f,
                          const int* __restrict__ det_n, int det_cap) {
    const int s = blockIdx.x;
    int m = det_n[s];
    if (m > P.C) m = P.C;
    const size_t base = (size_t)s * P.C;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        float x, y;
        if (det_f) { x = det_f[((size_t)s * det_cap + i) * 2]; y = det_f[((size_t)s * det_cap + i) * 2 + 1]; }
        else { x = (float)det_xy[((size_t)s * det_cap + i) * 2]; y = (float)det_xy[((size_t)s * det_cap + i) * 2 + 1]; }
        const size_t r = base + i;
        T.kp[r] = make_float2(x, y);
        T.track[r] = make_float2(x, y);
        T.state[r] = 0;
        T.cand[r] = 0;
        for (int k = 0; k < 3; k++) T.land[r * 3 + k] = nan64();
        double* p = T.pose + r * 12;
        for (int k = 0; k < 12; k++) p[k] = (k == 0 || k == 5 || k == 10) ? 1.0 : 0.0;
    }
    if (threadIdx.x == 0) { Q.n_rows[s] = m; Q.n_keep[s] = m; Q.n_tri[s] = 0; Q.num_features[s] = det_n[s]; Q.appended[s] = 0; }
}

// test hook: n_draws samples of Generator.choice(arange(N), size=4, replace=False) from a given PCG64 state
__global__ void pipe_rng_test_kernel(PipeRng* r, int N, int n_draws, int* out) {
    if (threadIdx.x || blockIdx.x) return;
    PipeRng st = *r;
    for (int i = 0; i < n_draws; i++) choice4(st, N, out + 4 * i);
    *r = st;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct vo_pipeline {
    vo_ctx* ctx = nullptr;
    vo_pipeline_params p{};
    PipeParams dp{};
    int n_levels = 0;
    size_t pitch0 = 0, frame_bytes = 0;
    unsigned char* base = nullptr;
    uint8_t* pyr[2] = {nullptr, nullptr};
    PipeTable tab[2]{};
    PipeSeq seq{};
    int cur = 0;                 // pyramid / table / detection set holding the CURRENT (old) frame
    long long steps = 0;
    bool primed = false;
    // KLT outputs
    float2* nxt = nullptr; uint8_t* status = nullptr; float* err = nullptr;
    uint8_t* inliers = nullptr;
    // detector
    double* resp = nullptr;
    int* det_xy[2] = {nullptr, nullptr}; int* det_n[2] = {nullptr, nullptr};
    int det_cap = 0;
    // results: two summary sets so that the download of step t overlaps step t + 1
    double* summary[2] = {nullptr, nullptr};
    // host entry point: staging for tightly packed frames (two sets), streams, events
    uint8_t* stage[2] = {nullptr, nullptr};
    // a step is spread over three streams: the caller's (pyramid, append, tracker, regroup), `side` (detector on the new
    // frame) and `pose_stream` (RANSAC + refinement, state update).  The latency-bound per-sequence kernels of step t
    // (one CTA per sequence) run under the tracker of step t + 1: nothing the tracker reads (keypoints, row counts) is
    // written after the regroup.
    cudaStream_t copy_stream = nullptr, down_stream = nullptr, side = nullptr, pose_stream = nullptr;
    cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_res[2] = {nullptr, nullptr};
    cudaEvent_t ev_level0 = nullptr, ev_det = nullptr, ev_fork = nullptr, ev_rg = nullptr, ev_upd = nullptr;
    int stage_next = 0, prefetched = 0, in_flight = 0, sub_next = 0;
    size_t pose_smem = 0;
    // two-view bootstrap (vo_pipeline_bootstrap_*): point pairs in, model / landmarks / masks out
    double *boot_p1 = nullptr, *boot_p2 = nullptr, *boot_F = nullptr, *boot_M = nullptr, *boot_land = nullptr;
    uint8_t *boot_mask = nullptr, *boot_fmask = nullptr;
    int *boot_n = nullptr, *boot_info = nullptr;
};

static int pipe_run_detector(vo_pipeline* pl, const uint8_t* pyr_level0, int set, cudaStream_t s) {
    const vo_pipeline_params& p = pl->p;
    vo_ctx* ctx = pl->ctx;
    if (p.detector == VO_DETECTOR_HARRIS) {
        int rc;
        if ((rc = vo_launch_harris_response(ctx, pyr_level0, p.n_seq, p.H, p.W, pl->pitch0, pl->frame_bytes, p.patch_size, p.kappa,
                                            pl->resp, s))) return rc;
        if ((rc = vo_launch_harris_nms(ctx, pl->resp, p.n_seq, p.H, p.W, p.nms_radius, p.det_max_corners, pl->det_xy[set], nullptr,
                                       s))) return rc;
        pipe_fill_int_kernel<<<vo_div_up(p.n_seq, 256), 256, 0, s>>>(pl->det_n[set], p.det_max_corners, p.n_seq);
        ctx->launches++;
        VO_CHECK_LAUNCH();
    } else if (p.detector == VO_DETECTOR_GFTT) {
        return vo_launch_gftt(ctx, pyr_level0, p.n_seq, p.H, p.W, pl->pitch0, pl->frame_bytes, p.det_max_corners, p.gftt_quality,
                              p.gftt_min_distance, p.gftt_block_size, (float*)pl->resp, (float*)pl->det_xy[set], pl->det_n[set], nullptr, s);
    }
    return VO_OK;
}

extern "C" {

int vo_pipeline_create(vo_ctx* ctx, const vo_pipeline_params* prm, vo_pipeline** out) {
    VO_REQUIRE(ctx && prm && out, "vo_pipeline_create: null argument");
    *out = nullptr;
    const vo_pipeline_params& p = *prm;
    VO_REQUIRE(p.n_seq >= 1 && p.H > 0 && p.W > 0, "vo_pipeline_create: bad sizes");
    VO_REQUIRE(p.capacity >= 32 && p.capacity <= 8192 && p.capacity % 32 == 0, "vo_pipeline_create: capacity must be a multiple of 32 in [32, 8192]");
    VO_REQUIRE(p.detector == VO_DETECTOR_NONE || p.detector == VO_DETECTOR_HARRIS || p.detector == VO_DETECTOR_GFTT,
               "vo_pipeline_create: unknown detector %d", p.detector);
    VO_REQUIRE(p.detector == VO_DETECTOR_NONE || (p.det_max_corners >= 1 && p.det_max_corners <= p.capacity),
               "vo_pipeline_create: det_max_corners must be in [1, capacity]");
    VO_REQUIRE(p.K[0] != 0.0 && p.K[4] != 0.0, "vo_pipeline_create: singular intrinsic matrix");
    VO_REQUIRE(p.tri_mode == 0 || p.tri_mode == 1, "vo_pipeline_create: tri_mode must be 0 or 1");
    VO_REQUIRE(p.ransac_confidence > 0.0 && p.ransac_confidence < 1.0 && p.ransac_max_iterations >= 1, "vo_pipeline_create: bad RANSAC parameters");
    VO_CUDA(cudaSetDevice(ctx->device));
    vo_pipeline* pl = new vo_pipeline();
    pl->ctx = ctx; pl->p = p;
    int lh[8], lw[8]; size_t lp[8], lo[8];
    int rc = vo_klt_layout_host(p.H, p.W, p.klt_max_level, p.klt_win, &pl->n_levels, lh, lw, lp, lo, &pl->frame_bytes);
    if (rc) { delete pl; return rc; }
    pl->pitch0 = lp[0];
    PipeParams& d = pl->dp;
    d.C = p.capacity;
    d.K = Intr{p.K[0], p.K[4], p.K[2], p.K[5]};
    for (int i = 0; i < 9; i++) { d.K9[i] = p.K[i]; d.Kinv9[i] = p.Kinv[i]; }
    d.thr = p.p3p_threshold; d.inclusive = p.p3p_inclusive; d.log1mconf = p.ransac_log1mconf != 0.0 ? p.ransac_log1mconf : log(1.0 - p.ransac_confidence);
    d.max_iter = p.ransac_max_iterations; d.refine = p.refine; d.conf = p.ransac_confidence;
    d.bearing_thr = p.bearing_threshold; d.tri_mode = p.tri_mode; d.err_thr = p.klt_error_threshold; d.redetect_frac = p.redetect_fraction;
    auto finish = [&]() -> int {
        const size_t S = p.n_seq, C = p.capacity, npx = (size_t)p.H * p.W;
        pl->det_cap = p.detector == VO_DETECTOR_NONE ? 1 : p.det_max_corners;
        const size_t DC = pl->det_cap;
        size_t off = 0;
        auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
        const size_t o_p0 = carve(S * pl->frame_bytes), o_p1 = carve(S * pl->frame_bytes);
        size_t o_t[2][6];
        for (int t = 0; t < 2; t++) {
            o_t[t][0] = carve(S * C * 8); o_t[t][1] = carve(S * C * 8); o_t[t][2] = carve(S * C);
            o_t[t][3] = carve(S * C * 24); o_t[t][4] = carve(S * C * 96); o_t[t][5] = carve(S * C);
        }
        const size_t o_nk = carve(S * 4);
        const size_t o_nr = carve(S * 4), o_nt = carve(S * 4), o_nf = carve(S * 4), o_ni = carve(S * 4), o_ap = carve(S * 4);
        const size_t o_cw = carve(S * 96), o_cp = carve(S * 96), o_wc = carve(S * 96), o_wp = carve(S * 96), o_pm = carve(S * 96);
        const size_t o_rng = carve(S * sizeof(PipeRng)), o_cnt = carve(S * VO_PIPE_NCOUNTS * 4);
        const size_t o_nx = carve(S * C * 8), o_st = carve(S * C), o_er = carve(S * C * 4), o_in = carve(S * C);
        const size_t o_resp = carve(p.detector == VO_DETECTOR_HARRIS ? S * npx * 8 : (p.detector == VO_DETECTOR_GFTT ? S * npx * 4 : 0));
        const size_t o_dx0 = carve(S * DC * 8), o_dx1 = carve(S * DC * 8), o_dn0 = carve(S * 4), o_dn1 = carve(S * 4);
        const size_t o_su0 = carve(S * VO_PIPE_SUMMARY_DOUBLES * 8), o_su1 = carve(S * VO_PIPE_SUMMARY_DOUBLES * 8);
        const size_t o_sg0 = carve(S * npx + 256), o_sg1 = carve(S * npx + 256);
        const size_t o_b1 = carve(S * C * 16), o_b2 = carve(S * C * 16), o_bF = carve(S * 72), o_bM = carve(S * 96), o_bl = carve(S * C * 24);
        const size_t o_bm = carve(S * C), o_bf = carve(S * C), o_bn = carve(S * 4), o_bi = carve(S * 16);
        cudaError_t e = cudaMalloc(&pl->base, off);
        if (e != cudaSuccess) { vo_set_error("vo_pipeline_create: cudaMalloc(%zu) -> %s", off, cudaGetErrorString(e)); return VO_ERR_CUDA; }
        unsigned char* b = pl->base;
        VO_CUDA(cudaMemsetAsync(b, 0, off, ctx->stream));
        pl->pyr[0] = b + o_p0; pl->pyr[1] = b + o_p1;
        for (int t = 0; t < 2; t++) {
            pl->tab[t].kp = (float2*)(b + o_t[t][0]); pl->tab[t].track = (float2*)(b + o_t[t][1]); pl->tab[t].state = b + o_t[t][2];
            pl->tab[t].land = (double*)(b + o_t[t][3]); pl->tab[t].pose = (double*)(b + o_t[t][4]); pl->tab[t].cand = b + o_t[t][5];
        }
        PipeSeq& q = pl->seq;
        q.n_keep = (int*)(b + o_nk);
        q.n_rows = (int*)(b + o_nr); q.n_tri = (int*)(b + o_nt); q.num_features = (int*)(b + o_nf); q.n_iterations = (int*)(b + o_ni);
        q.appended = (int*)(b + o_ap);
        q.c2w = (double*)(b + o_cw); q.c2w_prev = (double*)(b + o_cp); q.w2c = (double*)(b + o_wc); q.w2c_prev = (double*)(b + o_wp);
        q.p3p_model = (double*)(b + o_pm); q.rng = (PipeRng*)(b + o_rng); q.counts = (int*)(b + o_cnt);
        pl->nxt = (float2*)(b + o_nx); pl->status = b + o_st; pl->err = (float*)(b + o_er); pl->inliers = b + o_in;
        pl->resp = p.detector != VO_DETECTOR_NONE ? (double*)(b + o_resp) : nullptr;
        pl->det_xy[0] = (int*)(b + o_dx0); pl->det_xy[1] = (int*)(b + o_dx1); pl->det_n[0] = (int*)(b + o_dn0); pl->det_n[1] = (int*)(b + o_dn1);
        pl->summary[0] = (double*)(b + o_su0); pl->summary[1] = (double*)(b + o_su1);
        pl->stage[0] = b + o_sg0; pl->stage[1] = b + o_sg1;
        pl->boot_p1 = (double*)(b + o_b1); pl->boot_p2 = (double*)(b + o_b2); pl->boot_F = (double*)(b + o_bF); pl->boot_M = (double*)(b + o_bM);
        pl->boot_land = (double*)(b + o_bl); pl->boot_mask = b + o_bm; pl->boot_fmask = b + o_bf; pl->boot_n = (int*)(b + o_bn); pl->boot_info = (int*)(b + o_bi);
        // default per-sequence state: identity poses, the reference's initial iteration count (ransac.py:56)
        {
            std::vector<double> eye(S * 12, 0.0), eyew(S * 12, 0.0);
            for (size_t s = 0; s < S; s++) { eye[s * 12] = eye[s * 12 + 5] = eye[s * 12 + 10] = 1.0; eyew[s * 12] = eyew[s * 12 + 4] = eyew[s * 12 + 8] = 1.0; }
            for (double* dst : {q.c2w, q.c2w_prev}) VO_CUDA(cudaMemcpyAsync(dst, eye.data(), S * 96, cudaMemcpyHostToDevice, ctx->stream));
            for (double* dst : {q.w2c, q.w2c_prev, q.p3p_model}) VO_CUDA(cudaMemcpyAsync(dst, eyew.data(), S * 96, cudaMemcpyHostToDevice, ctx->stream));
            const double k0 = ceil(pl->dp.log1mconf / log(1.0 - pow(1.0 - p.ransac_outlier_ratio, 4.0)));
            const int it0 = (k0 < (double)p.ransac_max_iterations) ? (int)k0 : p.ransac_max_iterations;
            std::vector<int> iters(S, p.ransac_initial_iterations > 0 ? p.ransac_initial_iterations : it0);
            VO_CUDA(cudaMemcpyAsync(q.n_iterations, iters.data(), S * 4, cudaMemcpyHostToDevice, ctx->stream));
            VO_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        VO_CUDA(cudaStreamCreateWithFlags(&pl->copy_stream, cudaStreamNonBlocking));
        VO_CUDA(cudaStreamCreateWithFlags(&pl->down_stream, cudaStreamNonBlocking));
        {
            int lo_p = 0, hi_p = 0;
            VO_CUDA(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
            VO_CUDA(cudaStreamCreateWithPriority(&pl->side, cudaStreamNonBlocking, hi_p));
            VO_CUDA(cudaStreamCreateWithPriority(&pl->pose_stream, cudaStreamNonBlocking, hi_p));
        }
        VO_CUDA(cudaEventCreateWithFlags(&pl->ev_rg, cudaEventDisableTiming));
        VO_CUDA(cudaEventCreateWithFlags(&pl->ev_upd, cudaEventDisableTiming));
        for (int i = 0; i < 2; i++) {
            VO_CUDA(cudaEventCreateWithFlags(&pl->ev_up[i], cudaEventDisableTiming));
            VO_CUDA(cudaEventCreateWithFlags(&pl->ev_done[i], cudaEventDisableTiming));
            VO_CUDA(cudaEventCreateWithFlags(&pl->ev_res[i], cudaEventDisableTiming));
        }
        VO_CUDA(cudaEventCreateWithFlags(&pl->ev_level0, cudaEventDisableTiming));
        VO_CUDA(cudaEventCreateWithFlags(&pl->ev_det, cudaEventDisableTiming));
        VO_CUDA(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
        pl->pose_smem = (size_t)p.capacity * 41 + 64;
        VO_CUDA(cudaFuncSetAttribute(pipe_pose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->pose_smem));
        if (p.detector == VO_DETECTOR_HARRIS) {
            const int rc2 = vo_harris_nms_reserve(ctx, p.n_seq, p.H, p.W, p.nms_radius, p.det_max_corners);
            if (rc2) return rc2;
        }
        if (p.detector == VO_DETECTOR_GFTT) {
            const int rc2 = vo_gftt_reserve(ctx, p.n_seq, p.H, p.W, p.gftt_min_distance);
            if (rc2) return rc2;
        }
        return VO_OK;
    };
    rc = finish();
    if (rc) { vo_pipeline_destroy(pl); return rc; }
    *out = pl;
    return VO_OK;
}

void vo_pipeline_destroy(vo_pipeline* pl) {
    if (!pl) return;
    cudaSetDevice(pl->ctx->device);
    cudaStreamSynchronize(pl->ctx->stream);
    for (cudaStream_t s : {pl->copy_stream, pl->down_stream, pl->side, pl->pose_stream}) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    for (int i = 0; i < 2; i++) for (cudaEvent_t ev : {pl->ev_up[i], pl->ev_done[i], pl->ev_res[i]}) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : {pl->ev_level0, pl->ev_det, pl->ev_fork, pl->ev_rg, pl->ev_upd}) if (ev) cudaEventDestroy(ev);
    if (pl->base) cudaFree(pl->base);
    (void)cudaGetLastError();
    delete pl;
}

// Make `frames` the current (old) frame of every sequence: pyramid, detector, and -- with init_tables != 0 -- a fresh
// table from the detected corners (KLTTracker.__init__).
static int pipe_prime(vo_pipeline* pl, const uint8_t* d_frames, size_t pitch, size_t frame_stride, int init_tables, cudaStream_t s) {
    const vo_pipeline_params& p = pl->p;
    vo_ctx* ctx = pl->ctx;
    int rc;
    if ((rc = vo_launch_klt_pyramid(ctx, d_frames, p.n_seq, p.H, p.W, pitch, frame_stride, p.klt_max_level, p.klt_win,
                                    pl->pyr[pl->cur], s))) return rc;
    if ((rc = pipe_run_detector(pl, pl->pyr[pl->cur], pl->cur, s))) return rc;
    if (init_tables) {
        VO_REQUIRE(p.detector != VO_DETECTOR_NONE, "vo_pipeline_prime: init_tables needs a detector");
        const bool fl = p.detector == VO_DETECTOR_GFTT;      // Shi-Tomasi corners are float32, Harris keypoints int32
        pipe_init_from_det_kernel<<<p.n_seq, 256, 0, s>>>(pl->tab[pl->cur], pl->seq, pl->dp, fl ? nullptr : pl->det_xy[pl->cur],
                                                          fl ? (const float*)pl->det_xy[pl->cur] : nullptr, pl->det_n[pl->cur], pl->det_cap);
        ctx->launches++;
        VO_CHECK_LAUNCH();
    }
    pl->primed = true;
    return VO_OK;
}

int vo_pipeline_prime_dev(vo_pipeline* pl, const uint8_t* d_frames, size_t pitch, size_t frame_stride, int init_tables, void* stream) {
    VO_REQUIRE(pl && d_frames, "vo_pipeline_prime_dev: null argument");
    VO_CUDA(cudaSetDevice(pl->ctx->device));
    return pipe_prime(pl, d_frames, pitch, frame_stride, init_tables, stream ? (cudaStream_t)stream : pl->ctx->stream);
}

int vo_pipeline_prime_host(vo_pipeline* pl, const uint8_t* h_frames, int init_tables) {
    VO_REQUIRE(pl && h_frames, "vo_pipeline_prime_host: null argument");
    VO_REQUIRE(pl->in_flight == 0 && pl->prefetched == 0, "vo_pipeline_prime_host: steps are in flight");
    VO_CUDA(cudaSetDevice(pl->ctx->device));
    cudaStream_t s = pl->ctx->stream;
    const size_t npx = (size_t)pl->p.H * pl->p.W;
    VO_CUDA(cudaMemcpyAsync(pl->stage[0], h_frames, pl->p.n_seq * npx, cudaMemcpyHostToDevice, s));
    int rc = pipe_prime(pl, pl->stage[0], (size_t)pl->p.W, npx, init_tables, s);
    if (rc) return rc;
    VO_CUDA(cudaStreamSynchronize(s));
    return VO_OK;
}

// one step: frames (device) -> tables advanced, summary written to set `slot`.  Work is enqueued on s and on the two
// internal streams; ev_upd fires when the step's results (tables, summary) are complete.  `staging_free` (optional) is
// recorded as soon as d_frames has been consumed (after the pyramid's level-0 copy).
static int pipe_step(vo_pipeline* pl, const uint8_t* d_frames, size_t pitch, size_t frame_stride, int slot, cudaStream_t s,
                     cudaEvent_t staging_free = nullptr, cudaEvent_t slot_free = nullptr, const double* boot = nullptr) {
    const vo_pipeline_params& p = pl->p;
    vo_ctx* ctx = pl->ctx;
    VO_REQUIRE(pl->primed, "vo_pipeline step: call vo_pipeline_prime_* first (there is no previous frame yet)");
    const int cur = pl->cur, nx = 1 - cur;
    int rc;
    const bool fork = !ctx->env_frontend_serial;
    cudaStream_t sd = fork ? pl->side : s, sp = fork ? pl->pose_stream : s;
    // new frame -> pyramid; the detector runs on it on the side stream (its corners are what the NEXT step appends
    // when a sequence runs low: klt.py:207-230 detects on the old frame, which is this step's new frame)
    if ((rc = vo_launch_klt_pyramid(ctx, d_frames, p.n_seq, p.H, p.W, pitch, frame_stride, p.klt_max_level, p.klt_win,
                                    pl->pyr[nx], s, pl->ev_level0))) return rc;
    if (staging_free) VO_CUDA(cudaEventRecord(staging_free, s));
    // The detector of the previous step has left its corners (what this step's append may take).  Only the append and
    // what follows it wait for that: the pyramid above writes the slot of the frame before last, which nothing reads
    // any more (the previous tracker is ahead of it on this stream, the detector that read its level 0 finished before
    // that tracker started), so the level-0 copy and the pyramid slip in while the last band CTAs are still running.
    // (The wait has to be queued before this step's detector re-records ev_det.)
    if (fork && p.detector != VO_DETECTOR_NONE) VO_CUDA(cudaStreamWaitEvent(s, pl->ev_det, 0));
    if (p.detector != VO_DETECTOR_NONE && fork) {
        VO_CUDA(cudaStreamWaitEvent(sd, pl->ev_level0, 0));
        if ((rc = pipe_run_detector(pl, pl->pyr[nx], nx, sd))) return rc;
        VO_CUDA(cudaEventRecord(pl->ev_det, sd));
    }
    if (p.detector != VO_DETECTOR_NONE) {
        const bool fl = p.detector == VO_DETECTOR_GFTT;
        pipe_append_kernel<<<p.n_seq, 256, 0, s>>>(pl->tab[cur], pl->seq, pl->dp, fl ? nullptr : pl->det_xy[cur],
                                                   fl ? (const float*)pl->det_xy[cur] : nullptr, pl->det_n[cur], pl->det_cap);
        ctx->launches++;
        VO_CHECK_LAUNCH();
    } else {
        VO_CUDA(cudaMemsetAsync(pl->seq.appended, 0, (size_t)p.n_seq * 4, s));
    }
    if ((rc = vo_launch_klt_track(ctx, pl->pyr[cur], pl->pyr[nx], p.n_seq, p.H, p.W, p.klt_max_level, p.klt_win, p.klt_max_iters,
                                  p.klt_epsilon, p.klt_min_eig, (const float*)pl->tab[cur].kp, p.capacity, (float*)pl->nxt, pl->status,
                                  pl->err, s, pl->seq.n_rows))) return rc;
    // the regroup reads the states the previous step's update wrote
    if (fork) VO_CUDA(cudaStreamWaitEvent(s, pl->ev_upd, 0));
    pipe_regroup_kernel<<<p.n_seq, RG_THREADS, 0, s>>>(pl->tab[cur], pl->tab[nx], pl->seq, pl->dp, pl->nxt, pl->status, pl->err);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    if (fork) {
        VO_CUDA(cudaEventRecord(pl->ev_rg, s));
        VO_CUDA(cudaStreamWaitEvent(sp, pl->ev_rg, 0));
        if (slot_free) VO_CUDA(cudaStreamWaitEvent(sp, slot_free, 0));
    } else if (slot_free) {
        VO_CUDA(cudaStreamWaitEvent(s, slot_free, 0));
    }
    if (boot) {
        // main.py:203-231: the matched rows of the two frames -> relative pose, landmarks, inliers -> table
        pipe_boot_gather_kernel<<<p.n_seq, 256, 0, sp>>>(pl->tab[nx], pl->seq, pl->dp, pl->boot_p1, pl->boot_p2, pl->boot_n);
        ctx->launches++;
        VO_CHECK_LAUNCH();
        if ((rc = vo_launch_bootstrap(ctx, pl->boot_p1, pl->boot_p2, p.n_seq, p.capacity, pl->boot_n, p.K, boot[0], boot[1], (int)boot[2],
                                      pl->boot_F, pl->boot_M, pl->boot_land, pl->boot_mask, pl->boot_fmask, pl->boot_info, sp))) return rc;
        pipe_boot_apply_kernel<<<p.n_seq, UP_THREADS, 0, sp>>>(pl->tab[nx], pl->seq, pl->dp, pl->boot_M, pl->boot_land, pl->boot_mask,
                                                               pl->boot_info, pl->summary[slot]);
        ctx->launches++;
        VO_CHECK_LAUNCH();
    } else {
        pipe_pose_kernel<<<p.n_seq, PO_THREADS, pl->pose_smem, sp>>>(pl->tab[nx], pl->seq, pl->dp, pl->inliers);
        ctx->launches++;
        VO_CHECK_LAUNCH();
        pipe_update_kernel<<<p.n_seq, UP_THREADS, 0, sp>>>(pl->tab[nx], pl->seq, pl->dp, pl->inliers, pl->summary[slot]);
        ctx->launches++;
        VO_CHECK_LAUNCH();
    }
    VO_CUDA(cudaEventRecord(pl->ev_upd, sp));
    if (!fork && p.detector != VO_DETECTOR_NONE) { if ((rc = pipe_run_detector(pl, pl->pyr[nx], nx, s))) return rc; }
    pl->cur = nx;
    pl->steps++;
    return VO_OK;
}

// make `stream` wait for everything the steps enqueued so far left running on the internal streams
static int pipe_join(vo_pipeline* pl, cudaStream_t s) {
    if (!pl->ctx->env_frontend_serial && pl->steps > 0) {
        VO_CUDA(cudaStreamWaitEvent(s, pl->ev_upd, 0));
        if (pl->p.detector != VO_DETECTOR_NONE) VO_CUDA(cudaStreamWaitEvent(s, pl->ev_det, 0));
    }
    return VO_OK;
}

int vo_pipeline_sync_dev(vo_pipeline* pl, void* stream) {
    VO_REQUIRE(pl, "vo_pipeline_sync_dev: null argument");
    VO_CUDA(cudaSetDevice(pl->ctx->device));
    return pipe_join(pl, stream ? (cudaStream_t)stream : pl->ctx->stream);
}

int vo_pipeline_step_dev(vo_pipeline* pl, const uint8_t* d_frames, size_t pitch, size_t frame_stride, void* stream) {
    VO_REQUIRE(pl && d_frames, "vo_pipeline_step_dev: null argument");
    VO_CUDA(cudaSetDevice(pl->ctx->device));
    return pipe_step(pl, d_frames, pitch, frame_stride, 0, stream ? (cudaStream_t)stream : pl->ctx->stream);
}

int vo_pipeline_bootstrap_dev(vo_pipeline* pl, const uint8_t* d_frames, size_t pitch, size_t frame_stride, double threshold,
                              double confidence, int max_iters, void* stream) {
    VO_REQUIRE(pl && d_frames, "vo_pipeline_bootstrap_dev: null argument");
    VO_REQUIRE(max_iters >= 1, "vo_pipeline_bootstrap_dev: max_iters must be positive");
    VO_CUDA(cudaSetDevice(pl->ctx->device));
    const double boot[3] = {threshold, confidence, (double)max_iters};
    return pipe_step(pl, d_frames, pitch, frame_stride, 0, stream ? (cudaStream_t)stream : pl->ctx->stream, nullptr, nullptr, boot);
}

int vo_pipeline_bootstrap_host(vo_pipeline* pl, const uint8_t* h_frames, double threshold, double confidence, int max_iters,
                               double* h_summary) {
    VO_REQUIRE(pl && h_frames && h_summary, "vo_pipeline_bootstrap_host: null argument");
    VO_REQUIRE(pl->in_flight == 0 && pl->prefetched == 0, "vo_pipeline_bootstrap_host: submitted steps are still in flight");
    VO_REQUIRE(max_iters >= 1, "vo_pipeline_bootstrap_host: max_iters must be positive");
    VO_CUDA(cudaSetDevice(pl->ctx->device));
    cudaStream_t s = pl->ctx->stream;
    const size_t npx = (size_t)pl->p.H * pl->p.W;
    VO_CUDA(cudaMemcpyAsync(pl->stage[0], h_frames, pl->p.n_seq * npx, cudaMemcpyHostToDevice, s));
    const double boot[3] = {threshold, confidence, (double)max_iters};
    int rc = pipe_step(pl, pl->stage[0], (size_t)pl->p.W, npx, 0, s, nullptr, nullptr, boot);
    if (rc) return rc;
    if ((rc = pipe_join(pl, s))) return rc;
    VO_CUDA(cudaMemcpyAsync(h_summary, pl->summary[0], (size_t)pl->p.n_seq * VO_PIPE_SUMMARY_DOUBLES * 8, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaStreamSynchronize(s));
    return VO_OK;
}

const double* vo_pipeline_summary_dev(vo_pipeline* pl) { return pl ? pl->summary[0] : nullptr; }

static int pipe_upload(vo_pipeline* pl, int set, const uint8_t* h_frames) {
    const size_t npx = (size_t)pl->p.H * pl->p.W;
    VO_CUDA(cudaStreamWaitEvent(pl->copy_stream, pl->ev_done[set], 0));      // the step that read this set has finished
    VO_CUDA(cudaMemcpyAsync(pl->stage[set], h_frames, pl->p.n_seq * npx, cudaMemcpyHostToDevice, pl->copy_stream));
    VO_CUDA(cudaEventRecord(pl->ev_up[set], pl->copy_stream));
    return VO_OK;
}

int vo_pipeline_prefetch_host(vo_pipeline* pl, const uint8_t* h_frames) {
    VO_REQUIRE(pl && h_frames, "vo_pipeline_prefetch_host: null argument");
    VO_REQUIRE(pl->prefetched < 2, "vo_pipeline_prefetch_host: two uploads are already pending");
    VO_CUDA(cudaSetDevice(pl->ctx->device));
    const int set = pl->stage_next;
    int rc = pipe_upload(pl, set, h_frames);
    if (rc) return rc;
    pl->prefetched++; pl->stage_next = 1 - set;
    return VO_OK;
}

int vo_pipeline_submit_host(vo_pipeline* pl, const uint8_t* h_frames, double* h_summary) {
    VO_REQUIRE(pl && h_summary, "vo_pipeline_submit_host: null argument");
    VO_REQUIRE(pl->in_flight < 2, "vo_pipeline_submit_host: two steps are already in flight (call vo_pipeline_wait_host)");
    VO_CUDA(cudaSetDevice(pl->ctx->device));
    cudaStream_t s = pl->ctx->stream;
    int set;
    if (pl->prefetched) {
        set = (pl->stage_next + 2 - pl->prefetched) & 1;
        pl->prefetched--;
    } else {
        VO_REQUIRE(h_frames, "vo_pipeline_submit_host: no frames (neither prefetched nor given)");
        set = pl->stage_next;
        int rc = pipe_upload(pl, set, h_frames);
        if (rc) return rc;
        pl->stage_next = 1 - set;
    }
    const int slot = pl->sub_next;
    VO_CUDA(cudaStreamWaitEvent(s, pl->ev_up[set], 0));
    const size_t npx = (size_t)pl->p.H * pl->p.W;
    // ev_done[set]: the staged frames have been consumed; ev_res[slot]: summary set `slot` has left the device
    int rc = pipe_step(pl, pl->stage[set], (size_t)pl->p.W, npx, slot, s, pl->ev_done[set], pl->ev_res[slot]);
    if (rc) return rc;
    VO_CUDA(cudaStreamWaitEvent(pl->down_stream, pl->ev_upd, 0));
    if (pl->ctx->env_frontend_serial) VO_CUDA(cudaStreamWaitEvent(pl->down_stream, pl->ev_done[set], 0));
    VO_CUDA(cudaMemcpyAsync(h_summary, pl->summary[slot], (size_t)pl->p.n_seq * VO_PIPE_SUMMARY_DOUBLES * 8, cudaMemcpyDeviceToHost,
                            pl->down_stream));
    VO_CUDA(cudaEventRecord(pl->ev_res[slot], pl->down_stream));
    pl->in_flight++;
    pl->sub_next = 1 - slot;
    return VO_OK;
}

int vo_pipeline_wait_host(vo_pipeline* pl) {
    VO_REQUIRE(pl, "vo_pipeline_wait_host: null argument");
    VO_REQUIRE(pl->in_flight > 0, "vo_pipeline_wait_host: nothing was submitted");
    VO_CUDA(cudaSetDevice(pl->ctx->device));
    const int slot = (pl->sub_next + 2 - pl->in_flight) & 1;
    VO_CUDA(cudaEventSynchronize(pl->ev_res[slot]));
    pl->in_flight--;
    return VO_OK;
}

int vo_pipeline_step_host(vo_pipeline* pl, const uint8_t* h_frames, double* h_summary) {
    VO_REQUIRE(pl && pl->in_flight == 0, "vo_pipeline_step_host: submitted steps are still in flight");
    int rc = vo_pipeline_submit_host(pl, h_frames, h_summary);
    if (rc) return rc;
    return vo_pipeline_wait_host(pl);
}

// ---- table access (tests, the Python mirror of the reference's Features, bootstrap hand-over) ------------------
int vo_pipeline_read_table_host(vo_pipeline* pl, int seq, int* n_rows, float* h_kp, double* h_land, uint8_t* h_state, float* h_track,
                                double* h_pose, uint8_t* h_cand, double* h_c2w, int32_t* h_scalars, uint8_t* h_inliers,
                                uint64_t* h_rng) {
    VO_REQUIRE(pl && seq >= 0 && seq < pl->p.n_seq && n_rows, "vo_pipeline_read_table_host: bad argument");
    VO_CUDA(cudaSetDevice(pl->ctx->device));
    VO_CUDA(cudaDeviceSynchronize());
    const PipeTable& T = pl->tab[pl->cur];
    const size_t C = pl->p.capacity, b = (size_t)seq * C;
    int n = 0;
    VO_CUDA(cudaMemcpy(&n, pl->seq.n_rows + seq, 4, cudaMemcpyDeviceToHost));
    *n_rows = n;
    if (h_kp) VO_CUDA(cudaMemcpy(h_kp, T.kp + b, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (h_land) VO_CUDA(cudaMemcpy(h_land, T.land + b * 3, (size_t)n * 24, cudaMemcpyDeviceToHost));
    if (h_state) VO_CUDA(cudaMemcpy(h_state, T.state + b, (size_t)n, cudaMemcpyDeviceToHost));
    if (h_track) VO_CUDA(cudaMemcpy(h_track, T.track + b, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (h_pose) VO_CUDA(cudaMemcpy(h_pose, T.pose + b * 12, (size_t)n * 96, cudaMemcpyDeviceToHost));
    if (h_cand) VO_CUDA(cudaMemcpy(h_cand, T.cand + b, (size_t)n, cudaMemcpyDeviceToHost));
    if (h_c2w) {
        VO_CUDA(cudaMemcpy(h_c2w, pl->seq.c2w + (size_t)seq * 12, 96, cudaMemcpyDeviceToHost));
        VO_CUDA(cudaMemcpy(h_c2w + 12, pl->seq.c2w_prev + (size_t)seq * 12, 96, cudaMemcpyDeviceToHost));
        VO_CUDA(cudaMemcpy(h_c2w + 24, pl->seq.w2c + (size_t)seq * 12, 96, cudaMemcpyDeviceToHost));
        VO_CUDA(cudaMemcpy(h_c2w + 36, pl->seq.p3p_model + (size_t)seq * 12, 96, cudaMemcpyDeviceToHost));
    }
    if (h_scalars) {
        VO_CUDA(cudaMemcpy(h_scalars, pl->seq.num_features + seq, 4, cudaMemcpyDeviceToHost));
        VO_CUDA(cudaMemcpy(h_scalars + 1, pl->seq.n_iterations + seq, 4, cudaMemcpyDeviceToHost));
        VO_CUDA(cudaMemcpy(h_scalars + 2, pl->seq.n_tri + seq, 4, cudaMemcpyDeviceToHost));
        VO_CUDA(cudaMemcpy(h_scalars + 3, pl->seq.counts + (size_t)seq * VO_PIPE_NCOUNTS, VO_PIPE_NCOUNTS * 4, cudaMemcpyDeviceToHost));
    }
    if (h_inliers) {
        int nt = 0;
        VO_CUDA(cudaMemcpy(&nt, pl->seq.n_tri + seq, 4, cudaMemcpyDeviceToHost));
        VO_CUDA(cudaMemcpy(h_inliers, pl->inliers + b, (size_t)nt, cudaMemcpyDeviceToHost));
    }
    if (h_rng) {
        PipeRng r;
        VO_CUDA(cudaMemcpy(&r, pl->seq.rng + seq, sizeof(PipeRng), cudaMemcpyDeviceToHost));
        h_rng[0] = r.s_hi; h_rng[1] = r.s_lo; h_rng[2] = r.inc_hi; h_rng[3] = r.inc_lo; h_rng[4] = r.has32; h_rng[5] = r.u32;
    }
    return VO_OK;
}

int vo_pipeline_write_table_host(vo_pipeline* pl, int seq, int n_rows, const float* h_kp, const double* h_land, const uint8_t* h_state,
                                 const float* h_track, const double* h_pose, const double* h_c2w, int num_features, int n_iterations,
                                 const uint64_t* h_rng) {
    VO_REQUIRE(pl && seq >= 0 && seq < pl->p.n_seq, "vo_pipeline_write_table_host: bad sequence index");
    VO_REQUIRE(n_rows <= pl->p.capacity, "vo_pipeline_write_table_host: %d rows exceed the capacity %d", n_rows, pl->p.capacity);
    VO_REQUIRE(n_rows <= 0 || (h_kp && h_land && h_state && h_track && h_pose), "vo_pipeline_write_table_host: null column");
    VO_CUDA(cudaSetDevice(pl->ctx->device));
    VO_CUDA(cudaDeviceSynchronize());
    const PipeTable& T = pl->tab[pl->cur];
    const size_t C = pl->p.capacity, b = (size_t)seq * C, n = n_rows > 0 ? n_rows : 0;
    if (n) {
        VO_CUDA(cudaMemcpy(T.kp + b, h_kp, n * 8, cudaMemcpyHostToDevice));
        VO_CUDA(cudaMemcpy(T.land + b * 3, h_land, n * 24, cudaMemcpyHostToDevice));
        VO_CUDA(cudaMemcpy(T.state + b, h_state, n, cudaMemcpyHostToDevice));
        VO_CUDA(cudaMemcpy(T.track + b, h_track, n * 8, cudaMemcpyHostToDevice));
        VO_CUDA(cudaMemcpy(T.pose + b * 12, h_pose, n * 96, cudaMemcpyHostToDevice));
        VO_CUDA(cudaMemset(T.cand + b, 0, n));
    }
    if (n_rows >= 0) {                                       // n_rows < 0: scalars only, the rows stay
        VO_CUDA(cudaMemcpy(pl->seq.n_rows + seq, &n_rows, 4, cudaMemcpyHostToDevice));
        VO_CUDA(cudaMemcpy(pl->seq.n_keep + seq, &n_rows, 4, cudaMemcpyHostToDevice));
        int n_tri = 0;
        for (int i = 0; i < n_rows; i++) n_tri += h_state[i] == 2;
        VO_CUDA(cudaMemcpy(pl->seq.n_tri + seq, &n_tri, 4, cudaMemcpyHostToDevice));
    }
    if (num_features >= 0) VO_CUDA(cudaMemcpy(pl->seq.num_features + seq, &num_features, 4, cudaMemcpyHostToDevice));
    if (n_iterations > 0) VO_CUDA(cudaMemcpy(pl->seq.n_iterations + seq, &n_iterations, 4, cudaMemcpyHostToDevice));
    if (h_c2w) {
        // current camera-to-world pose (row-major 3x4); the world-to-camera form is its rigid inverse
        double w[12];
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) w[3 * i + j] = h_c2w[4 * j + i];
            w[9 + i] = -(h_c2w[i] * h_c2w[3] + h_c2w[4 + i] * h_c2w[7] + h_c2w[8 + i] * h_c2w[11]);
        }
        VO_CUDA(cudaMemcpy(pl->seq.c2w + (size_t)seq * 12, h_c2w, 96, cudaMemcpyHostToDevice));
        VO_CUDA(cudaMemcpy(pl->seq.c2w_prev + (size_t)seq * 12, h_c2w, 96, cudaMemcpyHostToDevice));
        VO_CUDA(cudaMemcpy(pl->seq.w2c + (size_t)seq * 12, w, 96, cudaMemcpyHostToDevice));
        VO_CUDA(cudaMemcpy(pl->seq.w2c_prev + (size_t)seq * 12, w, 96, cudaMemcpyHostToDevice));
    }
    if (h_rng) {
        PipeRng r;
        r.s_hi = h_rng[0]; r.s_lo = h_rng[1]; r.inc_hi = h_rng[2]; r.inc_lo = h_rng[3];
        r.has32 = (unsigned)h_rng[4]; r.u32 = (unsigned)h_rng[5]; r.pad0 = r.pad1 = 0;
        VO_CUDA(cudaMemcpy(pl->seq.rng + seq, &r, sizeof(PipeRng), cudaMemcpyHostToDevice));
    }
    return VO_OK;
}

// the detector's corners of the current frame (what the next step would append): [det_max_corners][2] (int32 for the
// Harris detector, float32 for Shi-Tomasi), count
int vo_pipeline_read_detections_host(vo_pipeline* pl, int seq, void* h_xy, int* n) {
    VO_REQUIRE(pl && seq >= 0 && seq < pl->p.n_seq && h_xy && n, "vo_pipeline_read_detections_host: bad argument");
    VO_REQUIRE(pl->p.detector != VO_DETECTOR_NONE, "vo_pipeline_read_detections_host: no detector configured");
    VO_CUDA(cudaSetDevice(pl->ctx->device));
    VO_CUDA(cudaDeviceSynchronize());
    VO_CUDA(cudaMemcpy(n, pl->det_n[pl->cur] + seq, 4, cudaMemcpyDeviceToHost));
    VO_CUDA(cudaMemcpy(h_xy, pl->det_xy[pl->cur] + (size_t)seq * pl->det_cap * 2, (size_t)pl->det_cap * 8, cudaMemcpyDeviceToHost));
    return VO_OK;
}

// p3p.py:188-213 for n_frames independent problems: minimise the reprojection error of the masked correspondences over
// the pose, starting from pose_in (R row-major | t, world -> camera).
int vo_refine_pose_dev(vo_ctx* ctx, const double* d_landmarks, const double* d_keypoints, const uint8_t* d_mask, int n_frames, int N,
                       const double* K9, const double* d_pose_in, double* d_pose_out, int32_t* d_iters, void* stream) {
    VO_REQUIRE(ctx && d_landmarks && d_keypoints && K9 && d_pose_in && d_pose_out, "vo_refine_pose_dev: null argument");
    VO_REQUIRE(n_frames >= 1 && N >= 1, "vo_refine_pose_dev: bad sizes");
    const size_t smem = (size_t)N * 41 + 64;
    VO_REQUIRE(smem <= 200 * 1024, "vo_refine_pose_dev: at most %d correspondences per problem", (int)((200 * 1024 - 64) / 41));
    VO_CUDA(cudaSetDevice(ctx->device));
    if (vo_ctx_once(ctx, VO_ATTR_PIPE))
        VO_CUDA(cudaFuncSetAttribute(refine_pose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    refine_pose_kernel<<<n_frames, PO_THREADS, smem, s>>>(d_landmarks, d_keypoints, d_mask, N, Intr{K9[0], K9[4], K9[2], K9[5]}, d_pose_in,
                                                          d_pose_out, d_iters);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    return VO_OK;
}

int vo_refine_pose_host(vo_ctx* ctx, const double* h_landmarks, const double* h_keypoints, const uint8_t* h_mask, int n_frames, int N,
                        const double* K9, const double* h_pose_in, double* h_pose_out, int32_t* h_iters) {
    VO_REQUIRE(ctx && h_landmarks && h_keypoints && K9 && h_pose_in && h_pose_out, "vo_refine_pose_host: null argument");
    VO_REQUIRE(n_frames >= 1 && N >= 1, "vo_refine_pose_host: bad sizes");
    VO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t F = n_frames;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_l = carve(F * N * 24), o_k = carve(F * N * 16), o_m = carve(F * N), o_pi = carve(F * 96), o_po = carve(F * 96), o_it = carve(F * 4);
    int rc = vo_buf_reserve(&ctx->scratch[13], off);
    if (rc) return rc;
    unsigned char* b = (unsigned char*)ctx->scratch[13].p;
    VO_CUDA(cudaMemcpyAsync(b + o_l, h_landmarks, F * N * 24, cudaMemcpyHostToDevice, s));
    VO_CUDA(cudaMemcpyAsync(b + o_k, h_keypoints, F * N * 16, cudaMemcpyHostToDevice, s));
    if (h_mask) VO_CUDA(cudaMemcpyAsync(b + o_m, h_mask, F * N, cudaMemcpyHostToDevice, s));
    VO_CUDA(cudaMemcpyAsync(b + o_pi, h_pose_in, F * 96, cudaMemcpyHostToDevice, s));
    if ((rc = vo_refine_pose_dev(ctx, (double*)(b + o_l), (double*)(b + o_k), h_mask ? b + o_m : nullptr, n_frames, N, K9, (double*)(b + o_pi),
                                 (double*)(b + o_po), (int32_t*)(b + o_it), s))) return rc;
    VO_CUDA(cudaMemcpyAsync(h_pose_out, b + o_po, F * 96, cudaMemcpyDeviceToHost, s));
    if (h_iters) VO_CUDA(cudaMemcpyAsync(h_iters, b + o_it, F * 4, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaStreamSynchronize(s));
    return VO_OK;
}

// test hook for the sample stream: state6 = {state_hi, state_lo, inc_hi, inc_lo, has_uint32, uinteger}
int vo_test_pcg64_choice4_host(vo_ctx* ctx, uint64_t* state6, int N, int n_draws, int32_t* h_out) {
    VO_REQUIRE(ctx && state6 && h_out && N >= 4 && n_draws >= 1, "vo_test_pcg64_choice4_host: bad argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    int rc = vo_buf_reserve(&ctx->scratch[12], sizeof(PipeRng) + (size_t)n_draws * 16 + 256);
    if (rc) return rc;
    PipeRng r;
    r.s_hi = state6[0]; r.s_lo = state6[1]; r.inc_hi = state6[2]; r.inc_lo = state6[3];
    r.has32 = (unsigned)state6[4]; r.u32 = (unsigned)state6[5]; r.pad0 = r.pad1 = 0;
    unsigned char* b = (unsigned char*)ctx->scratch[12].p;
    VO_CUDA(cudaMemcpyAsync(b, &r, sizeof(r), cudaMemcpyHostToDevice, ctx->stream));
    pipe_rng_test_kernel<<<1, 32, 0, ctx->stream>>>((PipeRng*)b, N, n_draws, (int*)(b + 256));
    ctx->launches++;
    VO_CHECK_LAUNCH();
    VO_CUDA(cudaMemcpyAsync(h_out, b + 256, (size_t)n_draws * 16, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaMemcpyAsync(&r, b, sizeof(r), cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    state6[0] = r.s_hi; state6[1] = r.s_lo; state6[2] = r.inc_hi; state6[3] = r.inc_lo; state6[4] = r.has32; state6[5] = r.u32;
    return VO_OK;
}

}  // extern "C"
