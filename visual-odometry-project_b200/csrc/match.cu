// Brute-force descriptor matching for sm_100a.
//
// Replaces  /root/reference/src/vo/features/harris.py:196-264  HarrisCornerDetector.matchDescriptor:
//   cv2.BFMatcher().knnMatch(desc1, desc2, k=2)  (L2 norm, float32), ratio test m < 0.85 n, and the
//   "first query wins" uniqueness of the train index (harris.py:252-258).
//
// The Harris descriptors are raw 8-bit patches (harris.py:160-194), so squared distances are exact
// integers:  |a - b|^2 = |a|^2 + |b|^2 - 2 a.b  with a.b accumulated by dp4a.  OpenCV sums the same
// integers in float32 (exact below 2^24) and takes sqrtf; the ratio test is evaluated in double exactly
// as the Python expression `m.distance < 0.85 * n.distance`.  k-NN order is (distance, train index), as
// cv2's batchDistance keeps the earlier index on ties.
//
// knn kernel: CTA = 4 warps x 8 queries; the query tile and 64-descriptor train tiles are staged in shared
// memory as 32-bit words (row pitch odd -> conflict free); a lane owns 2 train columns and keeps 8 x 2
// running dot products; the 2 best per query are merged across lanes with shuffles.
// select kernel: one CTA per frame applies the ratio test, resolves duplicates with atomicMin on the
// query index (= first come in the reference's loop) and compacts the pairs in query order.
#include "common.cuh"

namespace {

constexpr int MQ = 32;        // queries per CTA
constexpr int MT = 64;        // train descriptors per tile
constexpr int M_THREADS = 128;
constexpr int M_DW_MAX = 272; // descriptor length up to 1088 bytes ((2*16+1)^2 = 1089 is just over: r <= 15 -> 961)

__device__ __forceinline__ int dp4a_uu(unsigned int a, unsigned int b, int c) {
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// packs D bytes of one descriptor into words (zero padded) and returns |d|^2 contribution of the words it wrote
__device__ __forceinline__ unsigned int load_word(const uint8_t* __restrict__ d, int D, int w) {
    unsigned int v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int j = 4 * w + k;
        if (j < D) v |= (unsigned int)d[j] << (8 * k);
    }
    return v;
}

__device__ __forceinline__ void best2_insert(unsigned long long key, unsigned long long& b0, unsigned long long& b1) {
    if (key < b0) { b1 = b0; b0 = key; }
    else if (key < b1) b1 = key;
}

__global__ void __launch_bounds__(M_THREADS)
knn2_kernel(const uint8_t* __restrict__ q_desc, const uint8_t* __restrict__ t_desc, int Q, int T, int D,
            unsigned long long* __restrict__ best /*[F][Q][2] (dist2 << 32 | train idx)*/) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int DW = (D + 3) / 4;
    const int P = DW | 1;                                   // odd row pitch in words
    unsigned int* sq = reinterpret_cast<unsigned int*>(smem_raw);          // [MQ][P]
    unsigned int* stt = sq + MQ * P;                                       // [MT][P]
    int* nq = reinterpret_cast<int*>(stt + MT * P);                        // [MQ]
    int* nt = nq + MQ;                                                     // [MT]
    const int f = blockIdx.y, q0 = blockIdx.x * MQ;
    const uint8_t* qd = q_desc + (size_t)f * Q * D;
    const uint8_t* td = t_desc + (size_t)f * T * D;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < MQ * DW; i += M_THREADS) {
        const int r = i / DW, w = i - r * DW;
        sq[r * P + w] = (q0 + r < Q) ? load_word(qd + (size_t)(q0 + r) * D, D, w) : 0u;
    }
    __syncthreads();
    if (tid < MQ) {
        int s = 0;
        for (int w = 0; w < DW; w++) s = dp4a_uu(sq[tid * P + w], sq[tid * P + w], s);
        nq[tid] = s;
    }
    unsigned long long b0[8], b1[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { b0[k] = ~0ull; b1[k] = ~0ull; }

    for (int t0 = 0; t0 < T; t0 += MT) {
        __syncthreads();
        for (int i = tid; i < MT * DW; i += M_THREADS) {
            const int r = i / DW, w = i - r * DW;
            stt[r * P + w] = (t0 + r < T) ? load_word(td + (size_t)(t0 + r) * D, D, w) : 0u;
        }
        __syncthreads();
        if (tid < MT) {
            int s = 0;
            for (int w = 0; w < DW; w++) s = dp4a_uu(stt[tid * P + w], stt[tid * P + w], s);
            nt[tid] = s;
        }
        __syncthreads();
        int acc[8][2];
#pragma unroll
        for (int k = 0; k < 8; k++) { acc[k][0] = 0; acc[k][1] = 0; }
        const unsigned int* ta = stt + lane * P;
        const unsigned int* tb = stt + (lane + 32) * P;
        const unsigned int* qa = sq + (warp * 8) * P;
        for (int w = 0; w < DW; w++) {
            const unsigned int va = ta[w], vb = tb[w];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const unsigned int vq = qa[k * P + w];          // broadcast
                acc[k][0] = dp4a_uu(vq, va, acc[k][0]);
                acc[k][1] = dp4a_uu(vq, vb, acc[k][1]);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int t = t0 + lane + 32 * c;
                if (t < T) {
                    const int d2 = nq[warp * 8 + k] + nt[lane + 32 * c] - 2 * acc[k][c];
                    best2_insert(((unsigned long long)(unsigned int)d2 << 32) | (unsigned int)t, b0[k], b1[k]);
                }
            }
        }
    }
    // merge the per-lane best-2 across the warp
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long o0 = __shfl_xor_sync(0xFFFFFFFFu, b0[k], o);
            const unsigned long long o1 = __shfl_xor_sync(0xFFFFFFFFu, b1[k], o);
            best2_insert(o0, b0[k], b1[k]);
            best2_insert(o1, b0[k], b1[k]);
        }
        const int q = q0 + warp * 8 + k;
        if (lane == 0 && q < Q) {
            best[((size_t)f * Q + q) * 2] = b0[k];
            best[((size_t)f * Q + q) * 2 + 1] = b1[k];
        }
    }
}

// ratio test + first-come uniqueness + ordered compaction; one CTA per frame
__global__ void __launch_bounds__(1024)
match_select_kernel(const unsigned long long* __restrict__ best, int Q, int T, double ratio,
                    int* __restrict__ winner /*[F][T] scratch*/, int* __restrict__ pairs /*[F][Q][2]*/,
                    int* __restrict__ n_pairs /*[F]*/) {
    __shared__ int s_warp[32];
    __shared__ int s_run;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long* b = best + (size_t)f * Q * 2;
    int* win = winner + (size_t)f * T;
    for (int t = tid; t < T; t += 1024) win[t] = 0x7fffffff;
    if (tid == 0) s_run = 0;
    __syncthreads();
    // harris.py:254-258: for m, n in matches: if m.distance < 0.85 * n.distance and the train index is unused
    for (int q = tid; q < Q; q += 1024) {
        const unsigned long long m = b[2 * q], n = b[2 * q + 1];
        if (n == ~0ull) continue;                               // fewer than two neighbours: cv2 returns no pair to unpack
        const float dm = sqrtf((float)(unsigned int)(m >> 32)), dn = sqrtf((float)(unsigned int)(n >> 32));
        if ((double)dm < ratio * (double)dn) atomicMin(&win[(unsigned int)m], q);
    }
    __syncthreads();
    for (int base = 0; base < Q; base += 1024) {
        const int q = base + tid;
        bool keep = false;
        int t = 0;
        if (q < Q) {
            const unsigned long long m = b[2 * q], n = b[2 * q + 1];
            if (n != ~0ull) {
                t = (int)(unsigned int)m;
                const float dm = sqrtf((float)(unsigned int)(m >> 32)), dn = sqrtf((float)(unsigned int)(n >> 32));
                keep = ((double)dm < ratio * (double)dn) && win[t] == q;
            }
        }
        const unsigned int bal = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int off = s_run;
        for (int w = 0; w < warp; w++) off += s_warp[w];
        if (keep) {
            const int o = off + __popc(bal & ((1u << lane) - 1u));
            pairs[((size_t)f * Q + o) * 2] = q;
            pairs[((size_t)f * Q + o) * 2 + 1] = t;
        }
        __syncthreads();
        if (tid == 0) { int tot = 0; for (int w = 0; w < 32; w++) tot += s_warp[w]; s_run += tot; }
        __syncthreads();
    }
    if (tid == 0) n_pairs[f] = s_run;
}

}  // namespace

int vo_launch_match(vo_ctx* ctx, const uint8_t* d_q, const uint8_t* d_t, int n_frames, int Q, int T, int D, double ratio,
                    int* d_pairs, int* d_n_pairs, cudaStream_t stream) {
    VO_REQUIRE(n_frames >= 1 && Q >= 1 && T >= 1 && D >= 1 && (D + 3) / 4 <= M_DW_MAX, "match: bad sizes (descriptor length <= %d bytes)", 4 * M_DW_MAX);
    const int DW = (D + 3) / 4, P = DW | 1;
    const size_t smem = (size_t)(MQ + MT) * P * 4 + (MQ + MT) * 4;
    if (vo_ctx_once(ctx, VO_ATTR_MATCH))
        VO_CUDA(cudaFuncSetAttribute(knn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_best = carve((size_t)n_frames * Q * 16), o_win = carve((size_t)n_frames * T * 4);
    int rc = vo_buf_reserve(&ctx->scratch[10], off);
    if (rc) return rc;
    unsigned char* b = (unsigned char*)ctx->scratch[10].p;
    dim3 g(vo_div_up(Q, MQ), n_frames);
    knn2_kernel<<<g, M_THREADS, smem, stream>>>(d_q, d_t, Q, T, D, (unsigned long long*)(b + o_best));
    ctx->launches++;
    VO_CHECK_LAUNCH();
    match_select_kernel<<<n_frames, 1024, 0, stream>>>((unsigned long long*)(b + o_best), Q, T, ratio, (int*)(b + o_win),
                                                       d_pairs, d_n_pairs);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    return VO_OK;
}
