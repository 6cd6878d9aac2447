// Shi-Tomasi corner detector (cv2.goodFeaturesToTrack) for sm_100a.
//
// Replaces  /root/reference/src/vo/features/klt.py:24-26, 87-115  KLTTracker.find_corners:
//           cv2.goodFeaturesToTrack(img, maxCorners=500, qualityLevel=0.01, minDistance=8, blockSize=7)
// (the detector of the reference's KLT tracker mode: called once per sequence and whenever fewer than 80 % of the
// features survive, klt.py:207-230).
//
// The arithmetic is OpenCV's (imgproc corner.cpp / featureselect.cpp), float for float:
//   gftt_eig_kernel        Sobel with the smoothing side scaled by 1 / (4 * block * 255) -- fused multiply-adds where
//                          OpenCV's AVX2 filter code fuses, plain ones in its scalar tails (the last W mod 32 columns of
//                          the row filter) --, float32 products, block x block box sums in float64, min eigenvalue in
//                          float32 without FMA; per-frame maximum by atomicMax.  One CTA per 64 x 32 tile, everything
//                          staged in shared memory (REFLECT_101 twice: on the source for Sobel, on the products for the
//                          box filter).
//   gftt_candidates_kernel v > (float)(max * quality) and v >= its eight neighbours, image border excluded -> list of
//                          64-bit keys (value bits << 32 | pixel index): OpenCV's order (value down, address down) is
//                          the descending order of the keys.
//   gftt_select_kernel     one CTA per frame: the greedy minimum-distance pass as a parallel maximal-independent-set
//                          iteration (a candidate is taken once every closer, higher-priority candidate is rejected, and
//                          rejected once one of them is taken), neighbours found through a grid of minDistance-sized
//                          cells; then the maxCorners best taken candidates by a radix select + shared-memory bitonic sort.
// OpenCV's float64 box sums are running sums (s += new - old along a row, SUM += / -= rows down the image); here every
// window is summed directly, which is at least as accurate: the float32 results are identical except where a double
// sum lands within 1e-16 of a rounding boundary (the tests compare the maps bit for bit on the values that matter --
// above the quality threshold -- and the corner lists exactly).
#include "common.cuh"
#include "launchers.cuh"

namespace {

constexpr int GT_W = 64, GT_H = 32, GT_THREADS = 256;
constexpr int GT_MAXR = 7;                      // box radius <= 7 (blockSize <= 15)

__device__ __forceinline__ int refl101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * n - 2 - i; }
    return i;
}

__global__ void __launch_bounds__(GT_THREADS)
gftt_eig_kernel(const uint8_t* __restrict__ img, size_t pitch, size_t frame_stride, int H, int W, int r, float k1, float k0,
                float* __restrict__ eig, int* __restrict__ frame_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int RW = GT_W + 2 * r, RH = GT_H + 2 * r;        // region of products
    const int PW = RW + 2, PH = RH + 2;                    // source patch
    const int CP = RW + 1;                                 // plane pitch
    uint8_t* sP = smem_raw;                                                        // [PH][PW]
    float* sC = reinterpret_cast<float*>(smem_raw + ((PH * PW + 15) & ~15));       // 3 planes [RH][CP]
    double* sR = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(sC) + ((3 * RH * CP * 4 + 15) & ~15));   // [3][RH][GT_W]
    const int tx0 = blockIdx.x * GT_W, ty0 = blockIdx.y * GT_H, f = blockIdx.z;
    const uint8_t* src = img + (size_t)f * frame_stride;
    const int tid = threadIdx.x;
    const int nb = (W / 32) * 32;                          // OpenCV's row filter: vector body | scalar tail
    for (int i = tid; i < PH * PW; i += GT_THREADS) {
        const int py = i / PW, px = i - py * PW;
        const int gy = refl101(ty0 - r - 1 + py, H), gx = refl101(tx0 - r - 1 + px, W);
        sP[i] = src[(size_t)gy * pitch + gx];
    }
    __syncthreads();
    for (int i = tid; i < RH * RW; i += GT_THREADS) {
        const int ry = i / RW, rx = i - ry * RW;
        const int gy = ty0 - r + ry, gx = tx0 - r + rx;
        float xx = 0.f, xy = 0.f, yy = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const uint8_t* p = sP + (ry + 1) * PW + (rx + 1);
            float rowx[3], rowy[3];
#pragma unroll
            for (int d = -1; d <= 1; d++) {
                const float a = (float)p[d * PW - 1], b = (float)p[d * PW], c = (float)p[d * PW + 1];
                rowx[d + 1] = c - a;
                rowy[d + 1] = (gx < nb) ? __fmaf_rn(c, k1, __fmaf_rn(b, k0, __fmul_rn(a, k1)))
                                        : __fadd_rn(__fadd_rn(__fmul_rn(a, k1), __fmul_rn(b, k0)), __fmul_rn(c, k1));
            }
            const float dx = __fmaf_rn(__fadd_rn(rowx[0], rowx[2]), k1, __fmul_rn(rowx[1], k0));
            const float dy = __fsub_rn(rowy[2], rowy[0]);
            xx = __fmul_rn(dx, dx); xy = __fmul_rn(dx, dy); yy = __fmul_rn(dy, dy);
        }
        sC[ry * CP + rx] = xx; sC[RH * CP + ry * CP + rx] = xy; sC[2 * RH * CP + ry * CP + rx] = yy;
    }
    __syncthreads();
    // row sums (float64) for every region row that lies inside the image
    for (int i = tid; i < RH * GT_W; i += GT_THREADS) {
        const int ry = i / GT_W, c = i - ry * GT_W;
        const int gx = tx0 + c;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        if (gx < W) {
            for (int k = -r; k <= r; k++) {
                const int rx = refl101(gx + k, W) - (tx0 - r);
                s0 += (double)sC[ry * CP + rx]; s1 += (double)sC[RH * CP + ry * CP + rx]; s2 += (double)sC[2 * RH * CP + ry * CP + rx];
            }
        }
        sR[i] = s0; sR[RH * GT_W + i] = s1; sR[2 * RH * GT_W + i] = s2;
    }
    __syncthreads();
    float vmax = 0.f;
    for (int i = tid; i < GT_H * GT_W; i += GT_THREADS) {
        const int oy = i / GT_W, c = i - oy * GT_W;
        const int gy = ty0 + oy, gx = tx0 + c;
        if (gy >= H || gx >= W) continue;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int k = -r; k <= r; k++) {
            const int ry = refl101(gy + k, H) - (ty0 - r);
            s0 += sR[ry * GT_W + c]; s1 += sR[RH * GT_W + ry * GT_W + c]; s2 += sR[2 * RH * GT_W + ry * GT_W + c];
        }
        const float a = __fmul_rn((float)s0, 0.5f), b = (float)s1, cc = __fmul_rn((float)s2, 0.5f);
        const float t = __fsub_rn(a, cc);
        const float e = __fsub_rn(__fadd_rn(a, cc), __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), __fmul_rn(b, b))));
        eig[((size_t)f * H + gy) * W + gx] = e;
        vmax = fmaxf(vmax, e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if ((tid & 31) == 0 && vmax > 0.f) atomicMax(frame_max + f, __float_as_int(vmax));
}

__global__ void __launch_bounds__(256)
gftt_candidates_kernel(const float* __restrict__ eig, int H, int W, const int* __restrict__ frame_max, double quality,
                       unsigned long long* __restrict__ keys, unsigned int* __restrict__ counts, unsigned int cap) {
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < 1 || x >= W - 1 || y < 1 || y >= H - 1) return;
    const float mx = __int_as_float(frame_max[f]);
    const float thr = (float)((double)mx * quality);
    const float* e = eig + (size_t)f * H * W;
    const float v = e[(size_t)y * W + x];
    if (!(v > thr)) return;
    float m = v;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) m = fmaxf(m, e[(size_t)(y + dy) * W + x + dx]);
    if (v != m) return;
    const unsigned int slot = atomicAdd(counts + f, 1u);
    if (slot < cap) keys[(size_t)f * cap + slot] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned int)(y * W + x);
}

struct SelArgs {
    const unsigned long long* keys; const unsigned int* counts; unsigned int cap;
    int H, W, max_corners; double md2; int cell, gw, gh;
    unsigned int* cell_start;     // [F][gw * gh + 1]
    unsigned int* cell_items;     // [F][cap]
    unsigned char* state;         // [F][cap]  0 undecided, 1 taken, 2 rejected
    float* out_xy; int* out_n; unsigned int* stats;
};

constexpr int SEL_THREADS = 1024;
__global__ void __launch_bounds__(SEL_THREADS)
gftt_select_kernel(SelArgs A) {
    __shared__ unsigned int s_hist[256];
    __shared__ unsigned long long s_sort[1024];
    __shared__ unsigned int s_flag, s_n_taken;
    __shared__ unsigned long long s_prefix;
    const int f = blockIdx.x, tid = threadIdx.x;
    unsigned int n = A.counts[f];
    const bool overflow = n > A.cap;
    if (overflow) n = A.cap;
    const unsigned long long* keys = A.keys + (size_t)f * A.cap;
    const int ncell = A.gw * A.gh;
    unsigned int* cstart = A.cell_start + (size_t)f * (ncell + 1);
    unsigned int* citems = A.cell_items + (size_t)f * A.cap;
    unsigned char* state = A.state + (size_t)f * A.cap;
    // ---- grid of cells: counting sort of the candidates by cell
    for (int i = tid; i <= ncell; i += SEL_THREADS) cstart[i] = 0;
    __syncthreads();
    for (unsigned int i = tid; i < n; i += SEL_THREADS) {
        const unsigned int idx = (unsigned int)keys[i];
        const int y = idx / A.W, x = idx - y * A.W;
        atomicAdd(cstart + (y / A.cell) * A.gw + x / A.cell + 1, 1u);
        state[i] = 0;
    }
    __syncthreads();
    if (tid == 0) {                                         // exclusive prefix (a few thousand cells)
        unsigned int acc = 0;
        for (int i = 1; i <= ncell; i++) { acc += cstart[i]; cstart[i] = acc; }
    }
    __syncthreads();
    // fill: cstart[c + 1] currently holds the END of cell c; walk the ends down to the starts
    for (unsigned int i = tid; i < n; i += SEL_THREADS) {
        const unsigned int idx = (unsigned int)keys[i];
        const int y = idx / A.W, x = idx - y * A.W;
        const int c = (y / A.cell) * A.gw + x / A.cell;
        const unsigned int pos = atomicSub(cstart + c + 1, 1u) - 1u;
        citems[pos] = i;
    }
    __syncthreads();
    // now cstart[c + 1] == start of cell c; cell c spans [cstart[c + 1], c + 1 < ncell ? cstart[c + 2] : n)
    // ---- maximal independent set in priority order
    unsigned int rounds = 0;
    for (;;) {
        if (tid == 0) s_flag = 0;
        __syncthreads();
        for (unsigned int i = tid; i < n; i += SEL_THREADS) {
            if (state[i] != 0) continue;
            const unsigned long long ki = keys[i];
            const unsigned int idx = (unsigned int)ki;
            const int y = idx / A.W, x = idx - y * A.W;
            const int cx = x / A.cell, cy = y / A.cell;
            bool blocked = false, waiting = false;
            for (int yy = max(cy - 1, 0); yy <= min(cy + 1, A.gh - 1) && !blocked; yy++)
                for (int xx = max(cx - 1, 0); xx <= min(cx + 1, A.gw - 1) && !blocked; xx++) {
                    const int c = yy * A.gw + xx;
                    const unsigned int b = cstart[c + 1], e = (c + 1 < ncell) ? cstart[c + 2] : n;
                    for (unsigned int q = b; q < e; q++) {
                        const unsigned int j = citems[q];
                        const unsigned long long kj = keys[j];
                        if (kj <= ki) continue;                                    // lower priority (or itself)
                        const unsigned int jdx = (unsigned int)kj;
                        const int jy = jdx / A.W, jx = jdx - jy * A.W;
                        const float dx = (float)(x - jx), dy = (float)(y - jy);
                        if (!((double)(dx * dx + dy * dy) < A.md2)) continue;     // featureselect.cpp: float sum against the squared double
                        const unsigned char sj = state[j];
                        if (sj == 1) { blocked = true; break; }
                        if (sj == 0) waiting = true;
                    }
                }
            // decisions of this round become visible in the next one (states are read and written in the same sweep, but a
            // candidate only ever moves from undecided to its final state, and both rules stay valid under early visibility)
            if (blocked) { state[i] = 2; s_flag = 1; }
            else if (!waiting) { state[i] = 1; s_flag = 1; }
        }
        __syncthreads();
        rounds++;
        const unsigned int progressed = s_flag;
        __syncthreads();
        if (!progressed) break;
    }
    // ---- the max_corners best taken candidates: radix select on the 64-bit keys, then a bitonic sort
    if (tid == 0) s_n_taken = 0;
    __syncthreads();
    {
        unsigned int c = 0;
        for (unsigned int i = tid; i < n; i += SEL_THREADS) c += state[i] == 1;
        if (c) atomicAdd(&s_n_taken, c);
    }
    __syncthreads();
    const unsigned int n_taken = s_n_taken;
    const unsigned int want = (A.max_corners > 0 && (unsigned)A.max_corners < n_taken) ? (unsigned)A.max_corners : n_taken;
    const unsigned int K = want > 1024 ? 1024 : want;       // output capacity of this kernel (max_corners <= 1024)
    unsigned long long prefix = 0, mask = 0;
    if (n_taken > K) {
        unsigned int remaining = K;                         // find the K-th largest key among the taken ones
        for (int shift = 56; shift >= 0; shift -= 8) {
            if (tid < 256) s_hist[tid] = 0;
            __syncthreads();
            for (unsigned int i = tid; i < n; i += SEL_THREADS)
                if (state[i] == 1 && (keys[i] & mask) == prefix) atomicAdd(&s_hist[(keys[i] >> shift) & 255], 1u);
            __syncthreads();
            if (tid == 0) {
                unsigned int acc = 0; int b = 255;
                for (; b >= 0; b--) { if (acc + s_hist[b] >= remaining) break; acc += s_hist[b]; }
                s_prefix = prefix | ((unsigned long long)b << shift);
                s_flag = remaining - acc;
            }
            __syncthreads();
            prefix = s_prefix; remaining = s_flag; mask |= 0xffull << shift;
            __syncthreads();
        }
    }
    // gather keys >= threshold (prefix is the K-th largest key when a selection ran, 0 otherwise)
    if (tid == 0) s_flag = 0;
    s_sort[tid] = 0ull;
    __syncthreads();
    for (unsigned int i = tid; i < n; i += SEL_THREADS)
        if (state[i] == 1 && keys[i] >= prefix) { const unsigned int p = atomicAdd(&s_flag, 1u); if (p < 1024) s_sort[p] = keys[i]; }
    __syncthreads();
    for (unsigned int k = 2; k <= 1024; k <<= 1)             // descending bitonic sort of 1024 slots (zeros sink to the end)
        for (unsigned int j = k >> 1; j > 0; j >>= 1) {
            const unsigned int ixj = tid ^ j;
            if (ixj > (unsigned)tid) {
                const unsigned long long a = s_sort[tid], b = s_sort[ixj];
                const bool desc = (tid & k) == 0;
                if (desc ? (a < b) : (a > b)) { s_sort[tid] = b; s_sort[ixj] = a; }
            }
            __syncthreads();
        }
    if ((unsigned)tid < K) {
        const unsigned int idx = (unsigned int)s_sort[tid];
        const int y = idx / A.W, x = idx - y * A.W;
        A.out_xy[((size_t)f * A.max_corners + tid) * 2] = (float)x;
        A.out_xy[((size_t)f * A.max_corners + tid) * 2 + 1] = (float)y;
    }
    if (tid == 0) {
        A.out_n[f] = (int)K;
        if (A.stats) { A.stats[f * 4] = n; A.stats[f * 4 + 1] = n_taken; A.stats[f * 4 + 2] = rounds; A.stats[f * 4 + 3] = overflow ? 1u : 0u; }
    }
}

}  // namespace

static size_t gftt_eig_smem(int r) {
    const int RW = GT_W + 2 * r, RH = GT_H + 2 * r, PW = RW + 2, PH = RH + 2, CP = RW + 1;
    return ((size_t)(PH * PW + 15) & ~(size_t)15) + (((size_t)3 * RH * CP * 4 + 15) & ~(size_t)15) + (size_t)3 * RH * GT_W * 8 + 16;
}

struct GfttCarve { size_t o_max, o_cnt, o_keys, o_cs, o_ci, o_st, o_stats, total; unsigned int cap; int cell, gw, gh; };
static void gftt_carve(int n_frames, int H, int W, double min_distance, GfttCarve* c) {
    const size_t F = n_frames;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    // candidates are 3x3 local maxima: at most one per 2x2 block in generic images; plateaus of equal values can exceed
    // that, so the list is capped and the overflow reported
    size_t cap = ((size_t)H * W) / 4 + 1024;
    if (cap > (1u << 20)) cap = 1u << 20;
    c->cap = (unsigned int)cap;
    int cell = (int)lrint(min_distance);
    if (cell < 1) cell = 1;
    c->cell = cell; c->gw = (W + cell - 1) / cell; c->gh = (H + cell - 1) / cell;
    c->o_max = carve(F * 4); c->o_cnt = carve(F * 4); c->o_keys = carve(F * cap * 8);
    c->o_cs = carve(F * ((size_t)c->gw * c->gh + 2) * 4); c->o_ci = carve(F * cap * 4); c->o_st = carve(F * cap);
    c->o_stats = carve(F * 16);
    c->total = off;
}

int vo_gftt_reserve(vo_ctx* ctx, int n_frames, int H, int W, double min_distance) {
    GfttCarve cv;
    gftt_carve(n_frames, H, W, min_distance, &cv);
    return vo_buf_reserve(&ctx->scratch[9], cv.total);
}

int vo_launch_gftt_eig(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch, size_t frame_stride,
                       int block_size, float* d_eig, int* d_frame_max, cudaStream_t stream) {
    VO_REQUIRE(n_frames >= 1 && H >= 3 && W >= 3 && pitch >= (size_t)W, "gftt: bad shape");
    VO_REQUIRE(block_size >= 1 && (block_size & 1) && block_size / 2 <= GT_MAXR, "gftt: blockSize must be odd and <= %d", 2 * GT_MAXR + 1);
    VO_REQUIRE(n_frames <= 65535, "gftt: at most 65535 frames per call");
    const int r = block_size / 2;
    const double scale = 1.0 / (4.0 * block_size) / 255.0;
    const float k1 = (float)((double)1.0f * scale), k0 = (float)((double)2.0f * scale);
    const size_t smem = gftt_eig_smem(r);
    if (vo_ctx_once(ctx, VO_ATTR_GFTT))
        VO_CUDA(cudaFuncSetAttribute(gftt_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gftt_eig_smem(GT_MAXR)));
    VO_CUDA(cudaMemsetAsync(d_frame_max, 0, (size_t)n_frames * 4, stream));
    dim3 g(vo_div_up(W, GT_W), vo_div_up(H, GT_H), n_frames);
    gftt_eig_kernel<<<g, GT_THREADS, smem, stream>>>(d_img, pitch, frame_stride, H, W, r, k1, k0, d_eig, d_frame_max);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    return VO_OK;
}

int vo_launch_gftt(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch, size_t frame_stride,
                   int max_corners, double quality, double min_distance, int block_size, float* d_eig, float* d_xy, int* d_n,
                   unsigned int* d_stats_or_null, cudaStream_t stream) {
    VO_REQUIRE(max_corners >= 1 && max_corners <= 1024, "gftt: maxCorners must be in [1, 1024]");
    VO_REQUIRE(quality > 0.0 && min_distance >= 1.0, "gftt: qualityLevel must be positive and minDistance >= 1");
    VO_REQUIRE((size_t)H * W < (1ull << 31), "gftt: frame too large");
    GfttCarve cv;
    gftt_carve(n_frames, H, W, min_distance, &cv);
    int rc = vo_buf_reserve(&ctx->scratch[9], cv.total, stream);
    if (rc) return rc;
    unsigned char* b = (unsigned char*)ctx->scratch[9].p;
    if ((rc = vo_launch_gftt_eig(ctx, d_img, n_frames, H, W, pitch, frame_stride, block_size, d_eig, (int*)(b + cv.o_max), stream))) return rc;
    VO_CUDA(cudaMemsetAsync(b + cv.o_cnt, 0, (size_t)n_frames * 4, stream));
    dim3 g(vo_div_up(W, 256), H, n_frames);
    gftt_candidates_kernel<<<g, 256, 0, stream>>>(d_eig, H, W, (const int*)(b + cv.o_max), quality, (unsigned long long*)(b + cv.o_keys),
                                                  (unsigned int*)(b + cv.o_cnt), cv.cap);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    SelArgs a;
    a.keys = (const unsigned long long*)(b + cv.o_keys); a.counts = (const unsigned int*)(b + cv.o_cnt); a.cap = cv.cap;
    a.H = H; a.W = W; a.max_corners = max_corners; a.md2 = min_distance * min_distance;
    a.cell = cv.cell; a.gw = cv.gw; a.gh = cv.gh;
    a.cell_start = (unsigned int*)(b + cv.o_cs); a.cell_items = (unsigned int*)(b + cv.o_ci); a.state = b + cv.o_st;
    a.out_xy = d_xy; a.out_n = d_n; a.stats = d_stats_or_null ? d_stats_or_null : (unsigned int*)(b + cv.o_stats);
    gftt_select_kernel<<<n_frames, SEL_THREADS, 0, stream>>>(a);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    return VO_OK;
}
