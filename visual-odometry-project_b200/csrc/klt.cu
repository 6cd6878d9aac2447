// Pyramidal Lucas-Kanade tracker for sm_100a: one warp per keypoint.
//
// Replaces  /root/reference/src/vo/features/klt.py:233-239  cv2.calcOpticalFlowPyrLK(winSize=(17,17),
//           maxLevel=2, criteria=(EPS|COUNT, 10, 0.03))  and the status/error outputs used at klt.py:244-249.
//
// The algorithm is OpenCV's (video/src/lkpyramid.cpp, imgproc pyrDown): reflect-101 Gaussian
// pyramid, Scharr derivatives, 14-bit fixed-point bilinear patches, 2x2 normal equations, Newton
// steps with the EPS / oscillation stopping rules.  Per warp: the (win+3)^2 neighbourhood of the
// point is staged in shared memory, lanes stride over the win^2 window pixels, the Gram matrix
// and mismatch vector are reduced with warp shuffles as exact integers, and every lane carries
// the (uniform) float32 iteration state in OpenCV's operation order (--fmad=false).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int W_BITS = 14;
constexpr int KLT_MAX_LEVELS = 8;
constexpr int KLT_WARPS = 4;

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}

// cv2.pyrDown (uint8, 1 channel, BORDER_REFLECT_101): separable [1 4 6 4 1], streamed by warps, nothing in
// shared memory.  A lane owns one aligned 32-bit word of an input row (4 pixels = 2 outputs); the two horizontal
// sums are dp4a([1 4 6 4]) on the word (or on bytes funnelled in from the left neighbour by shuffle) plus the
// fifth tap; the warp walks down the rows keeping the five row sums of each output in registers, and rounds
// (sum + 128) >> 8 exactly like OpenCV's integer path.  The first and last lane only feed their neighbours
// (strips overlap by one word on each side), words that straddle the image border are assembled from
// reflected byte loads.
constexpr int PD_WARPS = 8, PD_ROWS = 24, PD_OUT_W = 60;   // 30 words -> 60 outputs per warp, 24 output rows
__global__ void __launch_bounds__(PD_WARPS * 32)
pyr_down_kernel(const uint8_t* __restrict__ src, int H, int W, size_t spitch, size_t sframe,
                uint8_t* __restrict__ dst, int dh, int dw, size_t dpitch, size_t dframe) {
    constexpr unsigned int FULL = 0xFFFFFFFFu, TAPS = 0x04060401u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x * PD_WARPS + warp;
    if (strip * PD_OUT_W >= dw) return;                   // warp-uniform
    const int wi = 30 * strip - 1 + lane;                 // this lane's word of the input row
    const int bx = 4 * wi;
    const bool fast = bx >= 0 && bx + 3 < W;
    const uint8_t* s = src + (size_t)blockIdx.z * sframe;
    uint8_t* d = dst + (size_t)blockIdx.z * dframe;
    const int oy0 = blockIdx.y * PD_ROWS, oy1 = min(dh, oy0 + PD_ROWS);
    int rx[4];
#pragma unroll
    for (int k = 0; k < 4; k++) rx[k] = reflect101(bx + k, W);
    auto load_word = [&](int iy) -> unsigned int {
        const uint8_t* r = s + (size_t)reflect101(iy, H) * spitch;
        if (fast) return *reinterpret_cast<const unsigned int*>(r + bx);
        return (unsigned int)r[rx[0]] | ((unsigned int)r[rx[1]] << 8) | ((unsigned int)r[rx[2]] << 16) | ((unsigned int)r[rx[3]] << 24);
    };
    auto hsum = [&](unsigned int w, int& ha, int& hb) {    // row sums of outputs 2 wi and 2 wi + 1
        const unsigned int prev = __shfl_up_sync(FULL, w, 1), next = __shfl_down_sync(FULL, w, 1);
        ha = (int)__dp4a(__byte_perm(prev, w, 0x5432), TAPS, (w >> 16) & 0xFFu);
        hb = (int)__dp4a(w, TAPS, next & 0xFFu);
    };
    int a0, a1, a2, a3, a4, b0, b1, b2, b3, b4;
    {
        const unsigned int w0 = load_word(2 * oy0 - 2), w1 = load_word(2 * oy0 - 1), w2 = load_word(2 * oy0);
        hsum(w0, a0, b0); hsum(w1, a1, b1); hsum(w2, a2, b2);
    }
    const bool writes = lane >= 1 && lane <= 30 && 2 * wi < dw;
    for (int y = oy0; y < oy1; y++) {
        const unsigned int w3 = load_word(2 * y + 1), w4 = load_word(2 * y + 2);
        hsum(w3, a3, b3); hsum(w4, a4, b4);
        const unsigned int oa = (unsigned int)(a0 + a4 + 4 * (a1 + a3) + 6 * a2 + 128) >> 8;
        const unsigned int ob = (unsigned int)(b0 + b4 + 4 * (b1 + b3) + 6 * b2 + 128) >> 8;
        if (writes) {
            uint8_t* o = d + (size_t)y * dpitch + 2 * wi;
            if (2 * wi + 1 < dw) *reinterpret_cast<unsigned short*>(o) = (unsigned short)(oa | (ob << 8));
            else *o = (uint8_t)oa;
        }
        a0 = a2; a1 = a3; a2 = a4; b0 = b2; b1 = b3; b2 = b4;
    }
}

// frame -> level-0 slot.  Destination rows are 16-byte aligned (pitch % 16 == 0); each thread assembles one
// 16-byte store from byte loads when the source row is not aligned the same way (tightly packed 1241-byte rows).
__global__ void __launch_bounds__(256)
copy_level0_kernel(const uint8_t* __restrict__ src, int H, int W, size_t spitch, size_t sframe,
                   uint8_t* __restrict__ dst, size_t dpitch, size_t dframe) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;      // a block covers several rows when a row needs few threads
    if (x >= W || y >= H) return;
    const uint8_t* s = src + (size_t)blockIdx.z * sframe + (size_t)y * spitch + x;
    uint8_t* d = dst + (size_t)blockIdx.z * dframe + (size_t)y * dpitch + x;
    if (x + 16 <= W) {
        uint4 v;
        if ((reinterpret_cast<uintptr_t>(s) & 15) == 0) {
            v = *reinterpret_cast<const uint4*>(s);
        } else {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
                w[k] = (uint32_t)s[4 * k] | ((uint32_t)s[4 * k + 1] << 8) | ((uint32_t)s[4 * k + 2] << 16) | ((uint32_t)s[4 * k + 3] << 24);
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        *reinterpret_cast<uint4*>(d) = v;
    } else {
        for (int k = 0; x + k < W; k++) d[k] = s[k];
    }
}

// cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) for 8-bit images (klt.py:57-62, 84-85): OpenCV's 15-bit fixed point,
// gray = (B * 3735 + G * 19235 + R * 9798 + 2^14) >> 15 (checked against cv2 on all 2^24 colours).  A thread converts
// four pixels: 12 interleaved bytes in, one 32-bit word out (destination rows are 16-byte aligned).
__global__ void __launch_bounds__(256)
bgr2gray_kernel(const uint8_t* __restrict__ src, int H, int W, size_t spitch, size_t sframe, uint8_t* __restrict__ dst,
                size_t dpitch, size_t dframe) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x >= W) return;
    const uint8_t* s = src + (size_t)blockIdx.z * sframe + (size_t)y * spitch + (size_t)x * 3;
    uint8_t* d = dst + (size_t)blockIdx.z * dframe + (size_t)y * dpitch + x;
    const int n = min(4, W - x);
    uint8_t px[12];
    if (n == 4 && (reinterpret_cast<uintptr_t>(s) & 3) == 0) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(s);
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
#pragma unroll
        for (int k = 0; k < 4; k++) { px[k] = (w0 >> (8 * k)) & 255; px[4 + k] = (w1 >> (8 * k)) & 255; px[8 + k] = (w2 >> (8 * k)) & 255; }
    } else {
        for (int k = 0; k < 3 * n; k++) px[k] = s[k];
    }
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t g = (px[3 * k] * 3735u + px[3 * k + 1] * 19235u + px[3 * k + 2] * 9798u + (1u << 14)) >> 15;
        out |= (k < n ? g : 0u) << (8 * k);
    }
    if (n == 4) *reinterpret_cast<uint32_t*>(d) = out;
    else for (int k = 0; k < n; k++) d[k] = (out >> (8 * k)) & 255;
}

struct PyrLayout {
    int n_levels;
    int h[KLT_MAX_LEVELS], w[KLT_MAX_LEVELS];
    size_t pitch[KLT_MAX_LEVELS], offset[KLT_MAX_LEVELS];
    size_t frame_bytes;
};

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// Stage the reflect-101 extended image region [x0, x0+n) x [y0, y0+n) into shared memory.
__device__ __forceinline__ void stage_patch(const uint8_t* __restrict__ img, int rows, int cols, size_t pitch,
                                            int x0, int y0, int n, uint8_t* sm, int lane) {
    for (int i = lane; i < n * n; i += 32) {
        const int py = i / n, px = i - py * n;
        sm[i] = img[(size_t)reflect101(y0 + py, rows) * pitch + reflect101(x0 + px, cols)];
    }
}

__global__ void __launch_bounds__(KLT_WARPS * 32)
klt_track_kernel(const uint8_t* __restrict__ pyr_prev, const uint8_t* __restrict__ pyr_next, PyrLayout lay,
                 int win, int max_iters, double eps2, double min_eig, const float* __restrict__ prev_pts, int n_pts,
                 float* __restrict__ next_pts, uint8_t* __restrict__ status, float* __restrict__ err,
                const int* __restrict__ n_valid) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pt = blockIdx.x * KLT_WARPS + warp;
    const int f = blockIdx.y;
    if (pt >= n_pts || (n_valid && pt >= n_valid[f])) return;   // per-frame row count of a resident feature table
    const int w2 = win * win;
    const int pn = win + 3;                 // staged I neighbourhood
    const int dn = win + 1;                 // integer positions where derivatives / J are needed
    // per-warp shared memory slices
    const size_t per_warp = (size_t)((pn * pn + 15) & ~15) + (size_t)dn * dn * 4 + (size_t)w2 * 2 + (size_t)w2 * 4;
    unsigned char* base = smem_raw + (size_t)warp * ((per_warp + 15) & ~(size_t)15);
    uint8_t* patch = base;                                                     // [pn*pn] u8 (I), reused for J [dn*dn]
    short* dpatch = reinterpret_cast<short*>(base + ((pn * pn + 15) & ~15));  // [dn*dn][2]
    short* Iw = dpatch + dn * dn * 2;                                          // [w2]
    short* dIw = Iw + w2;                                                      // [w2][2]

    const uint8_t* Ip = pyr_prev + (size_t)f * lay.frame_bytes;
    const uint8_t* Jp = pyr_next + (size_t)f * lay.frame_bytes;
    const size_t pidx = (size_t)f * n_pts + pt;
    const float px0 = prev_pts[2 * pidx], py0 = prev_pts[2 * pidx + 1];
    const float half = (float)(win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const int top = lay.n_levels - 1;
    bool st = true;
    float e_out = 0.f;
    float outx = 0.f, outy = 0.f;  // nextPts[ptidx]

    for (int level = top; level >= 0; level--) {
        const int cols = lay.w[level], rows = lay.h[level];
        const size_t pitch = lay.pitch[level];
        const uint8_t* I = Ip + lay.offset[level];
        const uint8_t* J = Jp + lay.offset[level];
        const float sc = (float)(1. / (1 << level));
        float prx = px0 * sc, pry = py0 * sc;
        float nx, ny;
        if (level == top) { nx = prx; ny = pry; }
        else { nx = outx * 2.f; ny = outy * 2.f; }
        outx = nx; outy = ny;
        prx -= half; pry -= half;
        const int ipx = (int)floorf(prx), ipy = (int)floorf(pry);
        if (ipx < -win || ipx >= cols || ipy < -win || ipy >= rows) {
            if (level == 0) { st = false; e_out = 0.f; }
            continue;
        }
        float a = prx - ipx, b = pry - ipy;
        int iw00 = __float2int_rn((1.f - a) * (1.f - b) * (1 << W_BITS));
        int iw01 = __float2int_rn(a * (1.f - b) * (1 << W_BITS));
        int iw10 = __float2int_rn((1.f - a) * b * (1 << W_BITS));
        int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;

        __syncwarp();
        stage_patch(I, rows, cols, pitch, ipx - 1, ipy - 1, pn, patch, lane);
        __syncwarp();
        // Scharr derivatives at the (win+1)^2 integer positions; zero outside the image
        for (int i = lane; i < dn * dn; i += 32) {
            const int qy = i / dn, qx = i - qy * dn;
            const int gx = ipx + qx, gy = ipy + qy;
            int dx = 0, dy = 0;
            if (gx >= 0 && gx < cols && gy >= 0 && gy < rows) {
                const uint8_t* r0 = patch + qy * pn + qx;  // top-left of the 3x3 neighbourhood
                const uint8_t* r1 = r0 + pn;
                const uint8_t* r2 = r1 + pn;
                const int t0m = ((int)r0[0] + (int)r2[0]) * 3 + (int)r1[0] * 10;
                const int t0p = ((int)r0[2] + (int)r2[2]) * 3 + (int)r1[2] * 10;
                const int t1m = (int)r2[0] - (int)r0[0], t1c = (int)r2[1] - (int)r0[1], t1p = (int)r2[2] - (int)r0[2];
                dx = t0p - t0m;
                dy = (t1p + t1m) * 3 + t1c * 10;
            }
            dpatch[2 * i] = (short)dx;
            dpatch[2 * i + 1] = (short)dy;
        }
        __syncwarp();
        int sA11 = 0, sA12 = 0, sA22 = 0;
        for (int i = lane; i < w2; i += 32) {
            const int y = i / win, x = i - y * win;
            const uint8_t* s0 = patch + (y + 1) * pn + (x + 1);
            const int ival = descale((int)s0[0] * iw00 + (int)s0[1] * iw01 + (int)s0[pn] * iw10 + (int)s0[pn + 1] * iw11,
                                     W_BITS - 5);
            const short* d0 = dpatch + 2 * (y * dn + x);
            const short* d1 = d0 + 2 * dn;
            const int ixval = descale((int)d0[0] * iw00 + (int)d0[2] * iw01 + (int)d1[0] * iw10 + (int)d1[2] * iw11, W_BITS);
            const int iyval = descale((int)d0[1] * iw00 + (int)d0[3] * iw01 + (int)d1[1] * iw10 + (int)d1[3] * iw11, W_BITS);
            Iw[i] = (short)ival;
            dIw[2 * i] = (short)ixval;
            dIw[2 * i + 1] = (short)iyval;
            sA11 += ixval * ixval;
            sA12 += ixval * iyval;
            sA22 += iyval * iyval;
        }
        const long long iA11 = warp_sum_ll(sA11), iA12 = warp_sum_ll(sA12), iA22 = warp_sum_ll(sA22);
        const float A11 = (float)iA11 * FLT_SCALE, A12 = (float)iA12 * FLT_SCALE, A22 = (float)iA22 * FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
        if ((double)minEig < min_eig || D < 1.1920928955078125e-07f) {
            if (level == 0) st = false;
            continue;
        }
        D = 1.f / D;
        nx -= half; ny -= half;
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < max_iters; j++) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -win || inx >= cols || iny < -win || iny >= rows) {
                if (level == 0) st = false;
                break;
            }
            a = nx - inx; b = ny - iny;
            iw00 = __float2int_rn((1.f - a) * (1.f - b) * (1 << W_BITS));
            iw01 = __float2int_rn(a * (1.f - b) * (1 << W_BITS));
            iw10 = __float2int_rn((1.f - a) * b * (1 << W_BITS));
            iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            __syncwarp();
            stage_patch(J, rows, cols, pitch, inx, iny, dn, patch, lane);
            __syncwarp();
            int sb1 = 0, sb2 = 0;
            for (int i = lane; i < w2; i += 32) {
                const int y = i / win, x = i - y * win;
                const uint8_t* s0 = patch + y * dn + x;
                const int diff = descale((int)s0[0] * iw00 + (int)s0[1] * iw01 + (int)s0[dn] * iw10 + (int)s0[dn + 1] * iw11,
                                         W_BITS - 5) - (int)Iw[i];
                sb1 += diff * (int)dIw[2 * i];
                sb2 += diff * (int)dIw[2 * i + 1];
            }
            const long long ib1 = warp_sum_ll(sb1), ib2 = warp_sum_ll(sb2);
            const float b1 = (float)ib1 * FLT_SCALE, b2 = (float)ib2 * FLT_SCALE;
            const float dx = (float)((A12 * b2 - A22 * b1) * D);
            const float dy = (float)((A12 * b1 - A11 * b2) * D);
            nx += dx; ny += dy;
            outx = nx + half; outy = ny + half;
            if ((double)dx * dx + (double)dy * dy <= eps2) break;
            if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
                outx -= dx * 0.5f;
                outy -= dy * 0.5f;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (st && level == 0) {
            const float fx = outx - half, fy = outy - half;
            const int inx = (int)floorf(fx), iny = (int)floorf(fy);
            if (inx < -win || inx >= cols || iny < -win || iny >= rows) {
                st = false;
                continue;
            }
            const float aa = fx - inx, bb = fy - iny;
            iw00 = __float2int_rn((1.f - aa) * (1.f - bb) * (1 << W_BITS));
            iw01 = __float2int_rn(aa * (1.f - bb) * (1 << W_BITS));
            iw10 = __float2int_rn((1.f - aa) * bb * (1 << W_BITS));
            iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            __syncwarp();
            stage_patch(J, rows, cols, pitch, inx, iny, dn, patch, lane);
            __syncwarp();
            int se = 0;
            for (int i = lane; i < w2; i += 32) {
                const int y = i / win, x = i - y * win;
                const uint8_t* s0 = patch + y * dn + x;
                const int diff = descale((int)s0[0] * iw00 + (int)s0[1] * iw01 + (int)s0[dn] * iw10 + (int)s0[dn + 1] * iw11,
                                         W_BITS - 5) - (int)Iw[i];
                se += diff < 0 ? -diff : diff;
            }
            const long long e = warp_sum_ll(se);
            e_out = (float)e * 1.f / (float)(32 * win * win);
        }
    }
    if (lane == 0) {
        next_pts[2 * pidx] = outx;
        next_pts[2 * pidx + 1] = outy;
        status[pidx] = st ? 1 : 0;
        err[pidx] = e_out;
    }
}


// ---------------------------------------------------------------------------------------------
// Fast path: window size known at compile time.  Same arithmetic as klt_track_kernel (the tests
// compare both with the oracle bit for bit); what changes is how the warp gets there:
//   * window values (I, Ix, Iy) live in registers, WIN*WIN/32 (+1) pixels per lane, static indexing;
//   * staging is division-free (compile-time divisors) with a branch-free single reflection
//     (pyramid levels are larger than the window, so one reflection is enough);
//   * warp sums use the REDUX unit on 16-bit halves (exact) instead of 64-bit shuffle trees;
//   * the J patch is re-staged only when the integer window position moves.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect1(int i, int n) {   // reflect-101 for |overshoot| <= n + 1 (levels are wider than the window)
    i = i < 0 ? -i : i;
    i = i >= n ? 2 * n - 2 - i : i;
    return i < 0 ? -i : i;
}

__device__ __forceinline__ long long warp_sum_exact(int v) {
    const int lo = v & 0xFFFF, hi = v >> 16;               // v == hi * 65536 + lo, also for negative v
    const int slo = __reduce_add_sync(0xFFFFFFFFu, lo);
    const int shi = __reduce_add_sync(0xFFFFFFFFu, hi);
    return (long long)shi * 65536ll + (long long)slo;
}

template <int N>
__device__ __forceinline__ void stage_patch_fast(const uint8_t* __restrict__ img, int rows, int cols, size_t pitch,
                                                 int x0, int y0, uint8_t* sm, int lane) {
    // element i = lane + 32 t -> (py, px), advanced without division; loads are batched (all issued, then stored)
    constexpr int T = (N * N + 31) / 32;
    int py = lane / N, px = lane - py * N;
    const bool inside = x0 >= 0 && y0 >= 0 && x0 + N <= cols && y0 + N <= rows;   // warp-uniform
    uint8_t v[T];
    if (inside) {
        const uint8_t* base = img + (size_t)y0 * pitch + x0;
#pragma unroll
        for (int t = 0; t < T; t++) {
            v[t] = (t * 32 + lane < N * N) ? base[(size_t)py * pitch + px] : (uint8_t)0;
            px += 32 % N; py += 32 / N;
            if (px >= N) { px -= N; py++; }
        }
    } else {
#pragma unroll
        for (int t = 0; t < T; t++) {
            v[t] = (t * 32 + lane < N * N) ? img[(size_t)reflect1(y0 + py, rows) * pitch + reflect1(x0 + px, cols)] : (uint8_t)0;
            px += 32 % N; py += 32 / N;
            if (px >= N) { px -= N; py++; }
        }
    }
#pragma unroll
    for (int t = 0; t < T; t++)
        if (t * 32 + lane < N * N) sm[t * 32 + lane] = v[t];
}

template <int WIN>
__global__ void __launch_bounds__(KLT_WARPS * 32, 4)
klt_track_fast(const uint8_t* __restrict__ pyr_prev, const uint8_t* __restrict__ pyr_next, PyrLayout lay,
               int max_iters, double eps2, double min_eig, const float* __restrict__ prev_pts, int n_pts,
               float* __restrict__ next_pts, uint8_t* __restrict__ status, float* __restrict__ err,
                const int* __restrict__ n_valid) {
    constexpr int W2 = WIN * WIN, PN = WIN + 3, DN = WIN + 1;
    constexpr int PPL = (W2 + 31) / 32;                        // window pixels per lane
    constexpr int PATCH_BYTES = (PN * PN + 15) & ~15;
    constexpr int PER_WARP = (PATCH_BYTES + DN * DN * 4 + 15) & ~15;
    __shared__ __align__(16) unsigned char smem[KLT_WARPS * PER_WARP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pt = blockIdx.x * KLT_WARPS + warp;
    const int f = blockIdx.y;
    if (pt >= n_pts || (n_valid && pt >= n_valid[f])) return;   // per-frame row count of a resident feature table
    uint8_t* patch = smem + warp * PER_WARP;                                  // I (PN^2), later J (DN^2)
    short* dpatch = reinterpret_cast<short*>(patch + PATCH_BYTES);             // [DN*DN][2]

    const uint8_t* Ip = pyr_prev + (size_t)f * lay.frame_bytes;
    const uint8_t* Jp = pyr_next + (size_t)f * lay.frame_bytes;
    const size_t pidx = (size_t)f * n_pts + pt;
    const float px0 = prev_pts[2 * pidx], py0 = prev_pts[2 * pidx + 1];
    const float half = (float)(WIN - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const int top = lay.n_levels - 1;
    bool st = true;
    float e_out = 0.f;
    float outx = 0.f, outy = 0.f;
    // window pixel k of this lane: offsets into the J patch (DN wide) and the I patch (PN wide); -1 = none
    int offJ[PPL], offI[PPL];
#pragma unroll
    for (int k = 0; k < PPL; k++) {
        const int i = k * 32 + lane;
        const int y = i / WIN, x = i - y * WIN;
        offJ[k] = i < W2 ? y * DN + x : -1;
        offI[k] = (y + 1) * PN + (x + 1);
    }

    for (int level = top; level >= 0; level--) {
        const int cols = lay.w[level], rows = lay.h[level];
        const size_t pitch = lay.pitch[level];
        const uint8_t* I = Ip + lay.offset[level];
        const uint8_t* J = Jp + lay.offset[level];
        const float sc = (float)(1. / (1 << level));
        float prx = px0 * sc, pry = py0 * sc;
        float nx, ny;
        if (level == top) { nx = prx; ny = pry; }
        else { nx = outx * 2.f; ny = outy * 2.f; }
        outx = nx; outy = ny;
        prx -= half; pry -= half;
        const int ipx = (int)floorf(prx), ipy = (int)floorf(pry);
        if (ipx < -WIN || ipx >= cols || ipy < -WIN || ipy >= rows) {
            if (level == 0) { st = false; e_out = 0.f; }
            continue;
        }
        float a = prx - ipx, b = pry - ipy;
        int iw00 = __float2int_rn((1.f - a) * (1.f - b) * (1 << W_BITS));
        int iw01 = __float2int_rn(a * (1.f - b) * (1 << W_BITS));
        int iw10 = __float2int_rn((1.f - a) * b * (1 << W_BITS));
        int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;

        __syncwarp();
        stage_patch_fast<PN>(I, rows, cols, pitch, ipx - 1, ipy - 1, patch, lane);
        __syncwarp();
        {
            int qy = lane / DN, qx = lane - qy * DN;
#pragma unroll 2
            for (int i = lane; i < DN * DN; i += 32) {
                const int gx = ipx + qx, gy = ipy + qy;
                int dx = 0, dy = 0;
                if (gx >= 0 && gx < cols && gy >= 0 && gy < rows) {
                    const uint8_t* r0 = patch + qy * PN + qx;
                    const uint8_t* r1 = r0 + PN;
                    const uint8_t* r2 = r1 + PN;
                    const int t0m = ((int)r0[0] + (int)r2[0]) * 3 + (int)r1[0] * 10;
                    const int t0p = ((int)r0[2] + (int)r2[2]) * 3 + (int)r1[2] * 10;
                    const int t1m = (int)r2[0] - (int)r0[0], t1c = (int)r2[1] - (int)r0[1], t1p = (int)r2[2] - (int)r0[2];
                    dx = t0p - t0m;
                    dy = (t1p + t1m) * 3 + t1c * 10;
                }
                dpatch[2 * i] = (short)dx;
                dpatch[2 * i + 1] = (short)dy;
                qx += 32 % DN; qy += 32 / DN;
                if (qx >= DN) { qx -= DN; qy++; }
            }
        }
        __syncwarp();
        int Iv[PPL], Ix[PPL], Iy[PPL];
        int sA11 = 0, sA12 = 0, sA22 = 0;
#pragma unroll
        for (int k = 0; k < PPL; k++) {
            Iv[k] = 0; Ix[k] = 0; Iy[k] = 0;
            if (offJ[k] >= 0) {
                const uint8_t* s0 = patch + offI[k];
                Iv[k] = descale((int)s0[0] * iw00 + (int)s0[1] * iw01 + (int)s0[PN] * iw10 + (int)s0[PN + 1] * iw11, W_BITS - 5);
                const short* d0 = dpatch + 2 * offJ[k];
                const short* d1 = d0 + 2 * DN;
                Ix[k] = descale((int)d0[0] * iw00 + (int)d0[2] * iw01 + (int)d1[0] * iw10 + (int)d1[2] * iw11, W_BITS);
                Iy[k] = descale((int)d0[1] * iw00 + (int)d0[3] * iw01 + (int)d1[1] * iw10 + (int)d1[3] * iw11, W_BITS);
                sA11 += Ix[k] * Ix[k];
                sA12 += Ix[k] * Iy[k];
                sA22 += Iy[k] * Iy[k];
            }
        }
        const long long iA11 = warp_sum_exact(sA11), iA12 = warp_sum_exact(sA12), iA22 = warp_sum_exact(sA22);
        const float A11 = (float)iA11 * FLT_SCALE, A12 = (float)iA12 * FLT_SCALE, A22 = (float)iA22 * FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * WIN * WIN);
        if ((double)minEig < min_eig || D < 1.1920928955078125e-07f) {
            if (level == 0) st = false;
            continue;
        }
        D = 1.f / D;
        nx -= half; ny -= half;
        float pdx = 0.f, pdy = 0.f;
        int sx = 0x7fffffff, sy = 0x7fffffff;                 // integer position of the staged J patch
        for (int j = 0; j < max_iters; j++) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -WIN || inx >= cols || iny < -WIN || iny >= rows) {
                if (level == 0) st = false;
                break;
            }
            a = nx - inx; b = ny - iny;
            iw00 = __float2int_rn((1.f - a) * (1.f - b) * (1 << W_BITS));
            iw01 = __float2int_rn(a * (1.f - b) * (1 << W_BITS));
            iw10 = __float2int_rn((1.f - a) * b * (1 << W_BITS));
            iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            if (inx != sx || iny != sy) {
                __syncwarp();
                stage_patch_fast<DN>(J, rows, cols, pitch, inx, iny, patch, lane);
                __syncwarp();
                sx = inx; sy = iny;
            }
            int sb1 = 0, sb2 = 0;
#pragma unroll
            for (int k = 0; k < PPL; k++) {
                if (offJ[k] >= 0) {
                    const uint8_t* s0 = patch + offJ[k];
                    const int diff = descale((int)s0[0] * iw00 + (int)s0[1] * iw01 + (int)s0[DN] * iw10 + (int)s0[DN + 1] * iw11,
                                             W_BITS - 5) - Iv[k];
                    sb1 += diff * Ix[k];
                    sb2 += diff * Iy[k];
                }
            }
            const long long ib1 = warp_sum_exact(sb1), ib2 = warp_sum_exact(sb2);
            const float b1 = (float)ib1 * FLT_SCALE, b2 = (float)ib2 * FLT_SCALE;
            const float dx = (float)((A12 * b2 - A22 * b1) * D);
            const float dy = (float)((A12 * b1 - A11 * b2) * D);
            nx += dx; ny += dy;
            outx = nx + half; outy = ny + half;
            if ((double)dx * dx + (double)dy * dy <= eps2) break;
            if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
                outx -= dx * 0.5f;
                outy -= dy * 0.5f;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (st && level == 0) {
            const float fx = outx - half, fy = outy - half;
            const int inx = (int)floorf(fx), iny = (int)floorf(fy);
            if (inx < -WIN || inx >= cols || iny < -WIN || iny >= rows) {
                st = false;
                continue;
            }
            const float aa = fx - inx, bb = fy - iny;
            iw00 = __float2int_rn((1.f - aa) * (1.f - bb) * (1 << W_BITS));
            iw01 = __float2int_rn(aa * (1.f - bb) * (1 << W_BITS));
            iw10 = __float2int_rn((1.f - aa) * bb * (1 << W_BITS));
            iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            if (inx != sx || iny != sy) {
                __syncwarp();
                stage_patch_fast<DN>(J, rows, cols, pitch, inx, iny, patch, lane);
                __syncwarp();
            }
            int se = 0;
#pragma unroll
            for (int k = 0; k < PPL; k++) {
                if (offJ[k] >= 0) {
                    const uint8_t* s0 = patch + offJ[k];
                    const int diff = descale((int)s0[0] * iw00 + (int)s0[1] * iw01 + (int)s0[DN] * iw10 + (int)s0[DN + 1] * iw11,
                                             W_BITS - 5) - Iv[k];
                    se += diff < 0 ? -diff : diff;
                }
            }
            const long long e = warp_sum_exact(se);
            e_out = (float)e * 1.f / (float)(32 * WIN * WIN);
        }
    }
    if (lane == 0) {
        next_pts[2 * pidx] = outx;
        next_pts[2 * pidx + 1] = outy;
        status[pidx] = st ? 1 : 0;
        err[pidx] = e_out;
    }
}

// warps (keypoints) per CTA.  A CTA holds its slot until its slowest warp converges, so small CTAs lose less: 8 -> 4 gave
// 1.5 % of the step, 4 -> 1 another 0.4 % (1.6718 -> 1.6653 ms; 2 is slower than both: 1.6765 ms).  32 one-warp CTAs fill
// an SM's 64-register budget exactly.
constexpr int KLT_SWARPS = 1;


// ---------------------------------------------------------------------------------------------
// Packed variant of the compact path: identical arithmetic, but the iteration loop (the bulk of the work)
// reads nothing but registers and two fixed-offset shared-memory words per pixel: the four bilinear taps of
// every window pixel of J are packed into one 32-bit register when the J patch is (re)staged, and the
// interpolation is two dp2a instructions (16-bit weights x 8-bit pixels, exact).  Pixel i = lane + 32 t sits
// at a compile-time offset, so the unrolled loop has no index arithmetic.
// ---------------------------------------------------------------------------------------------

// (win+3)^2 = 20 x 20 patch of klt_track_packed<17>: rows in pairs x 16 columns (10 loads per lane at a running
// offset) + the last four columns (3 loads per lane); when the patch crosses the border the generic path reflects.
__device__ __forceinline__ void stage_patch20(const uint8_t* __restrict__ img, int rows, int cols, size_t pitch,
                                              int x0, int y0, uint8_t* sm, int lane) {
    constexpr int N = 20;
    if (!(x0 >= 0 && y0 >= 0 && x0 + N <= cols && y0 + N <= rows)) {       // warp-uniform
        stage_patch_fast<N>(img, rows, cols, pitch, x0, y0, sm, lane);
        return;
    }
    const uint8_t* base = img + (size_t)y0 * pitch + x0;
    const int ip = (int)pitch, hi = lane >> 4, lx = lane & 15, qy = lane >> 2, qx = 16 + (lane & 3);
    uint8_t v[13];
    int off = hi * ip + lx;
#pragma unroll
    for (int t = 0; t < 10; t++) { v[t] = base[off]; off += 2 * ip; }
    off = qy * ip + qx;
#pragma unroll
    for (int u = 0; u < 3; u++) { v[10 + u] = (8 * u + qy < N) ? base[off] : (uint8_t)0; off += 8 * ip; }
    uint8_t* s0 = sm + hi * N + lx;
#pragma unroll
    for (int t = 0; t < 10; t++) s0[t * 2 * N] = v[t];
    uint8_t* s1 = sm + qy * N + qx;
#pragma unroll
    for (int u = 0; u < 3; u++) if (8 * u + qy < N) s1[u * 8 * N] = v[10 + u];
}

// dp2a with signed 16-bit halves of a and unsigned bytes of b: c + a.lo * b.byte0 + a.hi * b.byte1 (lo) / bytes 2, 3 (hi)
__device__ __forceinline__ int dp2a_lo_su(unsigned int a, unsigned int b, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(unsigned int a, unsigned int b, int c) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int WIN>
__global__ void __launch_bounds__(KLT_SWARPS * 32, 32 / KLT_SWARPS)
klt_track_packed(const uint8_t* __restrict__ pyr_prev, const uint8_t* __restrict__ pyr_next, PyrLayout lay,
                int max_iters, double eps2, double min_eig, const float* __restrict__ prev_pts, int n_pts,
                float* __restrict__ next_pts, uint8_t* __restrict__ status, float* __restrict__ err,
                const int* __restrict__ n_valid) {
    constexpr int W2 = WIN * WIN, PN = WIN + 3, DN = WIN + 1;
    constexpr int T = (W2 + 31) / 32;                      // window pixels per lane
    constexpr int PATCH_BYTES = (PN * PN + 15) & ~15;
    constexpr int GI_BYTES = T * 32 * 8;                   // padded to whole warps: the unrolled loops read every slot
    constexpr int AN = WIN + 2;                            // side of the interpolated patch of the fast setup
    constexpr int DP_BYTES = ((AN * AN > DN * DN ? AN * AN : DN * DN) * 4 + 15) & ~15;
    constexpr int PER_WARP = (PATCH_BYTES + DP_BYTES + GI_BYTES + 15) & ~15;
    __shared__ __align__(16) unsigned char smem[KLT_SWARPS * PER_WARP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pt = blockIdx.x * KLT_SWARPS + warp;
    const int f = blockIdx.y;
    if (pt >= n_pts || (n_valid && pt >= n_valid[f])) return;   // per-frame row count of a resident feature table
    uint8_t* patch = smem + warp * PER_WARP;                                  // I (PN^2), later J (DN^2)
    short* dpatch = reinterpret_cast<short*>(patch + PATCH_BYTES);             // [DN*DN][2]
    int* At = reinterpret_cast<int*>(patch + PATCH_BYTES);                     // [AN*AN] (fast setup, same storage)
    int2* GI = reinterpret_cast<int2*>(patch + PATCH_BYTES + DP_BYTES);        // [W2]  x = Ix | Iy << 16, y = I: one 64-bit load per slot

    const uint8_t* Ip = pyr_prev + (size_t)f * lay.frame_bytes;
    const uint8_t* Jp = pyr_next + (size_t)f * lay.frame_bytes;
    const size_t pidx = (size_t)f * n_pts + pt;
    const float px0 = prev_pts[2 * pidx], py0 = prev_pts[2 * pidx + 1];
    const float half = (float)(WIN - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const int top = lay.n_levels - 1;
    // Window pixel of (slot t, lane): slots 0..8 hold rows t (lanes 0-15) and 9 + t (lanes 16-31), columns 0..15 -- a lane
    // walks DOWN nine consecutive rows, so the bottom taps of one slot are the top taps of the next and a reload of the
    // J taps costs two byte loads per slot instead of four; slot 9 holds column 16 of row `lane`.  With t a compile-time constant every address is a per-lane base plus an
    // immediate, so the unrolled loops carry no index arithmetic.  (The generic setup below keeps running counters.)
    static_assert(WIN == 17, "the slot mapping of klt_track_packed is laid out for a 17 x 17 window");
    const int hi = lane >> 4, lx = lane & 15;
    auto slot_y = [&](int t) { return t < T - 1 ? t + 9 * hi : lane; };
    auto slot_x = [&](int t) { return t < T - 1 ? lx : 16; };
    unsigned int jw[T];
#pragma unroll
    for (int t = 0; t < T; t++) jw[t] = 0u;
    for (int i = W2 + lane; i < T * 32; i += 32) GI[i] = make_int2(0, 0);    // padding slots stay zero
    bool st = true;
    float e_out = 0.f;
    float outx = 0.f, outy = 0.f;

    for (int level = top; level >= 0; level--) {
        const int cols = lay.w[level], rows = lay.h[level];
        const size_t pitch = lay.pitch[level];
        const uint8_t* I = Ip + lay.offset[level];
        const uint8_t* J = Jp + lay.offset[level];
        const float sc = (float)(1. / (1 << level));
        float prx = px0 * sc, pry = py0 * sc;
        float nx, ny;
        if (level == top) { nx = prx; ny = pry; }
        else { nx = outx * 2.f; ny = outy * 2.f; }
        outx = nx; outy = ny;
        prx -= half; pry -= half;
        const int ipx = (int)floorf(prx), ipy = (int)floorf(pry);
        if (ipx < -WIN || ipx >= cols || ipy < -WIN || ipy >= rows) {
            if (level == 0) { st = false; e_out = 0.f; }
            continue;
        }
        float a = prx - ipx, b = pry - ipy;
        int iw00 = __float2int_rn((1.f - a) * (1.f - b) * (1 << W_BITS));
        int iw01 = __float2int_rn(a * (1.f - b) * (1 << W_BITS));
        int iw10 = __float2int_rn((1.f - a) * b * (1 << W_BITS));
        int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;

        __syncwarp();
        stage_patch20(I, rows, cols, pitch, ipx - 1, ipy - 1, patch, lane);
        __syncwarp();
        int sA11 = 0, sA12 = 0, sA22 = 0;
        if (ipx >= 0 && ipy >= 0 && ipx + DN <= cols && ipy + DN <= rows) {   // warp-uniform
            // Fast setup.  Every derivative sample of the window lies inside the image, so no sample is zeroed and
            // the integer identity  sum_k w_k Scharr(I)(p + k) = Scharr(sum_k w_k I(. + k))(p)  holds exactly: first
            // the un-descaled bilinear patch A (AN x AN, 32-bit), then per window pixel one 3x3 neighbourhood of A
            // gives I (centre), Ix and Iy.  Same integers as the reference order, a third of the instructions.
            {   // AN x AN = 19 x 19 positions: rows 0-9 (lanes 0-15) / 10-18 (lanes 16-31) x 16 columns, then columns 16..18.
                // A lane walks down its column: the two pixels of a patch row feed the A row above (weights iw10, iw11)
                // and the A row below (iw00, iw01) -- two byte loads per position instead of four, same integers.
                const uint8_t* p_lane = patch + 10 * hi * PN + lx;
                int* o_lane = At + 10 * hi * AN + lx;
                int top = (int)p_lane[0] * iw00 + (int)p_lane[1] * iw01;
#pragma unroll
                for (int k = 0; k < 10; k++) {
                    const uint8_t* s1 = p_lane + (k + 1) * PN;               // lanes 16-31, k = 9: one row past the patch (this warp's
                    const int q0 = (int)s1[0], q1 = (int)s1[1];               // own A area follows it); that A row is not stored
                    const int a = top + q0 * iw10 + q1 * iw11;
                    if (10 * hi + k < AN) o_lane[k * AN] = a;
                    top = q0 * iw00 + q1 * iw01;
                }
                // remaining 3 x 19 = 57 positions, column-major: position q = lane + 32 u -> (row q % 19, column 16 + q / 19)
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int q = lane + 32 * u, qc = q / AN, qr = q - qc * AN;
                    if (q < 3 * AN) {
                        const uint8_t* s0 = patch + qr * PN + 16 + qc;
                        At[qr * AN + 16 + qc] = (int)s0[0] * iw00 + (int)s0[1] * iw01 + (int)s0[PN] * iw10 + (int)s0[PN + 1] * iw11;
                    }
                }
            }
            __syncwarp();
            // A lane walks down its column: per row of A the horizontal difference d = right - left (Scharr x) and the
            // smoothed value sm = 3 left + 10 centre + 3 right (Scharr y) are formed once and slide through three
            // registers each; a slot then costs one new row (3 loads) instead of a 3x3 neighbourhood (9 loads).
            // Integer arithmetic regrouped, same integers:  ix = 3 (d0 + d2) + 10 d1,  iy = sm2 - sm0.
            auto emit = [&](int t, bool valid, int c1, int d0, int d1, int d2, int sm0, int sm2) {
                const int iv = valid ? descale(c1, W_BITS - 5) : 0;
                const int ix = valid ? descale((d0 + d2) * 3 + d1 * 10, W_BITS) : 0;
                const int iy = valid ? descale(sm2 - sm0, W_BITS) : 0;
                GI[t * 32 + lane] = make_int2((ix & 0xFFFF) | (iy << 16), iv);
                sA11 += ix * ix; sA12 += ix * iy; sA22 += iy * iy;
            };
            {
                const int* a_lane = At + 9 * hi * AN + lx;
                int d0, d1, d2, sm0, sm1, sm2, c1, c2;
                { const int l = a_lane[0], c = a_lane[1], r = a_lane[2]; d0 = r - l; sm0 = (l + r) * 3 + c * 10; }
                { const int l = a_lane[AN], c = a_lane[AN + 1], r = a_lane[AN + 2]; d1 = r - l; sm1 = (l + r) * 3 + c * 10; c1 = c; }
#pragma unroll
                for (int t = 0; t < T - 1; t++) {
                    const int* rw = a_lane + (t + 2) * AN;
                    { const int l = rw[0], c = rw[1], r = rw[2]; d2 = r - l; sm2 = (l + r) * 3 + c * 10; c2 = c; }
                    emit(t, slot_y(t) < WIN, c1, d0, d1, d2, sm0, sm2);
                    d0 = d1; d1 = d2; sm0 = sm1; sm1 = sm2; c1 = c2;
                }
            }
            {   // slot T - 1: column 16 of row `lane`
                const int* r0 = At + lane * AN + 16;
                const int* r1 = r0 + AN;
                const int* r2 = r1 + AN;
                const int a00 = r0[0], a01 = r0[1], a02 = r0[2], a10 = r1[0], a11 = r1[1], a12 = r1[2], a20 = r2[0], a21 = r2[1], a22 = r2[2];
                emit(T - 1, slot_y(T - 1) < WIN, a11, a02 - a00, a12 - a10, a22 - a20, (a00 + a02) * 3 + a01 * 10, (a20 + a22) * 3 + a21 * 10);
            }
        } else {
        {
            int qy = lane / DN, qx = lane - qy * DN;
#pragma unroll 2
            for (int i = lane; i < DN * DN; i += 32) {
                const int gx = ipx + qx, gy = ipy + qy;
                int dx = 0, dy = 0;
                if (gx >= 0 && gx < cols && gy >= 0 && gy < rows) {
                    const uint8_t* r0 = patch + qy * PN + qx;
                    const uint8_t* r1 = r0 + PN;
                    const uint8_t* r2 = r1 + PN;
                    const int t0m = ((int)r0[0] + (int)r2[0]) * 3 + (int)r1[0] * 10;
                    const int t0p = ((int)r0[2] + (int)r2[2]) * 3 + (int)r1[2] * 10;
                    const int t1m = (int)r2[0] - (int)r0[0], t1c = (int)r2[1] - (int)r0[1], t1p = (int)r2[2] - (int)r0[2];
                    dx = t0p - t0m;
                    dy = (t1p + t1m) * 3 + t1c * 10;
                }
                dpatch[2 * i] = (short)dx;
                dpatch[2 * i + 1] = (short)dy;
                qx += 32 % DN; qy += 32 / DN;
                if (qx >= DN) { qx -= DN; qy++; }
            }
        }
        __syncwarp();
        {
#pragma unroll 1
            for (int t = 0; t < T; t++) {
                const int y = slot_y(t), x = slot_x(t), i = t * 32 + lane;
                if (y >= WIN) { GI[i] = make_int2(0, 0); continue; }
                const uint8_t* s0 = patch + (y + 1) * PN + (x + 1);
                const int iv = descale((int)s0[0] * iw00 + (int)s0[1] * iw01 + (int)s0[PN] * iw10 + (int)s0[PN + 1] * iw11, W_BITS - 5);
                const short* d0 = dpatch + 2 * (y * DN + x);
                const short* d1 = d0 + 2 * DN;
                const int ix = descale((int)d0[0] * iw00 + (int)d0[2] * iw01 + (int)d1[0] * iw10 + (int)d1[2] * iw11, W_BITS);
                const int iy = descale((int)d0[1] * iw00 + (int)d0[3] * iw01 + (int)d1[1] * iw10 + (int)d1[3] * iw11, W_BITS);
                GI[i] = make_int2((ix & 0xFFFF) | (iy << 16), iv);
                sA11 += ix * ix; sA12 += ix * iy; sA22 += iy * iy;
            }
        }
        }
        const long long iA11 = warp_sum_exact(sA11), iA12 = warp_sum_exact(sA12), iA22 = warp_sum_exact(sA22);
        const float A11 = (float)iA11 * FLT_SCALE, A12 = (float)iA12 * FLT_SCALE, A22 = (float)iA22 * FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * WIN * WIN);
        if ((double)minEig < min_eig || D < 1.1920928955078125e-07f) {
            if (level == 0) st = false;
            continue;
        }
        D = 1.f / D;
        nx -= half; ny -= half;
        float pdx = 0.f, pdy = 0.f;
        int sx = 0x7fffffff, sy = 0x7fffffff;                 // integer position of the staged J patch
        // j == max_iters is the extra pass that measures the final patch error (level 0 only)
        bool want_err = false;
        for (int j = 0; j <= max_iters; j++) {
            float fx = nx, fy = ny;
            if (j == max_iters || want_err) {
                if (!(st && level == 0)) break;
                want_err = true;
                fx = outx - half; fy = outy - half;
            }
            const int inx = (int)floorf(fx), iny = (int)floorf(fy);
            if (inx < -WIN || inx >= cols || iny < -WIN || iny >= rows) {
                if (level == 0) st = false;
                break;
            }
            a = fx - inx; b = fy - iny;
            iw00 = __float2int_rn((1.f - a) * (1.f - b) * (1 << W_BITS));
            iw01 = __float2int_rn(a * (1.f - b) * (1 << W_BITS));
            iw10 = __float2int_rn((1.f - a) * b * (1 << W_BITS));
            iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            if (inx != sx || iny != sy) {                     // pack the four taps of every window pixel
                const bool inside = inx >= 0 && iny >= 0 && inx + DN <= cols && iny + DN <= rows;   // warp-uniform
                if (inside) {
                    const uint8_t* base = J + (size_t)iny * pitch + inx;
                    const int ip = (int)pitch;
                    // rows 9 hi .. 9 hi + 9 of this lane's column pair (the lanes of the lower half stop one row early:
                    // row 18 is outside the staged region and belongs to no window pixel)
                    const uint8_t* col = base + 9 * hi * ip + lx;
                    unsigned int pk[T];
#pragma unroll
                    for (int k = 0; k < T; k++) {
                        pk[k] = (k < T - 1 || hi == 0) ? ((unsigned)col[0] | ((unsigned)col[1] << 8)) : 0u;
                        col += ip;
                    }
#pragma unroll
                    for (int t = 0; t < T - 1; t++) jw[t] = slot_y(t) < WIN ? (pk[t] | (pk[t + 1] << 16)) : 0u;
                    {
                        const uint8_t* s0 = base + lane * ip + 16;
                        jw[T - 1] = slot_y(T - 1) < WIN ? ((unsigned)s0[0] | ((unsigned)s0[1] << 8) | ((unsigned)s0[ip] << 16) | ((unsigned)s0[ip + 1] << 24)) : 0u;
                    }
                } else {
                    __syncwarp();
                    stage_patch_fast<DN>(J, rows, cols, pitch, inx, iny, patch, lane);
                    __syncwarp();
                    const uint8_t* p_lane = patch + 9 * hi * DN + lx;
#pragma unroll
                    for (int t = 0; t < T; t++) {
                        const uint8_t* s0 = t < T - 1 ? p_lane + t * DN : patch + lane * DN + 16;
                        jw[t] = slot_y(t) < WIN ? ((unsigned)s0[0] | ((unsigned)s0[1] << 8) | ((unsigned)s0[DN] << 16) | ((unsigned)s0[DN + 1] << 24)) : 0u;
                    }
                }
                sx = inx; sy = iny;
            }
            // the weights are signed 16-bit (iw11 = 2^14 - the other three can be -1 after rounding), the pixels unsigned bytes
            const unsigned int w_lo = ((unsigned)iw00 & 0xFFFFu) | ((unsigned)iw01 << 16), w_hi = ((unsigned)iw10 & 0xFFFFu) | ((unsigned)iw11 << 16);
            int sb1 = 0, sb2 = 0, se = 0;
            if (!want_err) {
#pragma unroll
                for (int t = 0; t < T; t++) {                // slots past the window hold zeros: they add nothing
                    const int acc = dp2a_hi_su(w_hi, jw[t], dp2a_lo_su(w_lo, jw[t], 1 << (W_BITS - 6)));
                    const int2 gi = GI[t * 32 + lane];
                    const int diff = (acc >> (W_BITS - 5)) - gi.y;
                    const int g = gi.x;
                    sb1 += diff * (int)(short)(g & 0xFFFF);
                    sb2 += diff * (g >> 16);
                }
            } else {
#pragma unroll
                for (int t = 0; t < T; t++) {
                    const int acc = dp2a_hi_su(w_hi, jw[t], dp2a_lo_su(w_lo, jw[t], 1 << (W_BITS - 6)));
                    const int diff = (acc >> (W_BITS - 5)) - GI[t * 32 + lane].y;
                    se += diff < 0 ? -diff : diff;
                }
            }
            if (want_err) {
                const long long e = warp_sum_exact(se);
                e_out = (float)e * 1.f / (float)(32 * WIN * WIN);
                break;
            }
            const long long ib1 = warp_sum_exact(sb1), ib2 = warp_sum_exact(sb2);
            const float b1 = (float)ib1 * FLT_SCALE, b2 = (float)ib2 * FLT_SCALE;
            const float dx = (float)((A12 * b2 - A22 * b1) * D);
            const float dy = (float)((A12 * b1 - A11 * b2) * D);
            nx += dx; ny += dy;
            outx = nx + half; outy = ny + half;
            if ((double)dx * dx + (double)dy * dy <= eps2) { want_err = true; continue; }
            if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
                outx -= dx * 0.5f;
                outy -= dy * 0.5f;
                want_err = true;
                continue;
            }
            pdx = dx; pdy = dy;
        }
    }
    if (lane == 0) {
        next_pts[2 * pidx] = outx;
        next_pts[2 * pidx + 1] = outy;
        status[pidx] = st ? 1 : 0;
        err[pidx] = e_out;
    }
}

}  // namespace

static int klt_layout(int H, int W, int max_level, int win, PyrLayout* L) {
    if (max_level < 0 || max_level >= KLT_MAX_LEVELS) return 1;
    int h = H, w = W, level = 0;
    size_t off = 0;
    for (;;) {
        L->h[level] = h; L->w[level] = w;
        L->pitch[level] = ((size_t)w + 15) & ~(size_t)15;
        L->offset[level] = off;
        off += (L->pitch[level] * h + 255) & ~(size_t)255;
        if (level == max_level) break;
        const int nh = (h + 1) / 2, nw = (w + 1) / 2;
        if (nw <= win || nh <= win) break;
        h = nh; w = nw; level++;
    }
    L->n_levels = level + 1;
    L->frame_bytes = off;
    return 0;
}

int vo_klt_layout_host(int H, int W, int max_level, int win, int* n_levels, int* level_h, int* level_w,
                       size_t* level_pitch, size_t* level_offset, size_t* frame_bytes) {
    PyrLayout L;
    VO_REQUIRE(H >= 1 && W >= 1 && win >= 3 && win <= 31, "klt: bad image size / window (3..31)");
    VO_REQUIRE(klt_layout(H, W, max_level, win, &L) == 0, "klt: max_level must be in [0, %d)", KLT_MAX_LEVELS);
    *n_levels = L.n_levels;
    for (int l = 0; l < L.n_levels; l++) {
        if (level_h) level_h[l] = L.h[l];
        if (level_w) level_w[l] = L.w[l];
        if (level_pitch) level_pitch[l] = L.pitch[l];
        if (level_offset) level_offset[l] = L.offset[l];
    }
    *frame_bytes = L.frame_bytes;
    return VO_OK;
}

int vo_launch_klt_pyramid(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                          size_t frame_stride, int max_level, int win, uint8_t* d_pyr, cudaStream_t stream, cudaEvent_t level0_done) {
    PyrLayout L;
    VO_REQUIRE(H >= 1 && W >= 1 && win >= 3 && win <= 31 && n_frames >= 1, "klt pyramid: bad arguments");
    VO_REQUIRE(klt_layout(H, W, max_level, win, &L) == 0, "klt: max_level must be in [0, %d)", KLT_MAX_LEVELS);
    // level 0: copy unless the caller already placed the frames in the pyramid's level-0 slots
    if (!(d_img == d_pyr && pitch == L.pitch[0] && frame_stride == L.frame_bytes)) {
        const int chunks = vo_div_up(W, 16);
        const int bx = chunks >= 256 ? 256 : ((chunks + 31) & ~31), by = 256 / bx > 0 ? 256 / bx : 1;
        dim3 g(vo_div_up(chunks, bx), vo_div_up(H, by), n_frames), b(bx, by);
        copy_level0_kernel<<<g, b, 0, stream>>>(d_img, H, W, pitch, frame_stride, d_pyr, L.pitch[0], L.frame_bytes);
        ctx->launches++;
        VO_CHECK_LAUNCH();
    }
    if (level0_done) VO_CUDA(cudaEventRecord(level0_done, stream));   // level 0 is all the Harris detector reads
    for (int l = 1; l < L.n_levels; l++) {
        dim3 g(vo_div_up(vo_div_up(L.w[l], PD_OUT_W), PD_WARPS), vo_div_up(L.h[l], PD_ROWS), n_frames);
        pyr_down_kernel<<<g, PD_WARPS * 32, 0, stream>>>(d_pyr + L.offset[l - 1], L.h[l - 1], L.w[l - 1], L.pitch[l - 1],
                                               L.frame_bytes, d_pyr + L.offset[l], L.h[l], L.w[l], L.pitch[l],
                                               L.frame_bytes);
        ctx->launches++;
        VO_CHECK_LAUNCH();
    }
    return VO_OK;
}

int vo_launch_bgr2gray(vo_ctx* ctx, const uint8_t* d_bgr, int n_frames, int H, int W, size_t in_pitch, size_t in_frame_stride,
                       uint8_t* d_gray, size_t out_pitch, size_t out_frame_stride, cudaStream_t stream) {
    VO_REQUIRE(n_frames >= 1 && H >= 1 && W >= 1 && in_pitch >= (size_t)3 * W && out_pitch >= (size_t)W, "bgr2gray: bad shape / pitch");
    VO_REQUIRE(out_pitch % 4 == 0 && ((uintptr_t)d_gray % 4) == 0 && out_frame_stride % 4 == 0, "bgr2gray: destination rows must be 4-byte aligned");
    dim3 g(vo_div_up(vo_div_up(W, 4), 256), H, n_frames);
    bgr2gray_kernel<<<g, 256, 0, stream>>>(d_bgr, H, W, in_pitch, in_frame_stride, d_gray, out_pitch, out_frame_stride);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    return VO_OK;
}

int vo_launch_klt_track(vo_ctx* ctx, const uint8_t* d_pyr_prev, const uint8_t* d_pyr_next, int n_frames, int H, int W,
                        int max_level, int win, int max_iters, double epsilon, double min_eig,
                        const float* d_prev_pts, int n_pts, float* d_next_pts, uint8_t* d_status, float* d_err,
                        cudaStream_t stream, const int* d_n_valid) {
    PyrLayout L;
    VO_REQUIRE(H >= 1 && W >= 1 && win >= 3 && win <= 31 && n_frames >= 1 && n_pts >= 0, "klt track: bad arguments");
    VO_REQUIRE(klt_layout(H, W, max_level, win, &L) == 0, "klt: max_level must be in [0, %d)", KLT_MAX_LEVELS);
    if (n_pts == 0) return VO_OK;
    // criteria clamps of cv2.calcOpticalFlowPyrLK
    if (max_iters < 0) max_iters = 0;
    if (max_iters > 100) max_iters = 100;
    if (epsilon < 0.) epsilon = 0.;
    if (epsilon > 10.) epsilon = 10.;
    const int pn = win + 3, dn = win + 1, w2 = win * win;
    size_t per_warp = (size_t)((pn * pn + 15) & ~15) + (size_t)dn * dn * 4 + (size_t)w2 * 2 + (size_t)w2 * 4;
    per_warp = (per_warp + 15) & ~(size_t)15;
    const size_t smem = per_warp * KLT_WARPS;
    if (vo_ctx_once(ctx, VO_ATTR_KLT))
        VO_CUDA(cudaFuncSetAttribute(klt_track_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    dim3 g(vo_div_up(n_pts, KLT_WARPS), n_frames);
    // The specialised kernels stage patches with a single reflection (reflect1), which is only valid while every
    // level is wider than the patch: levels >= 1 are by construction (klt_layout), level 0 is checked here.
    const bool roomy = H > win + 3 && W > win + 3;
    if (win == 17 && roomy && !ctx->env_klt_generic) {
        dim3 gs(vo_div_up(n_pts, KLT_SWARPS), n_frames);
        klt_track_packed<17><<<gs, KLT_SWARPS * 32, 0, stream>>>(d_pyr_prev, d_pyr_next, L, max_iters, epsilon * epsilon, min_eig,
                                                                  d_prev_pts, n_pts, d_next_pts, d_status, d_err, d_n_valid);
        ctx->launches++;
        VO_CHECK_LAUNCH();
        return VO_OK;
    }
    if (win == 21 && roomy && !ctx->env_klt_generic) {
        klt_track_fast<21><<<g, KLT_WARPS * 32, 0, stream>>>(d_pyr_prev, d_pyr_next, L, max_iters, epsilon * epsilon, min_eig,
                                                              d_prev_pts, n_pts, d_next_pts, d_status, d_err, d_n_valid);
        ctx->launches++;
        VO_CHECK_LAUNCH();
        return VO_OK;
    }
    klt_track_kernel<<<g, KLT_WARPS * 32, smem, stream>>>(d_pyr_prev, d_pyr_next, L, win, max_iters, epsilon * epsilon,
                                                          min_eig, d_prev_pts, n_pts, d_next_pts, d_status, d_err, d_n_valid);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    return VO_OK;
}
