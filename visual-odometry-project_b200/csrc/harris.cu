// Harris corner detector for sm_100a: fused Sobel -> structure tensor -> box sums -> response,
// followed by an exact parallel evaluation of the reference's greedy non-maximum suppression.
//
// Replaces  /root/reference/src/vo/features/harris.py:86-158 (HarrisCornerDetector.extractKeypoints)
//           /root/reference/src/vo/features/harris.py:160-194 (extractDescriptors)
//
// Exactness: gradients, products and box sums are exact int32; the response is evaluated in
// float64 with individually rounded operations (__dmul_rn/__dadd_rn/__dsub_rn, no FMA
// contraction) in the same order as the numpy expression harris.py:123-126, so the response map
// is bit-identical to the reference's float64 `harris_scores`.
//
// NMS: the reference repeats {argmax; zero a (2r+1)^2 box} K times (harris.py:148-152).  That
// sequence equals the first K elements, in priority order (score desc, linear index asc), of the
// greedy maximal independent set of the "within Chebyshev distance r" graph.  It is computed here
// by (1) a dense local-maximum pass (local maxima are always picks), (2) a per-frame threshold = K-th
// best local maximum (nothing below it can be among the first K picks) and a bitmap of the boxes of the
// maxima above it, (3) one scan that lists the pixels above the threshold outside those boxes, (4) a
// per-frame kernel that walks that list in priority bands -- inside a band, rounds of "no undecided
// neighbour of higher priority -> pick, suppress the box" -- until K picks are known, then sorts them.
#include <cuda.h>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// Response kernel v1: 64x32 output tile per CTA, three shared-memory stages.
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int RT_W = 64;   // output tile width
constexpr int RT_H = 32;   // output tile height
constexpr int R_THREADS = 256;
constexpr int PR_MAX = 7;  // patch_size <= 15

__device__ __forceinline__ double harris_score(int a, int b, int c, double kappa) {
    // harris.py:123-127, one rounding per numpy ufunc call.
    const double sa = (double)a, sb = (double)b, sc = (double)c;
    const double trace = __dadd_rn(sa, sb);
    const double det = __dsub_rn(__dmul_rn(sa, sb), __dmul_rn(sc, sc));
    const double s = __dsub_rn(det, __dmul_rn(kappa, __dmul_rn(trace, trace)));
    return s < 0.0 ? 0.0 : s;
}

__global__ void __launch_bounds__(R_THREADS)
harris_response_tiled(const uint8_t* __restrict__ img, size_t pitch, size_t frame_stride,
                      int H, int W, int pr, double kappa, double* __restrict__ resp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int pad = pr + 1;
    const int in_w = RT_W + 2 * pad, in_h = RT_H + 2 * pad;
    const int pw = RT_W + 2 * pr, ph = RT_H + 2 * pr;  // product region
    int* pxx = reinterpret_cast<int*>(smem_raw);
    int* pyy = pxx + ph * pw;
    int* pxy = pyy + ph * pw;
    int* hxx = pxy + ph * pw;                          // [ph][RT_W]
    int* hyy = hxx + ph * RT_W;
    int* hxy = hyy + ph * RT_W;
    uint8_t* tile = reinterpret_cast<uint8_t*>(hxy + ph * RT_W);  // [in_h][in_w]

    const int x0 = blockIdx.x * RT_W, y0 = blockIdx.y * RT_H;
    const uint8_t* src = img + (size_t)blockIdx.z * frame_stride;
    double* dst = resp + (size_t)blockIdx.z * H * W;
    const int tid = threadIdx.x;

    for (int i = tid; i < in_h * in_w; i += R_THREADS) {
        const int ty = i / in_w, tx = i - ty * in_w;
        const int gy = y0 - pad + ty, gx = x0 - pad + tx;
        uint8_t v = 0;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = src[(size_t)gy * pitch + gx];
        tile[i] = v;
    }
    __syncthreads();
    // stage 1: Sobel + products at product-region position (py, px) <-> tile (py+1, px+1)
    for (int i = tid; i < ph * pw; i += R_THREADS) {
        const int py = i / pw, px = i - py * pw;
        const uint8_t* t0 = tile + py * in_w + px;
        const uint8_t* t1 = t0 + in_w;
        const uint8_t* t2 = t1 + in_w;
        const int gx = ((int)t0[0] + 2 * (int)t1[0] + (int)t2[0]) - ((int)t0[2] + 2 * (int)t1[2] + (int)t2[2]);
        const int gy = ((int)t0[0] + 2 * (int)t0[1] + (int)t0[2]) - ((int)t2[0] + 2 * (int)t2[1] + (int)t2[2]);
        pxx[i] = gx * gx;
        pyy[i] = gy * gy;
        pxy[i] = gx * gy;
    }
    __syncthreads();
    // stage 2: horizontal sliding sums, one (row, 8-column segment) per item
    const int ps = 2 * pr + 1;
    for (int i = tid; i < ph * (RT_W / 8); i += R_THREADS) {
        const int row = i / (RT_W / 8), seg = i - row * (RT_W / 8);
        const int* rxx = pxx + row * pw + seg * 8;
        const int* ryy = pyy + row * pw + seg * 8;
        const int* rxy = pxy + row * pw + seg * 8;
        int sxx = 0, syy = 0, sxy = 0;
        for (int k = 0; k < ps; k++) { sxx += rxx[k]; syy += ryy[k]; sxy += rxy[k]; }
        const int o = row * RT_W + seg * 8;
        hxx[o] = sxx; hyy[o] = syy; hxy[o] = sxy;
#pragma unroll
        for (int k = 1; k < 8; k++) {
            sxx += rxx[k + ps - 1] - rxx[k - 1];
            syy += ryy[k + ps - 1] - ryy[k - 1];
            sxy += rxy[k + ps - 1] - rxy[k - 1];
            hxx[o + k] = sxx; hyy[o + k] = syy; hxy[o + k] = sxy;
        }
    }
    __syncthreads();
    // stage 3: vertical sliding sums + response; thread = (column, 8-row segment)
    {
        const int col = tid % RT_W, rseg = tid / RT_W;  // 64 x 4
        int sxx = 0, syy = 0, sxy = 0;
        const int r0 = rseg * 8;
        for (int k = 0; k < ps; k++) {
            sxx += hxx[(r0 + k) * RT_W + col];
            syy += hyy[(r0 + k) * RT_W + col];
            sxy += hxy[(r0 + k) * RT_W + col];
        }
        const int gx = x0 + col;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int gy = y0 + r0 + k;
            if (k > 0) {
                sxx += hxx[(r0 + k + ps - 1) * RT_W + col] - hxx[(r0 + k - 1) * RT_W + col];
                syy += hyy[(r0 + k + ps - 1) * RT_W + col] - hyy[(r0 + k - 1) * RT_W + col];
                sxy += hxy[(r0 + k + ps - 1) * RT_W + col] - hxy[(r0 + k - 1) * RT_W + col];
            }
            if (gx < W && gy < H) {
                double s = 0.0;
                if (gx >= pad && gx < W - pad && gy >= pad && gy < H - pad) s = harris_score(sxx, syy, sxy, kappa);
                dst[(size_t)gy * W + gx] = s;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Response kernel v2 (patch_size 9, TMA-legal image layout): 128x54 output tile per CTA.
//   load   : one thread issues a 3-D TMA tile load (144 x 64 x 1 bytes, zero fill outside the image)
//   phase H: thread = (product row, 32-column segment).  The 3 image rows it needs are pulled with
//            conflict-free 128-bit shared loads into registers; it marches 42 columns, fully
//            unrolled: Sobel -> 3 products -> sliding 9-sums with a register ring -> H[row][x].
//   phase V: thread = (column, half tile).  Sliding vertical 9-sums over H with a register ring,
//            float64 score, coalesced 8-byte stores (a warp writes 256 contiguous bytes).
// Shared memory: 9.2 KB image + 3 x 62 x 129 x 4 B H-sums (row pitch 129 words: the row-per-lane
// stores of phase H and the column-per-lane loads of phase V are both bank-conflict free).
// ---------------------------------------------------------------------------------------------
constexpr int FT_W = 128, FT_H = 54, FT_ROWS = FT_H + 8;   // 62 product rows
constexpr int FT_IN_W = 144, FT_IN_H = 64;                 // TMA box
constexpr int FT_HP = 129;                                 // H row pitch (words)
constexpr int FT_THREADS = 256;
constexpr int FT_XSHIFT = 11;                              // (x0 - 5) % 16 == 0
constexpr int FT_IMG_BYTES = FT_IN_W * FT_IN_H;            // 9216
constexpr int FT_SMEM = FT_IMG_BYTES + 3 * FT_ROWS * FT_HP * 4 + 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// WC > 0: the image width is the compile-time constant WC (the row offsets of the stores become immediates:
// one instruction less per pixel); WC = 0: any width.
template <int WC>
__global__ void __launch_bounds__(FT_THREADS, 2)
harris_response_fast(const __grid_constant__ CUtensorMap tmap, int H, int W_rt, double kappa, double* __restrict__ resp,
                     int tiles_x, int tiles_y, int n_tiles) {
    const int W = WC > 0 ? WC : W_rt;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint8_t* img = smem_raw;
    int2* hab = reinterpret_cast<int2*>(smem_raw + FT_IMG_BYTES);          // (sum Ix^2, sum Iy^2): one 64-bit access
    int* hxy = reinterpret_cast<int*>(hab + FT_ROWS * FT_HP);              // sum Ix*Iy
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + FT_IMG_BYTES + 3 * FT_ROWS * FT_HP * 4);
    const int tid = threadIdx.x;

    // Persistent CTA (grid = 2 x SM count): tile t -> (frame, tile row, tile column).  The image box of the
    // next tile is requested as soon as phase H has consumed the current one, so its HBM latency hides
    // behind phase V.  Tiles start 11 columns left of a multiple of 128 so that the TMA box (which begins
    // 5 columns further left) starts on a 16-byte boundary, as cp.async.bulk.tensor needs for 1-byte data.
    // tile t -> (column bx, row by, frame fr), advanced by the grid stride without divisions (the decomposition of
    // the stride is computed once; a division per tile and thread was 8 % of the kernel's instructions)
    const int st_x = (int)gridDim.x % tiles_x, st_q = (int)gridDim.x / tiles_x, st_y = st_q % tiles_y, st_f = st_q / tiles_y;
    auto advance = [&](int& bx, int& by, int& fr) {
        bx += st_x;
        const int cx = bx >= tiles_x ? 1 : 0;
        bx -= cx ? tiles_x : 0;
        by += st_y + cx;
        const int cy = by >= tiles_y ? 1 : 0;
        by -= cy ? tiles_y : 0;
        fr += st_f + cy;
    };
    auto issue_load = [&](int bx, int by, int fr) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(FT_IMG_BYTES) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(smem_u32(img)), "l"(&tmap), "r"(bx * FT_W - FT_XSHIFT - 5), "r"(by * FT_H - 5), "r"(fr), "r"(smem_u32(mbar)) : "memory");
    };
    if (tid < FT_ROWS) { hab[tid * FT_HP + FT_W] = make_int2(0, 0); hxy[tid * FT_HP + FT_W] = 0; }   // the zero column (phase V)
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int tbx = (int)blockIdx.x % tiles_x, tby = ((int)blockIdx.x / tiles_x) % tiles_y, tfr = (int)blockIdx.x / (tiles_x * tiles_y);
    if (tid == 0 && (int)blockIdx.x < n_tiles) issue_load(tbx, tby, tfr);
    __syncthreads();
    uint32_t parity = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, parity ^= 1) {
    const int x0 = tbx * FT_W - FT_XSHIFT, y0 = tby * FT_H, f = tfr;
    advance(tbx, tby, tfr);                                  // now the coordinates of this CTA's next tile
    {   // all threads wait for the tile
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
        }
    }
    // ---------------- phase H ----------------
    {
        const int g = tid >> 6, r = tid & 63;
        if (r < FT_ROWS) {
            uint32_t w[3][12];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const uint4* src = reinterpret_cast<const uint4*>(img + (r + k) * FT_IN_W + 32 * g);
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const uint4 v = src[q];
                    w[k][4 * q] = v.x; w[k][4 * q + 1] = v.y; w[k][4 * q + 2] = v.z; w[k][4 * q + 3] = v.w;
                }
            }
            int ring_xx[9], ring_yy[9], ring_xy[9];
            int sxx = 0, syy = 0, sxy = 0;
            int2* oab = hab + r * FT_HP + 32 * g;
            int* oxy = hxy + r * FT_HP + 32 * g;
            // Sobel by 8-bit dot products on the packed pixels (no byte extraction): for product column m
            //   gx = sum_k {1,2,1}[k] * (p_k[m] - p_k[m+2]),   gy = sum_j {1,2,1}[j] * (p_0[m+j] - p_2[m+j])
            // (left - right, top - bottom: the signs cancel in the products, harris.py:103-113).  Pixels m .. m+2 lie in
            // the aligned word of image row k for m % 4 = 0, 1 and in that word funnel-shifted by two bytes for
            // m % 4 = 2, 3; the byte offset (0 or 1) inside the word is folded into the dp4a weights, so four
            // columns share one shift per image row.
            uint32_t s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
            for (int m = 0; m < 40; m++) {
                const int q = m >> 2;
                if ((m & 3) == 0) { s0 = w[0][q]; s1 = w[1][q]; s2 = w[2][q]; }
                if ((m & 3) == 2) {
                    s0 = __funnelshift_r(w[0][q], w[0][q + 1], 16);
                    s1 = __funnelshift_r(w[1][q], w[1][q + 1], 16);
                    s2 = __funnelshift_r(w[2][q], w[2][q + 1], 16);
                }
                int gx, gy;
                if ((m & 1) == 0) {
                    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(gx) : "r"(s0), "r"(0x00FF0001), "r"(0));
                    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(gx) : "r"(s1), "r"(0x00FE0002), "r"(gx));
                    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(gx) : "r"(s2), "r"(0x00FF0001), "r"(gx));
                    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(gy) : "r"(s0), "r"(0x00010201), "r"(0));
                    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(gy) : "r"(s2), "r"(0x00FFFEFF), "r"(gy));
                } else {
                    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(gx) : "r"(s0), "r"(0xFF000100), "r"(0));
                    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(gx) : "r"(s1), "r"(0xFE000200), "r"(gx));
                    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(gx) : "r"(s2), "r"(0xFF000100), "r"(gx));
                    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(gy) : "r"(s0), "r"(0x01020100), "r"(0));
                    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(gy) : "r"(s2), "r"(0xFFFEFF00), "r"(gy));
                }
                const int pxx = gx * gx, pyy = gy * gy, pxy = gx * gy;
                if (m >= 9) {
                    sxx += pxx - ring_xx[m % 9]; syy += pyy - ring_yy[m % 9]; sxy += pxy - ring_xy[m % 9];
                } else {
                    sxx += pxx; syy += pyy; sxy += pxy;
                }
                ring_xx[m % 9] = pxx; ring_yy[m % 9] = pyy; ring_xy[m % 9] = pxy;
                if (m >= 8) { oab[m - 8] = make_int2(sxx, syy); oxy[m - 8] = sxy; }
            }
        }
    }
    __syncthreads();
    if (tid == 0 && tile + (int)gridDim.x < n_tiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of img precede the async overwrite
        issue_load(tbx, tby, tfr);
    }
    // ---------------- phase V ----------------
    {
        const int col = tid & 127, half = tid >> 7;
        const int base = 27 * half;
        const int gx = x0 + col;
        // Columns inside the zero border of the score map (harris.py:129-137) read the spare column 128 of the
        // H arrays, which holds zeros: their sums are 0 and the score evaluates to +0 without a predicate.
        const bool x_in = gx >= 0 && gx < W, x_interior = gx >= 5 && gx < W - 5;
        const int hcol = x_interior ? col : FT_W;
        int rxx[9], ryy[9], rxy[9];
        int vxx = 0, vyy = 0, vxy = 0;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const int2 ab = hab[(base + k) * FT_HP + hcol];
            rxx[k] = ab.x; ryy[k] = ab.y; rxy[k] = hxy[(base + k) * FT_HP + hcol];
            vxx += rxx[k]; vyy += ryy[k]; vxy += rxy[k];
        }
        // Branch-free emission (straight-line blocks, so the scheduler can overlap the float64 chains of
        // neighbouring rows).  The clamp s < 0 -> 0 (harris.py:127) costs no compare: with h = s / 2,
        // |h| + h is s for s > 0 and +0 otherwise, exactly.  h comes out of the last subtraction for free:
        // fma(det, 0.5, -(kappa/2) * trace^2) rounds the same real number as det - kappa * trace^2, halved
        // (scaling by two commutes with rounding; the values are far from the subnormal range).
        // Tiles that touch neither the top nor the bottom of the image (CTA-uniform test) need no row predicate.
        const double half_kappa = 0.5 * kappa;
        const int gy0 = y0 + base;
        const unsigned int row_bytes = (unsigned int)W * (unsigned int)sizeof(double);
        char* dst0 = reinterpret_cast<char*>(resp + ((size_t)f * H + (size_t)gy0) * W + max(gx, 0));
        const bool y_interior_tile = (y0 >= 5) && (y0 + FT_H <= H - 5);
        if (y_interior_tile) {
            const unsigned int xs = x_in ? 1u : 0u;
#pragma unroll
            for (int i = 0; i < 27; i++) {
                const double sa = (double)vxx, sb = (double)vyy, sc = (double)vxy;
                const double trace = __dadd_rn(sa, sb);
                const double det = __dsub_rn(__dmul_rn(sa, sb), __dmul_rn(sc, sc));
                const double h = __fma_rn(det, 0.5, -__dmul_rn(half_kappa, __dmul_rn(trace, trace)));
                const double v = __dadd_rn(fabs(h), h);
                char* dst = dst0 + (size_t)row_bytes * (unsigned int)i;
                asm volatile(
                    "{\n\t.reg .pred r;\n\t"
                    "setp.ne.u32 r, %2, 0;\n\t"
                    "@r st.global.f64 [%0], %1;\n\t}"
                    ::"l"(dst), "d"(v), "r"(xs) : "memory");
                if (i < 26) {
                    const int2 nab = hab[(base + i + 9) * FT_HP + hcol];
                    const int nxx = nab.x, nyy = nab.y, nxy = hxy[(base + i + 9) * FT_HP + hcol];
                    vxx += nxx - rxx[i % 9]; vyy += nyy - ryy[i % 9]; vxy += nxy - rxy[i % 9];
                    rxx[i % 9] = nxx; ryy[i % 9] = nyy; rxy[i % 9] = nxy;
                }
            }
        } else {
            const int n_rows = x_in ? min(27, H - gy0) : 0;           // rows this thread may store
            const int i_lo = 5 - gy0;                                 // interior rows: i_lo <= i < i_lo + i_span
            const unsigned i_span = (unsigned)max(H - 10, 0);
#pragma unroll
            for (int i = 0; i < 27; i++) {
                const double sa = (double)vxx, sb = (double)vyy, sc = (double)vxy;
                const double trace = __dadd_rn(sa, sb);
                const double det = __dsub_rn(__dmul_rn(sa, sb), __dmul_rn(sc, sc));
                const double h = __fma_rn(det, 0.5, -__dmul_rn(half_kappa, __dmul_rn(trace, trace)));
                const double v = __dadd_rn(fabs(h), h);
                char* dst = dst0 + (size_t)row_bytes * (unsigned int)i;
                asm volatile(
                    "{\n\t.reg .pred q, r;\n\t.reg .f64 v;\n\t"
                    "setp.lt.u32 q, %2, %3;\n\t"
                    "selp.f64 v, %1, 0d0000000000000000, q;\n\t"
                    "setp.lt.s32 r, %4, %5;\n\t"
                    "@r st.global.f64 [%0], v;\n\t}"
                    ::"l"(dst), "d"(v), "r"((unsigned)(i - i_lo)), "r"(i_span), "r"(i), "r"(n_rows) : "memory");
                if (i < 26) {
                    const int2 nab = hab[(base + i + 9) * FT_HP + hcol];
                    const int nxx = nab.x, nyy = nab.y, nxy = hxy[(base + i + 9) * FT_HP + hcol];
                    vxx += nxx - rxx[i % 9]; vyy += nyy - ryy[i % 9]; vxy += nxy - rxy[i % 9];
                    rxx[i % 9] = nxx; ryy[i % 9] = nyy; rxy[i % 9] = nxy;
                }
            }
        }
    }
    __syncthreads();   // the H sums are free for the next tile's phase H
  }
}

// ---------------------------------------------------------------------------------------------
// NMS step 1: dense local-maximum detection (window radius r, raster-order tie break).
// ---------------------------------------------------------------------------------------------
constexpr int L_THREADS = 256;     // 8 warps = 8 adjacent column strips
constexpr int L_ROWS = 96;         // rows per warp (plus one block row above and below); multiple of 1, 2 and 3
constexpr int L_CHUNK = 4;         // block rows loaded ahead

// Scores are non-negative doubles, so their bit patterns order like unsigned 64-bit integers.
// A warp streams down a strip of 32 columns, one coalesced 256-byte row per load, nothing staged in shared
// memory.  The strip is cut into BxB blocks (B = 3 for r >= 5): a local maximum of the (2r+1)^2 window must be
// the maximum of its own block and of the 8 blocks around it (all nine lie inside the window when r >= 2B-1).
// Block maxima of a 32-bit monotone image of the score (high word) cost one max3 per B rows down the column and
// four shuffles across it, so the dense pass is a handful of instructions per pixel and rejects all but a
// fraction of a percent.  Survivors are queued and get the exact (2r+1)^2 test with full keys and the
// raster-order tie break, done by the whole warp on data that is L1/L2 resident.
// exact window test, one candidate per lane: window rows from the outside in (a candidate that survived the
// block test can only lose to the part of the window the nine blocks do not cover), a whole row of loads in
// flight, early exit per lane
template <int R>
__device__ __forceinline__ bool localmax_exact_lane(const double* __restrict__ src, int H, int W, int r_rt, int yc, int xc,
                                                    unsigned long long ck) {
    const int r = R > 0 ? R : r_rt;
    bool ok = true;
    for (int a = 0; a <= 2 * r && ok; a++) {
        const int dy = ((a & 1) ? 1 : -1) * (r - (a >> 1));
        const int y = yc + dy;
        if (y < 0 || y >= H) continue;
        const double* row = src + (size_t)y * W + xc;
        if (R > 0) {
            // high words first (4-byte loads, one compare each): a larger one ends the test, an equal one (rare) asks
            // for the low word.  Candidates away from the left / right border skip the column range checks.
            const unsigned int chi = (unsigned int)(ck >> 32), clo = (unsigned int)ck;
            const unsigned int* rowh = reinterpret_cast<const unsigned int*>(row) + 1;
            unsigned int qh[2 * R + 1];
            if (xc >= R && xc + R < W) {
#pragma unroll
                for (int dx = -R; dx <= R; dx++) qh[dx + R] = __ldg(rowh + 2 * dx);
            } else {
#pragma unroll
                for (int dx = -R; dx <= R; dx++) qh[dx + R] = (xc + dx >= 0 && xc + dx < W) ? __ldg(rowh + 2 * dx) : 0u;
            }
            bool gt = false;
            unsigned int tie = 0u;
#pragma unroll
            for (int dx = -R; dx <= R; dx++) {
                if (dy == 0 && dx == 0) continue;
                gt = gt || qh[dx + R] > chi;
                if (qh[dx + R] == chi) tie |= 1u << (dx + R);
            }
            if (gt) { ok = false; break; }
            while (tie) {                                    // equal high words: the low words decide (ties by raster order)
                const int dx = __ffs(tie) - 1 - R;
                tie &= tie - 1u;
                const unsigned int qlo = (xc + dx >= 0 && xc + dx < W) ? __ldg(rowh + 2 * dx - 1) : 0u;
                const bool before = (dy < 0) || (dy == 0 && dx < 0);
                if (before ? (qlo >= clo) : (qlo > clo)) ok = false;
            }
        } else {
            for (int dx = -r; dx <= r; dx++) {
                const unsigned long long qk = (xc + dx >= 0 && xc + dx < W) ? (unsigned long long)__double_as_longlong(__ldg(row + dx)) : 0ull;
                const bool before = (dy < 0) || (dy == 0 && dx < 0);
                if (!(dy == 0 && dx == 0) && (before ? (qk >= ck) : (qk > ck))) ok = false;
            }
        }
    }
    return ok;
}

template <int R, int B>
__global__ void __launch_bounds__(L_THREADS)
harris_localmax(const double* __restrict__ resp, int H, int W, int r_rt, unsigned int lm_cap,
                unsigned long long* __restrict__ lm_key, unsigned int* __restrict__ lm_idx,
                unsigned int* __restrict__ lm_count, unsigned int* __restrict__ cmap, int cbw, int cbh) {
    constexpr unsigned int FULL = 0xFFFFFFFFu;
    constexpr int NBLK = 32 / B, LANES = NBLK * B, OUT_W = (NBLK - 2) * B;   // the outer blocks are halo
    __shared__ unsigned int s_queue[L_THREADS / 32][L_CHUNK * B * OUT_W + 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x * (L_THREADS / 32) + warp;
    if (strip * OUT_W >= W) return;                    // warp-uniform
    const int f = blockIdx.z;
    const double* src = resp + (size_t)f * H * W;
    const int xs = strip * OUT_W - B + lane;
    const bool col_in = lane < LANES && xs >= 0 && xs < W;
    const bool decides = lane >= B && lane < LANES - B && xs < W;
    const int leader = lane - lane % B;
    const bool reject = (R > 0 ? R : r_rt) >= 2 * B - 1;   // the nine blocks lie inside the window (false only for r = 0)
    const int y0 = blockIdx.y * L_ROWS, y1 = min(H, y0 + L_ROWS);
    const int n_brows = (y1 - y0 + B - 1) / B + 2;     // block row k covers rows y0 - B + k*B ...
    unsigned int* queue = s_queue[warp];
    int n_q = 0;
    // the lane that owns a block of the coarse map (block row k of this CTA -> cm_lane[k * cbw])
    unsigned int* cm_lane = (cmap != nullptr && lane == leader && decides)
                                ? cmap + ((size_t)f * cbh + (y0 / B - 1)) * cbw + (strip * (NBLK - 2) - 1 + lane / B) : nullptr;
    unsigned int nb1 = 0u, nb2 = 0u;                   // 3-block row maxima of the two previous block rows
    unsigned int gprev[B];
#pragma unroll
    for (int i = 0; i < B; i++) gprev[i] = 0u;
    for (int k0 = 0; k0 < n_brows; k0 += L_CHUNK) {
        unsigned long long v[L_CHUNK * B];
        const int yf = y0 - B + k0 * B;                // first row of the chunk
        if (yf >= 0 && yf + L_CHUNK * B <= H && k0 + L_CHUNK <= n_brows) {   // interior chunk: no per-row tests
            const double* pr = src + (yf * W + xs);    // running row pointer: one 64-bit add per load
#pragma unroll
            for (int j = 0; j < L_CHUNK * B; j++) {
                v[j] = col_in ? (unsigned long long)__double_as_longlong(__ldg(pr)) : 0ull;
                pr += W;
            }
        } else {
#pragma unroll
            for (int j = 0; j < L_CHUNK * B; j++) {
                const int y = yf + j;
                v[j] = 0ull;
                if (k0 + j / B < n_brows && y >= 0 && y < H && col_in)
                    v[j] = (unsigned long long)__double_as_longlong(__ldg(src + (y * W + xs)));
            }
        }
#pragma unroll
        for (int c = 0; c < L_CHUNK; c++) {
            if (k0 + c >= n_brows) break;              // warp-uniform
            unsigned int g[B];
            unsigned int cm = 0u;
#pragma unroll
            for (int i = 0; i < B; i++) {
                // ceil(key / 2^32): a 32-bit monotone image of the key that is zero only for a zero key
                g[i] = (unsigned int)((v[c * B + i] + 0xFFFFFFFFull) >> 32);
                cm = max(cm, g[i]);
            }
            unsigned int t = cm;                       // block maximum, valid on the block's first lane
#pragma unroll
            for (int d = 1; d < B; d++) t = max(t, __shfl_down_sync(FULL, cm, d));
            // by-product: the block maxima (monotone 32-bit image of the score) as a coarse map of the frame; the
            // entry scan reads it instead of the score map and touches only the blocks at or above its threshold
            if (cm_lane != nullptr && k0 + c >= 1 && k0 + c <= n_brows - 2) cm_lane[(k0 + c) * cbw] = t;
            const unsigned int nb = max(t, max(__shfl_up_sync(FULL, t, B), __shfl_down_sync(FULL, t, B)));
            const unsigned int m9 = __shfl_sync(FULL, max(nb, max(nb1, nb2)), leader);   // 3x3 blocks around block row k-1
            // m9 >= every gprev[i], so equality means "not beaten by the nine blocks".  Cheap test first; the
            // exact candidate bits (with the segment's row range) only when some lane has a hit.
            bool hit = false;
#pragma unroll
            for (int i = 0; i < B; i++) hit = hit || (reject ? (gprev[i] == m9) : (gprev[i] != 0u));
            hit = hit && decides && (m9 != 0u || !reject);
            if (__any_sync(FULL, hit)) {
                const int yp = yf + (c - 1) * B;       // first row of the block row being decided
#pragma unroll
                for (int i = 0; i < B; i++) {
                    const bool cnd = hit && yp + i >= y0 && yp + i < y1 && (reject ? (gprev[i] == m9) : (gprev[i] != 0u));
                    const unsigned int mask = __ballot_sync(FULL, cnd);
                    if (cnd) queue[n_q + __popc(mask & ((1u << lane) - 1u))] = (unsigned int)((yp + i) * W + xs);
                    n_q += __popc(mask);
                }
            }
            nb2 = nb1; nb1 = nb;
#pragma unroll
            for (int i = 0; i < B; i++) gprev[i] = g[i];
        }
        __syncwarp();
        const bool last = k0 + L_CHUNK >= n_brows;
        while (n_q >= 32 || (last && n_q > 0)) {       // batches of 32 queued candidates, one per lane
            const int m = min(32, n_q);
            n_q -= m;
            if (lane < m) {
                const unsigned int p = queue[n_q + lane];
                const int yc = (int)(p / (unsigned)W), xc = (int)(p - (unsigned)yc * W);
                const unsigned long long k = (unsigned long long)__double_as_longlong(__ldg(src + p));
                if (localmax_exact_lane<R>(src, H, W, r_rt, yc, xc, k)) {
                    const unsigned int slot = atomicAdd(&lm_count[f], 1u);
                    if (slot < lm_cap) {
                        lm_key[(size_t)f * lm_cap + slot] = k;
                        lm_idx[(size_t)f * lm_cap + slot] = p;
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// NMS steps 2-4: one CTA per frame.
// ---------------------------------------------------------------------------------------------
constexpr int N_THREADS = 1024;

// priority: key descending, index ascending.
__device__ __forceinline__ bool prio_ge(unsigned long long k, unsigned int i, unsigned long long tk, unsigned int ti) {
    return k > tk || (k == tk && i <= ti);
}
__device__ __forceinline__ bool prio_gt(unsigned long long k, unsigned int i, unsigned long long tk, unsigned int ti) {
    return k > tk || (k == tk && i < ti);
}

// byte `b` (0 = most significant) of the 96-bit composite (key, ~idx): larger composite = higher priority
__device__ __forceinline__ unsigned int comp_byte(unsigned long long k, unsigned int i, int b) {
    if (b < 8) return (unsigned int)(k >> (56 - 8 * b)) & 0xFFu;
    return ((~i) >> (24 - 8 * (b - 8))) & 0xFFu;
}

// Block-wide radix select of the kth (1-based) highest-priority entry among n entries.
// Returns the threshold entry (tk, ti) to all threads.  Requires 1 <= kth <= n.
__device__ void block_select_kth(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ idxs,
                                 unsigned int n, unsigned int kth, unsigned int* hist /*smem[256]*/,
                                 unsigned int* s_misc /*smem[4]*/, unsigned long long* tk_out, unsigned int* ti_out) {
    __shared__ unsigned int s_wsum[8];
    unsigned long long pk = 0;  // chosen prefix of key
    unsigned int pi = 0;        // chosen prefix of ~idx
    unsigned int remaining = kth;
    for (int b = 0; b < 12; b++) {
        for (int j = threadIdx.x; j < 256; j += blockDim.x) hist[j] = 0;
        __syncthreads();
        for (unsigned int j = threadIdx.x; j < n; j += blockDim.x) {
            const unsigned long long k = keys[j];
            const unsigned int ni = ~idxs[j];
            bool match;
            if (b == 0) match = true;
            else if (b < 8) match = (k >> (64 - 8 * b)) == (pk >> (64 - 8 * b));
            else if (b == 8) match = (k == pk);
            else match = (k == pk) && ((ni >> (32 - 8 * (b - 8))) == (pi >> (32 - 8 * (b - 8))));
            if (match) atomicAdd(&hist[comp_byte(k, ~ni, b)], 1u);
        }
        __syncthreads();
        // which digit holds the `remaining`-th entry counting from the top?  parallel scan over the 256 bins
        // (thread t looks at digit 255 - t); s_misc[2..] carries the per-warp totals.
        {
            const unsigned int t = threadIdx.x, ln = t & 31u, wp = t >> 5;
            unsigned int v = 0, incl = 0;
            if (t < 256u) {
                v = hist[255u - t];
                incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (ln >= (unsigned)o) incl += u;
                }
                if (ln == 31u) s_wsum[wp] = incl;
            }
            __syncthreads();
            if (t < 256u) {
                unsigned int off = 0;
                for (unsigned int w = 0; w < wp; w++) off += s_wsum[w];
                incl += off;
                const unsigned int excl = incl - v;
                if (excl < remaining && incl >= remaining) { s_misc[0] = 255u - t; s_misc[1] = remaining - excl; }
            }
        }
        __syncthreads();
        const unsigned int d = s_misc[0];
        remaining = s_misc[1];
        if (b < 8) pk |= (unsigned long long)d << (56 - 8 * b);
        else pi |= d << (24 - 8 * (b - 8));
        __syncthreads();
    }
    *tk_out = pk;
    *ti_out = ~pi;
}

struct NmsArgs {
    const double* resp;            // [F][H][W]
    int H, W, r, K;
    unsigned int lm_cap;
    const unsigned long long* lm_key;  // [F][lm_cap]
    const unsigned int* lm_idx;
    const unsigned int* lm_count;      // [F]
    unsigned int* sup;                 // [F][bm_words] bitmap: pixel lies in the box of a pick
    unsigned int* mem;                 // [F][bm_words] bitmap: undecided entry of the current band (global fallback only)
    unsigned int bm_words;             // words per frame bitmap (H*W/32 + 2: the range helpers read one word ahead)
    uint4* ent_a;                      // [F][H*W] entries {pixel, key lo, key hi, -} in scan order
    unsigned int* ent_h;               // [F][H*W] high words of the entries' scores (all the binning pass needs)
    uint4* ent_b;                      // [F][H*W] the same entries ordered by priority bin (filled lazily from the top)
    unsigned long long* pick_key;      // [F][lm_cap]
    unsigned int* pick_idx;            // [F][lm_cap]
    int* kp_xy;                        // [F][K][2]
    unsigned int* stats;               // [F][4]: n_lm, n_entries, n_rounds, n_picks
    unsigned long long* thr_key;       // [F] threshold (score bits)
    unsigned int* thr_idx;             // [F] threshold (tie-break index)
    unsigned int* counters;            // [F][4]: entries, picks, min / max high word of the entries' keys
    int bitmaps_in_smem;
    unsigned int smem_bytes;           // dynamic shared memory of the band kernel
    unsigned int band;                 // entries per band (<= N_THREADS)
    // Speculative threshold.  The first K picks usually end far above the K-th best local maximum (most of them are not
    // local maxima of their window), so a call may start from the spec_rank[f]-th best local maximum instead -- the rank
    // the K-th pick of the PREVIOUS call had among the local maxima of frame f, with a margin.  Everything found above a
    // threshold is exact whatever the threshold is (a pixel's fate depends on higher priorities only); if fewer than K
    // picks lie above the speculative one, flag[f] asks for a second pass (pass = 1) from the safe threshold.
    const unsigned int* cmap;          // [F][cbh][cbw] maxima of the 3x3 blocks (ceil(score bits / 2^32)), or null
    int cbw, cbh;
    unsigned int* spec_rank;           // [F] carried between calls (0 = unknown), or null: no speculation
    unsigned int* flag;                // [F] 0 done / 1 second pass needed / 2 speculative pass running
    int pass;
};

// ---- bitmaps: bit p = pixel p (row-major, no row padding) ----
__device__ __forceinline__ bool bm_test(const unsigned int* bm, unsigned int p) { return (bm[p >> 5] >> (p & 31u)) & 1u; }
// set bits [b0, b0 + len), 1 <= len <= 31
__device__ __forceinline__ void bm_set_range(unsigned int* bm, unsigned int b0, int len) {
    const unsigned long long m = ((1ull << len) - 1ull) << (b0 & 31u);
    atomicOr(&bm[b0 >> 5], (unsigned int)m);
    if (m >> 32) atomicOr(&bm[(b0 >> 5) + 1], (unsigned int)(m >> 32));
}
// bits [b0, b0 + len) as the low bits of a word, 1 <= len <= 31
__device__ __forceinline__ unsigned int bm_get_range(const unsigned int* bm, unsigned int b0, int len) {
    const unsigned int w = b0 >> 5;
    return __funnelshift_r(bm[w], bm[w + 1], b0 & 31u) & ((1u << len) - 1u);
}

constexpr int NMS_SELECT_SMEM = 3584;   // local maxima staged in shared memory for the threshold select (42 KB)

// ---- NMS step 2 (one CTA per frame): threshold = K-th best local maximum; the maxima at or above it are picks
// ---- (a local maximum beats its whole window, so nothing can suppress it) and their boxes go into the bitmap.
__global__ void __launch_bounds__(N_THREADS)
harris_nms_select(NmsArgs a) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_misc[4];
    __shared__ unsigned int s_np;
    const int f = blockIdx.x;
    const int H = a.H, W = a.W, r = a.r, K = a.K;
    const unsigned long long* lmk = a.lm_key + (size_t)f * a.lm_cap;
    const unsigned int* lmi = a.lm_idx + (size_t)f * a.lm_cap;
    unsigned int* sup = a.sup + (size_t)f * a.bm_words;       // zeroed by the launcher
    unsigned long long* pk = a.pick_key + (size_t)f * a.lm_cap;
    unsigned int* pi = a.pick_idx + (size_t)f * a.lm_cap;
    const unsigned int n_lm = min(a.lm_count[f], a.lm_cap);
    const int tid = threadIdx.x;
    if (a.pass == 1) {                                        // second pass: only the frames whose speculation fell short
        if (a.flag[f] != 1u) return;
        for (unsigned int w = tid; w < a.bm_words; w += N_THREADS) sup[w] = 0u;
    }
    unsigned int rank = (unsigned int)K;                      // threshold = rank-th best local maximum
    if (a.pass == 0 && a.spec_rank != nullptr) {
        const unsigned int sr = a.spec_rank[f];
        if (sr >= 1u && sr < (unsigned int)K && n_lm >= sr) rank = sr;
    }
    const bool speculative = rank < (unsigned int)K;
    unsigned long long tk = 1ull;                             // "every positive score"
    unsigned int ti = 0xFFFFFFFFu;
    // the list of local maxima is staged in shared memory when it fits: the 12 radix passes then run at
    // shared-memory latency instead of one dependent global load per step
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned long long* kk = lmk;
    const unsigned int* ii = lmi;
    if (n_lm <= (unsigned)NMS_SELECT_SMEM) {
        unsigned long long* sk = reinterpret_cast<unsigned long long*>(smem_raw);
        unsigned int* si = reinterpret_cast<unsigned int*>(sk + NMS_SELECT_SMEM);
        for (unsigned int j = tid; j < n_lm; j += N_THREADS) { sk[j] = lmk[j]; si[j] = lmi[j]; }
        kk = sk; ii = si;
    }
    if (tid == 0) { s_np = 0; if (a.flag != nullptr && a.pass == 0) a.flag[f] = speculative ? 2u : 0u; }   // pass 1 keeps its 1 for scan / bands
    __syncthreads();
    if (n_lm >= rank && K > 0) block_select_kth(kk, ii, n_lm, rank, hist, s_misc, &tk, &ti);
    for (unsigned int j = tid; j < n_lm; j += N_THREADS) {
        const unsigned long long k = kk[j];
        const unsigned int i = ii[j];
        if (prio_ge(k, i, tk, ti)) {
            const unsigned int s = atomicAdd(&s_np, 1u);
            pk[s] = k; pi[s] = i;
        }
    }
    __syncthreads();
    const unsigned int n_p = s_np;
    const int win = 2 * r + 1;
    for (unsigned int it = tid; it < n_p * (unsigned)win; it += N_THREADS) {   // one box row per thread
        const unsigned int j = it / (unsigned)win;
        const int dy = (int)(it - j * (unsigned)win) - r;
        const unsigned int p = pi[j];
        const int py = (int)(p / (unsigned)W), px = (int)(p - (unsigned)py * W);
        const int y = py + dy;
        if (y < 0 || y >= H) continue;
        const int x0 = max(px - r, 0), x1 = min(px + r, W - 1);
        bm_set_range(sup, (unsigned int)(y * W + x0), x1 - x0 + 1);
    }
    if (tid == 0) {
        a.thr_key[f] = tk; a.thr_idx[f] = ti;
        a.counters[f * 4 + 0] = 0;               // entries (filled by the scan)
        a.counters[f * 4 + 1] = n_p;             // picks so far
        a.counters[f * 4 + 2] = 0xFFFFFFFFu;     // min / max high word of the entries' keys
        a.counters[f * 4 + 3] = 0u;
    }
}

// ---- NMS step 3 (grid-wide, one read of the score map): entries = pixels at or above the threshold outside
// ---- every box.  Only these can still become picks.
constexpr int SCAN_PER_THREAD = 8;

// LOOP = false: one CTA per chunk of 2048 pixels (grid.x = number of chunks); LOOP = true: the second pass of a speculative
// call -- few CTAs per frame that leave at once unless the frame asked for it, and walk the chunks otherwise
template <bool LOOP>
__global__ void __launch_bounds__(256, 8)
harris_nms_scan(NmsArgs a) {
    __shared__ unsigned int s_base;
    __shared__ unsigned int s_warp[8], s_min[8], s_max[8];
    // per warp: the pixels that pass the cheap test, then the kept ones {pixel, score lo, score hi}
    __shared__ unsigned int s_qp[8][32 * SCAN_PER_THREAD], s_ql[8][32 * SCAN_PER_THREAD], s_qh[8][32 * SCAN_PER_THREAD];
    const int f = blockIdx.y;
    if (LOOP && a.flag[f] != 1u) return;                   // second pass: only the frames whose speculation fell short
    const unsigned int npx = (unsigned int)a.H * a.W;
    const double* resp = a.resp + (size_t)f * npx;
    const unsigned int* sup = a.sup + (size_t)f * a.bm_words;
    uint4* ent = a.ent_a + (size_t)f * npx;
    unsigned int* enth = a.ent_h + (size_t)f * npx;
    const unsigned long long tk = a.thr_key[f];
    const unsigned int ti = a.thr_idx[f];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int lt = (1u << lane) - 1u;
    const unsigned int n_chunks = (npx + 256u * SCAN_PER_THREAD - 1u) / (256u * SCAN_PER_THREAD);
  for (unsigned int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const unsigned int base = chunk * (256u * SCAN_PER_THREAD) + threadIdx.x;
    const unsigned int cta_row = (chunk * (256u * SCAN_PER_THREAD)) / (unsigned)a.W;   // uniform: row of the CTA's first pixel
    unsigned long long k[SCAN_PER_THREAD];
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; j++) {            // all loads first (coalesced, 8 in flight per thread)
        const unsigned int p = base + j * 256u;
        k[j] = p < npx ? (unsigned long long)__double_as_longlong(__ldg(resp + p)) : 0ull;
    }
    // Cheap test first (high word against the threshold's): about a fifth of the pixels pass, in blobs, so nearly
    // every warp has some at every step -- evaluating the exact test under predication would cost all 32 lanes
    // each time.  The passing pixels are queued per warp instead and the exact comparison, the bitmap lookup and
    // the emission then run on dense lanes.
    const unsigned int thi = (unsigned int)(tk >> 32);
    unsigned int *qp = s_qp[warp], *ql = s_ql[warp], *qh = s_qh[warp];
    unsigned int nq = 0;
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; j++) {
        const bool maybe = (unsigned int)(k[j] >> 32) >= thi && k[j] != 0ull;
        const unsigned int m = __ballot_sync(0xFFFFFFFFu, maybe);
        if (maybe) { const unsigned int t = nq + __popc(m & lt); qp[t] = base + j * 256u; ql[t] = (unsigned int)k[j]; qh[t] = (unsigned int)(k[j] >> 32); }
        nq += __popc(m);
    }
    __syncwarp();
    unsigned int nk = 0, hmin = 0xFFFFFFFFu, hmax = 0u;    // kept entries are compacted to the front of the queue
    for (unsigned int i0 = 0; i0 < nq; i0 += 32) {
        const unsigned int i = i0 + lane;
        uint4 e = make_uint4(0u, 0u, 0u, 0u);
        bool keep = false;
        if (i < nq) {
            e = make_uint4(qp[i], ql[i], qh[i], 0u);
            keep = prio_ge(((unsigned long long)e.z << 32) | e.y, e.x, tk, ti) && !((__ldg(sup + (e.x >> 5)) >> (e.x & 31u)) & 1u);
        }
        const unsigned int m = __ballot_sync(0xFFFFFFFFu, keep);
        __syncwarp();                                        // all lanes have read their slot before slots <= i0 + 31 are rewritten
        if (keep) {
            const unsigned int t = nk + __popc(m & lt);
            qp[t] = e.x; ql[t] = e.y; qh[t] = e.z;
            hmin = min(hmin, e.z); hmax = max(hmax, e.z);
        }
        nk += __popc(m);
    }
    hmin = __reduce_min_sync(0xFFFFFFFFu, hmin);
    hmax = __reduce_max_sync(0xFFFFFFFFu, hmax);
    if (lane == 0) { s_warp[warp] = nk; s_min[warp] = hmin; s_max[warp] = hmax; }
    __syncthreads();
    if (threadIdx.x == 0) {                                  // one global atomic per CTA
        unsigned int t = 0, mn = 0xFFFFFFFFu, mx = 0u;
        for (int w = 0; w < 8; w++) {
            const unsigned int c = s_warp[w]; s_warp[w] = t; t += c;
            mn = min(mn, s_min[w]); mx = max(mx, s_max[w]);
        }
        s_base = t ? atomicAdd(&a.counters[f * 4 + 0], t) : 0u;
        if (t) { atomicMin(&a.counters[f * 4 + 2], mn); atomicMax(&a.counters[f * 4 + 3], mx); }
    }
    __syncthreads();
    const unsigned int o = s_base + s_warp[warp];
    for (unsigned int i = lane; i < nk; i += 32) {
        const uint4 e = make_uint4(qp[i], ql[i], qh[i], 0u);
        unsigned int py = cta_row, px = e.x - cta_row * (unsigned)a.W;   // no division: walk from the CTA's first row
        while (px >= (unsigned)a.W) { px -= (unsigned)a.W; py++; }
        enth[o + i] = e.z;
        ent[o + i] = make_uint4(e.x, e.y, e.z, (py << 16) | px);
    }
    if (!LOOP) break;
    __syncthreads();                                       // the queues and s_base are reused by the next chunk
  }
}

// ---- NMS step 3, coarse-map variant (3x3 blocks, radius >= 5): the same entry list from the block maxima the
// ---- local-maximum pass left behind.  A block whose maximum is below the threshold, or whose nine pixels all lie in
// ---- boxes, holds no entry and its part of the score map is never read.  With the speculative threshold about one
// ---- block in twenty survives, so the pass reads the coarse map (0.44 B/pixel) and a few percent of the score map
// ---- instead of all of it.  (From the safe threshold a third of the blocks survive, scattered: the dense scan is the
// ---- faster one then, and the second pass uses it.)  A warp owns SB_ROWS x 32 blocks: the block maxima and the box
// ---- bits of all of them are requested first, then the pixels of the surviving blocks are visited 128 at a time.
constexpr int SB_ROWS = 4;            // block rows per warp

__global__ void __launch_bounds__(256)
harris_nms_scan_blocks(NmsArgs a) {
    __shared__ unsigned char s_blk[8][SB_ROWS * 32];
    const int f = blockIdx.z;
    const int H = a.H, W = a.W, cbw = a.cbw, cbh = a.cbh;
    const unsigned int npx = (unsigned int)H * W;
    const double* resp = a.resp + (size_t)f * npx;
    const unsigned int* sup = a.sup + (size_t)f * a.bm_words;
    const unsigned int* cm = a.cmap + (size_t)f * cbw * cbh;
    uint4* ent = a.ent_a + (size_t)f * npx;
    unsigned int* enth = a.ent_h + (size_t)f * npx;
    const unsigned long long tk = a.thr_key[f];
    const unsigned int ti = a.thr_idx[f];
    const unsigned int gthr = (unsigned int)((tk + 0xFFFFFFFFull) >> 32);   // block maxima are ceil(key / 2^32)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int lt = (1u << lane) - 1u;
    const int gbx = blockIdx.x * 32 + lane;
    const int by0 = (blockIdx.y * 8 + warp) * SB_ROWS;
    if (by0 >= cbh) return;                                    // warp-uniform
    unsigned char* blk = s_blk[warp];
    const int x = 3 * gbx, len = min(3, W - x);
    unsigned int c[SB_ROWS];
#pragma unroll
    for (int rb = 0; rb < SB_ROWS; rb++)
        c[rb] = (gbx < cbw && by0 + rb < cbh) ? __ldg(cm + (size_t)(by0 + rb) * cbw + gbx) : 0u;
    bool active[SB_ROWS];
    unsigned int bits[SB_ROWS][3];
#pragma unroll
    for (int rb = 0; rb < SB_ROWS; rb++) {
        active[rb] = c[rb] >= gthr && c[rb] != 0u;
#pragma unroll
        for (int dy = 0; dy < 3; dy++) {
            const int y = 3 * (by0 + rb) + dy;
            bits[rb][dy] = (active[rb] && y < H) ? bm_get_range(sup, (unsigned int)(y * W + x), len) : 0xFFFFFFFFu;
        }
    }
    int nb = 0;
#pragma unroll
    for (int rb = 0; rb < SB_ROWS; rb++) {
        const unsigned int full = (1u << len) - 1u;            // all nine pixels inside boxes: nothing to emit
        const bool boxed = (bits[rb][0] & full) == full && (bits[rb][1] & full) == full && (bits[rb][2] & full) == full;
        const bool act = active[rb] && !boxed;
        const unsigned int m = __ballot_sync(0xFFFFFFFFu, act);
        if (act) blk[nb + __popc(m & lt)] = (unsigned char)((rb << 5) | lane);
        nb += __popc(m);
    }
    if (nb == 0) return;
    __syncwarp();
    const int n_items = nb * 9;
    unsigned int hmin = 0xFFFFFFFFu, hmax = 0u;
    bool any = false;
    for (int i0 = 0; i0 < n_items; i0 += 128) {
        unsigned int p[4];
        unsigned long long k[4];
        unsigned int sw[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + 32 * u + lane;
            p[u] = 0xFFFFFFFFu; k[u] = 0ull; sw[u] = 0xFFFFFFFFu;
            if (i < n_items) {
                const int j = i / 9, q = i - 9 * j, dy = q / 3, dx = q - 3 * dy;
                const int e = blk[j];
                const int yy = 3 * (by0 + (e >> 5)) + dy, xx = 3 * (blockIdx.x * 32 + (e & 31)) + dx;
                if (yy < H && xx < W) {
                    p[u] = (unsigned int)(yy * W + xx);
                    k[u] = (unsigned long long)__double_as_longlong(__ldg(resp + p[u]));
                    sw[u] = __ldg(sup + (p[u] >> 5));
                }
            }
        }
        bool keep[4];
        unsigned int mk[4], tot = 0;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            keep[u] = p[u] != 0xFFFFFFFFu && k[u] != 0ull && prio_ge(k[u], p[u], tk, ti) && !((sw[u] >> (p[u] & 31u)) & 1u);
            mk[u] = __ballot_sync(0xFFFFFFFFu, keep[u]);
            tot += __popc(mk[u]);
        }
        if (tot == 0) continue;                                // warp-uniform
        any = true;
        unsigned int o = 0;
        if (lane == 0) o = atomicAdd(&a.counters[f * 4 + 0], tot);   // few warps per frame get here: no same-address pile-up
        o = __shfl_sync(0xFFFFFFFFu, o, 0);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (keep[u]) {
                const unsigned int t = o + __popc(mk[u] & lt);
                const unsigned int hi = (unsigned int)(k[u] >> 32);
                const unsigned int py = p[u] / (unsigned)W, px = p[u] - py * (unsigned)W;
                enth[t] = hi;
                ent[t] = make_uint4(p[u], (unsigned int)k[u], hi, (py << 16) | px);
                hmin = min(hmin, hi); hmax = max(hmax, hi);
            }
            o += __popc(mk[u]);
        }
    }
    if (!any) return;
    hmin = __reduce_min_sync(0xFFFFFFFFu, hmin);
    hmax = __reduce_max_sync(0xFFFFFFFFu, hmax);
    if (lane == 0) { atomicMin(&a.counters[f * 4 + 2], hmin); atomicMax(&a.counters[f * 4 + 3], hmax); }
}

// ---- NMS step 4 (one CTA per frame): the entries are ordered into priority bins (high word of the score, 2048
// ---- bins over the frame's range) and the bins are processed from the top in bands of about a thousand
// ---- entries.  When a band starts, everything of higher priority is decided, so an entry competes only with
// ---- the undecided entries of its own band: it becomes a pick as soon as none of them with higher priority
// ---- lies in its window.  "Suppressed" and "undecided member of this band" are two bitmaps in shared memory,
// ---- so a window test is 2r+1 funnel shifts and a score load only for the rare neighbour found.  The loop
// ---- stops when the picks decided so far (local maxima + band picks down to the current bin) number K.
constexpr int NMS_BINS = 2048;
constexpr int NMS_BAND = 1024;       // entries per band at most (one per thread); the launcher halves it for <= 2048 keypoints
constexpr int NMS_NEW_CAP = 1024;    // new picks per round (the surplus waits for the next round)
constexpr int NMS_LAZY = 8192;       // entries put in order per scatter pass
constexpr int NMS_HASH = 2048;       // open-addressing table for the (at most NMS_BAND) undecided entries of a band
constexpr unsigned int NMS_EMPTY = 0xFFFFFFFFu;
constexpr size_t NMS_TABLES_BYTES = (size_t)(3 * NMS_BINS + NMS_NEW_CAP) * 4 + (size_t)NMS_HASH * 16 + (size_t)NMS_NEW_CAP * 8;

__device__ __forceinline__ unsigned int nms_hash(unsigned int p) { return (p * 2654435761u) >> 21; }   // 11 bits

#ifdef VO_NMS_TIMING
#define VO_NMS_T0() t_x = clock64()
#define VO_NMS_T1(acc) acc += clock64() - t_x
#else
#define VO_NMS_T0()
#define VO_NMS_T1(acc)
#endif

// score of an undecided entry of the current band (all of them are in the table {pixel, key lo, key hi, -})
__device__ __forceinline__ unsigned long long nms_lookup(const uint4* tab, unsigned int q) {
    unsigned int h = nms_hash(q);
    uint4 t = tab[h];
    for (int n = 0; t.x != q && t.x != NMS_EMPTY && n < NMS_HASH; n++) {
        h = (h + 1u) & (NMS_HASH - 1);
        t = tab[h];
    }
    return t.x == q ? (((unsigned long long)t.z << 32) | t.y) : 0ull;
}

// an undecided, unsuppressed entry of the band with higher priority inside the window of (py, px), or NMS_EMPTY
template <int R>
__device__ __forceinline__ unsigned int nms_find_blocker(const unsigned int* M, const unsigned int* S, const uint4* tab,
                                                         int H, int W, int r_rt, int py, int px,
                                                         unsigned int p, unsigned long long k) {
    const int r = R > 0 ? R : r_rt;
    const int x0 = max(px - r, 0), len = min(px + r, W - 1) - x0 + 1;
    if constexpr (R > 0) {
        unsigned int fld[2 * R + 1], any = 0u;
#pragma unroll
        for (int i = 0; i <= 2 * R; i++) {               // all window rows first: 2 loads each, independent
            const int y = py + i - R;
            fld[i] = (y >= 0 && y < H) ? bm_get_range(M, (unsigned int)(y * W + x0), len) : 0u;
        }
        fld[R] &= ~(1u << (px - x0));
#pragma unroll
        for (int i = 0; i <= 2 * R; i++) any |= fld[i];
        if (!any) return NMS_EMPTY;
#pragma unroll
        for (int i = 0; i <= 2 * R; i++) {
            unsigned int f = fld[i];
            const unsigned int q0 = (unsigned int)((py + i - R) * W + x0);
            while (f) {
                const unsigned int q = q0 + (unsigned int)(__ffs(f) - 1);
                f &= f - 1u;
                const bool dead = bm_test(S, q);             // its bit just lingers; both loads issue together
                const unsigned long long kq = nms_lookup(tab, q);
                if (!dead && prio_gt(kq, q, k, p)) return q;
            }
        }
        return NMS_EMPTY;
    } else {
        for (int dy = -r; dy <= r; dy++) {
            const int y = py + dy;
            if (y < 0 || y >= H) continue;
            const unsigned int q0 = (unsigned int)(y * W + x0);
            unsigned int f = bm_get_range(M, q0, len);
            if (dy == 0) f &= ~(1u << (px - x0));
            while (f) {
                const unsigned int q = q0 + (unsigned int)(__ffs(f) - 1);
                f &= f - 1u;
                const bool dead = bm_test(S, q);
                const unsigned long long kq = nms_lookup(tab, q);
                if (!dead && prio_gt(kq, q, k, p)) return q;
            }
        }
        return NMS_EMPTY;
    }
}

// the new picks of a round suppress their boxes (one box row per thread) and are appended to the pick list
__device__ __forceinline__ void nms_apply_picks(unsigned int* S, unsigned int* M, const unsigned int* newp, const unsigned long long* newk,
                                                unsigned int n_new, int H, int W, int r, unsigned long long* pk, unsigned int* pi,
                                                unsigned int picks_base) {
    const unsigned int win = 2u * r + 1u;
    for (unsigned int it = threadIdx.x; it < n_new * win; it += N_THREADS) {
        const unsigned int j = it / win;
        const int dy = (int)(it - j * win) - r;
        const unsigned int p = newp[j];
        const int py = (int)(p / (unsigned)W), px = (int)(p - (unsigned)py * W);
        const int y = py + dy;
        if (dy == 0) {
            atomicAnd(&M[p >> 5], ~(1u << (p & 31u)));
            pk[picks_base + j] = newk[j]; pi[picks_base + j] = p;
        }
        if (y < 0 || y >= H) continue;
        const int x0 = max(px - r, 0), x1 = min(px + r, W - 1);
        bm_set_range(S, (unsigned int)(y * W + x0), x1 - x0 + 1);
    }
}

template <int R>
__global__ void __launch_bounds__(N_THREADS)
harris_nms_bands(NmsArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_misc[4];
    __shared__ unsigned int s_cnt[8];          // 0/1: new picks, 2/3: blocked (double buffered by round), 4: picks, 5: gather
    __shared__ unsigned int s_ws[32], s_wl[32];
#ifdef VO_NMS_TIMING
    __shared__ unsigned int s_dbg[8];
    if (threadIdx.x < 8) s_dbg[threadIdx.x] = 0;
#endif
    const int f = blockIdx.x;
    if (a.pass == 1 && a.flag[f] != 1u) return;            // second pass: only the frames whose speculation fell short
    const int H = a.H, W = a.W, r = a.r, K = a.K;
    const unsigned int npx = (unsigned int)H * W;
    const double* resp = a.resp + (size_t)f * npx;
    const uint4* ent_a = a.ent_a + (size_t)f * npx;
    const unsigned int* ent_h = a.ent_h + (size_t)f * npx;
    uint4* ent_b = a.ent_b + (size_t)f * npx;
    unsigned long long* pk = a.pick_key + (size_t)f * a.lm_cap;
    unsigned int* pi = a.pick_idx + (size_t)f * a.lm_cap;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    unsigned int* endR = reinterpret_cast<unsigned int*>(smem_raw);   // [BINS] entries with rank <= R (rank 0 = top bin)
    unsigned int* cur = endR + NMS_BINS;                              // [BINS] scatter cursors
    unsigned int* lmR = cur + NMS_BINS;                               // [BINS] local-maximum picks with rank <= R
    unsigned int* newp = lmR + NMS_BINS;                              // [NEW_CAP]
    uint4* tab = reinterpret_cast<uint4*>(newp + NMS_NEW_CAP);        // [HASH] pixel -> score of the band's undecided entries
    unsigned long long* newk = reinterpret_cast<unsigned long long*>(tab + NMS_HASH);   // [NEW_CAP] scores of the new picks
    unsigned int* S = a.sup + (size_t)f * a.bm_words;
    unsigned int* M = a.mem ? a.mem + (size_t)f * a.bm_words : nullptr;
    if (a.bitmaps_in_smem) {
        unsigned int* gs = S;
        S = reinterpret_cast<unsigned int*>(newk + NMS_NEW_CAP);
        M = S + a.bm_words;
        for (unsigned int w = tid; w < a.bm_words; w += N_THREADS) { S[w] = gs[w]; M[w] = 0u; }
    }
    const unsigned int n = a.counters[f * 4 + 0];
    const unsigned int n_p0 = a.counters[f * 4 + 1];
    const unsigned int gmin = a.counters[f * 4 + 2], gmax = a.counters[f * 4 + 3];
    const unsigned int n_lm = min(a.lm_count[f], a.lm_cap);
    for (int j = tid; j < NMS_BINS; j += N_THREADS) { endR[j] = 0u; lmR[j] = 0u; }
    for (int j = tid; j < NMS_HASH; j += N_THREADS) tab[j].x = NMS_EMPTY;
    if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0; s_cnt[2] = 0; s_cnt[3] = 0; s_cnt[4] = n_p0; s_cnt[5] = 0; }
    __syncthreads();
    unsigned int rounds = 0;
    unsigned int n_picks_run = n_p0;           // picks so far (uniform across the CTA)
#ifdef VO_NMS_TIMING
    long long t_0 = clock64(), t_1 = t_0, t_2 = t_0, t_x = 0, acc_i = 0, acc_a = 0, acc_b = 0;
    unsigned int n_bands = 0;
#endif
    int shift = 0;                             // bin rank of a score: (gmax - high word) >> shift, in [0, NMS_BINS)
    if (n > 0) {
        while (((gmax - gmin) >> shift) >= (unsigned)NMS_BINS) shift++;
        // ---- counting sort of the entries by bin rank ----
        for (unsigned int j0 = 0; j0 < n; j0 += 8 * N_THREADS) {          // loads batched: this loop is latency bound
            unsigned int z[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { const unsigned int j = j0 + u * N_THREADS + tid; z[u] = j < n ? ent_h[j] : 0u; }
#pragma unroll
            for (int u = 0; u < 8; u++) if (j0 + u * N_THREADS + tid < n) atomicAdd(&endR[(gmax - z[u]) >> shift], 1u);
        }
        for (unsigned int j = tid; j < n_p0; j += N_THREADS) {
            const unsigned int h = (unsigned int)(pk[j] >> 32);
            const unsigned int rk = h >= gmax ? 0u : (gmax - h) >> shift;
            if (rk < (unsigned)NMS_BINS) atomicAdd(&lmR[rk], 1u);       // below every entry: never needed
        }
        __syncthreads();
        {   // inclusive scans over the 2048 ranks, two per thread
            const unsigned int a0 = endR[2 * tid], a1 = endR[2 * tid + 1];
            const unsigned int l0 = lmR[2 * tid], l1 = lmR[2 * tid + 1];
            unsigned int sa = a0 + a1, sl = l0 + l1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int va = __shfl_up_sync(0xFFFFFFFFu, sa, o), vl = __shfl_up_sync(0xFFFFFFFFu, sl, o);
                if (lane >= o) { sa += va; sl += vl; }
            }
            if (lane == 31) { s_ws[warp] = sa; s_wl[warp] = sl; }
            __syncthreads();
            unsigned int oa = 0, ol = 0;
            for (int w = 0; w < warp; w++) { oa += s_ws[w]; ol += s_wl[w]; }
            const unsigned int ea = oa + sa - (a0 + a1), el = ol + sl - (l0 + l1);   // exclusive
            cur[2 * tid] = ea; cur[2 * tid + 1] = ea + a0;
            endR[2 * tid] = ea + a0; endR[2 * tid + 1] = ea + a0 + a1;
            lmR[2 * tid] = el + l0; lmR[2 * tid + 1] = el + l0 + l1;
        }
        __syncthreads();
        // The ordered list is filled lazily from the top: the loop below usually stops after a few thousand entries,
        // so only the ranks that hold the first NMS_LAZY entries are scattered now (one pass over the high words, full
        // entries fetched for those ranks only); extend() adds the next ranks if the bands get that far.
        unsigned int sorted_upto = 0;                        // entries [0, sorted_upto) of ent_b are in place
        int r_sorted = -1;                                   // ... they are the ranks <= r_sorted
        auto extend = [&](unsigned int needed) {
            while (sorted_upto < needed && sorted_upto < n) {
                int lo = r_sorted + 1, hi = NMS_BINS - 1;    // smallest rank whose cumulative count reaches the target
                const unsigned int target = min(n, max(needed, sorted_upto + (unsigned)NMS_LAZY));
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (endR[mid] >= target) hi = mid; else lo = mid + 1; }
                const unsigned int r_lo = (unsigned int)(r_sorted + 1), r_hi = (unsigned int)lo;
                for (unsigned int j0 = 0; j0 < n; j0 += 8 * N_THREADS) {
                    unsigned int z[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) { const unsigned int j = j0 + u * N_THREADS + tid; z[u] = j < n ? ent_h[j] : 0u; }
                    // slots first, then the selected entries in two groups of four: their loads are in flight together
                    // (one load -> store chain per entry made this loop 14 % of the kernel)
                    unsigned int slot[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const unsigned int j = j0 + u * N_THREADS + tid;
                        const unsigned int rk = (gmax - z[u]) >> shift;
                        slot[u] = (j < n && rk >= r_lo && rk <= r_hi) ? atomicAdd(&cur[rk], 1u) : NMS_EMPTY;
                    }
#pragma unroll
                    for (int g = 0; g < 2; g++) {
                        uint4 e[4];
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (slot[4 * g + u] != NMS_EMPTY) e[u] = __ldg(&ent_a[j0 + (4 * g + u) * N_THREADS + tid]);
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (slot[4 * g + u] != NMS_EMPTY) ent_b[slot[4 * g + u]] = e[u];
                    }
                }
                r_sorted = lo;
                sorted_upto = endR[lo];
                __syncthreads();
            }
        };
        extend(1u);
#ifdef VO_NMS_TIMING
        t_1 = clock64();
#endif
        // ---- bands ----
        // band = as many whole ranks as fit NMS_BAND entries (at least one rank); its range depends only on the
        // bin table, so the entries of the next band are fetched while the current one is processed
        auto band_end = [&](unsigned int from, int r_from, int* r_out) -> unsigned int {
            int lo = r_from + 1, hi = NMS_BINS - 1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (endR[mid] > from) hi = mid; else lo = mid + 1; }
            int r_new = lo;                              // first non-empty rank
            hi = NMS_BINS - 1;
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (endR[mid] <= from + a.band) lo = mid; else hi = mid - 1; }
            if (endR[lo] <= from + a.band) r_new = max(r_new, lo);
            *r_out = r_new;
            return endR[r_new];
        };
        int par = 0, r_cur = -1, r_next = -1;
        unsigned int b0 = 0, b1 = band_end(0u, -1, &r_cur), b2 = b1;
        uint4 e_next = make_uint4(0u, 0u, 0u, 0u);
        extend(b1);
        if (b0 + tid < b1) e_next = ent_b[b0 + tid];
        while (true) {
            VO_NMS_T0();
            const uint4 ec = e_next;
            if (b1 < n) {
                b2 = band_end(b1, r_cur, &r_next);
                extend(b2);
                if (b1 + tid < b2) e_next = ent_b[b1 + tid];
            }
            const bool single = b1 - b0 <= (unsigned)N_THREADS;    // the normal case: one entry per thread, scores hashed
            if (single) {
                // ---------------- one entry per thread, kept in registers ----------------
                const unsigned int p = ec.x;
                const unsigned long long k = ((unsigned long long)ec.z << 32) | ec.y;
                const int py = (int)(ec.w >> 16), px = (int)(ec.w & 0xFFFFu);
                bool und = (b0 + tid < b1) && !bm_test(S, p);
                unsigned int my_h = 0, blocker = NMS_EMPTY;
                if (und) {
                    atomicOr(&M[p >> 5], 1u << (p & 31u));
                    my_h = nms_hash(p);
                    while (atomicCAS(&tab[my_h].x, NMS_EMPTY, p) != NMS_EMPTY) my_h = (my_h + 1u) & (NMS_HASH - 1);
                    tab[my_h].y = ec.y; tab[my_h].z = ec.z;
                }
                const bool inserted = und;
                __syncthreads();
                VO_NMS_T1(acc_i);
                while (true) {
                    VO_NMS_T0();
                    // phase A: an undecided entry is blocked while an undecided neighbour of higher priority exists
                    unsigned int blk = 0;
                    bool is_new = false;
                    if (und) {
                        if (bm_test(S, p)) {                                   // suppressed by a pick of the last round
                            und = false;
                            atomicAnd(&M[p >> 5], ~(1u << (p & 31u)));
                        } else {
                            if (blocker == NMS_EMPTY || !bm_test(M, blocker) || bm_test(S, blocker))
                                blocker = nms_find_blocker<R>(M, S, tab, H, W, r, py, px, p, k);
                            blk = blocker != NMS_EMPTY ? 1u : 0u;
                            is_new = blocker == NMS_EMPTY;
                        }
                    }
                    {   // one atomic per warp for the slots of the new picks
                        const unsigned int mnew = __ballot_sync(0xFFFFFFFFu, is_new);
                        unsigned int base_s = 0;
                        if (lane == 0 && mnew) base_s = atomicAdd(&s_cnt[par], (unsigned int)__popc(mnew));
                        base_s = __shfl_sync(0xFFFFFFFFu, base_s, 0);
                        if (is_new) {
                            const unsigned int s = base_s + __popc(mnew & ((1u << lane) - 1u));
                            if (s < (unsigned)NMS_NEW_CAP) { newp[s] = p; newk[s] = k; und = false; }
                            else blk = 1u;                                     // list full: next round
                        }
                    }
                    blk = __reduce_add_sync(0xFFFFFFFFu, blk);
                    if (lane == 0 && blk) atomicAdd(&s_cnt[2 + par], blk);
#ifdef VO_NMS_TIMING
                    {
                        const unsigned int nu = __popc(__ballot_sync(0xFFFFFFFFu, und || is_new));
                        if (lane == 0) {
                            const unsigned int dtw = (unsigned int)(clock64() - t_x);
                            atomicMax(&s_dbg[4 + (rounds & 1)], dtw); atomicAdd(&s_dbg[1], dtw); atomicAdd(&s_dbg[3], nu);
                        }
                    }
#endif
                    __syncthreads();
                    VO_NMS_T1(acc_a);
#ifdef VO_NMS_TIMING
                    if (tid == 0) { s_dbg[0] += s_dbg[4 + (rounds & 1)]; s_dbg[4 + ((rounds + 1) & 1)] = 0; }
#endif
                    VO_NMS_T0();
                    const unsigned int n_new = min(s_cnt[par], (unsigned)NMS_NEW_CAP);
                    const unsigned int n_blk = s_cnt[2 + par];
                    if (tid == 0) { s_cnt[par ^ 1] = 0; s_cnt[2 + (par ^ 1)] = 0; }
                    nms_apply_picks(S, M, newp, newk, n_new, H, W, r, pk, pi, n_picks_run);
                    n_picks_run += n_new;
                    if (n_blk == 0 && inserted) tab[my_h].x = NMS_EMPTY;       // last round: leave the table empty
                    __syncthreads();
                    VO_NMS_T1(acc_b);
                    par ^= 1;
                    rounds++;
                    if (n_blk == 0) break;
                }
            } else {
                // ---------------- a single oversize bin (massive ties): strips of entries, scores from global memory ----------------
                for (unsigned int j = b0 + tid; j < b1; j += N_THREADS) {
                    const unsigned int p = ent_b[j].x;
                    if (!bm_test(S, p)) atomicOr(&M[p >> 5], 1u << (p & 31u));
                }
                __syncthreads();
                while (true) {
                    unsigned int blk = 0;
                    for (unsigned int j = b0 + tid; j < b1; j += N_THREADS) {
                        const uint4 e = ent_b[j];
                        const unsigned int p = e.x;
                        if (!bm_test(M, p)) continue;
                        if (bm_test(S, p)) { atomicAnd(&M[p >> 5], ~(1u << (p & 31u))); continue; }
                        const unsigned long long k = ((unsigned long long)e.z << 32) | e.y;
                        const int py = (int)(p / (unsigned)W), px = (int)(p - (unsigned)py * W);
                        const int x0 = max(px - r, 0), len = min(px + r, W - 1) - x0 + 1;
                        bool blocked = false;
                        for (int dy = -r; dy <= r && !blocked; dy++) {
                            const int y = py + dy;
                            if (y < 0 || y >= H) continue;
                            const unsigned int q0 = (unsigned int)(y * W + x0);
                            unsigned int fld = bm_get_range(M, q0, len);
                            if (dy == 0) fld &= ~(1u << (px - x0));
                            while (fld && !blocked) {
                                const unsigned int q = q0 + (unsigned int)(__ffs(fld) - 1);
                                fld &= fld - 1u;
                                if (bm_test(S, q)) continue;                   // dead, its bit just lingers
                                blocked = prio_gt((unsigned long long)__double_as_longlong(resp[q]), q, k, p);
                            }
                        }
                        if (!blocked) {
                            const unsigned int s = atomicAdd(&s_cnt[par], 1u);
                            if (s < (unsigned)NMS_NEW_CAP) { newp[s] = p; newk[s] = k; } else blocked = true;
                        }
                        if (blocked) blk++;
                    }
                    blk = __reduce_add_sync(0xFFFFFFFFu, blk);
                    if (lane == 0 && blk) atomicAdd(&s_cnt[2 + par], blk);
                    __syncthreads();
                    const unsigned int n_new = min(s_cnt[par], (unsigned)NMS_NEW_CAP);
                    const unsigned int n_blk = s_cnt[2 + par];
                    if (tid == 0) { s_cnt[par ^ 1] = 0; s_cnt[2 + (par ^ 1)] = 0; }
                    nms_apply_picks(S, M, newp, newk, n_new, H, W, r, pk, pi, n_picks_run);
                    n_picks_run += n_new;
                    __syncthreads();
                    par ^= 1;
                    rounds++;
                    if (n_blk == 0) break;
                }
            }
#ifdef VO_NMS_TIMING
            n_bands++;
#endif
            if (lmR[r_cur] + (n_picks_run - n_p0) >= (unsigned)K) break;   // the K best picks are all known
            if (b1 >= n) break;
            b0 = b1; b1 = b2; r_cur = r_next;
        }
    }
    // ---- final: K best picks in priority order ----
    __syncthreads();
    const unsigned int n_picks = n_picks_run;
    if (tid == 0) {
        a.stats[f * 4 + 0] = n_lm; a.stats[f * 4 + 1] = n;
        a.stats[f * 4 + 2] = rounds; a.stats[f * 4 + 3] = n_picks;
    }
#ifdef VO_NMS_TIMING
    t_2 = clock64();
#endif
    unsigned int P2 = 1;
    while (P2 < (unsigned)max(K, 1)) P2 <<= 1;
    const unsigned int n_out = min(n_picks, (unsigned)K);
    unsigned long long fk = 0ull;
    unsigned int fi = 0xFFFFFFFFu;
    // a few more picks than K (the usual case): sort them all and keep the first K.  Otherwise select the K-th
    // best first (picks staged behind the sort area when they fit, so the 12 passes run at shared-memory latency).
    unsigned int PA = P2;
    while (PA < n_picks) PA <<= 1;
    const bool sort_all = PA <= 4096u && (size_t)PA * 12 <= a.smem_bytes;
    if (sort_all) P2 = PA;
    unsigned long long* sk = reinterpret_cast<unsigned long long*>(smem_raw);   // the band state is not needed any more
    unsigned int* si = reinterpret_cast<unsigned int*>(sk + P2);
    const unsigned long long* qk = pk;
    const unsigned int* qi = pi;
    if (!sort_all && n_picks > (unsigned)K) {
        if (((size_t)P2 + n_picks) * 12 + 16 <= a.smem_bytes) {
            unsigned long long* tk2 = reinterpret_cast<unsigned long long*>(smem_raw + (((size_t)P2 * 12 + 15) & ~(size_t)15));
            unsigned int* ti2 = reinterpret_cast<unsigned int*>(tk2 + n_picks);
            for (unsigned int j = tid; j < n_picks; j += N_THREADS) { tk2[j] = pk[j]; ti2[j] = pi[j]; }
            qk = tk2; qi = ti2;
            __syncthreads();
        }
        block_select_kth(qk, qi, n_picks, (unsigned)K, hist, s_misc, &fk, &fi);
    }
    __syncthreads();
    for (unsigned int j = tid; j < P2; j += N_THREADS) { sk[j] = 0ull; si[j] = 0xFFFFFFFFu; }
    __syncthreads();
    // Every pick is kept and the scores have bin ranks (the usual case): the order is "bin rank, then priority inside the
    // bin", so a counting sort by bin plus a rank among the handful of picks that share a bin replaces the 66 stages of
    // the bitonic network (25 of the kernel's ~175 us per frame).  A bin shared by many picks (plateaus of equal scores)
    // falls back to the network.
    unsigned long long* tk = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)P2 * 12);
    unsigned int* ti = reinterpret_cast<unsigned int*>(tk + P2);
    unsigned int* hb = ti + P2;                               // [BINS] picks per bin -> first slot of the bin
    unsigned int* hc = hb + NMS_BINS;                         // [BINS] scatter cursors
    bool bin_sort = (sort_all || n_picks <= (unsigned)K) && n_picks > 0 && P2 >= 2u && n_picks <= P2 &&
                    (size_t)P2 * 24 + (size_t)NMS_BINS * 8 + 16 <= a.smem_bytes;
    unsigned int b_max = 0u;                                  // bins: linear in the scores' high words between the picks' extremes
    int b_shift = 0;
    auto pick_bin = [&](unsigned long long k) -> unsigned int { return (b_max - (unsigned int)(k >> 32)) >> b_shift; };
    if (bin_sort) {
        if (tid == 0) { s_misc[2] = 0u; s_cnt[6] = 0xFFFFFFFFu; s_cnt[7] = 0u; }
        for (int j = tid; j < NMS_BINS; j += N_THREADS) hb[j] = 0u;
        __syncthreads();
        {
            unsigned int h_lo = 0xFFFFFFFFu, h_hi = 0u;
            for (unsigned int j = tid; j < n_picks; j += N_THREADS) {
                const unsigned int h = (unsigned int)(qk[j] >> 32);
                h_lo = min(h_lo, h); h_hi = max(h_hi, h);
            }
            h_lo = __reduce_min_sync(0xFFFFFFFFu, h_lo); h_hi = __reduce_max_sync(0xFFFFFFFFu, h_hi);
            if (lane == 0) { atomicMin(&s_cnt[6], h_lo); atomicMax(&s_cnt[7], h_hi); }
        }
        __syncthreads();
        b_max = s_cnt[7];
        while (((b_max - s_cnt[6]) >> b_shift) >= (unsigned)NMS_BINS) b_shift++;
        for (unsigned int j = tid; j < n_picks; j += N_THREADS) atomicAdd(&hb[pick_bin(qk[j])], 1u);
        __syncthreads();
        {   // exclusive scan over the 2048 bins, two per thread
            const unsigned int c0 = hb[2 * tid], c1 = hb[2 * tid + 1];
            if (max(c0, c1) > 48u) s_misc[2] = 1u;            // crowded bin: its ranking loop would be the slow part
            unsigned int sa = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int va = __shfl_up_sync(0xFFFFFFFFu, sa, o);
                if (lane >= o) sa += va;
            }
            if (lane == 31) s_ws[warp] = sa;
            __syncthreads();
            unsigned int oa = 0;
            for (int w = 0; w < warp; w++) oa += s_ws[w];
            const unsigned int ea = oa + sa - (c0 + c1);
            hb[2 * tid] = ea; hb[2 * tid + 1] = ea + c0;
            hc[2 * tid] = ea; hc[2 * tid + 1] = ea + c0;
        }
        __syncthreads();
        bin_sort = s_misc[2] == 0u;
    }
    if (bin_sort) {
        for (unsigned int j = tid; j < n_picks; j += N_THREADS) {
            const unsigned long long k = qk[j];
            const unsigned int slot = atomicAdd(&hc[pick_bin(k)], 1u);
            tk[slot] = k; ti[slot] = qi[j];
        }
        __syncthreads();
        for (unsigned int j = tid; j < n_picks; j += N_THREADS) {
            const unsigned long long k = tk[j];
            const unsigned int i = ti[j], b = pick_bin(k);
            const unsigned int lo = hb[b], hi = b + 1u < (unsigned int)NMS_BINS ? hb[b + 1u] : n_picks;
            unsigned int c = lo;
            for (unsigned int m = lo; m < hi; m++) c += prio_gt(tk[m], ti[m], k, i) ? 1u : 0u;
            sk[c] = k; si[c] = i;
        }
        __syncthreads();
    } else
    if (sort_all || n_picks <= (unsigned)K) {                 // every pick is kept: plain copy
        for (unsigned int j = tid; j < n_picks; j += N_THREADS) { sk[j] = qk[j]; si[j] = qi[j]; }
    } else {
        for (unsigned int j0 = 0; j0 < n_picks; j0 += N_THREADS) {   // one atomic per warp for the slots
            const unsigned int j = j0 + tid;
            unsigned long long k = 0ull;
            unsigned int i = 0u;
            bool keep = false;
            if (j < n_picks) { k = qk[j]; i = qi[j]; keep = prio_ge(k, i, fk, fi); }
            const unsigned int mk = __ballot_sync(0xFFFFFFFFu, keep);
            unsigned int base_s = 0;
            if (lane == 0 && mk) base_s = atomicAdd(&s_cnt[5], (unsigned int)__popc(mk));
            base_s = __shfl_sync(0xFFFFFFFFu, base_s, 0);
            const unsigned int sl = base_s + __popc(mk & ((1u << lane) - 1u));
            if (keep && sl < P2) { sk[sl] = k; si[sl] = i; }
        }
    }
    __syncthreads();
#ifdef VO_NMS_TIMING
    const long long t_f1 = clock64();
#endif
    // bitonic sort, highest priority first.  Element e = tid + 1024 u lives in registers; compare-exchanges with a
    // partner less than 32 elements away (45 of the 66 stages for 2048 elements) are warp shuffles, the others go
    // through shared memory.  P2 <= 4096 here: at most four elements per thread.
    if (bin_sort) {
        // already in order
    } else if (P2 <= 4u * N_THREADS) {
        unsigned long long kr[4];
        unsigned int ir[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned int e = tid + u * N_THREADS;
            kr[u] = e < P2 ? sk[e] : 0ull; ir[u] = e < P2 ? si[e] : 0xFFFFFFFFu;
        }
        const int n_u = (int)((P2 + N_THREADS - 1) / N_THREADS);     // elements per thread in use
        auto stage_shfl = [&](unsigned int size, unsigned int stride) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (u >= n_u) break;
                const unsigned int e = tid + u * N_THREADS;
                const unsigned long long ko = __shfl_xor_sync(0xFFFFFFFFu, kr[u], stride);
                const unsigned int io = __shfl_xor_sync(0xFFFFFFFFu, ir[u], stride);
                const bool lower = (e & stride) == 0, desc = (e & size) == 0;
                // (j, l) = (lower, upper) element of the pair; exchange when their order contradicts the direction
                const bool j_first = lower ? prio_gt(kr[u], ir[u], ko, io) : prio_gt(ko, io, kr[u], ir[u]);
                if (desc ? !j_first : j_first) { kr[u] = ko; ir[u] = io; }
            }
        };
        for (unsigned int size = 2; size <= P2; size <<= 1) {
            unsigned int stride = size >> 1;
            if (stride >= 32u) {
                __syncthreads();
#pragma unroll
                for (int u = 0; u < 4; u++) { const unsigned int e = tid + u * N_THREADS; if (e < P2) { sk[e] = kr[u]; si[e] = ir[u]; } }
                __syncthreads();
                for (; stride >= 32u; stride >>= 1) {
                    for (unsigned int j = tid; j < P2; j += N_THREADS) {
                        const unsigned int l = j ^ stride;
                        if (l > j) {
                            const bool desc = ((j & size) == 0);
                            const unsigned long long kj = sk[j], kl = sk[l];
                            const unsigned int ij = si[j], il = si[l];
                            const bool j_first = prio_gt(kj, ij, kl, il);  // j has higher priority
                            if (desc ? !j_first : j_first) { sk[j] = kl; sk[l] = kj; si[j] = il; si[l] = ij; }
                        }
                    }
                    __syncthreads();
                }
#pragma unroll
                for (int u = 0; u < 4; u++) { const unsigned int e = tid + u * N_THREADS; if (e < P2) { kr[u] = sk[e]; ir[u] = si[e]; } }
            }
            for (; stride > 0; stride >>= 1) stage_shfl(size, stride);
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; u++) { const unsigned int e = tid + u * N_THREADS; if (e < P2) { sk[e] = kr[u]; si[e] = ir[u]; } }
        __syncthreads();
    } else {
        for (unsigned int size = 2; size <= P2; size <<= 1) {
            for (unsigned int stride = size >> 1; stride > 0; stride >>= 1) {
                for (unsigned int j = tid; j < P2; j += N_THREADS) {
                    const unsigned int l = j ^ stride;
                    if (l > j) {
                        const bool desc = ((j & size) == 0);
                        const unsigned long long kj = sk[j], kl = sk[l];
                        const unsigned int ij = si[j], il = si[l];
                        const bool j_first = prio_gt(kj, ij, kl, il);  // j has higher priority
                        if (desc ? !j_first : j_first) {
                            if (!(kj == kl && ij == il)) { sk[j] = kl; sk[l] = kj; si[j] = il; si[l] = ij; }
                        }
                    }
                }
                __syncthreads();
            }
        }
    }
#ifdef VO_NMS_TIMING
    const long long t_f2 = clock64();
#endif
    // harris.py:150 with a negative slice start zeroes nothing: the same pixel is returned for
    // every remaining iteration.  Find the first such pick (if any).
    if (tid == 0) s_misc[0] = 0xFFFFFFFFu;
    __syncthreads();
    for (unsigned int j = tid; j < n_out; j += N_THREADS) {
        const unsigned int p = si[j];
        const int py = (int)(p / (unsigned)W), px = (int)(p - (unsigned)py * W);
        if (py < r || px < r) atomicMin(&s_misc[0], j);
    }
    __syncthreads();
    const unsigned int stuck = s_misc[0];
    int* out = a.kp_xy + (size_t)f * K * 2;
    for (unsigned int j = tid; j < (unsigned)K; j += N_THREADS) {
        unsigned int src = j;
        if (stuck != 0xFFFFFFFFu && j > stuck) src = stuck;
        int x = 0, y = 0;
        if (src < n_out) {
            const unsigned int p = si[src];
            y = (int)(p / (unsigned)W); x = (int)(p - (unsigned)y * W);
        }
        out[2 * j] = x; out[2 * j + 1] = y;
    }
    if (a.flag != nullptr) {
        // a speculative start is good when K picks were found above its threshold; the state for the next call is the rank
        // the K-th pick has among this frame's local maxima, with a margin (K = "start from the safe threshold")
        const bool fell_short = a.pass == 0 && a.flag[f] == 2u && n_picks < (unsigned)K;
        if (tid == 0) s_misc[1] = 0u;
        __syncthreads();
        if (!fell_short && a.spec_rank != nullptr && n_out >= (unsigned)K && K > 0) {
            const unsigned long long kk = sk[K - 1];
            const unsigned int ki = si[K - 1];
            const unsigned long long* lmk = a.lm_key + (size_t)f * a.lm_cap;
            const unsigned int* lmi = a.lm_idx + (size_t)f * a.lm_cap;
            unsigned int c = 0;
            for (unsigned int j = tid; j < n_lm; j += N_THREADS) c += prio_ge(lmk[j], lmi[j], kk, ki) ? 1u : 0u;
            c = __reduce_add_sync(0xFFFFFFFFu, c);
            if (lane == 0 && c) atomicAdd(&s_misc[1], c);
        }
        __syncthreads();
        if (tid == 0) {
            a.flag[f] = fell_short ? 1u : 0u;
            if (!fell_short && a.spec_rank != nullptr) {
                const unsigned int rk = s_misc[1];
                a.spec_rank[f] = (n_out >= (unsigned)K && rk > 0u) ? min((unsigned)K, rk + rk / 8u + 8u) : (unsigned)K;
            }
        }
    }
#ifdef VO_NMS_TIMING
    __syncthreads();
    if (tid == 0) {   // probe build only: cycles of {sort, bands, final}, bands processed
        a.stats[f * 4 + 0] = (unsigned int)(t_1 - t_0); a.stats[f * 4 + 1] = (unsigned int)(t_f1 - t_2);
        a.stats[f * 4 + 2] = (unsigned int)(t_f2 - t_f1); a.stats[f * 4 + 3] = (unsigned int)(clock64() - t_f2);
        (void)n_bands; (void)t_1; (void)t_2; (void)acc_i; (void)acc_b;
    }
#endif
}

// harris.py:160-194: raw (2r+1)^2 patches around each keypoint from a zero-padded image.
__global__ void harris_descriptors_kernel(const uint8_t* __restrict__ img, size_t pitch, size_t frame_stride,
                                          int H, int W, const int* __restrict__ kp_xy, int K, int r,
                                          uint8_t* __restrict__ desc) {
    const int f = blockIdx.y, k = blockIdx.x;
    const int d = 2 * r + 1;
    const int cx = kp_xy[((size_t)f * K + k) * 2], cy = kp_xy[((size_t)f * K + k) * 2 + 1];
    const uint8_t* src = img + (size_t)f * frame_stride;
    uint8_t* out = desc + ((size_t)f * K + k) * d * d;
    for (int i = threadIdx.x; i < d * d; i += blockDim.x) {
        const int dy = i / d - r, dx = i % d - r;
        const int y = cy + dy, x = cx + dx;
        uint8_t v = 0;
        if (y >= 0 && y < H && x >= 0 && x < W) v = src[(size_t)y * pitch + x];
        out[i] = v;
    }
}

__global__ void kp_to_points_kernel(const int* __restrict__ kp, size_t n2, float* __restrict__ pts) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) pts[i] = (float)kp[i];
}

}  // namespace

int vo_launch_kp_to_points(vo_ctx* ctx, const int* d_kp_xy, size_t n, float* d_pts, cudaStream_t stream) {
    if (n == 0) return VO_OK;
    const size_t n2 = n * 2;
    kp_to_points_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, stream>>>(d_kp_xy, n2, d_pts);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    return VO_OK;
}

// ---------------------------------------------------------------------------------------------
// Host-side launchers (called from the C ABI in abi.cu)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_tmapEncodeTiled)p;
    }
    return fn;
}

int vo_launch_harris_response(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                              size_t frame_stride, int patch_size, double kappa, double* d_resp,
                              cudaStream_t stream) {
    VO_REQUIRE(patch_size >= 1 && (patch_size & 1) && patch_size / 2 <= PR_MAX,
               "harris: patch_size must be odd and <= %d (got %d)", 2 * PR_MAX + 1, patch_size);
    const int pr = patch_size / 2, pad = pr + 1;
    VO_REQUIRE(H >= 2 * pad + 1 && W >= 2 * pad + 1, "harris: image %dx%d too small for patch_size %d", W, H, patch_size);
    VO_REQUIRE(n_frames >= 1 && pitch >= (size_t)W, "harris: bad n_frames/pitch");
    // fast path: 9x9 patch, 16-byte aligned base / pitch / frame stride (what TMA requires)
    const bool tma_ok = patch_size == 9 && ((uintptr_t)d_img % 16 == 0) && pitch % 16 == 0 &&
                        (frame_stride % 16 == 0 || n_frames == 1) && n_frames <= 65535 && tmap_encoder() != nullptr &&
                        !ctx->env_harris_no_tma;
    if (tma_ok) {
        CUtensorMap tmap;
        const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_frames};
        const cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)(n_frames == 1 ? pitch * H : frame_stride)};
        const cuuint32_t box[3] = {FT_IN_W, FT_IN_H, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = tmap_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)d_img, gdim, gstr, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) {
            if (vo_ctx_once(ctx, VO_ATTR_HARRIS_FAST)) {
#define VO_HR_ATTR(WC) VO_CUDA(cudaFuncSetAttribute(harris_response_fast<WC>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
                VO_HR_ATTR(0) VO_HR_ATTR(1226) VO_HR_ATTR(1241) VO_HR_ATTR(1920) VO_HR_ATTR(4096)
#undef VO_HR_ATTR
            }
            const int tiles_x = vo_div_up(W + FT_XSHIFT, FT_W), tiles_y = vo_div_up(H, FT_H);
            const long long n_tiles = (long long)tiles_x * tiles_y * n_frames;
            VO_REQUIRE(n_tiles < (1ll << 31), "harris: too many tiles");
            const int grid = (int)((n_tiles < 2ll * ctx->sm_count) ? n_tiles : 2ll * ctx->sm_count);
            // width-specialised instances (store offsets as immediates): KITTI (1241 in BASELINE.json, 1226 for the
            // reference's sequence 05), full HD, 4K DCI (BASELINE's stress config); any other width takes the generic one
#define VO_HR_LAUNCH(WC) harris_response_fast<WC><<<grid, FT_THREADS, FT_SMEM, stream>>>(tmap, H, W, kappa, d_resp, tiles_x, tiles_y, (int)n_tiles)
            if (W == 1241) VO_HR_LAUNCH(1241);
            else if (W == 1226) VO_HR_LAUNCH(1226);
            else if (W == 1920) VO_HR_LAUNCH(1920);
            else if (W == 4096) VO_HR_LAUNCH(4096);
            else VO_HR_LAUNCH(0);
#undef VO_HR_LAUNCH
            ctx->launches++;
            VO_CHECK_LAUNCH();
            return VO_OK;
        }
    }
    const int pw = RT_W + 2 * pr, ph = RT_H + 2 * pr;
    const size_t smem = (size_t)3 * ph * pw * 4 + (size_t)3 * ph * RT_W * 4 + (size_t)(RT_W + 2 * pad) * (RT_H + 2 * pad);
    if (vo_ctx_once(ctx, VO_ATTR_HARRIS_TILED))
        VO_CUDA(cudaFuncSetAttribute(harris_response_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    dim3 grid(vo_div_up(W, RT_W), vo_div_up(H, RT_H), n_frames);
    harris_response_tiled<<<grid, R_THREADS, smem, stream>>>(d_img, pitch, frame_stride, H, W, pr, kappa, d_resp);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    return VO_OK;
}

size_t vo_harris_lm_cap(int H, int W, int r) {
    return (size_t)vo_div_up(H, r + 1) * (size_t)vo_div_up(W, r + 1);
}

// scratch layout for NMS (per call, n_frames frames)
struct NmsCarve {
    size_t npx, lm_cap, bm_words, smem_bands, total;
    bool bm_smem;
    size_t o_lmk, o_lmi, o_cnt, o_sup, o_mem, o_ea, o_eb, o_eh, o_pk, o_pi, o_stats, o_tk, o_ti, o_ctr, o_flag, o_cm;
    int cbw, cbh;
};
static void nms_carve(int n_frames, int H, int W, int radius, int num_keypoints, NmsCarve* c) {
    const size_t npx = (size_t)H * W, F = n_frames;
    const size_t lm_cap = vo_harris_lm_cap(H, W, radius);
    const size_t bm_words = npx / 32 + 2;
    unsigned int P2 = 1;
    while (P2 < (unsigned)num_keypoints) P2 <<= 1;
    // shared memory of the band kernel: bin tables + new-pick list (+ the two bitmaps when they fit); the final
    // sort reuses the same storage
    const size_t smem_tables = NMS_TABLES_BYTES;
    const bool bm_smem = smem_tables + 2 * bm_words * 4 <= 220 * 1024;
    size_t smem_bands = smem_tables + (bm_smem ? 2 * bm_words * 4 : 0);
    if ((size_t)P2 * 12 > smem_bands) smem_bands = (size_t)P2 * 12;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    c->npx = npx; c->lm_cap = lm_cap; c->bm_words = bm_words; c->smem_bands = smem_bands; c->bm_smem = bm_smem;
    c->o_lmk = carve(F * lm_cap * 8); c->o_lmi = carve(F * lm_cap * 4); c->o_cnt = carve(F * 4);
    c->o_sup = carve(F * bm_words * 4); c->o_mem = carve(bm_smem ? 0 : F * bm_words * 4);
    c->o_ea = carve(F * npx * 16); c->o_eb = carve(F * npx * 16); c->o_eh = carve(F * npx * 4);
    c->o_pk = carve(F * lm_cap * 8); c->o_pi = carve(F * lm_cap * 4);
    c->o_stats = carve(F * 16);
    c->o_tk = carve(F * 8); c->o_ti = carve(F * 4); c->o_ctr = carve(F * 16); c->o_flag = carve(F * 4);
    c->cbw = vo_div_up(W, 3); c->cbh = vo_div_up(H, 3);
    c->o_cm = carve(radius >= 5 ? F * (size_t)c->cbw * c->cbh * 4 : 0);     // coarse map of the 3x3 block maxima
    c->total = off;
}

// Reserve the NMS working memory ahead of time (resident objects call this at creation, so that their steps never
// allocate and stay capturable into a CUDA graph).
int vo_harris_nms_reserve(vo_ctx* ctx, int n_frames, int H, int W, int radius, int num_keypoints) {
    NmsCarve cv;
    nms_carve(n_frames, H, W, radius, num_keypoints, &cv);
    const int rc = vo_buf_reserve(&ctx->scratch[16], (size_t)n_frames * 4);   // carried speculation state
    if (rc) return rc;
    return vo_buf_reserve(&ctx->scratch[0], cv.total);
}

int vo_launch_harris_nms(vo_ctx* ctx, const double* d_resp, int n_frames, int H, int W, int radius,
                         int num_keypoints, int* d_kp_xy, unsigned int* d_stats_or_null, cudaStream_t stream) {
    VO_REQUIRE(radius >= 0 && radius <= 15, "harris nms: radius must be in [0, 15] (got %d)", radius);
    VO_REQUIRE(num_keypoints >= 1 && num_keypoints <= 16384, "harris nms: num_keypoints must be in [1, 16384]");
    VO_REQUIRE(H >= 2 * radius + 1 && W >= 2 * radius + 1, "harris nms: image smaller than the suppression box");
    VO_REQUIRE(H < 65536 && W < 65536, "harris nms: frame sides must be below 65536");
    NmsCarve cv;
    nms_carve(n_frames, H, W, radius, num_keypoints, &cv);
    const size_t npx = cv.npx, lm_cap = cv.lm_cap, F = n_frames, bm_words = cv.bm_words, smem_bands = cv.smem_bands, off = cv.total;
    const bool bm_smem = cv.bm_smem;
    const size_t o_lmk = cv.o_lmk, o_lmi = cv.o_lmi, o_cnt = cv.o_cnt, o_sup = cv.o_sup, o_mem = cv.o_mem, o_ea = cv.o_ea,
                 o_eb = cv.o_eb, o_eh = cv.o_eh, o_pk = cv.o_pk, o_pi = cv.o_pi, o_stats = cv.o_stats, o_tk = cv.o_tk,
                 o_ti = cv.o_ti, o_ctr = cv.o_ctr;
    int rc = vo_buf_reserve(&ctx->scratch[0], off, stream);
    if (rc) return rc;
    unsigned char* base = (unsigned char*)ctx->scratch[0].p;
    VO_CUDA(cudaMemsetAsync(base + o_cnt, 0, F * 4, stream));
    VO_CUDA(cudaMemsetAsync(base + o_sup, 0, F * bm_words * 4, stream));
    if (!bm_smem) VO_CUDA(cudaMemsetAsync(base + o_mem, 0, F * bm_words * 4, stream));

    if (vo_ctx_once(ctx, VO_ATTR_NMS)) {
        VO_CUDA(cudaFuncSetAttribute(harris_nms_bands<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        VO_CUDA(cudaFuncSetAttribute(harris_nms_bands<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    }
    {
        unsigned long long* lmk = (unsigned long long*)(base + o_lmk);
        unsigned int* lmi = (unsigned int*)(base + o_lmi);
        unsigned int* lmc = (unsigned int*)(base + o_cnt);
        const int bs = radius >= 5 ? 3 : (radius >= 3 ? 2 : 1);      // block size: the 3x3 blocks must fit the window
        const int out_w = (32 / bs - 2) * bs;
        // the coarse map is worth its 12 us only when the scan can use it: with a speculative threshold
        unsigned int* cm = (bs == 3 && !ctx->env_nms_no_spec && num_keypoints >= 64) ? (unsigned int*)(base + cv.o_cm) : nullptr;
        dim3 g1(vo_div_up(vo_div_up(W, out_w), L_THREADS / 32), vo_div_up(H, L_ROWS), n_frames);
        if (radius == 5) harris_localmax<5, 3><<<g1, L_THREADS, 0, stream>>>(d_resp, H, W, radius, (unsigned)lm_cap, lmk, lmi, lmc, cm, cv.cbw, cv.cbh);
        else if (bs == 3) harris_localmax<0, 3><<<g1, L_THREADS, 0, stream>>>(d_resp, H, W, radius, (unsigned)lm_cap, lmk, lmi, lmc, cm, cv.cbw, cv.cbh);
        else if (bs == 2) harris_localmax<0, 2><<<g1, L_THREADS, 0, stream>>>(d_resp, H, W, radius, (unsigned)lm_cap, lmk, lmi, lmc, nullptr, 0, 0);
        else harris_localmax<0, 1><<<g1, L_THREADS, 0, stream>>>(d_resp, H, W, radius, (unsigned)lm_cap, lmk, lmi, lmc, nullptr, 0, 0);
    }
    ctx->launches++;
    VO_CHECK_LAUNCH();
    NmsArgs a;
    a.resp = d_resp; a.H = H; a.W = W; a.r = radius; a.K = num_keypoints; a.lm_cap = (unsigned)lm_cap;
    a.lm_key = (unsigned long long*)(base + o_lmk); a.lm_idx = (unsigned int*)(base + o_lmi);
    a.lm_count = (unsigned int*)(base + o_cnt);
    a.sup = (unsigned int*)(base + o_sup); a.mem = bm_smem ? nullptr : (unsigned int*)(base + o_mem);
    a.bm_words = (unsigned)bm_words;
    a.ent_a = (uint4*)(base + o_ea); a.ent_b = (uint4*)(base + o_eb); a.ent_h = (unsigned int*)(base + o_eh);
    a.pick_key = (unsigned long long*)(base + o_pk); a.pick_idx = (unsigned int*)(base + o_pi);
    a.kp_xy = d_kp_xy;
    a.stats = d_stats_or_null ? d_stats_or_null : (unsigned int*)(base + o_stats);
    a.thr_key = (unsigned long long*)(base + o_tk); a.thr_idx = (unsigned int*)(base + o_ti);
    a.counters = (unsigned int*)(base + o_ctr);
    a.bitmaps_in_smem = bm_smem ? 1 : 0;
    a.smem_bytes = (unsigned)smem_bands;
    // measured: up to ~2000 keypoints per frame half-size bands win (fewer band members per window, shorter rounds:
    // 1.932 -> 1.917 ms per step at K = 1000), at 10 000 keypoints (4096x2160) full bands do (1494 vs 1336 frames/s)
    a.band = (unsigned)((ctx->nms_band >= 32 && ctx->nms_band <= N_THREADS) ? ctx->nms_band : (num_keypoints <= 2048 ? NMS_BAND / 2 : NMS_BAND));
    // carried state of the speculative threshold: one rank per frame slot, reset when the shape of the calls changes
    a.flag = (unsigned int*)(base + cv.o_flag);
    a.spec_rank = nullptr;
    a.cmap = nullptr; a.cbw = cv.cbw; a.cbh = cv.cbh;
    if (!ctx->env_nms_no_spec && num_keypoints >= 64) {
        const unsigned long long key = ((unsigned long long)n_frames << 48) ^ ((unsigned long long)H << 34) ^ ((unsigned long long)W << 20) ^
                                       ((unsigned long long)num_keypoints << 5) ^ (unsigned long long)radius;
        const void* before = ctx->scratch[16].p;
        if ((rc = vo_buf_reserve(&ctx->scratch[16], F * 4, stream))) return rc;
        if (ctx->scratch[16].p != before || ctx->nms_state_key != key) {
            VO_CUDA(cudaMemsetAsync(ctx->scratch[16].p, 0, F * 4, stream));
            ctx->nms_state_key = key;
        }
        a.spec_rank = (unsigned int*)ctx->scratch[16].p;
        if (radius >= 5) a.cmap = (const unsigned int*)(base + cv.o_cm);
    }
    const dim3 g3(vo_div_up((int)npx, 256 * SCAN_PER_THREAD), n_frames);
    for (int pass = 0; pass < ((a.spec_rank) ? 2 : 1); pass++) {   // pass 1: the frames whose speculation fell short (usually none)
        a.pass = pass;
        harris_nms_select<<<n_frames, N_THREADS, NMS_SELECT_SMEM * 12, stream>>>(a);
        ctx->launches++;
        VO_CHECK_LAUNCH();
        if (pass == 0 && a.cmap != nullptr) harris_nms_scan_blocks<<<dim3(vo_div_up(cv.cbw, 32), vo_div_up(cv.cbh, 8 * SB_ROWS), n_frames), 256, 0, stream>>>(a);
        else if (pass == 0) harris_nms_scan<false><<<g3, 256, 0, stream>>>(a);
        else harris_nms_scan<true><<<dim3(16, n_frames), 256, 0, stream>>>(a);
        ctx->launches++;
        VO_CHECK_LAUNCH();
        if (radius == 5) harris_nms_bands<5><<<n_frames, N_THREADS, smem_bands, stream>>>(a);
        else harris_nms_bands<0><<<n_frames, N_THREADS, smem_bands, stream>>>(a);
        ctx->launches++;
        VO_CHECK_LAUNCH();
    }
    return VO_OK;
}

int vo_launch_harris_descriptors(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                                 size_t frame_stride, const int* d_kp_xy, int K, int r, uint8_t* d_desc,
                                 cudaStream_t stream) {
    VO_REQUIRE(r >= 0 && r <= 32 && K >= 1, "harris descriptors: bad radius / K");
    dim3 grid(K, n_frames);
    harris_descriptors_kernel<<<grid, 128, 0, stream>>>(d_img, pitch, frame_stride, H, W, d_kp_xy, K, r, d_desc);
    ctx->launches++;
    VO_CHECK_LAUNCH();
    return VO_OK;
}
