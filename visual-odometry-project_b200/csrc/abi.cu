// extern "C" boundary of libvo_b200.so -- see include/vo_b200.h for the contract.
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/vo_b200.h"
#include "common.cuh"
#include "launchers.cuh"

static thread_local char g_err[1024] = "";

void vo_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Growing a scratch buffer frees and reallocates (a device-wide synchronisation, illegal under stream capture):
// resident objects (vo_frontend / vo_pipeline) reserve their worst case at creation so that steps never get here.
int vo_buf_reserve(VoBuf* b, size_t bytes, cudaStream_t launch_stream) {
    if (bytes <= b->cap) return VO_OK;
    if (launch_stream) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(launch_stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
            vo_set_error("scratch growth to %zu bytes requested during stream capture: reserve it before capturing", bytes);
            return VO_ERR_CAPACITY;
        }
    }
    if (b->p) VO_CUDA(cudaFree(b->p));
    b->p = nullptr;
    b->cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    VO_CUDA(cudaMalloc(&b->p, want));
    b->cap = want;
    return VO_OK;
}

int vo_pinned_reserve(VoBuf* b, size_t bytes) {
    if (bytes <= b->cap) return VO_OK;
    if (b->p) VO_CUDA(cudaFreeHost(b->p));
    b->p = nullptr;
    b->cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    VO_CUDA(cudaMallocHost(&b->p, want));
    b->cap = want;
    return VO_OK;
}

// Measured FP64 ceiling for the roofline of the P3P kernels: 8 independent DFMA chains per thread, every SM full.
__global__ void __launch_bounds__(256)
dfma_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
        a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

static inline cudaStream_t pick_stream(vo_ctx* ctx, void* stream) {
    return stream ? (cudaStream_t)stream : ctx->stream;
}

extern "C" {

int vo_abi_version(void) { return 2; }
const char* vo_last_error(void) { return g_err; }

int vo_ctx_create(vo_ctx** out, int device) {
    if (!out) { vo_set_error("vo_ctx_create: out == NULL"); return VO_ERR_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        vo_set_error("vo_ctx_create: no CUDA device (%s); this library has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return VO_ERR_NO_DEVICE;
    }
    VO_REQUIRE(device >= 0 && device < n, "vo_ctx_create: device %d out of range [0,%d)", device, n);
    VO_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VO_CUDA(cudaGetDeviceProperties(&prop, device));
    VO_REQUIRE(prop.major == 10, "vo_ctx_create: built for sm_100a, device is sm_%d%d", prop.major, prop.minor);
    vo_ctx* c = new vo_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    {   // environment switches are read here, once; launch paths only look at the context
        const char* e;
        c->env_harris_no_tma = (e = getenv("VO_HARRIS_NO_TMA")) && e[0] == '1';
        c->env_klt_generic = (e = getenv("VO_KLT_GENERIC")) && e[0] == '1';
        c->env_frontend_serial = (e = getenv("VO_FRONTEND_SERIAL")) && e[0] == '1';
        c->nms_band = (e = getenv("VO_NMS_BAND")) ? atoi(e) : 0;
        c->env_nms_no_spec = (e = getenv("VO_NMS_NO_SPEC")) && e[0] == '1';
    }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        vo_set_error("vo_ctx_create: cudaStreamCreate failed");
        delete c;
        return VO_ERR_CUDA;
    }
    *out = c;
    return VO_OK;
}

void vo_ctx_destroy(vo_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& b : ctx->scratch) if (b.p) cudaFree(b.p);
    for (auto& b : ctx->pinned) if (b.p) cudaFreeHost(b.p);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

unsigned long long vo_ctx_launch_count(const vo_ctx* ctx) { return ctx ? ctx->launches : 0ull; }

int vo_ctx_synchronize(vo_ctx* ctx) {
    VO_REQUIRE(ctx, "null context");
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    return VO_OK;
}

void* vo_ctx_stream(vo_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int vo_copy_to_host(vo_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
    VO_REQUIRE(ctx && h_dst && d_src, "vo_copy_to_host: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    VO_CUDA(cudaDeviceSynchronize());
    VO_CUDA(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return VO_OK;
}

// FP64 fused multiply-add rate of this GPU in GFLOP/s (2 flop per DFMA), measured with CUDA events.
int vo_test_dfma_peak(vo_ctx* ctx, double* gflops) {
    VO_REQUIRE(ctx && gflops, "vo_test_dfma_peak: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
    int rc = vo_buf_reserve(&ctx->scratch[12], (size_t)blocks * threads * 8);
    if (rc) return rc;
    cudaEvent_t e0, e1;
    VO_CUDA(cudaEventCreate(&e0)); VO_CUDA(cudaEventCreate(&e1));
    dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((double*)ctx->scratch[12].p, iters);       // warm-up
    VO_CUDA(cudaEventRecord(e0, ctx->stream));
    for (int r = 0; r < 5; r++) dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((double*)ctx->scratch[12].p, iters);
    VO_CUDA(cudaEventRecord(e1, ctx->stream));
    VO_CUDA(cudaEventSynchronize(e1));
    ctx->launches += 6;
    float ms = 0.f;
    VO_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *gflops = 5.0 * blocks * threads * (double)iters * 8.0 * 2.0 / (ms * 1e-3) / 1e9;
    return VO_OK;
}

// Page-locked host memory for the frame uploads of the pipelined entry points.  write_combined != 0 asks for
// cudaHostAllocWriteCombined: the CPU only ever writes these buffers (frames decoded / copied into them), and
// write-combined pages are not snooped during the GPU's reads, which helps when several GPUs pull frames at once.
int vo_host_alloc(void** out, size_t bytes, int write_combined) {
    VO_REQUIRE(out && bytes > 0, "vo_host_alloc: bad argument");
    *out = nullptr;
    VO_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
    return VO_OK;
}
int vo_host_free(void* p) {
    if (p) VO_CUDA(cudaFreeHost(p));
    return VO_OK;
}

// ------------------------------------------------------------------------------------------
// Harris
// ------------------------------------------------------------------------------------------
int vo_harris_response_dev(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                           size_t frame_stride, int patch_size, double kappa, double* d_resp, void* stream) {
    VO_REQUIRE(ctx && d_img && d_resp, "vo_harris_response_dev: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    return vo_launch_harris_response(ctx, d_img, n_frames, H, W, pitch, frame_stride, patch_size, kappa, d_resp,
                                     pick_stream(ctx, stream));
}

int vo_harris_nms_dev(vo_ctx* ctx, const double* d_resp, int n_frames, int H, int W, int nms_radius,
                      int num_keypoints, int32_t* d_kp_xy, uint32_t* d_stats, void* stream) {
    VO_REQUIRE(ctx && d_resp && d_kp_xy, "vo_harris_nms_dev: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    return vo_launch_harris_nms(ctx, d_resp, n_frames, H, W, nms_radius, num_keypoints, d_kp_xy, d_stats,
                                pick_stream(ctx, stream));
}

int vo_harris_detect_dev(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                         size_t frame_stride, int patch_size, double kappa, int nms_radius, int num_keypoints,
                         double* d_resp, int32_t* d_kp_xy, void* stream) {
    int rc = vo_harris_response_dev(ctx, d_img, n_frames, H, W, pitch, frame_stride, patch_size, kappa, d_resp, stream);
    if (rc) return rc;
    return vo_harris_nms_dev(ctx, d_resp, n_frames, H, W, nms_radius, num_keypoints, d_kp_xy, nullptr, stream);
}

int vo_harris_descriptors_dev(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                              size_t frame_stride, const int32_t* d_kp_xy, int K, int desc_radius,
                              uint8_t* d_desc, void* stream) {
    VO_REQUIRE(ctx && d_img && d_kp_xy && d_desc, "vo_harris_descriptors_dev: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    return vo_launch_harris_descriptors(ctx, d_img, n_frames, H, W, pitch, frame_stride, d_kp_xy, K, desc_radius,
                                        d_desc, pick_stream(ctx, stream));
}

int vo_harris_detect_host(vo_ctx* ctx, const uint8_t* h_img, int n_frames, int H, int W, int patch_size,
                          double kappa, int nms_radius, int num_keypoints, int desc_radius, double* h_resp,
                          int32_t* h_kp_xy, uint8_t* h_desc) {
    VO_REQUIRE(ctx && h_img && h_kp_xy, "vo_harris_detect_host: null argument");
    VO_REQUIRE(n_frames >= 1 && H > 0 && W > 0, "vo_harris_detect_host: bad shape");
    VO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t pitch = ((size_t)W + 15) & ~(size_t)15;  // 16-byte rows: TMA-legal layout in HBM
    const size_t fstride = pitch * H;
    const size_t npx = (size_t)H * W;
    const int d = 2 * desc_radius + 1;
    int rc;
    if ((rc = vo_buf_reserve(&ctx->scratch[1], fstride * n_frames))) return rc;
    if ((rc = vo_buf_reserve(&ctx->scratch[2], npx * n_frames * sizeof(double)))) return rc;
    if ((rc = vo_buf_reserve(&ctx->scratch[3], (size_t)n_frames * num_keypoints * 2 * sizeof(int32_t)))) return rc;
    uint8_t* d_img = (uint8_t*)ctx->scratch[1].p;
    double* d_resp = (double*)ctx->scratch[2].p;
    int32_t* d_kp = (int32_t*)ctx->scratch[3].p;
    VO_CUDA(cudaMemcpy2DAsync(d_img, pitch, h_img, W, W, (size_t)H * n_frames, cudaMemcpyHostToDevice, s));
    if ((rc = vo_harris_detect_dev(ctx, d_img, n_frames, H, W, pitch, fstride, patch_size, kappa, nms_radius,
                                   num_keypoints, d_resp, d_kp, s))) return rc;
    VO_CUDA(cudaMemcpyAsync(h_kp_xy, d_kp, (size_t)n_frames * num_keypoints * 2 * sizeof(int32_t),
                            cudaMemcpyDeviceToHost, s));
    if (h_resp) VO_CUDA(cudaMemcpyAsync(h_resp, d_resp, npx * n_frames * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (h_desc) {
        VO_REQUIRE(desc_radius >= 0, "vo_harris_detect_host: bad descriptor radius");
        const size_t db = (size_t)n_frames * num_keypoints * d * d;
        if ((rc = vo_buf_reserve(&ctx->scratch[4], db))) return rc;
        if ((rc = vo_harris_descriptors_dev(ctx, d_img, n_frames, H, W, pitch, fstride, d_kp, num_keypoints,
                                            desc_radius, (uint8_t*)ctx->scratch[4].p, s))) return rc;
        VO_CUDA(cudaMemcpyAsync(h_desc, ctx->scratch[4].p, db, cudaMemcpyDeviceToHost, s));
    }
    VO_CUDA(cudaStreamSynchronize(s));
    return VO_OK;
}

int vo_harris_descriptors_host(vo_ctx* ctx, const uint8_t* h_img, int H, int W, const int32_t* h_kp_xy, int K,
                               int desc_radius, uint8_t* h_desc) {
    VO_REQUIRE(ctx && h_img && h_kp_xy && h_desc && H > 0 && W > 0 && K >= 1, "vo_harris_descriptors_host: bad argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t pitch = ((size_t)W + 15) & ~(size_t)15;
    const int d = 2 * desc_radius + 1;
    const size_t db = (size_t)K * d * d;
    int rc;
    if ((rc = vo_buf_reserve(&ctx->scratch[1], pitch * H))) return rc;
    if ((rc = vo_buf_reserve(&ctx->scratch[3], (size_t)K * 2 * sizeof(int32_t)))) return rc;
    if ((rc = vo_buf_reserve(&ctx->scratch[4], db))) return rc;
    VO_CUDA(cudaMemcpy2DAsync(ctx->scratch[1].p, pitch, h_img, W, W, H, cudaMemcpyHostToDevice, s));
    VO_CUDA(cudaMemcpyAsync(ctx->scratch[3].p, h_kp_xy, (size_t)K * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    if ((rc = vo_harris_descriptors_dev(ctx, (uint8_t*)ctx->scratch[1].p, 1, H, W, pitch, pitch * H,
                                        (int32_t*)ctx->scratch[3].p, K, desc_radius, (uint8_t*)ctx->scratch[4].p, s))) return rc;
    VO_CUDA(cudaMemcpyAsync(h_desc, ctx->scratch[4].p, db, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaStreamSynchronize(s));
    return VO_OK;
}

int vo_match_descriptors_dev(vo_ctx* ctx, const uint8_t* d_desc1, const uint8_t* d_desc2, int n_frames, int Q, int T,
                             int D, double ratio, int32_t* d_pairs, int32_t* d_n_pairs, void* stream) {
    VO_REQUIRE(ctx && d_desc1 && d_desc2 && d_pairs && d_n_pairs, "vo_match_descriptors_dev: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    return vo_launch_match(ctx, d_desc1, d_desc2, n_frames, Q, T, D, ratio, d_pairs, d_n_pairs, pick_stream(ctx, stream));
}

int vo_match_descriptors_host(vo_ctx* ctx, const uint8_t* h_desc1, const uint8_t* h_desc2, int n_frames, int Q, int T,
                              int D, double ratio, int32_t* h_pairs, int32_t* h_n_pairs) {
    VO_REQUIRE(ctx && h_desc1 && h_desc2 && h_pairs && h_n_pairs, "vo_match_descriptors_host: null argument");
    VO_REQUIRE(n_frames >= 1 && Q >= 1 && T >= 1 && D >= 1, "vo_match_descriptors_host: bad sizes");
    VO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t F = n_frames;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_a = carve(F * Q * D), o_b = carve(F * T * D), o_p = carve(F * Q * 8), o_n = carve(F * 4);
    int rc = vo_buf_reserve(&ctx->scratch[11], off);
    if (rc) return rc;
    unsigned char* b = (unsigned char*)ctx->scratch[11].p;
    VO_CUDA(cudaMemcpyAsync(b + o_a, h_desc1, F * Q * D, cudaMemcpyHostToDevice, s));
    VO_CUDA(cudaMemcpyAsync(b + o_b, h_desc2, F * T * D, cudaMemcpyHostToDevice, s));
    if ((rc = vo_launch_match(ctx, b + o_a, b + o_b, n_frames, Q, T, D, ratio, (int*)(b + o_p), (int*)(b + o_n), s))) return rc;
    VO_CUDA(cudaMemcpyAsync(h_pairs, b + o_p, F * Q * 8, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaMemcpyAsync(h_n_pairs, b + o_n, F * 4, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaStreamSynchronize(s));
    return VO_OK;
}

// ------------------------------------------------------------------------------------------
// Shi-Tomasi corners (cv2.goodFeaturesToTrack)
// ------------------------------------------------------------------------------------------
int vo_gftt_dev(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch, size_t frame_stride, int max_corners,
                double quality_level, double min_distance, int block_size, float* d_eig, float* d_xy, int32_t* d_n, uint32_t* d_stats,
                void* stream) {
    VO_REQUIRE(ctx && d_img && d_eig && d_xy && d_n, "vo_gftt_dev: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    return vo_launch_gftt(ctx, d_img, n_frames, H, W, pitch, frame_stride, max_corners, quality_level, min_distance, block_size, d_eig,
                          d_xy, d_n, d_stats, pick_stream(ctx, stream));
}

int vo_gftt_host(vo_ctx* ctx, const uint8_t* h_img, int n_frames, int H, int W, int max_corners, double quality_level,
                 double min_distance, int block_size, float* h_eig, float* h_xy, int32_t* h_n, uint32_t* h_stats) {
    VO_REQUIRE(ctx && h_img && h_xy && h_n, "vo_gftt_host: null argument");
    VO_REQUIRE(n_frames >= 1 && H >= 3 && W >= 3 && max_corners >= 1, "vo_gftt_host: bad shape");
    VO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t pitch = ((size_t)W + 15) & ~(size_t)15, fstride = pitch * H, npx = (size_t)H * W, F = n_frames;
    int rc;
    if ((rc = vo_buf_reserve(&ctx->scratch[1], fstride * F))) return rc;
    if ((rc = vo_buf_reserve(&ctx->scratch[2], npx * F * sizeof(float)))) return rc;
    if ((rc = vo_buf_reserve(&ctx->scratch[3], F * max_corners * 8 + F * 4 + F * 16 + 1024))) return rc;
    uint8_t* d_img = (uint8_t*)ctx->scratch[1].p;
    float* d_eig = (float*)ctx->scratch[2].p;
    float* d_xy = (float*)ctx->scratch[3].p;
    int32_t* d_n = (int32_t*)((unsigned char*)ctx->scratch[3].p + ((F * max_corners * 8 + 255) & ~(size_t)255));
    uint32_t* d_st = (uint32_t*)(d_n + ((F + 63) & ~(size_t)63));
    VO_CUDA(cudaMemcpy2DAsync(d_img, pitch, h_img, W, W, (size_t)H * F, cudaMemcpyHostToDevice, s));
    if ((rc = vo_launch_gftt(ctx, d_img, n_frames, H, W, pitch, fstride, max_corners, quality_level, min_distance, block_size, d_eig, d_xy,
                             d_n, d_st, s))) return rc;
    VO_CUDA(cudaMemcpyAsync(h_xy, d_xy, F * max_corners * 8, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaMemcpyAsync(h_n, d_n, F * 4, cudaMemcpyDeviceToHost, s));
    if (h_eig) VO_CUDA(cudaMemcpyAsync(h_eig, d_eig, npx * F * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (h_stats) VO_CUDA(cudaMemcpyAsync(h_stats, d_st, F * 16, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaStreamSynchronize(s));
    return VO_OK;
}

// ------------------------------------------------------------------------------------------
// KLT
// ------------------------------------------------------------------------------------------
int vo_klt_pyramid_layout(int H, int W, int max_level, int win, int* n_levels, int* level_h, int* level_w,
                          size_t* level_pitch, size_t* level_offset, size_t* frame_bytes) {
    VO_REQUIRE(n_levels && frame_bytes, "vo_klt_pyramid_layout: null argument");
    return vo_klt_layout_host(H, W, max_level, win, n_levels, level_h, level_w, level_pitch, level_offset, frame_bytes);
}

int vo_klt_build_pyramid_dev(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                             size_t frame_stride, int max_level, int win, uint8_t* d_pyr, void* stream) {
    VO_REQUIRE(ctx && d_img && d_pyr, "vo_klt_build_pyramid_dev: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    return vo_launch_klt_pyramid(ctx, d_img, n_frames, H, W, pitch, frame_stride, max_level, win, d_pyr,
                                 pick_stream(ctx, stream));
}

int vo_klt_track_dev(vo_ctx* ctx, const uint8_t* d_pyr_prev, const uint8_t* d_pyr_next, int n_frames, int H,
                     int W, int max_level, int win, int max_iters, double epsilon, double min_eig_threshold,
                     const float* d_prev_pts, int n_pts, float* d_next_pts, uint8_t* d_status, float* d_err,
                     void* stream) {
    VO_REQUIRE(ctx && d_pyr_prev && d_pyr_next && (n_pts == 0 || (d_prev_pts && d_next_pts && d_status && d_err)),
               "vo_klt_track_dev: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    return vo_launch_klt_track(ctx, d_pyr_prev, d_pyr_next, n_frames, H, W, max_level, win, max_iters, epsilon,
                               min_eig_threshold, d_prev_pts, n_pts, d_next_pts, d_status, d_err,
                               pick_stream(ctx, stream));
}

// 64-bit content hash of a host image batch (identifies "the same frame as last call's next")
static unsigned long long host_hash(const uint8_t* p, size_t n) {
    unsigned long long h0 = 0x9E3779B97F4A7C15ull, h1 = 0xC2B2AE3D27D4EB4Full, h2 = 0x165667B19E3779F9ull, h3 = 0x27D4EB2F165667C5ull;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        unsigned long long w[4];
        memcpy(w, p + i, 32);
        h0 = (h0 ^ w[0]) * 0x100000001B3ull; h1 = (h1 ^ w[1]) * 0x100000001B3ull;
        h2 = (h2 ^ w[2]) * 0x100000001B3ull; h3 = (h3 ^ w[3]) * 0x100000001B3ull;
        h0 ^= h0 >> 29; h1 ^= h1 >> 31; h2 ^= h2 >> 27; h3 ^= h3 >> 33;
    }
    for (; i < n; i++) h0 = (h0 ^ p[i]) * 0x100000001B3ull;
    return (h0 ^ (h1 << 1) ^ (h2 << 2) ^ (h3 << 3)) + n;
}

// upload one image batch (1 or 3 channels, tightly packed) into the level-0 slots of pyramid `dst` and build the levels
static int klt_stage_pyramid(vo_ctx* ctx, const uint8_t* h_img, int channels, int n_frames, int H, int W, int max_level, int win,
                             size_t pitch0, size_t fb, uint8_t* dst, cudaStream_t s) {
    int rc;
    if (channels == 1) {
        for (int f = 0; f < n_frames; f++)
            VO_CUDA(cudaMemcpy2DAsync(dst + f * fb, pitch0, h_img + (size_t)f * H * W, W, W, H, cudaMemcpyHostToDevice, s));
    } else {
        const size_t bytes = (size_t)n_frames * H * W * 3;
        if ((rc = vo_buf_reserve(&ctx->scratch[14], bytes + 16))) return rc;
        VO_CUDA(cudaMemcpyAsync(ctx->scratch[14].p, h_img, bytes, cudaMemcpyHostToDevice, s));
        if ((rc = vo_launch_bgr2gray(ctx, (const uint8_t*)ctx->scratch[14].p, n_frames, H, W, (size_t)W * 3, (size_t)H * W * 3, dst,
                                     pitch0, fb, s))) return rc;
    }
    return vo_launch_klt_pyramid(ctx, dst, n_frames, H, W, pitch0, fb, max_level, win, dst, s);
}

static int klt_track_host_impl(vo_ctx* ctx, const uint8_t* h_prev, const uint8_t* h_next, int channels, int n_frames, int H, int W,
                               int max_level, int win, int max_iters, double epsilon, double min_eig_threshold,
                               const float* h_prev_pts, int n_pts, float* h_next_pts, uint8_t* h_status, float* h_err) {
    VO_REQUIRE(ctx && h_prev && h_next, "vo_klt_track_host: null argument");
    VO_REQUIRE(channels == 1 || channels == 3, "vo_klt_track_host: 1 (gray) or 3 (BGR) channels");
    VO_REQUIRE(n_pts == 0 || (h_prev_pts && h_next_pts && h_status && h_err), "vo_klt_track_host: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    int n_levels = 0, lh[8], lw[8];
    size_t lp[8], lo[8], fb = 0;
    int rc = vo_klt_layout_host(H, W, max_level, win, &n_levels, lh, lw, lp, lo, &fb);
    if (rc) return rc;
    if (n_pts == 0) return VO_OK;
    const size_t pts_b = (size_t)n_frames * n_pts * 2 * sizeof(float);
    const size_t img_bytes = (size_t)n_frames * H * W * channels;
    auto& kc = ctx->klt_cache;
    const bool same_shape = kc.valid && kc.H == H && kc.W == W && kc.channels == channels && kc.n_frames == n_frames &&
                            kc.max_level == max_level && kc.win == win && kc.bytes == img_bytes;
    const unsigned long long hp = host_hash(h_prev, img_bytes), hn = host_hash(h_next, img_bytes);
    const bool need_a = fb * n_frames > ctx->scratch[5].cap, need_b = fb * n_frames > ctx->scratch[6].cap;
    if ((rc = vo_buf_reserve(&ctx->scratch[5], fb * n_frames))) return rc;
    if ((rc = vo_buf_reserve(&ctx->scratch[6], fb * n_frames))) return rc;
    if ((rc = vo_buf_reserve(&ctx->scratch[7], 2 * pts_b + (size_t)n_frames * n_pts * 8 + 1024))) return rc;
    const bool hit = same_shape && !need_a && !need_b && kc.hash == hp;
    const int prev_buf = hit ? kc.which : 0;
    uint8_t* pa = (uint8_t*)ctx->scratch[5 + prev_buf].p;
    uint8_t* pb = (uint8_t*)ctx->scratch[5 + (1 - prev_buf)].p;
    float* d_prev = (float*)ctx->scratch[7].p;
    float* d_next = d_prev + (size_t)n_frames * n_pts * 2;
    float* d_err = d_next + (size_t)n_frames * n_pts * 2;
    uint8_t* d_st = (uint8_t*)(d_err + (size_t)n_frames * n_pts);
    kc.valid = false;
    if (hit) kc.hits++;
    else if ((rc = klt_stage_pyramid(ctx, h_prev, channels, n_frames, H, W, max_level, win, lp[0], fb, pa, s))) return rc;
    if ((rc = klt_stage_pyramid(ctx, h_next, channels, n_frames, H, W, max_level, win, lp[0], fb, pb, s))) return rc;
    VO_CUDA(cudaMemcpyAsync(d_prev, h_prev_pts, pts_b, cudaMemcpyHostToDevice, s));
    if ((rc = vo_launch_klt_track(ctx, pa, pb, n_frames, H, W, max_level, win, max_iters, epsilon, min_eig_threshold,
                                  d_prev, n_pts, d_next, d_st, d_err, s))) return rc;
    VO_CUDA(cudaMemcpyAsync(h_next_pts, d_next, pts_b, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaMemcpyAsync(h_err, d_err, (size_t)n_frames * n_pts * sizeof(float), cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaMemcpyAsync(h_status, d_st, (size_t)n_frames * n_pts, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaStreamSynchronize(s));
    kc.hash = hn; kc.bytes = img_bytes; kc.H = H; kc.W = W; kc.channels = channels; kc.n_frames = n_frames;
    kc.max_level = max_level; kc.win = win; kc.which = 1 - prev_buf; kc.valid = true;
    return VO_OK;
}

int vo_klt_track_host(vo_ctx* ctx, const uint8_t* h_prev, const uint8_t* h_next, int n_frames, int H, int W,
                      int max_level, int win, int max_iters, double epsilon, double min_eig_threshold,
                      const float* h_prev_pts, int n_pts, float* h_next_pts, uint8_t* h_status, float* h_err) {
    return klt_track_host_impl(ctx, h_prev, h_next, 1, n_frames, H, W, max_level, win, max_iters, epsilon, min_eig_threshold,
                               h_prev_pts, n_pts, h_next_pts, h_status, h_err);
}

int vo_klt_track_bgr_host(vo_ctx* ctx, const uint8_t* h_prev_bgr, const uint8_t* h_next_bgr, int n_frames, int H, int W,
                          int max_level, int win, int max_iters, double epsilon, double min_eig_threshold,
                          const float* h_prev_pts, int n_pts, float* h_next_pts, uint8_t* h_status, float* h_err) {
    return klt_track_host_impl(ctx, h_prev_bgr, h_next_bgr, 3, n_frames, H, W, max_level, win, max_iters, epsilon, min_eig_threshold,
                               h_prev_pts, n_pts, h_next_pts, h_status, h_err);
}

unsigned long long vo_klt_cache_hits(const vo_ctx* ctx) { return ctx ? ctx->klt_cache.hits : 0ull; }

int vo_bgr2gray_dev(vo_ctx* ctx, const uint8_t* d_bgr, int n_frames, int H, int W, size_t in_pitch, size_t in_frame_stride,
                    uint8_t* d_gray, size_t out_pitch, size_t out_frame_stride, void* stream) {
    VO_REQUIRE(ctx && d_bgr && d_gray, "vo_bgr2gray_dev: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    return vo_launch_bgr2gray(ctx, d_bgr, n_frames, H, W, in_pitch, in_frame_stride, d_gray, out_pitch, out_frame_stride,
                              pick_stream(ctx, stream));
}

int vo_bgr2gray_host(vo_ctx* ctx, const uint8_t* h_bgr, int n_frames, int H, int W, uint8_t* h_gray) {
    VO_REQUIRE(ctx && h_bgr && h_gray && n_frames >= 1 && H >= 1 && W >= 1, "vo_bgr2gray_host: bad argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t pitch = ((size_t)W + 15) & ~(size_t)15, in_b = (size_t)n_frames * H * W * 3;
    int rc;
    if ((rc = vo_buf_reserve(&ctx->scratch[14], in_b + 16))) return rc;
    if ((rc = vo_buf_reserve(&ctx->scratch[15], pitch * H * n_frames))) return rc;
    VO_CUDA(cudaMemcpyAsync(ctx->scratch[14].p, h_bgr, in_b, cudaMemcpyHostToDevice, s));
    if ((rc = vo_launch_bgr2gray(ctx, (const uint8_t*)ctx->scratch[14].p, n_frames, H, W, (size_t)W * 3, (size_t)H * W * 3,
                                 (uint8_t*)ctx->scratch[15].p, pitch, pitch * H, s))) return rc;
    VO_CUDA(cudaMemcpy2DAsync(h_gray, W, ctx->scratch[15].p, pitch, W, (size_t)H * n_frames, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaStreamSynchronize(s));
    return VO_OK;
}

// ------------------------------------------------------------------------------------------
// P3P + RANSAC
// ------------------------------------------------------------------------------------------
int vo_p3p_ransac_score_dev(vo_ctx* ctx, const double* d_landmarks, const double* d_keypoints, int n_frames,
                            int N, const double* K9, const int32_t* d_sample_idx, int n_hyp, double threshold,
                            int inclusive, double* d_models, uint8_t* d_valid, int32_t* d_counts, void* stream) {
    VO_REQUIRE(ctx && d_landmarks && d_keypoints && K9 && d_sample_idx && d_models && d_valid && d_counts,
               "vo_p3p_ransac_score_dev: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    return vo_launch_p3p_score(ctx, d_landmarks, d_keypoints, n_frames, N, K9, d_sample_idx, n_hyp, threshold, inclusive,
                               d_models, d_valid, d_counts, pick_stream(ctx, stream));
}

int vo_p3p_ransac_select_dev(vo_ctx* ctx, const double* d_landmarks, const double* d_keypoints, int n_frames,
                             int N, const double* K9, const double* d_models, const uint8_t* d_valid,
                             const int32_t* d_counts, int n_hyp, double threshold, int inclusive, const int32_t* d_iters_for_count,
                             int initial_iters, int start_n, int start_best, int32_t* d_best4, int32_t* d_consumed,
                             int32_t* d_iters_out, uint8_t* d_inliers, double* d_best_model, void* stream) {
    VO_REQUIRE(ctx && d_landmarks && d_keypoints && K9 && d_models && d_valid && d_counts && d_iters_for_count &&
               d_best4 && d_consumed && d_iters_out && d_inliers && d_best_model,
               "vo_p3p_ransac_select_dev: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    return vo_launch_p3p_select(ctx, d_landmarks, d_keypoints, n_frames, N, K9, d_models, d_valid, d_counts, n_hyp,
                                threshold, inclusive, d_iters_for_count, initial_iters, start_n, start_best, d_best4, d_consumed,
                                d_iters_out, d_inliers, d_best_model, pick_stream(ctx, stream));
}

int vo_p3p_ransac_host(vo_ctx* ctx, const double* h_landmarks, const double* h_keypoints, int n_frames, int N,
                       const double* K9, const int32_t* h_sample_idx, int n_hyp, double threshold, int inclusive,
                       const int32_t* h_iters_for_count, int initial_iters, int start_n, int start_best,
                       int32_t* h_best4, int32_t* h_consumed, int32_t* h_iters_out, uint8_t* h_inliers,
                       double* h_best_model, int32_t* h_counts, uint8_t* h_valid, double* h_models) {
    VO_REQUIRE(ctx && h_landmarks && h_keypoints && K9 && h_sample_idx && h_iters_for_count && h_best4 &&
               h_consumed && h_iters_out && h_inliers && h_best_model, "vo_p3p_ransac_host: null argument");
    VO_REQUIRE(n_frames >= 1 && N >= 4 && n_hyp >= 1, "vo_p3p_ransac_host: bad sizes");
    VO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t F = n_frames;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_l = carve(F * N * 24), o_k = carve(F * N * 16), o_s = carve(F * n_hyp * 16);
    const size_t o_m = carve(F * n_hyp * 96), o_v = carve(F * n_hyp), o_c = carve(F * n_hyp * 4);
    const size_t o_t = carve((size_t)(N + 1) * 4), o_b = carve(F * 16), o_con = carve(F * 4), o_it = carve(F * 4);
    const size_t o_in = carve(F * N), o_bm = carve(F * 96);
    int rc = vo_buf_reserve(&ctx->scratch[8], off);
    if (rc) return rc;
    unsigned char* b = (unsigned char*)ctx->scratch[8].p;
    VO_CUDA(cudaMemcpyAsync(b + o_l, h_landmarks, F * N * 24, cudaMemcpyHostToDevice, s));
    VO_CUDA(cudaMemcpyAsync(b + o_k, h_keypoints, F * N * 16, cudaMemcpyHostToDevice, s));
    VO_CUDA(cudaMemcpyAsync(b + o_s, h_sample_idx, F * n_hyp * 16, cudaMemcpyHostToDevice, s));
    VO_CUDA(cudaMemcpyAsync(b + o_t, h_iters_for_count, (size_t)(N + 1) * 4, cudaMemcpyHostToDevice, s));
    if ((rc = vo_launch_p3p_score(ctx, (double*)(b + o_l), (double*)(b + o_k), n_frames, N, K9, (int*)(b + o_s), n_hyp,
                                  threshold, inclusive, (double*)(b + o_m), b + o_v, (int*)(b + o_c), s))) return rc;
    if ((rc = vo_launch_p3p_select(ctx, (double*)(b + o_l), (double*)(b + o_k), n_frames, N, K9, (double*)(b + o_m),
                                   b + o_v, (int*)(b + o_c), n_hyp, threshold, inclusive, (int*)(b + o_t), initial_iters, start_n,
                                   start_best, (int*)(b + o_b), (int*)(b + o_con), (int*)(b + o_it), b + o_in,
                                   (double*)(b + o_bm), s))) return rc;
    VO_CUDA(cudaMemcpyAsync(h_best4, b + o_b, F * 16, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaMemcpyAsync(h_consumed, b + o_con, F * 4, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaMemcpyAsync(h_iters_out, b + o_it, F * 4, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaMemcpyAsync(h_inliers, b + o_in, F * N, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaMemcpyAsync(h_best_model, b + o_bm, F * 96, cudaMemcpyDeviceToHost, s));
    if (h_counts) VO_CUDA(cudaMemcpyAsync(h_counts, b + o_c, F * n_hyp * 4, cudaMemcpyDeviceToHost, s));
    if (h_valid) VO_CUDA(cudaMemcpyAsync(h_valid, b + o_v, F * n_hyp, cudaMemcpyDeviceToHost, s));
    if (h_models) VO_CUDA(cudaMemcpyAsync(h_models, b + o_m, F * n_hyp * 96, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaStreamSynchronize(s));
    return VO_OK;
}

// ------------------------------------------------------------------------------------------
// Triangulation
// ------------------------------------------------------------------------------------------
int vo_triangulate_dev(vo_ctx* ctx, const double* d_p1, const double* d_p2, int n, const double* d_proj1,
                       int proj1_per_point, const double* d_proj2, int mode, double* d_out, void* stream) {
    VO_REQUIRE(ctx && (n == 0 || (d_p1 && d_p2 && d_proj1 && d_proj2 && d_out)), "vo_triangulate_dev: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    return vo_launch_triangulate(ctx, d_p1, d_p2, n, d_proj1, proj1_per_point, d_proj2, 0, mode, d_out,
                                 pick_stream(ctx, stream));
}

int vo_triangulate_host(vo_ctx* ctx, const double* h_p1, const double* h_p2, int n, const double* h_proj1,
                        int proj1_per_point, const double* h_proj2, int mode, double* h_out) {
    VO_REQUIRE(ctx && n >= 0, "vo_triangulate_host: bad arguments");
    if (n == 0) return VO_OK;
    VO_REQUIRE(h_p1 && h_p2 && h_proj1 && h_proj2 && h_out, "vo_triangulate_host: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t np1 = proj1_per_point ? (size_t)n : 1;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_a = carve((size_t)n * 16), o_b = carve((size_t)n * 16), o_c1 = carve(np1 * 96), o_c2 = carve(96);
    const size_t o_o = carve((size_t)n * 24);
    int rc = vo_buf_reserve(&ctx->scratch[9], off);
    if (rc) return rc;
    unsigned char* b = (unsigned char*)ctx->scratch[9].p;
    VO_CUDA(cudaMemcpyAsync(b + o_a, h_p1, (size_t)n * 16, cudaMemcpyHostToDevice, s));
    VO_CUDA(cudaMemcpyAsync(b + o_b, h_p2, (size_t)n * 16, cudaMemcpyHostToDevice, s));
    VO_CUDA(cudaMemcpyAsync(b + o_c1, h_proj1, np1 * 96, cudaMemcpyHostToDevice, s));
    VO_CUDA(cudaMemcpyAsync(b + o_c2, h_proj2, 96, cudaMemcpyHostToDevice, s));
    if ((rc = vo_launch_triangulate(ctx, (double*)(b + o_a), (double*)(b + o_b), n, (double*)(b + o_c1),
                                    proj1_per_point, (double*)(b + o_c2), 0, mode, (double*)(b + o_o), s))) return rc;
    VO_CUDA(cudaMemcpyAsync(h_out, b + o_o, (size_t)n * 24, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaStreamSynchronize(s));
    return VO_OK;
}

// ------------------------------------------------------------------------------------------
// Two-view bootstrap
// ------------------------------------------------------------------------------------------
int vo_bootstrap_dev(vo_ctx* ctx, const double* d_p1, const double* d_p2, int n_seq, int N, const int32_t* d_n_pts,
                     const double* K9, double threshold, double confidence, int max_iters, double* d_F, double* d_M,
                     double* d_landmarks, uint8_t* d_mask, uint8_t* d_f_mask, int32_t* d_info, void* stream) {
    VO_REQUIRE(ctx && d_p1 && d_p2 && K9 && d_F && d_M && d_landmarks && d_mask && d_info, "vo_bootstrap_dev: null argument");
    VO_CUDA(cudaSetDevice(ctx->device));
    return vo_launch_bootstrap(ctx, d_p1, d_p2, n_seq, N, d_n_pts, K9, threshold, confidence, max_iters, d_F, d_M,
                               d_landmarks, d_mask, d_f_mask, d_info, pick_stream(ctx, stream));
}

int vo_bootstrap_host(vo_ctx* ctx, const double* h_p1, const double* h_p2, int n_seq, int N, const int32_t* h_n_pts,
                      const double* K9, double threshold, double confidence, int max_iters, double* h_F, double* h_M,
                      double* h_landmarks, uint8_t* h_mask, uint8_t* h_f_mask, int32_t* h_info) {
    VO_REQUIRE(ctx && h_p1 && h_p2 && K9 && h_F && h_M && h_landmarks && h_mask && h_info, "vo_bootstrap_host: null argument");
    VO_REQUIRE(n_seq >= 1 && N >= 1, "vo_bootstrap_host: n_seq and N must be positive");
    VO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t S = (size_t)n_seq, rows = S * (size_t)N;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_a = carve(rows * 16), o_b = carve(rows * 16), o_n = carve(S * 4), o_F = carve(S * 72), o_M = carve(S * 96);
    const size_t o_l = carve(rows * 24), o_m = carve(rows), o_f = carve(rows), o_i = carve(S * 16);
    int rc = vo_buf_reserve(&ctx->scratch[9], off);
    if (rc) return rc;
    unsigned char* b = (unsigned char*)ctx->scratch[9].p;
    VO_CUDA(cudaMemcpyAsync(b + o_a, h_p1, rows * 16, cudaMemcpyHostToDevice, s));
    VO_CUDA(cudaMemcpyAsync(b + o_b, h_p2, rows * 16, cudaMemcpyHostToDevice, s));
    if (h_n_pts) VO_CUDA(cudaMemcpyAsync(b + o_n, h_n_pts, S * 4, cudaMemcpyHostToDevice, s));
    if ((rc = vo_launch_bootstrap(ctx, (double*)(b + o_a), (double*)(b + o_b), n_seq, N, h_n_pts ? (int*)(b + o_n) : nullptr, K9,
                                  threshold, confidence, max_iters, (double*)(b + o_F), (double*)(b + o_M), (double*)(b + o_l),
                                  b + o_m, b + o_f, (int*)(b + o_i), s))) return rc;
    VO_CUDA(cudaMemcpyAsync(h_F, b + o_F, S * 72, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaMemcpyAsync(h_M, b + o_M, S * 96, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaMemcpyAsync(h_landmarks, b + o_l, rows * 24, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaMemcpyAsync(h_mask, b + o_m, rows, cudaMemcpyDeviceToHost, s));
    if (h_f_mask) VO_CUDA(cudaMemcpyAsync(h_f_mask, b + o_f, rows, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaMemcpyAsync(h_info, b + o_i, S * 16, cudaMemcpyDeviceToHost, s));
    VO_CUDA(cudaStreamSynchronize(s));
    return VO_OK;
}

// ------------------------------------------------------------------------------------------
// cv::RNG subsets for the host-driven OpenCV-style RANSAC (control flow only: no device work)
// ------------------------------------------------------------------------------------------
int vo_cv_rng_subsets_host(uint64_t* state, int n_points, int model_points, int count, int32_t* out) {
    VO_REQUIRE(state && out && count >= 0, "vo_cv_rng_subsets_host: null argument");
    VO_REQUIRE(model_points >= 1 && model_points <= 16 && n_points >= model_points, "vo_cv_rng_subsets_host: need 1 <= model_points <= 16 and n_points >= model_points");
    uint64_t st = *state;
    for (int c = 0; c < count; c++) {
        int32_t* id = out + (size_t)c * model_points;
        for (int i = 0; i < model_points; i++) {
            for (;;) {
                st = (uint64_t)(uint32_t)st * 4164903690ull + (st >> 32);       // cv::RNG::next
                const int32_t v = (int32_t)((uint32_t)st % (uint32_t)n_points);  // uniform(0, n)
                bool dup = false;
                for (int j = 0; j < i; j++) dup |= (id[j] == v);
                if (!dup) { id[i] = v; break; }
            }
        }
    }
    *state = st;
    return VO_OK;
}

}  // extern "C"
