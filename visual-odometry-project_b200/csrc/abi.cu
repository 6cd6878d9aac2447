// extern "C" boundary of libvo_b200.so -- see include/vo_b200.h for the contract.
#include <stdarg.h>

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

int vo_buf_reserve(VoBuf* b, size_t bytes) {
    if (bytes <= b->cap) return VO_OK;
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

static inline cudaStream_t pick_stream(vo_ctx* ctx, void* stream) {
    return stream ? (cudaStream_t)stream : ctx->stream;
}

extern "C" {

int vo_abi_version(void) { return 1; }
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
    VO_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
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

}  // extern "C"
