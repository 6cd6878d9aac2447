// Shared helpers for the B200 (sm_100a) visual-odometry front end.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define VO_OK 0
#define VO_ERR_CUDA 1
#define VO_ERR_ARG 2
#define VO_ERR_CAPACITY 3
#define VO_ERR_NO_DEVICE 4

void vo_set_error(const char* fmt, ...);

#define VO_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            vo_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,              \
                         cudaGetErrorString(e__));                                 \
            return VO_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)

#define VO_CHECK_LAUNCH() VO_CUDA(cudaGetLastError())

#define VO_REQUIRE(cond, ...)                                                      \
    do {                                                                           \
        if (!(cond)) {                                                             \
            vo_set_error(__VA_ARGS__);                                             \
            return VO_ERR_ARG;                                                     \
        }                                                                          \
    } while (0)

// Growable device scratch owned by a context.
struct VoBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct vo_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;  // context-owned stream used by the *_host entry points
    VoBuf scratch[17];               // [16]: per-frame state the NMS carries from one call to the next (harris.cu)
    VoBuf pinned[4];
    unsigned long long launches = 0;  // kernels launched through this context
    // One-time opt-ins (cudaFuncSetAttribute is per device): a bit per kernel family, set by vo_ctx_once().
    unsigned attr_done = 0;
    // Environment switches, read once in vo_ctx_create (never in a launch path).
    bool env_harris_no_tma = false;   // VO_HARRIS_NO_TMA=1: generic three-stage response kernel
    bool env_klt_generic = false;     // VO_KLT_GENERIC=1: runtime-window tracker for every window size
    bool env_frontend_serial = false; // VO_FRONTEND_SERIAL=1: no fork/join across streams
    int nms_band = 0;                 // VO_NMS_BAND (0 = built-in default)
    bool env_nms_no_spec = false;     // VO_NMS_NO_SPEC=1: never start the NMS from a speculative threshold
    unsigned long long nms_state_key = 0;   // shape the carried NMS state belongs to
    // vo_klt_track_*_host keeps the pyramid of the last `next` image: a tracker that is called with consecutive frame
    // pairs (klt.py:233-239) uploads and pyramids every frame once, not twice.
    struct { unsigned long long hash = 0; size_t bytes = 0; int H = 0, W = 0, channels = 0, n_frames = 0, max_level = 0, win = 0;
             int which = 0; bool valid = false; unsigned long long hits = 0; } klt_cache;
};

// A context is bound to ONE device and its launchers carve working memory from ctx-owned scratch: calls that
// share a vo_ctx must be ordered on one stream (or be externally serialised).  Use one vo_ctx per stream.
enum { VO_ATTR_HARRIS_FAST = 1, VO_ATTR_HARRIS_TILED = 2, VO_ATTR_NMS = 4, VO_ATTR_KLT = 8, VO_ATTR_MATCH = 16,
       VO_ATTR_GFTT = 32, VO_ATTR_PIPE = 64, VO_ATTR_BOOT = 128 };
static inline bool vo_ctx_once(vo_ctx* ctx, unsigned bit) {
    if (ctx->attr_done & bit) return false;
    ctx->attr_done |= bit;
    return true;
}

int vo_buf_reserve(VoBuf* b, size_t bytes, cudaStream_t launch_stream = nullptr);
int vo_pinned_reserve(VoBuf* b, size_t bytes);

static inline int vo_div_up(int a, int b) { return (a + b - 1) / b; }
