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
    VoBuf scratch[16];
    VoBuf pinned[4];
    unsigned long long launches = 0;  // kernels launched through this context
};

int vo_buf_reserve(VoBuf* b, size_t bytes);
int vo_pinned_reserve(VoBuf* b, size_t bytes);

static inline int vo_div_up(int a, int b) { return (a + b - 1) / b; }
