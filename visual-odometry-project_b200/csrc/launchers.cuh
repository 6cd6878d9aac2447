// Host-side launch functions implemented next to their kernels.
#pragma once
#include "common.cuh"

// harris.cu
int vo_launch_harris_response(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                              size_t frame_stride, int patch_size, double kappa, double* d_resp,
                              cudaStream_t stream);
size_t vo_harris_lm_cap(int H, int W, int r);
int vo_launch_harris_nms(vo_ctx* ctx, const double* d_resp, int n_frames, int H, int W, int radius,
                         int num_keypoints, int* d_kp_xy, unsigned int* d_stats_or_null, cudaStream_t stream);
int vo_launch_harris_descriptors(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                                 size_t frame_stride, const int* d_kp_xy, int K, int r, uint8_t* d_desc,
                                 cudaStream_t stream);
