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
int vo_harris_nms_reserve(vo_ctx* ctx, int n_frames, int H, int W, int radius, int num_keypoints);
int vo_launch_harris_descriptors(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                                 size_t frame_stride, const int* d_kp_xy, int K, int r, uint8_t* d_desc,
                                 cudaStream_t stream);

// klt.cu
int vo_klt_layout_host(int H, int W, int max_level, int win, int* n_levels, int* level_h, int* level_w,
                       size_t* level_pitch, size_t* level_offset, size_t* frame_bytes);
int vo_launch_klt_pyramid(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                          size_t frame_stride, int max_level, int win, uint8_t* d_pyr, cudaStream_t stream, cudaEvent_t level0_done = nullptr);
int vo_launch_klt_track(vo_ctx* ctx, const uint8_t* d_pyr_prev, const uint8_t* d_pyr_next, int n_frames, int H, int W,
                        int max_level, int win, int max_iters, double epsilon, double min_eig,
                        const float* d_prev_pts, int n_pts, float* d_next_pts, uint8_t* d_status, float* d_err,
                        cudaStream_t stream, const int* d_n_valid = nullptr);
int vo_launch_bgr2gray(vo_ctx* ctx, const uint8_t* d_bgr, int n_frames, int H, int W, size_t in_pitch, size_t in_frame_stride,
                       uint8_t* d_gray, size_t out_pitch, size_t out_frame_stride, cudaStream_t stream);
// p3p.cu
int vo_launch_p3p_score(vo_ctx* ctx, const double* d_landmarks, const double* d_keypoints, int n_frames, int N,
                        const double* K9, const int* d_sample_idx, int n_hyp, double threshold, int inclusive, double* d_models,
                        unsigned char* d_valid, int* d_counts, cudaStream_t stream);
int vo_launch_p3p_select(vo_ctx* ctx, const double* d_landmarks, const double* d_keypoints, int n_frames, int N,
                         const double* K9, const double* d_models, const unsigned char* d_valid, const int* d_counts,
                         int n_hyp, double threshold, int inclusive, const int* d_iters_for_count, int initial_iters, int start_n,
                         int start_best, int* d_best4, int* d_consumed, int* d_iters_out, unsigned char* d_inliers,
                         double* d_best_model, cudaStream_t stream);
// triangulation.cu
int vo_launch_triangulate(vo_ctx* ctx, const double* d_p1, const double* d_p2, int n, const double* d_proj1,
                          int proj1_per_point, const double* d_proj2, int proj2_group, int mode, double* d_out,
                          cudaStream_t stream);
// harris.cu: int32 (x, y) keypoints -> float32 points (the KLT input format)
int vo_launch_kp_to_points(vo_ctx* ctx, const int* d_kp_xy, size_t n, float* d_pts, cudaStream_t stream);
// match.cu
int vo_launch_match(vo_ctx* ctx, const uint8_t* d_q, const uint8_t* d_t, int n_frames, int Q, int T, int D, double ratio,
                    int* d_pairs, int* d_n_pairs, cudaStream_t stream);
// gftt.cu
int vo_gftt_reserve(vo_ctx* ctx, int n_frames, int H, int W, double min_distance);
int vo_launch_gftt_eig(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch, size_t frame_stride,
                       int block_size, float* d_eig, int* d_frame_max, cudaStream_t stream);
int vo_launch_gftt(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch, size_t frame_stride,
                   int max_corners, double quality, double min_distance, int block_size, float* d_eig, float* d_xy, int* d_n,
                   unsigned int* d_stats_or_null, cudaStream_t stream);
// bootstrap.cu
int vo_launch_bootstrap(vo_ctx* ctx, const double* d_p1, const double* d_p2, int n_seq, int N, const int* d_n_pts,
                        const double* K9, double threshold, double confidence, int max_iters, double* d_F, double* d_M,
                        double* d_landmarks, unsigned char* d_mask, unsigned char* d_f_mask, int* d_info, cudaStream_t stream);
