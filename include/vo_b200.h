/*
 * vo_b200.h -- C ABI of the B200-native visual-odometry front end (libvo_b200.so).
 *
 * Drop-in boundary for the data-parallel hot path of saegsali/visual-odometry-project.
 * Every entry point names the reference interface it replaces (file:line under
 * /root/reference).  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *  - `*_dev` functions take DEVICE pointers and a cudaStream_t (passed as void*; NULL = the
 *    context's own stream) and are asynchronous.  `*_host` functions take HOST pointers, stage
 *    through pinned memory, run on the context's stream and return after the result is on the
 *    host.
 *  - All functions return 0 on success, non-zero on failure; vo_last_error() describes it.
 *  - Images are uint8, row-major, `pitch` bytes between rows, `frame_stride` bytes between the
 *    frames of a batch.  Points are (x, y) = (column, row), as in the reference's (N, 2, 1) arrays.
 *  - There is no CPU fallback: without a CUDA device vo_ctx_create fails.
 *  - A vo_ctx belongs to one device and owns the working memory its launchers carve from: calls that share a
 *    context must be ordered on ONE stream (use one context per stream).  Resident objects (vo_frontend,
 *    vo_pipeline) reserve their working memory at creation; a `_dev` call that would have to grow it while its
 *    stream is being captured into a CUDA graph fails with a clear error instead of allocating.
 */
#ifndef VO_B200_H
#define VO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vo_ctx vo_ctx;

/* ---- context ------------------------------------------------------------------------------ */
int vo_ctx_create(vo_ctx** out, int device);
void vo_ctx_destroy(vo_ctx* ctx);
const char* vo_last_error(void);
int vo_abi_version(void);
/* number of kernels launched through this context so far (bench.py's gpu_launches) */
unsigned long long vo_ctx_launch_count(const vo_ctx* ctx);
int vo_ctx_synchronize(vo_ctx* ctx);
/* the context's stream as a cudaStream_t (for event timing on the launching stream) */
void* vo_ctx_stream(vo_ctx* ctx);
/* synchronous device -> host copy of a resident result buffer (after a device-wide sync) */
int vo_copy_to_host(vo_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);

/* page-locked host buffers for the *_host entry points (write_combined: CPU-write-only staging, not snooped by GPU reads) */
int vo_host_alloc(void** out, size_t bytes, int write_combined);
int vo_host_free(void* p);
/* measured FP64 fused-multiply-add rate of the device in GFLOP/s (the ceiling the P3P kernels are compared with) */
int vo_test_dfma_peak(vo_ctx* ctx, double* gflops);

/* ---- Harris: src/vo/features/harris.py ---------------------------------------------------- */
/* harris.py:102-137  float64 score map [n_frames][H][W], zero border of patch_size/2+1 pixels.  */
int vo_harris_response_dev(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                           size_t frame_stride, int patch_size, double kappa, double* d_resp, void* stream);
/* harris.py:139-152  greedy non-maximum suppression -> int32 [n_frames][num_keypoints][2] = (x, y),
 * in the reference's selection order.  d_stats (optional, may be NULL): uint32 [n_frames][4] =
 * {local maxima, entries above the threshold outside all boxes, rounds, picks}.                  */
int vo_harris_nms_dev(vo_ctx* ctx, const double* d_resp, int n_frames, int H, int W, int nms_radius,
                      int num_keypoints, int32_t* d_kp_xy, uint32_t* d_stats, void* stream);
/* harris.py:86-158  HarrisCornerDetector.extractKeypoints = response + NMS.                     */
int vo_harris_detect_dev(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                         size_t frame_stride, int patch_size, double kappa, int nms_radius, int num_keypoints,
                         double* d_resp, int32_t* d_kp_xy, void* stream);
/* harris.py:160-194  extractDescriptors: uint8 [n_frames][K][(2r+1)^2] patches, zero padded.    */
int vo_harris_descriptors_dev(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                              size_t frame_stride, const int32_t* d_kp_xy, int K, int desc_radius,
                              uint8_t* d_desc, void* stream);
/* Host-buffer variant of extractKeypoints (+ optional response / descriptors).  h_resp and h_desc
 * may be NULL.  h_img is tightly packed (pitch == W).                                           */
int vo_harris_detect_host(vo_ctx* ctx, const uint8_t* h_img, int n_frames, int H, int W, int patch_size,
                          double kappa, int nms_radius, int num_keypoints, int desc_radius, double* h_resp,
                          int32_t* h_kp_xy, uint8_t* h_desc);

/* harris.py:160-194 alone, host buffers: one tightly packed frame, int32 (x, y) keypoints [K][2].  */
int vo_harris_descriptors_host(vo_ctx* ctx, const uint8_t* h_img, int H, int W, const int32_t* h_kp_xy, int K,
                               int desc_radius, uint8_t* h_desc);

/* harris.py:196-264  matchDescriptor for 8-bit descriptors: cv2.BFMatcher().knnMatch(k=2) (L2), the ratio
 * test m.distance < ratio * n.distance (0.85 in the reference) and first-come uniqueness of the train index.
 * desc uint8 [n_frames][Q|T][D]; pairs int32 [n_frames][Q][2] = (query, train) in query order, the first
 * n_pairs[f] rows valid.                                                                            */
int vo_match_descriptors_dev(vo_ctx* ctx, const uint8_t* d_desc1, const uint8_t* d_desc2, int n_frames, int Q, int T,
                             int D, double ratio, int32_t* d_pairs, int32_t* d_n_pairs, void* stream);
int vo_match_descriptors_host(vo_ctx* ctx, const uint8_t* h_desc1, const uint8_t* h_desc2, int n_frames, int Q, int T,
                              int D, double ratio, int32_t* h_pairs, int32_t* h_n_pairs);

/* ---- Shi-Tomasi corners: src/vo/features/klt.py:24-26, 87-115 (cv2.goodFeaturesToTrack) -------------------------- */
/* cv2.goodFeaturesToTrack(img, maxCorners, qualityLevel, minDistance, blockSize=block_size) with the default 3x3 Sobel
 * aperture and no mask (the reference passes an all-255 mask).  OpenCV's float arithmetic is reproduced operation by
 * operation (see csrc/gftt.cu), so the corner list -- float32 [n_frames][max_corners][2] = (x, y) in OpenCV's order, the
 * first n[f] rows valid -- is cv2's.  d_eig receives cv2.cornerMinEigenVal's map, float32 [n_frames][H][W].
 * stats (optional) uint32 [n_frames][4] = {candidates, corners before the maxCorners cut, selection rounds, overflow}. */
int vo_gftt_dev(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch, size_t frame_stride, int max_corners,
                double quality_level, double min_distance, int block_size, float* d_eig, float* d_xy, int32_t* d_n, uint32_t* d_stats,
                void* stream);
int vo_gftt_host(vo_ctx* ctx, const uint8_t* h_img, int n_frames, int H, int W, int max_corners, double quality_level,
                 double min_distance, int block_size, float* h_eig, float* h_xy, int32_t* h_n, uint32_t* h_stats);

/* ---- KLT: src/vo/features/klt.py:233-239 (cv2.calcOpticalFlowPyrLK) ------------------------- */
/* Gaussian 5x5 pyramid (cv2.pyrDown, BORDER_REFLECT_101) of a batch of frames, frame-major: level l
 * of frame f lives at d_pyr + f * frame_bytes + level_offset[l] with row pitch level_pitch[l]
 * (arrays of up to 8 entries filled by vo_klt_pyramid_layout; the pyramid stops before a level that
 * is not larger than the window, as cv2.buildOpticalFlowPyramid does).  If d_img already is the
 * level-0 slot (d_img == d_pyr, pitch == level_pitch[0], frame_stride == frame_bytes) no copy is made. */
int vo_klt_pyramid_layout(int H, int W, int max_level, int win, int* n_levels, int* level_h, int* level_w,
                          size_t* level_pitch, size_t* level_offset, size_t* frame_bytes);
int vo_klt_build_pyramid_dev(vo_ctx* ctx, const uint8_t* d_img, int n_frames, int H, int W, size_t pitch,
                             size_t frame_stride, int max_level, int win, uint8_t* d_pyr, void* stream);
/* Pyramidal Lucas-Kanade, one warp per point.  Points float32 [n_frames][n_pts][2]; outputs
 * next points, status (uint8) and the L1 patch error (float32), as cv2 returns them.             */
int vo_klt_track_dev(vo_ctx* ctx, const uint8_t* d_pyr_prev, const uint8_t* d_pyr_next, int n_frames, int H,
                     int W, int max_level, int win, int max_iters, double epsilon, double min_eig_threshold,
                     const float* d_prev_pts, int n_pts, float* d_next_pts, uint8_t* d_status, float* d_err,
                     void* stream);
int vo_klt_track_host(vo_ctx* ctx, const uint8_t* h_prev, const uint8_t* h_next, int n_frames, int H, int W,
                      int max_level, int win, int max_iters, double epsilon, double min_eig_threshold,
                      const float* h_prev_pts, int n_pts, float* h_next_pts, uint8_t* h_status, float* h_err);

/* The same with BGR frames (uint8 [n_frames][H][W][3], as cv2.imread returns them): the conversion klt.py:57-62 does
 * with cv2.cvtColor(BGR2GRAY) runs on the device (OpenCV's 15-bit fixed point, bit-exact) straight into the pyramid.
 * Both host entry points keep the pyramid of the last `next` batch and reuse it when the following call's `prev` has
 * the same content (a tracker fed consecutive frame pairs uploads every frame once); vo_klt_cache_hits counts reuses. */
int vo_klt_track_bgr_host(vo_ctx* ctx, const uint8_t* h_prev_bgr, const uint8_t* h_next_bgr, int n_frames, int H, int W,
                          int max_level, int win, int max_iters, double epsilon, double min_eig_threshold,
                          const float* h_prev_pts, int n_pts, float* h_next_pts, uint8_t* h_status, float* h_err);
unsigned long long vo_klt_cache_hits(const vo_ctx* ctx);
/* klt.py:57-62, 84-85  cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) for uint8 images, bit-exact.                          */
int vo_bgr2gray_dev(vo_ctx* ctx, const uint8_t* d_bgr, int n_frames, int H, int W, size_t in_pitch, size_t in_frame_stride,
                    uint8_t* d_gray, size_t out_pitch, size_t out_frame_stride, void* stream);
int vo_bgr2gray_host(vo_ctx* ctx, const uint8_t* h_bgr, int n_frames, int H, int W, uint8_t* h_gray);

/* ---- P3P + RANSAC: src/vo/pose_estimation/p3p.py:51-108, src/vo/algorithms/ransac.py:69-129 --- */
/* For every hypothesis h (4 sample indices: 3 for P3P, the 4th disambiguates, as cv2.solvePnP with
 * SOLVEPNP_P3P does) solve the pose, count reprojection inliers (squared pixel error < threshold,
 * p3p.py:104-108 / ransac.py:104-106; with inclusive != 0 the rule of cv2.solvePnPRansac, which the reference's
 * use_opencv=True path runs (p3p.py:142-151): projections rounded to float32, squared distance in float32,
 * err <= (float)threshold with threshold = reprojectionError^2; the caller passes landmarks rounded to float32, as
 * solvePnPRansac converts them), and return per-hypothesis models, validity and counts.
 * landmarks float64 [n_frames][N][3], keypoints float64 [n_frames][N][2], K9 float64 [9] row-major (HOST pointer),
 * sample_idx int32 [n_frames][n_hyp][4].  models float64 [n_frames][n_hyp][12] = R (row-major) | t. */
int vo_p3p_ransac_score_dev(vo_ctx* ctx, const double* d_landmarks, const double* d_keypoints, int n_frames,
                            int N, const double* K9, const int32_t* d_sample_idx, int n_hyp, double threshold,
                            int inclusive, double* d_models, uint8_t* d_valid, int32_t* d_counts, void* stream);
/* ransac.py:90-121 sequential scan with the adaptive iteration count.  d_iters_for_count is the
 * host-computed table n_iterations(best_n_inliers), best_n_inliers = 0..N (ransac.py:58-67,115-120);
 * initial_iters / start_n / start_best carry the loop state in (ransac.py:56,81,84; start_best = -1
 * and start_n = 0 for a fresh call).  Outputs per frame: d_best4 int32 [4] = {best hypothesis index
 * or -1 if none beat start_best, best inlier count, n (iterations counted), exhausted flag: 1 if the
 * hypotheses ran out before n reached n_iterations}; d_consumed = hypotheses consumed (valid and
 * invalid, i.e. RNG draws); d_iters_out = n_iterations at exit; inlier mask uint8 [N]; model [12]. */
int vo_p3p_ransac_select_dev(vo_ctx* ctx, const double* d_landmarks, const double* d_keypoints, int n_frames,
                             int N, const double* K9, const double* d_models, const uint8_t* d_valid,
                             const int32_t* d_counts, int n_hyp, double threshold, int inclusive, const int32_t* d_iters_for_count,
                             int initial_iters, int start_n, int start_best, int32_t* d_best4, int32_t* d_consumed,
                             int32_t* d_iters_out, uint8_t* d_inliers, double* d_best_model, void* stream);
/* Host-buffer variant: score + select.  h_counts / h_valid / h_models may be NULL.  K9 and
 * h_iters_for_count (int32 [N+1]) are host pointers in every variant.                           */
int vo_p3p_ransac_host(vo_ctx* ctx, const double* h_landmarks, const double* h_keypoints, int n_frames, int N,
                       const double* K9, const int32_t* h_sample_idx, int n_hyp, double threshold, int inclusive,
                       const int32_t* h_iters_for_count, int initial_iters, int start_n, int start_best,
                       int32_t* h_best4, int32_t* h_consumed, int32_t* h_iters_out, uint8_t* h_inliers,
                       double* h_best_model, int32_t* h_counts, uint8_t* h_valid, double* h_models);

/* p3p.py:188-213  _nonlinear_refinement for n_frames independent problems: minimise the sum of squared reprojection
 * distances of the correspondences with mask != 0 (NULL = all) over the 6-dof pose, starting from pose_in (R row-major |
 * t, world -> camera).  The reference runs scipy.optimize.least_squares over the twist with a numeric Jacobian and stops
 * on ftol = 1e-8; this is a damped Gauss-Newton on SE(3) with the analytic Jacobian run to a 1e-11 step: same cost, its
 * minimum (never a higher cost than the reference's result; see DESIGN.md for the measured distance).  One CTA per
 * problem, block-wide reductions; iters (optional) = steps tried.                                                    */
int vo_refine_pose_dev(vo_ctx* ctx, const double* d_landmarks, const double* d_keypoints, const uint8_t* d_mask, int n_frames, int N,
                       const double* K9, const double* d_pose_in, double* d_pose_out, int32_t* d_iters, void* stream);
int vo_refine_pose_host(vo_ctx* ctx, const double* h_landmarks, const double* h_keypoints, const uint8_t* h_mask, int n_frames, int N,
                        const double* K9, const double* h_pose_in, double* h_pose_out, int32_t* h_iters);

/* ---- Triangulation: src/vo/landmarks/triangulation.py:352-389, 38-86 -------------------------- */
/* Linear (DLT) triangulation, one point per thread, one-sided Jacobi SVD in registers.
 * p1, p2 float64 [n][2]; proj1 float64 [n or 1][12] (row-major 3x4; per-point when
 * proj1_per_point != 0, as triangulate_candidates uses); proj2 float64 [12].
 * mode 0: 6x4 system [p1]x C1 ; [p2]x C2   (triangulation.py:379-387, use_opencv=False)
 * mode 1: 4x4 system x*P3-P1, y*P3-P2      (cv2.triangulatePoints, triangulation.py:59-74)
 * out float64 [n][3] (cartesian).                                                              */
int vo_triangulate_dev(vo_ctx* ctx, const double* d_p1, const double* d_p2, int n, const double* d_proj1,
                       int proj1_per_point, const double* d_proj2, int mode, double* d_out, void* stream);
int vo_triangulate_host(vo_ctx* ctx, const double* h_p1, const double* h_p2, int n, const double* h_proj1,
                        int proj1_per_point, const double* h_proj2, int mode, double* h_out);

/* Host helper of the drop-in P3PPoseEstimator's default path (p3p.py:142-151, cv2.solvePnPRansac): `count` subsets of
 * `model_points` distinct indices below n_points as RANSACPointSetRegistrator::getSubset draws them from cv::RNG
 * (state in / out; OpenCV seeds every run with 2^64 - 1).  Control flow only -- the models and inlier counts of the
 * subsets come from vo_p3p_ransac_*.  out int32 [count][model_points].                                           */
int vo_cv_rng_subsets_host(uint64_t* state, int n_points, int model_points, int count, int32_t* out);

/* ---- Two-view bootstrap: src/vo/landmarks/triangulation.py:88-350 (as src/main.py:185-222 configures it) ---- */
/* LandmarksTriangulator(use_ransac=True, use_opencv=True).triangulate_matches for n_seq independent frame pairs, one
 * CTA each: cv2.findFundamentalMat(FM_RANSAC, threshold, confidence) restated (float32 points, cv::RNG subsets of 7
 * with the collinearity check, 7-point solver, float32 epipolar errors, RANSACUpdateNumIters; triangulation.py:126-134),
 * E = K^T F K (:224-243), the four [R | t] candidates (:245-277), cheirality vote by DLT triangulation of the F inliers
 * and the landmarks of ALL matches with the winner (:279-350).
 * p1, p2 float64 [n_seq][N][2] (frame 1 / frame 2 pixels); n_pts int32 [n_seq] or NULL (= N); 15 <= n <= 8192 (below 15
 * points OpenCV switches to LMedS: info[0] = 0 is returned).  Outputs: F float64 [n_seq][9] (F33 = 1), M float64
 * [n_seq][12] (row-major 3x4, frame 1 -> frame 2), landmarks float64 [n_seq][N][3] (frame-1 coordinates), mask uint8
 * [n_seq][N] (F inlier and in front of both cameras = triangulate_matches' third result), f_mask uint8 [n_seq][N] or
 * NULL (F inliers = _find_fundamental_matrix_ransac's second result), info int32 [n_seq][4] = {model found, RANSAC
 * iterations, F inliers, cheirality-valid inliers}.                                                               */
int vo_bootstrap_dev(vo_ctx* ctx, const double* d_p1, const double* d_p2, int n_seq, int N, const int32_t* d_n_pts,
                     const double* K9, double threshold, double confidence, int max_iters, double* d_F, double* d_M,
                     double* d_landmarks, uint8_t* d_mask, uint8_t* d_f_mask, int32_t* d_info, void* stream);
int vo_bootstrap_host(vo_ctx* ctx, const double* h_p1, const double* h_p2, int n_seq, int N, const int32_t* h_n_pts,
                      const double* K9, double threshold, double confidence, int max_iters, double* h_F, double* h_M,
                      double* h_landmarks, uint8_t* h_mask, uint8_t* h_f_mask, int32_t* h_info);

/* ---- The front end as one resident object: src/main.py:248-287 (loop body) --------------------- */
/* Per-sequence state (pyramids of the previous / current frame, last keypoints) stays in HBM; one
 * call advances n_seq independent sequences by one frame: pyramid -> KLT (previous keypoints into
 * the new frame, klt.py:233-239) -> Harris on the new frame (harris.py:86-158) -> P3P-RANSAC on
 * the supplied 3D-2D correspondences (p3p.py:123-186, use_opencv=False) -> DLT triangulation of the
 * supplied tracks (triangulation.py:38-86).                                                       */
typedef struct vo_frontend vo_frontend;
typedef struct {
    int n_seq, H, W;
    int patch_size; double kappa; int nms_radius; int num_keypoints;          /* harris.py:16-34   */
    int klt_win, klt_max_level, klt_max_iters; double klt_epsilon, klt_min_eig; /* klt.py:29-33    */
    int n_corr, n_hyp; double p3p_threshold;                                   /* p3p.py:14-40      */
    int n_tri, tri_mode;                                                       /* triangulation.py  */
} vo_frontend_params;
typedef struct {
    double* d_resp; int32_t* d_kp_xy; float* d_tracked; uint8_t* d_status; float* d_err;
    int32_t* d_best4; uint8_t* d_inliers; double* d_pose; double* d_tri_out; int32_t* d_counts;
    uint8_t* d_cur_pyramid; size_t pyr_pitch0, pyr_frame_bytes;
} vo_frontend_outputs_t;
int vo_frontend_create(vo_ctx* ctx, const vo_frontend_params* params, vo_frontend** out);
void vo_frontend_destroy(vo_frontend* fe);
/* device pointers of the resident result buffers (valid until destroy) */
int vo_frontend_outputs(vo_frontend* fe, vo_frontend_outputs_t* out);
/* level-0 slots the next step will read: upload frames here to skip the copy kernel */
uint8_t* vo_frontend_next_frame_slot(vo_frontend* fe, size_t* pitch, size_t* frame_stride);
/* Device-resident step (asynchronous on `stream`).  d_frames uint8 [n_seq] frames with the given
 * pitch / frame_stride; P3P inputs as vo_p3p_ransac_*; triangulation inputs [n_seq * n_tri] points,
 * per-point start projections [n_seq * n_tri][12], one end projection per sequence [n_seq][12].  */
int vo_frontend_step_dev(vo_frontend* fe, const uint8_t* d_frames, size_t pitch, size_t frame_stride,
                         const double* d_landmarks, const double* d_kp2d, const double* K9,
                         const int32_t* d_sample_idx, const int32_t* d_iters_table, int initial_iters,
                         const double* d_tri_p1, const double* d_tri_p2, const double* d_tri_proj1,
                         const double* d_tri_proj2, void* stream);
/* Optional: start uploading the inputs of the NEXT step on the context's copy stream (double-buffered
 * staging) so that the transfer overlaps the step that is computing now.  Up to two uploads may be pending;
 * step_host consumes the oldest one (its h_* input pointers are then ignored and may be NULL).  Call order:
 * prefetch(step 0); then per step t: prefetch(step t+1); step_host(step t).                          */
int vo_frontend_prefetch_host(vo_frontend* fe, const uint8_t* h_frames, const double* h_landmarks, const double* h_kp2d,
                              const int32_t* h_sample_idx, const int32_t* h_iters_table, const double* h_tri_p1,
                              const double* h_tri_p2, const double* h_tri_proj1, const double* h_tri_proj2);
/* Host-buffer step: uploads the frames (tightly packed) and the P3P / triangulation inputs, runs
 * the step, downloads keypoints int32 [n_seq][K][2], tracked points / status / err of the previous
 * keypoints, best4 / inlier masks / poses [n_seq][12] and landmarks [n_seq * n_tri][3]; returns when
 * the results are on the host.  h_tracked, h_status, h_err, h_best4, h_inliers may be NULL.       */
int vo_frontend_step_host(vo_frontend* fe, const uint8_t* h_frames, const double* h_landmarks, const double* h_kp2d,
                          const double* K9, const int32_t* h_sample_idx, const int32_t* h_iters_table,
                          int initial_iters, const double* h_tri_p1, const double* h_tri_p2,
                          const double* h_tri_proj1, const double* h_tri_proj2, int32_t* h_kp_xy, float* h_tracked,
                          uint8_t* h_status, float* h_err, int32_t* h_best4, uint8_t* h_inliers, double* h_pose,
                          double* h_tri_out);
/* Pipelined form of vo_frontend_step_host (main.py:248-287 for a stream of frames): submit enqueues the step and
 * the download of its results into the given host buffers and returns at once; wait blocks until the oldest
 * submitted step's results are on the host.  Up to two steps may be in flight, so the GPU never idles between
 * steps and the download of step t runs under the compute of step t+1 (outputs alternate between two device
 * sets).  Call order: prefetch(0); per step t: prefetch(t+1); submit(t); wait() for step t-1.  The host buffers
 * of a submitted step must stay untouched until its wait returns.                                       */
int vo_frontend_submit_host(vo_frontend* fe, const uint8_t* h_frames, const double* h_landmarks, const double* h_kp2d,
                            const double* K9, const int32_t* h_sample_idx, const int32_t* h_iters_table,
                            int initial_iters, const double* h_tri_p1, const double* h_tri_p2,
                            const double* h_tri_proj1, const double* h_tri_proj2, int32_t* h_kp_xy, float* h_tracked,
                            uint8_t* h_status, float* h_err, int32_t* h_best4, uint8_t* h_inliers, double* h_pose,
                            double* h_tri_out);
int vo_frontend_wait_host(vo_frontend* fe);


/* ---- The chained, device-resident pipeline: src/main.py:248-287 with the reference's data flow --------------- */
/* One feature table per sequence stays in HBM (rows: keypoint, landmark, state, track start, start pose, candidate
 * flag -- the columns of src/vo/primitives/features.py).  One step advances every sequence by one frame:
 *   re-detection rule + append (klt.py:207-230) -> pyramidal LK (klt.py:233-239) -> status / error filter
 *   (klt.py:244-249) -> Matches regrouping (matches.py) -> P3P-RANSAC on the triangulated rows with the numpy
 *   PCG64(2023) sample stream, adaptive stop and carried state (ransac.py:69-129, p3p.py:123-186) -> pose refinement
 *   (p3p.py:188-213; damped Gauss-Newton to the minimum of the same cost) -> State.update_with_world_pose,
 *   reset_outliers, compute_candidates (state.py:39-178) -> triangulate_candidates (triangulation.py:38-86) ->
 *   update_with_world_landmarks + _check_landmarks (state.py:70-110).
 * Only the frames are uploaded; a step returns, per sequence, the pose and a few counters.                       */
typedef struct vo_pipeline vo_pipeline;
#define VO_DETECTOR_NONE 0     /* the host appends corners itself (vo_pipeline_write_table_host)                 */
#define VO_DETECTOR_HARRIS 1   /* HarrisCornerDetector.extractKeypoints (harris.py:86-158) on every new frame    */
#define VO_DETECTOR_GFTT 2     /* cv2.goodFeaturesToTrack (klt.py:24-26, 87-115), the reference's KLT-mode detector */
#define VO_PIPE_NCOUNTS 12
#define VO_PIPE_SUMMARY_DOUBLES 18   /* per sequence: 12 doubles pose + VO_PIPE_NCOUNTS int32 counters            */
/* counters: 0 rows after the step, 1 rows tracked, 2 rows kept by the status/error filter, 3 P3P population,
 * 4 inliers of the winning model, 5 candidates triangulated, 6 triangulated rows after the step,
 * 7 flags (1 re-detected, 2 no pose: fewer than 4 landmarks or no model, 4 table full on append, 8 sample cap hit),
 * 8 RANSAC n_iterations after the step, 9 samples drawn, 10 landmarks dropped behind a camera, 11 refinement steps */
typedef struct {
    int n_seq, H, W;
    int capacity;                                    /* table rows per sequence, multiple of 32                    */
    int klt_win, klt_max_level, klt_max_iters;       /* klt.py:29-33                                               */
    double klt_epsilon, klt_min_eig;
    float klt_error_threshold;                       /* klt.py:36  _error_threshold                                */
    double redetect_fraction;                        /* klt.py:211 (0.8)                                           */
    int detector, det_max_corners;                   /* VO_DETECTOR_*; corners per detection                       */
    int patch_size; double kappa; int nms_radius;    /* harris.py:16-34                                            */
    double gftt_quality, gftt_min_distance; int gftt_block_size;   /* klt.py:24-26                                 */
    double K[9], Kinv[9];                            /* intrinsics and their inverse AS THE HOST COMPUTES IT
                                                        (camera.py:92: np.linalg.inv in K's own dtype)             */
    double p3p_threshold; int p3p_inclusive;         /* p3p.py:20.  inclusive = 0: the reference's own RANSAC (ransac.py: numpy
                                                        PCG64 stream, err < threshold, state carried between frames);
                                                        1: cv2.solvePnPRansac restated (p3p.py:142-151: cv::RNG subsets, float32
                                                        points and errors, err <= threshold = reprojectionError^2,
                                                        RANSACUpdateNumIters; nothing carried between frames)        */
    double ransac_confidence, ransac_outlier_ratio;  /* p3p.py:22-23                                               */
    double ransac_log1mconf;                         /* log(1 - confidence) as the host evaluates it (0 = compute) */
    int ransac_max_iterations, ransac_initial_iterations; /* p3p.py:24; 0 = derive from the outlier ratio          */
    int refine;                                      /* p3p.py:25 nonlinear_refinement                             */
    double bearing_threshold;                        /* state.py:9                                                 */
    int tri_mode;                                    /* 1 = cv2.triangulatePoints' system (use_opencv=True), 0 = reference's */
} vo_pipeline_params;
int vo_pipeline_create(vo_ctx* ctx, const vo_pipeline_params* params, vo_pipeline** out);
void vo_pipeline_destroy(vo_pipeline* pl);
/* Make `frames` the current frame of every sequence (pyramid + detector).  init_tables != 0 also starts every table
 * from the detected corners, all unmatched (KLTTracker.__init__, klt.py:40-50).                                  */
int vo_pipeline_prime_dev(vo_pipeline* pl, const uint8_t* d_frames, size_t pitch, size_t frame_stride, int init_tables, void* stream);
int vo_pipeline_prime_host(vo_pipeline* pl, const uint8_t* h_frames, int init_tables);
/* Device-resident step (asynchronous): d_frames uint8 [n_seq] frames.  The summary [n_seq][VO_PIPE_SUMMARY_DOUBLES]
 * of the last step_dev is at vo_pipeline_summary_dev().                                                          */
int vo_pipeline_step_dev(vo_pipeline* pl, const uint8_t* d_frames, size_t pitch, size_t frame_stride, void* stream);
/* The two-view bootstrap of src/main.py:203-231 as one step of the resident pipeline: the primed frame's corners are
 * tracked into `frames` (klt.py:233-266), the matched rows go through vo_bootstrap_* (triangulation.py:88-350, threshold /
 * confidence as main.py:185-193 passes them: 0.25, 0.999; max_iters = OpenCV's default 1000), then update_with_local_pose,
 * update_with_local_landmarks and reset_outliers (state.py:25-110, 167-178) are applied to the table.  The summary row
 * holds the new pose; counters: 3 = matched pairs, 4 = inliers that became landmarks, 6 = triangulated rows, 8 = RANSAC
 * iterations, 7 |= 2 when no model was found (fewer than 15 pairs or a degenerate scene; the table is left matched). */
int vo_pipeline_bootstrap_dev(vo_pipeline* pl, const uint8_t* d_frames, size_t pitch, size_t frame_stride, double threshold,
                              double confidence, int max_iters, void* stream);
int vo_pipeline_bootstrap_host(vo_pipeline* pl, const uint8_t* h_frames, double threshold, double confidence, int max_iters,
                               double* h_summary);
const double* vo_pipeline_summary_dev(vo_pipeline* pl);
/* A step's detector and its pose / state-update kernels run on internal streams and may still be in flight when
 * step_dev returns control to `stream` (they run under the NEXT step's tracker).  vo_pipeline_sync_dev makes `stream`
 * wait for them: call it before timing events, before reading the summary, before destroying the frames' memory.   */
int vo_pipeline_sync_dev(vo_pipeline* pl, void* stream);
/* Host-buffer steps, pipelined like vo_frontend_*: prefetch uploads the NEXT step's frames (tightly packed) under the
 * current step, submit enqueues the step and the download of its summary, wait blocks for the oldest submitted step. */
int vo_pipeline_prefetch_host(vo_pipeline* pl, const uint8_t* h_frames);
int vo_pipeline_submit_host(vo_pipeline* pl, const uint8_t* h_frames, double* h_summary);
int vo_pipeline_wait_host(vo_pipeline* pl);
int vo_pipeline_step_host(vo_pipeline* pl, const uint8_t* h_frames, double* h_summary);
/* Table access (synchronous; bootstrap hand-over, tests, the Python mirror of Features).  Columns: kp float32 [n][2],
 * land float64 [n][3], state uint8 [n], track float32 [n][2], pose float64 [n][12] (camera-to-world, row-major 3x4),
 * cand uint8 [n]; h_c2w float64 [48] = current and previous camera-to-world pose (3x4), the estimator's world-to-camera
 * pose and the unrefined RANSAC model (both R row-major | t); h_scalars int32 [3 + VO_PIPE_NCOUNTS] = {num_features,
 * n_iterations, P3P population, counters}; h_inliers uint8 [P3P population]; h_rng uint64 [6].  Any may be NULL.    */
int vo_pipeline_read_table_host(vo_pipeline* pl, int seq, int* n_rows, float* h_kp, double* h_land, uint8_t* h_state, float* h_track,
                                double* h_pose, uint8_t* h_cand, double* h_c2w, int32_t* h_scalars, uint8_t* h_inliers,
                                uint64_t* h_rng);
/* Replace a sequence's table (e.g. after the host-side two-view bootstrap, main.py:203-231); n_rows < 0 keeps the
 * rows and only sets the scalars that are given.  num_features < 0 and
 * n_iterations <= 0 keep the current values; h_c2w (3x4) and h_rng (numpy PCG64 state: state hi/lo, inc hi/lo,
 * has_uint32, uinteger) may be NULL.                                                                              */
int vo_pipeline_write_table_host(vo_pipeline* pl, int seq, int n_rows, const float* h_kp, const double* h_land, const uint8_t* h_state,
                                 const float* h_track, const double* h_pose, const double* h_c2w, int num_features, int n_iterations,
                                 const uint64_t* h_rng);
/* corners of the current frame as the detector left them: int32 (Harris) or float32 (Shi-Tomasi) [det_max_corners][2] */
int vo_pipeline_read_detections_host(vo_pipeline* pl, int seq, void* h_xy, int* n);
/* Test hook: n_draws samples of numpy's Generator(PCG64).choice(arange(N), size=4, replace=False) (ransac.py:92-94)
 * from the given generator state; the state is advanced in place.                                                 */
int vo_test_pcg64_choice4_host(vo_ctx* ctx, uint64_t* state6, int N, int n_draws, int32_t* h_out);


#ifdef __cplusplus
}
#endif
#endif /* VO_B200_H */
