// The VO front end as one resident object: per-sequence state (image pyramids of the previous and
// the current frame, last keypoints) stays in HBM; one call advances every sequence by one frame:
//   pyramid -> KLT (previous keypoints into the new frame) -> Harris on the new frame ->
//   P3P-RANSAC on the supplied 3D-2D correspondences -> DLT triangulation of the supplied tracks.
// This is the loop body of /root/reference/src/main.py:248-287 (tracker.trackFeatures,
// pose_estimator.estimate_pose, triangulator.triangulate_candidates) for S independent sequences.
// Host code only: it sequences the launchers of harris.cu / klt.cu / p3p.cu / triangulation.cu on
// one stream, so the whole step is capturable in a CUDA graph.
#include "../../include/vo_b200.h"
#include <cstdlib>
#include <initializer_list>

#include "common.cuh"
#include "launchers.cuh"

struct vo_frontend {
    vo_ctx* ctx;
    vo_frontend_params p;
    int n_levels;
    size_t pitch0, frame_bytes;
    unsigned char* base;      // one allocation, carved below
    uint8_t* pyr[2];
    int cur;
    long long steps;
    double* resp;
    int32_t* kp;
    float *pts_prev, *pts_next, *err;
    uint8_t* status;
    double* models; uint8_t* valid; int32_t* counts;
    int32_t *best4, *consumed, *iters_out; uint8_t* inliers; double* pose;
    double* tri_out;
    // staging for the host entry point
    double *s_landmarks, *s_kp2d, *s_tri_p1, *s_tri_p2, *s_tri_proj1, *s_tri_proj2;
    int32_t *s_samples, *s_table;
    uint8_t* s_frames;        // tightly packed upload target of the host entry point
    unsigned char* stage[2];  // two staging sets (frames + P3P + triangulation inputs) for upload / compute overlap
    size_t so_fr, so_l, so_k, so_s, so_tb, so_t1, so_t2, so_tp1, so_tp2, stage_bytes;
    int stage_next;           // set the next prefetch / upload writes (sets alternate)
    int prefetched;           // number of uploaded-but-not-yet-consumed steps (0..2), oldest first
    cudaStream_t copy_stream;  // uploads / downloads of the host entry point overlap compute on the main stream
    cudaEvent_t ev_up[8], ev_done[8];
    // the stages of a step form a fork-join graph: {pyramid -> KLT}, {Harris response -> NMS}, {P3P -> triangulation};
    // the latency-bound per-frame kernels (NMS bands, RANSAC replay) then run under the issue-bound ones
    cudaStream_t side[2];
    cudaEvent_t ev_fork, ev_level0, ev_harris, ev_pose;
    int overlap;               // 0: everything on one stream (VO_FRONTEND_SERIAL=1)
    // pipelined host entry point: the step's outputs alternate between two sets so that the download of step t
    // (on its own stream) runs under the compute of step t+1
    struct OutSet { int32_t* kp; float* pts_next; float* err; uint8_t* status; int32_t* best4; uint8_t* inliers; double* pose; double* tri_out; } outs[2];
    cudaStream_t down_stream;
    cudaEvent_t ev_res[2];
    int in_flight, sub_next;
};

extern "C" {

int vo_frontend_create(vo_ctx* ctx, const vo_frontend_params* prm, vo_frontend** out) {
    VO_REQUIRE(ctx && prm && out, "vo_frontend_create: null argument");
    *out = nullptr;
    const vo_frontend_params& p = *prm;
    VO_REQUIRE(p.n_seq >= 1 && p.H > 0 && p.W > 0 && p.num_keypoints >= 1 && p.n_corr >= 4 && p.n_hyp >= 1 && p.n_tri >= 0,
               "vo_frontend_create: bad sizes");
    VO_CUDA(cudaSetDevice(ctx->device));
    vo_frontend* fe = new vo_frontend();
    fe->ctx = ctx; fe->p = p; fe->cur = 0; fe->steps = 0;
    int lh[8], lw[8]; size_t lp[8], lo[8];
    int rc = vo_klt_layout_host(p.H, p.W, p.klt_max_level, p.klt_win, &fe->n_levels, lh, lw, lp, lo, &fe->frame_bytes);
    if (rc) { delete fe; return rc; }
    fe->pitch0 = lp[0];
    const size_t S = p.n_seq, K = p.num_keypoints, N = p.n_corr, Hn = p.n_hyp, T = p.n_tri, npx = (size_t)p.H * p.W;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_p0 = carve(S * fe->frame_bytes), o_p1 = carve(S * fe->frame_bytes), o_resp = carve(S * npx * 8);
    const size_t o_kp = carve(S * K * 8), o_pp = carve(S * K * 8), o_pn = carve(S * K * 8), o_err = carve(S * K * 4);
    const size_t o_st = carve(S * K), o_m = carve(S * Hn * 96), o_v = carve(S * Hn), o_c = carve(S * Hn * 4);
    const size_t o_b4 = carve(S * 16), o_con = carve(S * 4), o_it = carve(S * 4), o_in = carve(S * N), o_pose = carve(S * 96);
    const size_t o_to = carve(S * T * 24 + 256);
    const size_t o2_kp = carve(S * K * 8), o2_pn = carve(S * K * 8), o2_err = carve(S * K * 4), o2_st = carve(S * K);
    const size_t o2_b4 = carve(S * 16), o2_in = carve(S * N), o2_pose = carve(S * 96), o2_to = carve(S * T * 24 + 256);
    const size_t o_sl = carve(S * N * 24), o_sk = carve(S * N * 16), o_ss = carve(S * Hn * 16), o_stb = carve((N + 1) * 4);
    const size_t o_fr = carve(S * npx + 256);
    const size_t o_t1 = carve(S * T * 16 + 256), o_t2 = carve(S * T * 16 + 256), o_tp1 = carve(S * T * 96 + 256), o_tp2 = carve(S * 96);
    cudaError_t e = cudaMalloc(&fe->base, off);
    if (e != cudaSuccess) {
        vo_set_error("vo_frontend_create: cudaMalloc(%zu) -> %s", off, cudaGetErrorString(e));
        delete fe;
        return VO_ERR_CUDA;
    }
    // everything below may fail half way: the lambda returns the error and vo_frontend_destroy releases what exists
    auto finish = [&]() -> int {
    unsigned char* b = fe->base;
    VO_CUDA(cudaMemsetAsync(b, 0, off, ctx->stream));
    fe->pyr[0] = b + o_p0; fe->pyr[1] = b + o_p1; fe->resp = (double*)(b + o_resp);
    fe->kp = (int32_t*)(b + o_kp); fe->pts_prev = (float*)(b + o_pp); fe->pts_next = (float*)(b + o_pn);
    fe->err = (float*)(b + o_err); fe->status = b + o_st;
    fe->models = (double*)(b + o_m); fe->valid = b + o_v; fe->counts = (int32_t*)(b + o_c);
    fe->best4 = (int32_t*)(b + o_b4); fe->consumed = (int32_t*)(b + o_con); fe->iters_out = (int32_t*)(b + o_it);
    fe->inliers = b + o_in; fe->pose = (double*)(b + o_pose); fe->tri_out = (double*)(b + o_to);
    fe->s_landmarks = (double*)(b + o_sl); fe->s_kp2d = (double*)(b + o_sk); fe->s_samples = (int32_t*)(b + o_ss);
    fe->s_table = (int32_t*)(b + o_stb); fe->s_tri_p1 = (double*)(b + o_t1); fe->s_tri_p2 = (double*)(b + o_t2);
    fe->s_tri_proj1 = (double*)(b + o_tp1); fe->s_tri_proj2 = (double*)(b + o_tp2);
    fe->s_frames = b + o_fr;
    fe->outs[0] = {fe->kp, fe->pts_next, fe->err, fe->status, fe->best4, fe->inliers, fe->pose, fe->tri_out};
    fe->outs[1] = {(int32_t*)(b + o2_kp), (float*)(b + o2_pn), (float*)(b + o2_err), b + o2_st, (int32_t*)(b + o2_b4), b + o2_in,
                   (double*)(b + o2_pose), (double*)(b + o2_to)};
    fe->in_flight = 0; fe->sub_next = 0;
    {
        size_t o = 0;
        auto cv = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
        fe->so_fr = cv(S * npx + 256); fe->so_l = cv(S * N * 24); fe->so_k = cv(S * N * 16); fe->so_s = cv(S * Hn * 16);
        fe->so_tb = cv((N + 1) * 4); fe->so_t1 = cv(S * T * 16 + 256); fe->so_t2 = cv(S * T * 16 + 256);
        fe->so_tp1 = cv(S * T * 96 + 256); fe->so_tp2 = cv(S * 96);
        fe->stage_bytes = o;
        for (int i = 0; i < 2; i++) VO_CUDA(cudaMalloc(&fe->stage[i], o));
        fe->stage_next = 0; fe->prefetched = 0;
    }
    VO_CUDA(cudaStreamCreateWithFlags(&fe->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 8; i++) {
        VO_CUDA(cudaEventCreateWithFlags(&fe->ev_up[i], cudaEventDisableTiming));
        VO_CUDA(cudaEventCreateWithFlags(&fe->ev_done[i], cudaEventDisableTiming));
    }
    {   // the side streams get the highest priority: their per-frame, latency-bound kernels are then scheduled
        // into the SM slots the tracker's short-lived CTAs free, instead of queueing behind its whole grid
        int prio_lo = 0, prio_hi = 0;
        VO_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        for (int i = 0; i < 2; i++) VO_CUDA(cudaStreamCreateWithPriority(&fe->side[i], cudaStreamNonBlocking, prio_hi));
    }
    VO_CUDA(cudaStreamCreateWithFlags(&fe->down_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) VO_CUDA(cudaEventCreateWithFlags(&fe->ev_res[i], cudaEventDisableTiming));
    VO_CUDA(cudaEventCreateWithFlags(&fe->ev_fork, cudaEventDisableTiming));
    VO_CUDA(cudaEventCreateWithFlags(&fe->ev_level0, cudaEventDisableTiming));
    VO_CUDA(cudaEventCreateWithFlags(&fe->ev_harris, cudaEventDisableTiming));
    VO_CUDA(cudaEventCreateWithFlags(&fe->ev_pose, cudaEventDisableTiming));
    fe->overlap = !ctx->env_frontend_serial;
    // the detector's working memory is reserved now: a step never allocates (and is capturable into a CUDA graph)
    { const int rc2 = vo_harris_nms_reserve(ctx, p.n_seq, p.H, p.W, p.nms_radius, p.num_keypoints); if (rc2) return rc2; }
    VO_CUDA(cudaStreamSynchronize(ctx->stream));
    return VO_OK;
    };
    rc = finish();
    if (rc) { vo_frontend_destroy(fe); return rc; }
    *out = fe;
    return VO_OK;
}

// Tolerates a partially constructed object (vo_frontend_create's failure path): every handle is checked.
void vo_frontend_destroy(vo_frontend* fe) {
    if (!fe) return;
    cudaSetDevice(fe->ctx->device);
    cudaStreamSynchronize(fe->ctx->stream);
    if (fe->copy_stream) cudaStreamSynchronize(fe->copy_stream);
    for (int i = 0; i < 8; i++) { if (fe->ev_up[i]) cudaEventDestroy(fe->ev_up[i]); if (fe->ev_done[i]) cudaEventDestroy(fe->ev_done[i]); }
    for (int i = 0; i < 2; i++) if (fe->side[i]) { cudaStreamSynchronize(fe->side[i]); cudaStreamDestroy(fe->side[i]); }
    for (cudaEvent_t ev : {fe->ev_fork, fe->ev_level0, fe->ev_harris, fe->ev_pose}) if (ev) cudaEventDestroy(ev);
    if (fe->down_stream) { cudaStreamSynchronize(fe->down_stream); cudaStreamDestroy(fe->down_stream); }
    for (int i = 0; i < 2; i++) if (fe->ev_res[i]) cudaEventDestroy(fe->ev_res[i]);
    if (fe->copy_stream) cudaStreamDestroy(fe->copy_stream);
    if (fe->stage[0]) cudaFree(fe->stage[0]);
    if (fe->stage[1]) cudaFree(fe->stage[1]);
    if (fe->base) cudaFree(fe->base);
    (void)cudaGetLastError();
    delete fe;
}

int vo_frontend_outputs(vo_frontend* fe, vo_frontend_outputs_t* o) {
    VO_REQUIRE(fe && o, "vo_frontend_outputs: null argument");
    o->d_resp = fe->resp; o->d_kp_xy = fe->kp; o->d_tracked = fe->pts_next; o->d_status = fe->status; o->d_err = fe->err;
    o->d_best4 = fe->best4; o->d_inliers = fe->inliers; o->d_pose = fe->pose; o->d_tri_out = fe->tri_out;
    o->d_counts = fe->counts; o->d_cur_pyramid = fe->pyr[fe->cur];
    o->pyr_pitch0 = fe->pitch0; o->pyr_frame_bytes = fe->frame_bytes;
    return VO_OK;
}

// the level-0 slots of the pyramid the NEXT step will fill (so a caller can upload frames in place)
uint8_t* vo_frontend_next_frame_slot(vo_frontend* fe, size_t* pitch, size_t* frame_stride) {
    if (!fe) return nullptr;
    if (pitch) *pitch = fe->pitch0;
    if (frame_stride) *frame_stride = fe->frame_bytes;
    return fe->pyr[1 - fe->cur];
}

// All stages for sequences [s0, s0 + n) on stream s.  Inputs are indexed from sequence 0.
static int frontend_run_range(vo_frontend* fe, int s0, int n, const uint8_t* d_frames, size_t pitch, size_t frame_stride,
                              const double* d_landmarks, const double* d_kp2d, const double* K9,
                              const int32_t* d_sample_idx, const int32_t* d_iters_table, int initial_iters,
                              const double* d_tri_p1, const double* d_tri_p2, const double* d_tri_proj1,
                              const double* d_tri_proj2, cudaStream_t s) {
    vo_ctx* ctx = fe->ctx;
    const vo_frontend_params& p = fe->p;
    const int nxt = 1 - fe->cur;
    const size_t K = p.num_keypoints, N = p.n_corr, Hn = p.n_hyp, T = p.n_tri, npx = (size_t)p.H * p.W;
    uint8_t* pyr_new = fe->pyr[nxt] + (size_t)s0 * fe->frame_bytes;
    const uint8_t* pyr_old = fe->pyr[fe->cur] + (size_t)s0 * fe->frame_bytes;
    int rc;
    // streams: s carries pyramid -> KLT; sh the Harris detector; sp pose + triangulation.  Both side streams
    // fork from s and join it again, so the step is still one unit of work on s (and capturable into a graph).
    cudaStream_t sh = fe->overlap ? fe->side[0] : s, sp = fe->overlap ? fe->side[1] : s;
    if (fe->overlap) {
        VO_CUDA(cudaEventRecord(fe->ev_fork, s));
        VO_CUDA(cudaStreamWaitEvent(sp, fe->ev_fork, 0));
    }
    // 1. pyramid of the new frames (level 0 copied unless already uploaded into the slot)
    if ((rc = vo_launch_klt_pyramid(ctx, d_frames + (size_t)s0 * frame_stride, n, p.H, p.W, pitch, frame_stride,
                                    p.klt_max_level, p.klt_win, pyr_new, s, fe->overlap ? fe->ev_level0 : nullptr))) return rc;
    if (fe->overlap) VO_CUDA(cudaStreamWaitEvent(sh, fe->ev_level0, 0));
    // 2. KLT: last frame's keypoints into the new frame (klt.py:233-239)
    if (fe->steps > 0) {
        if ((rc = vo_launch_klt_track(ctx, pyr_old, pyr_new, n, p.H, p.W, p.klt_max_level, p.klt_win, p.klt_max_iters,
                                      p.klt_epsilon, p.klt_min_eig, fe->pts_prev + (size_t)s0 * K * 2, p.num_keypoints,
                                      fe->pts_next + (size_t)s0 * K * 2, fe->status + (size_t)s0 * K,
                                      fe->err + (size_t)s0 * K, s))) return rc;
    }
    // 3. Harris on the new frame (harris.py:86-158); its keypoints seed the next step's tracking
    if ((rc = vo_launch_harris_response(ctx, pyr_new, n, p.H, p.W, fe->pitch0, fe->frame_bytes, p.patch_size, p.kappa,
                                        fe->resp + (size_t)s0 * npx, sh))) return rc;
    if ((rc = vo_launch_harris_nms(ctx, fe->resp + (size_t)s0 * npx, n, p.H, p.W, p.nms_radius, p.num_keypoints,
                                   fe->kp + (size_t)s0 * K * 2, nullptr, sh))) return rc;
    // 4. P3P + RANSAC (p3p.py:123-186 with use_opencv=False)
    if ((rc = vo_launch_p3p_score(ctx, d_landmarks + (size_t)s0 * N * 3, d_kp2d + (size_t)s0 * N * 2, n, p.n_corr, K9,
                                  d_sample_idx + (size_t)s0 * Hn * 4, p.n_hyp, p.p3p_threshold, 0,
                                  fe->models + (size_t)s0 * Hn * 12, fe->valid + (size_t)s0 * Hn,
                                  fe->counts + (size_t)s0 * Hn, sp))) return rc;
    if ((rc = vo_launch_p3p_select(ctx, d_landmarks + (size_t)s0 * N * 3, d_kp2d + (size_t)s0 * N * 2, n, p.n_corr, K9,
                                   fe->models + (size_t)s0 * Hn * 12, fe->valid + (size_t)s0 * Hn,
                                   fe->counts + (size_t)s0 * Hn, p.n_hyp, p.p3p_threshold, 0, d_iters_table, initial_iters, 0, -1,
                                   fe->best4 + (size_t)s0 * 4, fe->consumed + s0, fe->iters_out + s0,
                                   fe->inliers + (size_t)s0 * N, fe->pose + (size_t)s0 * 12, sp))) return rc;
    // 5. triangulation of new landmarks (triangulation.py:38-86)
    if (p.n_tri > 0) {
        VO_REQUIRE(d_tri_p1 && d_tri_p2 && d_tri_proj1 && d_tri_proj2, "vo_frontend_step: null triangulation input");
        if ((rc = vo_launch_triangulate(ctx, d_tri_p1 + (size_t)s0 * T * 2, d_tri_p2 + (size_t)s0 * T * 2, n * p.n_tri,
                                        d_tri_proj1 + (size_t)s0 * T * 12, 1, d_tri_proj2 + (size_t)s0 * 12, p.n_tri,
                                        p.tri_mode, fe->tri_out + (size_t)s0 * T * 3, sp))) return rc;
    }
    // join: the new keypoints replace the tracked set only after the tracker has read it
    if (fe->overlap) {
        VO_CUDA(cudaEventRecord(fe->ev_harris, sh));
        VO_CUDA(cudaEventRecord(fe->ev_pose, sp));
        VO_CUDA(cudaStreamWaitEvent(s, fe->ev_harris, 0));
        VO_CUDA(cudaStreamWaitEvent(s, fe->ev_pose, 0));
    }
    if ((rc = vo_launch_kp_to_points(ctx, fe->kp + (size_t)s0 * K * 2, (size_t)n * K, fe->pts_prev + (size_t)s0 * K * 2, s))) return rc;
    return VO_OK;
}

int vo_frontend_step_dev(vo_frontend* fe, const uint8_t* d_frames, size_t pitch, size_t frame_stride,
                         const double* d_landmarks, const double* d_kp2d, const double* K9,
                         const int32_t* d_sample_idx, const int32_t* d_iters_table, int initial_iters,
                         const double* d_tri_p1, const double* d_tri_p2, const double* d_tri_proj1,
                         const double* d_tri_proj2, void* stream) {
    VO_REQUIRE(fe && d_frames && d_landmarks && d_kp2d && K9 && d_sample_idx && d_iters_table,
               "vo_frontend_step_dev: null argument");
    VO_CUDA(cudaSetDevice(fe->ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : fe->ctx->stream;
    int rc = frontend_run_range(fe, 0, fe->p.n_seq, d_frames, pitch, frame_stride, d_landmarks, d_kp2d, K9, d_sample_idx,
                                d_iters_table, initial_iters, d_tri_p1, d_tri_p2, d_tri_proj1, d_tri_proj2, s);
    if (rc) return rc;
    fe->cur = 1 - fe->cur;
    fe->steps++;
    return VO_OK;
}

// enqueue the upload of one step's inputs into staging set `set` on the copy stream
static int frontend_upload(vo_frontend* fe, int set, const uint8_t* h_frames, const double* h_landmarks,
                           const double* h_kp2d, const int32_t* h_sample_idx, const int32_t* h_iters_table,
                           const double* h_tri_p1, const double* h_tri_p2, const double* h_tri_proj1,
                           const double* h_tri_proj2) {
    const vo_frontend_params& p = fe->p;
    const size_t S = p.n_seq, N = p.n_corr, Hn = p.n_hyp, T = p.n_tri, npx = (size_t)p.H * p.W;
    unsigned char* b = fe->stage[set];
    cudaStream_t cs = fe->copy_stream;
    VO_CUDA(cudaStreamWaitEvent(cs, fe->ev_done[set], 0));   // the step that read this set before has finished
    // one contiguous copy per array (a pitched 2-D copy of 1241-byte rows is several times slower over PCIe);
    // the pyramid builder re-pitches the frames into the level-0 slots on the device.
    VO_CUDA(cudaMemcpyAsync(b + fe->so_fr, h_frames, S * npx, cudaMemcpyHostToDevice, cs));
    VO_CUDA(cudaMemcpyAsync(b + fe->so_l, h_landmarks, S * N * 24, cudaMemcpyHostToDevice, cs));
    VO_CUDA(cudaMemcpyAsync(b + fe->so_k, h_kp2d, S * N * 16, cudaMemcpyHostToDevice, cs));
    VO_CUDA(cudaMemcpyAsync(b + fe->so_s, h_sample_idx, S * Hn * 16, cudaMemcpyHostToDevice, cs));
    VO_CUDA(cudaMemcpyAsync(b + fe->so_tb, h_iters_table, (N + 1) * 4, cudaMemcpyHostToDevice, cs));
    if (T > 0) {
        VO_CUDA(cudaMemcpyAsync(b + fe->so_t1, h_tri_p1, S * T * 16, cudaMemcpyHostToDevice, cs));
        VO_CUDA(cudaMemcpyAsync(b + fe->so_t2, h_tri_p2, S * T * 16, cudaMemcpyHostToDevice, cs));
        VO_CUDA(cudaMemcpyAsync(b + fe->so_tp1, h_tri_proj1, S * T * 96, cudaMemcpyHostToDevice, cs));
        VO_CUDA(cudaMemcpyAsync(b + fe->so_tp2, h_tri_proj2, S * 96, cudaMemcpyHostToDevice, cs));
    }
    VO_CUDA(cudaEventRecord(fe->ev_up[set], cs));
    return VO_OK;
}

int vo_frontend_prefetch_host(vo_frontend* fe, const uint8_t* h_frames, const double* h_landmarks, const double* h_kp2d,
                              const int32_t* h_sample_idx, const int32_t* h_iters_table, const double* h_tri_p1,
                              const double* h_tri_p2, const double* h_tri_proj1, const double* h_tri_proj2) {
    VO_REQUIRE(fe && h_frames && h_landmarks && h_kp2d && h_sample_idx && h_iters_table, "vo_frontend_prefetch_host: null argument");
    VO_REQUIRE(fe->prefetched < 2, "vo_frontend_prefetch_host: two prefetched steps are already pending");
    if (fe->p.n_tri > 0) VO_REQUIRE(h_tri_p1 && h_tri_p2 && h_tri_proj1 && h_tri_proj2, "vo_frontend_prefetch_host: null triangulation buffer");
    VO_CUDA(cudaSetDevice(fe->ctx->device));
    const int set = fe->stage_next;
    int rc = frontend_upload(fe, set, h_frames, h_landmarks, h_kp2d, h_sample_idx, h_iters_table, h_tri_p1, h_tri_p2,
                             h_tri_proj1, h_tri_proj2);
    if (rc) return rc;
    fe->prefetched++; fe->stage_next = 1 - set;
    return VO_OK;
}

int vo_frontend_submit_host(vo_frontend* fe, const uint8_t* h_frames, const double* h_landmarks, const double* h_kp2d,
                            const double* K9, const int32_t* h_sample_idx, const int32_t* h_iters_table,
                            int initial_iters, const double* h_tri_p1, const double* h_tri_p2,
                            const double* h_tri_proj1, const double* h_tri_proj2, int32_t* h_kp_xy, float* h_tracked,
                            uint8_t* h_status, float* h_err, int32_t* h_best4, uint8_t* h_inliers, double* h_pose,
                            double* h_tri_out) {
    VO_REQUIRE(fe && K9 && h_kp_xy && h_pose, "vo_frontend_submit_host: null argument");
    VO_REQUIRE(fe->in_flight < 2, "vo_frontend_submit_host: two submitted steps are already in flight (call vo_frontend_wait_host)");
    vo_ctx* ctx = fe->ctx;
    VO_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream, ds = fe->down_stream;
    const vo_frontend_params& p = fe->p;
    const size_t S = p.n_seq, K = p.num_keypoints, N = p.n_corr, T = p.n_tri, npx = (size_t)p.H * p.W;
    if (T > 0) VO_REQUIRE(h_tri_out, "vo_frontend_submit_host: null triangulation output");
    int set;
    if (fe->prefetched) {                 // inputs were uploaded by vo_frontend_prefetch_host (host pointers may be NULL)
        set = (fe->stage_next + 2 - fe->prefetched) & 1;     // oldest pending set
        fe->prefetched--;
    } else {
        VO_REQUIRE(h_frames && h_landmarks && h_kp2d && h_sample_idx && h_iters_table, "vo_frontend_submit_host: null input");
        if (T > 0) VO_REQUIRE(h_tri_p1 && h_tri_p2 && h_tri_proj1 && h_tri_proj2, "vo_frontend_submit_host: null triangulation buffer");
        set = fe->stage_next;
        int rc = frontend_upload(fe, set, h_frames, h_landmarks, h_kp2d, h_sample_idx, h_iters_table, h_tri_p1, h_tri_p2,
                                 h_tri_proj1, h_tri_proj2);
        if (rc) return rc;
        fe->stage_next = 1 - set;
    }
    const bool had_prev = fe->steps > 0;
    unsigned char* b = fe->stage[set];
    // this step writes output set `slot`; its previous contents (two submits ago) must have left the device
    const int slot = fe->sub_next;
    const vo_frontend::OutSet& o = fe->outs[slot];
    VO_CUDA(cudaStreamWaitEvent(s, fe->ev_res[slot], 0));
    fe->kp = o.kp; fe->pts_next = o.pts_next; fe->err = o.err; fe->status = o.status;
    fe->best4 = o.best4; fe->inliers = o.inliers; fe->pose = o.pose; fe->tri_out = o.tri_out;
    VO_CUDA(cudaStreamWaitEvent(s, fe->ev_up[set], 0));
    int rc = frontend_run_range(fe, 0, (int)S, b + fe->so_fr, (size_t)p.W, npx, (const double*)(b + fe->so_l),
                                (const double*)(b + fe->so_k), K9, (const int32_t*)(b + fe->so_s),
                                (const int32_t*)(b + fe->so_tb), initial_iters, (const double*)(b + fe->so_t1),
                                (const double*)(b + fe->so_t2), (const double*)(b + fe->so_tp1),
                                (const double*)(b + fe->so_tp2), s);
    if (rc) return rc;
    fe->cur = 1 - fe->cur;
    fe->steps++;
    VO_CUDA(cudaEventRecord(fe->ev_done[set], s));      // staging set free again, outputs complete
    // results come back on their own stream, under the next step's compute
    VO_CUDA(cudaStreamWaitEvent(ds, fe->ev_done[set], 0));
    VO_CUDA(cudaMemcpyAsync(h_kp_xy, o.kp, S * K * 8, cudaMemcpyDeviceToHost, ds));
    if (had_prev && h_tracked) VO_CUDA(cudaMemcpyAsync(h_tracked, o.pts_next, S * K * 8, cudaMemcpyDeviceToHost, ds));
    if (had_prev && h_status) VO_CUDA(cudaMemcpyAsync(h_status, o.status, S * K, cudaMemcpyDeviceToHost, ds));
    if (had_prev && h_err) VO_CUDA(cudaMemcpyAsync(h_err, o.err, S * K * 4, cudaMemcpyDeviceToHost, ds));
    if (h_best4) VO_CUDA(cudaMemcpyAsync(h_best4, o.best4, S * 16, cudaMemcpyDeviceToHost, ds));
    if (h_inliers) VO_CUDA(cudaMemcpyAsync(h_inliers, o.inliers, S * N, cudaMemcpyDeviceToHost, ds));
    VO_CUDA(cudaMemcpyAsync(h_pose, o.pose, S * 96, cudaMemcpyDeviceToHost, ds));
    if (T > 0) VO_CUDA(cudaMemcpyAsync(h_tri_out, o.tri_out, S * T * 24, cudaMemcpyDeviceToHost, ds));
    VO_CUDA(cudaEventRecord(fe->ev_res[slot], ds));
    fe->in_flight++;
    fe->sub_next = 1 - slot;
    return VO_OK;
}

int vo_frontend_wait_host(vo_frontend* fe) {
    VO_REQUIRE(fe, "vo_frontend_wait_host: null argument");
    VO_REQUIRE(fe->in_flight > 0, "vo_frontend_wait_host: nothing was submitted");
    VO_CUDA(cudaSetDevice(fe->ctx->device));
    const int slot = (fe->sub_next + 2 - fe->in_flight) & 1;   // oldest submitted step
    VO_CUDA(cudaEventSynchronize(fe->ev_res[slot]));
    fe->in_flight--;
    return VO_OK;
}

int vo_frontend_step_host(vo_frontend* fe, const uint8_t* h_frames, const double* h_landmarks, const double* h_kp2d,
                          const double* K9, const int32_t* h_sample_idx, const int32_t* h_iters_table,
                          int initial_iters, const double* h_tri_p1, const double* h_tri_p2,
                          const double* h_tri_proj1, const double* h_tri_proj2, int32_t* h_kp_xy, float* h_tracked,
                          uint8_t* h_status, float* h_err, int32_t* h_best4, uint8_t* h_inliers, double* h_pose,
                          double* h_tri_out) {
    VO_REQUIRE(fe && fe->in_flight == 0, "vo_frontend_step_host: submitted steps are still in flight (call vo_frontend_wait_host)");
    int rc = vo_frontend_submit_host(fe, h_frames, h_landmarks, h_kp2d, K9, h_sample_idx, h_iters_table, initial_iters, h_tri_p1,
                                     h_tri_p2, h_tri_proj1, h_tri_proj2, h_kp_xy, h_tracked, h_status, h_err, h_best4, h_inliers,
                                     h_pose, h_tri_out);
    if (rc) return rc;
    return vo_frontend_wait_host(fe);
}

}  // extern "C"
