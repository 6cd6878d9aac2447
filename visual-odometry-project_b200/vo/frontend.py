"""Resident VO front end (vo_frontend_* of include/vo_b200.h): S independent sequences advance one
frame per call -- pyramid, KLT, Harris, P3P-RANSAC, triangulation -- with all per-sequence state
kept in HBM.  Mirrors the loop body of the reference's src/main.py:248-287."""
import ctypes as C

import numpy as np

from . import _native as nat


class Frontend:
    def __init__(self, n_seq, H, W, num_keypoints=1000, patch_size=9, kappa=0.09, nms_radius=5,
                 klt_win=17, klt_max_level=2, klt_max_iters=10, klt_epsilon=0.03, klt_min_eig=1e-4,
                 n_corr=1000, n_hyp=512, p3p_threshold=1.25, n_tri=1000, tri_mode=1, ctx=None):
        self.ctx = ctx if ctx is not None else nat.default_context(0)
        self.params = nat.FrontendParams(n_seq, H, W, patch_size, kappa, nms_radius, num_keypoints, klt_win,
                                         klt_max_level, klt_max_iters, klt_epsilon, klt_min_eig, n_corr, n_hyp,
                                         p3p_threshold, n_tri, tri_mode)
        self._h = C.c_void_p()
        nat.check(nat.lib().vo_frontend_create(self.ctx.handle, C.byref(self.params), C.byref(self._h)),
                  "vo_frontend_create")

    def outputs(self) -> nat.FrontendOutputs:
        o = nat.FrontendOutputs()
        nat.check(nat.lib().vo_frontend_outputs(self._h, C.byref(o)), "vo_frontend_outputs")
        return o

    def next_frame_slot(self):
        pitch, stride = C.c_size_t(), C.c_size_t()
        p = nat.lib().vo_frontend_next_frame_slot(self._h, C.byref(pitch), C.byref(stride))
        return int(p), pitch.value, stride.value

    def step_dev(self, d_frames, pitch, frame_stride, d_landmarks, d_kp2d, K9, d_samples, d_table, initial_iters,
                 d_tri_p1, d_tri_p2, d_tri_proj1, d_tri_proj2, stream=0):
        """All d_* are raw device addresses (ints); K9 a host float64[9] array."""
        rc = nat.lib().vo_frontend_step_dev(self._h, d_frames, pitch, frame_stride, d_landmarks, d_kp2d, nat.ptr(K9),
                                            d_samples, d_table, int(initial_iters), d_tri_p1, d_tri_p2, d_tri_proj1,
                                            d_tri_proj2, stream or None)
        nat.check(rc, "vo_frontend_step_dev")

    def step_host(self, frames, landmarks, kp2d, K9, samples, table, initial_iters, tri_p1, tri_p2, tri_proj1,
                  tri_proj2, out):
        """numpy (ideally pinned) host buffers in, results into the numpy arrays of `out`
        (keys: kp_xy, tracked, status, err, best4, inliers, pose, tri_out)."""
        def hp(a):
            return a.ctypes.data_as(C.c_void_p) if a is not None else None
        rc = nat.lib().vo_frontend_step_host(
            self._h, hp(frames), hp(landmarks), hp(kp2d), hp(K9), hp(samples), hp(table), int(initial_iters),
            hp(tri_p1), hp(tri_p2), hp(tri_proj1), hp(tri_proj2), hp(out["kp_xy"]), hp(out.get("tracked")),
            hp(out.get("status")), hp(out.get("err")), hp(out.get("best4")), hp(out.get("inliers")), hp(out["pose"]),
            hp(out.get("tri_out")))
        nat.check(rc, "vo_frontend_step_host")

    def submit_host(self, frames, landmarks, kp2d, K9, samples, table, initial_iters, tri_p1, tri_p2, tri_proj1,
                    tri_proj2, out):
        """Pipelined step_host: enqueue the step and the download of its results into `out`, return at once.
        `wait_host()` blocks until the oldest submitted step's results are in its `out` arrays; at most two steps
        may be in flight, and their `out` arrays must not be touched until the matching wait returns."""
        def hp(a):
            return a.ctypes.data_as(C.c_void_p) if a is not None else None
        rc = nat.lib().vo_frontend_submit_host(
            self._h, hp(frames), hp(landmarks), hp(kp2d), hp(K9), hp(samples), hp(table), int(initial_iters),
            hp(tri_p1), hp(tri_p2), hp(tri_proj1), hp(tri_proj2), hp(out["kp_xy"]), hp(out.get("tracked")),
            hp(out.get("status")), hp(out.get("err")), hp(out.get("best4")), hp(out.get("inliers")), hp(out["pose"]),
            hp(out.get("tri_out")))
        nat.check(rc, "vo_frontend_submit_host")

    def wait_host(self):
        nat.check(nat.lib().vo_frontend_wait_host(self._h), "vo_frontend_wait_host")

    def prefetch_host(self, frames, landmarks, kp2d, samples, table, tri_p1, tri_p2, tri_proj1, tri_proj2):
        """Start uploading the NEXT step's inputs while the current step computes (see vo_frontend_prefetch_host)."""
        def hp(a):
            return a.ctypes.data_as(C.c_void_p) if a is not None else None
        rc = nat.lib().vo_frontend_prefetch_host(self._h, hp(frames), hp(landmarks), hp(kp2d), hp(samples), hp(table),
                                                 hp(tri_p1), hp(tri_p2), hp(tri_proj1), hp(tri_proj2))
        nat.check(rc, "vo_frontend_prefetch_host")

    def close(self):
        if self._h:
            nat.lib().vo_frontend_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
