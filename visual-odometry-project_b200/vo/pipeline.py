"""The chained, device-resident pipeline (vo_pipeline_* of include/vo_b200.h): the loop body of the reference's
src/main.py:248-287 for S independent sequences with every feature table kept in HBM.  A step uploads only the frames
and returns the pose and a few counters per sequence; `read_table` / `write_table` expose a sequence's table with the
columns of the reference's Features (src/vo/primitives/features.py) for the bootstrap hand-over and for tests."""
import ctypes as C

import numpy as np

from . import _native as nat

NCOUNTS = 12
SUMMARY_DOUBLES = 18
DETECTOR_NONE, DETECTOR_HARRIS, DETECTOR_GFTT = 0, 1, 2
COUNT_NAMES = ("n_rows", "n_tracked", "n_kept", "p3p_N", "n_inliers", "n_candidates", "n_tri", "flags", "n_iterations",
               "draws", "n_behind", "gn_iters")


class PipelineParams(C.Structure):
    """vo_pipeline_params of include/vo_b200.h."""
    _fields_ = [("n_seq", C.c_int), ("H", C.c_int), ("W", C.c_int), ("capacity", C.c_int),
                ("klt_win", C.c_int), ("klt_max_level", C.c_int), ("klt_max_iters", C.c_int),
                ("klt_epsilon", C.c_double), ("klt_min_eig", C.c_double), ("klt_error_threshold", C.c_float),
                ("redetect_fraction", C.c_double), ("detector", C.c_int), ("det_max_corners", C.c_int),
                ("patch_size", C.c_int), ("kappa", C.c_double), ("nms_radius", C.c_int),
                ("gftt_quality", C.c_double), ("gftt_min_distance", C.c_double), ("gftt_block_size", C.c_int),
                ("K", C.c_double * 9), ("Kinv", C.c_double * 9),
                ("p3p_threshold", C.c_double), ("p3p_inclusive", C.c_int),
                ("ransac_confidence", C.c_double), ("ransac_outlier_ratio", C.c_double), ("ransac_log1mconf", C.c_double),
                ("ransac_max_iterations", C.c_int), ("ransac_initial_iterations", C.c_int), ("refine", C.c_int),
                ("bearing_threshold", C.c_double), ("tri_mode", C.c_int)]


def rng_state6(rng: np.random.Generator) -> np.ndarray:
    """numpy PCG64 generator state as the six uint64 words the device sampler carries."""
    st = rng.bit_generator.state
    if st["bit_generator"] != "PCG64":
        raise ValueError("the reference's RANSAC uses numpy's default PCG64 generator")
    s, inc = st["state"]["state"], st["state"]["inc"]
    m = (1 << 64) - 1
    return np.array([s >> 64, s & m, inc >> 64, inc & m, st["has_uint32"], st["uinteger"]], dtype=np.uint64)


def rng_from_state6(state6) -> np.random.Generator:
    rng = np.random.default_rng(0)
    st = rng.bit_generator.state
    w = [int(x) for x in state6]
    st["state"]["state"] = (w[0] << 64) | w[1]
    st["state"]["inc"] = (w[2] << 64) | w[3]
    st["has_uint32"], st["uinteger"] = w[4], w[5]
    rng.bit_generator.state = st
    return rng


class Pipeline:
    def __init__(self, n_seq, H, W, K, *, capacity=2048, detector=DETECTOR_HARRIS, det_max_corners=1000, patch_size=9,
                 kappa=0.09, nms_radius=5, gftt_quality=0.01, gftt_min_distance=8.0, gftt_block_size=7, klt_win=17, klt_max_level=2, klt_max_iters=10, klt_epsilon=0.03,
                 klt_min_eig=1e-4, klt_error_threshold=100.0, redetect_fraction=0.8, p3p_threshold=1.25,
                 p3p_opencv=False, confidence=0.9999, outlier_ratio=0.9, max_iterations=10000, refine=True,
                 bearing_threshold=0.0075, tri_opencv=True, rng_seed=2023, ctx=None):
        self.ctx = ctx if ctx is not None else nat.default_context(0)
        K = np.asarray(K)
        Kinv = np.linalg.inv(K)                        # camera.py:92: in K's own dtype (float32 for the KITTI loader)
        p = PipelineParams()
        p.n_seq, p.H, p.W, p.capacity = int(n_seq), int(H), int(W), int(capacity)
        p.klt_win, p.klt_max_level, p.klt_max_iters = int(klt_win), int(klt_max_level), int(klt_max_iters)
        p.klt_epsilon, p.klt_min_eig, p.klt_error_threshold = float(klt_epsilon), float(klt_min_eig), float(klt_error_threshold)
        p.redetect_fraction = float(redetect_fraction)
        p.detector, p.det_max_corners = int(detector), int(det_max_corners)
        p.patch_size, p.kappa, p.nms_radius = int(patch_size), float(kappa), int(nms_radius)
        p.gftt_quality, p.gftt_min_distance, p.gftt_block_size = float(gftt_quality), float(gftt_min_distance), int(gftt_block_size)
        for i, v in enumerate(np.asarray(K, dtype=np.float64).reshape(9)):
            p.K[i] = float(v)
        for i, v in enumerate(np.asarray(Kinv, dtype=np.float64).reshape(9)):
            p.Kinv[i] = float(v)
        # p3p.py:149 (OpenCV rule: squared error <= reprojectionError^2) vs ransac.py:105 (error < inlier_threshold)
        p.p3p_threshold = float(p3p_threshold) ** 2 if p3p_opencv else float(p3p_threshold)
        p.p3p_inclusive = int(bool(p3p_opencv))
        p.ransac_confidence, p.ransac_outlier_ratio = float(confidence), float(outlier_ratio)
        p.ransac_log1mconf = float(np.log(1 - confidence))
        p.ransac_max_iterations = int(min(max_iterations, 2 ** 31 - 1))
        k0 = np.ceil(np.log(1 - confidence) / np.log(1 - (1 - outlier_ratio) ** 4))     # ransac.py:56-67
        p.ransac_initial_iterations = int(min(p.ransac_max_iterations, int(k0)))
        p.refine, p.bearing_threshold, p.tri_mode = int(bool(refine)), float(bearing_threshold), 1 if tri_opencv else 0
        self.params = p
        self.n_seq, self.capacity = int(n_seq), int(capacity)
        self._h = C.c_void_p()
        nat.check(nat.lib().vo_pipeline_create(self.ctx.handle, C.byref(p), C.byref(self._h)), "vo_pipeline_create")
        if rng_seed is not None:                       # ransac.py:52: every estimator starts from default_rng(2023)
            st = rng_state6(np.random.default_rng(rng_seed))
            for s in range(self.n_seq):
                self.write_table(s, rng=st)

    # -- frames ---------------------------------------------------------------------------------
    def _frames(self, frames):
        a = np.ascontiguousarray(frames, dtype=np.uint8)
        if a.ndim == 2:
            a = a[None]
        if a.shape != (self.n_seq, self.params.H, self.params.W):
            raise ValueError(f"frames must be ({self.n_seq}, {self.params.H}, {self.params.W}) uint8, got {a.shape}")
        return a

    def prime(self, frames, init_tables=True):
        a = self._frames(frames)
        nat.check(nat.lib().vo_pipeline_prime_host(self._h, nat.ptr(a), int(bool(init_tables))), "vo_pipeline_prime_host")

    def step(self, frames):
        """Advance every sequence by one frame; returns the summary (see summary_dict)."""
        a = self._frames(frames)
        out = np.empty((self.n_seq, SUMMARY_DOUBLES), dtype=np.float64)
        nat.check(nat.lib().vo_pipeline_step_host(self._h, nat.ptr(a), nat.ptr(out)), "vo_pipeline_step_host")
        return self.summary_dict(out)

    def bootstrap(self, frames, threshold=0.25, confidence=0.999, max_iters=1000):
        """main.py:203-231 on the device: track the primed frame's corners into `frames`, estimate the relative pose
        (cv2.findFundamentalMat's RANSAC restated, essential-matrix decomposition, cheirality vote), triangulate the
        inliers and apply the state updates; returns the summary (pose = the second camera)."""
        a = self._frames(frames)
        out = np.empty((self.n_seq, SUMMARY_DOUBLES), dtype=np.float64)
        nat.check(nat.lib().vo_pipeline_bootstrap_host(self._h, nat.ptr(a), float(threshold), float(confidence), int(max_iters),
                                                       nat.ptr(out)), "vo_pipeline_bootstrap_host")
        return self.summary_dict(out)

    @staticmethod
    def summary_dict(raw):
        raw = np.ascontiguousarray(raw)
        counts = raw[:, 12:].copy().view(np.int32).reshape(raw.shape[0], -1)[:, :NCOUNTS]
        d = {"pose": raw[:, :12].reshape(-1, 3, 4).copy(), "counts": counts}
        d.update({name: counts[:, i] for i, name in enumerate(COUNT_NAMES)})
        return d

    # pipelined host API (pinned numpy buffers recommended)
    def prefetch(self, frames):
        nat.check(nat.lib().vo_pipeline_prefetch_host(self._h, nat.ptr(frames)), "vo_pipeline_prefetch_host")

    def submit(self, frames, summary_out):
        nat.check(nat.lib().vo_pipeline_submit_host(self._h, nat.ptr(frames) if frames is not None else None,
                                                    nat.ptr(summary_out)), "vo_pipeline_submit_host")

    def wait(self):
        nat.check(nat.lib().vo_pipeline_wait_host(self._h), "vo_pipeline_wait_host")

    def prime_dev(self, d_frames, pitch, frame_stride, init_tables=True, stream=0):
        nat.check(nat.lib().vo_pipeline_prime_dev(self._h, d_frames, pitch, frame_stride, int(bool(init_tables)), stream or None),
                  "vo_pipeline_prime_dev")

    def step_dev(self, d_frames, pitch, frame_stride, stream=0):
        nat.check(nat.lib().vo_pipeline_step_dev(self._h, d_frames, pitch, frame_stride, stream or None), "vo_pipeline_step_dev")

    def sync_dev(self, stream=0):
        """Make `stream` wait for the detector / pose / update kernels the last steps left running on internal streams."""
        nat.check(nat.lib().vo_pipeline_sync_dev(self._h, stream or None), "vo_pipeline_sync_dev")

    def summary_dev(self) -> int:
        return int(nat.lib().vo_pipeline_summary_dev(self._h) or 0)

    # -- tables ---------------------------------------------------------------------------------
    def read_table(self, seq):
        """The sequence's table as numpy arrays shaped like the reference's Features columns."""
        Cn = self.capacity
        n = C.c_int()
        kp = np.empty((Cn, 2), np.float32); land = np.empty((Cn, 3)); state = np.empty(Cn, np.uint8)
        track = np.empty((Cn, 2), np.float32); pose = np.empty((Cn, 12)); cand = np.empty(Cn, np.uint8)
        c2w = np.empty(48); scal = np.empty(3 + NCOUNTS, np.int32); inl = np.zeros(Cn, np.uint8); rng = np.empty(6, np.uint64)
        nat.check(nat.lib().vo_pipeline_read_table_host(self._h, int(seq), C.byref(n), nat.ptr(kp), nat.ptr(land), nat.ptr(state),
                                                        nat.ptr(track), nat.ptr(pose), nat.ptr(cand), nat.ptr(c2w), nat.ptr(scal),
                                                        nat.ptr(inl), nat.ptr(rng)), "vo_pipeline_read_table_host")
        n = n.value
        pose4 = np.zeros((n, 4, 4)); pose4[:, :3, :] = pose[:n].reshape(n, 3, 4); pose4[:, 3, 3] = 1.0
        nanrow = np.isnan(pose[:n]).any(1)
        pose4[nanrow] = np.nan                                   # matches.py:193-206: NaN poses for triangulated rows
        def m4(v):
            T = np.eye(4); T[:3] = v.reshape(3, 4); return T
        def model(v):
            return v[:9].reshape(3, 3).copy(), v[9:].reshape(3, 1).copy()
        ntri = int(scal[2])
        return dict(n=n, kp=kp[:n].copy(), land=land[:n].copy(), state=state[:n].astype(int), track=track[:n].astype(np.float64),
                    pose=pose4, cand=cand[:n].astype(bool), curr_pose=m4(c2w[:12]), prev_pose=m4(c2w[12:24]),
                    w2c=model(c2w[24:36]), p3p_model=model(c2w[36:48]), num_features=int(scal[0]), n_iterations=int(scal[1]),
                    p3p_N=ntri, counts=dict(zip(COUNT_NAMES, scal[3:].tolist())), inliers=inl[:ntri].astype(bool), rng=rng)

    def write_table(self, seq, kp=None, land=None, state=None, track=None, pose=None, curr_pose=None, num_features=-1,
                    n_iterations=0, rng=None):
        """Replace the table of one sequence (columns as in read_table; pose (n, 4, 4) camera-to-world)."""
        if kp is None:
            n, a = -1, [None] * 5                                   # scalars only: the rows stay
        else:
            kp = np.ascontiguousarray(np.asarray(kp, dtype=np.float32).reshape(-1, 2))
            n = kp.shape[0]
            land = np.ascontiguousarray(np.asarray(land, dtype=np.float64).reshape(n, 3))
            state = np.ascontiguousarray(np.asarray(state).astype(np.uint8).reshape(n))
            track = np.ascontiguousarray(np.asarray(track, dtype=np.float32).reshape(n, 2))
            pose = np.ascontiguousarray(np.asarray(pose, dtype=np.float64).reshape(n, 4, 4)[:, :3, :].reshape(n, 12))
            a = [nat.ptr(kp), nat.ptr(land), nat.ptr(state), nat.ptr(track), nat.ptr(pose)]
        cw = None if curr_pose is None else np.ascontiguousarray(np.asarray(curr_pose, dtype=np.float64)[:3, :].reshape(12))
        rg = None if rng is None else np.ascontiguousarray(np.asarray(rng, dtype=np.uint64).reshape(6))
        nat.check(nat.lib().vo_pipeline_write_table_host(self._h, int(seq), n, *a, nat.ptr(cw) if cw is not None else None,
                                                         int(num_features), int(n_iterations), nat.ptr(rg) if rg is not None else None),
                  "vo_pipeline_write_table_host")

    def read_detections(self, seq):
        xy = np.empty((self.params.det_max_corners, 2), np.float32 if self.params.detector == DETECTOR_GFTT else np.int32)
        n = C.c_int()
        nat.check(nat.lib().vo_pipeline_read_detections_host(self._h, int(seq), nat.ptr(xy), C.byref(n)), "vo_pipeline_read_detections_host")
        return xy[: n.value].copy()

    def close(self):
        if self._h:
            nat.lib().vo_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pcg64_choice4(state6, N, n_draws, ctx=None):
    """Test hook (vo_test_pcg64_choice4_host): (samples int32 (n_draws, 4), advanced state6)."""
    ctx = ctx if ctx is not None else nat.default_context(0)
    st = np.ascontiguousarray(np.asarray(state6, dtype=np.uint64).copy())
    out = np.empty((n_draws, 4), np.int32)
    nat.check(nat.lib().vo_test_pcg64_choice4_host(ctx.handle, nat.ptr(st), int(N), int(n_draws), nat.ptr(out)), "vo_test_pcg64_choice4_host")
    return out, st
