"""ORACLE -- TEST INFRASTRUCTURE ONLY: the reference's per-frame loop on flat numpy arrays.

Restates, for the KLT tracker mode of the reference,
    src/main.py:248-287                  loop body (track -> P3P -> State updates -> triangulate)
    src/vo/features/klt.py:191-280       track_features (re-detection rule, status / error filter)
    src/vo/features/klt.py:117-189       update_features (appending fresh corners)
    src/vo/primitives/matches.py:10-212  Matches (stable regrouping by state, track / pose hand-over)
    src/vo/primitives/state.py           update_with_world_pose, reset_outliers, compute_candidates,
                                         update_with_world_landmarks, _check_landmarks
    src/vo/pose_estimation/p3p.py:123-213  estimate_pose + _nonlinear_refinement
    src/vo/landmarks/triangulation.py:38-86 triangulate_candidates
with one feature table per sequence (rows: keypoint, landmark, state, track start, start pose,
candidate flag).  The stages themselves are the oracle's (oracle.klt_track, oracle.RansacP3P,
oracle.triangulate); bookkeeping is numpy.

Pinned by tests/golden/loop.npz: the reference's own classes run headless over the six KITTI frames
of /root/reference/tests/test_data (tests/golden/make_golden.py::make_loop), table by table.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np

import oracle


# ----------------------------------------------------------------------------------------------
# pose refinement
# ----------------------------------------------------------------------------------------------
def refine_scipy(R, t, X, uv, K):
    """p3p.py:188-213 as written there: scipy.optimize.least_squares over the twist, residual =
    per-point reprojection distance, numeric Jacobian, expm / logm for the parametrisation."""
    from scipy.linalg import expm, logm
    from scipy.optimize import least_squares
    K = np.asarray(K, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64).reshape(-1, 3)
    uv = np.asarray(uv, dtype=np.float64).reshape(-1, 2)
    H = np.eye(4)
    H[:3, :3] = R
    H[:3, 3] = np.asarray(t).reshape(3)

    def to_H(tw):
        se = np.zeros((4, 4))
        w = tw[3:]
        se[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
        se[:3, 3] = tw[:3]
        return expm(se)

    def res(tw):
        Hg = to_H(tw)
        cam = X @ Hg[:3, :3].T + Hg[:3, 3]
        p = cam @ K.T
        return np.linalg.norm(uv - p[:, :2] / p[:, 2:], axis=1)

    se = logm(H)
    x0 = np.concatenate([se[:3, 3], [-se[1, 2], se[0, 2], -se[0, 1]]]).real
    Hg = to_H(least_squares(res, x0=x0).x)
    return Hg[:3, :3], Hg[:3, 3:]


GN_MAX_ITERS = 30
GN_STEP_TOL = 1e-11


def _so3_exp(w):
    th = np.sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2])
    Wx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + Wx
    return np.eye(3) + (np.sin(th) / th) * Wx + ((1 - np.cos(th)) / (th * th)) * (Wx @ Wx)


def _gn_system(R, t, X, uv, fx, fy, cx, cy):
    cam = X @ R.T + t
    iz = 1.0 / cam[:, 2]
    xn, yn = cam[:, 0] * iz, cam[:, 1] * iz
    ru = uv[:, 0] - (fx * xn + cx)
    rv = uv[:, 1] - (fy * yn + cy)
    # d(u, v) / d(delta): delta = (v, w), Xc' = Xc + w x Xc + v
    Ju = np.stack([fx * iz, np.zeros_like(iz), -fx * xn * iz,
                   -fx * xn * yn, fx * (1 + xn * xn), -fx * yn], 1)
    Jv = np.stack([np.zeros_like(iz), fy * iz, -fy * yn * iz,
                   -fy * (1 + yn * yn), fy * xn * yn, fy * xn], 1)
    Hm = Ju.T @ Ju + Jv.T @ Jv
    g = Ju.T @ ru + Jv.T @ rv
    return Hm, g, float(ru @ ru + rv @ rv)


def _cost(R, t, X, uv, fx, fy, cx, cy):
    cam = X @ R.T + t
    iz = 1.0 / cam[:, 2]
    ru = uv[:, 0] - (fx * cam[:, 0] * iz + cx)
    rv = uv[:, 1] - (fy * cam[:, 1] * iz + cy)
    return float(ru @ ru + rv @ rv)


def refine_gn(R, t, X, uv, K):
    """The refinement the CUDA pipeline runs in place of p3p.py:188-213: damped Gauss-Newton on SE(3)
    (left perturbation, analytic Jacobian of the 2N reprojection residuals).  It minimises the same
    cost as the reference's least_squares call (sum of squared reprojection distances), so both end
    in the same minimum; the reference stops at scipy's ftol = 1e-8, this one at a 1e-11 step."""
    K = np.asarray(K, dtype=np.float64)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    X = np.asarray(X, dtype=np.float64).reshape(-1, 3)
    uv = np.asarray(uv, dtype=np.float64).reshape(-1, 2)
    R = np.array(R, dtype=np.float64).reshape(3, 3)
    t = np.array(t, dtype=np.float64).reshape(3)
    lam = 0.0
    Hm, g, c = _gn_system(R, t, X, uv, fx, fy, cx, cy)
    for _ in range(GN_MAX_ITERS):
        A = Hm + lam * np.diag(np.diag(Hm))
        try:
            d = np.linalg.solve(A, g)
        except np.linalg.LinAlgError:
            break
        if not np.all(np.isfinite(d)):
            break
        E = _so3_exp(d[3:])
        R2, t2 = E @ R, E @ t + d[:3]
        c2 = _cost(R2, t2, X, uv, fx, fy, cx, cy)
        if c2 <= c:
            R, t = R2, t2
            lam = lam * 0.1 if lam > 1e-9 else 0.0
            Hm, g, c = _gn_system(R, t, X, uv, fx, fy, cx, cy)
            if np.max(np.abs(d)) < GN_STEP_TOL:
                break
        else:
            lam = 1e-4 if lam == 0.0 else lam * 10.0
            if lam > 1e8:
                break
    return R, t.reshape(3, 1)


# ----------------------------------------------------------------------------------------------
# the loop
# ----------------------------------------------------------------------------------------------
def _inv_pose(T):
    return np.linalg.inv(T)


class LoopOracle:
    """One sequence.  Table columns (row i = feature i):
        kp        float32 (n, 2)   keypoint in the current frame (klt.py:241 next_pts)
        land      float64 (n, 3)   world landmark or NaN
        state     int     (n,)     0 unmatched / 1 matched / 2 triangulated (features.py:41-43)
        track     float64 (n, 2)   keypoint at the start of the track (NaN once triangulated)
        pose      float64 (n, 4, 4) camera-to-world pose at the start of the track
        cand      bool    (n,)     candidate for triangulation in this frame
    """

    def __init__(self, K, detector, *, p3p_opencv=False, refine="scipy", tri_opencv=True, inlier_threshold=1.25,
                 outlier_ratio=0.9, confidence=0.9999, max_iterations=10000, bearing_threshold=0.0075,
                 error_threshold=100.0, redetect_fraction=0.8, klt=None, klt_params=None, eager_detector=False, capacity=None):
        self.K = np.asarray(K)                               # dtype kept: the reference's KITTI K is float32 (loader.py:94-96)
        self.Kinv = np.linalg.inv(self.K)                    # camera.py:92 (float32 inverse when K is float32)
        self.K64 = np.asarray(K, dtype=np.float64)
        self.detector = detector                             # callable(gray) -> float32 (m, 2) corners
        self.p3p_opencv = p3p_opencv
        self.refine = refine
        self.tri_opencv = tri_opencv
        self.inlier_threshold = inlier_threshold
        self.bearing_threshold = bearing_threshold
        self.error_threshold = error_threshold
        self.redetect_fraction = redetect_fraction
        thr = inlier_threshold ** 2 if p3p_opencv else inlier_threshold        # p3p.py:149 vs ransac.py:105
        self.ransac = oracle.RansacP3P(self.K64, thr, outlier_ratio, confidence, max_iterations, inclusive=p3p_opencv)
        self.klt = klt or oracle.klt_track
        self.klt_params = klt_params or dict(win=17, max_level=2, max_iters=10, epsilon=0.03)   # klt.py:29-33
        # eager_detector: run the detector on every new frame and keep its corners for the next step's re-detection
        # (what the CUDA pipeline does; the result of a step is the same, klt.py:207-230 detects on that very frame)
        self.capacity = capacity            # rows the CUDA pipeline's table can hold (an append is cut to the room left)
        self.eager_detector = eager_detector
        self._cached_det = None
        self.curr_pose = np.eye(4)
        self.prev_pose = None
        self.num_features = None
        self.kp = self.land = self.state = self.track = self.pose = self.cand = None

    # -- table ------------------------------------------------------------------------------
    def init_detect(self, gray):
        """KLTTracker.__init__ (klt.py:40-50): fresh corners, everything unmatched."""
        kp = np.asarray(self.detector(gray), dtype=np.float32).reshape(-1, 2)
        self.num_features = kp.shape[0]
        self.set_table(kp, np.full((len(kp), 3), np.nan), np.zeros(len(kp), int), kp.astype(np.float64),
                       np.stack([np.eye(4)] * len(kp)) if len(kp) else np.empty((0, 4, 4)))

    def set_table(self, kp, land, state, track, pose, cand=None, curr_pose=None, prev_pose=None, num_features=None):
        self.kp = np.array(kp, dtype=np.float32).reshape(-1, 2)
        n = len(self.kp)
        self.land = np.array(land, dtype=np.float64).reshape(n, 3)
        self.state = np.array(state).astype(int).reshape(n)
        self.track = np.array(track, dtype=np.float64).reshape(n, 2)
        self.pose = np.array(pose, dtype=np.float64).reshape(n, 4, 4)
        self.cand = np.zeros(n, bool) if cand is None else np.array(cand, dtype=bool).reshape(n)
        if curr_pose is not None:
            self.curr_pose = np.array(curr_pose, dtype=np.float64)
        if prev_pose is not None:
            self.prev_pose = np.array(prev_pose, dtype=np.float64)
        if num_features is not None:
            self.num_features = int(num_features)

    def table(self):
        return dict(kp=self.kp.copy(), land=self.land.copy(), state=self.state.copy(), track=self.track.copy(),
                    pose=self.pose.copy(), cand=self.cand.copy(), curr_pose=self.curr_pose.copy())

    # -- one frame --------------------------------------------------------------------------
    def step(self, prev_gray, new_gray):
        info = {}
        # klt.py:207-230: too few features left -> detect on the OLD frame and append (all-255 mask: duplicates allowed)
        n = len(self.kp)
        info["redetect"] = bool(n < self.num_features * self.redetect_fraction)
        if info["redetect"]:
            fresh = self._cached_det if self._cached_det is not None else self.detector(prev_gray)
            fresh = np.asarray(fresh, dtype=np.float32).reshape(-1, 2)
            self.num_features = fresh.shape[0]                              # klt.py:114 (find_corners side effect)
            info["overflow"] = self.capacity is not None and n + len(fresh) > self.capacity
            if info["overflow"]:
                fresh = fresh[: self.capacity - n]
            m = len(fresh)
            self.kp = np.concatenate([self.kp, fresh])
            self.land = np.concatenate([self.land, np.full((m, 3), np.nan)])
            self.state = np.concatenate([self.state, np.zeros(m, int)])
            self.track = np.concatenate([self.track, fresh.astype(np.float64)])
            self.pose = np.concatenate([self.pose, np.stack([np.eye(4)] * m) if m else np.empty((0, 4, 4))])
            self.cand = np.concatenate([self.cand, np.zeros(m, bool)])
        # klt.py:233-249
        nxt, status, err = self.klt(prev_gray, new_gray, self.kp, **self.klt_params)
        keep = status.astype(bool) & (err < np.float32(self.error_threshold))
        info["n_tracked"], info["n_kept"] = len(self.kp), int(keep.sum())
        old_kp, old_state = self.kp[keep], self.state[keep]
        old_land, old_track, old_pose = self.land[keep], self.track[keep], self.pose[keep]
        nxt = nxt[keep]
        # matches.py:31-38 with identity pairs: stable regrouping [state 2 | state 1 | state 0]
        order = np.concatenate([np.flatnonzero(old_state == 2), np.flatnonzero(old_state == 1), np.flatnonzero(old_state == 0)])
        n2, n1, n0 = int((old_state == 2).sum()), int((old_state == 1).sum()), int((old_state == 0).sum())
        kp = nxt[order]
        land = np.concatenate([old_land[order[:n2]], np.full((n1 + n0, 3), np.nan)])       # matches.py:152-164
        state = np.concatenate([np.full(n2, 2), np.ones(n1 + n0, int)])                     # matches.py:166-173
        track = np.concatenate([np.full((n2, 2), np.nan), old_track[order[n2:n2 + n1]],     # matches.py:175-191 / 83-90
                                old_kp[order[n2 + n1:]].astype(np.float64)])
        pose = np.concatenate([np.full((n2, 4, 4), np.nan), old_pose[order[n2:]]])          # matches.py:193-206
        cand = np.zeros(len(kp), bool)                                                       # fresh Features (klt.py:256)
        # main.py:254-262: P3P on the triangulated rows
        N = n2
        info["p3p_N"] = N
        # the reference raises when fewer than 4 landmarks are left (Generator.choice) and loops forever when no sample
        # yields a model; the CUDA pipeline reports "no pose" and keeps the last one -- restated here for those cases
        model, inl = (None, np.zeros(N, bool))
        if N >= 4 and self.p3p_opencv:
            # p3p.py:142-151: cv2.solvePnPRansac (its own generator, restarted at every call; nothing carried over)
            model, inl, it = oracle.cv_solve_pnp_ransac_p3p(land[:N], kp[:N], self.K64, self.inlier_threshold,
                                                            self.ransac.confidence, self.ransac.max_iterations)
            self.ransac.draws += it
            info["cv_iterations"] = it
        elif N >= 4:
            model, inl = self.ransac.find_best_model(land[:N], kp[:N].astype(np.float64),
                                                     draw_cap=10 * min(self.ransac.max_iterations, 1 << 24) + 4096)
        info["ransac_n_iterations"] = self.ransac.n_iterations
        info["ransac_draws"] = self.ransac.draws
        info["no_pose"] = model is None
        self.prev_pose = self.curr_pose
        if model is None:
            inl = np.zeros(N, bool)
            info["inliers"] = inl
            outliers = np.zeros(len(kp), bool)
            info["p3p_R"], info["p3p_t"] = np.linalg.inv(self.curr_pose)[:3, :3], np.linalg.inv(self.curr_pose)[:3, 3]
        else:
            R, t = model
            info["p3p_R"], info["p3p_t"] = np.array(R), np.array(t).reshape(3)
            if self.refine == "scipy":
                R, t = refine_scipy(R, t, land[:N][inl], kp[:N][inl], self.K)
            elif self.refine == "gn":
                R, t = refine_gn(R, t, land[:N][inl], kp[:N][inl], self.K)
            info["inliers"] = inl.copy()
            outliers = np.zeros(len(kp), bool)
            outliers[:N] = ~inl                                                              # main.py:264-265
            # state.py:19-23, 39-51
            Tcw = np.concatenate([np.concatenate([np.asarray(R), np.asarray(t).reshape(3, 1)], 1), [[0, 0, 0, 1]]], 0)
            self.curr_pose = np.linalg.inv(Tcw)
        pose[state == 0] = self.curr_pose                                                    # features.py:228 (no such rows here)

        def reset(mask):                                                                     # state.py:167-178
            state[mask] = 0
            track[mask] = kp[mask]
            pose[mask] = self.curr_pose
        reset(outliers)
        # state.py:139-165 + 180-229
        sel = (state == 1) & (model is not None)
        ends = np.concatenate([kp[sel].astype(np.float64), np.ones((int(sel.sum()), 1))], 1)
        starts = np.concatenate([track[sel], np.ones((int(sel.sum()), 1))], 1)
        d1 = np.einsum("nij,nj->ni", pose[sel][:, :3, :3], starts @ self.Kinv.T)
        d2 = (ends @ self.Kinv.T) @ self.curr_pose[:3, :3].T
        with np.errstate(invalid="ignore"):
            ang = np.arccos(np.sum(d1 * d2, -1) / (np.linalg.norm(d1, axis=-1) * np.linalg.norm(d2, axis=-1)))
        info["angles"] = ang
        cand[sel] = ang >= self.bearing_threshold
        info["n_candidates"] = int(cand.sum())
        # main.py:276-284
        if cand.sum() > 0:
            proj1 = self.K @ np.linalg.inv(pose[cand])[:, :3]                                # triangulation.py:53-57
            proj2 = self.K @ np.linalg.inv(self.curr_pose)[:3]
            Xw = oracle.triangulate(track[cand], kp[cand].astype(np.float64), proj1, proj2, mode=1 if self.tri_opencv else 0)
            info["new_landmarks"] = Xw
            land[cand] = Xw                                                                  # state.py:83-84
            state[cand] = 2
            hom = np.concatenate([land, np.ones((len(land), 1))], 1)                         # state.py:92-110
            with np.errstate(invalid="ignore"):
                zc = (hom @ np.linalg.inv(self.curr_pose).T)
                zp = (hom @ np.linalg.inv(self.prev_pose).T)
                behind = (zc[:, 2] / zc[:, 3] < 0) | (zp[:, 2] / zp[:, 3] < 0)
            land[behind] = np.nan
            reset(behind)
            info["n_behind"] = int(behind.sum())
        self.kp, self.land, self.state, self.track, self.pose, self.cand = kp, land, state, track, pose, cand
        self._cached_det = self.detector(new_gray) if self.eager_detector else None
        info["n"] = len(kp)
        info["n_tri"] = int((state == 2).sum())
        return info
