"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement (plain C + numpy) of the reference's hot path, used solely as
the checker by tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  The product package
(``visual-odometry-project_b200/``) never imports this module.

Each wrapper names the reference lines it restates; see the C sources for the
arithmetic notes.  Parity pin: tests/golden/*.npz (made by
tests/golden/make_golden.py importing the reference itself in the build
container) -- checked by tests/test_oracle_golden.py.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        _lib = C.CDLL(path)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


# ----------------------------------------------------------------------------
# Harris  (reference: src/vo/features/harris.py:86-194)
# ----------------------------------------------------------------------------
def harris_response(img: np.ndarray, patch_size: int = 9, kappa: float = 0.09) -> np.ndarray:
    """float64 (H, W) score map of harris.py:102-137 (zero border of patch_size//2+1)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    out = np.empty((H, W), dtype=np.float64)
    rc = lib().oracle_harris_response(_p(img, C.c_uint8), H, W, W, int(patch_size),
                                      C.c_double(kappa), _p(out, C.c_double))
    if rc != 0:
        raise ValueError(f"oracle_harris_response rc={rc}")
    return out


def harris_nms(scores: np.ndarray, num_keypoints: int, radius: int) -> np.ndarray:
    """Greedy argmax/zero-box selection of harris.py:148-152 -> int32 (K, 2) as (x, y)."""
    s = np.array(scores, dtype=np.float64, order="C", copy=True)
    H, W = s.shape
    kp = np.empty((num_keypoints, 2), dtype=np.int32)
    lib().oracle_harris_nms(_p(s, C.c_double), H, W, int(radius), int(num_keypoints), _p(kp, C.c_int32))
    return kp


def harris_keypoints(img, num_keypoints=1000, patch_size=9, kappa=0.09, radius=5):
    resp = harris_response(img, patch_size, kappa)
    return harris_nms(resp, num_keypoints, radius), resp


def harris_descriptors(img: np.ndarray, kp_xy: np.ndarray, r: int = 9) -> np.ndarray:
    """harris.py:160-194 -> float64 (K, (2r+1)^2)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    kp = np.ascontiguousarray(kp_xy, dtype=np.int32)
    H, W = img.shape
    d = 2 * r + 1
    out = np.empty((kp.shape[0], d * d), dtype=np.float64)
    lib().oracle_harris_descriptors(_p(img, C.c_uint8), H, W, W, _p(kp, C.c_int32), kp.shape[0], int(r),
                                    _p(out, C.c_double))
    return out


def match_descriptors(desc1, desc2, ratio=0.85):
    """matchDescriptor of harris.py:196-264 for 8-bit descriptors: cv2.BFMatcher().knnMatch(k=2) with the L2 norm
    (exact integer squared distances, float32 sqrt), neighbours ordered by (distance, train index) as cv2's
    batchDistance does, the ratio test in double, and first-come uniqueness of the train index."""
    a = np.asarray(desc1).reshape(len(desc1), -1).astype(np.int64)
    b = np.asarray(desc2).reshape(len(desc2), -1).astype(np.int64)
    d2 = (a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2 * (a @ b.T)
    order = np.lexsort((np.broadcast_to(np.arange(b.shape[0]), d2.shape), d2), axis=1)[:, :2]
    used = np.zeros(b.shape[0], dtype=bool)
    good = []
    for q in range(a.shape[0]):
        m, n = order[q]
        dm = float(np.sqrt(np.float32(d2[q, m])))
        dn = float(np.sqrt(np.float32(d2[q, n])))
        if dm < ratio * dn and not used[m]:
            good.append([q, m])
            used[m] = True
    return np.array(good, dtype=np.int64).reshape(-1, 2)


# ----------------------------------------------------------------------------
# Shi-Tomasi corners  (reference: src/vo/features/klt.py:24-26, 87-115 -> cv2.goodFeaturesToTrack)
# ----------------------------------------------------------------------------
def min_eigen_val(img: np.ndarray, block: int = 7) -> np.ndarray:
    """cv2.cornerMinEigenVal(img, block, ksize=3) -> float32 (H, W)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    out = np.empty((H, W), dtype=np.float32)
    rc = lib().oracle_min_eigen_val(_p(img, C.c_uint8), H, W, W, int(block), _p(out, C.c_float))
    if rc != 0:
        raise ValueError(f"oracle_min_eigen_val rc={rc}")
    return out


def gftt_select(eig: np.ndarray, max_corners=500, quality=0.01, min_distance=8.0, return_candidates=False):
    """Thresholding, 3x3 local maxima, ordering and the minimum-distance greedy pass of cv2.goodFeaturesToTrack on a
    given eigenvalue map -> float32 (n, 2) corners as (x, y)."""
    e = np.ascontiguousarray(eig, dtype=np.float32)
    H, W = e.shape
    xy = np.empty((H * W // 2 + 64, 2), dtype=np.float32)
    nc = C.c_int(0)
    n = lib().oracle_gftt_select(_p(e, C.c_float), H, W, int(max_corners), C.c_double(quality), C.c_double(min_distance),
                                 _p(xy, C.c_float), C.byref(nc))
    if n < 0:
        raise ValueError(f"oracle_gftt_select rc={n}")
    return (xy[:n].copy(), nc.value) if return_candidates else xy[:n].copy()


def good_features_to_track(img, max_corners=500, quality=0.01, min_distance=8.0, block=7):
    """cv2.goodFeaturesToTrack(img, maxCorners, qualityLevel, minDistance, blockSize=block) -> float32 (n, 2)."""
    return gftt_select(min_eigen_val(img, block), max_corners, quality, min_distance)


# ----------------------------------------------------------------------------
# P3P + RANSAC  (reference: src/vo/pose_estimation/p3p.py, src/vo/algorithms/ransac.py)
# ----------------------------------------------------------------------------
def p3p_solve4(X4, uv4, K, return_all=False):
    """model_fn of p3p.py:51-79: pose (R 3x3, t 3x1) from 4 correspondences or None."""
    X4 = np.ascontiguousarray(np.asarray(X4, dtype=np.float64).reshape(4, 3))
    uv4 = np.ascontiguousarray(np.asarray(uv4, dtype=np.float64).reshape(4, 2))
    K = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
    m = np.zeros(12)
    allm = np.zeros((4, 12))
    n_all = C.c_int(0)
    ok = lib().oracle_p3p_solve4(_p(X4, C.c_double), _p(uv4, C.c_double), _p(K, C.c_double), _p(m, C.c_double),
                                 _p(allm, C.c_double), C.byref(n_all))
    if return_all:
        return [(allm[i, :9].reshape(3, 3).copy(), allm[i, 9:].reshape(3, 1).copy()) for i in range(n_all.value)]
    if not ok:
        return None
    return m[:9].reshape(3, 3).copy(), m[9:].reshape(3, 1).copy()


def reproj_errors(R, t, landmarks, keypoints, K):
    """error_fn of p3p.py:81-108: squared pixel reprojection error per correspondence."""
    m = np.concatenate([np.asarray(R, dtype=np.float64).reshape(9), np.asarray(t, dtype=np.float64).reshape(3)])
    L = np.ascontiguousarray(np.asarray(landmarks, dtype=np.float64).reshape(-1, 3))
    P = np.ascontiguousarray(np.asarray(keypoints, dtype=np.float64).reshape(-1, 2))
    K = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
    err = np.empty(L.shape[0])
    lib().oracle_reproj_errors(_p(m, C.c_double), _p(L, C.c_double), _p(P, C.c_double), L.shape[0],
                               _p(K, C.c_double), _p(err, C.c_double))
    return err


def p3p_ransac_score(landmarks, keypoints, K, sample_idx, threshold):
    """Models, validity and inlier counts for every sample-index set (ransac.py:92-106)."""
    L = np.ascontiguousarray(np.asarray(landmarks, dtype=np.float64).reshape(-1, 3))
    P = np.ascontiguousarray(np.asarray(keypoints, dtype=np.float64).reshape(-1, 2))
    K = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
    S = np.ascontiguousarray(sample_idx, dtype=np.int32).reshape(-1, 4)
    n = S.shape[0]
    models = np.empty((n, 12))
    valid = np.empty(n, dtype=np.uint8)
    counts = np.empty(n, dtype=np.int32)
    lib().oracle_p3p_ransac_score(_p(L, C.c_double), _p(P, C.c_double), L.shape[0], _p(K, C.c_double),
                                  _p(S, C.c_int32), n, C.c_double(threshold), _p(models, C.c_double),
                                  _p(valid, C.c_uint8), _p(counts, C.c_int32))
    return models, valid, counts


def ransac_iterations_table(N, s, confidence, max_iterations):
    """n_iterations as a function of best_n_inliers = 0..N (ransac.py:58-67 and 113-120).

    Evaluated with numpy *scalars*, one value at a time, exactly as the reference does, so the
    ceil() lands on the same integer."""
    out = np.empty(N + 1, dtype=np.int64)
    for c in range(N + 1):
        ratio = 1 - np.int64(c) / N
        ratio = min(max(ratio, 0.01), 0.99)
        k = np.ceil(np.log(1 - confidence) / np.log(1 - (1 - ratio) ** s))
        out[c] = int(min(max_iterations, int(k)))
    return out


def ransac_initial_iterations(s, outlier_ratio, confidence, max_iterations):
    """ransac.py:55-67."""
    k = np.ceil(np.log(1 - confidence) / np.log(1 - (1 - outlier_ratio) ** s))
    return min(max_iterations, int(k))


def ransac_scan(valid, counts, table, initial_iters, start_n=0, start_best=-1):
    """The sequential loop of ransac.py:90-121 over pre-scored hypotheses.

    Returns (best_h or -1, consumed, n_iterations_at_exit, n, best_count, exhausted)."""
    n, best, best_h = start_n, start_best, -1
    n_iter = initial_iters
    h = 0
    H = len(counts)
    while n < n_iter:
        if h >= H:
            return best_h, h, n_iter, n, best, True
        if not valid[h]:
            h += 1
            continue
        if counts[h] > best:
            best, best_h = int(counts[h]), h
            n_iter = int(table[best])
        n += 1
        h += 1
    return best_h, h, n_iter, n, best, False


# ----------------------------------------------------------------------------
# Triangulation  (reference: src/vo/landmarks/triangulation.py:352-389 and 38-86)
# ----------------------------------------------------------------------------
def _skew(p):
    """helpers.py:58-85 for homogeneous pixels (N, 3)."""
    S = np.zeros((p.shape[0], 3, 3))
    S[:, 0, 1] = -p[:, 2]
    S[:, 0, 2] = p[:, 1]
    S[:, 1, 0] = p[:, 2]
    S[:, 1, 2] = -p[:, 0]
    S[:, 2, 0] = -p[:, 1]
    S[:, 2, 1] = p[:, 0]
    return S


def triangulate(p1, p2, C1, C2, mode=0):
    """Linear triangulation.  p1, p2: (N, 2); C1: (3, 4) or (N, 3, 4); C2: (3, 4) -> (N, 3).

    mode 0: triangulation.py:379-387 -- A = [[p1]x C1; [p2]x C2] (6x4), numpy SVD, last right
            singular vector, dehomogenised (helpers.py:19-29).
    mode 1: cv2.triangulatePoints' system (4x4: x*P[2]-P[0], y*P[2]-P[1] per view), as used by
            triangulate_candidates with use_opencv=True (triangulation.py:59-74), same SVD."""
    p1 = np.asarray(p1, dtype=np.float64).reshape(-1, 2)
    p2 = np.asarray(p2, dtype=np.float64).reshape(-1, 2)
    N = p1.shape[0]
    C1 = np.broadcast_to(np.asarray(C1, dtype=np.float64), (N, 3, 4)) if np.ndim(C1) == 2 else np.asarray(C1, dtype=np.float64)
    C2 = np.broadcast_to(np.asarray(C2, dtype=np.float64), (N, 3, 4))
    if mode == 0:
        h1 = np.concatenate([p1, np.ones((N, 1))], axis=1)
        h2 = np.concatenate([p2, np.ones((N, 1))], axis=1)
        A = np.concatenate([_skew(h1) @ C1, _skew(h2) @ C2], axis=1)
    else:
        A = np.stack([p1[:, 0:1] * C1[:, 2] - C1[:, 0], p1[:, 1:2] * C1[:, 2] - C1[:, 1],
                      p2[:, 0:1] * C2[:, 2] - C2[:, 0], p2[:, 1:2] * C2[:, 2] - C2[:, 1]], axis=1)
    _, _, Vh = np.linalg.svd(A, full_matrices=False)
    P = Vh[:, -1, :]
    return P[:, :3] / P[:, 3:]


# ----------------------------------------------------------------------------
# KLT  (reference: src/vo/features/klt.py:233-249 -> cv2.calcOpticalFlowPyrLK)
# ----------------------------------------------------------------------------
def pyr_down(img: np.ndarray) -> np.ndarray:
    """cv2.pyrDown for uint8 single channel."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    out = np.empty(((H + 1) // 2, (W + 1) // 2), dtype=np.uint8)
    lib().oracle_pyr_down(_p(img, C.c_uint8), H, W, W, _p(out, C.c_uint8), out.shape[1])
    return out


def klt_track(prev, nxt, pts, win=17, max_level=2, max_iters=10, epsilon=0.03, min_eig=1e-4):
    """cv2.calcOpticalFlowPyrLK(prev, nxt, pts, None, winSize=(win, win), maxLevel=max_level,
    criteria=(EPS|COUNT, max_iters, epsilon)) -> (next_pts float32 (N, 2), status uint8 (N,), err float32 (N,))."""
    prev = np.ascontiguousarray(prev, dtype=np.uint8)
    nxt = np.ascontiguousarray(nxt, dtype=np.uint8)
    assert prev.shape == nxt.shape and prev.ndim == 2
    pts = np.ascontiguousarray(np.asarray(pts, dtype=np.float32).reshape(-1, 2))
    n = pts.shape[0]
    out = np.zeros((n, 2), dtype=np.float32)
    status = np.zeros(n, dtype=np.uint8)
    err = np.zeros(n, dtype=np.float32)
    H, W = prev.shape
    rc = lib().oracle_klt_track(_p(prev, C.c_uint8), _p(nxt, C.c_uint8), H, W, W, int(max_level), int(win),
                                int(max_iters), C.c_double(epsilon), C.c_double(min_eig), _p(pts, C.c_float), n,
                                _p(out, C.c_float), _p(status, C.c_uint8), _p(err, C.c_float))
    if rc < 0:
        raise ValueError("oracle_klt_track: bad arguments")
    return out, status, err


class RansacP3P:
    """Literal sequential restatement of RANSAC.find_best_model (ransac.py:69-129) with the P3P
    model_fn / error_fn of p3p.py:51-108, including the state the reference keeps between calls
    (rng stream seeded with 2023, outlier_ratio, n_iterations)."""

    def __init__(self, K, inlier_threshold, outlier_ratio=0.9, confidence=0.99, max_iterations=np.inf, inclusive=False):
        self.K = np.asarray(K, dtype=np.float64)
        self.inclusive = inclusive          # True: `<=` (cv2.solvePnPRansac's rule, see vo_b200.h), False: ransac.py:105
        self.s = 4
        self.inlier_threshold = inlier_threshold
        self.outlier_ratio = outlier_ratio
        self.confidence = confidence
        self.max_iterations = max_iterations
        self.rng = np.random.default_rng(2023)                      # ransac.py:52
        self.n_iterations = min(max_iterations, self._n_iter())     # ransac.py:56
        self.draws = 0

    def _n_iter(self):
        k = np.ceil(np.log(1 - self.confidence) / np.log(1 - (1 - self.outlier_ratio) ** self.s))
        return int(k)

    def find_best_model(self, landmarks, keypoints, draw_cap=None):
        L = np.asarray(landmarks, dtype=np.float64).reshape(-1, 3)
        P = np.asarray(keypoints, dtype=np.float64).reshape(-1, 2)
        N = L.shape[0]
        best_n, best_inl, best_model, n = -1, None, None, 0
        drawn = 0
        while n < self.n_iterations:
            if draw_cap is not None and drawn >= draw_cap:      # not in the reference (it would spin); see loop.py
                break
            idxs = self.rng.choice(np.arange(N), replace=False, size=self.s)
            self.draws += 1
            drawn += 1
            model = p3p_solve4(L[idxs], P[idxs], self.K)
            if model is None:
                continue
            e = reproj_errors(model[0], model[1], L, P, self.K)
            inl = (e <= self.inlier_threshold) if self.inclusive else (e < self.inlier_threshold)
            c = inl.sum()
            if c > best_n:
                best_n, best_inl, best_model = c, inl, model
                self.outlier_ratio = 1 - best_n / N
                self.outlier_ratio = min(max(self.outlier_ratio, 0.01), 0.99)
                self.n_iterations = int(min(self.max_iterations, self._n_iter()))
            n += 1
        return best_model, best_inl


# ----------------------------------------------------------------------------
# cv2.solvePnPRansac(flags=SOLVEPNP_P3P)  (reference: src/vo/pose_estimation/p3p.py:142-165, the use_opencv=True path
# that src/main.py takes).  OpenCV is a third-party dependency (not vendored in /root/reference); restated from
# calib3d/src/solvepnp.cpp (solvePnPRansac, PnPRansacCallback) and calib3d/src/ptsetreg.cpp
# (RANSACPointSetRegistrator::run / getSubset / findInliers, RANSACUpdateNumIters) and core's cv::RNG:
#   - object and image points are converted to float32 before anything else;
#   - every run starts from RNG(0xffffffffffffffff); a subset is 4 distinct draws of uniform(0, N) = next() % N,
#     next(): state = (uint32)state * 4164903690 + (state >> 32);
#   - the model is solvePnP(P3P) of the subset (three points solve, the fourth picks the root); a failed subset still
#     counts as an iteration;
#   - the error of a point is the squared distance between the float32 image point and the float32-rounded projection,
#     evaluated in float32; inliers are err <= (float)(reprojectionError^2);
#   - a model is kept when its inlier count exceeds max(best, 3); then niters = RANSACUpdateNumIters(confidence,
#     (N - good) / N, 4, niters) with cvRound;
# Pinned (tests/test_oracle_golden.py): inlier masks EQUAL to cv2.solvePnPRansac's on random problems and on the
# reference's own KITTI run (tests/golden/loop.npz cv_f3_inliers).
# ----------------------------------------------------------------------------
class CvRNG:
    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else a + self.next() % (b - a)


def cv_update_num_iters(p, ep, model_points, max_iters):
    """RANSACUpdateNumIters (ptsetreg.cpp)."""
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    tiny = float(np.finfo(np.float64).tiny)
    num = max(1.0 - p, tiny)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < tiny:
        return 0
    num, denom = float(np.log(num)), float(np.log(denom))
    return max_iters if (denom >= 0 or -num >= max_iters * (-denom)) else int(np.rint(num / denom))


def cv_subset4(rng: CvRNG, N: int):
    """RANSACPointSetRegistrator::getSubset for modelPoints = 4 (no degeneracy check for PnP)."""
    idx = []
    for _ in range(4):
        while True:
            v = rng.uniform(0, N)
            if v not in idx:
                break
        idx.append(v)
    return idx


def cv_reproj_errors_f32(R, t, L32, P32, K):
    """PnPRansacCallback::computeError: float32 projections, float32 squared distance."""
    cam = L32 @ np.asarray(R, dtype=np.float64).T + np.asarray(t, dtype=np.float64).reshape(3)
    with np.errstate(all="ignore"):
        iz = np.where(cam[:, 2] != 0, 1.0 / cam[:, 2], 1.0)          # cvProjectPoints2: z = z ? 1./z : 1; x *= z
        x = (cam[:, 0] * iz) * K[0, 0] + K[0, 2]
        y = (cam[:, 1] * iz) * K[1, 1] + K[1, 2]
    dx = P32[:, 0] - x.astype(np.float32)
    dy = P32[:, 1] - y.astype(np.float32)
    return dx * dx + dy * dy


def cv_solve_pnp_ransac_p3p(landmarks, keypoints, K, reproj_error, confidence, max_iters):
    """-> (best model (R, t) or None, inlier mask, iterations run)."""
    K = np.asarray(K, dtype=np.float64)
    L32 = np.asarray(landmarks, dtype=np.float64).reshape(-1, 3).astype(np.float32).astype(np.float64)
    P32 = np.asarray(keypoints).reshape(-1, 2).astype(np.float32)
    N = L32.shape[0]
    rng = CvRNG()
    niters, best, best_mask, best_model, it = int(max_iters), 0, np.zeros(N, bool), None, 0
    thr = np.float32(reproj_error * reproj_error)
    while it < niters:
        idx = cv_subset4(rng, N)
        it += 1
        m = p3p_solve4(L32[idx], P32[idx].astype(np.float64), K)
        if m is None:
            continue
        with np.errstate(invalid="ignore"):
            mask = cv_reproj_errors_f32(m[0], m[1], L32, P32, K) <= thr
        good = int(mask.sum())
        if good > max(best, 3):
            best, best_mask, best_model = good, mask, m
            niters = cv_update_num_iters(confidence, (N - good) / N, 4, niters)
    return best_model, best_mask, it
