"""numpy-in / numpy-out wrappers over the host-buffer C ABI (include/vo_b200.h).

These are what the `vo` classes call.  Every function runs hand-written CUDA on the GPU through
libvo_b200.so; nothing here computes on the CPU beyond shaping arguments.
"""
import ctypes as C

import numpy as np

from . import _native as nat


def _ctx(ctx=None):
    return ctx if ctx is not None else nat.default_context(0)


# ------------------------------------------------------------------------------------------------
# Harris  (reference: src/vo/features/harris.py)
# ------------------------------------------------------------------------------------------------
def harris_detect(img, num_keypoints=1000, patch_size=9, kappa=0.09, nms_radius=5, desc_radius=None,
                  want_response=False, ctx=None):
    """extractKeypoints (+ optionally extractDescriptors) for one frame (H, W) or a batch (F, H, W).

    Returns (kp_xy int32 [..., K, 2], response float64 [..., H, W] | None, desc uint8 [..., K, D] | None).
    """
    ctx = _ctx(ctx)
    a = np.ascontiguousarray(img, dtype=np.uint8)
    single = a.ndim == 2
    if single:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("harris_detect: image must be (H, W) or (F, H, W) uint8 grayscale")
    F, H, W = a.shape
    kp = np.empty((F, num_keypoints, 2), dtype=np.int32)
    resp = np.empty((F, H, W), dtype=np.float64) if want_response else None
    desc = None
    dr = 0
    if desc_radius is not None:
        dr = int(desc_radius)
        desc = np.empty((F, num_keypoints, (2 * dr + 1) ** 2), dtype=np.uint8)
    rc = nat.lib().vo_harris_detect_host(
        ctx.handle, nat.ptr(a), F, H, W, int(patch_size), C.c_double(kappa), int(nms_radius), int(num_keypoints),
        dr, nat.ptr(resp) if resp is not None else None, nat.ptr(kp), nat.ptr(desc) if desc is not None else None)
    nat.check(rc, "vo_harris_detect_host")
    if single:
        return kp[0], (resp[0] if resp is not None else None), (desc[0] if desc is not None else None)
    return kp, resp, desc


# ------------------------------------------------------------------------------------------------
# Shi-Tomasi corners  (reference: src/vo/features/klt.py:24-26, 87-115 -> cv2.goodFeaturesToTrack)
# ------------------------------------------------------------------------------------------------
def good_features_to_track(img, max_corners=500, quality_level=0.01, min_distance=8, block_size=7, want_eig=False,
                           want_stats=False, ctx=None):
    """cv2.goodFeaturesToTrack for one gray frame (H, W) or a batch (F, H, W) on the GPU.

    Returns float32 corners (n, 2) = (x, y) in OpenCV's order (a list of such arrays for a batch), plus the
    eigenvalue map float32 (cv2.cornerMinEigenVal) and / or the per-frame statistics when asked for."""
    ctx = _ctx(ctx)
    a = np.ascontiguousarray(img, dtype=np.uint8)
    single = a.ndim == 2
    if single:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("good_features_to_track: image must be (H, W) or (F, H, W) uint8 grayscale")
    F, H, W = a.shape
    xy = np.zeros((F, max_corners, 2), np.float32)
    n = np.zeros(F, np.int32)
    eig = np.empty((F, H, W), np.float32) if want_eig else None
    st = np.zeros((F, 4), np.uint32)
    rc = nat.lib().vo_gftt_host(ctx.handle, nat.ptr(a), F, H, W, int(max_corners), C.c_double(quality_level),
                                C.c_double(min_distance), int(block_size), nat.ptr(eig) if want_eig else None, nat.ptr(xy),
                                nat.ptr(n), nat.ptr(st))
    nat.check(rc, "vo_gftt_host")
    corners = [xy[f, : n[f]].copy() for f in range(F)]
    out = [corners[0] if single else corners]
    if want_eig:
        out.append(eig[0] if single else eig)
    if want_stats:
        out.append(st[0] if single else st)
    return out[0] if len(out) == 1 else tuple(out)


# ------------------------------------------------------------------------------------------------
# KLT  (reference: src/vo/features/klt.py:233-239 -> cv2.calcOpticalFlowPyrLK)
# ------------------------------------------------------------------------------------------------
def klt_track(prev, nxt, pts, win=17, max_level=2, max_iters=10, epsilon=0.03, min_eig=1e-4, ctx=None):
    """Pyramidal LK for one frame pair (H, W) / (H, W, 3) BGR or a batch (F, H, W) / (F, H, W, 3); pts (N, 2) or
    (F, N, 2) float32.  BGR input is converted on the device (cv2.cvtColor's arithmetic, klt.py:57-62).

    Returns (next_pts float32, status uint8, err float32) shaped like cv2's outputs (without the
    middle singleton axis)."""
    ctx = _ctx(ctx)
    a = np.ascontiguousarray(prev, dtype=np.uint8)
    b = np.ascontiguousarray(nxt, dtype=np.uint8)
    if a.shape != b.shape:
        raise ValueError("klt_track: prev and next images must have the same shape")
    bgr = a.shape[-1] == 3 and a.ndim in (3, 4) and not (a.ndim == 3 and np.asarray(pts).ndim == 3)
    single = a.ndim == (3 if bgr else 2)
    if single:
        a, b = a[None], b[None]
    F, H, W = a.shape[:3]
    p = np.ascontiguousarray(np.asarray(pts, dtype=np.float32).reshape(F, -1, 2))
    n = p.shape[1]
    out = np.zeros((F, n, 2), dtype=np.float32)
    status = np.zeros((F, n), dtype=np.uint8)
    err = np.zeros((F, n), dtype=np.float32)
    fn = nat.lib().vo_klt_track_bgr_host if bgr else nat.lib().vo_klt_track_host
    rc = fn(ctx.handle, nat.ptr(a), nat.ptr(b), F, H, W, int(max_level), int(win), int(max_iters), C.c_double(epsilon),
            C.c_double(min_eig), nat.ptr(p), n, nat.ptr(out), nat.ptr(status), nat.ptr(err))
    nat.check(rc, "vo_klt_track_host")
    if single:
        return out[0], status[0], err[0]
    return out, status, err


def bgr2gray(img, ctx=None):
    """cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) for uint8 (H, W, 3) or (F, H, W, 3), on the device, bit-exact."""
    ctx = _ctx(ctx)
    a = np.ascontiguousarray(img, dtype=np.uint8)
    single = a.ndim == 3
    if single:
        a = a[None]
    if a.ndim != 4 or a.shape[-1] != 3:
        raise ValueError("bgr2gray: image must be (H, W, 3) or (F, H, W, 3) uint8")
    F, H, W = a.shape[:3]
    out = np.empty((F, H, W), np.uint8)
    nat.check(nat.lib().vo_bgr2gray_host(ctx.handle, nat.ptr(a), F, H, W, nat.ptr(out)), "vo_bgr2gray_host")
    return out[0] if single else out


# ------------------------------------------------------------------------------------------------
# P3P + RANSAC  (reference: src/vo/pose_estimation/p3p.py, src/vo/algorithms/ransac.py)
# ------------------------------------------------------------------------------------------------
def p3p_ransac(landmarks, keypoints, K, sample_idx, threshold, iters_for_count, initial_iters, start_n=0,
               start_best=-1, want_all=False, inclusive=False, ctx=None):
    """Score every 4-index sample set and replay the reference's adaptive loop over them.

    landmarks (N, 3) / keypoints (N, 2) or batched (F, N, .); sample_idx (H, 4) or (F, H, 4) int32.
    inclusive: count squared errors equal to the threshold as inliers (cv2.solvePnPRansac's rule).
    Returns a dict with best (index or -1), best_count, n, exhausted, consumed, n_iterations,
    inliers (bool), R, t and, with want_all, counts / valid / models for every hypothesis."""
    ctx = _ctx(ctx)
    L = np.asarray(landmarks, dtype=np.float64)
    single = L.ndim == 2 or (L.ndim == 3 and L.shape[-1] == 1)
    L = np.ascontiguousarray(L.reshape((1, -1, 3)) if single else L.reshape(L.shape[0], -1, 3))
    F, N = L.shape[0], L.shape[1]
    P = np.ascontiguousarray(np.asarray(keypoints, dtype=np.float64).reshape(F, N, 2))
    S = np.ascontiguousarray(np.asarray(sample_idx, dtype=np.int32).reshape(F, -1, 4))
    Hn = S.shape[1]
    K9 = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
    i32max = np.iinfo(np.int32).max
    table = np.ascontiguousarray(np.minimum(np.asarray(iters_for_count).reshape(-1), i32max).astype(np.int32))
    initial_iters = int(min(initial_iters, i32max))
    if table.shape[0] != N + 1:
        raise ValueError("p3p_ransac: iters_for_count must have N + 1 entries")
    best4 = np.zeros((F, 4), dtype=np.int32)
    consumed = np.zeros(F, dtype=np.int32)
    iters_out = np.zeros(F, dtype=np.int32)
    inl = np.zeros((F, N), dtype=np.uint8)
    bm = np.zeros((F, 12), dtype=np.float64)
    counts = np.zeros((F, Hn), dtype=np.int32) if want_all else None
    valid = np.zeros((F, Hn), dtype=np.uint8) if want_all else None
    models = np.zeros((F, Hn, 12), dtype=np.float64) if want_all else None
    rc = nat.lib().vo_p3p_ransac_host(
        ctx.handle, nat.ptr(L), nat.ptr(P), F, N, nat.ptr(K9), nat.ptr(S), Hn, C.c_double(threshold), int(bool(inclusive)),
        nat.ptr(table),
        int(initial_iters), int(start_n), int(start_best), nat.ptr(best4), nat.ptr(consumed), nat.ptr(iters_out),
        nat.ptr(inl), nat.ptr(bm), nat.ptr(counts) if want_all else None, nat.ptr(valid) if want_all else None,
        nat.ptr(models) if want_all else None)
    nat.check(rc, "vo_p3p_ransac_host")
    res = dict(best=best4[:, 0], best_count=best4[:, 1], n=best4[:, 2], exhausted=best4[:, 3].astype(bool),
               consumed=consumed, n_iterations=iters_out, inliers=inl.astype(bool), R=bm[:, :9].reshape(F, 3, 3),
               t=bm[:, 9:].reshape(F, 3, 1), counts=counts, valid=valid, models=models)
    if single:
        res = {k: (v[0] if v is not None else None) for k, v in res.items()}
    return res


def refine_pose(landmarks, keypoints, K, R, t, mask=None, ctx=None):
    """_nonlinear_refinement (p3p.py:188-213) on the GPU: (R 3x3, t 3x1) minimising the reprojection error of the
    (masked) correspondences, started at (R, t).  Also returns the number of Gauss-Newton steps."""
    ctx = _ctx(ctx)
    L = np.ascontiguousarray(np.asarray(landmarks, dtype=np.float64).reshape(-1, 3))
    P = np.ascontiguousarray(np.asarray(keypoints, dtype=np.float64).reshape(-1, 2))
    N = L.shape[0]
    if P.shape[0] != N:
        raise ValueError("refine_pose: landmarks and keypoints differ in length")
    K9 = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
    pin = np.ascontiguousarray(np.concatenate([np.asarray(R, dtype=np.float64).reshape(9), np.asarray(t, dtype=np.float64).reshape(3)]))
    pout = np.empty(12)
    it = np.zeros(1, np.int32)
    m = None if mask is None else np.ascontiguousarray(np.asarray(mask).astype(np.uint8).reshape(N))
    rc = nat.lib().vo_refine_pose_host(ctx.handle, nat.ptr(L), nat.ptr(P), nat.ptr(m) if m is not None else None, 1, N, nat.ptr(K9),
                                       nat.ptr(pin), nat.ptr(pout), nat.ptr(it))
    nat.check(rc, "vo_refine_pose_host")
    return pout[:9].reshape(3, 3).copy(), pout[9:].reshape(3, 1).copy(), int(it[0])


# ------------------------------------------------------------------------------------------------
# Triangulation  (reference: src/vo/landmarks/triangulation.py:352-389, 38-86)
# ------------------------------------------------------------------------------------------------
def triangulate(p1, p2, C1, C2, mode=0, ctx=None):
    """DLT triangulation: p1, p2 (N, 2[, 1]); C1 (3, 4) or (N, 3, 4); C2 (3, 4) -> (N, 3) float64."""
    ctx = _ctx(ctx)
    a = np.ascontiguousarray(np.asarray(p1, dtype=np.float64).reshape(-1, 2))
    b = np.ascontiguousarray(np.asarray(p2, dtype=np.float64).reshape(-1, 2))
    if a.shape != b.shape:
        raise ValueError("triangulate: input points dimension mismatch")
    n = a.shape[0]
    c1 = np.asarray(C1, dtype=np.float64)
    per_point = 1 if c1.ndim == 3 else 0
    if per_point and c1.shape[0] != n:
        raise ValueError("triangulate: per-point projection matrices must match the number of points")
    c1 = np.ascontiguousarray(c1.reshape(-1, 12))
    c2 = np.ascontiguousarray(np.asarray(C2, dtype=np.float64).reshape(12))
    out = np.empty((n, 3), dtype=np.float64)
    if n == 0:
        return out
    rc = nat.lib().vo_triangulate_host(ctx.handle, nat.ptr(a), nat.ptr(b), n, nat.ptr(c1), per_point, nat.ptr(c2),
                                       int(mode), nat.ptr(out))
    nat.check(rc, "vo_triangulate_host")
    return out


def bootstrap(points1, points2, K, threshold, confidence, max_iters=1000, ctx=None):
    """Two-view bootstrap of triangulation.py:88-350 (use_ransac=True, use_opencv=True) on the GPU: cv2.findFundamentalMat's
    RANSAC, essential-matrix decomposition, cheirality vote, landmarks of all matches.
    points (N, 2[, 1]) -> dict(found, F (3, 3), M (3, 4), landmarks (N, 3), mask (N,), f_mask (N,), iterations)."""
    ctx = _ctx(ctx)
    a = np.ascontiguousarray(np.asarray(points1, dtype=np.float64).reshape(-1, 2))
    b = np.ascontiguousarray(np.asarray(points2, dtype=np.float64).reshape(-1, 2))
    if a.shape != b.shape:
        raise ValueError("bootstrap: input points dimension mismatch")
    n = a.shape[0]
    if n < 15:
        raise ValueError("bootstrap: cv2.findFundamentalMat(FM_RANSAC) needs at least 15 matches for its RANSAC path "
                         f"(got {n}); the GPU path restates that one only")
    K9 = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
    F, M = np.empty(9, np.float64), np.empty(12, np.float64)
    land = np.empty((n, 3), np.float64)
    mask, f_mask = np.empty(n, np.uint8), np.empty(n, np.uint8)
    info = np.zeros(4, np.int32)
    rc = nat.lib().vo_bootstrap_host(ctx.handle, nat.ptr(a), nat.ptr(b), 1, n, None, nat.ptr(K9), float(threshold),
                                     float(confidence), int(max_iters), nat.ptr(F), nat.ptr(M), nat.ptr(land),
                                     nat.ptr(mask), nat.ptr(f_mask), nat.ptr(info))
    nat.check(rc, "vo_bootstrap_host")
    return {"found": bool(info[0]), "F": F.reshape(3, 3), "M": M.reshape(3, 4), "landmarks": land,
            "mask": mask.astype(bool), "f_mask": f_mask.astype(bool), "iterations": int(info[1]),
            "n_f_inliers": int(info[2]), "n_valid": int(info[3])}


def harris_descriptors(img, kp_xy, desc_radius=9, ctx=None):
    """extractDescriptors alone (harris.py:160-194): uint8 (K, (2r+1)^2) patches for given keypoints."""
    ctx = _ctx(ctx)
    a = np.ascontiguousarray(img, dtype=np.uint8)
    kp = np.ascontiguousarray(kp_xy, dtype=np.int32).reshape(-1, 2)
    H, W = a.shape
    d = 2 * int(desc_radius) + 1
    out = np.empty((kp.shape[0], d * d), dtype=np.uint8)
    if kp.shape[0] == 0:
        return out
    rc = nat.lib().vo_harris_descriptors_host(ctx.handle, nat.ptr(a), H, W, nat.ptr(kp), kp.shape[0], int(desc_radius),
                                              nat.ptr(out))
    nat.check(rc, "vo_harris_descriptors_host")
    return out


def match_descriptors(desc1, desc2, ratio=0.85, ctx=None):
    """matchDescriptor for 8-bit descriptors (harris.py:196-264): (M, 2) int array of (query, train) pairs."""
    ctx = _ctx(ctx)
    a = np.ascontiguousarray(desc1, dtype=np.uint8)
    b = np.ascontiguousarray(desc2, dtype=np.uint8)
    a = a.reshape(a.shape[0], -1)
    b = b.reshape(b.shape[0], -1)
    if a.shape[1] != b.shape[1]:
        raise ValueError("match_descriptors: descriptor lengths differ")
    Q, T, D = a.shape[0], b.shape[0], a.shape[1]
    pairs = np.zeros((Q, 2), dtype=np.int32)
    n = np.zeros(1, dtype=np.int32)
    rc = nat.lib().vo_match_descriptors_host(ctx.handle, nat.ptr(a), nat.ptr(b), 1, Q, T, D, C.c_double(ratio),
                                             nat.ptr(pairs), nat.ptr(n))
    nat.check(rc, "vo_match_descriptors_host")
    return pairs[: int(n[0])].astype(np.int64)
