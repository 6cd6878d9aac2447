"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Two-view bootstrap of the reference: src/vo/landmarks/triangulation.py:88-350 as src/main.py:185-222 runs it
(``use_ransac=True, use_opencv=True``): ``cv2.findFundamentalMat(FM_RANSAC)`` -> essential matrix -> four
[R | t] candidates -> cheirality vote by DLT triangulation -> landmarks of all points.

``cv2.findFundamentalMat`` is a third-party dependency (pinned ``opencv-python==4.8.1.78`` by the reference,
4.13.0 in this image); its published algorithm (modules/calib3d/src/fundam.cpp, ptsetreg.cpp) is restated here:
  - points converted to float32; RANSACPointSetRegistrator with modelPoints = 7, cv::RNG seeded with 2^64-1 per call;
  - getSubset: 7 distinct ``rng.uniform(0, N)`` draws (redraw on repetition), then FMEstimatorCallback::checkSubset:
    the subset is redrawn when its LAST point is collinear with (or too close to) two earlier ones, in either image;
  - run7Point: points normalised (centroid, mean distance sqrt(2)), null space of the 7x9 system, cubic in lambda of
    det(lambda f1 + (1 - lambda) f2), solveCubic's closed forms (root order min, max, mid for three real roots),
    F scaled to F33 = 1 before and after de-normalisation;
  - computeError: max of the two squared point-line distances, rounded to float32; inlier when err <= (float)thr^2;
  - a model is kept when its inlier count exceeds max(best, 6); niters = RANSACUpdateNumIters(confidence,
    (N - good) / N, 7, niters).
The null-space basis of OpenCV's Jacobi SVD is not reproducible (the two zero singular values leave it free), so the
fundamental matrices agree with cv2's to rounding (~1e-9 relative) and, per sample, possibly in a different order
(which matters only for exact ties in the inlier count).  Pinned live against cv2 in tests/test_oracle_golden.py:
FM_7POINT solutions, RANSAC masks.
"""
import numpy as np

from . import CvRNG, cv_update_num_iters, triangulate

FLT_EPSILON = float(np.finfo(np.float32).eps)
DBL_EPSILON = float(np.finfo(np.float64).eps)


def cv_solve_cubic(c):
    """cv::solveCubic for c[0] x^3 + c[1] x^2 + c[2] x + c[3] -> list of real roots in OpenCV's order."""
    a0, a1, a2, a3 = (float(v) for v in c)
    if a0 == 0:
        if a1 == 0:
            if a2 == 0:
                return []
            return [-a3 / a2]
        d = a2 * a2 - 4 * a1 * a3
        if d < 0:
            return []
        d = np.sqrt(d)
        q1, q2 = (-a2 + d) * 0.5, (a2 + d) * -0.5
        if abs(q1) > abs(q2):
            r = [q1 / a1, a3 / q1]
        else:
            r = [q2 / a1, a3 / q2]
        return r if d > 0 else r[:1]
    a0 = 1.0 / a0
    a1, a2, a3 = a1 * a0, a2 * a0, a3 * a0
    Q = (a1 * a1 - 3 * a2) * (1.0 / 9)
    R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1.0 / 54)
    Qc = Q * Q * Q
    d = Qc - R * R
    if d > 0:
        theta = np.arccos(R / np.sqrt(Qc))
        t0, t1, t2 = -2 * np.sqrt(Q), theta * (1.0 / 3), a1 * (1.0 / 3)
        return [t0 * np.cos(t1) - t2, t0 * np.cos(t1 + 2 * np.pi / 3) - t2, t0 * np.cos(t1 + 4 * np.pi / 3) - t2]
    if d == 0:
        if R >= 0:
            x0, x1 = -2 * np.cbrt(R) - a1 / 3, np.cbrt(R) - a1 / 3
        else:
            x0, x1 = 2 * np.cbrt(-R) - a1 / 3, -np.cbrt(-R) - a1 / 3
        return [x0] if x0 == x1 else [x0, x1]
    d = np.sqrt(-d)
    e = np.cbrt(d + abs(R))
    if R > 0:
        e = -e
    return [(e + Q / e) - a1 * (1.0 / 3)]


def _det3_pencil(A, B):
    """coefficients (c0, c1, c2, c3) of det(lambda A + B) = c0 l^3 + c1 l^2 + c2 l + c3 (3x3, row-major 9-vectors)."""
    A, B = A.reshape(3, 3), B.reshape(3, 3)

    def det3(r0, r1, r2):
        return float(np.dot(r0, np.cross(r1, r2)))
    c0 = det3(A[0], A[1], A[2])
    c1 = det3(B[0], A[1], A[2]) + det3(A[0], B[1], A[2]) + det3(A[0], A[1], B[2])
    c2 = det3(A[0], B[1], B[2]) + det3(B[0], A[1], B[2]) + det3(B[0], B[1], A[2])
    c3 = det3(B[0], B[1], B[2])
    return c0, c1, c2, c3


def cv_fm_7point(m1, m2):
    """run7Point (fundam.cpp): 7 float32 point pairs -> list of up to three 3x3 fundamental matrices."""
    m1 = np.asarray(m1, np.float32).reshape(7, 2).astype(np.float64)
    m2 = np.asarray(m2, np.float32).reshape(7, 2).astype(np.float64)
    c1, c2 = m1.mean(0), m2.mean(0)
    s1 = np.sqrt(((m1 - c1) ** 2).sum(1)).mean()
    s2 = np.sqrt(((m2 - c2) ** 2).sum(1)).mean()
    if s1 < FLT_EPSILON or s2 < FLT_EPSILON:
        return []
    s1, s2 = np.sqrt(2.0) / s1, np.sqrt(2.0) / s2
    a, b = (m1 - c1) * s1, (m2 - c2) * s2
    x0, y0, x1, y1 = a[:, 0], a[:, 1], b[:, 0], b[:, 1]
    A = np.stack([x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, np.ones(7)], 1)   # (m2,1)^T F (m1,1) = 0
    Vt = np.linalg.svd(A, full_matrices=True)[2]
    f1, f2 = Vt[7].copy(), Vt[8].copy()
    f1 -= f2                                      # F ~ lambda f1 + f2
    roots = cv_solve_cubic(_det3_pencil(f1, f2))
    T1 = np.array([[s1, 0, -s1 * c1[0]], [0, s1, -s1 * c1[1]], [0, 0, 1]])
    T2 = np.array([[s2, 0, -s2 * c2[0]], [0, s2, -s2 * c2[1]], [0, 0, 1]])
    out = []
    for lam in roots:
        mu = 1.0
        s = f1[8] * lam + f2[8]
        F = np.empty(9)
        if abs(s) > DBL_EPSILON:
            mu = 1.0 / s
            lam = lam * mu
            F[8] = 1.0
        else:
            F[8] = 0.0
        F[:8] = f1[:8] * lam + f2[:8] * mu
        F = T2.T @ F.reshape(3, 3) @ T1
        if abs(F[2, 2]) > FLT_EPSILON:
            F = F * (1.0 / F[2, 2])
        out.append(F)
    return out


def fm_7point_elimination(m1, m2):
    """The same solver with the null space taken the way the CUDA kernel takes it (Gauss-Jordan elimination with complete
    pivoting, basis orthonormalised) instead of an SVD: the solutions are the singular members of the pencil the null space
    spans, whatever its basis.  tests/test_oracle_golden.py checks that OpenCV's RANSAC ends identically with either."""
    m1 = np.asarray(m1, np.float32).reshape(7, 2).astype(np.float64)
    m2 = np.asarray(m2, np.float32).reshape(7, 2).astype(np.float64)
    c1, c2 = m1.mean(0), m2.mean(0)
    s1 = np.sqrt(((m1 - c1) ** 2).sum(1)).mean()
    s2 = np.sqrt(((m2 - c2) ** 2).sum(1)).mean()
    if s1 < FLT_EPSILON or s2 < FLT_EPSILON:
        return []
    s1, s2 = np.sqrt(2.0) / s1, np.sqrt(2.0) / s2
    a, b = (m1 - c1) * s1, (m2 - c2) * s2
    x0, y0, x1, y1 = a[:, 0], a[:, 1], b[:, 0], b[:, 1]
    A = np.stack([x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, np.ones(7)], 1)
    perm = list(range(9))
    for k in range(7):
        sub = np.abs(A[k:, k:])
        if not sub.max() > 0:
            return []
        i, j = np.unravel_index(np.argmax(sub), sub.shape)
        i, j = i + k, j + k
        A[[k, i]] = A[[i, k]]
        A[:, [k, j]] = A[:, [j, k]]
        perm[k], perm[j] = perm[j], perm[k]
        A[k, k:] *= 1.0 / A[k, k]
        for r in range(7):
            if r != k and A[r, k] != 0:
                A[r, k:] -= A[r, k] * A[k, k:]
    f1, f2 = np.zeros(9), np.zeros(9)
    for k in range(7):
        f1[perm[k]], f2[perm[k]] = -A[k, 7], -A[k, 8]
    f1[perm[7]] = 1.0
    f2[perm[8]] = 1.0
    f1 /= np.linalg.norm(f1)
    f2 -= (f1 @ f2) * f1
    f2 /= np.linalg.norm(f2)
    f1 = f1 - f2
    roots = cv_solve_cubic(_det3_pencil(f1, f2))
    T1 = np.array([[s1, 0, -s1 * c1[0]], [0, s1, -s1 * c1[1]], [0, 0, 1]])
    T2 = np.array([[s2, 0, -s2 * c2[0]], [0, s2, -s2 * c2[1]], [0, 0, 1]])
    out = []
    for lam in roots:
        mu = 1.0
        s = f1[8] * lam + f2[8]
        F = np.empty(9)
        if abs(s) > DBL_EPSILON:
            mu = 1.0 / s
            lam = lam * mu
            F[8] = 1.0
        else:
            F[8] = 0.0
        F[:8] = f1[:8] * lam + f2[:8] * mu
        F = T2.T @ F.reshape(3, 3) @ T1
        if abs(F[2, 2]) > FLT_EPSILON:
            F = F * (1.0 / F[2, 2])
        out.append(F)
    return out


def _collinear_last(pts):
    """haveCollinearPoints (fundam.cpp): the LAST point against every pair of earlier ones."""
    i = len(pts) - 1
    for j in range(i):
        dx1, dy1 = float(pts[j, 0]) - float(pts[i, 0]), float(pts[j, 1]) - float(pts[i, 1])
        for k in range(j):
            dx2, dy2 = float(pts[k, 0]) - float(pts[i, 0]), float(pts[k, 1]) - float(pts[i, 1])
            if abs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def cv_fm_subset(rng: CvRNG, m1, m2, max_attempts=10000):
    """RANSACPointSetRegistrator::getSubset with FMEstimatorCallback::checkSubset -> 7 indices or None."""
    N = m1.shape[0]
    for _ in range(max_attempts):
        idx = []
        for _ in range(7):
            while True:
                v = rng.uniform(0, N)
                if v not in idx:
                    break
            idx.append(v)
        if not _collinear_last(m1[idx]) and not _collinear_last(m2[idx]):
            return idx
    return None


def cv_fm_errors_f32(F, m1, m2):
    """FMEstimatorCallback::computeError: max of the squared distances to the two epipolar lines, as float32."""
    F = np.asarray(F, np.float64).reshape(3, 3)
    x1, y1 = m1[:, 0].astype(np.float64), m1[:, 1].astype(np.float64)
    x2, y2 = m2[:, 0].astype(np.float64), m2[:, 1].astype(np.float64)
    with np.errstate(all="ignore"):
        a = F[0, 0] * x1 + F[0, 1] * y1 + F[0, 2]
        b = F[1, 0] * x1 + F[1, 1] * y1 + F[1, 2]
        c = F[2, 0] * x1 + F[2, 1] * y1 + F[2, 2]
        s2 = 1.0 / (a * a + b * b)
        d2 = x2 * a + y2 * b + c
        a = F[0, 0] * x2 + F[1, 0] * y2 + F[2, 0]
        b = F[0, 1] * x2 + F[1, 1] * y2 + F[2, 1]
        c = F[0, 2] * x2 + F[1, 2] * y2 + F[2, 2]
        s1 = 1.0 / (a * a + b * b)
        d1 = x1 * a + y1 * b + c
        return np.maximum(d1 * d1 * s1, d2 * d2 * s2).astype(np.float32)


def cv_find_fundamental_ransac(points1, points2, threshold, confidence, max_iters=1000, solver=cv_fm_7point):
    """cv2.findFundamentalMat(points1, points2, FM_RANSAC, threshold, confidence) for N >= 15 (triangulation.py:126-134)
    -> (F or None, bool mask, iterations run)."""
    m1 = np.asarray(points1).reshape(-1, 2).astype(np.float32)
    m2 = np.asarray(points2).reshape(-1, 2).astype(np.float32)
    N = m1.shape[0]
    assert N >= 15, "below 15 points OpenCV switches to LMedS"
    if threshold <= 0:
        threshold = 3.0
    if confidence < DBL_EPSILON or confidence > 1 - DBL_EPSILON:
        confidence = 0.99
    rng = CvRNG()
    thr = np.float32(threshold * threshold)
    niters, best, best_mask, best_F, it = int(max_iters), 0, np.zeros(N, bool), None, 0
    while it < niters:
        idx = cv_fm_subset(rng, m1, m2)
        if idx is None:
            break
        it += 1
        for F in solver(m1[idx], m2[idx]):
            with np.errstate(invalid="ignore"):
                mask = cv_fm_errors_f32(F, m1, m2) <= thr
            good = int(mask.sum())
            if good > max(best, 6):
                best, best_mask, best_F = good, mask, F
                niters = cv_update_num_iters(confidence, (N - good) / N, 7, niters)
    return best_F, best_mask, it


def decompose_essential(E):
    """triangulation.py:245-277 -> (4, 3, 4) candidates [R | t]."""
    U, _, Vh = np.linalg.svd(E)
    T = U[:, 2:]
    Wm = np.array([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]])
    R = [U @ Wm @ Vh, U @ Wm.T @ Vh]
    R = [r * (-1.0 if np.linalg.det(r) < 0 else 1.0) for r in R]
    M = np.zeros((4, 3, 4))
    for i in range(2):
        for j in range(2):
            M[2 * i + j] = np.concatenate([R[j], (-1) ** i * T], axis=-1)
    return M


def relative_pose(points1, points2, K, F, inliers):
    """triangulation.py:279-350 after the fundamental matrix is known -> (M (3,4), landmarks (N,3), mask (N,))."""
    p1 = np.asarray(points1, np.float64).reshape(-1, 2)
    p2 = np.asarray(points2, np.float64).reshape(-1, 2)
    K = np.asarray(K, np.float64)
    E = K.T @ F @ K
    q1, q2 = p1[inliers], p2[inliers]
    M1 = np.hstack([np.eye(3), np.zeros((3, 1))])
    best_valid, best_mask, best_M = -1, None, None
    for M2 in decompose_essential(E):
        X = triangulate(q1, q2, K @ M1, K @ M2, mode=0).reshape(-1, 3)
        X2 = X @ M2[:, :3].T + M2[:, 3]
        front = (X[:, 2] >= 0) & (X2[:, 2] >= 0)
        if front.sum() > best_valid:
            best_valid, best_mask, best_M = int(front.sum()), front, M2
    land = triangulate(p1, p2, K @ M1, K @ best_M, mode=0).reshape(-1, 3)
    mask = np.zeros(p1.shape[0], bool)
    mask[inliers] = best_mask
    return best_M, land, mask


def bootstrap(points1, points2, K, threshold, confidence):
    """LandmarksTriangulator(use_ransac=True, use_opencv=True).triangulate_matches on raw point arrays."""
    F, inl, it = cv_find_fundamental_ransac(points1, points2, threshold, confidence)
    if F is None:
        return None, None, None, None, it
    M, land, mask = relative_pose(points1, points2, K, F, inl)
    return F, M, land, mask, it
