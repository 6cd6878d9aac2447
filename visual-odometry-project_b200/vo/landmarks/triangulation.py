"""Landmark triangulation and two-view bootstrapping (reference: src/vo/landmarks/triangulation.py).

GPU: _linear_triangulation and triangulate_candidates (vo_triangulate_*), and the two-view bootstrap as
src/main.py configures it (use_ransac=True, use_opencv=True): cv2.findFundamentalMat's RANSAC restated, the
essential-matrix decomposition, the cheirality vote and the landmarks in ONE call (vo_bootstrap_host).  The
other bootstrap configurations (the reference's own 8-point RANSAC, no RANSAC) keep the reference's host
arithmetic; their triangulations still run on the GPU."""
import numpy as np

from vo import _ops
from vo.algorithms import RANSAC
from vo.helpers import normalize_points, to_homogeneous_coordinates

__all__ = ["LandmarksTriangulator"]


class LandmarksTriangulator:
    def __init__(self, camera1, camera2, use_ransac: bool = True, outlier_ratio: float = 0.9,
                 ransac_threshold: float = 3.0, ransac_confidence=0.99, use_opencv: bool = True) -> None:
        self.camera1, self.camera2 = camera1, camera2
        self._use_ransac = use_ransac
        self._outlier_ratio = outlier_ratio
        self._ransac_reproj_threshold = ransac_threshold
        self._ransac_confidence = ransac_confidence
        self._use_opencv = use_opencv

    # ---- hot path -----------------------------------------------------------------------------
    def triangulate_candidates(self, features, current_pose: np.ndarray) -> np.ndarray:
        """Triangulate every candidate track from its start (own pose) and end (current pose)
        (triangulation.py:38-86).  use_opencv picks cv2.triangulatePoints' 4x4 system, otherwise the
        6x4 cross-product system; both are solved on the GPU."""
        sel = features.candidate_mask
        starts, ends = features.tracks[sel], features.keypoints[sel]
        proj1 = self.camera1.intrinsic_matrix @ np.linalg.inv(features.poses[sel])[:, :3]
        proj2 = self.camera2.intrinsic_matrix @ np.linalg.inv(current_pose)[:3]
        if starts.shape[0] == 0:
            return np.empty((0, 3, 1))
        X = _ops.triangulate(starts, ends, proj1, proj2, mode=1 if self._use_opencv else 0)
        return X.reshape(-1, 3, 1)

    def _linear_triangulation(self, points1, points2, C1, C2):
        """DLT with the 6x4 system [p1]x C1 ; [p2]x C2 (triangulation.py:352-389) -> (N, 3, 1)."""
        assert points1.shape == points2.shape, "Input points dimension mismatch"
        assert points1.shape[1] == 2, "Points must have two rows for (u,v)"
        assert points1.shape[2] == 1, "Points must be a column vector"
        assert C1.shape == (3, 4) and C2.shape == (3, 4), "Matrix C1 and C2 must be 3 rows and 4 columns [R T]"
        return _ops.triangulate(points1, points2, C1, C2, mode=0).reshape(-1, 3, 1)

    # ---- bootstrap ----------------------------------------------------------------------------
    def triangulate_matches(self, matches):
        """Relative pose + landmarks from the matched keypoints of two frames (triangulation.py:88-108)."""
        p1 = matches.frame1.features.matched_candidate_inliers_keypoints
        p2 = matches.frame2.features.matched_candidate_inliers_keypoints
        return self._find_relative_pose(p1, p2)

    def _gpu_bootstrap(self, points1, points2):
        K1, K2 = self.camera1.intrinsic_matrix, self.camera2.intrinsic_matrix
        if not np.array_equal(K1, K2):
            raise NotImplementedError("the GPU bootstrap assumes one camera (main.py passes the same Camera twice)")
        r = _ops.bootstrap(points1, points2, K1, self._ransac_reproj_threshold, self._ransac_confidence)
        if not r["found"]:
            raise RuntimeError("bootstrap: the fundamental-matrix RANSAC found no model")
        return r

    def _find_fundamental_matrix_ransac(self, points1, points2):
        """F and its inlier mask (triangulation.py:110-163); use_opencv: cv2.findFundamentalMat's RANSAC on the GPU."""
        if self._use_opencv:
            r = self._gpu_bootstrap(points1, points2)
            return r["F"], r["f_mask"]
        n1, T1 = normalize_points(points1)
        n2, T2 = normalize_points(points2)

        def fit(sample):
            return self._find_fundamental_matrix(sample[:, 0], sample[:, 1], is_normalized=True)

        def algebraic_error(F, pts):
            a, b = to_homogeneous_coordinates(pts[:, 0]), to_homogeneous_coordinates(pts[:, 1])
            return np.sum((b.transpose((0, 2, 1)) @ F @ a) ** 2, axis=(1, 2))

        search = RANSAC(s_points=8, population=np.stack([n1, n2], axis=1), model_fn=fit, error_fn=algebraic_error,
                        inlier_threshold=self._ransac_reproj_threshold, outlier_ratio=self._outlier_ratio,
                        confidence=self._ransac_confidence)
        F, inl = search.find_best_model()
        return T2.T @ F @ T1, inl

    def _find_fundamental_matrix(self, points1, points2, is_normalized: bool = False):
        """Normalised 8-point algorithm (triangulation.py:165-222)."""
        assert points1.shape == points2.shape, "Input points dimension mismatch"
        assert points1.shape[0] >= 8, "Not enough points for 8-point algorithm"
        assert points1.shape[1] == 2, "Points must have two rows for (u,v)"
        assert points1.shape[2] == 1, "Points must be a column vector"
        if self._use_opencv:
            import cv2
            return cv2.findFundamentalMat(points1=points1, points2=points2, method=cv2.FM_8POINT)[0]
        if not is_normalized:
            points1, T1 = normalize_points(points1)
            points2, T2 = normalize_points(points2)
        a, b = to_homogeneous_coordinates(points1), to_homogeneous_coordinates(points2)
        Q = np.stack([np.kron(a[i], b[i]).T.ravel() for i in range(a.shape[0])])
        F = np.linalg.svd(Q, full_matrices=True)[2][-1].reshape(3, 3).T
        U, S, Vh = np.linalg.svd(F)
        S[-1] = 0                                   # rank 2
        F = U @ np.diag(S) @ Vh
        return F if is_normalized else T2.T @ F @ T1

    def _find_essential_matrix(self, points1, points2):
        K1, K2 = self.camera1.intrinsic_matrix, self.camera2.intrinsic_matrix
        if self._use_ransac:
            F, inl = self._find_fundamental_matrix_ransac(points1, points2)
            return K2.T @ F @ K1, inl
        return K2.T @ self._find_fundamental_matrix(points1, points2) @ K1

    def _decompose_essential_matrix(self, E: np.ndarray) -> np.ndarray:
        """The four [R | t] candidates of E = [t]x R (triangulation.py:245-277)."""
        U, _, Vh = np.linalg.svd(E)
        t = U[:, 2:]
        Wm = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]])
        rots = [U @ Wm @ Vh, U @ Wm.T @ Vh]
        rots = [R * (-1 if np.linalg.det(R) < 0 else 1) for R in rots]
        M = np.zeros((4, 3, 4))
        for i in range(2):
            for j in range(2):
                M[2 * i + j] = np.concatenate([rots[j], (-1) ** i * t], axis=-1)
        return M

    def _find_relative_pose(self, points1, points2):
        """M = [R t] from frame 1 to frame 2, landmarks in frame-1 coordinates and (with RANSAC) the
        inlier mask (triangulation.py:279-350)."""
        if self._use_ransac and self._use_opencv:          # main.py's configuration: everything in one GPU call
            r = self._gpu_bootstrap(points1, points2)
            return r["M"], r["landmarks"].reshape(-1, 3, 1), r["mask"]
        if self._use_ransac:
            E, inl = self._find_essential_matrix(points1, points2)
            q1, q2 = points1[inl], points2[inl]
        else:
            E = self._find_essential_matrix(points1, points2)
            q1, q2 = points1, points2
        K1, K2 = self.camera1.intrinsic_matrix, self.camera2.intrinsic_matrix
        M1 = np.hstack((np.eye(3), np.zeros((3, 1))))
        best_valid, best_mask, best_M = -1, None, None
        for M2 in self._decompose_essential_matrix(E):
            X = self._linear_triangulation(q1, q2, K1 @ M1, K2 @ M2)
            X2 = M2[:, :3] @ X + M2[:, 3:]
            in_front = ((X[:, -1] >= 0) & (X2[:, -1] >= 0)).flatten()      # cheirality in both cameras
            if in_front.sum() > best_valid:
                best_valid, best_mask, best_M = in_front.sum(), in_front, M2
        landmarks = self._linear_triangulation(points1, points2, K1 @ M1, K2 @ best_M)
        if self._use_ransac:
            mask = np.zeros((points1.shape[0],), dtype=bool)
            mask[inl] = best_mask
            return best_M, landmarks, mask
        return best_M, landmarks
