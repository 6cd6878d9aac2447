"""P3P pose estimation (reference: src/vo/pose_estimation/p3p.py).

estimate_pose runs the P3P minimal solver and the reprojection inlier count for whole batches of
RANSAC hypotheses on the GPU (vo_p3p_ransac_* in include/vo_b200.h) and replays the sequential
loop over them on the host:
  * use_opencv=False: the reference's own RANSAC (ransac.py): numpy rng 2023, adaptive iteration
    count, the outlier_ratio / n_iterations state that survives between calls.  Same rng position,
    iteration count, winning model and inlier mask as the reference.
  * use_opencv=True (the default, what src/main.py runs): cv2.solvePnPRansac's RANSAC restated
    (vo/algorithms/cv_ransac.py): OpenCV's fixed-seed generator and subset rule, float32 points,
    float32 errors with `<=`, RANSACUpdateNumIters.  The inlier mask is cv2.solvePnPRansac's own.
The refinement (p3p.py:188-213) is a Gauss-Newton on the GPU (vo_refine_pose_host)."""
import numpy as np

from vo import _ops
from vo.algorithms import RANSAC

__all__ = ["P3PPoseEstimator"]


class P3PPoseEstimator:
    FIRST_BATCH = 256      # hypotheses scored speculatively per GPU call
    NEXT_BATCH = 2048

    def __init__(self, intrinsic_matrix: np.ndarray, inlier_threshold: float, use_opencv: bool = True,
                 outlier_ratio: float = 0.9, confidence: float = 0.99, max_iterations: int = 10000,
                 nonlinear_refinement: bool = True) -> None:
        self._use_opencv = use_opencv
        self.intrinsic_matrix = intrinsic_matrix
        self.inlier_threshold = inlier_threshold
        self.outlier_ratio = outlier_ratio
        self.confidence = confidence
        self.max_iterations = max_iterations
        self.nonlinear_refinement = nonlinear_refinement
        # Same object the reference builds (p3p.py:110-121); its callables are never invoked here,
        # the GPU evaluates model_fn / error_fn for whole batches of samples.
        self.ransac = RANSAC(s_points=4, population=None, model_fn=None, error_fn=None,
                             inlier_threshold=self.inlier_threshold, outlier_ratio=self.outlier_ratio,
                             confidence=self.confidence, max_iterations=self.max_iterations, p3p=True)

    def estimate_pose(self, features):
        """((R 3x3, t 3x1) mapping world points into the camera, boolean inlier mask)  (p3p.py:123-186)."""
        points_3d, points_2d = features.landmarks, features.keypoints
        assert points_3d.shape[1] == 3 and points_2d.shape[1] == 2, "Invalid shape."
        assert points_3d is not None and points_2d is not None, "3D landmarks and 2D keypoints must be provided."
        if self._use_opencv:
            model, inliers = self._gpu_ransac_opencv(points_3d, points_2d)
            assert model is not None, "OpenCV P3P failed"                       # p3p.py:153
        else:
            model, inliers = self._gpu_ransac(points_3d, points_2d)
        if model is not None and self.nonlinear_refinement:
            model = self._nonlinear_refinement(points_3d[inliers], points_2d[inliers], model)
        return model, inliers

    def _gpu_ransac_opencv(self, points_3d, points_2d):
        """cv2.solvePnPRansac(flags=SOLVEPNP_P3P, iterationsCount=max_iterations, reprojectionError=inlier_threshold,
        confidence=confidence) (p3p.py:142-151): OpenCV's loop on the host, models and counts of its subsets on the GPU."""
        from vo.algorithms.cv_ransac import CvRNG, subsets, update_num_iters
        L32 = np.asarray(points_3d, dtype=np.float64).reshape(-1, 3).astype(np.float32).astype(np.float64)   # solvePnPRansac: CV_32F
        P32 = np.asarray(points_2d).reshape(-1, 2).astype(np.float32).astype(np.float64)
        N = L32.shape[0]
        if N <= 4:
            return None, None
        thr2 = float(self.inlier_threshold) ** 2
        table = np.full(N + 1, np.iinfo(np.int32).max, dtype=np.int64)      # no adaptive stop inside the batch: replayed below
        rng = CvRNG()
        niters = int(min(self.max_iterations, np.iinfo(np.int32).max))
        it, best, best_sample, best_model = 0, 0, None, None
        batch = self.FIRST_BATCH
        while it < niters:
            want = int(min(batch, niters - it))
            samples = subsets(rng, N, want)
            r = _ops.p3p_ransac(L32, P32, self.intrinsic_matrix, samples, thr2, table, np.iinfo(np.int32).max, want_all=True,
                                inclusive=True)
            for j in range(want):
                if not it < niters:
                    break
                it += 1                                         # a subset without a model still is an iteration
                if not r["valid"][j]:
                    continue
                good = int(r["counts"][j])
                if good > max(best, 3):
                    best, best_sample, best_model = good, samples[j].copy(), r["models"][j].copy()
                    niters = update_num_iters(self.confidence, (N - good) / N, 4, niters)
            batch = self.NEXT_BATCH
        if best_model is None:
            return None, None
        r = _ops.p3p_ransac(L32, P32, self.intrinsic_matrix, best_sample[None], thr2, table, np.iinfo(np.int32).max, inclusive=True)
        return (best_model[:9].reshape(3, 3).copy(), best_model[9:].reshape(3, 1).copy()), r["inliers"].copy()

    def _gpu_ransac(self, points_3d, points_2d):
        rs = self.ransac
        N = points_3d.shape[0]
        table = rs.iterations_table(N)
        thr = self.inlier_threshold                                     # ransac.py:105: squared error < threshold
        n, best_count, best = 0, -1, None
        batch = self.FIRST_BATCH
        drawn, draw_cap = 0, 10 * int(min(self.max_iterations, 1 << 24)) + 4096
        while n < rs.n_iterations:
            if drawn >= draw_cap:
                # ransac.py:90-103 would spin forever when no sample yields a model (degenerate landmarks);
                # here that would be an endless stream of GPU round trips, so give up loudly instead.
                raise RuntimeError(f"P3P-RANSAC: {drawn} samples drawn, {n} valid models: degenerate correspondences")
            state0 = rs.rng.bit_generator.state
            want = int(min(batch, max(1, rs.n_iterations - n)))
            samples = np.stack([rs.draw(N) for _ in range(want)]).astype(np.int32)
            r = _ops.p3p_ransac(points_3d, points_2d, self.intrinsic_matrix, samples, thr, table,
                                rs.n_iterations, start_n=n, start_best=best_count)
            n, consumed = int(r["n"]), int(r["consumed"])
            drawn += consumed
            if int(r["best"]) >= 0:
                best_count = int(r["best_count"])
                best = ((r["R"].copy(), r["t"].copy()), r["inliers"].copy())
                # ransac.py:113-120 (adaptive update after an improvement)
                rs.outlier_ratio = min(max(1 - np.int64(best_count) / N, 0.01), 0.99)
                rs.n_iterations = int(r["n_iterations"])
            if consumed < want:
                # the loop stopped inside this batch: put the rng where the reference's would be
                rs.rng.bit_generator.state = state0
                for _ in range(consumed):
                    rs.draw(N)
                break
            batch = self.NEXT_BATCH
        if best is None:
            return None, None
        return best

    def _nonlinear_refinement(self, points_3d, points_2d, best_model):
        """Minimise the reprojection residuals over the 6-dof pose (p3p.py:188-213) on the GPU: a damped Gauss-Newton
        on SE(3) run to the minimum of the cost the reference hands to scipy.optimize.least_squares (which stops on
        ftol about 1e-3 short of it; DESIGN.md has the measured distances)."""
        R, t, _ = _ops.refine_pose(points_3d, points_2d, self.intrinsic_matrix, best_model[0], np.asarray(best_model[1]).reshape(3))
        return R, t
