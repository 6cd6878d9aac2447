"""Control flow of cv2.solvePnPRansac's RANSAC (OpenCV calib3d: RANSACPointSetRegistrator::run / getSubset,
RANSACUpdateNumIters; core: cv::RNG), which the reference's default pose path calls (src/vo/pose_estimation/p3p.py:142-151).

Only the sequential bookkeeping lives here (which subsets are drawn, when a model is kept, when the loop stops); the
models and inlier counts of whole batches of subsets come from the GPU (vo_p3p_ransac_host with OpenCV's float32 inlier
rule).  With it the default path returns cv2.solvePnPRansac's own inlier mask instead of an approximation of it."""
import numpy as np

__all__ = ["CvRNG", "update_num_iters", "subset4", "subsets"]


class CvRNG:
    """cv::RNG, multiply-with-carry; every RANSAC run starts from RNG((uint64)-1)."""

    def __init__(self, state: int = 0xFFFFFFFFFFFFFFFF) -> None:
        self.state = state

    def next(self) -> int:
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a: int, b: int) -> int:
        return a if a == b else a + self.next() % (b - a)


def update_num_iters(p: float, ep: float, model_points: int, max_iters: int) -> int:
    """RANSACUpdateNumIters: iterations needed for confidence p at outlier ratio ep (cvRound = round half to even)."""
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    tiny = float(np.finfo(np.float64).tiny)
    num = max(1.0 - p, tiny)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < tiny:
        return 0
    num, denom = float(np.log(num)), float(np.log(denom))
    return max_iters if (denom >= 0 or -num >= max_iters * (-denom)) else int(np.rint(num / denom))


def subset4(rng: CvRNG, n: int):
    """getSubset for four model points: distinct draws of uniform(0, n), redrawn on a repeat."""
    idx = []
    for _ in range(4):
        while True:
            v = rng.uniform(0, n)
            if v not in idx:
                break
        idx.append(v)
    return idx


def subsets(rng: CvRNG, n: int, count: int, model_points: int = 4) -> np.ndarray:
    """`count` consecutive getSubset results as int32 (count, model_points); advances `rng` (native loop:
    vo_cv_rng_subsets_host -- the Python one above costs more than the GPU batch it feeds)."""
    from vo import _native as nat
    st = np.array([rng.state], dtype=np.uint64)
    out = np.empty((count, model_points), dtype=np.int32)
    nat.check(nat.lib().vo_cv_rng_subsets_host(nat.ptr(st), int(n), int(model_points), int(count), nat.ptr(out)), "vo_cv_rng_subsets_host")
    rng.state = int(st[0])
    return out
