"""RANSAC (reference: src/vo/algorithms/ransac.py).

`RANSAC.find_best_model` keeps the reference's exact sequential semantics (rng seeded with 2023,
adaptive iteration count, and the outlier_ratio / n_iterations state that survives between calls).
For arbitrary Python model_fn / error_fn callables it is a host loop, as in the reference.  The P3P
instance built by vo.pose_estimation.P3PPoseEstimator does not call those callables: it evaluates
batches of hypotheses on the GPU (vo_p3p_ransac_*) and replays the same loop over the batch, which
yields the same best model, inlier mask, iteration count and rng position."""
from typing import Callable

import numpy as np

__all__ = ["RANSAC"]


class RANSAC:
    def __init__(self, s_points: int, population, model_fn: Callable, error_fn: Callable, inlier_threshold: float,
                 outlier_ratio: float = 0.9, confidence: float = 0.99, max_iterations: int = np.inf,
                 adaptive: bool = True, p3p: bool = False) -> None:
        self.s = s_points
        self.population = np.array(population)
        self.model_fn = model_fn
        self.error_fn = error_fn
        self.inlier_threshold = inlier_threshold
        self.outlier_ratio = outlier_ratio
        self.confidence = confidence
        self.adaptive = adaptive
        self.p3p = p3p
        self.rng = np.random.default_rng(2023)                                   # ransac.py:52
        self.max_iterations = max_iterations
        self.n_iterations = min(max_iterations, self.compute_n_iterations())     # ransac.py:56
        self._tables = {}

    def compute_n_iterations(self) -> int:
        """ceil(log(1 - p) / log(1 - (1 - eps)^s))  (ransac.py:58-67)."""
        k = np.ceil(np.log(1 - self.confidence) / np.log(1 - (1 - self.outlier_ratio) ** self.s))
        return int(k)

    def iterations_table(self, n_population: int) -> np.ndarray:
        """n_iterations after an improvement to `c` inliers, c = 0..N, evaluated with the same numpy
        scalar expressions as ransac.py:113-120 (so ceil() lands on the same integer)."""
        key = (n_population, self.s, self.confidence, self.max_iterations)
        tab = self._tables.get(key)
        if tab is None:
            saved = self.outlier_ratio
            tab = np.empty(n_population + 1, dtype=np.int64)
            for c in range(n_population + 1):
                self.outlier_ratio = min(max(1 - np.int64(c) / n_population, 0.01), 0.99)
                tab[c] = int(min(self.max_iterations, self.compute_n_iterations()))
            self.outlier_ratio = saved
            self._tables[key] = tab
        return tab

    def draw(self, n_population: int) -> np.ndarray:
        """One sample of s distinct indices, consuming the rng exactly like ransac.py:92-94."""
        return self.rng.choice(np.arange(n_population), replace=False, size=self.s)

    def find_best_model(self, population=None):
        """Host loop for generic callables (ransac.py:69-129)."""
        if population is not None:
            self.population = np.array(population)
        assert self.population is not None, "Population must be provided"
        best_count, best_inliers, best_model, n = -1, None, None, 0
        size = len(self.population)
        while n < self.n_iterations:
            sample = self.population[self.draw(size)]
            model = self.model_fn(sample)
            if model is None:
                continue
            inliers = self.error_fn(model, self.population) < self.inlier_threshold
            count = inliers.sum()
            if count > best_count:
                best_count, best_inliers, best_model = count, inliers, model
                if self.adaptive:
                    self.outlier_ratio = min(max(1 - best_count / size, 0.01), 0.99)
                    self.n_iterations = int(min(self.max_iterations, self.compute_n_iterations()))
            n += 1
        if not self.p3p:   # refit on all inliers, except for P3P (ransac.py:123-127)
            best_model = self.model_fn(self.population[best_inliers])
        return best_model, best_inliers
