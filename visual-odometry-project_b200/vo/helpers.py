"""Small coordinate helpers (reference: src/vo/helpers.py).  Host-side numpy, not on the hot path."""
import numpy as np
from scipy.linalg import expm, logm


def to_homogeneous_coordinates(points: np.ndarray) -> np.ndarray:
    """(N, D, 1) -> (N, D+1, 1) with a trailing one (helpers.py:5-15)."""
    assert points.ndim == 3, "Points must have three dimensions"
    ones = np.ones((points.shape[0], 1, 1))
    return np.concatenate((points, ones), axis=-2)


def to_cartesian_coordinates(points: np.ndarray) -> np.ndarray:
    """(N, D+1, 1) -> (N, D, 1), dividing by the last coordinate (helpers.py:18-28)."""
    assert points.ndim == 3, "Points must have three dimensions"
    return points[:, :-1] / points[:, -1:]


def normalize_points(points: np.ndarray):
    """Hartley normalisation (helpers.py:31-55): zero mean, RMS distance sqrt(D).  Returns (points', T)."""
    dim = points.shape[1]
    centroid = np.mean(points, axis=0, keepdims=True)
    sigma = np.sqrt(np.mean(np.sum((points - centroid) ** 2, axis=-2)))
    scale = np.sqrt(dim) / sigma
    T = np.diag([scale] * dim + [1])
    T[:-1, -1:] = -scale * centroid.reshape(dim, 1)
    moved = to_cartesian_coordinates(T @ to_homogeneous_coordinates(points))
    return moved, T


def to_skew_symmetric_matrix(v: np.ndarray) -> np.ndarray:
    """[v]x for a (3, 1) vector or a batch (N, 3, 1) (helpers.py:58-85)."""
    assert (v.ndim == 2 and v.shape == (3, 1)) or (v.ndim == 3 and v.shape[1:] == (3, 1)), \
        "Vector must be a single 3D vector or an array of 3D vectors"
    n = v.shape[0] if v.ndim == 3 else 1
    x, y, z = v[..., 0, 0], v[..., 1, 0], v[..., 2, 0]
    out = np.zeros((n, 3, 3))
    out[:, 0, 1], out[:, 0, 2] = -z, y
    out[:, 1, 0], out[:, 1, 2] = z, -x
    out[:, 2, 0], out[:, 2, 1] = -y, x
    return out.squeeze() if v.ndim == 2 else out


def twist_to_H_matrix(twist):
    """[v; w] -> 4x4 rigid transform via the matrix exponential (helpers.py:88-105)."""
    v, w = twist[:3], twist[3:]
    se = np.zeros((4, 4))
    se[:3, :3] = to_skew_symmetric_matrix(w.reshape(3, 1))
    se[:3, 3] = v
    return expm(se)


def skew_matrix_to_cross(M):
    """The vector x with M y = x cross y (helpers.py:133-146)."""
    return np.array([-M[1, 2], M[0, 2], -M[0, 1]])


def H_matrix_to_twist(H):
    """4x4 rigid transform -> [v; w] via the matrix logarithm (helpers.py:108-130)."""
    se = logm(H)
    return np.concatenate([se[:3, 3], skew_matrix_to_cross(se[:3, :3])])
