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
