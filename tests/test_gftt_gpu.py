"""GPU: cv2.goodFeaturesToTrack (the detector of the reference's KLT tracker mode, klt.py:24-26, 87-115) on the device.

Bar: the corner list -- coordinates AND order -- equals cv2's own on the same image, and equals the oracle's (which is
pinned bit for bit to cv2.cornerMinEigenVal / goodFeaturesToTrack in tests/test_oracle_golden.py).  The eigenvalue map is
compared bit for bit wherever it can matter (values above the quality threshold); elsewhere OpenCV's running float64 box
sums carry rounding drift that a direct sum does not reproduce (documented in csrc/gftt.cu), so the bar there is 1e-5
of the frame maximum."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, synthetic_image

pytestmark = pytest.mark.gpu

PARAMS = dict(maxCorners=500, qualityLevel=0.01, minDistance=8, blockSize=7)          # klt.py:24-26


def _kitti(i):
    import cv2
    return cv2.imread(os.path.join(GOLDEN, "kitti05", f"{i:06d}.png"), cv2.IMREAD_GRAYSCALE)


def _check(img, ctx, **kw):
    import cv2
    from vo import _ops
    p = dict(PARAMS, **kw)
    want = cv2.goodFeaturesToTrack(img, **p)
    want = want.reshape(-1, 2) if want is not None else np.empty((0, 2), np.float32)
    got, eig, st = _ops.good_features_to_track(img, p["maxCorners"], p["qualityLevel"], p["minDistance"], p["blockSize"],
                                                want_eig=True, want_stats=True, ctx=ctx)
    e_or = oracle.min_eigen_val(img, p["blockSize"])
    thr = np.float32(float(e_or.max()) * p["qualityLevel"])
    hot = e_or > thr
    assert np.array_equal(eig[hot], e_or[hot])
    assert np.abs(eig - e_or).max() <= 1e-5 * max(float(e_or.max()), 1e-30)
    assert np.array_equal(got, want), (len(got), len(want))
    assert np.array_equal(got, oracle.gftt_select(e_or, p["maxCorners"], p["qualityLevel"], p["minDistance"]))
    assert st[3] == 0
    return got, st


def test_gftt_kitti_frames_equal_cv2(ctx, golden):
    g = golden("loop")
    for i in range(6):
        got, st = _check(_kitti(i), ctx)
        assert np.array_equal(got, g[f"gftt_{i}"])            # the corners the reference's tracker started from
    assert np.array_equal(g["gftt_0"], g["init_kp"])


@pytest.mark.parametrize("shape,seed", [((120, 200), 1), ((376, 1241), 2), ((64, 31), 3), ((50, 97), 4), ((33, 33), 5), ((200, 1215), 6)])
def test_gftt_synthetic_shapes(ctx, shape, seed):
    """ragged widths (vector body / scalar tail of OpenCV's row filter), tiny frames, dense textures (11k candidates)"""
    _check(synthetic_image(shape[0], shape[1], seed), ctx)


def test_gftt_parameters_and_edge_cases(ctx):
    img = synthetic_image(160, 240, 9)
    _check(img, ctx, maxCorners=50)
    _check(img, ctx, maxCorners=1000, minDistance=3, qualityLevel=0.001)
    _check(img, ctx, blockSize=3, minDistance=12)
    _check(img, ctx, blockSize=5, minDistance=1)
    flat = np.full((40, 60), 77, np.uint8)                      # no gradient: no corners
    got, st = _check(flat, ctx)
    assert len(got) == 0
    steps = np.zeros((64, 96), np.uint8)                        # exact ties: equal corners at mirrored places
    steps[16:48, 24:72] = 200
    _check(steps, ctx)
    _check(np.ascontiguousarray(steps.T), ctx)


def test_gftt_batch_and_tracker_class(ctx):
    from vo import _ops
    from vo.features import KLTTracker
    from vo.primitives import Frame
    import cv2
    frames = np.stack([_kitti(i) for i in range(3)])
    got = _ops.good_features_to_track(frames, ctx=ctx)
    for i in range(3):
        assert np.array_equal(got[i], cv2.goodFeaturesToTrack(frames[i], **PARAMS).reshape(-1, 2))
    bgr = cv2.imread(os.path.join(GOLDEN, "kitti05", "000000.png"))
    trk = KLTTracker(Frame(bgr))                                # klt.py:40-50 through the drop-in class
    assert np.array_equal(trk.frame.features.keypoints.reshape(-1, 2), got[0])
