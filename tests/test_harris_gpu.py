"""GPU parity: CUDA Harris (through the C ABI) vs the oracle and the reference's golden vectors.
Bit-exact: float64 response map, keypoint coordinates in selection order, descriptors."""
import numpy as np
import pytest

import oracle
from conftest import synthetic_image

pytestmark = pytest.mark.gpu


def _detect(ctx, img, K, r=5, ps=9, kappa=0.09, dr=None):
    from vo import _ops
    return _ops.harris_detect(img, K, ps, kappa, r, dr, want_response=True, ctx=ctx)


def test_golden_crop(ctx, golden):
    g = golden("harris")
    for K, r in [(150, 5), (400, 3)]:
        kp, resp, desc = _detect(ctx, g["crop"], K, r, dr=9)
        assert np.array_equal(resp, g["crop_resp"])
        assert np.array_equal(kp, g[f"crop_kp_K{K}_r{r}"])
        assert np.array_equal(desc, g[f"crop_desc_K{K}_r{r}"])


def test_golden_zero_fill(ctx, golden):
    g = golden("harris")
    kp, _, _ = _detect(ctx, g["blank"], 40)
    assert np.array_equal(kp, g["blank_kp40"])


@pytest.mark.parametrize("shape,K,r,ps", [
    ((97, 131), 60, 5, 9),      # ragged, not a multiple of any tile
    ((120, 160), 300, 2, 5),
    ((64, 64), 50, 5, 3),
    ((200, 300), 500, 7, 9),    # r > patch_radius + 1: numpy's negative-slice case can trigger
    ((376, 1241), 1000, 5, 9),  # KITTI-shaped, BASELINE configs[1]
])
def test_vs_oracle(ctx, shape, K, r, ps):
    img = synthetic_image(shape[0], shape[1], seed=shape[0] * 7 + K)
    kp, resp, desc = _detect(ctx, img, K, r, ps, dr=4)
    resp_o = oracle.harris_response(img, ps, 0.09)
    assert np.array_equal(resp, resp_o)
    kp_o = oracle.harris_nms(resp_o, K, r)
    assert np.array_equal(kp, kp_o)
    assert np.array_equal(desc, oracle.harris_descriptors(img, kp_o, 4).astype(np.uint8))


def test_batch_matches_single(ctx):
    imgs = np.stack([synthetic_image(150, 210, seed=s) for s in range(5)])
    kp_b, resp_b, _ = _detect(ctx, imgs, 120)
    for f in range(5):
        kp, resp, _ = _detect(ctx, imgs[f], 120)
        assert np.array_equal(kp_b[f], kp)
        assert np.array_equal(resp_b[f], resp)


def test_ties_and_flat(ctx):
    # periodic pattern: many exactly equal scores -> tie-break by raster index must match numpy.argmax
    img = np.zeros((96, 128), np.uint8)
    img[::16, :] = 255
    img[:, ::16] = 255
    kp, resp, _ = _detect(ctx, img, 80)
    resp_o = oracle.harris_response(img)
    assert np.array_equal(resp, resp_o)
    assert np.array_equal(kp, oracle.harris_nms(resp_o, 80, 5))
    # constant image: no corners at all
    kp, resp, _ = _detect(ctx, np.full((40, 50), 7, np.uint8), 10)
    assert not resp.any() and not kp.any()


def test_errors(ctx):
    from vo._native import VoNativeError
    with pytest.raises(VoNativeError):
        _detect(ctx, np.zeros((8, 8), np.uint8), 10)       # smaller than the patch
    with pytest.raises(VoNativeError):
        _detect(ctx, np.zeros((64, 64), np.uint8), 10, ps=8)  # even patch size


def test_match_descriptors(ctx, golden):
    """GPU matcher vs the reference's cv2.BFMatcher run (golden) and vs the oracle on random 8-bit descriptors."""
    from vo import _ops
    g = golden("harris")
    pairs = _ops.match_descriptors(g["match_desc1"], g["match_desc2"], ctx=ctx)
    assert np.array_equal(pairs, g["match_pairs"])
    rng = np.random.default_rng(3)
    for Q, T, D in [(57, 133, 361), (1000, 1000, 361), (40, 2, 9), (300, 70, 25)]:
        base = rng.integers(0, 256, (max(Q, T), D))
        d1 = np.clip(base[:Q] + rng.integers(-6, 7, (Q, D)), 0, 255).astype(np.uint8)
        d2 = np.clip(base[rng.permutation(max(Q, T))[:T]] + rng.integers(-6, 7, (T, D)), 0, 255).astype(np.uint8)
        d2[T // 2] = d2[0]                                   # exact duplicates: ties broken by train index
        got = _ops.match_descriptors(d1, d2, ctx=ctx)
        assert np.array_equal(got, oracle.match_descriptors(d1, d2)), (Q, T, D)
