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
    ((376, 1241), 1000, 5, 9),  # KITTI-shaped, BASELINE configs[1] (width-specialised response kernel)
    ((120, 1226), 300, 5, 9),   # the other width-specialised instances: KITTI sequence 05, full HD, 4K DCI (odd heights:
    ((75, 1920), 300, 5, 9),    # the last tile row is partial)
    ((61, 4096), 300, 5, 9),
    ((75, 1921), 300, 5, 9),    # ... and the generic instance next to one of them
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


def _nms_dev(ctx, resp, K, r):
    import ctypes as C
    import torch
    from vo import _native as nat
    a = np.ascontiguousarray(resp, np.float64)
    if a.ndim == 2:
        a = a[None]
    F, H, W = a.shape
    d = torch.from_numpy(a).cuda()
    kp = torch.empty((F, K, 2), dtype=torch.int32, device="cuda")
    st = torch.zeros((F, 4), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    nat.check(nat.lib().vo_harris_nms_dev(ctx.handle, d.data_ptr(), F, H, W, r, K, kp.data_ptr(), st.data_ptr(), None), "nms")
    ctx.synchronize()
    return kp.cpu().numpy(), st.cpu().numpy()


@pytest.mark.parametrize("r", [0, 1, 2, 3, 4, 5, 7])
def test_nms_arbitrary_maps(ctx, r):
    """vo_harris_nms_dev on score maps that do not come from the response kernel: plateaus (ties), huge dynamic
    range, denormals, fewer local maxima than K -- against the literal argmax loop."""
    rng = np.random.default_rng(100 + r)
    H, W = 90, 140
    maps = []
    m = rng.random((H, W)) * (rng.random((H, W)) > 0.3)
    maps.append((m, 200))
    maps.append((rng.integers(0, 4, (H, W)).astype(np.float64), 150))          # plateaus: ties everywhere
    maps.append((10.0 ** rng.uniform(-200, 200, (H, W)) * (rng.random((H, W)) > 0.5), 300))
    maps.append((rng.integers(0, 50, (H, W)).astype(np.float64) * 5e-324, 100))  # denormals: high word is zero
    yy, xx = np.mgrid[0:H, 0:W]
    maps.append((np.exp(-((yy - 40) ** 2 + (xx - 60) ** 2) / 300.0) + 0.5 * np.exp(-((yy - 70) ** 2 + (xx - 20) ** 2) / 80.0), 120))
    maps.append((np.full((H, W), 3.0), 40))                                    # one plateau
    for i, (m, K) in enumerate(maps):
        if r == 0 and i in (1, 5):
            K = 30
        kp, _ = _nms_dev(ctx, m, K, r)
        assert np.array_equal(kp[0], oracle.harris_nms(np.ascontiguousarray(m), K, r)), (r, i)


def test_nms_large_frame_global_bitmaps(ctx):
    """a frame whose two bitmaps do not fit in shared memory takes the global-memory bitmaps"""
    rng = np.random.default_rng(5)
    H, W = 1000, 1900
    m = rng.random((H, W)) ** 8
    from scipy.ndimage import uniform_filter
    m = uniform_filter(m, 7) * (rng.random((H, W)) > 0.2)
    kp, st = _nms_dev(ctx, m, 700, 5)
    assert np.array_equal(kp[0], oracle.harris_nms(np.ascontiguousarray(m), 700, 5))


def test_nms_random_soak(ctx):
    """100 random (size, radius, K, score-map family) cases against the literal argmax loop: sparse noise, few-valued
    plateaus, smoothed blobs, 600 decades of dynamic range, quantised ramps"""
    from scipy.ndimage import uniform_filter
    rng = np.random.default_rng(2024)
    n = 0
    for it in range(100):
        H, W, r = int(rng.integers(24, 200)), int(rng.integers(24, 260)), int(rng.integers(0, 8))
        kind = it % 6
        if kind == 0:
            m = rng.random((H, W)) * (rng.random((H, W)) > rng.random())
        elif kind == 1:
            m = rng.integers(0, int(rng.integers(2, 9)), (H, W)).astype(np.float64)
        elif kind == 2:
            m = uniform_filter(rng.random((H, W)) ** 6, int(rng.integers(2, 9))) * (rng.random((H, W)) > 0.1)
        elif kind == 3:
            m = 10.0 ** rng.uniform(-300, 300, (H, W)) * (rng.random((H, W)) > 0.4)
        elif kind == 4:
            m = np.round(uniform_filter(rng.random((H, W)), 5) * 20) / 20
        else:
            m = uniform_filter(rng.random((H, W)) ** 3, 3)
        K = int(rng.integers(1, max(2, H * W // ((r + 1) ** 2) // 2 + 2)))
        m = np.ascontiguousarray(m)
        kp, _ = _nms_dev(ctx, m, K, r)
        assert np.array_equal(kp[0], oracle.harris_nms(m, K, r)), (it, kind, H, W, r, K)
        n += 1
    assert n == 100


def test_full_kitti_frame_vs_reference(ctx, golden):
    """The full frame the reference ships (1226x370, BASELINE configs[0]): the keypoints of the reference's own
    HarrisCornerDetector, K = 200 (its tests/test_harris.py) and K = 1000 (main.py's default), in selection order."""
    import os
    import cv2
    from conftest import GOLDEN
    from vo import _ops
    g = golden("loop")
    img = cv2.imread(os.path.join(GOLDEN, "kitti05", "000000.png"), cv2.IMREAD_GRAYSCALE)
    for K in (200, 1000):
        kp, _, _ = _ops.harris_detect(img, K, ctx=ctx)
        assert np.array_equal(kp, g[f"harris_full_kp{K}"]), K
    # the drop-in class on the same frame
    from vo.features import HarrisCornerDetector
    from vo.primitives import Frame
    fr = HarrisCornerDetector(num_keypoints=200).extractKeypoints(Frame(img.copy()))
    assert np.array_equal(fr.features.keypoints.reshape(-1, 2).astype(np.int32), g["harris_full_kp200"])


def test_speculative_threshold_is_invisible(ctx):
    """The NMS starts from the rank the K-th pick had in the previous call of the same frame slot (harris.cu,
    "speculative threshold").  Whatever the history -- the same frame again (speculation holds), a frame with far fewer
    strong corners (it falls short: second pass), a blank frame, then the first one again -- the keypoints are the oracle's."""
    a = synthetic_image(200, 320, seed=5)
    b = (synthetic_image(200, 320, seed=6) // 8 + 100).astype(np.uint8)       # low contrast: other score scale
    b[60:140, 100:220] = synthetic_image(80, 120, seed=7)                      # ... with one busy region
    blank = np.full((200, 320), 77, np.uint8)
    want = {id(im): oracle.harris_nms(oracle.harris_response(im, 9, 0.09), 300, 5) for im in (a, b, blank)}
    for im in (a, a, a, b, b, a, blank, a, b):
        kp, _, _ = _detect(ctx, im, 300)
        assert np.array_equal(kp, want[id(im)])
    # batches: every slot carries its own state
    batch1, batch2 = np.stack([a, b, blank]), np.stack([b, blank, a])
    for batch in (batch1, batch1, batch2, batch2, batch1):
        kp, _, _ = _detect(ctx, batch, 300)
        for f in range(3):
            assert np.array_equal(kp[f], want[id((a, b, blank)[[0, 1, 2][f]] if batch is batch1 else (b, blank, a)[f])])
