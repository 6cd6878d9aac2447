"""GPU parity: CUDA pyramidal LK (C ABI) vs the oracle (bit-exact: same integer sums, same float32
op order) and vs cv2.calcOpticalFlowPyrLK golden vectors (status identical, tracks within 1e-2 px)."""
import numpy as np
import pytest

import oracle
from conftest import synthetic_image

pytestmark = pytest.mark.gpu


def _track(ctx, a, b, pts, **kw):
    from vo import _ops
    return _ops.klt_track(a, b, pts, ctx=ctx, **kw)


def test_golden_vs_cv2(ctx, golden):
    g = golden("klt")
    for suffix, kw in [("", {}), ("_w21_l3", dict(win=21, max_level=3, max_iters=30, epsilon=0.01))]:
        nxt, st, err = _track(ctx, g["prev"], g["next"], g["pts"], **kw)
        st_ref = g["status" + suffix]
        assert np.array_equal(st, st_ref)
        ok = st_ref == 1
        d = np.abs(nxt - g["next_pts" + suffix]).max(axis=1)[ok]
        assert d.max() < 1e-2                       # BASELINE north_star tolerance: 1e-2 px
        assert np.median(d) < 1e-4
        assert np.abs(err - g["err" + suffix])[ok].max() < 5e-2


def _shifted_pair(h, w, seed, dx, dy):
    big = synthetic_image(h + 40, w + 40, seed)
    a = np.ascontiguousarray(big[20:20 + h, 20:20 + w])
    b = np.ascontiguousarray(big[20 - dy:20 - dy + h, 20 - dx:20 - dx + w])
    return a, b


@pytest.mark.parametrize("shape,win,lvl,n", [
    ((97, 131), 17, 2, 200),
    ((376, 1241), 17, 2, 1000),     # BASELINE configs[1]
    ((240, 320), 21, 3, 300),
    ((60, 70), 9, 4, 100),          # pyramid truncated by the window-size rule
    ((120, 160), 17, 0, 100),       # no pyramid
])
def test_vs_oracle_bitexact(ctx, shape, win, lvl, n):
    a, b = _shifted_pair(shape[0], shape[1], seed=shape[0] + win, dx=3, dy=-2)
    rng = np.random.default_rng(n)
    pts = np.stack([rng.uniform(-win, shape[1] + win, n), rng.uniform(-win, shape[0] + win, n)], 1).astype(np.float32)
    nxt, st, err = _track(ctx, a, b, pts, win=win, max_level=lvl)
    nxt_o, st_o, err_o = oracle.klt_track(a, b, pts, win=win, max_level=lvl)
    assert np.array_equal(st, st_o)
    assert np.array_equal(nxt, nxt_o)               # bit-exact float32
    assert np.array_equal(err, err_o)
    good = st_o == 1
    assert good.sum() > n // 4
    # property: a pure integer shift is recovered for well-textured interior points
    inner = good & (pts[:, 0] > 2 * win) & (pts[:, 0] < shape[1] - 2 * win) & \
        (pts[:, 1] > 2 * win) & (pts[:, 1] < shape[0] - 2 * win) & (err_o < 2)
    if inner.sum():
        flow = nxt[inner] - pts[inner]
        assert np.median(np.abs(flow - np.array([3, -2], np.float32))) < 0.1


def test_batch_and_empty(ctx):
    pairs = [_shifted_pair(100, 140, s, 1, 1) for s in range(3)]
    a = np.stack([p[0] for p in pairs])
    b = np.stack([p[1] for p in pairs])
    rng = np.random.default_rng(0)
    pts = np.stack([rng.uniform(0, 140, (3, 64)), rng.uniform(0, 100, (3, 64))], -1).astype(np.float32)
    nxt, st, err = _track(ctx, a, b, pts)
    for f in range(3):
        n1, s1, e1 = _track(ctx, a[f], b[f], pts[f])
        assert np.array_equal(n1, nxt[f]) and np.array_equal(s1, st[f]) and np.array_equal(e1, err[f])
    n0, s0, e0 = _track(ctx, a[0], b[0], np.zeros((0, 2), np.float32))
    assert n0.shape == (0, 2) and s0.shape == (0,)


def test_near_integer_positions(ctx):
    """window corners a hair off the pixel grid: the fourth bilinear weight 2^14 - (w00 + w01 + w10) rounds to -1
    or 0, which the packed dp2a path must treat as a signed 16-bit value (integer shifts keep the tracked
    positions near the grid during the iterations too)"""
    a, b = _shifted_pair(200, 260, 11, 2, -1)
    rng = np.random.default_rng(4)
    base = np.stack([rng.integers(20, 240, 600), rng.integers(20, 180, 600)], axis=1).astype(np.float32)
    eps = rng.choice(np.array([0.0, 1e-5, 2e-5, 3e-5, 6e-5, 1e-4, -1e-5, -3e-5, 0.5, 0.99997, 3.8e-5, 4.6e-5], np.float32), size=(600, 2))
    pts = (base + eps).astype(np.float32)
    nxt, st, err = _track(ctx, a, b, pts)
    nxt_o, st_o, err_o = oracle.klt_track(a, b, pts)
    assert np.array_equal(st, st_o) and np.array_equal(nxt, nxt_o) and np.array_equal(err, err_o)


@pytest.mark.parametrize("shape", [(97, 131), (376, 1241), (61, 63), (64, 256), (130, 243), (33, 1000), (12, 40), (23, 17)])
def test_pyramid_levels_vs_oracle(ctx, shape):
    """vo_klt_build_pyramid_dev level by level against the oracle's cv2.pyrDown restatement (bit-exact), sizes whose
    last input word straddles the border in every way (W % 4 = 0..3), two frames per call"""
    import ctypes as C
    import torch
    from vo import _native as nat
    H, W = shape
    imgs = np.ascontiguousarray(np.stack([synthetic_image(H, W, seed=H + W), synthetic_image(H, W, seed=H * W)]))   # C order
    L = nat.lib()
    nl = C.c_int()
    lh, lw = (C.c_int * 8)(), (C.c_int * 8)()
    lp, lo = (C.c_size_t * 8)(), (C.c_size_t * 8)()
    fb = C.c_size_t()
    nat.check(L.vo_klt_pyramid_layout(H, W, 3, 5, C.byref(nl), lh, lw, lp, lo, C.byref(fb)), "layout")
    d_img = torch.from_numpy(imgs).cuda()
    d_pyr = torch.zeros((2, fb.value), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    nat.check(L.vo_klt_build_pyramid_dev(ctx.handle, d_img.data_ptr(), 2, H, W, W, H * W, 3, 5, d_pyr.data_ptr(), None), "pyramid")
    ctx.synchronize()
    pyr = d_pyr.cpu().numpy()
    assert nl.value >= 2
    for f in range(2):
        ref = imgs[f]
        for l in range(nl.value):
            if l > 0:
                ref = oracle.pyr_down(ref)
            got = pyr[f, lo[l]:lo[l] + lh[l] * lp[l]].reshape(lh[l], lp[l])[:, :lw[l]]
            assert ref.shape == (lh[l], lw[l])
            assert np.array_equal(got, ref), (shape, f, l)


def test_bgr2gray_bitexact_vs_cv2(ctx):
    """klt.py:57-62 converts with cv2.cvtColor(BGR2GRAY); the device kernel must reproduce OpenCV's fixed point on
    every colour (a 4096x4096 image holding all 2^24 of them) and on ragged widths."""
    import cv2
    from vo import _ops
    b, g, r = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    img = np.ascontiguousarray(np.stack([b, g, r], -1).reshape(4096, 4096, 3))
    assert np.array_equal(_ops.bgr2gray(img, ctx=ctx), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    rng = np.random.default_rng(0)
    for shape in [(3, 37, 53, 3), (1, 5, 1, 3), (2, 64, 1241, 3)]:
        a = rng.integers(0, 256, shape, dtype=np.uint8)
        want = np.stack([cv2.cvtColor(x, cv2.COLOR_BGR2GRAY) for x in a])
        assert np.array_equal(_ops.bgr2gray(a, ctx=ctx), want)


def test_bgr_input_and_pyramid_cache(ctx, golden):
    """BGR frames through vo_klt_track_bgr_host equal the gray path, and a second call whose `prev` is the last call's
    `next` reuses the resident pyramid (same results)."""
    from vo import _native as nat, _ops
    g = golden("klt")
    c0, c1 = g["prev"], g["next"]
    rng = np.random.default_rng(1)
    b0 = np.stack([c0, c0, c0], -1)
    b1 = np.stack([c1, c1, c1], -1)
    b2 = np.roll(b1, 2, axis=1)
    pts = g["pts"]
    want = _ops.klt_track(c0, c1, pts, ctx=ctx)
    got = _ops.klt_track(b0, b1, pts, ctx=ctx)
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
    h0 = nat.lib().vo_klt_cache_hits(ctx.handle)
    a2 = _ops.klt_track(b1, b2, got[0], ctx=ctx)          # prev == last next: cache hit
    assert nat.lib().vo_klt_cache_hits(ctx.handle) == h0 + 1
    _ops.klt_track(b0, b1, pts, ctx=ctx)                  # miss (prev differs)
    assert nat.lib().vo_klt_cache_hits(ctx.handle) == h0 + 1
    a3 = _ops.klt_track(np.ascontiguousarray(b1[..., 0]), np.ascontiguousarray(b2[..., 0]), got[0], ctx=ctx)   # gray, no cache
    for a, b in zip(a2, a3):
        assert np.array_equal(a, b)
