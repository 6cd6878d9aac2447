"""GPU: BASELINE.json configs[3] -- large-frame Harris / KLT stress (4096x2160, 10k tracked keypoints, 4-level
pyramid).  The oracle checks what it can finish in seconds; the rest is checked through properties of the
reference's greedy selection that do not depend on size."""
import numpy as np
import pytest

import oracle
from conftest import synthetic_image

pytestmark = pytest.mark.gpu

H, W = 2160, 4096


@pytest.fixture(scope="module")
def big_pair():
    rng = np.random.default_rng(77)
    # cheap large texture: tile a seeded 540x1024 texture with per-tile intensity changes + rectangles
    base = synthetic_image(540 + 16, 1024 + 16, seed=77).astype(np.int32)
    big = np.zeros((H + 16, W + 16), np.int32)
    for ty in range(4):
        for tx in range(4):
            big[ty * 540:(ty + 1) * 540 + 16, tx * 1024:(tx + 1) * 1024 + 16] = (base * (0.6 + 0.1 * ((ty * 4 + tx) % 5))).astype(np.int32)
    for _ in range(3000):
        y, x = rng.integers(0, H - 40), rng.integers(0, W - 40)
        big[y:y + rng.integers(6, 40), x:x + rng.integers(6, 40)] = rng.integers(0, 256)
    big = np.clip(big, 0, 255).astype(np.uint8)
    a = np.ascontiguousarray(big[8:8 + H, 8:8 + W])
    b = np.ascontiguousarray(big[6:6 + H, 5:5 + W])      # moved by (+3, +2) pixels
    return a, b


def test_harris_large_frame(ctx, big_pair):
    from vo import _ops
    a, _ = big_pair
    K, r = 10000, 5
    kp, resp, _ = _ops.harris_detect(a, K, want_response=True, ctx=ctx)
    resp_o = oracle.harris_response(a)
    assert np.array_equal(resp, resp_o)                                   # float64 score map, bit-exact
    # the first 300 picks against the literal argmax / zero-box loop
    assert np.array_equal(kp[:300], oracle.harris_nms(resp_o, 300, r))
    # properties of the greedy selection for all 10k picks
    s = resp_o[kp[:, 1], kp[:, 0]]
    assert (s > 0).all() and (np.diff(s) <= 0).all()                      # non-increasing, all real corners
    order = np.lexsort((kp[:, 0], kp[:, 1]))
    ys, xs = kp[order, 1].astype(np.int64), kp[order, 0].astype(np.int64)
    grid = {}
    for y, x in zip(ys, xs):
        cy, cx = y // (r + 1), x // (r + 1)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                for (yy, xx) in grid.get((cy + dy, cx + dx), ()):
                    assert max(abs(yy - y), abs(xx - x)) > r          # no pick inside another pick's box
        grid.setdefault((cy, cx), []).append((y, x))
    # a pick is the best pixel of its own box among pixels not suppressed by earlier picks
    for k in (0, 10, 999, 5000, 9999):
        x, y = kp[k]
        box = resp_o[max(0, y - r):y + r + 1, max(0, x - r):x + r + 1].copy()
        earlier = kp[:k]
        near = earlier[(np.abs(earlier[:, 0] - x) <= 2 * r) & (np.abs(earlier[:, 1] - y) <= 2 * r)]
        for ex, ey in near:
            y0, y1 = max(0, ey - r) - max(0, y - r), min(H, ey + r + 1) - max(0, y - r)
            x0, x1 = max(0, ex - r) - max(0, x - r), min(W, ex + r + 1) - max(0, x - r)
            box[max(0, y0):max(0, y1), max(0, x0):max(0, x1)] = 0
        assert box.max() == s[k]


def test_klt_large_frame_4_levels(ctx, big_pair):
    from vo import _ops
    a, b = big_pair
    kp, _, _ = _ops.harris_detect(a, 10000, ctx=ctx)
    pts = kp.astype(np.float32)
    nxt, st, err = _ops.klt_track(a, b, pts, win=17, max_level=3, ctx=ctx)
    nxt_o, st_o, err_o = oracle.klt_track(a, b, pts, win=17, max_level=3)
    assert np.array_equal(st, st_o) and np.array_equal(nxt, nxt_o) and np.array_equal(err, err_o)   # bit-exact
    good = (st == 1) & (err < 1.5)
    assert good.mean() > 0.6
    flow = nxt[good] - pts[good]
    assert np.abs(np.median(flow, axis=0) - np.array([3, 2], np.float32)).max() < 0.05              # the known shift
