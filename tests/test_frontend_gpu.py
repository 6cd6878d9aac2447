"""GPU: the resident front-end object (vo_frontend_*) must produce, for every sequence, exactly what
the individual entry points produce (which the other tests pin to the oracle / reference)."""
import numpy as np
import pytest

from conftest import synthetic_image

pytestmark = pytest.mark.gpu


def _inputs(S, H, W, N, Hn, T, seed):
    rng = np.random.default_rng(seed)
    K = np.array([[400.0, 0, W / 2], [0, 400.0, H / 2], [0, 0, 1]])
    L = rng.uniform(-3, 3, (S, N, 3)); L[..., 2] += 9
    uv = np.einsum("ij,snj->sni", K, L); uv = uv[..., :2] / uv[..., 2:] + rng.normal(0, 0.3, (S, N, 2))
    uv[:, : N // 4] += rng.uniform(-40, 40, (S, N // 4, 2))
    samples = np.stack([np.argsort(rng.random((Hn, N)), axis=1)[:, :4] for _ in range(S)]).astype(np.int32)
    X = rng.uniform(-3, 3, (S, T, 3)); X[..., 2] += 10
    Xh = np.concatenate([X, np.ones((S, T, 1))], -1)
    proj2 = np.stack([K @ np.hstack([np.eye(3), np.array([[-0.6], [0.0], [0.05]])])] * S)
    M = np.repeat(np.hstack([np.eye(3), np.zeros((3, 1))])[None, None], S, 0).repeat(T, 1)
    M[..., 3] += rng.normal(0, 0.05, (S, T, 3))
    proj1 = K @ M
    a = np.einsum("stij,stj->sti", proj1, Xh); b = np.einsum("sij,stj->sti", proj2, Xh)
    p1 = a[..., :2] / a[..., 2:]; p2 = b[..., :2] / b[..., 2:]
    return dict(K=K, landmarks=L, kp2d=uv, samples=samples, tri_p1=np.ascontiguousarray(p1), tri_p2=np.ascontiguousarray(p2),
                tri_proj1=np.ascontiguousarray(proj1.reshape(S, T, 12)), tri_proj2=np.ascontiguousarray(proj2.reshape(S, 12)), X=X)


@pytest.mark.parametrize("use_prefetch", [False, True])
def test_frontend_matches_individual_ops(ctx, use_prefetch):
    from vo import _ops
    from vo.frontend import Frontend
    S, H, W, KP, N, Hn, T = 3, 120, 200, 150, 300, 128, 64
    big = [synthetic_image(H + 20, W + 20, seed=40 + s) for s in range(S)]
    frames = [np.stack([np.ascontiguousarray(b[10 + t:10 + t + H, 10 + 2 * t:10 + 2 * t + W]) for b in big]) for t in range(3)]
    g = _inputs(S, H, W, N, Hn, T, seed=5)
    table = np.full(N + 1, 10 ** 6, np.int32)
    K9 = np.ascontiguousarray(g["K"].reshape(9))
    fe = Frontend(S, H, W, num_keypoints=KP, n_corr=N, n_hyp=Hn, p3p_threshold=1.5, n_tri=T, tri_mode=1, ctx=ctx)
    outs = dict(kp_xy=np.zeros((S, KP, 2), np.int32), tracked=np.zeros((S, KP, 2), np.float32), status=np.zeros((S, KP), np.uint8),
                err=np.zeros((S, KP), np.float32), best4=np.zeros((S, 4), np.int32), inliers=np.zeros((S, N), np.uint8),
                pose=np.zeros((S, 12)), tri_out=np.zeros((S * T, 3)))
    args = (g["landmarks"], g["kp2d"], g["samples"], table, g["tri_p1"], g["tri_p2"], g["tri_proj1"], g["tri_proj2"])
    prev_kp = None
    if use_prefetch:
        fe.prefetch_host(frames[0], *args)
    for t in range(3):
        if use_prefetch:
            if t + 1 < 3:
                fe.prefetch_host(frames[t + 1], *args)
            fe.step_host(None, None, None, K9, None, None, 10 ** 6, None, None, None, None, outs)
        else:
            fe.step_host(frames[t], g["landmarks"], g["kp2d"], K9, g["samples"], table, 10 ** 6, g["tri_p1"], g["tri_p2"],
                         g["tri_proj1"], g["tri_proj2"], outs)
        kp_ref, _, _ = _ops.harris_detect(frames[t], KP, ctx=ctx)
        assert np.array_equal(outs["kp_xy"], kp_ref)
        if prev_kp is not None:
            nxt, st, err = _ops.klt_track(frames[t - 1], frames[t], prev_kp.astype(np.float32), ctx=ctx)
            assert np.array_equal(outs["status"], st) and np.array_equal(outs["tracked"], nxt) and np.array_equal(outs["err"], err)
        prev_kp = kp_ref
        r = _ops.p3p_ransac(g["landmarks"], g["kp2d"], g["K"], g["samples"], 1.5, table, 10 ** 6, ctx=ctx)
        assert np.array_equal(outs["best4"][:, 0], r["best"]) and np.array_equal(outs["inliers"].astype(bool), r["inliers"])
        assert np.array_equal(outs["pose"][:, :9].reshape(S, 3, 3), r["R"])
        for s in range(S):
            X = _ops.triangulate(g["tri_p1"][s], g["tri_p2"][s], g["tri_proj1"][s].reshape(T, 3, 4), g["tri_proj2"][s].reshape(3, 4), mode=1, ctx=ctx)
            assert np.array_equal(outs["tri_out"][s * T:(s + 1) * T], X)
            assert np.allclose(X, g["X"][s], atol=1e-6)
    fe.close()


def test_frontend_pipelined_matches_sequential(ctx):
    """prefetch + submit + wait with two steps in flight gives, step for step, what step_host gives"""
    from vo.frontend import Frontend
    S, H, W, KP, N, Hn, T, n_steps = 2, 120, 200, 150, 300, 128, 64, 6
    big = [synthetic_image(H + 30, W + 30, seed=70 + s) for s in range(S)]
    frames = [np.stack([np.ascontiguousarray(b[10 + t:10 + t + H, 10 + 2 * t:10 + 2 * t + W]) for b in big]) for t in range(n_steps)]
    g = _inputs(S, H, W, N, Hn, T, seed=9)
    table = np.full(N + 1, 10 ** 6, np.int32)
    K9 = np.ascontiguousarray(g["K"].reshape(9))
    args = (g["landmarks"], g["kp2d"], g["samples"], table, g["tri_p1"], g["tri_p2"], g["tri_proj1"], g["tri_proj2"])

    def new_outs():
        return dict(kp_xy=np.zeros((S, KP, 2), np.int32), tracked=np.zeros((S, KP, 2), np.float32), status=np.zeros((S, KP), np.uint8),
                    err=np.zeros((S, KP), np.float32), best4=np.zeros((S, 4), np.int32), inliers=np.zeros((S, N), np.uint8),
                    pose=np.zeros((S, 12)), tri_out=np.zeros((S * T, 3)))

    ref = []
    fe = Frontend(S, H, W, num_keypoints=KP, n_corr=N, n_hyp=Hn, p3p_threshold=1.5, n_tri=T, tri_mode=1, ctx=ctx)
    for t in range(n_steps):
        o = new_outs()
        fe.step_host(frames[t], g["landmarks"], g["kp2d"], K9, g["samples"], table, 10 ** 6, g["tri_p1"], g["tri_p2"],
                     g["tri_proj1"], g["tri_proj2"], o)
        ref.append(o)
    fe.close()

    fe = Frontend(S, H, W, num_keypoints=KP, n_corr=N, n_hyp=Hn, p3p_threshold=1.5, n_tri=T, tri_mode=1, ctx=ctx)
    got = [new_outs() for _ in range(n_steps)]
    fe.prefetch_host(frames[0], *args)
    for t in range(n_steps):
        if t + 1 < n_steps:
            fe.prefetch_host(frames[t + 1], *args)
        fe.submit_host(None, None, None, K9, None, None, 10 ** 6, None, None, None, None, got[t])
        if t > 0:
            fe.wait_host()
    fe.wait_host()
    with pytest.raises(Exception):
        fe.wait_host()                               # nothing in flight any more
    fe.close()
    for t in range(n_steps):
        for k in ref[t]:
            if t == 0 and k in ("tracked", "status", "err"):
                continue                             # no previous keypoints at the first step
            assert np.array_equal(got[t][k], ref[t][k]), (t, k)


def test_frontend_submit_limits(ctx):
    """at most two submitted steps in flight; step_host refuses to run while submits are pending"""
    from vo.frontend import Frontend
    from vo._native import VoNativeError
    S, H, W, KP, N, Hn, T = 1, 96, 128, 50, 100, 32, 16
    frame = np.stack([synthetic_image(H, W, seed=3)])
    g = _inputs(S, H, W, N, Hn, T, seed=2)
    table = np.full(N + 1, 10 ** 6, np.int32)
    K9 = np.ascontiguousarray(g["K"].reshape(9))
    fe = Frontend(S, H, W, num_keypoints=KP, n_corr=N, n_hyp=Hn, p3p_threshold=1.5, n_tri=T, tri_mode=1, ctx=ctx)

    def outs():
        return dict(kp_xy=np.zeros((S, KP, 2), np.int32), tracked=np.zeros((S, KP, 2), np.float32), status=np.zeros((S, KP), np.uint8),
                    err=np.zeros((S, KP), np.float32), best4=np.zeros((S, 4), np.int32), inliers=np.zeros((S, N), np.uint8),
                    pose=np.zeros((S, 12)), tri_out=np.zeros((S * T, 3)))

    args = (frame, g["landmarks"], g["kp2d"], K9, g["samples"], table, 10 ** 6, g["tri_p1"], g["tri_p2"], g["tri_proj1"], g["tri_proj2"])
    o = [outs() for _ in range(3)]
    fe.submit_host(*args, o[0])
    fe.submit_host(*args, o[1])
    with pytest.raises(VoNativeError):
        fe.submit_host(*args, o[2])
    with pytest.raises(VoNativeError):
        fe.step_host(*args, o[2])
    fe.wait_host()
    fe.wait_host()
    assert np.array_equal(o[0]["kp_xy"], o[1]["kp_xy"])     # same frame twice: same detections
    fe.step_host(*args, o[2])
    assert np.array_equal(o[2]["kp_xy"], o[0]["kp_xy"])
    fe.close()
