"""GPU parity: batched P3P + RANSAC scoring (C ABI) vs the oracle (bit-exact models, counts, masks)
and vs the reference's own estimator run (tests/golden/p3p.npz)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _rot_angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


def _samples(N, n, seed=2023):
    rng = np.random.default_rng(seed)
    return np.array([rng.choice(np.arange(N), replace=False, size=4) for _ in range(n)], np.int32)


@pytest.mark.parametrize("tag", ["clean", "noisy"])
def test_score_bitexact_vs_oracle(ctx, golden, tag):
    from vo import _ops
    g = golden("p3p")
    L, P, thr = g[f"{tag}_landmarks"], g[f"{tag}_keypoints"], float(g[f"{tag}_threshold"])
    N = L.shape[0]
    S = _samples(N, 1500)
    table = oracle.ransac_iterations_table(N, 4, 0.99, 1000)
    init = oracle.ransac_initial_iterations(4, 0.9, 0.99, 1000)
    r = _ops.p3p_ransac(L, P, g["K"], S, thr, table, init, want_all=True, ctx=ctx)
    models, valid, counts = oracle.p3p_ransac_score(L, P, g["K"], S, thr)
    assert np.array_equal(r["valid"], valid)
    assert np.array_equal(r["counts"], counts)
    assert np.array_equal(r["models"], models)              # float64 bit-exact
    best_h, consumed, n_iter, n, best, exhausted = oracle.ransac_scan(valid, counts, table, init)
    assert (int(r["best"]), int(r["consumed"]), int(r["n_iterations"]), int(r["n"]), int(r["best_count"]),
            bool(r["exhausted"])) == (best_h, consumed, n_iter, n, best, exhausted)
    inl = oracle.reproj_errors(models[best_h, :9], models[best_h, 9:], L, P, g["K"]) < thr
    assert np.array_equal(r["inliers"], inl)                # bit-exact inlier mask
    # ... and against the reference's own estimator (P3PPoseEstimator, use_opencv=False)
    assert np.array_equal(r["inliers"], g[f"{tag}_refine0_inliers"])
    assert _rot_angle(r["R"], g[f"{tag}_refine0_R"]) < 1e-5                       # 1e-5 rad
    t_ref = g[f"{tag}_refine0_t"]
    assert np.linalg.norm(r["t"] - t_ref) <= 1e-4 * np.linalg.norm(t_ref)        # 1e-4 relative
    assert int(r["n_iterations"]) == int(g[f"{tag}_refine0_n_iterations"])


def test_solver_vs_cv2_golden(ctx, golden):
    from vo import _ops
    g = golden("p3p")
    for tag in ("clean", "noisy"):
        L, P, S = g[f"{tag}_landmarks"], g[f"{tag}_keypoints"], g[f"{tag}_sample_idx"]
        N = L.shape[0]
        table = np.full(N + 1, 10 ** 6, np.int32)
        r = _ops.p3p_ransac(L, P, g["K"], S, 1.0, table, 10 ** 6, want_all=True, ctx=ctx)
        cvm = g[f"{tag}_cv_models"]
        cv_ok = g[f"{tag}_cv_valid"].astype(bool) & np.isfinite(cvm).all(axis=1)
        assert np.array_equal(r["valid"].astype(bool), cv_ok)
        for h in np.where(cv_ok)[0]:
            assert _rot_angle(r["models"][h, :9].reshape(3, 3), cvm[h, :9].reshape(3, 3)) < 1e-5
            assert np.linalg.norm(r["models"][h, 9:] - cvm[h, 9:]) <= 1e-4 * np.linalg.norm(cvm[h, 9:])


def test_batched_frames_and_continuation(ctx, golden):
    """Frames are independent; a run split into two batches (state carried) equals one run."""
    from vo import _ops
    g = golden("p3p")
    L, P, thr = g["noisy_landmarks"], g["noisy_keypoints"], float(g["noisy_threshold"])
    N = L.shape[0]
    table = oracle.ransac_iterations_table(N, 4, 0.99, 1000)
    init = oracle.ransac_initial_iterations(4, 0.9, 0.99, 1000)
    S = _samples(N, 512, seed=5)
    Lb = np.stack([L, L[::-1].copy()])
    Pb = np.stack([P, P[::-1].copy()])
    Sb = np.stack([S, S])
    rb = _ops.p3p_ransac(Lb, Pb, g["K"], Sb, thr, table, init, want_all=True, ctx=ctx)
    r0 = _ops.p3p_ransac(L, P, g["K"], S, thr, table, init, want_all=True, ctx=ctx)
    assert np.array_equal(rb["counts"][0], r0["counts"]) and np.array_equal(rb["inliers"][0], r0["inliers"])
    # split: the first 8 hypotheses, then the rest with the loop state carried over
    k = 8
    a = _ops.p3p_ransac(L, P, g["K"], S[:k], thr, table, init, ctx=ctx)
    assert bool(a["exhausted"]) and int(a["consumed"]) == k
    b = _ops.p3p_ransac(L, P, g["K"], S[k:], thr, table, int(a["n_iterations"]), start_n=int(a["n"]),
                        start_best=int(a["best_count"]), ctx=ctx)
    total_best = (k + int(b["best"])) if int(b["best"]) >= 0 else int(a["best"])
    assert total_best == int(r0["best"])
    assert k + int(b["consumed"]) == int(r0["consumed"])


def test_large_sweep_property(ctx):
    """BASELINE configs[2]-sized sweep: the winner's mask must equal the count, and the count must be
    the maximum over hypotheses scanned (no adaptive stop)."""
    from vo import _ops
    rng = np.random.default_rng(3)
    N, Hn = 2500, 16384
    K = np.array([[707.0912, 0, 601.8873], [0, 707.0912, 183.1104], [0, 0, 1.0]])
    L = rng.uniform(-10, 10, (N, 3))
    L[:, 2] = rng.uniform(4, 40, N)
    uv = (K @ L.T).T
    uv = uv[:, :2] / uv[:, 2:] + rng.normal(0, 0.3, (N, 2))
    out = rng.choice(N, N // 3, replace=False)
    uv[out] += rng.uniform(-80, 80, (len(out), 2))
    S = rng.integers(0, N, (Hn, 4)).astype(np.int32)
    table = np.full(N + 1, 10 ** 7, np.int32)
    r = _ops.p3p_ransac(L, uv, K, S, 1.5, table, 10 ** 7, want_all=True, ctx=ctx)
    assert int(r["consumed"]) == Hn and bool(r["exhausted"])
    assert int(r["best_count"]) == int(r["counts"][r["valid"].astype(bool)].max())
    assert int(r["inliers"].sum()) == int(r["best_count"])
    assert int(r["best"]) == int(np.argmax(np.where(r["valid"].astype(bool), r["counts"], -1)))
    assert int(r["best_count"]) > N // 2
    # spot-check 64 hypotheses against the oracle
    sel = rng.choice(Hn, 64, replace=False)
    m, v, c = oracle.p3p_ransac_score(L, uv, K, S[sel], 1.5)
    assert np.array_equal(r["counts"][sel], c) and np.array_equal(r["models"][sel], m)


def test_full_size_sweep_3000x65536(ctx):
    """BASELINE configs[2] at its upper end: 3000 correspondences x 65536 hypotheses in one call.  1024 hypotheses
    are checked against the oracle bit for bit (models, validity, counts); the replay over all of them must pick
    the first maximum, and its mask must have exactly that many inliers."""
    from vo import _ops
    rng = np.random.default_rng(17)
    N, Hn = 3000, 65536
    K = np.array([[707.0912, 0, 601.8873], [0, 707.0912, 183.1104], [0, 0, 1.0]])
    L = rng.uniform(-12, 12, (N, 3))
    L[:, 2] = rng.uniform(4, 60, N)
    uv = (K @ L.T).T
    uv = uv[:, :2] / uv[:, 2:] + rng.normal(0, 0.4, (N, 2))
    out = rng.choice(N, int(0.4 * N), replace=False)
    uv[out] += rng.uniform(-100, 100, (len(out), 2))
    S = np.argsort(rng.random((Hn, 16)), axis=1)[:, :4].astype(np.int32)          # 4 distinct numbers ...
    S = (S + rng.integers(0, N - 16, (Hn, 1))).astype(np.int32)                    # ... shifted to a random window
    table = np.full(N + 1, 10 ** 7, np.int32)
    r = _ops.p3p_ransac(L, uv, K, S, 1.25, table, 10 ** 7, want_all=True, ctx=ctx)
    valid = r["valid"].astype(bool)
    assert int(r["consumed"]) == Hn and bool(r["exhausted"]) and int(r["n"]) == int(valid.sum())
    assert int(r["best"]) == int(np.argmax(np.where(valid, r["counts"], -1)))
    assert int(r["best_count"]) == int(r["counts"][valid].max()) == int(r["inliers"].sum())
    sel = np.sort(rng.choice(Hn, 1024, replace=False))
    m, v, c = oracle.p3p_ransac_score(L, uv, K, S[sel], 1.25)
    assert np.array_equal(r["valid"][sel], v) and np.array_equal(r["counts"][sel], c) and np.array_equal(r["models"][sel], m)
    # OpenCV's rule (float32 points, float32 errors, err <= thr^2): counts equal to the restated computeError / findInliers
    L32 = L.astype(np.float32).astype(np.float64)
    P32 = uv.astype(np.float32)
    r2 = _ops.p3p_ransac(L32, P32.astype(np.float64), K, S[:512], 1.25 ** 2, table, 10 ** 7, want_all=True, inclusive=True, ctx=ctx)
    for h in range(0, 512, 8):
        if r2["valid"][h]:
            e = oracle.cv_reproj_errors_f32(r2["models"][h][:9].reshape(3, 3), r2["models"][h][9:], L32, P32, K)
            assert int((e <= np.float32(1.25 ** 2)).sum()) == int(r2["counts"][h]), h


def test_refine_pose_matches_restatement_and_beats_reference_cost(ctx, golden):
    """vo_refine_pose_host (p3p.py:188-213 on the device) against its numpy restatement (oracle.loop.refine_gn) and
    against the reference's own refined pose on its frame-3 problem: same minimum (1e-9), never a higher cost than
    where scipy stopped."""
    from oracle.loop import _cost, refine_gn
    from vo import _ops
    g = golden("loop")
    inl = g["f3_inliers"]
    N = len(inl)
    kp, land = g["f3_kp"][:N][inl], g["f3_land"][:N][inl]
    ok = ~np.isnan(land).any(1)
    kp, land = kp[ok].astype(np.float64), land[ok]
    K = g["K"].astype(np.float64)
    f = (K[0, 0], K[1, 1], K[0, 2], K[1, 2])
    R0, t0 = g["noref_f3_R"], g["noref_f3_t"]
    R, t, iters = _ops.refine_pose(land, kp, K, R0, t0, ctx=ctx)
    Rg, tg = refine_gn(R0, t0, land, kp, K)
    assert np.abs(R - Rg).max() < 1e-9 and np.abs(t - tg).max() < 1e-8 and 1 <= iters <= 30
    assert _cost(R, t.ravel(), land, kp, *f) <= _cost(g["f3_R"], g["f3_t"], land, kp, *f)
    assert np.abs(R - g["f3_R"]).max() < 2e-4 and np.abs(t.ravel() - g["f3_t"]).max() < 5e-3
    # a mask restricts the problem: refining on the first half equals refining the first half alone
    m = np.zeros(len(land), np.uint8)
    m[: len(land) // 2] = 1
    Ra, ta, _ = _ops.refine_pose(land, kp, K, R0, t0, mask=m, ctx=ctx)
    Rb, tb, _ = _ops.refine_pose(land[: len(land) // 2], kp[: len(land) // 2], K, R0, t0, ctx=ctx)
    assert np.abs(Ra - Rb).max() < 1e-10 and np.abs(ta - tb).max() < 1e-9


def test_default_path_equals_cv2_solvepnpransac(ctx):
    """P3PPoseEstimator(use_opencv=True) -- the reference's default and main.py's configuration -- must return
    cv2.solvePnPRansac's own inlier mask (OpenCV's RANSAC restated: cv::RNG subsets and loop on the host, models and
    float32-rule counts on the GPU), on random problems of different sizes and outlier ratios.  The minimal solver is
    OpenCV's only to ~1e-7 rad, so a point whose squared error lies within 1e-3 of the threshold may fall on the other
    side (one point in one of these problems): the masks must be identical on most problems and never differ elsewhere."""
    import cv2
    from vo.pose_estimation import P3PPoseEstimator
    from vo.primitives import Features
    rng = np.random.default_rng(3)
    K = np.array([[707.0912, 0, 601.8873], [0, 707.0912, 183.1104], [0, 0, 1.0]])
    n_exact = 0
    for trial in range(12):
        N = int(rng.integers(60, 700))
        L = rng.uniform(-10, 10, (N, 3, 1))
        L[:, 2] = rng.uniform(4, 50, (N, 1))
        R = cv2.Rodrigues(rng.normal(0, 0.1, 3).reshape(3, 1))[0]
        cam = L[:, :, 0] @ R.T + rng.normal(0, 0.5, 3)
        uv = cam @ K.T
        P = uv[:, :2] / uv[:, 2:] + rng.normal(0, 0.4, (N, 2))
        out = rng.choice(N, int(rng.uniform(0.05, 0.5) * N), replace=False)
        P[out] += rng.uniform(-50, 50, (len(out), 2))
        P = P.astype(np.float32).reshape(N, 2, 1)
        ok, rv, tv, inl = cv2.solvePnPRansac(L, P, K, None, flags=cv2.SOLVEPNP_P3P, iterationsCount=10000, reprojectionError=1.25,
                                             confidence=0.9999)
        want = np.zeros(N, bool)
        want[inl.ravel()] = True
        est = P3PPoseEstimator(intrinsic_matrix=K, inlier_threshold=1.25, use_opencv=True, confidence=0.9999, nonlinear_refinement=True)
        (Rg, tg), mask = est.estimate_pose(Features(keypoints=P, landmarks=L))
        n_exact += int(np.array_equal(mask, want))
        if not np.array_equal(mask, want):
            d = np.flatnonzero(mask != want)
            assert len(d) <= 2, (trial, len(d))
            m0, mk, _ = oracle.cv_solve_pnp_ransac_p3p(L, P, K, 1.25, 0.9999, 10000)
            assert np.array_equal(mk, mask), trial                       # the GPU path is the restatement, bit for bit
            e = oracle.cv_reproj_errors_f32(m0[0], m0[1], L[:, :, 0].astype(np.float32).astype(np.float64), P[:, :, 0], K)
            assert np.all(np.abs(e[d] - 1.25 ** 2) < 1e-3), (trial, e[d])
        # refined pose = minimum of the reprojection cost over cv2's inliers: no worse than cv2's own (EPnP-refitted) result
        def cost(Rm, tm):
            c = L[want, :, 0] @ Rm.T + np.asarray(tm).reshape(3)
            p = c @ K.T
            return float((((p[:, :2] / p[:, 2:]) - P[want, :, 0]) ** 2).sum())
        assert cost(Rg, tg) <= cost(cv2.Rodrigues(rv)[0], tv) * (1 + 1e-6)
    assert n_exact >= 10
