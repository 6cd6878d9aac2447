"""CPU: pin the oracle to the reference's own outputs (tests/golden/*.npz, produced by
tests/golden/make_golden.py importing /root/reference)."""
import numpy as np

import oracle


def test_harris_response_bitexact_vs_reference(golden):
    g = golden("harris")
    resp = oracle.harris_response(g["crop"], 9, 0.09)
    assert resp.dtype == np.float64
    assert np.array_equal(resp, g["crop_resp"])  # bit-exact float64


def test_harris_keypoints_vs_reference(golden):
    g = golden("harris")
    for K, r in [(150, 5), (400, 3)]:
        kp, _ = oracle.harris_keypoints(g["crop"], K, 9, 0.09, r)
        assert np.array_equal(kp, g[f"crop_kp_K{K}_r{r}"])
        desc = oracle.harris_descriptors(g["crop"], kp, 9)
        assert np.array_equal(desc.astype(np.uint8), g[f"crop_desc_K{K}_r{r}"])


def test_harris_zero_fill_vs_reference(golden):
    g = golden("harris")
    kp, _ = oracle.harris_keypoints(g["blank"], 40)
    assert np.array_equal(kp, g["blank_kp40"])
    assert (kp[-1] == 0).all()  # harris.py:149 returns index 0 once every score is zero


# ---------------------------------------------------------------------------- KLT
def test_pyr_down_bitexact_vs_cv2(golden):
    g = golden("klt")
    p1 = oracle.pyr_down(g["prev"])
    assert np.array_equal(p1, g["pyr1"])
    assert np.array_equal(oracle.pyr_down(p1), g["pyr2"])


def _klt_compare(nxt, st, err, g, suffix=""):
    st_ref = g["status" + suffix]
    assert np.array_equal(st, st_ref)                       # status identical (incl. out-of-image points)
    ok = st_ref == 1
    d = np.abs(nxt - g["next_pts" + suffix]).max(axis=1)[ok]
    assert d.max() < 1e-2, d.max()                          # BASELINE tolerance: 1e-2 px
    assert np.median(d) < 1e-4
    assert np.abs(err - g["err" + suffix])[ok].max() < 5e-2


def test_klt_vs_cv2(golden):
    g = golden("klt")
    nxt, st, err = oracle.klt_track(g["prev"], g["next"], g["pts"])
    _klt_compare(nxt, st, err, g)
    nxt, st, err = oracle.klt_track(g["prev"], g["next"], g["pts"], win=21, max_level=3, max_iters=30, epsilon=0.01)
    _klt_compare(nxt, st, err, g, "_w21_l3")


# ---------------------------------------------------------------------------- P3P
def _rot_angle(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return float(np.arccos(np.clip(c, -1, 1)))


def test_p3p_solver_vs_cv2(golden):
    g = golden("p3p")
    for tag in ("clean", "noisy"):
        L, P, S = g[f"{tag}_landmarks"], g[f"{tag}_keypoints"], g[f"{tag}_sample_idx"]
        for h in range(S.shape[0]):
            m = oracle.p3p_solve4(L[S[h]], P[S[h]], g["K"])
            cvm = g[f"{tag}_cv_models"][h]
            # cv2 4.13 flags some samples without a real solution as "success" with a NaN pose;
            # a usable cv2 model is one that is both flagged valid and finite.
            cv_ok = bool(g[f"{tag}_cv_valid"][h]) and bool(np.isfinite(cvm).all())
            assert (m is not None) == cv_ok
            if m is None:
                continue
            assert _rot_angle(m[0], cvm[:9].reshape(3, 3)) < 1e-5            # BASELINE: 1e-5 rad
            assert np.linalg.norm(m[1].ravel() - cvm[9:]) <= 1e-4 * np.linalg.norm(cvm[9:])  # 1e-4 relative


def test_p3p_ransac_vs_reference(golden):
    """The reference's own estimator run (rng 2023, adaptive stop, state carried into a second call)."""
    g = golden("p3p")
    for tag in ("clean", "noisy"):
        L, P, thr = g[f"{tag}_landmarks"], g[f"{tag}_keypoints"], float(g[f"{tag}_threshold"])
        r = oracle.RansacP3P(g["K"], thr, 0.9, 0.99, 1000)
        for call in ("", "_second"):
            model, inl = r.find_best_model(L, P)
            assert np.array_equal(inl, g[f"{tag}_refine0{call}_inliers"])    # bit-exact mask
            assert _rot_angle(model[0], g[f"{tag}_refine0{call}_R"]) < 1e-5
            t_ref = g[f"{tag}_refine0{call}_t"]
            assert np.linalg.norm(model[1] - t_ref) <= 1e-4 * np.linalg.norm(t_ref)
            if call == "":
                assert r.n_iterations == int(g[f"{tag}_refine0_n_iterations"])
                assert r.outlier_ratio == float(g[f"{tag}_refine0_outlier_ratio"])


def test_p3p_batched_scan_equals_sequential(golden):
    g = golden("p3p")
    L, P, thr = g["noisy_landmarks"], g["noisy_keypoints"], float(g["noisy_threshold"])
    N = L.shape[0]
    seq = oracle.RansacP3P(g["K"], thr, 0.9, 0.99, 1000)
    model, inl = seq.find_best_model(L, P)
    rng = np.random.default_rng(2023)
    S = np.array([rng.choice(np.arange(N), replace=False, size=4) for _ in range(1000)], np.int32)
    models, valid, counts = oracle.p3p_ransac_score(L, P, g["K"], S, thr)
    table = oracle.ransac_iterations_table(N, 4, 0.99, 1000)
    init = oracle.ransac_initial_iterations(4, 0.9, 0.99, 1000)
    best_h, consumed, n_iter, n, best, exhausted = oracle.ransac_scan(valid, counts, table, init)
    assert not exhausted and consumed == seq.draws and n_iter == seq.n_iterations
    assert np.array_equal(models[best_h, :9].reshape(3, 3), model[0])
    assert best == inl.sum()


# ---------------------------------------------------------------------------- triangulation
def test_triangulation_vs_reference(golden):
    g = golden("triangulation")
    for tag in ("clean", "noisy"):
        X = oracle.triangulate(g[f"{tag}_p1"], g[f"{tag}_p2"], g["C1"], g["C2"], mode=0)
        assert np.array_equal(X, g[f"{tag}_linear"])       # same numpy SVD -> identical
        K = g["C1"][:, :3]
        proj1 = K @ np.linalg.inv(g[f"{tag}_cand_poses"])[:, :3]
        proj2 = K @ np.linalg.inv(g[f"{tag}_cand_current_pose"])[:3]
        X0 = oracle.triangulate(g[f"{tag}_cand_tracks"], g[f"{tag}_p2"], proj1, proj2, mode=0)
        X1 = oracle.triangulate(g[f"{tag}_cand_tracks"], g[f"{tag}_p2"], proj1, proj2, mode=1)
        assert np.allclose(X0, g[f"{tag}_cand_cv0"], rtol=0, atol=1e-9)
        assert np.allclose(X1, g[f"{tag}_cand_cv1"], rtol=1e-7, atol=1e-7)   # cv2's float64 SVD


def test_match_descriptors_vs_reference(golden):
    g = golden("harris")
    pairs = oracle.match_descriptors(g["match_desc1"], g["match_desc2"])
    assert np.array_equal(pairs, g["match_pairs"])          # cv2.BFMatcher + the reference's loop
    assert len(pairs) == int(g["match_n_matched"])
