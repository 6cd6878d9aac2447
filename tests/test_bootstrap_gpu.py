"""GPU: the two-view bootstrap (vo_bootstrap_*: cv2.findFundamentalMat's RANSAC restated + essential-matrix decomposition
+ cheirality vote + landmarks) against the REFERENCE's LandmarksTriangulator (tests/golden/bootstrap.npz) and the oracle.

Tolerances: the masks are integers -> equal; F to 1e-7 relative (the device takes the null space of the 7x9 system by
elimination, OpenCV by a Jacobi SVD: the solutions agree to rounding); [R | t] to 1e-7; landmarks to 1e-5 relative."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

TAGS = ("kitti_", "syn0_", "syn1_", "syn2_")


def test_bootstrap_equals_the_reference():
    from vo import _ops
    g = np.load(os.path.join(GOLDEN, "bootstrap.npz"))
    for tag in TAGS:
        r = _ops.bootstrap(g[tag + "p1"], g[tag + "p2"], g["K"], float(g[tag + "thr"]), float(g[tag + "conf"]))
        assert r["found"], tag
        assert np.array_equal(r["f_mask"], g[tag + "f_inl"]), (tag, int((r["f_mask"] != g[tag + "f_inl"]).sum()))
        assert np.abs(r["F"] - g[tag + "F"]).max() <= 1e-7 * np.abs(g[tag + "F"]).max(), tag
        assert np.array_equal(r["mask"], g[tag + "inl"]), tag
        assert np.abs(r["M"] - g[tag + "M"]).max() < 1e-7, tag
        ref = g[tag + "land"]
        assert np.nanmax(np.abs(r["landmarks"] - ref) / (1 + np.abs(ref))) < 1e-5, tag
        assert r["n_f_inliers"] == int(g[tag + "f_inl"].sum()) and r["n_valid"] == int(g[tag + "inl"].sum())


def test_bootstrap_drop_in_class():
    """LandmarksTriangulator(use_ransac=True, use_opencv=True) of the package: _find_fundamental_matrix_ransac and
    _find_relative_pose return what the reference's class returned (shapes included)."""
    from vo.landmarks import LandmarksTriangulator
    from vo.sensors import Camera
    g = np.load(os.path.join(GOLDEN, "bootstrap.npz"))
    cam = Camera(intrinsic_matrix=g["K"])
    tri = LandmarksTriangulator(camera1=cam, camera2=cam, use_ransac=True, use_opencv=True, outlier_ratio=0.9,
                                ransac_threshold=0.25, ransac_confidence=0.999)
    p1, p2 = g["kitti_p1"].reshape(-1, 2, 1), g["kitti_p2"].reshape(-1, 2, 1)
    F, inl = tri._find_fundamental_matrix_ransac(p1, p2)
    assert F.shape == (3, 3) and inl.dtype == bool and np.array_equal(inl, g["kitti_f_inl"])
    M, land, mask = tri._find_relative_pose(p1, p2)
    assert M.shape == (3, 4) and land.shape == (p1.shape[0], 3, 1) and np.array_equal(mask, g["kitti_inl"])
    assert np.abs(M - g["kitti_M"]).max() < 1e-7


def test_bootstrap_batched_random_vs_oracle():
    """Several sequences in one launch (ragged point counts) against the oracle; the accepted model, the iteration count
    and both masks must agree on every problem."""
    import ctypes as C
    import torch
    from oracle import bootstrap as ob
    from vo import _native as nat
    K = np.array([[718.856, 0, 607.19], [0, 718.856, 185.2], [0, 0, 1]])
    rng = np.random.default_rng(5)
    S, N = 6, 700
    P1, P2, n = np.zeros((S, N, 2)), np.zeros((S, N, 2)), np.array([700, 15, 333, 64, 500, 128], np.int32)
    for s in range(S):
        X = np.c_[rng.uniform(-10, 10, N), rng.uniform(-3, 3, N), rng.uniform(6, 40, N)]
        w = rng.uniform(-0.05, 0.05, 3)
        th = np.linalg.norm(w); k = w / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        t = np.array([0.1, -0.05, -1.0]) + rng.normal(0, 0.05, 3)
        a = (K @ X.T).T; b = (K @ (X @ R.T + t).T).T
        P1[s] = a[:, :2] / a[:, 2:] + rng.normal(0, 0.2, (N, 2))
        P2[s] = b[:, :2] / b[:, 2:] + rng.normal(0, 0.2, (N, 2))
        no = int(0.3 * n[s]); P2[s, :no] += rng.uniform(-40, 40, (no, 2))
    dev = torch.device("cuda", 0)
    ctx = nat.default_context(0)
    d1, d2, dn = torch.from_numpy(P1).to(dev), torch.from_numpy(P2).to(dev), torch.from_numpy(n).to(dev)
    dF = torch.empty((S, 9), dtype=torch.float64, device=dev); dM = torch.empty((S, 12), dtype=torch.float64, device=dev)
    dL = torch.empty((S, N, 3), dtype=torch.float64, device=dev)
    dm = torch.empty((S, N), dtype=torch.uint8, device=dev); df = torch.empty((S, N), dtype=torch.uint8, device=dev)
    di = torch.empty((S, 4), dtype=torch.int32, device=dev)
    K9 = np.ascontiguousarray(K.reshape(9))
    rc = nat.lib().vo_bootstrap_dev(ctx.handle, d1.data_ptr(), d2.data_ptr(), S, N, dn.data_ptr(), nat.ptr(K9), 0.5, 0.999, 1000,
                                    dF.data_ptr(), dM.data_ptr(), dL.data_ptr(), dm.data_ptr(), df.data_ptr(), di.data_ptr(),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream))
    nat.check(rc, "vo_bootstrap_dev")
    torch.cuda.synchronize()
    info, mask, fmask = di.cpu().numpy(), dm.cpu().numpy().astype(bool), df.cpu().numpy().astype(bool)
    for s in range(S):
        F, M, land, m, it = ob.bootstrap(P1[s, :n[s]], P2[s, :n[s]], K, 0.5, 0.999)
        assert info[s, 0] == 1 and info[s, 1] == it, (s, info[s], it)
        assert np.abs(dF[s].cpu().numpy().reshape(3, 3) - F).max() <= 1e-7 * np.abs(F).max(), s
        assert np.array_equal(mask[s, :n[s]], m) and not mask[s, n[s]:].any(), s
        assert np.abs(dM[s].cpu().numpy().reshape(3, 4) - M).max() < 1e-7, s
        got = dL[s, :n[s]].cpu().numpy()
        assert np.nanmax(np.abs(got - land) / (1 + np.abs(land))) < 1e-5, s
        assert fmask[s, :n[s]].sum() == info[s, 2]


def test_bootstrap_too_few_points_is_reported():
    from vo import _ops
    with pytest.raises(ValueError):
        _ops.bootstrap(np.zeros((10, 2)), np.zeros((10, 2)), np.eye(3), 0.25, 0.999)


def test_bootstrap_edge_cases_vs_oracle():
    """The smallest population OpenCV's RANSAC path accepts (15), duplicated matches (getSubset's collinearity / proximity
    check redraws), a planar scene (F not unique) and gross outliers only (few inliers): iteration counts and masks equal
    to the oracle's."""
    from oracle import bootstrap as ob
    from vo import _ops
    K = np.array([[718.856, 0, 607.19], [0, 718.856, 185.2], [0, 0, 1]])
    rng = np.random.default_rng(11)

    def scene(N, planar=False, noise=0.2, n_out=0):
        X = np.c_[rng.uniform(-10, 10, N), np.full(N, 1.6) if planar else rng.uniform(-3, 3, N), rng.uniform(6, 40, N)]
        w = rng.uniform(-0.05, 0.05, 3); th = np.linalg.norm(w); k = w / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        t = np.array([0.1, -0.05, -1.0]) + rng.normal(0, 0.05, 3)
        a = (K @ X.T).T; b = (K @ (X @ R.T + t).T).T
        p1 = a[:, :2] / a[:, 2:] + rng.normal(0, noise, (N, 2)); p2 = b[:, :2] / b[:, 2:] + rng.normal(0, noise, (N, 2))
        p2[:n_out] += rng.uniform(-40, 40, (n_out, 2))
        return p1, p2
    cases = {"fifteen": scene(15), "planar": scene(400, planar=True, n_out=80), "mostly_outliers": scene(200, n_out=170)}
    p1, p2 = scene(120, n_out=20)
    cases["duplicates"] = (np.r_[p1, p1[:60]], np.r_[p2, p2[:60]])
    for name, (p1, p2) in cases.items():
        r = _ops.bootstrap(p1, p2, K, 0.5, 0.999)
        F, M, land, mask, it = ob.bootstrap(p1, p2, K, 0.5, 0.999)
        if F is None:
            assert not r["found"], name
            continue
        assert r["found"] and r["iterations"] == it, (name, r["iterations"], it)
        f_inl = ob.cv_fm_errors_f32(F, p1.astype(np.float32), p2.astype(np.float32)) <= np.float32(0.25)
        assert np.array_equal(r["f_mask"], f_inl), name
        assert np.array_equal(r["mask"], mask), name
        if name != "planar":                                 # on a plane E is ill-conditioned: the masks agree, the pose need not to 1e-7
            assert np.abs(r["M"] - M).max() < 1e-6, name
