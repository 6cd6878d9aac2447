"""GPU: the drop-in `vo` classes, written like the reference's own tests (tests/test_harris.py,
tests/test_p3p.py, tests/test_triangulation.py) plus parity against arrays the reference produced."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rot_angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


def _cameras():
    from vo.sensors import Camera
    K = np.array([[500, 0, 320], [0, 500, 240], [0, 0, 1]], dtype=float)
    th1, th2 = np.pi / 8, np.pi / 32
    R = np.array([[np.cos(th1), -np.sin(th1), 0], [np.sin(th1), np.cos(th1), 0], [0, 0, 1]])
    R = R @ np.array([[np.cos(th2), 0, np.sin(th2)], [0, 1, 0], [-np.sin(th2), 0, np.cos(th2)]])
    return Camera(K, R=np.eye(3), t=np.zeros((3, 1))), Camera(K, R=R, t=np.array([[1, 1, -1]], dtype=float).T)


# ---------------------------------------------------------------------------- Harris (tests/test_harris.py)
def test_harris_detector_on_kitti_crops(ctx, golden):
    from vo.features import HarrisCornerDetector
    from vo.primitives import Frame, Matches
    g, gh = golden("klt"), golden("harris")
    f1, f2 = Frame(np.stack([g["prev"]] * 3, -1)), Frame(np.stack([g["next"]] * 3, -1))
    det = HarrisCornerDetector(f1, num_keypoints=200)
    m = det.featureMatcher(f1, f2)
    assert isinstance(m, Matches)
    assert m.frame1.features.keypoints.shape[0] <= 200 and m.frame1.features.descriptors.shape[0] <= 200
    assert m.frame1.features.descriptors.shape == m.frame2.features.descriptors.shape
    n_matched = int((m.frame1.features.state == 1).sum())
    assert n_matched > 20
    # the first frame moved by about a pixel: matched keypoints must be close
    d = np.abs(m.frame1.features.matched_inliers_keypoints - m.frame2.features.matched_inliers_keypoints).reshape(-1, 2)
    assert np.median(d.max(axis=1)) <= 3
    # keypoints / descriptors are the reference's (bit-exact)
    det = HarrisCornerDetector(num_keypoints=150)
    fr = det.extractDescriptors(det.extractKeypoints(Frame(gh["crop"].copy())))
    assert np.array_equal(fr.features.keypoints.reshape(-1, 2).astype(np.int32), gh["crop_kp_K150_r5"])
    assert np.array_equal(fr.features.descriptors.reshape(150, -1).astype(np.uint8), gh["crop_desc_K150_r5"])
    with pytest.raises(AssertionError):
        det.extractKeypoints(fr)          # "Frame already has features"


# ---------------------------------------------------------------------------- P3P (tests/test_p3p.py)
def test_estimate_pose_like_reference_test(ctx):
    from vo.pose_estimation import P3PPoseEstimator
    from vo.primitives import Features
    cam1, cam2 = _cameras()
    rng = np.random.default_rng(2023)
    L = rng.uniform(-1, 1, size=(1000, 3, 1))
    L[:, 2] = L[:, 2] * 5 + 10
    p2 = cam2.project_points_world_frame(L)
    est = P3PPoseEstimator(intrinsic_matrix=cam2.intrinsic_matrix, use_opencv=False, inlier_threshold=1,
                           outlier_ratio=0.9, confidence=0.99, max_iterations=1000)
    (R, t), inl = est.estimate_pose(Features(keypoints=p2, landmarks=L))
    assert R.shape == (3, 3) and t.shape == (3, 1)
    assert np.allclose(R, cam2.R, atol=1e-3) and np.allclose(t, cam2.t, atol=1e-3)
    assert inl.all()


@pytest.mark.parametrize("tag", ["clean", "noisy"])
def test_estimate_pose_vs_reference_run(ctx, golden, tag):
    """Same rng stream, iteration count, inlier mask and pose as the reference's estimator, on the first
    and on a second call of the same object (state carried over)."""
    from vo.pose_estimation import P3PPoseEstimator
    from vo.primitives import Features
    g = golden("p3p")
    L = g[f"{tag}_landmarks"].reshape(-1, 3, 1)
    P = g[f"{tag}_keypoints"].reshape(-1, 2, 1)
    for refine in (0, 1):
        est = P3PPoseEstimator(intrinsic_matrix=g["K"], use_opencv=False, inlier_threshold=float(g[f"{tag}_threshold"]),
                               outlier_ratio=0.9, confidence=0.99, max_iterations=1000, nonlinear_refinement=bool(refine))
        est.FIRST_BATCH = 16              # force the multi-batch / rng-rewind path
        for call in ("", "_second"):
            (R, t), inl = est.estimate_pose(Features(keypoints=P.copy(), landmarks=L.copy()))
            k = f"{tag}_refine{refine}{call}"
            assert np.array_equal(inl, g[k + "_inliers"])
            assert _rot_angle(R, g[k + "_R"]) < 1e-5
            assert np.linalg.norm(t - g[k + "_t"]) <= 1e-4 * np.linalg.norm(g[k + "_t"])
            if call == "":
                assert est.ransac.n_iterations == int(g[k + "_n_iterations"])
                assert est.ransac.outlier_ratio == float(g[k + "_outlier_ratio"])


# ---------------------------------------------------------------------------- triangulation (tests/test_triangulation.py)
def test_relative_pose_and_triangulation_like_reference_test(ctx):
    from vo.landmarks import LandmarksTriangulator
    from vo.primitives import Features, Frame, Matches
    cam1, cam2 = _cameras()
    tri = LandmarksTriangulator(camera1=cam1, camera2=cam2, use_ransac=False, use_opencv=False)
    rng = np.random.default_rng(2023)
    for _ in range(3):
        L = rng.uniform(-1, 1, size=(1000, 3, 1))
        L[:, 2] = L[:, 2] * 5 + 10
        p1, p2 = cam1.project_points_world_frame(L), cam2.project_points_world_frame(L)
        ok = (np.all((0 <= p1) & (p1 <= 400), axis=-2) & np.all((0 <= p1) & (p2 <= 400), axis=-2)).flatten()
        L, p1, p2 = L[ok], p1[ok], p2[ok]
        M2, _ = tri._find_relative_pose(p1, p2)
        assert np.allclose(M2[:3, :3], cam2.R)
        assert np.allclose(M2[:3, 3:] / np.linalg.norm(M2[:3, 3:]), cam2.t / np.linalg.norm(cam2.t))
        M2[:3, 3:] *= np.linalg.norm(cam2.t) / np.linalg.norm(M2[:3, 3:])
        c2_T_w = np.vstack([M2, [0, 0, 0, 1]]) @ cam1.c_T_w
        X = tri._linear_triangulation(p1, p2, C1=cam1.intrinsic_matrix @ cam1.c_T_w[:3], C2=cam2.intrinsic_matrix @ c2_T_w[:3])
        assert np.allclose(L, X, atol=1e-4)
        m = Matches(Frame(None, features=Features(keypoints=p1)), Frame(None, features=Features(keypoints=p2)),
                    matches=np.stack([np.arange(len(p1))] * 2, axis=-1))
        m.frame2.features.candidate_mask = np.ones(p1.shape[0], dtype=bool)
        assert np.all(m.frame2.features.tracks == p1) and np.all(m.frame2.features.poses == np.eye(4))
        for use_cv in (False, True):
            tri2 = LandmarksTriangulator(camera1=cam1, camera2=cam2, use_ransac=False, use_opencv=use_cv)
            Xc = tri2.triangulate_candidates(m.frame2.features, np.linalg.inv(cam2.c_T_w))
            assert np.allclose(L, Xc, atol=1e-4)


def test_triangulate_candidates_vs_reference_run(ctx, golden):
    from vo.landmarks import LandmarksTriangulator
    from vo.primitives import Features
    g = golden("triangulation")
    cam1, cam2 = _cameras()
    for tag in ("clean", "noisy"):
        f = Features(keypoints=g[f"{tag}_p2"].reshape(-1, 2, 1).copy())
        f.tracks = g[f"{tag}_cand_tracks"].reshape(-1, 2, 1).copy()
        f.poses = g[f"{tag}_cand_poses"].copy()
        f.candidate_mask = np.ones(f.length, dtype=bool)
        for use_cv in (False, True):
            tri = LandmarksTriangulator(cam1, cam2, use_ransac=False, use_opencv=use_cv)
            X = tri.triangulate_candidates(f, g[f"{tag}_cand_current_pose"]).reshape(-1, 3)
            ref = g[f"{tag}_cand_cv{int(use_cv)}"]
            assert np.abs(X - ref).max() <= 1e-7 * max(1.0, np.abs(ref).max())


# ---------------------------------------------------------------------------- KLT tracker end to end
def test_klt_tracker_vs_reference_run(ctx, golden):
    from vo.features import KLTTracker
    from vo.primitives import Frame
    g = golden("klt")
    f0, f1 = Frame(np.stack([g["prev"]] * 3, -1)), Frame(np.stack([g["next"]] * 3, -1))
    trk = KLTTracker(f0)
    assert np.array_equal(f0.features.keypoints.reshape(-1, 2).astype(np.float32), g["tracker_init_kp"])
    m = trk.track_features(f0, f1)
    kp1 = m.frame1.features.keypoints.reshape(-1, 2)
    kp2 = m.frame2.features.keypoints.reshape(-1, 2)
    assert kp1.shape == g["tracker_kp1"].shape          # same points survive the status / error filter
    assert np.array_equal(kp1.astype(np.float32), g["tracker_kp1"])
    assert np.abs(kp2 - g["tracker_kp2"]).max() < 1e-2  # 1e-2 px
    assert (m.frame2.features.state == 1).all()


def test_two_contexts_in_one_process():
    """Per-device kernel attributes (dynamic shared memory opt-ins) are tracked per context: a second context -- on
    another GPU when the box has one, else on the same -- must be able to run every large-shared-memory kernel."""
    import torch
    from conftest import synthetic_image
    from vo import _native as nat, _ops
    dev = 1 if torch.cuda.device_count() > 1 else 0
    c2 = nat.Context(dev)
    img = synthetic_image(128, 192, seed=3)
    a = _ops.harris_detect(img, 60, desc_radius=9, ctx=c2)
    b = _ops.harris_detect(img, 60, desc_radius=9, ctx=nat.default_context(0))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    pairs = _ops.match_descriptors(a[2], b[2], ctx=c2)
    assert len(pairs) > 30
    c2.close()
