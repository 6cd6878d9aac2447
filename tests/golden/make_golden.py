"""Generate the golden fixtures under tests/golden/ by importing the REFERENCE itself.

Run in the build container only (needs /root/reference and cv2):
    python tests/golden/make_golden.py
The GPU box has no /root/reference, so the outputs (*.npz, small) are committed.
matplotlib / pytransform3d are not installed here and are only used by the reference's plotting
code, so they are stubbed before import.
"""
import os
import sys
import types

import numpy as np

for _name in ["matplotlib", "matplotlib.pyplot", "matplotlib.lines", "mpl_toolkits", "pytransform3d",
              "pytransform3d.transformations", "pytransform3d.plot_utils", "pytransform3d.camera",
              "pytransform3d.rotations"]:
    sys.modules.setdefault(_name, types.ModuleType(_name))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "src"))
OUT = os.path.dirname(os.path.abspath(__file__))

import cv2  # noqa: E402


def kitti_gray(i):
    img = cv2.imread(f"{REF}/tests/test_data/kitti/05/image_0/{i:06d}.png")
    return cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)


def make_harris():
    from vo.features.harris import HarrisCornerDetector
    from vo.primitives import Frame

    out = {}
    g = kitti_gray(0)
    # (a) full KITTI frame, test_harris.py's configuration (num_keypoints=200): keypoints only
    det = HarrisCornerDetector(num_keypoints=200)
    fr = det.extractKeypoints(Frame(g.copy()))
    out["full_shape"] = np.array(g.shape)
    out["full_kp200"] = fr.features.keypoints.reshape(-1, 2).astype(np.int32)
    # (b) a 192x320 crop that travels with the repo: keypoints, descriptors and the score map
    crop = np.ascontiguousarray(g[100:292, 400:720])
    out["crop"] = crop
    for K, r in [(150, 5), (400, 3)]:
        det = HarrisCornerDetector(num_keypoints=K, nonmaximum_supression_radius=r)
        fr = det.extractKeypoints(Frame(crop.copy()))
        fr = det.extractDescriptors(fr)
        out[f"crop_kp_K{K}_r{r}"] = fr.features.keypoints.reshape(-1, 2).astype(np.int32)
        out[f"crop_desc_K{K}_r{r}"] = fr.features.descriptors.reshape(K, -1).astype(np.uint8)
    # score map of the crop, recomputed with the reference's exact expressions (harris.py:102-137)
    from scipy import signal
    sx = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]])
    sy = np.array([[-1, -2, -1], [0, 0, 0], [1, 2, 1]])
    Ix = signal.convolve2d(sx, crop, mode="valid", boundary="symm")
    Iy = signal.convolve2d(sy, crop, mode="valid", boundary="symm")
    p = np.ones((9, 9))
    a = signal.convolve2d(p, Ix ** 2, mode="valid", boundary="symm")
    b = signal.convolve2d(p, Iy ** 2, mode="valid", boundary="symm")
    c = signal.convolve2d(p, Ix * Iy, mode="valid", boundary="symm")
    s = a * b - c ** 2 - 0.09 * ((a + b) ** 2)
    s[s < 0] = 0
    out["crop_resp"] = np.pad(s, [(5, 5), (5, 5)])
    # few-corner case: K larger than the number of selectable corners -> (0, 0) fill
    blank = np.zeros((64, 96), np.uint8)
    blank[20:40, 30:60] = 200
    det = HarrisCornerDetector(num_keypoints=40)
    fr = det.extractKeypoints(Frame(blank.copy()))
    out["blank"] = blank
    out["blank_kp40"] = fr.features.keypoints.reshape(-1, 2).astype(np.int32)
    # matchDescriptor of the reference on two consecutive KITTI crops (harris.py:196-264)
    crop2 = np.ascontiguousarray(kitti_gray(1)[100:292, 400:720])
    det = HarrisCornerDetector(num_keypoints=300)
    fa = det.extractDescriptors(det.extractKeypoints(Frame(crop.copy())))
    fb = det.extractDescriptors(det.extractKeypoints(Frame(crop2.copy())))
    out["crop2"] = crop2
    out["match_desc1"] = fa.features.descriptors.reshape(300, -1).astype(np.uint8)
    out["match_desc2"] = fb.features.descriptors.reshape(300, -1).astype(np.uint8)
    bf = cv2.BFMatcher()
    used = np.zeros(300)
    good = []
    for m, n in bf.knnMatch(out["match_desc1"].astype(np.float32), out["match_desc2"].astype(np.float32), k=2):
        if m.distance < 0.85 * n.distance and used[m.trainIdx] == 0:
            good.append([m.queryIdx, m.trainIdx])
            used[m.trainIdx] = 1
    out["match_pairs"] = np.array(good, np.int32)
    mt = det.matchDescriptor(fa, fb)
    out["match_n_matched"] = int((mt.frame1.features.state == 1).sum())
    np.savez_compressed(os.path.join(OUT, "harris.npz"), **out)
    print("harris.npz", {k: np.shape(v) for k, v in out.items()})


def make_klt():
    """klt.py:29-33 parameters; cv2.calcOpticalFlowPyrLK is the reference's own call (klt.py:233-239)."""
    from vo.features.klt import KLTTracker
    from vo.primitives import Frame

    out = {}
    g0, g1 = kitti_gray(0), kitti_gray(1)
    c0 = np.ascontiguousarray(g0[100:292, 400:720])
    c1 = np.ascontiguousarray(g1[100:292, 400:720])
    out["prev"], out["next"] = c0, c1
    lk = dict(winSize=(17, 17), maxLevel=2, criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 10, 0.03))
    pts = cv2.goodFeaturesToTrack(c0, maxCorners=500, qualityLevel=0.01, minDistance=8, blockSize=7).reshape(-1, 2)
    rng = np.random.default_rng(7)
    H, W = c0.shape
    edge = np.stack([rng.uniform(-12, W + 12, 120), rng.uniform(-12, H + 12, 120)], 1).astype(np.float32)
    pts = np.concatenate([pts, edge]).astype(np.float32)
    nxt, st, err = cv2.calcOpticalFlowPyrLK(c0, c1, pts.reshape(-1, 1, 2), None, **lk)
    out["pts"], out["next_pts"], out["status"], out["err"] = pts, nxt.reshape(-1, 2), st.ravel(), err.ravel()
    # a 4-level pyramid / other window, as BASELINE config 4 uses
    lk2 = dict(winSize=(21, 21), maxLevel=3, criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 30, 0.01))
    nxt, st, err = cv2.calcOpticalFlowPyrLK(c0, c1, pts.reshape(-1, 1, 2), None, **lk2)
    out["next_pts_w21_l3"], out["status_w21_l3"], out["err_w21_l3"] = nxt.reshape(-1, 2), st.ravel(), err.ravel()
    out["pyr1"] = cv2.pyrDown(c0)
    out["pyr2"] = cv2.pyrDown(out["pyr1"])
    # the reference class end to end (BGR frames; equal channels convert back to the same gray)
    f0 = Frame(np.stack([c0] * 3, -1))
    f1 = Frame(np.stack([c1] * 3, -1))
    np.random.seed(0)
    trk = KLTTracker(f0)
    out["tracker_init_kp"] = f0.features.keypoints.reshape(-1, 2).astype(np.float32)
    m = trk.track_features(f0, f1)
    out["tracker_kp1"] = m.frame1.features.keypoints.reshape(-1, 2).astype(np.float32)
    out["tracker_kp2"] = m.frame2.features.keypoints.reshape(-1, 2).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "klt.npz"), **out)
    print("klt.npz", {k: v.shape for k, v in out.items()}, "tracked", int(out["status"].sum()))


def _cameras():
    from vo.sensors import Camera
    K = np.array([[500, 0, 320], [0, 500, 240], [0, 0, 1]], dtype=float)
    th1, th2 = np.pi / 8, np.pi / 32
    R = np.array([[np.cos(th1), -np.sin(th1), 0], [np.sin(th1), np.cos(th1), 0], [0, 0, 1]])
    R = R @ np.array([[np.cos(th2), 0, np.sin(th2)], [0, 1, 0], [-np.sin(th2), 0, np.cos(th2)]])
    t = np.array([[1, 1, -1]], dtype=float).T
    return Camera(K, R=np.eye(3), t=np.zeros((3, 1))), Camera(K, R=R, t=t)


def make_p3p():
    """tests/test_p3p.py's configuration through the reference's own P3PPoseEstimator (use_opencv=False)."""
    from vo.pose_estimation import P3PPoseEstimator
    from vo.primitives import Features

    out = {}
    cam1, cam2 = _cameras()
    out["K"], out["R_true"], out["t_true"] = cam2.intrinsic_matrix, cam2.R, cam2.t
    rng = np.random.default_rng(2023)
    for tag, N, noise, outl, thr in [("clean", 1000, 0.0, 0.0, 1.0), ("noisy", 600, 0.4, 0.35, 2.0)]:
        L = rng.uniform(-1, 1, size=(N, 3, 1))
        L[:, 2] = L[:, 2] * 5 + 10
        p2 = cam2.project_points_world_frame(L)
        p2 = p2 + rng.normal(0, noise, p2.shape) if noise > 0 else p2
        n_out = int(outl * N)
        if n_out:
            idx = rng.choice(N, n_out, replace=False)
            p2[idx] += rng.uniform(-60, 60, (n_out, 2, 1))
        out[f"{tag}_landmarks"], out[f"{tag}_keypoints"], out[f"{tag}_threshold"] = L.reshape(N, 3), p2.reshape(N, 2), thr
        for refine in (False, True):
            est = P3PPoseEstimator(intrinsic_matrix=cam2.intrinsic_matrix, use_opencv=False, inlier_threshold=thr,
                                   outlier_ratio=0.9, confidence=0.99, max_iterations=1000,
                                   nonlinear_refinement=refine)
            (R, t), inl = est.estimate_pose(Features(keypoints=p2.copy(), landmarks=L.copy()))
            k = f"{tag}_refine{int(refine)}"
            out[k + "_R"], out[k + "_t"], out[k + "_inliers"] = R, np.asarray(t).reshape(3, 1), inl
            out[k + "_n_iterations"] = est.ransac.n_iterations
            out[k + "_outlier_ratio"] = est.ransac.outlier_ratio
            # second call on the same estimator: rng / n_iterations / outlier_ratio state carries over
            (R, t), inl = est.estimate_pose(Features(keypoints=p2.copy(), landmarks=L.copy()))
            out[k + "_second_R"], out[k + "_second_t"], out[k + "_second_inliers"] = R, np.asarray(t).reshape(3, 1), inl
        # the solver alone: cv2.solvePnP on the first 96 sample sets of the reference's rng stream
        r = np.random.default_rng(2023)
        S, valid, models = [], [], []
        for _ in range(96):
            ids = r.choice(np.arange(N), replace=False, size=4)
            ok, rv, tv = cv2.solvePnP(L[ids], p2[ids], cam2.intrinsic_matrix, None, flags=cv2.SOLVEPNP_P3P)
            S.append(ids)
            valid.append(ok)
            models.append(np.concatenate([cv2.Rodrigues(rv)[0].ravel(), tv.ravel()]) if ok else np.zeros(12))
        out[f"{tag}_sample_idx"], out[f"{tag}_cv_valid"], out[f"{tag}_cv_models"] = np.array(S, np.int32), np.array(valid), np.array(models)
    np.savez_compressed(os.path.join(OUT, "p3p.npz"), **out)
    print("p3p.npz", {k: np.shape(v) for k, v in out.items()})


def make_triangulation():
    """tests/test_triangulation.py's two-camera configuration through LandmarksTriangulator."""
    from vo.landmarks.triangulation import LandmarksTriangulator
    from vo.primitives import Features

    out = {}
    cam1, cam2 = _cameras()
    rng = np.random.default_rng(2023)
    N = 400
    L = rng.uniform(-1, 1, size=(N, 3, 1))
    L[:, 2] = L[:, 2] * 5 + 10
    out["landmarks"] = L.reshape(N, 3)
    out["C1"], out["C2"] = cam1.projection_matrix, cam2.projection_matrix
    for tag, noise in [("clean", 0.0), ("noisy", 0.5)]:
        p1 = cam1.project_points_world_frame(L) + rng.normal(0, noise, (N, 2, 1))
        p2 = cam2.project_points_world_frame(L) + rng.normal(0, noise, (N, 2, 1))
        out[f"{tag}_p1"], out[f"{tag}_p2"] = p1.reshape(N, 2), p2.reshape(N, 2)
        tri = LandmarksTriangulator(cam1, cam2, use_ransac=False, use_opencv=False)
        out[f"{tag}_linear"] = tri._linear_triangulation(p1, p2, cam1.projection_matrix, cam2.projection_matrix).reshape(N, 3)
        # triangulate_candidates: per-point start poses (triangulation.py:50-57), both code paths
        poses = np.stack([np.eye(4)] * N)
        poses[:, :3, 3] = rng.normal(0, 0.05, (N, 3))        # slightly different start pose per track
        K = cam1.intrinsic_matrix
        starts = np.stack([to_pixels(K, np.linalg.inv(poses[i])[:3], L[i]) for i in range(N)]) + rng.normal(0, noise, (N, 2, 1))
        feats = Features(keypoints=p2.copy())
        feats.tracks = starts
        feats.poses = poses
        feats.candidate_mask = np.ones(N, dtype=bool)
        cur = np.linalg.inv(cam2.c_T_w)
        out[f"{tag}_cand_tracks"], out[f"{tag}_cand_poses"], out[f"{tag}_cand_current_pose"] = starts.reshape(N, 2), poses, cur
        for use_cv in (False, True):
            tri = LandmarksTriangulator(cam1, cam2, use_ransac=False, use_opencv=use_cv)
            out[f"{tag}_cand_cv{int(use_cv)}"] = tri.triangulate_candidates(feats, cur).reshape(N, 3)
    np.savez_compressed(os.path.join(OUT, "triangulation.npz"), **out)
    print("triangulation.npz", {k: np.shape(v) for k, v in out.items()})


def make_bookkeeping():
    """Matches / State bookkeeping of the reference on a mixed-state feature table."""
    from vo.primitives import Features, Frame, Matches, State
    from vo.sensors import Camera

    rng = np.random.default_rng(11)
    n1, n2 = 40, 36
    kp1 = rng.uniform(0, 300, (n1, 2, 1))
    kp2 = rng.uniform(0, 300, (n2, 2, 1))
    f1 = Features(kp1.copy())
    st = rng.integers(0, 3, n1).astype(float)
    f1.state = st.copy()
    lm = np.full((n1, 3, 1), np.nan)
    lm[st == 2] = rng.uniform(-3, 3, (int((st == 2).sum()), 3, 1)) + np.array([[0], [0], [12.0]])
    f1.landmarks = lm.copy()
    tr = kp1.copy()
    tr[st == 1] += rng.normal(0, 3, (int((st == 1).sum()), 2, 1))
    f1.tracks = tr.copy()
    poses = np.stack([np.eye(4)] * n1)
    poses[:, :3, 3] = rng.normal(0, 0.1, (n1, 3))
    f1.poses = poses.copy()
    f1.descriptors = rng.integers(0, 255, (n1, 5)).astype(float)
    f2 = Features(kp2.copy())
    f2.descriptors = rng.integers(0, 255, (n2, 5)).astype(float)
    m = np.stack([rng.permutation(n1)[:25], rng.permutation(n2)[:25]], 1)
    out = dict(kp1=kp1, kp2=kp2, state1=st, land1=lm, tracks1=tr, poses1=poses, desc1=f1.descriptors.copy(),
               desc2=f2.descriptors.copy(), matches=m)
    K = np.array([[500, 0, 150], [0, 500, 150], [0, 0, 1.0]])
    fr1, fr2 = Frame(None, features=f1, sensor=Camera(K)), Frame(None, features=f2, sensor=Camera(K))
    mt = Matches(fr1, fr2, m)
    for tag, f in (("a", mt.frame1.features), ("b", mt.frame2.features)):
        out.update({f"{tag}_kp": f.keypoints.copy(), f"{tag}_state": f.state.copy(), f"{tag}_land": f.landmarks.copy(),
                    f"{tag}_tracks": f.tracks.copy(), f"{tag}_poses": f.poses.copy(), f"{tag}_desc": f.descriptors.copy()})
    state = State(fr1)
    state.update_from_matches(mt)
    pose_w2c = np.eye(4)
    pose_w2c[:3, 3] = [0.3, -0.1, 0.2]
    state.update_with_world_pose(pose_w2c[:3])
    outl = np.zeros(mt.frame2.features.length, dtype=bool)
    outl[::7] = True
    state.reset_outliers(outl)
    state.compute_candidates()
    f = state.curr_frame.features
    out.update(dict(s_pose=state.get_pose(), s_state=f.state, s_tracks=f.tracks, s_poses=f.poses, s_cand=f.candidate_mask,
                    s_outliers=outl, s_pose_in=pose_w2c[:3]))
    np.savez_compressed(os.path.join(OUT, "bookkeeping.npz"), **out)
    print("bookkeeping.npz", {k: np.shape(v) for k, v in out.items()})


def _ref_loop(p3p_opencv, mode="klt", refine=True, n_frames=6):
    """The reference's src/main.py:185-287 without the plotting: bootstrap on frames 0 and 2, then the loop body for
    the remaining frames.  Returns the bootstrap table and one record per loop frame."""
    from vo.features import Tracker
    from vo.landmarks import LandmarksTriangulator
    from vo.pose_estimation import P3PPoseEstimator
    from vo.primitives import Features, Frame, State
    from vo.sensors import Camera

    vals = open(f"{REF}/tests/test_data/kitti/05/calib.txt").readlines()[1].split(" ")[1:]      # loader.py:86-96
    K = np.array([np.float32(v) for v in vals]).reshape(3, 4)[:, :3]
    camera = Camera(intrinsic_matrix=K)
    frames = []
    for i in range(n_frames):
        f = Frame(cv2.imread(f"{REF}/tests/test_data/kitti/05/image_0/{i:06d}.png"), sensor=camera, intrinsics=K)
        f.frame_id = i
        frames.append(f)
    np.random.seed(0)                                # uids only (klt.py:70-72)
    triangulator = LandmarksTriangulator(camera1=camera, camera2=camera, use_ransac=True, use_opencv=True,
                                         outlier_ratio=0.9, ransac_threshold=0.25, ransac_confidence=0.999)   # main.py:185-193
    pose_estimator = P3PPoseEstimator(use_opencv=p3p_opencv, intrinsic_matrix=K, inlier_threshold=1.25,
                                      outlier_ratio=0.9, confidence=0.9999, nonlinear_refinement=refine)      # main.py:194-201
    it = iter(frames)
    init_frame = next(it)
    state = State(init_frame)
    next(it)
    new_frame = next(it)
    tracker = Tracker(init_frame, mode=mode)
    init_kp = init_frame.features.keypoints.reshape(-1, 2).copy()
    matches = tracker.trackFeatures(state.curr_frame, new_frame)
    state.update_from_matches(matches)
    M, landmarks, inliers = triangulator.triangulate_matches(matches)
    outliers = np.zeros(shape=(matches.frame2.features.length,), dtype=bool)
    outliers[matches.frame2.features.match_inliers] = ~inliers
    state.update_with_local_pose(M)
    inliers_mask = np.zeros_like(matches.frame2.features.matched_candidate_inliers).astype(bool)
    inliers_mask[matches.frame2.features.matched_candidate_inliers] = inliers
    state.update_with_local_landmarks(landmarks[inliers], inliers_mask)
    state.reset_outliers(outliers)

    def table(prefix):
        f = state.curr_frame.features
        return {prefix + "kp": f.keypoints.reshape(-1, 2).copy(), prefix + "land": f.landmarks.reshape(-1, 3).copy(),
                prefix + "state": f.state.astype(np.int8), prefix + "track": f.tracks.reshape(-1, 2).copy(),
                prefix + "pose": f.poses.copy(), prefix + "cand": f.candidate_mask.copy(),
                prefix + "curr_pose": state.get_pose().copy()}
    out = {"K": K, "init_kp": init_kp, "num_features": tracker._tracker._num_features}
    out.update(table("boot_"))
    for new_frame in it:
        matches = tracker.trackFeatures(state.curr_frame, new_frame)
        (rmatrix, tvec), inl = pose_estimator.estimate_pose(Features(
            keypoints=matches.frame2.features.triangulated_inliers_keypoints,
            landmarks=matches.frame2.features.triangulated_inliers_landmarks))
        outliers = np.zeros(shape=(matches.frame2.features.length,), dtype=bool)
        outliers[matches.frame2.features.triangulate_inliers] = ~inl
        state.update_from_matches(matches)
        state.update_with_world_pose(np.concatenate((rmatrix, tvec), axis=1))
        state.reset_outliers(outliers)
        state.compute_candidates()
        n_candidates = np.sum(state.curr_frame.features.candidate_mask)
        if n_candidates > 0:
            lw = triangulator.triangulate_candidates(state.curr_frame.features, current_pose=state.get_pose())
            state.update_with_world_landmarks(lw, matches.frame2.features.candidate_mask)
        pre = f"f{new_frame.frame_id}_"
        out.update(table(pre))
        out[pre + "inliers"] = np.asarray(inl).copy()
        out[pre + "R"], out[pre + "t"] = np.asarray(rmatrix).copy(), np.asarray(tvec).reshape(3).copy()
        out[pre + "n_candidates"] = int(n_candidates)
        if not p3p_opencv:
            out[pre + "n_iterations"] = pose_estimator.ransac.n_iterations
    return out


def make_loop():
    """The end-to-end pin (BASELINE configs[0]): the reference's main loop, KLT tracker mode, on the six KITTI frames it
    ships.  `loop.npz` holds the run with use_opencv=False for P3P (reproducible: numpy rng 2023), `cv_*` keys the final
    poses / inlier masks of the use_opencv=True run main.py itself uses (cv2.solvePnPRansac, tolerance pin only).
    The frames are copied next to the fixtures so the GPU box (no /root/reference) can replay the loop."""
    import shutil
    dst = os.path.join(OUT, "kitti05")
    os.makedirs(dst, exist_ok=True)
    for i in range(6):
        shutil.copyfile(f"{REF}/tests/test_data/kitti/05/image_0/{i:06d}.png", os.path.join(dst, f"{i:06d}.png"))
    out = _ref_loop(p3p_opencv=False)
    cvrun = _ref_loop(p3p_opencv=True)
    for k, v in cvrun.items():
        if k.startswith("f") and k.split("_", 1)[1] in ("inliers", "R", "t", "curr_pose", "n_candidates", "state"):
            out["cv_" + k] = v
    norefine = _ref_loop(p3p_opencv=False, refine=False)
    for k, v in norefine.items():
        if k.startswith("f") and k.split("_", 1)[1] in ("R", "t", "curr_pose", "inliers"):
            out["noref_" + k] = v
    # full-frame Harris goldens (BASELINE configs[0]/[1] shape): K = 200 (test_harris.py) and K = 1000 (main.py default)
    from vo.features.harris import HarrisCornerDetector
    from vo.primitives import Frame
    g = kitti_gray(0)
    for Kk in (200, 1000):
        det = HarrisCornerDetector(num_keypoints=Kk)
        out[f"harris_full_kp{Kk}"] = det.extractKeypoints(Frame(g.copy())).features.keypoints.reshape(-1, 2).astype(np.int32)
    # Shi-Tomasi corners of every frame (klt.py:24-26, 99-100): cv2's own output, the detector of the KLT mode
    for i in range(6):
        out[f"gftt_{i}"] = cv2.goodFeaturesToTrack(kitti_gray(i), maxCorners=500, qualityLevel=0.01, minDistance=8,
                                                   blockSize=7).reshape(-1, 2)
    np.savez_compressed(os.path.join(OUT, "loop.npz"), **out)
    print("loop.npz", {k: np.shape(v) for k, v in out.items() if not k.startswith(("cv_", "noref_"))})


def make_bootstrap():
    """The two-view bootstrap (triangulation.py:88-350) as main.py:185-193 configures it, run by the REFERENCE's
    LandmarksTriangulator: (a) on the matches of its own KLT tracker between KITTI frames 0 and 2 (main.py:203-222),
    (b) on synthetic two-view problems with outliers.  Inputs and every result travel in bootstrap.npz."""
    from vo.features import Tracker
    from vo.landmarks import LandmarksTriangulator
    from vo.primitives import Frame, State
    from vo.sensors import Camera

    vals = open(f"{REF}/tests/test_data/kitti/05/calib.txt").readlines()[1].split(" ")[1:]
    K = np.array([np.float32(v) for v in vals]).reshape(3, 4)[:, :3]
    camera = Camera(intrinsic_matrix=K)
    out = {"K": K}

    def run(tag, p1, p2, thr, conf):
        tri = LandmarksTriangulator(camera1=camera, camera2=camera, use_ransac=True, use_opencv=True, outlier_ratio=0.9,
                                    ransac_threshold=thr, ransac_confidence=conf)
        F, f_inl = tri._find_fundamental_matrix_ransac(p1, p2)
        M, land, inl = tri._find_relative_pose(p1, p2)
        out.update({tag + "p1": p1.reshape(-1, 2), tag + "p2": p2.reshape(-1, 2), tag + "thr": thr, tag + "conf": conf,
                    tag + "F": F, tag + "f_inl": f_inl, tag + "M": M, tag + "land": land.reshape(-1, 3), tag + "inl": inl})
        print(tag, p1.shape[0], "points", int(f_inl.sum()), "F inliers", int(inl.sum()), "valid")

    frames = []
    for i in range(3):
        f = Frame(cv2.imread(f"{REF}/tests/test_data/kitti/05/image_0/{i:06d}.png"), sensor=camera, intrinsics=K)
        f.frame_id = i
        frames.append(f)
    np.random.seed(0)
    state = State(frames[0])
    tracker = Tracker(frames[0], mode="klt")
    matches = tracker.trackFeatures(state.curr_frame, frames[2])
    run("kitti_", matches.frame1.features.matched_candidate_inliers_keypoints.astype(np.float64),
        matches.frame2.features.matched_candidate_inliers_keypoints.astype(np.float64), 0.25, 0.999)
    rng = np.random.default_rng(7)
    for j, (N, frac, thr, conf) in enumerate([(300, 0.2, 0.25, 0.999), (1000, 0.4, 1.0, 0.99), (64, 0.1, 3.0, 0.999)]):
        X = np.c_[rng.uniform(-10, 10, N), rng.uniform(-3, 3, N), rng.uniform(6, 40, N)]
        R, _ = cv2.Rodrigues(rng.uniform(-0.05, 0.05, 3))
        t = np.array([0.1, -0.05, -1.0]) + rng.normal(0, 0.05, 3)
        a = (K.astype(np.float64) @ X.T).T
        b = (K.astype(np.float64) @ (X @ R.T + t).T).T
        p1 = a[:, :2] / a[:, 2:] + rng.normal(0, 0.1, (N, 2))
        p2 = b[:, :2] / b[:, 2:] + rng.normal(0, 0.1, (N, 2))
        no = int(frac * N)
        p2[:no] += rng.uniform(-40, 40, (no, 2))
        run(f"syn{j}_", p1.reshape(-1, 2, 1), p2.reshape(-1, 2, 1), thr, conf)
    np.savez_compressed(os.path.join(OUT, "bootstrap.npz"), **out)


def to_pixels(K, M, X):
    x = K @ (M[:, :3] @ X + M[:, 3:])
    return x[:2] / x[2:]


if __name__ == "__main__":
    which = sys.argv[1:] or ["harris", "klt", "p3p", "triangulation", "bookkeeping", "loop", "bootstrap"]
    for w in which:
        globals()["make_" + w]()
