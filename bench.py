#!/usr/bin/env python
"""VO front-end benchmark: frames/s of the full, CHAINED step (Harris + KLT + P3P-RANSAC + triangulation with the
reference's per-frame data flow, src/main.py:248-287) on synthetic KITTI-shaped 1241x376 grayscale sequences, plus the
Harris response kernel's HBM roofline.

    python bench.py --gpus 1 --steps 300 --warmup 5          # this repo's CUDA path (vo_pipeline_*)
    python bench.py --impl reference --steps 3 --warmup 1    # the reference algorithm on the box's host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --workload p3p_sweep | stress            # the other BASELINE configs (extra lines, same contract)

One "step" advances S independent sequences (S per GPU, `config.sequences_per_gpu`) by one frame: the detector runs on
the new frame, the sequence's feature table (resident in HBM) is tracked into it, the pose comes from P3P-RANSAC on the
rows that carry landmarks, outliers restart their tracks, tracks with enough parallax are triangulated.  `value` is
measured with the frames already in HBM (CUDA events on the launching stream, max over ranks); `e2e` goes through the
host-buffer C ABI: every step uploads its frames from pinned host memory (nothing else) and downloads the poses and
counters.  Sequences are sharded across ranks with no data-path collective (weak scaling).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "visual-odometry-project_b200"))
sys.path.insert(0, ROOT)

H, W = 376, 1241
K_INTR = np.array([[707.0912, 0, 601.8873], [0, 707.0912, 183.1104], [0, 0, 1.0]])
METRIC = "VO frames/s @1241x376 (Harris+KLT+P3P-RANSAC+triang)"
DEPTH = 20.0                  # the static scene is a textured plane at this depth
FLOW = (3, 1)                 # camera translation per frame, in pixels of image motion (x, y)
KP, CAPACITY = 1000, 2048     # Harris keypoints per detection (harris.py:16-34), table rows per sequence
P3P_CONF, P3P_MAX_ITER, P3P_THR = 0.9999, 10000, 1.25          # src/main.py:194-201
N_MOVERS = 22                 # independently moving patches per sequence: their features are the RANSAC outliers


# ------------------------------------------------------------------------------------------------
# synthetic world (seeded; the same generator feeds the CUDA arm, the e2e arm and the CPU arm)
# ------------------------------------------------------------------------------------------------
def make_texture(seed, th=1024, tw=3072):
    import cv2
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (th, tw)).astype(np.float32)
    img = cv2.GaussianBlur(img, (0, 0), 2.0)
    img = (img - img.min()) / (img.max() - img.min()) * 255
    for _ in range(2500):
        y, x = int(rng.integers(0, th - 40)), int(rng.integers(0, tw - 40))
        hh, ww = int(rng.integers(6, 36)), int(rng.integers(6, 36))
        img[y:y + hh, x:x + ww] = rng.integers(0, 256)
    img = cv2.GaussianBlur(img, (0, 0), 0.8)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def frame_from_texture(tex, s, t):
    """Frame t of sequence s without movers: a crop that moves FLOW px per frame (used by the stage-level tools)."""
    th, tw = tex.shape
    ox = (37 * s + FLOW[0] * t) % (tw - W)
    oy = (11 * s + FLOW[1] * t) % (th - H)
    return tex[oy:oy + H, ox:ox + W]


class World:
    """S sequences of a camera translating in front of a textured plane (depth DEPTH, FLOW pixels of image motion per
    frame) with N_MOVERS independently moving patches each.  Frame `t` of every sequence is a pure function of t, so a
    pool of P time steps can be walked back and forth (0, 1, .., P-1, P-2, .., 0, 1, ..) as an endless, geometrically
    consistent sequence."""

    def __init__(self, seed, S, P):
        self.S, self.P = S, P
        self.tex = make_texture(seed)
        self.tex2 = make_texture(seed + 7919, 512, 1024)
        rng = np.random.default_rng(seed + 1)
        th, tw = self.tex.shape
        self.ox = (37 * np.arange(S)) % (tw - W - FLOW[0] * P)
        self.oy = (11 * np.arange(S)) % (th - H - FLOW[1] * P)
        m = N_MOVERS
        self.mw, self.mh = rng.integers(40, 100, (S, m)), rng.integers(30, 70, (S, m))
        self.mx, self.my = rng.uniform(0, W - 100, (S, m)), rng.uniform(0, H - 70, (S, m))
        v = rng.uniform(2.0, 5.0, (S, m)) * rng.choice([-1, 1], (S, m))
        self.mvx = v - FLOW[0] * (v < 0)                      # at least 2 px/frame away from the background flow (-FLOW)
        self.mvy = rng.integers(-3, 4, (S, m)).astype(float)
        self.msx, self.msy = rng.integers(0, 1024 - 100, (S, m)), rng.integers(0, 512 - 70, (S, m))

    def frame(self, s, t):
        oy, ox = self.oy[s] + FLOW[1] * t, self.ox[s] + FLOW[0] * t
        f = self.tex[oy:oy + H, ox:ox + W].copy()
        for j in range(N_MOVERS):
            x, y = int(round(self.mx[s, j] + self.mvx[s, j] * t)), int(round(self.my[s, j] + self.mvy[s, j] * t))
            w, h = int(self.mw[s, j]), int(self.mh[s, j])
            x0, y0, x1, y1 = max(x, 0), max(y, 0), min(x + w, W), min(y + h, H)
            if x1 > x0 and y1 > y0:
                sx, sy = int(self.msx[s, j]) + (x0 - x), int(self.msy[s, j]) + (y0 - y)
                f[y0:y1, x0:x1] = self.tex2[sy:sy + (y1 - y0), sx:sx + (x1 - x0)]
        return f

    def pool(self):
        out = np.empty((self.P, self.S, H, W), np.uint8)
        for t in range(self.P):
            for s in range(self.S):
                out[t, s] = self.frame(s, t)
        return out

    def time_index(self, i):
        """pool index of step i of the endless back-and-forth walk"""
        if self.P == 1:
            return 0
        k = i % (2 * (self.P - 1))
        return k if k < self.P else 2 * (self.P - 1) - k

    @staticmethod
    def landmarks_of(kp_xy):
        """World points (camera-0 frame) of frame-0 pixels on the static plane."""
        kp = np.asarray(kp_xy, dtype=np.float64).reshape(-1, 2)
        return np.stack([(kp[:, 0] - K_INTR[0, 2]) * DEPTH / K_INTR[0, 0], (kp[:, 1] - K_INTR[1, 2]) * DEPTH / K_INTR[1, 1],
                         np.full(len(kp), DEPTH)], 1)

    @staticmethod
    def true_position(t):
        return np.array([FLOW[0] * t * DEPTH / K_INTR[0, 0], FLOW[1] * t * DEPTH / K_INTR[1, 1], 0.0])


def pipeline_kwargs():
    return dict(capacity=CAPACITY, det_max_corners=KP, p3p_threshold=P3P_THR, p3p_opencv=False, confidence=P3P_CONF,
                max_iterations=P3P_MAX_ITER, refine=os.environ.get("VO_BENCH_NO_REFINE") != "1", tri_opencv=True)


def start_tables(pl, S):
    """Hand-over after the (host-side, one-off) bootstrap: every corner of frame 0 gets the landmark of the static plane
    under it (corners on moving patches get a wrong one: they are the first outliers)."""
    for s in range(S):
        t = pl.read_table(s)
        n = t["n"]
        pl.write_table(s, t["kp"], World.landmarks_of(t["kp"]), np.full(n, 2), t["kp"], np.stack([np.eye(4)] * n),
                       curr_pose=np.eye(4), num_features=KP)


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Polls NVML every few ms from a thread while the timed region runs (nvidia-smi -lms is too coarse for a region of
    a fraction of a second); falls back to one nvidia-smi query."""

    def __init__(self, gpu_index):
        self.idx, self.sm, self.reasons, self.max_mhz, self._stop, self.th = gpu_index, [], set(), None, False, None

    def _poll(self):
        import pynvml as nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
        while not self._stop:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception:
            self.th = None

    def stop(self):
        self._stop = True
        if self.th is not None:
            self.th.join(timeout=1)
        if self.sm:
            return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(self.max_mhz), "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": "NVML polled every 4 ms during the timed region"}
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.idx}", "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            a, b = [float(x) for x in out.strip().split(",")]
            return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "source": "nvidia-smi after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on host cores (loop oracle = restated main.py loop; cv2 for the call the reference
# itself makes into OpenCV, scipy for its refinement).  Only these functions touch oracle/.
# ------------------------------------------------------------------------------------------------
def _cv_klt(prev, nxt, pts, win, max_level, max_iters, epsilon):
    import cv2
    n, s, e = cv2.calcOpticalFlowPyrLK(prev, nxt, np.ascontiguousarray(pts).reshape(-1, 1, 2), None, winSize=(win, win),
                                       maxLevel=max_level, criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, max_iters, epsilon))
    return n.reshape(-1, 2), s.ravel(), e.ravel()


def cpu_make_sequence(world, s):
    import oracle
    from oracle.loop import LoopOracle
    det = lambda im: oracle.harris_keypoints(im, KP, 9, 0.09, 5)[0].astype(np.float32)      # harris.py:86-158
    lo = LoopOracle(K_INTR, detector=det, refine="scipy", klt=_cv_klt, inlier_threshold=P3P_THR, confidence=P3P_CONF,
                    max_iterations=P3P_MAX_ITER, eager_detector=True)
    f0 = world.frame(s, 0)
    lo.init_detect(f0)
    lo.set_table(lo.kp, World.landmarks_of(lo.kp), np.full(len(lo.kp), 2), lo.kp, lo.pose, curr_pose=np.eye(4), num_features=KP)
    lo._cached_det = det(f0)
    return lo


def _cpu_worker(conn, seed, n_seq, P, mine):
    """One host process of the CPU arm: owns the sequences `mine` and advances them on command."""
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        pass
    import oracle
    oracle.lib()
    world = World(seed, n_seq, P)
    seqs = {s: cpu_make_sequence(world, s) for s in mine}
    frames = {(s, t): world.frame(s, t) for s in mine for t in range(P)}
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg is None:
            break
        a, b = msg
        infos = [seqs[s].step(frames[(s, a)], frames[(s, b)]) for s in mine]
        conn.send([(i["n_tracked"], i["p3p_N"], i["n_candidates"]) for i in infos])
    conn.close()


def cpu_pipeline_fps(n_seq, steps, warmup, procs, seed=1234):
    """frames/s of the CPU arm: `n_seq` sequences advance one frame per step on `procs` host processes (the reference
    is single-threaded Python; independent sequences in independent processes is how it would be scaled on a host)."""
    import multiprocessing as mp
    import oracle
    oracle.lib()                                   # build the C port once, before the workers need it
    mpc = mp.get_context("spawn")                  # the parent may hold a CUDA context: never fork it
    P = min(8, warmup + steps + 1)
    world = World(seed, 1, P)
    workers = []
    for w in range(procs):
        mine = list(range(w, n_seq, procs))
        if not mine:
            continue
        pc, cc = mpc.Pipe()
        pr = mpc.Process(target=_cpu_worker, args=(cc, seed, n_seq, P, mine), daemon=True)
        pr.start()
        workers.append((pr, pc))
    for _, pc in workers:
        assert pc.recv() == "ready"
    times, stats = [], []
    for it in range(warmup + steps):
        a, b = world.time_index(it), world.time_index(it + 1)
        t0 = time.perf_counter()
        for _, pc in workers:
            pc.send((a, b))
        res = [r for _, pc in workers for r in pc.recv()]
        if it >= warmup:
            times.append(time.perf_counter() - t0)
            stats.append(np.mean(np.array(res, dtype=np.float64), axis=0))
    for pr, pc in workers:
        pc.send(None)
    for pr, pc in workers:
        pr.join(timeout=10)
    ms = 1e3 * float(np.mean(times))
    return n_seq / (ms / 1e3), ms, np.mean(stats, axis=0).tolist()


# ------------------------------------------------------------------------------------------------
def class_api_loop(world, n_frames=12):
    """main.py:248-287 through the drop-in `vo` classes (one sequence, host arrays between the stages, one C-ABI call per
    stage): Tracker("klt").trackFeatures -> P3PPoseEstimator.estimate_pose -> State updates -> triangulate_candidates.
    Returns frames/s over the timed frames (the first two are warm-up)."""
    from vo.features import Tracker
    from vo.landmarks import LandmarksTriangulator
    from vo.pose_estimation import P3PPoseEstimator
    from vo.primitives import Features, Frame, State
    from vo.sensors import Camera
    cam = Camera(intrinsic_matrix=K_INTR)
    frames = []
    for t in range(n_frames + 1):
        g = world.frame(0, world.time_index(t))
        frames.append(Frame(np.ascontiguousarray(np.stack([g, g, g], -1)), sensor=cam, intrinsics=K_INTR))   # BGR, as cv2.imread gives
    tri = LandmarksTriangulator(camera1=cam, camera2=cam, use_ransac=True, use_opencv=True, outlier_ratio=0.9, ransac_threshold=0.25,
                                ransac_confidence=0.999)
    est = P3PPoseEstimator(use_opencv=True, intrinsic_matrix=K_INTR, inlier_threshold=P3P_THR, outlier_ratio=0.9, confidence=P3P_CONF,
                           nonlinear_refinement=True)
    state = State(frames[0])
    tracker = Tracker(frames[0], mode="klt")
    f = frames[0].features                      # hand-over in place of the two-view bootstrap: the plane's landmarks
    n = f.length
    f.landmarks = World.landmarks_of(f.keypoints.reshape(-1, 2)).reshape(n, 3, 1)
    f.state = np.full(n, 2.0)
    f.state[::5] = 0.0                          # some tracks still untriangulated, as after the reference's bootstrap
    f.landmarks[f.state == 0] = np.nan
    times = []
    for t in range(1, n_frames + 1):
        t0 = time.perf_counter()
        matches = tracker.trackFeatures(state.curr_frame, frames[t])
        (R, tv), inl = est.estimate_pose(Features(keypoints=matches.frame2.features.triangulated_inliers_keypoints,
                                                  landmarks=matches.frame2.features.triangulated_inliers_landmarks))
        outliers = np.zeros(shape=(matches.frame2.features.length,), dtype=bool)
        outliers[matches.frame2.features.triangulate_inliers] = ~inl
        state.update_from_matches(matches)
        state.update_with_world_pose(np.concatenate((R, tv), axis=1))
        state.reset_outliers(outliers)
        state.compute_candidates()
        if np.sum(state.curr_frame.features.candidate_mask) > 0:
            lw = tri.triangulate_candidates(state.curr_frame.features, current_pose=state.get_pose())
            state.update_with_world_landmarks(lw, matches.frame2.features.candidate_mask)
        if t > 2:
            times.append(time.perf_counter() - t0)
    return 1.0 / float(np.mean(times)), float(np.linalg.norm(state.get_pose()[:3, 3] - World.true_position(world.time_index(n_frames))))


def class_api_cpu_reference(world, n_frames=4):
    """The same loop on the host, as the reference runs it (loop oracle: cv2.goodFeaturesToTrack / calcOpticalFlowPyrLK,
    restated cv2.solvePnPRansac, scipy least_squares, numpy bookkeeping), one sequence, one process."""
    import cv2
    import oracle
    from oracle.loop import LoopOracle
    det = lambda im: cv2.goodFeaturesToTrack(im, maxCorners=500, qualityLevel=0.01, minDistance=8, blockSize=7).reshape(-1, 2)
    lo = LoopOracle(K_INTR, detector=det, refine="scipy", klt=_cv_klt, p3p_opencv=True, inlier_threshold=P3P_THR, confidence=P3P_CONF,
                    max_iterations=P3P_MAX_ITER)
    fr = [world.frame(0, world.time_index(t)) for t in range(n_frames + 1)]
    lo.init_detect(fr[0])
    n = len(lo.kp)
    state = np.full(n, 2)
    state[::5] = 0
    land = World.landmarks_of(lo.kp)
    land[state == 0] = np.nan
    lo.set_table(lo.kp, land, state, lo.kp, lo.pose, curr_pose=np.eye(4), num_features=n)
    times = []
    for t in range(1, n_frames + 1):
        t0 = time.perf_counter()
        lo.step(fr[t - 1], fr[t])
        if t > 1:
            times.append(time.perf_counter() - t0)
    return 1.0 / float(np.mean(times))


def bootstrap_problems(S, N, seed=77):
    """S synthetic two-view problems (a 3-D point cloud, not the bench's plane: a plane leaves F undetermined), 0.3 px
    noise, 30 % gross outliers: float64 (S, N, 2) pixel pairs and the true relative poses."""
    rng = np.random.default_rng(seed)
    P1, P2, Ms = np.empty((S, N, 2)), np.empty((S, N, 2)), np.empty((S, 3, 4))
    for s in range(S):
        X = np.c_[rng.uniform(-10, 10, N), rng.uniform(-3, 3, N), rng.uniform(6, 40, N)]
        w = rng.uniform(-0.05, 0.05, 3)
        th = np.linalg.norm(w); k = w / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        t = np.array([0.1, -0.05, -1.0]) + rng.normal(0, 0.05, 3)
        a = (K_INTR @ X.T).T
        b = (K_INTR @ (X @ R.T + t).T).T
        P1[s] = a[:, :2] / a[:, 2:] + rng.normal(0, 0.3, (N, 2))
        P2[s] = b[:, :2] / b[:, 2:] + rng.normal(0, 0.3, (N, 2))
        no = int(0.3 * N)
        P2[s, :no] += rng.uniform(-40, 40, (no, 2))
        Ms[s] = np.c_[R, t / np.linalg.norm(t)]
    return P1, P2, Ms


def bootstrap_leg(ctx, dev, stream, tstream, S, N=1000, with_cpu=False):
    """vo_bootstrap_dev on S problems of N matches (threshold 1.0 px, confidence 0.999), device-resident inputs; beside it
    the reference's own call (cv2.findFundamentalMat + the decomposition / triangulations on the host, one core)."""
    import torch
    from vo import _native as nat
    L = nat.lib()
    P1, P2, Ms = bootstrap_problems(S, N)
    d1, d2 = torch.from_numpy(P1).to(dev), torch.from_numpy(P2).to(dev)
    dF = torch.empty((S, 9), dtype=torch.float64, device=dev); dM = torch.empty((S, 12), dtype=torch.float64, device=dev)
    dL = torch.empty((S, N, 3), dtype=torch.float64, device=dev)
    dm = torch.empty((S, N), dtype=torch.uint8, device=dev); di = torch.empty((S, 4), dtype=torch.int32, device=dev)
    K9 = np.ascontiguousarray(K_INTR.reshape(9))

    def call():
        nat.check(L.vo_bootstrap_dev(ctx.handle, d1.data_ptr(), d2.data_ptr(), S, N, None, nat.ptr(K9), 1.0, 0.999, 1000,
                                     dF.data_ptr(), dM.data_ptr(), dL.data_ptr(), dm.data_ptr(), None, di.data_ptr(), stream), "vo_bootstrap_dev")
    call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    for _ in range(3):
        call()
    e1.record(tstream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    M = dM.cpu().numpy().reshape(S, 3, 4)
    info = di.cpu().numpy()
    rot_err = float(np.median([np.abs(M[s][:, :3] - Ms[s][:, :3]).max() for s in range(S)]))
    dir_err = float(np.median([np.abs(M[s][:, 3] - Ms[s][:, 3]).max() for s in range(S)]))
    out = {"ms_per_call": ms, "problems_per_call": S, "matches_per_problem": N, "problems_per_s": S / (ms / 1e3),
           "ransac_iterations_mean": float(info[:, 1].mean()), "models_found": int(info[:, 0].sum()),
           "median_rotation_error": rot_err, "median_translation_direction_error": dir_err,
           "note": "vo_bootstrap_dev: cv2.findFundamentalMat's RANSAC restated + essential-matrix decomposition + cheirality "
                   "vote + landmarks, one CTA per problem, inputs resident in HBM"}
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import cv2
        n_cpu = min(S, 8)
        t0 = time.perf_counter()
        for s in range(n_cpu):
            cv2.findFundamentalMat(points1=P1[s].reshape(-1, 2, 1), points2=P2[s].reshape(-1, 2, 1), method=cv2.FM_RANSAC,
                                   ransacReprojThreshold=1.0, confidence=0.999)
        dt = time.perf_counter() - t0
        out["cpu_cv2_findFundamentalMat_problems_per_s"] = n_cpu / dt
        out["cpu_sample"] = f"{n_cpu} of the same problems through cv2.findFundamentalMat alone (the reference's call, triangulation.py:126-134) on one host core"
    return out


def bind_near_gpu(local_rank):
    """Pin this process -- and with it the pinned staging buffers it allocates next -- to the NUMA node of its GPU, so
    that uploads do not cross the inter-socket link.  CPU affinity first (NVML's ideal set when it is a proper subset of
    what we may use); independently of that the MEMORY policy is set to the GPU's node with set_mempolicy(MPOL_PREFERRED),
    which works even when the cpuset does not contain that node's CPUs.  Best effort: returns a short description."""
    notes = []
    node = None
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local_rank]) if vis and vis.split(",")[local_rank].strip().isdigit() else local_rank
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        try:
            bus = nv.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            path = f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node"
            node = int(open(path).read().strip())
        except Exception:
            node = None
        if node is None or node < 0:                  # containers often hide sysfs' numa_node: ask NVML for the memory affinity
            try:
                m = nv.nvmlDeviceGetMemoryAffinity(h, 4, nv.NVML_AFFINITY_SCOPE_NODE)
                bits = [64 * i + b for i, w in enumerate(m) for b in range(64) if (int(w) >> b) & 1]
                node = bits[0] if len(bits) == 1 else None
            except Exception:
                node = None
        n_words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            notes.append(f"cpus: bound to {len(use)} near gpu {phys}")
        else:
            notes.append(f"cpus: no binding ({len(cpus)} ideal, {len(allowed)} allowed)")
    except Exception as ex:   # NVML or the affinity call not available: run unbound
        notes.append(f"cpus: no binding ({type(ex).__name__})")
    try:
        if node is not None and node >= 0:
            libc = C.CDLL(None, use_errno=True)
            MPOL_PREFERRED = 1
            maskbits = (C.c_ulong * 16)()
            maskbits[node // 64] = 1 << (node % 64)
            rc = libc.syscall(238, MPOL_PREFERRED, maskbits, C.c_ulong(16 * 64 + 1))       # __NR_set_mempolicy (x86_64)
            notes.append(f"memory: preferred numa node {node}" if rc == 0 else f"memory: set_mempolicy failed (errno {C.get_errno()})")
        else:
            notes.append("memory: gpu numa node unknown")
    except Exception as ex:
        notes.append(f"memory: no policy ({type(ex).__name__})")
    return "; ".join(notes)


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def workload_text():
    return ("synthetic KITTI-shaped 1241x376 grayscale sequences (camera translating over a textured plane, "
            f"{N_MOVERS} independently moving patches per sequence), chained per-frame flow of src/main.py:248-287: "
            f"Harris detection of {KP} keypoints on every frame, pyramidal KLT (win 17, 3 levels) of the sequence's "
            f"feature table (up to {CAPACITY} rows), P3P-RANSAC (numpy PCG64 stream, adaptive stop, conf {P3P_CONF}) on the "
            "triangulated rows + pose refinement, bearing-angle candidates, DLT triangulation, cheirality check")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300, help="timed steps (300 steps of 148 frames = 0.6 s on one B200)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="vo", choices=["vo", "p3p_sweep", "stress"])
    ap.add_argument("--seqs", type=int, default=148, help="independent sequences per GPU")
    ap.add_argument("--pool", type=int, default=8, help="distinct time steps per sequence, walked back and forth (inputs > L2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    ap.add_argument("--no-single-sequence", action="store_true", help="skip the S = 1 leg (keeps an ncu launch list to one batch size)")
    ap.add_argument("--no-extras", action="store_true", help="only the contract line's own measurements")
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        n_seq = cores
        # a CPU step takes ~1.4 s: keep the whole run within a few minutes whatever K the caller asks for
        ref_steps, ref_warmup = max(1, min(args.steps, 20)), min(args.warmup, 2)
        fps, ms, st = cpu_pipeline_fps(n_seq, ref_steps, ref_warmup, cores)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": ref_steps, "warmup": ref_warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(), "frames_per_step": n_seq, "steps_requested": args.steps,
                       "mean_rows_tracked": st[0], "mean_p3p_population": st[1], "mean_candidates": st[2]},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{n_seq} sequences advance one frame per step on {cores} host processes: oracle C port of "
                                       "harris.py, cv2.calcOpticalFlowPyrLK (the reference's own call), oracle C port of "
                                       "cv2.solvePnP(P3P) + ransac.py's loop with the numpy rng, scipy least_squares refinement "
                                       "(p3p.py:188-213), numpy bookkeeping of matches.py / state.py, numpy SVD triangulation"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch
    import torch.distributed as dist
    from vo import _native as nat
    from vo.pipeline import DETECTOR_HARRIS, SUMMARY_DOUBLES, Pipeline

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_near_gpu(local_rank)         # host thread + pinned buffers on the GPU's own socket (matters from 4 GPUs up)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = nat.Context(local_rank)
    L = nat.lib()
    if args.workload != "vo":
        import bench_extra
        line = bench_extra.run(args, ctx, dev, rank, world, ClockSampler(local_rank), load_peaks())
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return
    S, P = args.seqs, max(2, args.pool)

    # ---- synthetic inputs -------------------------------------------------------------------
    wd = World(1234 + rank, S, P)
    pool_h = wd.pool()
    pool_pinned = torch.from_numpy(pool_h).pin_memory()
    pitch = (W + 15) & ~15
    pool_d = torch.zeros((P, S, H, pitch), dtype=torch.uint8, device=dev)
    pool_d[..., :W] = pool_pinned.to(dev)
    tstream = torch.cuda.Stream(device=dev)          # explicit stream: kernels and timing events share it
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def new_pipeline(n_seq):
        pl = Pipeline(n_seq, H, W, K_INTR, detector=DETECTOR_HARRIS, ctx=ctx, **pipeline_kwargs())
        pl.prime_dev(pool_d[0].data_ptr(), pitch, H * pitch, init_tables=True, stream=stream)
        torch.cuda.synchronize()
        start_tables(pl, n_seq)
        return pl

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm (`value`) -----------------------------------------------------
    pl = new_pipeline(S)

    def step_dev(p, i, n_seq=S):
        fr = pool_d[wd.time_index(i)]
        p.step_dev(fr.data_ptr(), pitch, H * pitch, stream=stream)

    for i in range(1, warmup + 1):
        step_dev(pl, i)
    pl.sync_dev(stream)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    for i in range(args.steps):
        step_dev(pl, warmup + 1 + i)
    pl.sync_dev(stream)                      # the last step's pose / update / detector kernels (internal streams) count
    e1.record(tstream)
    barrier()
    launches = ctx.launch_count() - l0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * S * args.steps / (ms_total / 1e3)

    # sanity + workload statistics from the last step's summary: poses finite and on the true trajectory
    summ_raw = np.empty((S, SUMMARY_DOUBLES))
    ctx.copy_to_host(summ_raw, pl.summary_dev())
    summ = Pipeline.summary_dict(summ_raw)
    t_true = World.true_position(wd.time_index(warmup + args.steps))
    pos_err = np.linalg.norm(summ["pose"][:, :, 3] - t_true, axis=1)
    # monocular VO drifts (landmarks triangulated over two or three frames of 0.085 units of baseline each at depth 20);
    # "valid" = finite, still on the trajectory (a frame of motion is 0.09 units) and most rows carrying landmarks
    if not (np.isfinite(summ["pose"]).all() and np.median(pos_err) < 0.5 and (summ["n_tri"] > 100).mean() > 0.9):
        raise SystemExit(f"bench.py: the timed steps did not produce valid poses (median position error {np.median(pos_err):.3g})")
    stats = {k: float(np.mean(summ[k])) for k in ("n_tracked", "n_kept", "p3p_N", "n_inliers", "n_candidates", "n_tri", "draws", "gn_iters")}
    stats["redetections_last_step"] = int((summ["flags"] & 1).sum())
    stats["median_position_error"] = float(np.median(pos_err))

    # ---- Harris response kernel alone: live roofline ------------------------------------------
    resp_d = torch.empty((S, H, W), dtype=torch.float64, device=dev)

    def harris_only(i):
        fr = pool_d[i % P]
        nat.check(L.vo_harris_response_dev(ctx.handle, fr.data_ptr(), S, H, W, pitch, H * pitch, 9,
                                           C.c_double(0.09), resp_d.data_ptr(), stream), "harris")
    for i in range(3):
        harris_only(i)
    torch.cuda.synchronize()
    n_rf = 20
    e0.record(tstream)
    for i in range(n_rf):
        harris_only(3 + i)
    e1.record(tstream)
    torch.cuda.synchronize()
    harris_ms = e0.elapsed_time(e1) / n_rf
    algo_bytes = S * H * W * (1 + 8)                 # uint8 pixel in, float64 score out
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = algo_bytes / (harris_ms / 1e3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "harris_traffic.json"))).get("dram_bytes_per_launch_at_S")
    except Exception:
        pass
    roofline = {"kernel": "harris_response", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": harris_ms,
                "note": "9 B/pixel (uint8 in, float64 score out) x S frames per launch; traffic from the committed ncu capture "
                        "(profiles/harris_traffic.json, refreshed by tools/refresh_roofline.sh)"}
    del resp_d
    # the dominant kernel of the step (the tracker, ~53 % of the throughput-bound time) is issue bound, not HBM bound: its
    # roofline-style entry is the issue-slot utilisation of the committed ncu capture (not measured in this run)
    roofline_klt = None
    try:
        kj = json.load(open(os.path.join(ROOT, "profiles", "klt_issue.json")))
        roofline_klt = {"kernel": kj["kernel"], "bound": "issue", "achieved": kj["ipc"], "peak": 4.0, "unit": "warp instructions / cycle / SM",
                        "frac": kj["issue_active_pct"] / 100.0, "launch_ms": kj["duration_ms"], "warp_instructions": kj["warp_instructions"],
                        "note": "issue slots active (smsp__issue_active) from the ncu --set full capture committed under profiles/; "
                                "DRAM traffic of the kernel is 171 MB per launch (2.5 % of the HBM peak)", "source": kj["source"]}
    except Exception:
        pass

    # ---- one sequence alone (latency-bound): device-resident steps of a single 1241x376 stream -----
    single = None
    if rank == 0 and not args.no_single_sequence and not args.no_extras:
        pl1 = new_pipeline(1)
        for i in range(1, 6):
            step_dev(pl1, i)
        torch.cuda.synchronize()
        n1 = 100
        e0.record(tstream)
        for i in range(n1):
            step_dev(pl1, 6 + i)
        pl1.sync_dev(stream)
        e1.record(tstream)
        torch.cuda.synchronize()
        ms1 = e0.elapsed_time(e1) / n1
        single = {"frames_per_s": 1e3 / ms1, "ms_per_frame": ms1, "note": "S = 1: one sequence, back-to-back device-resident steps"}
        pl1.close()

    # ---- end-to-end arm through the host-buffer C ABI ------------------------------------------
    pl2 = new_pipeline(S)
    if os.environ.get("VO_BENCH_WC", "1") == "1":
        # upload staging in write-combined page-locked memory (the CPU only writes it): not snooped during the GPU's reads
        pool_np = nat.pinned_empty(pool_h.shape, np.uint8, write_combined=True)
        pool_np[...] = pool_h
    else:
        pool_np = pool_pinned.numpy()
    outs = [torch.zeros((S, SUMMARY_DOUBLES), dtype=torch.float64).pin_memory().numpy() for _ in range(2)]

    def submit(i):
        # documented call order (include/vo_b200.h): the upload of step i+1 is queued, then step i (whose frames were
        # uploaded during step i-1) is submitted together with the download of its summary.  Every step moves h2d + d2h
        # bytes; with two steps in flight the GPU always has the next step queued.
        pl2.prefetch(pool_np[wd.time_index(i + 1)])
        pl2.submit(None, outs[i & 1])

    def run_pipelined(first, n):
        for i in range(n):
            submit(first + i)
            if i > 0:
                pl2.wait()                # the summary of step first + i - 1 is on the host
        if n > 0:
            pl2.wait()

    e2e_steps = args.e2e_steps or args.steps
    pl2.prefetch(pool_np[wd.time_index(1)])
    run_pipelined(1, warmup)
    barrier()
    t0 = time.perf_counter()
    run_pipelined(1 + warmup, e2e_steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * S * e2e_steps / e2e_s
    h2d = S * H * W
    d2h = S * SUMMARY_DOUBLES * 8
    last = Pipeline.summary_dict(outs[(warmup + e2e_steps) & 1])
    if not np.isfinite(last["pose"]).all():
        raise SystemExit("bench.py: the end-to-end arm returned non-finite poses")
    pl2.close()

    # ---- the drop-in class API (one sequence, host arrays between the stages) ------------------
    class_api = None
    if rank == 0 and not args.no_extras:
        try:
            fps_c, err_c = class_api_loop(wd)
            class_api = {"frames_per_s": fps_c, "position_error": err_c,
                         "note": "src/main.py:248-287 through the vo classes (KLT mode, use_opencv=True, refinement on): one sequence, "
                                 "one C-ABI call per stage, numpy bookkeeping on the host; ~500 Shi-Tomasi corners as the reference's KLT mode"}
            if world == 1 and not args.no_cpu_baseline:
                class_api["cpu_reference_frames_per_s"] = class_api_cpu_reference(wd)
        except Exception as ex:
            class_api = {"frames_per_s": None, "note": f"failed: {type(ex).__name__}: {ex}"}

    # ---- two-view bootstrap (main.py:203-231, once per sequence): S problems in one launch ----------
    boot = None
    if rank == 0 and not args.no_extras:
        try:
            boot = bootstrap_leg(ctx, dev, stream, tstream, S, with_cpu=(world == 1 and not args.no_cpu_baseline))
        except Exception as ex:
            boot = {"ms_per_call": None, "note": f"failed: {type(ex).__name__}: {ex}"}

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            fps, ms, st = cpu_pipeline_fps(cores, 2, 1, cores)
            cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                   "sample": f"{cores} sequences x 2 timed steps (1 warm-up) on {cores} host processes: the loop oracle (C port of harris.py "
                             "and of the P3P solver, cv2.calcOpticalFlowPyrLK, ransac.py's loop, scipy least_squares, numpy "
                             f"bookkeeping / SVD), same per-frame workload; rows tracked {st[0]:.0f}, P3P population {st[1]:.0f}"}
        except Exception as ex:  # the oracle is a checker; its absence must not break the product bench
            cpu = {"value": None, "unit": "frames/s", "cores": cores, "kind": "port", "sample": f"unavailable: {ex}"}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(), "sequences_per_gpu": S, "frames_per_step": world * S,
                       "l2_policy": f"inputs larger than L2: {P} time steps of {S} frames walked back and forth "
                                    f"({P * S * H * pitch / 1e6:.0f} MB) + {S * H * W * 8 / 1e6:.0f} MB score maps per step",
                       "parallelism": f"{world} x independent sequence shards, no collective", "host_binding": numa,
                       "per_frame_means_last_step": stats},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "h2d_gbs_per_gpu": h2d * e2e_steps / e2e_s / 1e9,
                    "api": "vo_pipeline_prefetch_host + vo_pipeline_submit_host + vo_pipeline_wait_host: only the frames go up "
                           "(pinned host buffers), poses + counters come back, every step; two steps in flight"},
            "gpu_launches": int(launches) * world,
            "roofline": roofline, "roofline_tracker": roofline_klt, "cpu_baseline": cpu, "clocks": clocks, "single_sequence": single, "class_api": class_api,
            "bootstrap": boot,
        }))
    pl.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
