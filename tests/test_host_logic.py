"""CPU: host-side logic of the `vo` package (no GPU compute): helpers, primitives, RANSAC loop, the
C ABI's exported symbols, sharding across ranks (gloo, world_size 2).  Mirrors the reference's
tests/test_helpers.py, tests/test_features.py, tests/test_ransac.py."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT


# ---------------------------------------------------------------- helpers (reference tests/test_helpers.py)
def test_homogeneous_cartesian_roundtrip():
    from vo.helpers import to_cartesian_coordinates, to_homogeneous_coordinates
    pts = np.array([[[1], [2]], [[3], [4]]])
    hom = to_homogeneous_coordinates(pts)
    assert np.array_equal(hom, np.array([[[1], [2], [1]], [[3], [4], [1]]]))
    assert np.array_equal(to_cartesian_coordinates(np.array([[[1], [2], [1]], [[3], [4], [2]]])),
                          np.array([[[1], [2]], [[1.5], [2]]]))
    with pytest.warns(RuntimeWarning):
        to_cartesian_coordinates(np.array([[[1], [2], [0]], [[3], [4], [0]]]))


@pytest.mark.parametrize("D,N", [(2, 4), (3, 3), (2, 1000)])
def test_normalize_points(D, N):
    from vo.helpers import normalize_points, to_cartesian_coordinates, to_homogeneous_coordinates
    pts = np.random.default_rng(2023).normal(-3, 10, size=(N, D, 1))
    norm, T = normalize_points(pts)
    assert np.allclose(np.mean(norm, axis=0), 0)
    assert np.allclose(np.sqrt(np.mean(np.sum(norm ** 2, axis=-2))), np.sqrt(D))
    assert np.allclose(norm, to_cartesian_coordinates(T @ to_homogeneous_coordinates(pts)))


def test_skew_and_twist():
    from vo.helpers import H_matrix_to_twist, to_skew_symmetric_matrix, twist_to_H_matrix
    assert np.array_equal(to_skew_symmetric_matrix(np.array([[1, 2, 3]]).T), np.array([[0, -3, 2], [3, 0, -1], [-2, 1, 0]]))
    batch = to_skew_symmetric_matrix(np.array([[1, 2, 3], [4, 5, 6]]).reshape(2, 3, 1))
    assert np.array_equal(batch[1], np.array([[0, -6, 5], [6, 0, -4], [-5, 4, 0]]))
    tw = np.array([0.1, -0.2, 0.3, 0.05, 0.02, -0.04])
    assert np.allclose(H_matrix_to_twist(twist_to_H_matrix(tw)), tw)


# ---------------------------------------------------------------- primitives (reference tests/test_features.py)
def test_features_basics():
    from vo.primitives import Features
    kp = np.array([[1, 2], [3, 4]]).reshape(-1, 2, 1)
    lm = np.array([[9, 10, 11], [12, 13, 14]]).reshape(-1, 3, 1)
    f = Features(kp, lm)
    assert np.array_equal(f.keypoints, kp) and np.array_equal(f.landmarks, lm)
    g = Features(kp)
    assert g.descriptors is None and np.all(np.isnan(g.landmarks)) and g.length == 2 and np.allclose(g.state, 0)
    g.mask(np.array([True, False]))
    assert g.length == 1 and g.tracks.shape == (1, 2, 1) and g.poses.shape == (1, 4, 4)
    with pytest.raises(AssertionError):
        Features(np.zeros((3, 2)))


def test_matches_and_state_vs_reference(golden):
    """Block re-ordering of Matches and the State updates, against arrays produced by the reference."""
    from vo.primitives import Features, Frame, Matches, State
    from vo.sensors import Camera
    g = golden("bookkeeping")
    f1 = Features(g["kp1"].copy())
    f1.state, f1.landmarks, f1.tracks, f1.poses = g["state1"].copy(), g["land1"].copy(), g["tracks1"].copy(), g["poses1"].copy()
    f1.descriptors = g["desc1"].copy()
    f2 = Features(g["kp2"].copy())
    f2.descriptors = g["desc2"].copy()
    K = np.array([[500, 0, 150], [0, 500, 150], [0, 0, 1.0]])
    fr1, fr2 = Frame(None, features=f1, sensor=Camera(K)), Frame(None, features=f2, sensor=Camera(K))
    mt = Matches(fr1, fr2, g["matches"])
    for tag, f in (("a", mt.frame1.features), ("b", mt.frame2.features)):
        for name, col in (("kp", f.keypoints), ("state", f.state), ("land", f.landmarks), ("tracks", f.tracks),
                          ("poses", f.poses), ("desc", f.descriptors)):
            assert np.array_equal(col, g[f"{tag}_{name}"], equal_nan=True), (tag, name)
    st = State(fr1)
    st.update_from_matches(mt)
    st.update_with_world_pose(g["s_pose_in"])
    st.reset_outliers(g["s_outliers"])
    st.compute_candidates()
    f = st.curr_frame.features
    assert np.allclose(st.get_pose(), g["s_pose"], atol=0, rtol=0)
    assert np.array_equal(f.state, g["s_state"]) and np.array_equal(f.candidate_mask, g["s_cand"])
    assert np.array_equal(f.tracks, g["s_tracks"], equal_nan=True) and np.array_equal(f.poses, g["s_poses"], equal_nan=True)


# ---------------------------------------------------------------- RANSAC (reference tests/test_ransac.py)
def test_ransac_parabola():
    from vo.algorithms import RANSAC
    rng = np.random.default_rng(2023)
    poly = rng.uniform(size=[3, 1])
    extremum = -poly[1] / (2 * poly[0])
    xstart = extremum - 0.5
    lowest, highest = np.polyval(poly, extremum), np.polyval(poly, xstart)
    yspan = highest - lowest
    max_noise = 0.1 * yspan
    x = rng.uniform(size=[1, 20]) + xstart
    y = np.polyval(poly, x)
    y = y + (rng.uniform(size=y.shape) - 0.5) * 2 * max_noise
    data = np.concatenate([np.concatenate([x, rng.uniform(size=[1, 10]) + xstart], axis=1),
                           np.concatenate([y, rng.uniform(size=[1, 10]) * yspan + lowest], axis=1)], axis=0).T
    r = RANSAC(3, data, lambda s: np.polyfit(s[:, 0], s[:, 1], 2),
               lambda p, pts: np.abs(np.polyval(p, pts[:, 0]) - pts[:, 1]), max_noise + 1e-5, 1 / 3, 0.99)
    model, inl = r.find_best_model()
    xs = np.linspace(data[:, 0].min(), data[:, 0].max(), 100)
    assert np.allclose(np.polyval(poly, xs), np.polyval(model, xs), atol=2e-3)   # the reference's own tolerance


def test_ransac_iteration_table_matches_oracle():
    import oracle
    from vo.algorithms import RANSAC
    r = RANSAC(4, None, None, None, 1.0, 0.9, 0.99, 1000, p3p=True)
    assert np.array_equal(r.iterations_table(600), oracle.ransac_iterations_table(600, 4, 0.99, 1000))
    assert r.n_iterations == oracle.ransac_initial_iterations(4, 0.9, 0.99, 1000)


# ---------------------------------------------------------------- C ABI
def test_abi_exports_every_declared_symbol():
    from vo import _native as nat
    header = open(os.path.join(ROOT, "include", "vo_b200.h")).read()
    names = set(re.findall(r"\b(vo_[a-z0-9_]+)\s*\(", header))
    names -= {"vo_ctx", "vo_frontend"}
    assert len(names) >= 25
    L = nat.lib()
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert L.vo_abi_version() == 2
    out = subprocess.run(["nm", "-D", "--defined-only", nat.LIB_PATH], capture_output=True, text=True).stdout
    for n in names:
        assert f" T {n}" in out, n


def test_no_cpu_fallback_and_no_oracle_in_product():
    """Without a GPU every compute entry point must fail loudly; the product never imports oracle/."""
    import torch
    from vo import _native as nat
    if not torch.cuda.is_available():
        with pytest.raises(nat.VoNativeError, match="no CUDA device"):
            nat.Context(0)
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("the test oracle", ""), f


# ---------------------------------------------------------------- multi-rank host logic (gloo, world_size 2)
def test_shard_range():
    from vo.sharding import shard_range
    for n, w in [(8, 2), (7, 3), (1, 4), (148, 8)]:
        blocks = [shard_range(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys, time
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from vo.sharding import shard_range, job_frames_per_second
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(9, rank, world)            # 9 sequences over 2 ranks -> 5 + 4, no exchange of data
secs = 0.010 * (rank + 1)                       # rank 1 is the slow one
t = torch.tensor([secs], dtype=torch.float64)
dist.barrier()
dist.all_reduce(t, op=dist.ReduceOp.MAX)        # timing is the max over ranks
n = torch.tensor([hi - lo], dtype=torch.int64)
dist.all_reduce(n, op=dist.ReduceOp.SUM)
if rank == 0:
    print("RESULT", int(n.item()), float(t.item()), job_frames_per_second([5, 4], [0.010, 0.020]))
dist.destroy_process_group()
"""


def test_two_rank_sharding_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), PKG],
                         capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("RESULT")][0].split()
    assert int(line[1]) == 9 and abs(float(line[2]) - 0.020) < 1e-12 and abs(float(line[3]) - 450.0) < 1e-9


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores) runs without a GPU and prints one JSON line
    with the keys the driver reads"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_cv_rng_subsets_native_equals_python():
    """vo_cv_rng_subsets_host (the native getSubset loop the default pose path uses) against the literal Python loop:
    same subsets, same generator state afterwards."""
    from vo.algorithms.cv_ransac import CvRNG, subset4, subsets
    for n, count in [(37, 500), (5, 100), (1000, 256), (4, 10)]:
        a, b = CvRNG(), CvRNG()
        want = np.array([subset4(a, n) for _ in range(count)], dtype=np.int32)
        got = subsets(b, n, count)
        assert np.array_equal(got, want) and a.state == b.state, (n, count)
