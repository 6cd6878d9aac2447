"""GPU parity: DLT triangulation (C ABI) vs the oracle (numpy SVD, the reference's own solver) and
the reference's golden outputs.  Tolerance: 1e-9 relative on these well-conditioned geometries."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _close(a, b, rel=1e-9):
    return np.abs(a - b).max() <= rel * max(1.0, np.abs(b).max())


def test_golden_reference(ctx, golden):
    from vo import _ops
    g = golden("triangulation")
    for tag in ("clean", "noisy"):
        X = _ops.triangulate(g[f"{tag}_p1"], g[f"{tag}_p2"], g["C1"], g["C2"], mode=0, ctx=ctx)
        assert _close(X, g[f"{tag}_linear"])
        K = g["C1"][:, :3]
        proj1 = K @ np.linalg.inv(g[f"{tag}_cand_poses"])[:, :3]
        proj2 = K @ np.linalg.inv(g[f"{tag}_cand_current_pose"])[:3]
        X0 = _ops.triangulate(g[f"{tag}_cand_tracks"], g[f"{tag}_p2"], proj1, proj2, mode=0, ctx=ctx)
        X1 = _ops.triangulate(g[f"{tag}_cand_tracks"], g[f"{tag}_p2"], proj1, proj2, mode=1, ctx=ctx)
        assert _close(X0, g[f"{tag}_cand_cv0"])
        assert _close(X1, g[f"{tag}_cand_cv1"], rel=1e-7)
    # the reference's own acceptance test (tests/test_triangulation.py): atol 1e-4 to ground truth
    X = _ops.triangulate(g["clean_p1"], g["clean_p2"], g["C1"], g["C2"], mode=0, ctx=ctx)
    assert np.allclose(X, g["landmarks"], atol=1e-4)


@pytest.mark.parametrize("n", [1, 33, 5000])
def test_vs_oracle_random(ctx, n):
    from vo import _ops
    rng = np.random.default_rng(n)
    K = np.array([[707.0912, 0, 601.8873], [0, 707.0912, 183.1104], [0, 0, 1.0]])
    X = rng.uniform(-8, 8, (n, 3))
    X[:, 2] = rng.uniform(5, 60, n)
    C2 = K @ np.hstack([np.eye(3), np.array([[-0.7], [0.02], [0.1]])])
    C1 = np.stack([K @ np.hstack([np.eye(3), rng.normal(0, 0.2, (3, 1))]) for _ in range(n)])
    def proj(C, X):
        x = np.einsum("...ij,nj->...ni" if C.ndim == 2 else "nij,nj->ni", C, np.hstack([X, np.ones((n, 1))]))
        return x[:, :2] / x[:, 2:]
    p1 = proj(C1, X) + rng.normal(0, 0.3, (n, 2))
    p2 = proj(C2, X) + rng.normal(0, 0.3, (n, 2))
    for mode in (0, 1):
        got = _ops.triangulate(p1, p2, C1, C2, mode=mode, ctx=ctx)
        ref = oracle.triangulate(p1, p2, C1, C2, mode=mode)
        rel = np.abs(got - ref).max(axis=1) / np.maximum(1.0, np.abs(ref).max(axis=1))
        assert rel.max() < 1e-8, rel.max()
    assert _ops.triangulate(np.zeros((0, 2)), np.zeros((0, 2)), C2, C2, ctx=ctx).shape == (0, 3)
