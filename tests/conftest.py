import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "visual-odometry-project_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


@pytest.fixture(scope="session")
def ctx():
    """One vo_ctx on cuda:0 for the GPU tests; fails loudly if the library or GPU is missing."""
    from vo import _native as nat
    return nat.default_context(0)


def synthetic_image(h, w, seed, blobs=True):
    """Seeded textured test image: smoothed noise plus rectangles (corners)."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h, w)).astype(np.float64)
    # cheap separable blur so the response has smooth hills like a natural image
    k = np.array([1, 4, 6, 4, 1], dtype=np.float64) / 16.0
    for _ in range(2):
        img = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, img)
        img = np.apply_along_axis(lambda c: np.convolve(c, k, mode="same"), 0, img)
    img = (img - img.min()) / (img.max() - img.min()) * 255.0
    if blobs:
        for _ in range(max(4, (h * w) // 4000)):
            y, x = rng.integers(0, h - 8), rng.integers(0, w - 8)
            hh, ww = rng.integers(4, 24), rng.integers(4, 24)
            img[y:y + hh, x:x + ww] = rng.integers(0, 256)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)
