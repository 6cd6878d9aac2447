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
    np.savez_compressed(os.path.join(OUT, "harris.npz"), **out)
    print("harris.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    which = sys.argv[1:] or ["harris", "klt", "p3p", "triangulation"]
    for w in which:
        globals()["make_" + w]()
