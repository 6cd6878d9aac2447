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
