"""numpy-in / numpy-out wrappers over the host-buffer C ABI (include/vo_b200.h).

These are what the `vo` classes call.  Every function runs hand-written CUDA on the GPU through
libvo_b200.so; nothing here computes on the CPU beyond shaping arguments.
"""
import ctypes as C

import numpy as np

from . import _native as nat


def _ctx(ctx=None):
    return ctx if ctx is not None else nat.default_context(0)


# ------------------------------------------------------------------------------------------------
# Harris  (reference: src/vo/features/harris.py)
# ------------------------------------------------------------------------------------------------
def harris_detect(img, num_keypoints=1000, patch_size=9, kappa=0.09, nms_radius=5, desc_radius=None,
                  want_response=False, ctx=None):
    """extractKeypoints (+ optionally extractDescriptors) for one frame (H, W) or a batch (F, H, W).

    Returns (kp_xy int32 [..., K, 2], response float64 [..., H, W] | None, desc uint8 [..., K, D] | None).
    """
    ctx = _ctx(ctx)
    a = np.ascontiguousarray(img, dtype=np.uint8)
    single = a.ndim == 2
    if single:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("harris_detect: image must be (H, W) or (F, H, W) uint8 grayscale")
    F, H, W = a.shape
    kp = np.empty((F, num_keypoints, 2), dtype=np.int32)
    resp = np.empty((F, H, W), dtype=np.float64) if want_response else None
    desc = None
    dr = 0
    if desc_radius is not None:
        dr = int(desc_radius)
        desc = np.empty((F, num_keypoints, (2 * dr + 1) ** 2), dtype=np.uint8)
    rc = nat.lib().vo_harris_detect_host(
        ctx.handle, nat.ptr(a), F, H, W, int(patch_size), C.c_double(kappa), int(nms_radius), int(num_keypoints),
        dr, nat.ptr(resp) if resp is not None else None, nat.ptr(kp), nat.ptr(desc) if desc is not None else None)
    nat.check(rc, "vo_harris_detect_host")
    if single:
        return kp[0], (resp[0] if resp is not None else None), (desc[0] if desc is not None else None)
    return kp, resp, desc
