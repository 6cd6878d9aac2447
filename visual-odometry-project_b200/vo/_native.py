"""ctypes binding of libvo_b200.so (include/vo_b200.h).

This is the only way the Python `vo` package reaches the GPU: there is no CPU fallback and no
import of the test oracle.  If the shared library is missing or no B200 is visible, every
compute entry point raises.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libvo_b200.so")

_lib = None
_lock = threading.Lock()


class VoNativeError(RuntimeError):
    pass


def _sig(fn, restype, argtypes):
    fn.restype = restype
    fn.argtypes = argtypes


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise VoNativeError(
                f"{LIB_PATH} not found: build it with `python visual-odometry-project_b200/build.py` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, i32, u64, sz, dbl = C.c_void_p, C.c_int, C.c_ulonglong, C.c_size_t, C.c_double
        _sig(L.vo_abi_version, i32, [])
        _sig(L.vo_last_error, C.c_char_p, [])
        _sig(L.vo_ctx_create, i32, [C.POINTER(vp), i32])
        _sig(L.vo_ctx_destroy, None, [vp])
        _sig(L.vo_ctx_launch_count, u64, [vp])
        _sig(L.vo_ctx_synchronize, i32, [vp])
        _sig(L.vo_ctx_stream, vp, [vp])
        _sig(L.vo_copy_to_host, i32, [vp, vp, vp, sz])
        _sig(L.vo_harris_response_dev, i32, [vp, vp, i32, i32, i32, sz, sz, i32, dbl, vp, vp])
        _sig(L.vo_harris_nms_dev, i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp])
        _sig(L.vo_harris_detect_dev, i32, [vp, vp, i32, i32, i32, sz, sz, i32, dbl, i32, i32, vp, vp, vp])
        _sig(L.vo_harris_descriptors_dev, i32, [vp, vp, i32, i32, i32, sz, sz, vp, i32, i32, vp, vp])
        _sig(L.vo_harris_detect_host, i32, [vp, vp, i32, i32, i32, i32, dbl, i32, i32, i32, vp, vp, vp])
        _sig(L.vo_harris_descriptors_host, i32, [vp, vp, i32, i32, vp, i32, i32, vp])
        _optional = {
            "vo_klt_pyramid_layout": (i32, [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
            "vo_klt_build_pyramid_dev": (i32, [vp, vp, i32, i32, i32, sz, sz, i32, i32, vp, vp]),
            "vo_klt_track_dev": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, dbl, dbl, vp, i32, vp, vp, vp, vp]),
            "vo_klt_track_host": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, dbl, dbl, vp, i32, vp, vp, vp]),
            "vo_p3p_ransac_score_dev": (i32, [vp, vp, vp, i32, i32, vp, vp, i32, dbl, i32, vp, vp, vp, vp]),
            "vo_p3p_ransac_select_dev": (i32, [vp, vp, vp, i32, i32, vp, vp, vp, vp, i32, dbl, i32, vp, i32, i32, i32,
                                               vp, vp, vp, vp, vp, vp]),
            "vo_p3p_ransac_host": (i32, [vp, vp, vp, i32, i32, vp, vp, i32, dbl, i32, vp, i32, i32, i32,
                                         vp, vp, vp, vp, vp, vp, vp, vp]),
            "vo_triangulate_dev": (i32, [vp, vp, vp, i32, vp, i32, vp, i32, vp, vp]),
            "vo_triangulate_host": (i32, [vp, vp, vp, i32, vp, i32, vp, i32, vp]),
            "vo_cv_rng_subsets_host": (i32, [vp, i32, i32, i32, vp]),
            "vo_bootstrap_dev": (i32, [vp, vp, vp, i32, i32, vp, vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp, vp]),
            "vo_bootstrap_host": (i32, [vp, vp, vp, i32, i32, vp, vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp]),
        }
        _optional.update({
            "vo_match_descriptors_dev": (i32, [vp, vp, vp, i32, i32, i32, i32, dbl, vp, vp, vp]),
            "vo_match_descriptors_host": (i32, [vp, vp, vp, i32, i32, i32, i32, dbl, vp, vp]),
            "vo_frontend_create": (i32, [vp, vp, C.POINTER(vp)]),
            "vo_frontend_destroy": (None, [vp]),
            "vo_frontend_outputs": (i32, [vp, vp]),
            "vo_frontend_next_frame_slot": (vp, [vp, C.POINTER(sz), C.POINTER(sz)]),
            "vo_frontend_step_dev": (i32, [vp, vp, sz, sz, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]),
            "vo_frontend_prefetch_host": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
            "vo_frontend_step_host": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp,
                                            vp, vp, vp, vp, vp, vp, vp, vp]),
            "vo_frontend_submit_host": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp,
                                              vp, vp, vp, vp, vp, vp, vp, vp]),
            "vo_frontend_wait_host": (i32, [vp]),
            "vo_pipeline_create": (i32, [vp, vp, C.POINTER(vp)]),
            "vo_pipeline_destroy": (None, [vp]),
            "vo_pipeline_prime_dev": (i32, [vp, vp, sz, sz, i32, vp]),
            "vo_pipeline_prime_host": (i32, [vp, vp, i32]),
            "vo_pipeline_step_dev": (i32, [vp, vp, sz, sz, vp]),
            "vo_pipeline_bootstrap_dev": (i32, [vp, vp, sz, sz, dbl, dbl, i32, vp]),
            "vo_pipeline_bootstrap_host": (i32, [vp, vp, dbl, dbl, i32, vp]),
            "vo_pipeline_summary_dev": (vp, [vp]),
            "vo_pipeline_sync_dev": (i32, [vp, vp]),
            "vo_pipeline_prefetch_host": (i32, [vp, vp]),
            "vo_pipeline_submit_host": (i32, [vp, vp, vp]),
            "vo_pipeline_wait_host": (i32, [vp]),
            "vo_pipeline_step_host": (i32, [vp, vp, vp]),
            "vo_pipeline_read_table_host": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
            "vo_pipeline_write_table_host": (i32, [vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp]),
            "vo_pipeline_read_detections_host": (i32, [vp, i32, vp, vp]),
            "vo_test_pcg64_choice4_host": (i32, [vp, vp, i32, i32, vp]),
            "vo_klt_track_bgr_host": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, dbl, dbl, vp, i32, vp, vp, vp]),
            "vo_klt_cache_hits": (u64, [vp]),
            "vo_bgr2gray_dev": (i32, [vp, vp, i32, i32, i32, sz, sz, vp, sz, sz, vp]),
            "vo_bgr2gray_host": (i32, [vp, vp, i32, i32, i32, vp]),
            "vo_gftt_dev": (i32, [vp, vp, i32, i32, i32, sz, sz, i32, dbl, dbl, i32, vp, vp, vp, vp, vp]),
            "vo_gftt_host": (i32, [vp, vp, i32, i32, i32, i32, dbl, dbl, i32, vp, vp, vp, vp]),
            "vo_test_dfma_peak": (i32, [vp, vp]),
            "vo_host_alloc": (i32, [C.POINTER(vp), sz, i32]),
            "vo_host_free": (i32, [vp]),
            "vo_refine_pose_dev": (i32, [vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp]),
            "vo_refine_pose_host": (i32, [vp, vp, vp, vp, i32, i32, vp, vp, vp, vp]),
        })
        for name, (rt, at) in _optional.items():
            if hasattr(L, name):  # all are present in a complete build; tests/test_abi.py checks that
                _sig(getattr(L, name), rt, at)
        _lib = L
        return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().vo_last_error().decode("utf-8", "replace")
        raise VoNativeError(f"{what} failed (rc={rc}): {msg}")


class Context:
    """Owns a vo_ctx (device scratch + stream) on one GPU."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().vo_ctx_create(C.byref(self._h), int(device)), "vo_ctx_create")
        self.device = int(device)

    @property
    def handle(self):
        return self._h

    def launch_count(self) -> int:
        return int(lib().vo_ctx_launch_count(self._h))

    def synchronize(self) -> None:
        check(lib().vo_ctx_synchronize(self._h), "vo_ctx_synchronize")

    def stream(self) -> int:
        return int(lib().vo_ctx_stream(self._h) or 0)

    def copy_to_host(self, dst: np.ndarray, d_src: int) -> np.ndarray:
        """Synchronous D2H copy of a resident result buffer into a contiguous numpy array."""
        check(lib().vo_copy_to_host(self._h, ptr(dst), C.c_void_p(d_src), dst.nbytes), "vo_copy_to_host")
        return dst

    def close(self):
        if self._h:
            lib().vo_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx


def pinned_empty(shape, dtype=np.uint8, write_combined=False) -> np.ndarray:
    """numpy array over page-locked host memory (vo_host_alloc); freed when the array is collected."""
    import weakref
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    check(lib().vo_host_alloc(C.byref(p), n, int(bool(write_combined))), "vo_host_alloc")
    buf = (C.c_ubyte * n).from_address(p.value)
    a = np.frombuffer(buf, dtype=dtype).reshape(shape)
    weakref.finalize(buf, lib().vo_host_free, p)
    return a


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class FrontendParams(C.Structure):
    """vo_frontend_params of include/vo_b200.h."""
    _fields_ = [("n_seq", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("patch_size", C.c_int), ("kappa", C.c_double), ("nms_radius", C.c_int), ("num_keypoints", C.c_int),
                ("klt_win", C.c_int), ("klt_max_level", C.c_int), ("klt_max_iters", C.c_int),
                ("klt_epsilon", C.c_double), ("klt_min_eig", C.c_double),
                ("n_corr", C.c_int), ("n_hyp", C.c_int), ("p3p_threshold", C.c_double),
                ("n_tri", C.c_int), ("tri_mode", C.c_int)]


class FrontendOutputs(C.Structure):
    """vo_frontend_outputs_t of include/vo_b200.h (device pointers)."""
    _fields_ = [("d_resp", C.c_void_p), ("d_kp_xy", C.c_void_p), ("d_tracked", C.c_void_p), ("d_status", C.c_void_p),
                ("d_err", C.c_void_p), ("d_best4", C.c_void_p), ("d_inliers", C.c_void_p), ("d_pose", C.c_void_p),
                ("d_tri_out", C.c_void_p), ("d_counts", C.c_void_p), ("d_cur_pyramid", C.c_void_p),
                ("pyr_pitch0", C.c_size_t), ("pyr_frame_bytes", C.c_size_t)]
