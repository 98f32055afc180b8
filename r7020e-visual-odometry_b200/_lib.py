"""ctypes loader for libvo_b200.so (the C ABI in include/vo_b200.h).

The library is built in-tree (``python __graft_entry__.py`` or ``make -C csrc``).  There is no CPU
fallback: if the shared object is missing, or no sm_100 device is present when a compute call is
made, the call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libvo_b200.so")
_LIB = None


class VoError(RuntimeError):
    pass


class SiftOpts(C.Structure):
    _fields_ = [("contrast_threshold", C.c_float), ("edge_threshold", C.c_float),
                ("num_layers_in_octave", C.c_int), ("sigma", C.c_float), ("index_base", C.c_int)]


class Keypoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32)]


class MatchOpts(C.Structure):
    _fields_ = [("match_threshold", C.c_float), ("max_ratio", C.c_float), ("unique", C.c_int),
                ("index_base", C.c_int)]


class P3POpts(C.Structure):
    _fields_ = [("max_num_trials", C.c_int), ("confidence", C.c_double),
                ("max_reproj_error", C.c_double), ("seed", C.c_uint64), ("adaptive", C.c_int)]


class FramesOpts(C.Structure):
    _fields_ = [("sift", SiftOpts), ("match", MatchOpts), ("p3p", P3POpts),
                ("max_keypoints", C.c_int), ("first_frame", C.c_int), ("col_major", C.c_int)]


# every symbol include/vo_b200.h declares (tests check that the .so exports each one)
EXPORTS = [
    "vo_version", "vo_last_error", "vo_ctx_create", "vo_ctx_destroy", "vo_ctx_sync",
    "vo_ctx_stream", "vo_profile_enable", "vo_kernel_launches", "vo_profile_count", "vo_profile_get", "vo_frames_dev", "vo_sift", "vo_sift_batch", "vo_sift_stack", "vo_match", "vo_match_top2", "vo_match_dev",
    "vo_match_top2_dev", "vo_match_best2_dev", "vo_match_best2_gather_dev", "vo_landmarks_prepare", "vo_peer_alloc", "vo_peer_open", "vo_peer_close", "vo_peer_free", "vo_match_stats", "vo_match_debug_gemm", "vo_triangulate", "vo_p3p",
    "vo_frames", "vo_frames_landmarks", "vo_frames_use_graph", "vo_frames_graph_state", "vo_png_info", "vo_png_decode_gray8", "vo_png_read_batch", "vo_inflate_zlib", "vo_png_decode_batch_dev", "vo_png_read_batch_dev",
]


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise VoError(f"{SO_PATH} is missing: build it with `python __graft_entry__.py` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(SO_PATH)
        L.vo_last_error.restype = C.c_char_p
        L.vo_ctx_stream.restype = C.c_void_p
        L.vo_ctx_stream.argtypes = [C.c_void_p]
        L.vo_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.vo_ctx_destroy.argtypes = [C.c_void_p]
        L.vo_ctx_destroy.restype = None
        L.vo_ctx_sync.argtypes = [C.c_void_p]
        L.vo_kernel_launches.restype = C.c_longlong
        L.vo_kernel_launches.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def check(rc, allow=()):
    if rc != 0 and rc not in allow:
        raise VoError(f"libvo_b200 error {rc}: {lib().vo_last_error().decode()}")
    return rc
