"""ctypes binding of the CPU oracle (oracle/libvo_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does (see oracle/vo_oracle.h).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _KP(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32)]


class _SiftOpts(C.Structure):
    _fields_ = [("n_octave_layers", C.c_int), ("contrast_threshold", C.c_float),
                ("edge_threshold", C.c_float), ("sigma", C.c_float)]


class _MatchOpts(C.Structure):
    _fields_ = [("match_threshold", C.c_float), ("max_ratio", C.c_float), ("unique", C.c_int),
                ("index_base", C.c_int)]


class _P3POpts(C.Structure):
    _fields_ = [("max_num_trials", C.c_int), ("confidence", C.c_double),
                ("max_reproj_error", C.c_double), ("seed", C.c_uint64), ("adaptive", C.c_int)]


KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"),
                     ("response", "f4"), ("octave", "i4")])


def build(force=False):
    so = os.path.join(_HERE, "libvo_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("sift.c", "match.c", "geom.c", "vo_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libvo_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.vo_oracle_expf.restype = C.c_float
        _LIB.vo_oracle_expf.argtypes = [C.c_float]
        _LIB.vo_oracle_atan2deg.restype = C.c_float
        _LIB.vo_oracle_atan2deg.argtypes = [C.c_float, C.c_float]
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def sift(img, n_octave_layers=3, contrast_threshold=0.04, edge_threshold=10.0, sigma=1.6,
         capacity=32768):
    """img: HxW uint8 (row-major). Returns (kps structured array, desc n x 128 float32)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    kps = np.zeros(capacity, dtype=KP_DTYPE)
    desc = np.zeros((capacity, 128), dtype=np.float32)
    o = _SiftOpts(n_octave_layers, contrast_threshold, edge_threshold, sigma)
    n = lib().vo_oracle_sift(_p(img, C.c_uint8), img.shape[0], img.shape[1], img.shape[1],
                             C.byref(o), capacity, kps.ctypes.data_as(C.POINTER(_KP)),
                             _p(desc, C.c_float))
    if n < 0:
        raise RuntimeError("oracle sift: capacity too small")
    return kps[:n].copy(), desc[:n].copy()


def blur(src, sigma):
    src = np.ascontiguousarray(src, dtype=np.float32)
    dst = np.empty_like(src)
    lib().vo_oracle_blur(_p(src, C.c_float), _p(dst, C.c_float), src.shape[0], src.shape[1],
                         C.c_float(sigma))
    return dst


def base_image(img, sigma=1.6):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.empty((img.shape[0] * 2, img.shape[1] * 2), dtype=np.float32)
    lib().vo_oracle_base_image(_p(img, C.c_uint8), img.shape[0], img.shape[1], img.shape[1],
                               C.c_float(sigma), _p(out, C.c_float))
    return out


def gauss_kernel(sigma):
    taps = np.zeros(64, dtype=np.float32)
    r = C.c_int(0)
    lib().vo_oracle_gauss_kernel(C.c_float(sigma), C.byref(r), _p(taps, C.c_float))
    return r.value, taps[: r.value + 1].copy()


def match(f1, f2, match_threshold=1.0, max_ratio=0.6, unique=False, index_base=0):
    """f1: n1 x d, f2: n2 x d float32 (row-major numpy).  Returns (pairs P x 2 uint32, metric)."""
    f1 = np.ascontiguousarray(f1, dtype=np.float32)
    f2 = np.ascontiguousarray(f2, dtype=np.float32)
    n1, n2 = f1.shape[0], f2.shape[0]
    dim = f1.shape[1] if f1.ndim == 2 else 128
    i1 = np.zeros(max(n1, 1), dtype=np.uint32)
    i2 = np.zeros(max(n1, 1), dtype=np.uint32)
    m = np.zeros(max(n1, 1), dtype=np.float32)
    o = _MatchOpts(match_threshold, max_ratio, int(unique), index_base)
    p = lib().vo_oracle_match(_p(f1, C.c_float), n1, _p(f2, C.c_float), n2, dim, 0, C.byref(o),
                              _p(i1, C.c_uint32), _p(i2, C.c_uint32), _p(m, C.c_float))
    return np.stack([i1[:p], i2[:p]], axis=1), m[:p].copy()


def match_top2(f1, f2):
    f1 = np.ascontiguousarray(f1, dtype=np.float32)
    f2 = np.ascontiguousarray(f2, dtype=np.float32)
    n1, n2, dim = f1.shape[0], f2.shape[0], f1.shape[1]
    j1 = np.zeros(n1, dtype=np.uint32)
    s1 = np.zeros(n1, dtype=np.float32)
    s2 = np.zeros(n1, dtype=np.float32)
    lib().vo_oracle_match_top2(_p(f1, C.c_float), n1, _p(f2, C.c_float), n2, dim, 0,
                               _p(j1, C.c_uint32), _p(s1, C.c_float), _p(s2, C.c_float))
    return j1, s1, s2


def triangulate(pts1, pts2, P1, P2):
    pts1 = np.ascontiguousarray(pts1, dtype=np.float64).reshape(-1, 2)
    pts2 = np.ascontiguousarray(pts2, dtype=np.float64).reshape(-1, 2)
    P1 = np.ascontiguousarray(P1, dtype=np.float64).reshape(3, 4)
    P2 = np.ascontiguousarray(P2, dtype=np.float64).reshape(3, 4)
    n = pts1.shape[0]
    xyz = np.zeros((n, 3)); err = np.zeros(n); valid = np.zeros(n, dtype=np.uint8)
    lib().vo_oracle_triangulate(_p(pts1, C.c_double), _p(pts2, C.c_double), n, _p(P1, C.c_double),
                                _p(P2, C.c_double), _p(xyz, C.c_double), _p(err, C.c_double),
                                _p(valid, C.c_uint8))
    return xyz, err, valid.astype(bool)


def p3p(img_pts, world_pts, K4, max_num_trials=1000, confidence=99.0, max_reproj_error=1.0,
        seed=0, adaptive=True):
    """Returns dict(A 4x4, inliers bool n, status, n_inliers, best_trial, trials_run)."""
    ip = np.ascontiguousarray(img_pts, dtype=np.float64).reshape(-1, 2)
    wp = np.ascontiguousarray(world_pts, dtype=np.float64).reshape(-1, 3)
    K = np.ascontiguousarray(K4, dtype=np.float64).reshape(4)
    n = ip.shape[0]
    A = np.zeros((4, 4)); inl = np.zeros(max(n, 1), dtype=np.uint8)
    ni, bt, tr = C.c_int(0), C.c_int(0), C.c_int(0)
    o = _P3POpts(max_num_trials, confidence, max_reproj_error, seed, int(adaptive))
    st = lib().vo_oracle_p3p(_p(ip, C.c_double), _p(wp, C.c_double), n, _p(K, C.c_double),
                             C.byref(o), _p(A, C.c_double), _p(inl, C.c_uint8), C.byref(ni),
                             C.byref(bt), C.byref(tr))
    return dict(A=A, inliers=inl[:n].astype(bool), status=st, n_inliers=ni.value,
                best_trial=bt.value, trials_run=tr.value)


def p3p_solve(f3, X3):
    f = np.ascontiguousarray(f3, dtype=np.float64).reshape(9)
    X = np.ascontiguousarray(X3, dtype=np.float64).reshape(9)
    R = np.zeros((4, 9)); t = np.zeros((4, 3))
    lib().vo_oracle_p3p_solve.restype = C.c_int
    n = lib().vo_oracle_p3p_solve(_p(f, C.c_double), _p(X, C.c_double), _p(R, C.c_double),
                                  _p(t, C.c_double))
    return R[:n].reshape(n, 3, 3), t[:n]


def sample4(seed, trial, n):
    idx = np.zeros(4, dtype=np.uint32)
    lib().vo_oracle_sample4(C.c_uint64(seed), C.c_uint32(trial), C.c_uint32(n), _p(idx, C.c_uint32))
    return idx


def philox(c, k):
    out = np.zeros(4, dtype=np.uint32)
    lib().vo_oracle_philox4x32(*[C.c_uint32(int(v)) for v in c], C.c_uint32(int(k[0])),
                               C.c_uint32(int(k[1])), _p(out, C.c_uint32))
    return out
