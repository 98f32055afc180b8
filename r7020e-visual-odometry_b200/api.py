"""Host-side mirror of the toolbox calls on the reference's hot path, bound to libvo_b200.so.

Same names, argument meaning and error behaviour as the MATLAB calls in VO.m, so the VO loop
(vo.py) and the parity tests read like the reference:

    detectSIFTFeatures / extractFeatures(...,"Method","SIFT")   VO.m:79-84   -> vo_sift
    matchFeatures                                               VO.m:87,...  -> vo_match
    triangulate                                                 VO.m:114-115 -> vo_triangulate
    estworldpose                                                VO.m:123-127 -> vo_p3p

Indices are 0-based by default (NumPy); pass ``index_base=1`` for MATLAB numbering.  No function
here computes on the CPU: everything goes through the C ABI and fails loudly without the CUDA
library / a B200.
"""
import ctypes as C
import numpy as np

from . import _lib
from ._lib import VoError, check

KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"),
                     ("response", "f4"), ("octave", "i4")])

_default_ctx = None
_process_device = None   # libvo_b200 runs one process per GPU: the device of the first context


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Context:
    """Owns the device state of one GPU (vo_ctx): one stream and all device buffers.  Several
    contexts on the same device may run concurrently from different host threads; all contexts of
    a process live on one device (vo_ctx_create refuses a second device)."""

    def __init__(self, device=0):
        global _process_device
        self._h = C.c_void_p()
        check(_lib.lib().vo_ctx_create(int(device), C.byref(self._h)))
        self.device = device
        if _process_device is None:
            _process_device = int(device)

    def close(self):
        if self._h:
            _lib.lib().vo_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def sync(self):
        check(_lib.lib().vo_ctx_sync(self._h))

    def profile_enable(self, on=True):
        """Bracket every kernel stage with CUDA events (see vo_profile_enable); resets the totals."""
        check(_lib.lib().vo_profile_enable(self._h, 1 if on else 0))

    def profile(self):
        """{stage: dict(ms, launches, bytes, flops)} accumulated since profile_enable()."""
        L = _lib.lib()
        out = {}
        for i in range(L.vo_profile_count(self._h)):
            name = C.create_string_buffer(64)
            ms, by, fl = C.c_double(), C.c_double(), C.c_double()
            ln = C.c_longlong()
            check(L.vo_profile_get(self._h, i, name, 64, C.byref(ms), C.byref(ln), C.byref(by), C.byref(fl)))
            out[name.value.decode()] = dict(ms=ms.value, launches=ln.value, bytes=by.value, flops=fl.value)
        return out

    def kernel_launches(self):
        return int(_lib.lib().vo_kernel_launches(self._h))

    def use_frames_graph(self, enable=True):
        """vo_frames_use_graph: replay the frame loop's launch sequence as a CUDA graph (captured at the second call with
        the same batch shape and options)."""
        check(_lib.lib().vo_frames_use_graph(self._h, 1 if enable else 0))

    def frames_graph_state(self):
        """0 = off, 1 = on (nothing captured yet), 2 = a captured graph is being replayed, -1 = capture failed (plain launches)."""
        return int(_lib.lib().vo_frames_graph_state(self._h))

    @property
    def stream(self):
        """cudaStream_t (int) the host-pointer entry points launch on."""
        return _lib.lib().vo_ctx_stream(self._h)

    def match_stats(self):
        s = (C.c_int * 4)()
        check(_lib.lib().vo_match_stats(self._h, s))
        return dict(exact_integer_path=bool(s[0]), rowscan_rows=s[1], gemm_launches=s[2], k_extent=s[3])


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0 if _process_device is None else _process_device)
    return _default_ctx


class SIFTPoints:
    """Subset of MATLAB's SIFTPoints: Location (n x 2 single), Scale, Orientation, Metric,
    Octave, Layer; indexable like ``pts(idx)``."""

    def __init__(self, kps, index_base=0, features=None):
        self.kps = kps
        self.index_base = index_base
        self._features = features

    @property
    def Location(self):
        return np.stack([self.kps["x"], self.kps["y"]], axis=1)

    @property
    def Count(self):
        return len(self.kps)

    @property
    def Scale(self):
        return self.kps["size"] * 0.5

    @property
    def Orientation(self):
        return np.deg2rad(self.kps["angle"])

    @property
    def Metric(self):
        return self.kps["response"]

    @property
    def Octave(self):
        o = self.kps["octave"] & 255
        return np.where(o < 128, o, o - 256).astype(np.int32)

    @property
    def Layer(self):
        return ((self.kps["octave"] >> 8) & 255).astype(np.int32)

    def __len__(self):
        return len(self.kps)

    def __getitem__(self, idx):
        f = None if self._features is None else self._features[idx]
        return SIFTPoints(self.kps[idx], self.index_base, f)


def _sift_opts(ContrastThreshold, EdgeThreshold, NumLayersInOctave, Sigma, index_base):
    return _lib.SiftOpts(float(ContrastThreshold), float(EdgeThreshold), int(NumLayersInOctave),
                         float(Sigma), int(index_base))


def detectSIFTFeatures(I, ContrastThreshold=0.0, EdgeThreshold=10.0, NumLayersInOctave=3, Sigma=1.6,
                       index_base=0, capacity=16384, ctx=None):
    """detectSIFTFeatures(I) (VO.m:79-80).  Detection and description are one fused device call;
    the descriptors ride along in the returned SIFTPoints for extractFeatures.
    ContrastThreshold <= 0 selects the default 0.04/3 (MATLAB prints it as 0.0133)."""
    ctx = ctx or default_context()
    I = np.asarray(I)
    if I.dtype != np.uint8:
        if I.dtype.kind == "f":
            I = np.clip(np.rint(I * 255.0), 0, 255).astype(np.uint8)
        else:
            raise VoError("detectSIFTFeatures: image must be uint8 or floating point in [0,1]")
    if I.ndim != 2:
        raise VoError("detectSIFTFeatures: image must be 2-D grayscale")
    col_major = 0
    if I.flags.f_contiguous and not I.flags.c_contiguous:
        col_major, ld = 1, I.shape[0]
    else:
        I = np.ascontiguousarray(I)
        ld = I.shape[1]
    kps = np.zeros(capacity, dtype=KP_DTYPE)
    desc = np.zeros((capacity, 128), dtype=np.float32)
    n = C.c_int(0)
    o = _sift_opts(ContrastThreshold, EdgeThreshold, NumLayersInOctave, Sigma, index_base)
    rc = _lib.lib().vo_sift(ctx.handle, _p(I, C.c_uint8), I.shape[0], I.shape[1], ld, col_major,
                            C.byref(o), capacity, kps.ctypes.data_as(C.POINTER(_lib.Keypoint)),
                            _p(desc, C.c_float), C.byref(n))
    check(rc)
    m = n.value
    return SIFTPoints(kps[:m].copy(), index_base, desc[:m].copy())


def extractFeatures(I, points, Method="SIFT"):
    """[features, validPoints] = extractFeatures(I, points, "Method", "SIFT") (VO.m:83-84)."""
    if Method != "SIFT":
        raise VoError("extractFeatures: only Method='SIFT' is on the reference's path")
    if points._features is None:
        raise VoError("extractFeatures: points must come from detectSIFTFeatures (fused call)")
    return points._features, points


def sift_batch(imgs, ContrastThreshold=0.0, EdgeThreshold=10.0, NumLayersInOctave=3, Sigma=1.6,
               index_base=0, capacity=8192, ctx=None):
    """Batched detect+describe over imgs[n, rows, cols] uint8.  Returns a list of SIFTPoints."""
    ctx = ctx or default_context()
    imgs = np.ascontiguousarray(imgs, dtype=np.uint8)
    n_img, rows, cols = imgs.shape
    kps = np.zeros((n_img, capacity), dtype=KP_DTYPE)
    desc = np.zeros((n_img, capacity, 128), dtype=np.float32)
    n = np.zeros(n_img, dtype=np.int32)
    o = _sift_opts(ContrastThreshold, EdgeThreshold, NumLayersInOctave, Sigma, index_base)
    check(_lib.lib().vo_sift_batch(ctx.handle, _p(imgs, C.c_uint8), n_img, rows, cols, C.byref(o),
                                   capacity, kps.ctypes.data_as(C.POINTER(_lib.Keypoint)),
                                   _p(desc, C.c_float), _p(n, C.c_int)))
    return [SIFTPoints(kps[i, :n[i]].copy(), index_base, desc[i, :n[i]].copy()) for i in range(n_img)]


def matchFeatures(features1, features2, MatchThreshold=1.0, MaxRatio=0.6, Unique=False,
                  index_base=0, return_metric=False, ctx=None):
    """indexPairs = matchFeatures(features1, features2) (VO.m:87, 283, 293, 311, 323).
    Exhaustive, SSD on unit-normalised rows; returns P x 2 uint32 (ascending in column 0)."""
    ctx = ctx or default_context()
    f1 = np.asarray(features1, dtype=np.float32)
    f2 = np.asarray(features2, dtype=np.float32)
    if f1.ndim != 2 or f2.ndim != 2 or f1.shape[1] != f2.shape[1]:
        raise VoError("matchFeatures: features must be N1 x D and N2 x D with equal D")
    col_major = 0
    if (f1.flags.f_contiguous and f2.flags.f_contiguous and not f1.flags.c_contiguous
            and not f2.flags.c_contiguous):
        col_major = 1
    else:
        f1 = np.ascontiguousarray(f1)
        f2 = np.ascontiguousarray(f2)
    n1, n2, dim = f1.shape[0], f2.shape[0], f1.shape[1]
    i1 = np.zeros(max(n1, 1), dtype=np.uint32)
    i2 = np.zeros(max(n1, 1), dtype=np.uint32)
    m = np.zeros(max(n1, 1), dtype=np.float32)
    p = C.c_int(0)
    o = _lib.MatchOpts(float(MatchThreshold), float(MaxRatio), int(bool(Unique)), int(index_base))
    check(_lib.lib().vo_match(ctx.handle, _p(f1, C.c_float), n1, _p(f2, C.c_float), n2, dim, col_major,
                              C.byref(o), _p(i1, C.c_uint32), _p(i2, C.c_uint32), _p(m, C.c_float),
                              C.byref(p)))
    pairs = np.stack([i1[:p.value], i2[:p.value]], axis=1)
    if return_metric:
        return pairs, m[:p.value].copy()
    return pairs


def match_top2(features1, features2, ctx=None):
    """Per-row nearest column, its SSD and the second-nearest SSD (no thresholding)."""
    ctx = ctx or default_context()
    f1 = np.ascontiguousarray(features1, dtype=np.float32)
    f2 = np.ascontiguousarray(features2, dtype=np.float32)
    n1, n2, dim = f1.shape[0], f2.shape[0], f1.shape[1]
    j1 = np.zeros(max(n1, 1), dtype=np.uint32)
    s1 = np.zeros(max(n1, 1), dtype=np.float32)
    s2 = np.zeros(max(n1, 1), dtype=np.float32)
    check(_lib.lib().vo_match_top2(ctx.handle, _p(f1, C.c_float), n1, _p(f2, C.c_float), n2, dim, 0,
                                   _p(j1, C.c_uint32), _p(s1, C.c_float), _p(s2, C.c_float)))
    return j1[:n1], s1[:n1], s2[:n1]


def match_debug_gemm(features1, features2, ctx=None):
    ctx = ctx or default_context()
    f1 = np.ascontiguousarray(features1, dtype=np.float32)
    f2 = np.ascontiguousarray(features2, dtype=np.float32)
    c = np.zeros((f1.shape[0], f2.shape[0]), dtype=np.float32)
    check(_lib.lib().vo_match_debug_gemm(ctx.handle, _p(f1, C.c_float), f1.shape[0], _p(f2, C.c_float),
                                         f2.shape[0], f1.shape[1], _p(c, C.c_float)))
    return c


def triangulate(matchedPoints1, matchedPoints2, camProjection1, camProjection2, full=False, ctx=None):
    """worldPoints = triangulate(pts1, pts2, P1, P2) with 3x4 projection matrices (VO.m:114-115).
    Accepts 1x2 or N x 2 single/double; output has the class of the points."""
    ctx = ctx or default_context()
    p1 = np.asarray(matchedPoints1)
    p2 = np.asarray(matchedPoints2)
    is_double = 1 if p1.dtype == np.float64 else 0
    dt = np.float64 if is_double else np.float32
    p1 = np.ascontiguousarray(p1, dtype=dt).reshape(-1, 2)
    p2 = np.ascontiguousarray(p2, dtype=dt).reshape(-1, 2)
    if p1.shape != p2.shape:
        raise VoError("triangulate: point sets must have the same size")
    P1 = np.ascontiguousarray(camProjection1, dtype=np.float64)
    P2 = np.ascontiguousarray(camProjection2, dtype=np.float64)
    if P1.shape == (4, 3):
        P1, P2 = np.ascontiguousarray(P1.T), np.ascontiguousarray(P2.T)   # legacy camMatrix
    if P1.shape != (3, 4) or P2.shape != (3, 4):
        raise VoError("triangulate: projection matrices must be 3x4")
    n = p1.shape[0]
    xyz = np.zeros((n, 3), dtype=dt)
    err = np.zeros(n, dtype=dt)
    valid = np.zeros(max(n, 1), dtype=np.uint8)
    check(_lib.lib().vo_triangulate(ctx.handle, p1.ctypes.data_as(C.c_void_p), p2.ctypes.data_as(C.c_void_p),
                                    n, is_double, 0, _p(P1, C.c_double), _p(P2, C.c_double),
                                    xyz.ctypes.data_as(C.c_void_p), err.ctypes.data_as(C.c_void_p),
                                    _p(valid, C.c_uint8)))
    if full:
        return xyz, err, valid[:n].astype(bool)
    return xyz


class rigidtform3d:
    """4x4 premultiply rigid transform, like MATLAB's rigidtform3d (VO.m:58, 130)."""

    def __init__(self, A=None):
        self.A = np.eye(4) if A is None else np.asarray(A, dtype=np.float64).reshape(4, 4)

    @property
    def R(self):
        return self.A[:3, :3]

    @property
    def Translation(self):
        return self.A[:3, 3]

    def transformPointsForward(self, pts):
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
        return pts @ self.A[:3, :3].T + self.A[:3, 3]


def estworldpose(imagePoints, worldPoints, intrinsics, MaxNumTrials=1000, Confidence=99.0,
                 MaxReprojectionError=1.0, Seed=0, Adaptive=True, full=False, ctx=None):
    """worldPose = estworldpose(imagePoints, worldPoints, intrinsics) (VO.m:123-127).
    intrinsics: (fx, fy, cx, cy) or a 3x3 K.  Raises like MATLAB when no status is requested
    (full=False) and the estimate fails."""
    ctx = ctx or default_context()
    ip = np.ascontiguousarray(imagePoints, dtype=np.float64).reshape(-1, 2)
    wp = np.ascontiguousarray(worldPoints, dtype=np.float64).reshape(-1, 3)
    if ip.shape[0] != wp.shape[0]:
        raise VoError("estworldpose: imagePoints and worldPoints must have the same count")
    K = np.asarray(intrinsics, dtype=np.float64)
    if K.shape == (3, 3):
        K = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]])
    K = np.ascontiguousarray(K.reshape(4))
    n = ip.shape[0]
    A = np.zeros((4, 4))
    inl = np.zeros(max(n, 1), dtype=np.uint8)
    status = C.c_int(0)
    info = (C.c_int * 3)()
    o = _lib.P3POpts(int(MaxNumTrials), float(Confidence), float(MaxReprojectionError), int(Seed),
                     1 if Adaptive else 0)
    check(_lib.lib().vo_p3p(ctx.handle, _p(ip, C.c_double), _p(wp, C.c_double), n, 0, _p(K, C.c_double),
                            C.byref(o), _p(A, C.c_double), _p(inl, C.c_uint8), C.byref(status), info))
    if full:
        return dict(A=A, inliers=inl[:n].astype(bool), status=status.value, n_inliers=info[0],
                    best_trial=info[1], trials_run=info[2])
    if status.value == 1:
        raise VoError("estworldpose: not enough points (need at least 4)")
    if status.value == 2:
        raise VoError("estworldpose: not enough inliers")
    return rigidtform3d(A)
