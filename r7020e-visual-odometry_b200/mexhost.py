"""A stand-in MATLAB host for the MEX gateways (csrc/mex/*.mexa64).

MATLAB is not installed here, so the gateways are compiled against a ``mex.h`` stand-in and
driven through ``libmexshim.so`` (csrc/mex/mexshim.cpp), which implements the mx*/mex* calls they use.  This
module is that driver: it builds column-major ``mxArray`` inputs from NumPy arrays, calls a gateway's
``mexFunction`` with MATLAB's error semantics (``mexErrMsgIdAndTxt`` unwinds the call) and converts the outputs
back.  ``MexOps`` plugs the gateways into the line-by-line mirror of VO.m (vo.VisualOdometry), which is exactly
what a MATLAB session does after swapping the six toolbox calls; ``frames`` is the batched
gateway ``vo_frames_mex``.  Used by tests/test_mex_gpu.py and by bench.py's ``e2e_dropin`` leg.
"""
import ctypes as C
import os
import time

import numpy as np

D = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "mex")
CLS = {np.dtype("float64"): 6, np.dtype("float32"): 7, np.dtype("uint8"): 9, np.dtype("int32"): 12, np.dtype("uint32"): 13,
       np.dtype("int64"): 14, np.dtype("uint64"): 15}
NP = {6: np.float64, 7: np.float32, 9: np.uint8, 12: np.int32, 13: np.uint32, 14: np.int64, 15: np.uint64, 3: np.uint8}


class MexError(RuntimeError):
    pass


class Host:
    def __init__(self):
        so = os.path.join(D, "libmexshim.so")
        if not os.path.exists(so):
            raise MexError(f"{so} is missing: build it with `python __graft_entry__.py`")
        self.shim = C.CDLL(so, mode=C.RTLD_GLOBAL)
        s = self.shim
        s.shim_from_buffer.restype = C.c_void_p
        s.shim_from_buffer.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_void_p]
        s.shim_from_buffer_nd.restype = C.c_void_p
        s.shim_from_buffer_nd.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
        s.mxCreateString.restype = C.c_void_p
        s.mxGetData.restype = C.c_void_p
        s.mxGetData.argtypes = [C.c_void_p]
        s.mxDestroyArray.argtypes = [C.c_void_p]
        for f in ("mxGetM", "mxGetN", "mxGetNumberOfDimensions"):
            getattr(s, f).restype = C.c_size_t
            getattr(s, f).argtypes = [C.c_void_p]
        s.mxGetDimensions.restype = C.POINTER(C.c_size_t)
        s.mxGetDimensions.argtypes = [C.c_void_p]
        s.mxGetClassID.argtypes = [C.c_void_p]
        s.shim_last_error_id.restype = C.c_char_p
        s.shim_last_error_msg.restype = C.c_char_p
        s.shim_call.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        self.gates = {}

    def mx(self, a):
        """NumPy array / str -> mxArray* (column-major copy, like a MATLAB variable)."""
        if isinstance(a, str):
            return self.shim.mxCreateString(a.encode())
        a = np.asarray(a)
        if a.ndim <= 2:
            a = np.asfortranarray(np.atleast_2d(a))
            return self.shim.shim_from_buffer(CLS[a.dtype], a.shape[0], a.shape[1], a.ctypes.data_as(C.c_void_p))
        a = np.asfortranarray(a)
        dims = (C.c_size_t * a.ndim)(*a.shape)
        return self.shim.shim_from_buffer_nd(CLS[a.dtype], a.ndim, dims, a.ctypes.data_as(C.c_void_p))

    def free(self, h):
        self.shim.mxDestroyArray(h)

    def call(self, gate, nlhs, *args, keep_inputs=False):
        """outs = gate(args...).  Arguments that are ints are taken as mxArray* made by ``mx`` (kept alive by
        the caller, as MATLAB keeps a workspace variable)."""
        if gate not in self.gates:
            self.gates[gate] = C.CDLL(os.path.join(D, gate + ".mexa64"))
        fn = C.cast(self.gates[gate].mexFunction, C.c_void_p)
        own = [not isinstance(a, int) for a in args]
        hs = [a if isinstance(a, int) else self.mx(a) for a in args]
        prhs = (C.c_void_p * len(args))(*hs)
        plhs = (C.c_void_p * max(nlhs, 1))()
        rc = self.shim.shim_call(fn, nlhs, plhs, len(args), prhs)
        for h, o in zip(hs, own):
            if o:
                self.free(h)
        if rc:
            raise MexError(self.shim.shim_last_error_id().decode() + ": " + self.shim.shim_last_error_msg().decode())
        outs = []
        for k in range(max(nlhs, 1)):
            nd = self.shim.mxGetNumberOfDimensions(plhs[k])
            dp = self.shim.mxGetDimensions(plhs[k])
            shape = tuple(int(dp[i]) for i in range(nd))
            cls = self.shim.mxGetClassID(plhs[k])
            cnt = int(np.prod(shape))
            if cnt:
                buf = (C.c_char * (cnt * np.dtype(NP[cls]).itemsize)).from_address(self.shim.mxGetData(plhs[k]))
                outs.append(np.frombuffer(bytes(buf), dtype=NP[cls]).reshape(shape, order="F").copy())
            else:
                outs.append(np.zeros(shape, dtype=NP[cls]))
            self.free(plhs[k])
        return outs


class MexOps:
    """The toolbox calls of VO.m answered by the MEX gateways (operator set of vo.VisualOdometry)."""

    def __init__(self, host=None, seed=0):
        self.host = host or Host()
        self.seed = seed

    def detect_and_extract(self, img):
        desc, loc = self.host.call("vo_sift_mex", 2, img)
        return desc, loc

    def detect_and_extract_pair(self, lf, rf):
        """[desc, loc, ~, ~, ~, ~, ~, count] = vo_sift_mex(cat(3, lf, rf)): both images of VO.m:79-84 in one call."""
        o = self.host.call("vo_sift_mex", 8, np.stack([lf, rf], axis=2))
        desc, loc, cnt = o[0], o[1], o[7][:, 0]
        return (desc[:cnt[0]], loc[:cnt[0]]), (desc[cnt[0]:], loc[cnt[0]:])

    def matchFeatures(self, f1, f2):
        pairs, = self.host.call("vo_match_mex", 1, f1, f2)
        return pairs.astype(np.int64) - 1                        # MATLAB 1-based -> NumPy indexing of the mirror

    def triangulate(self, p1, p2, P1, P2):
        xyz, = self.host.call("vo_triangulate_mex", 1, np.asarray(p1, np.float64), np.asarray(p2, np.float64),
                              np.asarray(P1, np.float64), np.asarray(P2, np.float64))
        return xyz

    def estworldpose(self, image_points, world_points, K4, frame_index):
        seed = (self.seed + frame_index * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        A, inl, st = self.host.call("vo_p3p_mex", 3, np.asarray(image_points, np.float64), np.asarray(world_points, np.float64),
                                    np.asarray(K4, np.float64), "Seed", np.uint64(seed))
        return dict(A=A, inliers=inl[:, 0].astype(bool), status=int(st[0, 0]), n_inliers=int(inl.sum()))


def frames(host, left, right, P1, P2, seed=0, first_frame=0, handles=None):
    """[relA, status, counts] = vo_frames_mex(L, R, P1, P2, 'Seed', seed, 'FirstFrame', first_frame).
    left/right: [n, rows, cols] uint8 (NumPy); the MATLAB-side arrays are H x W x N.  ``handles`` = (L, R)
    mxArray* made beforehand with ``stack_handles`` (a MATLAB session already holds its image stack)."""
    if handles is None:
        L, R = np.transpose(left, (1, 2, 0)), np.transpose(right, (1, 2, 0))
    else:
        L, R = handles
    relA, status, counts = host.call("vo_frames_mex", 3, L, R, np.asarray(P1, np.float64), np.asarray(P2, np.float64),
                                     "Seed", np.uint64(seed), "FirstFrame", np.float64(first_frame))
    return np.ascontiguousarray(np.transpose(relA, (2, 0, 1))), status[:, 0], np.ascontiguousarray(counts.T)


def stack_handles(host, left, right):
    return host.mx(np.transpose(left, (1, 2, 0))), host.mx(np.transpose(right, (1, 2, 0)))


def bench_dropin(sleft, sright, n_percall=24, n_batched=129, batch=32):
    """bench.py's e2e_dropin leg: wall-clock frames/s of (a) the six-call loop of VO.m through the per-call
    gateways and (b) the batched gateway, both through the stand-in host; (b) is also compared bit for bit with
    vo_frames called directly."""
    from . import synth, vo
    host = Host()
    P0, P1 = synth.KITTI_P0, synth.KITTI_P1
    out = {}
    # (a) literal drop-in: one gateway call per toolbox call
    g = vo.VisualOdometry(P0, P1, MexOps(host, seed=1))
    for i in range(3):
        g.step(sleft[i], sright[i])                          # warm-up (plans, buffers)
    g = vo.VisualOdometry(P0, P1, MexOps(host, seed=1))
    t0 = time.perf_counter()
    for i in range(n_percall):
        g.step(sleft[i], sright[i])
    dt = time.perf_counter() - t0
    out["percall"] = dict(value=(n_percall - 1) / dt, unit="frames/s", frames=n_percall - 1, ms_per_frame=1e3 * dt / (n_percall - 1),
                          calls_per_frame=9, note="vo_sift_mex x2, vo_match_mex x5, vo_triangulate_mex, vo_p3p_mex per frame; "
                                                  "MATLAB-side indexing done in NumPy; includes the stand-in host's own copies of every "
                                                  "input into a fresh mxArray and of every output back (0.7 ms per 3 k x 128 descriptor "
                                                  "matrix: tools/mex_percall_probe.py), which a MATLAB session does not make")
    # (b) batched gateway: stacks of batch+1 frames (one-frame halo), the stacks already MATLAB arrays
    nb = (min(n_batched, len(sleft)) - 1) // batch
    hs = [stack_handles(host, sleft[b * batch: b * batch + batch + 1], sright[b * batch: b * batch + batch + 1]) for b in range(nb)]
    frames(host, None, None, P0, P1, seed=1, first_frame=0, handles=hs[0])       # warm-up
    t0 = time.perf_counter()
    res = [frames(host, None, None, P0, P1, seed=1, first_frame=b * batch, handles=hs[b]) for b in range(nb)]
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    hs2 = stack_handles(host, sleft[:batch + 1], sright[:batch + 1])
    t_stack = time.perf_counter() - t1
    host.free(hs2[0]); host.free(hs2[1])
    for h in hs:
        host.free(h[0]); host.free(h[1])
    same = True
    for b in range(nb):
        ref = vo.run_frames(sleft[b * batch: b * batch + batch + 1], sright[b * batch: b * batch + batch + 1], P0, P1, seed=1,
                            first_frame=b * batch)
        same = same and np.array_equal(res[b][0], ref[0]) and np.array_equal(res[b][1], ref[1]) and np.array_equal(res[b][2], ref[2])
    out["batched"] = dict(value=nb * batch / dt, unit="frames/s", frames=nb * batch, frames_per_call=batch, ms_per_call=1e3 * dt / nb,
                          ms_to_build_one_stack_pair=1e3 * t_stack,
                          note="vo_frames_mex(L, R, P1, P2): H x W x N uint8 stacks held by the host (pageable memory), one call "
                               "per batch, one batch at a time")
    out["value"] = out["batched"]["value"]
    out["unit"] = "frames/s"
    out["equals_vo_frames"] = bool(same)
    return out
