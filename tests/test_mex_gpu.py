"""The four MEX gateways driven by a fake MATLAB host (csrc/mex/mexshim.cpp): column-major inputs,
1-based uint32 indexPairs, class-preserving triangulate, estworldpose-style erroring."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu
D = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "r7020e-visual-odometry_b200", "csrc", "mex")
CLS = {np.dtype("float64"): 6, np.dtype("float32"): 7, np.dtype("uint8"): 9, np.dtype("int32"): 12, np.dtype("uint32"): 13}
NP = {6: np.float64, 7: np.float32, 9: np.uint8, 12: np.int32, 13: np.uint32, 3: np.uint8}


class Host:
    def __init__(self):
        self.shim = C.CDLL(os.path.join(D, "libmexshim.so"), mode=C.RTLD_GLOBAL)
        s = self.shim
        s.shim_from_buffer.restype = C.c_void_p
        s.shim_from_buffer.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_void_p]
        s.mxCreateString.restype = C.c_void_p
        s.mxGetData.restype = C.c_void_p
        s.mxGetData.argtypes = [C.c_void_p]
        for f in ("mxGetM", "mxGetN"):
            getattr(s, f).restype = C.c_size_t
            getattr(s, f).argtypes = [C.c_void_p]
        s.mxGetClassID.argtypes = [C.c_void_p]
        s.shim_last_error_id.restype = C.c_char_p
        s.shim_last_error_msg.restype = C.c_char_p
        s.shim_call.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        self.gates = {}

    def mx(self, a):
        if isinstance(a, str):
            return self.shim.mxCreateString(a.encode())
        a = np.asfortranarray(np.atleast_2d(a))                   # MATLAB arrays are column-major
        return self.shim.shim_from_buffer(CLS[a.dtype], a.shape[0], a.shape[1], a.ctypes.data_as(C.c_void_p))

    def call(self, gate, nlhs, *args):
        if gate not in self.gates:
            self.gates[gate] = C.CDLL(os.path.join(D, gate + ".mexa64"))
        fn = C.cast(self.gates[gate].mexFunction, C.c_void_p)
        prhs = (C.c_void_p * len(args))(*[self.mx(a) for a in args])
        plhs = (C.c_void_p * max(nlhs, 1))()
        if self.shim.shim_call(fn, nlhs, plhs, len(args), prhs):
            raise RuntimeError(self.shim.shim_last_error_id().decode() + ": " + self.shim.shim_last_error_msg().decode())
        outs = []
        for k in range(max(nlhs, 1)):
            m, n, cls = self.shim.mxGetM(plhs[k]), self.shim.mxGetN(plhs[k]), self.shim.mxGetClassID(plhs[k])
            buf = (C.c_char * (m * n * np.dtype(NP[cls]).itemsize)).from_address(self.shim.mxGetData(plhs[k])) if m * n else b""
            outs.append(np.frombuffer(bytes(buf), dtype=NP[cls]).reshape((n, m)).T.copy())
        return outs


@pytest.fixture(scope="module")
def host():
    return Host()


def test_match_gateway(host):
    from conftest import correlated_pair
    f1, f2 = correlated_pair(500, 600, seed=8)
    pairs, metric = host.call("vo_match_mex", 2, f1, f2)
    op, om = oracle.match(f1, f2, index_base=1)
    assert pairs.dtype == np.uint32 and pairs.shape == op.shape
    assert np.array_equal(pairs, op) and np.array_equal(metric[:, 0].view(np.uint32), om.view(np.uint32))
    p2, = host.call("vo_match_mex", 1, f1, f2, "MaxRatio", np.float64(0.9), "MatchThreshold", np.float64(5.0))
    assert np.array_equal(p2, oracle.match(f1, f2, max_ratio=0.9, match_threshold=5.0, index_base=1)[0])
    with pytest.raises(RuntimeError, match="vo:match:class"):
        host.call("vo_match_mex", 1, f1.astype(np.float64), f2.astype(np.float64))
    e, = host.call("vo_match_mex", 1, np.zeros((0, 128), np.float32), f2)
    assert e.shape == (0, 2)


def test_sift_gateway(host):
    from vo_b200 import synth
    img = synth.texture(150, 210, seed=2)
    desc, loc, scale, orient, metric, octave, layer = host.call("vo_sift_mex", 7, img)
    okp, odesc = oracle.sift(img)
    assert desc.shape == (len(okp), 128) and desc.dtype == np.float32
    assert np.array_equal(desc, odesc)
    assert np.array_equal(loc, np.stack([okp["x"], okp["y"]], 1) + np.float32(1.0))       # 1-based Location
    assert np.allclose(scale[:, 0], okp["size"] * 0.5) and octave.dtype == np.int32
    assert set(np.unique(layer)) <= {1, 2, 3}


def test_triangulate_and_p3p_gateways(host):
    from vo_b200 import synth
    rng = np.random.default_rng(0)
    X = np.c_[rng.uniform(-10, 10, 50), rng.uniform(-2, 2, 50), rng.uniform(5, 50, 50)]
    pr = lambda P: (lambda h: h[:, :2] / h[:, 2:])(np.c_[X, np.ones(50)] @ P.T)
    x1, x2 = pr(synth.KITTI_P0), pr(synth.KITTI_P1)
    xyz, err, valid = host.call("vo_triangulate_mex", 3, x1, x2, synth.KITTI_P0, synth.KITTI_P1)
    assert xyz.dtype == np.float64 and np.allclose(xyz, X, rtol=1e-9) and valid.all()
    one, = host.call("vo_triangulate_mex", 1, x1[:1].astype(np.float32), x2[:1].astype(np.float32), synth.KITTI_P0, synth.KITTI_P1)
    assert one.dtype == np.float32 and one.shape == (1, 3)               # the reference's 1x2 single call (VO.m:114)
    legacy, = host.call("vo_triangulate_mex", 1, x1, x2, synth.KITTI_P0.T.copy(), synth.KITTI_P1.T.copy())
    assert np.allclose(legacy, X, rtol=1e-9)                              # 4x3 camMatrix form
    A, inl, status = host.call("vo_p3p_mex", 3, x1, X, synth.KITTI_K4, "Seed", np.float64(3))
    o = oracle.p3p(x1, X, synth.KITTI_K4, seed=3)
    assert status[0, 0] == 0 and np.allclose(A, o["A"], atol=1e-9) and np.array_equal(inl[:, 0].astype(bool), o["inliers"])
    with pytest.raises(RuntimeError, match="vo:p3p:notEnoughPts"):
        host.call("vo_p3p_mex", 1, x1[:3], X[:3], synth.KITTI_K4)         # errors like estworldpose
    _, _, st = host.call("vo_p3p_mex", 3, x1[:3], X[:3], synth.KITTI_K4)
    assert st[0, 0] == 1
