"""The MEX gateways driven by a stand-in MATLAB host (vo_b200/mexhost.py over csrc/mex/mexshim.cpp):
column-major inputs, 1-based uint32 indexPairs, class-preserving triangulate, estworldpose-style erroring,
H x W x N stacks for vo_sift_mex and the batched loop gateway vo_frames_mex."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu
from vo_b200.mexhost import Host, MexOps, frames as mex_frames


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


def test_sift_gateway_stack(host):
    """vo_sift_mex(cat(3, lf, rf)): both images of VO.m:79-84 in one call; rows concatenated, count per image."""
    from vo_b200 import synth
    l, r = synth.shift_stream(1, seed=3, h=120, w=200)
    o = host.call("vo_sift_mex", 8, np.stack([l[0], r[0]], axis=2))
    cnt = o[7][:, 0]
    a = host.call("vo_sift_mex", 7, l[0]); b = host.call("vo_sift_mex", 7, r[0])
    assert cnt.tolist() == [len(a[0]), len(b[0])] and len(a[0]) > 50
    for k in range(7):
        assert np.array_equal(o[k], np.concatenate([a[k], b[k]], axis=0))
    with pytest.raises(RuntimeError, match="vo:sift:class"):
        host.call("vo_sift_mex", 1, np.zeros((20, 20), np.int32))


def test_frames_gateway_equals_vo_frames(host, ctx):
    """[relA, status, counts] = vo_frames_mex(L, R, P1, P2, ...) on H x W x N stacks is bit-identical to vo_frames
    on the same frames (and to the column-major form of the C ABI), relA ready for rigidtform3d (column-major)."""
    from vo_b200 import synth, vo
    left, right = synth.shift_stream(5, seed=6, h=188, w=620)
    want = vo.run_frames(left, right, synth.KITTI_P0, synth.KITTI_P1, seed=7, first_frame=3, ctx=ctx)
    got = mex_frames(host, left, right, synth.KITTI_P0, synth.KITTI_P1, seed=7, first_frame=3)
    assert (want[1][1:] == 0).all() and (want[2][1:, 6] > 30).all()
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    lt = np.ascontiguousarray(np.transpose(left, (0, 2, 1))); rt = np.ascontiguousarray(np.transpose(right, (0, 2, 1)))
    cm = vo.run_frames(lt, rt, synth.KITTI_P0, synth.KITTI_P1, seed=7, first_frame=3, ctx=ctx, col_major=True)
    for a, b in zip(cm, want):
        assert np.array_equal(a, b)
    # 4x3 legacy camMatrix form and argument errors
    relA, = host.call("vo_frames_mex", 1, np.transpose(left, (1, 2, 0)), np.transpose(right, (1, 2, 0)),
                      synth.KITTI_P0.T.copy(), synth.KITTI_P1.T.copy(), "Seed", np.uint64(7), "FirstFrame", np.float64(3))
    assert np.array_equal(np.transpose(relA, (2, 0, 1)), want[0])
    with pytest.raises(RuntimeError, match="vo:frames:size"):
        host.call("vo_frames_mex", 1, np.transpose(left, (1, 2, 0)), np.transpose(right[:4], (1, 2, 0)), synth.KITTI_P0, synth.KITTI_P1)
    with pytest.raises(RuntimeError, match="vo:frames:class"):
        host.call("vo_frames_mex", 1, left[0].astype(np.float32), right[0].astype(np.float32), synth.KITTI_P0, synth.KITTI_P1)


def test_six_call_loop_through_gateways_equals_batched(host, ctx):
    """The literal drop-in: VO.m's loop with every toolbox call answered by a gateway (MexOps) gives the same
    relative poses as the batched loop."""
    from vo_b200 import synth, vo
    left, right = synth.shift_stream(4, seed=8, h=188, w=620)
    g = vo.VisualOdometry(synth.KITTI_P0, synth.KITTI_P1, MexOps(host, seed=5))
    rels = [g.step(left[i], right[i]) for i in range(4)]
    want = vo.run_frames(left, right, synth.KITTI_P0, synth.KITTI_P1, seed=5, ctx=ctx)
    for i in range(1, 4):
        assert np.array_equal(rels[i], want[0][i])
