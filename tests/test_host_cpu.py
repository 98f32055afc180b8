"""CPU suite: host-side logic -- sharding arithmetic (gloo, world_size 2), the KITTI evaluator, the
synthetic generator, and the VO loop mirror driven by the oracle."""
import os
import sys

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_frame_and_row_chunks():
    from vo_b200 import shard
    for n in (1, 2, 7, 33, 4541):
        for w in (1, 2, 3, 4, 8):
            ch = shard.frame_chunks(n, w)
            assert ch[0][0] == 1 and ch[-1][1] == max(n, 1)
            assert all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
            sizes = [h - l for l, h in ch]
            assert max(sizes) - min(sizes) <= 1
    assert shard.row_chunks(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vo_b200 import shard
    from oracle import oracle
    # (1) frame sharding: fake per-frame result = f(frame index) so stitching errors are visible
    n_frames = 11

    def load(lo, hi):
        return np.arange(lo, hi), np.arange(lo, hi)

    def run(left, right, first):
        n = len(left)
        rel = np.tile(np.eye(4), (n, 1, 1))
        rel[:, 0, 3] = left * 10.0 + 1.0
        assert left[0] == first
        return rel, np.full(n, 0), None
    rel, status = shard.run_sequence_sharded(load, n_frames, run, rank, world, dist, batch=3)
    ok1 = np.array_equal(rel[1:, 0, 3], np.arange(1, n_frames) * 10.0 + 1.0) and np.array_equal(rel[0], np.eye(4))
    # (2) row-sharded best-2 + all-gather equals the unsharded result bit for bit
    rng = np.random.default_rng(0)
    q_ = np.rint(np.abs(rng.normal(0, 40, (37, 128)))).astype(np.float32)
    l_ = np.rint(np.abs(rng.normal(0, 40, (50, 128)))).astype(np.float32)
    j1, s1, s2 = shard.match_top2_row_sharded(q_, l_, oracle.match_top2, rank, world, dist)
    r1, r2, r3 = oracle.match_top2(q_, l_)
    ok2 = np.array_equal(j1, r1) and np.array_equal(s1.view(np.uint32), r2.view(np.uint32)) and np.array_equal(s2.view(np.uint32), r3.view(np.uint32))
    q.put((rank, bool(ok1), bool(ok2)))
    dist.destroy_process_group()


def test_sharding_with_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(60)
    assert sorted(res) == [(0, True, True), (1, True, True)]


def test_kitti_eval_on_reference_ground_truth():
    from vo_b200 import kitti_eval
    k = np.load(os.path.join(G, "kitti00_reference_data.npz"))
    gt = np.tile(np.eye(4), (len(k["poses"]), 1, 1)); gt[:, :3, :] = k["poses"]      # kitti/poses/00.txt
    t, r, n = kitti_eval.kitti_errors(gt, gt)
    assert n > 0 and t < 1e-12 and r < 1e-7
    assert kitti_eval.xz_error(gt, gt).max() == 0
    # a 1 % scale error on every step -> t_err = 1 %
    rel = [np.linalg.inv(gt[i - 1]) @ gt[i] for i in range(1, len(gt))]
    for a in rel:
        a[:3, 3] *= 1.01
    est = [np.eye(4)]
    for a in rel:
        est.append(est[-1] @ a)
    t, r, n = kitti_eval.kitti_errors(np.array(est), gt)
    assert 0.006 < t <= 0.0101 and r < 1e-6       # chord <= path length on curved segments
    # reference's off-by-one (PlotOnMap.m compares pose k with truth row k): error is one step length
    e = kitti_eval.xz_error(np.array(est)[1:], gt)
    assert 0.3 < np.median(e[50:]) < 2.0


def test_synth_streams():
    from vo_b200 import synth
    l, r = synth.shift_stream(3, seed=1, h=64, w=200, disparity=12, shift=3)
    assert l.shape == (3, 64, 200) and l.dtype == np.uint8
    assert np.array_equal(l[0][:, 12:], r[0][:, :-12])              # x_left - x_right = +12
    assert np.array_equal(l[1][:, :-3], l[0][:, 3:])                # +3 px per frame
    assert np.array_equal(synth.texture(32, 48, 5), synth.texture(32, 48, 5))
    a, z = synth.shift_stream_truth()
    assert abs(z - 32.18) < 0.01 and abs(a[0, 3] - 0.1343) < 1e-3


def test_vo_loop_with_oracle_recovers_motion():
    from vo_b200 import vo, synth
    from oracle_ops import OracleOps
    l, r = synth.shift_stream(3, seed=11, h=188, w=620)
    o = vo.VisualOdometry(synth.KITTI_P0, synth.KITTI_P1, OracleOps(seed=1))
    for i in range(3):
        rel = o.step(l[i], r[i])
    truth, _ = synth.shift_stream_truth()
    assert np.allclose(rel[:3, 3], truth[:3, 3], atol=0.03) and np.allclose(rel[:3, :3], np.eye(3), atol=5e-3)
    assert len(o.all_poses) == 2 and np.allclose(o.all_poses[-1], o.all_poses[0] @ rel)
    assert o.log[-1]["k4"] <= o.log[-1]["k3"] <= min(o.log[-1]["k1"], o.log[-1]["k2"])


def test_landmark_map_mirrors_the_reference_loop():
    """VO.m:147-161 + CreateLandmarksFromFeatures.m, restated literally (scalar loops, MATLAB growth
    semantics) and compared with the batched mirror in vo.py (oracle triangulation on both sides)."""
    from vo_b200 import vo, synth
    from oracle_ops import OracleOps
    ops = OracleOps(seed=0)
    rng = np.random.default_rng(3)
    P1, P2 = np.asarray(synth.KITTI_P0, float).reshape(3, 4), np.asarray(synth.KITTI_P1, float).reshape(3, 4)
    n = 41
    X = np.column_stack([rng.uniform(-10, 10, n), rng.uniform(-2, 2, n), rng.uniform(2, 150, n)])
    X[5, 2] = -4.0                                     # behind the camera
    def proj(P):
        h = np.c_[X, np.ones(n)] @ P.T
        return (h[:, :2] / h[:, 2:]).astype(np.float32)
    fl, fr = proj(P1), proj(P2)
    ang = 0.3
    pose = np.eye(4); pose[:3, :3] = [[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]]; pose[:3, 3] = [1, 2, 3]
    cur = rng.normal(size=(4, 3))
    for m in (n, n - 1, 1, 0):
        got = vo.create_landmarks_from_features(ops, fl[:m], fr[:m], P1, P2, pose, cur)
        lm = np.zeros((2, 3))                          # zeros(size(features_l, 2), 3)
        for i in range(1, m + 1, 2):                   # for i = 1:2:size(features_l, 1)
            c = ops.triangulate(fl[i - 1:i], fr[i - 1:i], P1, P2)[0]
            if c[2] < 0 or c[2] > 80:
                continue
            if i > len(lm):
                lm = np.vstack([lm, np.zeros((i - len(lm), 3))])
            lm[i - 1] = pose[:3, :3] @ c + pose[:3, 3]
        want = np.vstack([cur, lm])
        assert got.shape == want.shape, (m, got.shape, want.shape)
        assert np.allclose(got, want, rtol=0, atol=1e-9)
    # new-landmark selection: "old" as soon as any tracked point shares the x OR the y coordinate
    l = np.array([[1., 2.], [3., 4.], [5., 6.], [7., 8.]]); r = l + 100
    old_l = np.array([[3., 99.], [50., 6.]]); old_r = np.array([[107., 0.]])
    idx = vo.new_landmark_indices(l, r, old_l, old_r)
    want = [k for k in range(4) if not (old_l == l[k]).any() and not (old_r == r[k]).any()]
    assert idx.tolist() == want == [0]


def test_sequence_runner_and_pipeline_host_logic(tmp_path, monkeypatch):
    """Host logic of io.run_sequence (batch cutting with a one-frame halo, double-buffered decode) and of
    vo.FramePipeline (work queue over several contexts, results in input order) with the GPU call replaced
    by a stand-in that only records what it was given."""
    cv2 = pytest.importorskip("cv2")
    from vo_b200 import io, vo
    n, h, w = 11, 20, 30
    rng = np.random.default_rng(0)
    left = rng.integers(0, 256, (n, h, w), dtype=np.uint8); right = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    lf, rf = [], []
    for i in range(n):
        for name, arr, lst in (("l", left, lf), ("r", right, rf)):
            p = os.path.join(tmp_path, f"{name}{i:04d}.png"); assert cv2.imwrite(p, arr[i]); lst.append(p)
    calls = []

    def fake_run_frames(l, r, P1, P2, seed=0, first_frame=0, max_keypoints=8192, ctx=None, device_ptrs=None):
        m = len(l)
        calls.append((first_frame, m))
        assert np.array_equal(l, left[first_frame:first_frame + m]) and np.array_equal(r, right[first_frame:first_frame + m])
        rel = np.tile(np.eye(4), (m, 1, 1)); rel[:, 0, 3] = np.arange(first_frame, first_frame + m)   # frame index in tx
        return rel, np.zeros(m, np.int32), np.full((m, 8), first_frame, np.int32)
    monkeypatch.setattr(vo, "run_frames", fake_run_frames)
    rel, status, counts = io.run_sequence(lf, rf, np.eye(3, 4), np.eye(3, 4), batch=4, pinned=False)
    assert calls == [(0, 4), (3, 5), (7, 4)]                       # [lo, hi) with the one-frame halo
    assert np.array_equal(rel[:, 0, 3], np.arange(n))              # every frame's pose comes from the batch that owns it
    assert counts[:4, 0].tolist() == [0] * 4 and counts[4:8, 0].tolist() == [3] * 4 and counts[8:, 0].tolist() == [7] * 3

    # a PNG that fails to decode in a LATER batch (decoded by the background worker) must raise before that
    # buffer reaches vo_frames -- never silently return poses computed from a stale buffer
    from vo_b200 import VoError
    calls.clear()
    good = open(lf[9], "rb").read()
    open(lf[9], "wb").write(good[:len(good) // 2])
    with pytest.raises(VoError):
        io.run_sequence(lf, rf, np.eye(3, 4), np.eye(3, 4), batch=4, pinned=False)
    assert (7, 4) not in calls                                     # the batch holding frame 9 never ran
    open(lf[9], "wb").write(good)

    class FakeCtx:
        def close(self):
            pass
    monkeypatch.setattr(vo.api, "Context", lambda device=0: FakeCtx())
    pipe = vo.FramePipeline(depth=3)
    batches = [(left[i:i + 2], right[i:i + 2], i) for i in range(0, 10, 2)]
    out = pipe.map(batches, np.eye(3, 4), np.eye(3, 4))
    assert [int(o[2][0, 0]) for o in out] == [0, 2, 4, 6, 8]       # results in input order
    pipe.close()

    def boom(*a, **k):
        raise RuntimeError("device fault")
    monkeypatch.setattr(vo, "run_frames", boom)
    with pytest.raises(RuntimeError, match="device fault"):
        vo.FramePipeline(depth=2).map(batches, np.eye(3, 4), np.eye(3, 4))


def test_street_world_is_a_consistent_static_scene():
    """The benchmark workload (synth.StreetWorld): deterministic, and geometrically exact -- a pixel of the left
    image, lifted to 3-D with the renderer's own depth, lands on the same grey value in the right image (P1) and
    in the next frame's left image (relative ground-truth pose): one static world, exact stereo and motion."""
    cv2 = pytest.importorskip("cv2")
    from vo_b200 import synth
    d = np.load(os.path.join(G, "kitti00_reference_data.npz"))
    gt = np.tile(np.eye(4), (12, 1, 1)); gt[:, :3, :] = d["poses"][:12]
    w = synth.StreetWorld(gt, seed=3)
    l0, r0, zl0, zr0 = w.render(0, depth=True)
    l0b, r0b = synth.StreetWorld(gt, seed=3).render(0)
    assert np.array_equal(l0, l0b) and np.array_equal(r0, r0b)
    assert (l0 > 0).mean() > 0.5 and not np.array_equal(l0, r0)
    l1, _, zl1, _ = w.render(1, depth=True)
    K = synth.KITTI_P0[:, :3]
    ys, xs = np.mgrid[20:356:5, 20:1221:7]
    Z = zl0[ys, xs]
    ok = np.isfinite(Z)
    X = np.stack([(xs - K[0, 2]) * Z / K[0, 0], (ys - K[1, 2]) * Z / K[1, 1], Z], axis=-1)       # camera-0 coordinates
    rel = np.linalg.inv(gt[1]) @ gt[0]                                                            # camera 0 -> camera 1
    for img, zb, Rt in ((r0, zr0, np.c_[np.eye(3), [-synth.KITTI_BASELINE, 0, 0]]), (l1, zl1, rel[:3])):
        Y = X @ Rt[:, :3].T + Rt[:, 3]
        with np.errstate(invalid="ignore", divide="ignore"):
            u = K[0, 0] * Y[..., 0] / Y[..., 2] + K[0, 2]; v = K[1, 1] * Y[..., 1] / Y[..., 2] + K[1, 2]
        inside = ok & (u > 1) & (u < 1239) & (v > 1) & (v < 374)
        ui = np.clip(np.rint(np.nan_to_num(u)), 0, 1240).astype(int); vi = np.clip(np.rint(np.nan_to_num(v)), 0, 375).astype(int)
        with np.errstate(invalid="ignore"):
            vis = inside & (np.abs(zb[vi, ui] - Y[..., 2]) < 0.02 * Y[..., 2])                     # not occluded in the other view
        a = l0[ys, xs][vis].astype(np.float64)
        b = cv2.remap(img, np.nan_to_num(u).astype(np.float32), np.nan_to_num(v).astype(np.float32), cv2.INTER_LINEAR)[vis].astype(np.float64)
        assert vis.sum() > 2000 and np.corrcoef(a, b)[0, 1] > 0.9, (vis.sum(), np.corrcoef(a, b)[0, 1])


def test_descriptor_sets_kinds():
    from vo_b200 import synth
    for kind in ("integer", "ties", "float"):
        f1, f2 = synth.descriptor_sets(kind, 300, 400, seed=5)
        assert f1.shape == (300, 128) and f2.shape == (400, 128) and f1.dtype == np.float32
        g1, g2 = synth.descriptor_sets(kind, 300, 400, seed=5)
        assert np.array_equal(f1, g1) and np.array_equal(f2, g2)
    f1, f2 = synth.descriptor_sets("ties", 300, 400, seed=5)
    assert len(np.unique(f2, axis=0)) < 400 and np.array_equal(f1, np.rint(f1))
    f1, _ = synth.descriptor_sets("float", 300, 400, seed=5)
    assert np.allclose(np.linalg.norm(f1, axis=1), 1.0, atol=1e-5) and not np.array_equal(f1, np.rint(f1))


def test_bench_oracle_row_checker_equals_the_oracle():
    """bench.py spreads the oracle's best-2 search over worker processes (column chunks) and applies the acceptance
    tests of oracle/match.c to the merged scores: both must equal the oracle called directly, bit for bit."""
    import bench
    from oracle import oracle
    from vo_b200 import synth
    try:
        for kind in ("integer", "ties", "float"):
            f1, f2 = synth.descriptor_sets(kind, 64, 20000, seed=11)
            j, s1, s2 = bench.oracle_top2_rows(f1, f2)
            oj, os1, os2 = oracle.match_top2(f1, f2)
            assert np.array_equal(j, oj) and np.array_equal(s1.view(np.uint32), os1.view(np.uint32))
            assert np.array_equal(s2.view(np.uint32), os2.view(np.uint32))
            keep = bench.oracle_keep(s1, s2, len(f2))
            pairs, metric = oracle.match(f1, f2)
            assert np.array_equal(np.nonzero(keep)[0].astype(np.uint32), pairs[:, 0]) and np.array_equal(j[keep], pairs[:, 1])
    finally:
        if bench._POOL is not None:
            bench._POOL.terminate(); bench._POOL = None


def test_one_term_match_margin_covers_the_worst_bf16_rounding():
    """The general-float match contracts the bf16 high halves alone when a score bound decides the rows and certifies with
    |approx - exact| <= (1/128 + 1/2048) ||a|| ||b|| (csrc/vo_match.cu: ONE_TERM_EPS).  The worst case -- every element just
    below a bf16 rounding midpoint, all products of one sign -- stays inside it, as do random and scaled inputs."""
    def bf16(x):
        u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
        return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16).astype(np.uint32).view(np.float32)
    eps = 1.0 / 128 + 1.0 / 2048
    rng = np.random.default_rng(0)
    m = np.float32(1 + 2.0 ** -8 * (1 - 1e-4))
    sets = [(np.full((4, 128), m, np.float32), np.full((4, 128), m, np.float32)),
            (np.abs(rng.standard_normal((500, 128))).astype(np.float32), np.abs(rng.standard_normal((500, 128))).astype(np.float32)),
            ((rng.standard_normal((500, 128)) * 300).astype(np.float32), (rng.standard_normal((500, 128)) * 0.01).astype(np.float32))]
    worst = 0.0
    for a, b in sets:
        exact = np.einsum("ik,ik->i", a.astype(np.float64), b.astype(np.float64))
        approx = np.einsum("ik,ik->i", bf16(a).astype(np.float64), bf16(b).astype(np.float64))
        rel = np.abs(approx - exact) / (np.linalg.norm(a.astype(np.float64), axis=1) * np.linalg.norm(b.astype(np.float64), axis=1))
        worst = max(worst, float(rel.max()))
    assert 0.0077 < worst < eps, worst      # the adversarial rows reach 2^-7; the margin keeps 6 % of room
