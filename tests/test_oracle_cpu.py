"""CPU suite: the oracle against its golden vectors (OpenCV 4.13, NumPy float64, published known
answers, the reference's data files).  These pin the oracle; the GPU tests then compare the CUDA
path with the oracle."""
import os

import numpy as np
import pytest

from oracle import oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("tag", ["small", "wide", "kitti_size"])
def test_sift_oracle_vs_opencv(tag):
    """kitti_size: a 376 x 1241 frame of the benchmark's street sequence, i.e. the BASELINE image size."""
    from scipy.spatial import cKDTree
    g = np.load(os.path.join(G, f"sift_cv2_{tag}.npz"))
    okp, odesc = oracle.sift(g["image"])
    ck, cd = g["kps"], g["desc"].astype(np.float32)
    assert abs(len(okp) - len(ck)) <= 0.02 * len(ck)
    tree = cKDTree(ck[:, :2])
    o = np.stack([okp["x"], okp["y"]], 1)
    hit, dl2, exact = 0, [], 0
    for i in range(len(o)):
        best = None
        for c in tree.query_ball_point(o[i], 0.5):            # north_star: within 0.5 px
            da = abs(((okp["angle"][i] - ck[c, 3]) + 180.0) % 360.0 - 180.0)
            if da < 2.0 and abs(okp["size"][i] - ck[c, 2]) < 0.1 * ck[c, 2]:
                if best is None or da < best[0]:
                    best = (da, c)
        if best is None:
            continue
        hit += 1
        a = odesc[i] / np.linalg.norm(odesc[i]); b = cd[best[1]] / np.linalg.norm(cd[best[1]])
        dl2.append(np.linalg.norm(a - b)); exact += np.array_equal(odesc[i], cd[best[1]])
    assert hit / len(o) >= 0.95                                # >= 95 % repeatability
    assert np.percentile(dl2, 90) <= 0.02                      # descriptor L2 tolerance (unit vectors)
    assert exact / len(o) >= 0.8                               # most descriptors are bit-identical to OpenCV
    assert np.all(np.diff(okp["x"]) >= 0)                      # OpenCV output order
    assert np.array_equal(odesc, np.rint(odesc)) and odesc.max() <= 255 and odesc.min() >= 0


@pytest.mark.parametrize("tag,kw", [("layers2", dict(n_octave_layers=2)), ("layers4", dict(n_octave_layers=4)),
                                    ("sigma1p2", dict(sigma=1.2)),
                                    ("contrast_edge", dict(contrast_threshold=0.08, edge_threshold=5.0))])
def test_sift_oracle_options_vs_opencv(tag, kw):
    """The option handling (layers per octave, sigma, contrast / edge thresholds) is pinned too."""
    from scipy.spatial import cKDTree
    g = np.load(os.path.join(G, "sift_cv2_options.npz"))
    ck, cd = g[f"kps_{tag}"], g[f"desc_{tag}"].astype(np.float32)
    okp, odesc = oracle.sift(g["image"], **kw)
    assert len(okp) == len(ck)
    tree = cKDTree(ck[:, :2])
    hit = exact = 0
    for i in range(len(okp)):
        for c in tree.query_ball_point([okp["x"][i], okp["y"][i]], 0.5):
            da = abs(((okp["angle"][i] - ck[c, 3]) + 180.0) % 360.0 - 180.0)
            if da < 2.0 and abs(okp["size"][i] - ck[c, 2]) < 0.1 * ck[c, 2]:
                hit += 1; exact += np.array_equal(odesc[i], cd[c])
                break
    assert hit == len(okp)                                     # every keypoint within 0.5 px / 2 deg / 10 % size
    assert exact / len(okp) >= 0.9                             # descriptors bit-identical to OpenCV's


def test_blur_and_base_image_vs_opencv():
    g = np.load(os.path.join(G, "blur_cv2.npz"))
    f = g["image"].astype(np.float32)
    for key in g.files:
        if key.startswith("blur_"):
            assert np.abs(oracle.blur(f, float(key[5:])) - g[key]).max() < 2e-4
    assert np.abs(oracle.base_image(g["image"]) - g["base"]).max() < 2e-4
    r, taps = oracle.gauss_kernel(1.2263)
    assert r == 5 and abs(taps[0] + 2 * taps[1:].sum() - 1.0) < 1e-6


def test_math_primitives():
    xs = np.linspace(-30, 0, 4001, dtype=np.float32)
    e = np.array([oracle.lib().vo_oracle_expf(float(x)) for x in xs])
    assert np.max(np.abs(e - np.exp(xs.astype(np.float64))) / np.exp(xs.astype(np.float64))) < 4e-7
    rng = np.random.default_rng(0)
    for _ in range(500):
        y, x = rng.normal(size=2)
        a = oracle.lib().vo_oracle_atan2deg(float(y), float(x))
        t = np.degrees(np.arctan2(y, x)) % 360.0
        assert abs(((a - t) + 180) % 360 - 180) < 0.35          # OpenCV fastAtan2 accuracy class


def test_match_vs_float64_bruteforce():
    g = np.load(os.path.join(G, "match_f64.npz"))
    f1, f2 = g["f1"].astype(np.float32), g["f2"].astype(np.float32)
    j1, s1, s2 = oracle.match_top2(f1, f2)
    clear = (g["s2"] - g["s1"]) > 1e-5                         # rows where FP32 cannot flip the order
    assert np.array_equal(j1[clear], g["j1"][clear])
    assert np.allclose(s1, g["s1"], atol=2e-6) and np.allclose(s2, g["s2"], atol=2e-6)
    pairs, metric = oracle.match(f1, f2)
    margin = (np.abs(g["s1"] - 0.04) > 1e-5) & (np.abs(g["s1"] - 0.6 * g["s2"]) > 1e-5)
    keep = g["keep"]
    got = np.zeros(len(f1), bool); got[pairs[:, 0]] = True
    assert np.array_equal(got[margin], keep[margin])
    assert np.all(np.diff(pairs[:, 0].astype(np.int64)) > 0)   # ascending in the first column
    assert pairs.dtype == np.uint32 and metric.dtype == np.float32


def test_match_vs_opencv_bfmatcher():
    """An independent implementation of the exhaustive matcher: cv2.BFMatcher (L2, k = 2) on unit-normalised rows, with the
    matchFeatures rules applied to its distances (SSD = d^2 <= 0.04, SSD ratio <= 0.6).  Same pairs as the oracle, except
    rows whose decision sits within rounding of a threshold; the metrics agree to float precision."""
    cv2 = pytest.importorskip("cv2")
    from conftest import correlated_pair
    f1, f2 = correlated_pair(1500, 1800, seed=5)
    pairs, metric = oracle.match(f1, f2)
    u1 = (f1 / np.linalg.norm(f1, axis=1, keepdims=True)).astype(np.float32)
    u2 = (f2 / np.linalg.norm(f2, axis=1, keepdims=True)).astype(np.float32)
    knn = cv2.BFMatcher(cv2.NORM_L2).knnMatch(u1, u2, k=2)
    cv_pairs, cv_metric, near = [], [], set()
    for i, (a, b) in enumerate(knn):
        s1, s2 = a.distance ** 2, b.distance ** 2
        if min(abs(s1 - 0.04), abs(s1 / max(s2, 1e-6) - 0.6)) < 1e-4:
            near.add(i)
        if s1 <= 0.04 and s1 / max(s2, 1e-6) <= 0.6:
            cv_pairs.append((i, a.trainIdx)); cv_metric.append(s1)
    mine = {int(i): (int(j), float(m)) for (i, j), m in zip(pairs, metric)}
    theirs = {i: (j, m) for (i, j), m in zip(cv_pairs, cv_metric)}
    assert len(theirs) > 300
    for i in set(mine) | set(theirs):
        if i in near:
            continue
        assert i in mine and i in theirs, i
        assert mine[i][0] == theirs[i][0]
        assert abs(mine[i][1] - theirs[i][1]) < 2e-5


def test_match_semantics_edge_cases():
    rng = np.random.default_rng(1)
    f = np.rint(np.abs(rng.normal(0, 40, (6, 128)))).astype(np.float32)
    # duplicates: nearest = lowest index, second-nearest SSD = 0 -> ambiguous -> rejected
    f2 = np.vstack([f[0], f[0], f[1:]])
    j1, s1, s2 = oracle.match_top2(f[:1], f2)
    assert j1[0] == 0 and s1[0] <= 3e-7 and s2[0] <= 3e-7
    assert len(oracle.match(f[:1], f2)[0]) == 0
    # one candidate only: no ratio test
    assert len(oracle.match(f[:1], f[:1])[0]) == 1
    # empty sets
    assert oracle.match(np.zeros((0, 128), np.float32), f)[0].shape == (0, 2)
    assert oracle.match(f, np.zeros((0, 128), np.float32))[0].shape == (0, 2)
    # Unique = forward-backward consistency, 1-based numbering
    a = oracle.match(f, f, unique=True, index_base=1)[0]
    assert np.array_equal(a[:, 0], a[:, 1]) and a.min() == 1
    # scale invariance: rows are unit-normalised first
    p0 = oracle.match(f, f2, max_ratio=1.0)[0]
    p1 = oracle.match(f * 2.0, f2 * 0.5, max_ratio=1.0)[0]
    assert np.array_equal(p0, p1)


def test_triangulate_vs_opencv_and_closed_form():
    g = np.load(os.path.join(G, "triangulate_cv2.npz"))
    k = np.load(os.path.join(G, "kitti00_reference_data.npz"))
    P0, P1 = k["calib"]                                        # kitti/00/calib.txt:1-2
    xyz, err, valid = oracle.triangulate(g["x1"], g["x2"], P0, P1)
    rel = np.linalg.norm(xyz - g["xyz"], axis=1) / np.linalg.norm(g["xyz"], axis=1)
    assert rel.max() < 1e-9                                    # north_star tolerance is 1e-4
    assert valid.all()
    assert abs(-P1[0, 3] / P1[0, 0] - 0.53717) < 1e-4          # baseline from the reference's calib
    X = g["truth"]
    proj = lambda P: (lambda h: h[:, :2] / h[:, 2:])(np.c_[X, np.ones(len(X))] @ P.T)
    x1, x2 = proj(P0), proj(P1)
    xyz, _, _ = oracle.triangulate(x1, x2, P0, P1)
    assert np.allclose(xyz[:, 2], P0[0, 0] * 0.5371657 / (x1[:, 0] - x2[:, 0]), rtol=1e-6)


def test_philox_known_answer():
    assert list(oracle.philox([0, 0, 0, 0], [0, 0])) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert list(oracle.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    for n in (4, 5, 100):
        for t in range(50):
            idx = oracle.sample4(7, t, n)
            assert len(set(idx.tolist())) == 4 and idx.max() < n


def test_p3p_and_msac():
    import cv2
    K4 = np.array([718.856, 718.856, 607.1928, 185.2157])
    rng = np.random.default_rng(2)
    for _ in range(100):
        R = cv2.Rodrigues(rng.normal(0, 0.2, 3))[0]; t = rng.normal(0, 1, 3)
        Xw = np.c_[rng.uniform(-10, 10, 3), rng.uniform(-3, 3, 3), rng.uniform(5, 40, 3)]
        Xc = Xw @ R.T + t
        f = Xc / np.linalg.norm(Xc, axis=1, keepdims=True)
        Rs, ts = oracle.p3p_solve(f, Xw)
        assert 1 <= len(Rs) <= 4
        assert min(np.abs(Rs[i] - R).max() + np.abs(ts[i] - t).max() for i in range(len(Rs))) < 1e-6
        for Ri in Rs:
            assert np.allclose(Ri @ Ri.T, np.eye(3), atol=1e-9) and np.linalg.det(Ri) > 0
    n = 300
    R = cv2.Rodrigues(np.array([0.01, -0.03, 0.005]))[0]; t = np.array([0.05, -0.02, -0.8])
    Xw = np.c_[rng.uniform(-15, 15, n), rng.uniform(-3, 2, n), rng.uniform(5, 60, n)]
    Xc = Xw @ R.T + t
    uv = np.c_[K4[0] * Xc[:, 0] / Xc[:, 2] + K4[2], K4[1] * Xc[:, 1] / Xc[:, 2] + K4[3]] + rng.normal(0, 0.2, (n, 2))
    out = rng.random(n) < 0.3
    uv[out] += rng.normal(0, 30, (out.sum(), 2))
    r = oracle.p3p(uv, Xw, K4, seed=42)
    A = np.eye(4); A[:3, :3] = R.T; A[:3, 3] = -R.T @ t
    assert r["status"] == 0 and np.abs(r["A"] - A).max() < 0.05
    assert r["trials_run"] < 1000                               # adaptive stop
    assert (r["inliers"] & ~out).sum() > 0.8 * (~out).sum()
    full = oracle.p3p(uv, Xw, K4, seed=42, adaptive=False)
    assert full["trials_run"] == 1000 and full["n_inliers"] >= r["n_inliers"] - 5
    assert oracle.p3p(uv[:3], Xw[:3], K4)["status"] == 1        # fewer than 4 points
    same = oracle.p3p(uv, Xw, K4, seed=42)
    assert same["best_trial"] == r["best_trial"] and np.array_equal(same["A"], r["A"])   # seeded => reproducible


def test_msac_pose_vs_opencv_solvepnpransac():
    """estworldpose end to end against an independent robust estimator: cv2.solvePnPRansac (P3P minimal solver, 1 px
    reprojection threshold) on KITTI-like data with 30 % outliers.  Different samplers and scorings, so the comparison is
    what a user sees: the same pose (to the noise level) and the same inlier set (up to borderline points)."""
    cv2 = pytest.importorskip("cv2")
    K4 = np.array([718.856, 718.856, 607.1928, 185.2157])
    Km = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
    rng = np.random.default_rng(12)
    for trial in range(5):
        n = 400
        R = cv2.Rodrigues(rng.normal(0, 0.03, 3))[0]; t = np.array([rng.normal(0, 0.1), rng.normal(0, 0.05), -rng.uniform(0.3, 1.5)])
        Xw = np.c_[rng.uniform(-15, 15, n), rng.uniform(-3, 2, n), rng.uniform(5, 60, n)]
        Xc = Xw @ R.T + t
        uv = np.c_[K4[0] * Xc[:, 0] / Xc[:, 2] + K4[2], K4[1] * Xc[:, 1] / Xc[:, 2] + K4[3]] + rng.normal(0, 0.15, (n, 2))
        out = rng.random(n) < 0.3
        uv[out] += rng.normal(0, 40, (int(out.sum()), 2))
        r = oracle.p3p(uv, Xw, K4, seed=trial)
        ok, rv, tv, inl = cv2.solvePnPRansac(Xw, uv, Km, None, iterationsCount=1000, reprojectionError=1.0, confidence=0.99,
                                             flags=cv2.SOLVEPNP_P3P)
        assert ok and r["status"] == 0
        Rc = cv2.Rodrigues(rv)[0]
        A_cv = np.eye(4); A_cv[:3, :3] = Rc.T; A_cv[:3, 3] = (-Rc.T @ tv).ravel()      # camera pose in the world, as estworldpose returns it
        assert np.abs(r["A"][:3, :3] - A_cv[:3, :3]).max() < 2e-3
        assert np.linalg.norm(r["A"][:3, 3] - A_cv[:3, 3]) < 0.05
        cv_in = np.zeros(n, bool); cv_in[inl.ravel()] = True
        both = (r["inliers"] & cv_in).sum()
        assert both > 0.9 * max(r["inliers"].sum(), cv_in.sum())
        assert not (r["inliers"] & out & (np.linalg.norm(uv - (np.c_[K4[0] * Xc[:, 0] / Xc[:, 2] + K4[2], K4[1] * Xc[:, 1] / Xc[:, 2] + K4[3]]), axis=1) > 5)).any()


def test_p3p_solutions_equal_opencv_solvep3p():
    """The oracle's P3P (Gao) against OpenCV's solveP3P on 200 three-point problems: the same number of
    solutions, and each of them equal to 1e-5 (golden vectors: tests/golden/p3p_cv2.npz)."""
    g = np.load(os.path.join(G, "p3p_cv2.npz"))
    K = g["K"]
    total = 0
    for X, uv, n, Rc, tc in zip(g["X"], g["uv"], g["n"], g["R"], g["t"]):
        ray = np.c_[(uv[:, 0] - K[0, 2]) / K[0, 0], (uv[:, 1] - K[1, 2]) / K[1, 1], np.ones(3)]
        f = ray / np.linalg.norm(ray, axis=1, keepdims=True)
        Rs, ts = oracle.p3p_solve(f, X)
        assert len(Rs) == n
        for i in range(n):
            assert min(np.abs(Rs[j] - Rc[i]).max() + np.abs(ts[j] - tc[i]).max() for j in range(len(Rs))) < 1e-5
        total += n
    assert total > 300


def test_find_remaining_points_index_chain():
    """VO.m:280-334 semantics on a toy problem where every match is known: the four gathers must
    leave all eight arrays row-aligned."""
    from vo_b200 import vo
    from oracle_ops import OracleOps
    rng = np.random.default_rng(4)
    base = np.rint(np.abs(rng.normal(0, 40, (40, 128)))).astype(np.float32)
    ident = lambda n: np.arange(n, dtype=np.float32)[:, None].repeat(2, 1)
    old_ids = rng.permutation(40)[:30]
    old = dict(l_desc=base[old_ids], r_desc=base[old_ids], l_pos=ident(40)[old_ids], r_pos=ident(40)[old_ids] + 0.5)
    cl, cr = rng.permutation(40)[:35], rng.permutation(40)[:33]
    cur = dict(l_desc=base[cl], r_desc=base[cr], l_pos=ident(40)[cl], r_pos=ident(40)[cr] + 0.5)
    c, o, lm, rm, ks = vo.find_remaining_points(OracleOps(), old, cur)
    expect = set(old_ids) & set(cl) & set(cr)
    assert set(c["l_pos"][:, 0].astype(int)) == expect
    assert np.array_equal(c["l_pos"], o["l_pos"]) and np.array_equal(c["r_pos"], o["r_pos"])
    assert np.array_equal(c["l_pos"] + 0.5, c["r_pos"])
    assert ks[3] == len(expect) and len(c["l_desc"]) == len(o["r_desc"]) == len(expect)


def test_oracle_trajectory_vs_an_opencv_operator_loop():
    """The whole path against independent implementations: the reference's loop (vo.VisualOdometry, the line-by-line
    mirror of VO.m) with every toolbox call answered by OpenCV (tests/cv2_ops.py: cv2.SIFT, BFMatcher + matchFeatures
    rules, triangulatePoints, solvePnPRansac) on the first 16 rendered street frames, against the oracle's golden
    trajectory (tests/golden/street_oracle_225.npz).  Same number of tracked points per frame to within 1 %, relative
    poses equal to the scatter of two RANSAC samplers (centimetres / hundredths of a degree over 0.86 m steps)."""
    pytest.importorskip("cv2")
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from vo_b200 import vo, synth
    from cv2_ops import Cv2Ops
    g = np.load(os.path.join(G, "street_oracle_225.npz"))
    n = 16
    left, right, gt = bench.street_frames(n)
    v = vo.VisualOdometry(synth.KITTI_P0, synth.KITTI_P1, Cv2Ops(seed=1))
    dts, dRs = [], []
    for i in range(n):
        a = v.step(left[i], right[i])
        if i == 0:
            continue
        o = g["rel"][i]
        assert v.log[-1]["status"] == 0 and g["status"][i] == 0
        assert abs(int(v.log[-1]["k4"]) - int(g["tracked"][i - 1])) <= max(5, 0.01 * g["tracked"][i - 1])
        dts.append(np.linalg.norm(a[:3, 3] - o[:3, 3]))
        dRs.append(np.degrees(np.arccos(np.clip((np.trace(a[:3, :3].T @ o[:3, :3]) - 1) / 2, -1, 1))))
    assert max(dts) < 0.06 and np.median(dts) < 0.02, dts          # metres per 0.86 m step
    assert max(dRs) < 0.2 and np.median(dRs) < 0.06, dRs          # degrees
