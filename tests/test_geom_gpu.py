"""GPU parity of vo_triangulate / vo_p3p against the CPU oracle and closed-form known answers."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu

P0 = np.array([[718.856, 0, 607.1928, 0], [0, 718.856, 185.2157, 0], [0, 0, 1, 0.0]])
P1 = P0.copy()
P1[0, 3] = -386.1448
K4 = np.array([718.856, 718.856, 607.1928, 185.2157])


def _proj(P, X):
    h = np.c_[X, np.ones(len(X))] @ P.T
    return h[:, :2] / h[:, 2:]


def _scene(n, seed, noise=0.0):
    rng = np.random.default_rng(seed)
    X = np.c_[rng.uniform(-20, 20, n), rng.uniform(-3, 2, n), rng.uniform(4, 80, n)]
    return X, _proj(P0, X) + rng.normal(0, noise, (n, 2)), _proj(P1, X) + rng.normal(0, noise, (n, 2))


@pytest.mark.parametrize("n", [1, 2, 130, 1000])
def test_triangulate_matches_oracle(ctx, n):
    import vo_b200
    X, x1, x2 = _scene(n, n, noise=0.3)
    xyz, err, valid = vo_b200.triangulate(x1, x2, P0, P1, full=True, ctx=ctx)
    oxyz, oerr, ovalid = oracle.triangulate(x1, x2, P0, P1)
    # tolerance from north_star: 1e-4 relative; the kernel shares the oracle's op order, so ~1e-12
    assert np.max(np.linalg.norm(xyz - oxyz, axis=1) / np.linalg.norm(oxyz, axis=1)) < 1e-9
    assert np.allclose(err, oerr, rtol=1e-6, atol=1e-9)
    assert np.array_equal(valid, ovalid)


def test_triangulate_known_answer_and_single(ctx):
    import vo_b200
    X, x1, x2 = _scene(200, 3)
    xyz = vo_b200.triangulate(x1, x2, P0, P1, ctx=ctx)
    assert np.max(np.abs(xyz - X) / np.abs(X).max()) < 1e-9
    # closed form for a rectified pair: Z = f*b/(x1-x2)
    z = 718.856 * (386.1448 / 718.856) / (x1[:, 0] - x2[:, 0])
    assert np.allclose(xyz[:, 2], z, rtol=1e-9)
    # float32 1x2 inputs like the reference's loop (VO.m:114): output class single
    s = vo_b200.triangulate(x1[0].astype(np.float32), x2[0].astype(np.float32), P0, P1, ctx=ctx)
    assert s.dtype == np.float32 and s.shape == (1, 3)
    assert np.allclose(s[0], X[0], rtol=2e-3)
    assert vo_b200.triangulate(np.zeros((0, 2)), np.zeros((0, 2)), P0, P1, ctx=ctx).shape == (0, 3)


def _pose_problem(n, seed, outliers=0.3, noise=0.2):
    import cv2
    rng = np.random.default_rng(seed)
    R = cv2.Rodrigues(rng.normal(0, 0.03, 3))[0]
    t = np.array([0.05, -0.02, -0.8]) + rng.normal(0, 0.05, 3)
    Xw = np.c_[rng.uniform(-15, 15, n), rng.uniform(-3, 2, n), rng.uniform(5, 60, n)]
    Xc = Xw @ R.T + t
    uv = np.c_[K4[0] * Xc[:, 0] / Xc[:, 2] + K4[2], K4[1] * Xc[:, 1] / Xc[:, 2] + K4[3]]
    uv += rng.normal(0, noise, (n, 2))
    out = rng.random(n) < outliers
    uv[out] += rng.normal(0, 30, (out.sum(), 2))
    A = np.eye(4)
    A[:3, :3] = R.T
    A[:3, 3] = -R.T @ t
    return uv, Xw, A, ~out


@pytest.mark.parametrize("n,seed", [(4, 1), (50, 2), (300, 3), (1500, 4)])
def test_p3p_matches_oracle(ctx, n, seed):
    import vo_b200
    uv, Xw, A_true, good = _pose_problem(n, seed, outliers=0.0 if n == 4 else 0.3)
    for adaptive in (True, False):
        g = vo_b200.estworldpose(uv, Xw, K4, Seed=42 + seed, Adaptive=adaptive, full=True, ctx=ctx)
        o = oracle.p3p(uv, Xw, K4, seed=42 + seed, adaptive=adaptive)
        assert g["status"] == o["status"]
        assert g["best_trial"] == o["best_trial"]
        assert g["trials_run"] == o["trials_run"]
        assert g["n_inliers"] == o["n_inliers"]
        assert np.array_equal(g["inliers"], o["inliers"])
        assert np.allclose(g["A"], o["A"], rtol=0, atol=1e-9)
    if n >= 50:
        assert np.abs(g["A"] - A_true).max() < 0.05
        assert (g["inliers"] & good).sum() >= 0.8 * good.sum()


def test_p3p_status_and_errors(ctx):
    import vo_b200
    uv, Xw, _, _ = _pose_problem(3, 1)
    r = vo_b200.estworldpose(uv, Xw, K4, full=True, ctx=ctx)
    assert r["status"] == 1 and np.array_equal(r["A"], np.eye(4))
    with pytest.raises(vo_b200.VoError):
        vo_b200.estworldpose(uv, Xw, K4, ctx=ctx)               # mirrors estworldpose erroring
    rng = np.random.default_rng(0)
    r = vo_b200.estworldpose(rng.uniform(0, 1000, (40, 2)), rng.uniform(-5, 5, (40, 3)) + [0, 0, 20],
                             K4, full=True, ctx=ctx)
    o = oracle.p3p(r["inliers"].astype(float)[:0].reshape(0, 2), np.zeros((0, 3)), K4)
    assert o["status"] == 1
    assert r["status"] in (0, 2)
