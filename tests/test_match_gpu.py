"""GPU parity of vo_match (tcgen05 GEMM + exact re-rank) against the CPU oracle: indices and
metrics must be BIT-EXACT on identical descriptor inputs (north_star)."""
import numpy as np
import pytest

from conftest import correlated_pair, sift_like_descriptors
from oracle import oracle

pytestmark = pytest.mark.gpu


def _check(ctx, f1, f2, **kw):
    import vo_b200
    pairs, metric = vo_b200.matchFeatures(f1, f2, return_metric=True, ctx=ctx, **kw)
    okw = dict(match_threshold=kw.get("MatchThreshold", 1.0), max_ratio=kw.get("MaxRatio", 0.6),
               unique=kw.get("Unique", False), index_base=kw.get("index_base", 0))
    opairs, ometric = oracle.match(f1, f2, **okw)
    assert pairs.shape == opairs.shape, (pairs.shape, opairs.shape)
    assert np.array_equal(pairs, opairs)
    assert np.array_equal(metric.view(np.uint32), ometric.view(np.uint32))
    return pairs


def test_debug_gemm_is_exact_integer_dot(ctx):
    import vo_b200.api as api
    f1 = sift_like_descriptors(300, 11)
    f2 = sift_like_descriptors(700, 12)
    c = api.match_debug_gemm(f1, f2, ctx=ctx)
    ref = f1.astype(np.float64) @ f2.astype(np.float64).T
    assert np.array_equal(c.astype(np.float64), ref)


@pytest.mark.parametrize("n1,n2", [(1, 1), (1, 2), (2, 1), (3, 3), (5, 300), (127, 129), (128, 256),
                                   (129, 257), (1000, 777), (2048, 2048), (3000, 3100)])
def test_top2_bit_exact(ctx, n1, n2):
    import vo_b200.api as api
    f1, f2 = correlated_pair(n1, n2, seed=100 + n1 + n2)
    j1, s1, s2 = api.match_top2(f1, f2, ctx=ctx)
    oj1, os1, os2 = oracle.match_top2(f1, f2)
    assert np.array_equal(j1, oj1)
    assert np.array_equal(s1.view(np.uint32), os1.view(np.uint32))
    assert np.array_equal(s2.view(np.uint32), os2.view(np.uint32))
    assert ctx.match_stats()["exact_integer_path"]


@pytest.mark.parametrize("n1,n2", [(64, 64), (500, 900), (3000, 2500)])
def test_match_default_bit_exact(ctx, n1, n2):
    pairs = _check(ctx, *correlated_pair(n1, n2, seed=7 + n1))
    if n1 >= 500:
        assert len(pairs) > 50            # the synthetic pairs really do match


def test_match_options(ctx):
    f1, f2 = correlated_pair(800, 800, seed=3)
    _check(ctx, f1, f2, MatchThreshold=10.0, MaxRatio=0.8)
    _check(ctx, f1, f2, index_base=1)
    _check(ctx, f1, f2, Unique=True)


def test_empty_and_tiny(ctx):
    import vo_b200
    e = np.zeros((0, 128), dtype=np.float32)
    f = sift_like_descriptors(10, 1)
    assert vo_b200.matchFeatures(e, f, ctx=ctx).shape == (0, 2)
    assert vo_b200.matchFeatures(f, e, ctx=ctx).shape == (0, 2)
    _check(ctx, f[:1], f[:1])          # n2 == 1: no ratio test
    _check(ctx, f, f[:2])


def test_adversarial_ties(ctx):
    """5 % duplicated rows and rows differing by +-1 in one bin (SURVEY.md 8d config 4 ii): ties must
    resolve to the lowest index, exactly like the oracle; triple duplicates go through the row scan."""
    rng = np.random.default_rng(5)
    f2 = sift_like_descriptors(1500, 21)
    dup = rng.permutation(1500)[:150]
    f2[dup[:75]] = f2[dup[75:150]]                       # exact duplicates
    f2[dup[:10]] = f2[dup[10]]                           # a 10-fold duplicate
    near = rng.permutation(1500)[:75]
    f2[near, rng.integers(0, 128, 75)] += 1.0
    f1 = f2[rng.permutation(1500)[:900]].copy()
    f1[::3, 5] = np.clip(f1[::3, 5] + 1, 0, 255)
    import vo_b200.api as api
    j1, s1, s2 = api.match_top2(f1, f2, ctx=ctx)
    oj1, os1, os2 = oracle.match_top2(f1, f2)
    assert np.array_equal(j1, oj1)
    assert np.array_equal(s1.view(np.uint32), os1.view(np.uint32))
    assert np.array_equal(s2.view(np.uint32), os2.view(np.uint32))
    _check(ctx, f1, f2, MaxRatio=1.0)


def test_general_float_path(ctx):
    """Non-integer descriptors (unit-norm floats, as MATLAB might hand over): split-bf16 GEMM +
    exact FP32 re-rank must still reproduce the oracle bit for bit."""
    rng = np.random.default_rng(9)
    f2 = np.abs(rng.standard_normal((1200, 128))).astype(np.float32)
    f2 /= np.linalg.norm(f2, axis=1, keepdims=True)
    f1 = f2[rng.permutation(1200)[:700]] + rng.normal(0, 0.01, (700, 128)).astype(np.float32)
    f1 = f1.astype(np.float32)
    import vo_b200.api as api
    j1, s1, s2 = api.match_top2(f1, f2, ctx=ctx)
    st = ctx.match_stats()
    assert not st["exact_integer_path"]
    oj1, os1, os2 = oracle.match_top2(f1, f2)
    assert np.array_equal(j1, oj1)
    assert np.array_equal(s1.view(np.uint32), os1.view(np.uint32))
    assert np.array_equal(s2.view(np.uint32), os2.view(np.uint32))
    assert st["rowscan_rows"] < 0.2 * len(f1) and st["k_extent"] == 384      # exact top-2 of every row: three bf16 terms
    _check(ctx, f1, f2)
    st = ctx.match_stats()                       # matchFeatures: the score bound decides the rows, one bf16 term suffices
    assert not st["exact_integer_path"] and st["k_extent"] == 128 and st["rowscan_rows"] < 0.05 * len(f1)
    # unit-norm float copies of SIFT-like rows at a size with several column tiles and row panels
    from vo_b200 import synth
    h1, h2 = synth.descriptor_sets("float", 3000, 5000, seed=3)
    _check(ctx, h1, h2)
    st = ctx.match_stats()
    assert st["k_extent"] == 128 and st["rowscan_rows"] < 0.05 * len(h1)
    _check(ctx, h1, h2, Unique=True)
    _check(ctx, h1, h2, MatchThreshold=10.0, MaxRatio=0.9)       # a loose bound: more candidates per row
    # signed, non-normalised general floats
    g1 = rng.standard_normal((333, 128)).astype(np.float32) * 3.0
    g2 = rng.standard_normal((555, 128)).astype(np.float32) * 0.5
    _check(ctx, g1, g2, MatchThreshold=100.0, MaxRatio=1.0)


def test_general_float_decisions_near_the_thresholds(ctx):
    """Unit-norm float rows whose best scores are spread over 0 .. 0.1 (threshold 0.04) and whose landmark set holds
    near-duplicates (ratios around 0.6): the one-term general path has to prove every keep / reject or fall back to the
    exact scan -- pairs and metric bits equal the oracle's either way."""
    rng = np.random.default_rng(21)
    def unit(x):
        return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    def jitter(x, t):       # rows at squared distance ~t from the rows of x
        d = rng.standard_normal(x.shape)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        return unit(x + d * np.sqrt(t)[:, None])
    base = unit(np.abs(rng.standard_normal((1500, 128))))
    sib = jitter(base[:700], rng.uniform(0.0, 0.1, 700))
    f2 = np.concatenate([base, sib]).astype(np.float32)
    src = rng.permutation(len(f2))[:1800]
    f1 = jitter(f2[src], rng.uniform(0.0, 0.1, len(src)))
    pairs = _check(ctx, f1, f2)
    st = ctx.match_stats()
    assert not st["exact_integer_path"] and st["k_extent"] == 128
    assert 100 < len(pairs) < len(f1)
    assert st["rowscan_rows"] < 0.25 * len(f1), st
    _check(ctx, f1, f2, Unique=True)
    _check(ctx, f1, f2, MaxRatio=0.9)
    _check(ctx, f1, f2, MatchThreshold=2.0, MaxRatio=0.8)


def test_col_major_inputs(ctx):
    f1, f2 = correlated_pair(400, 500, seed=77)
    import vo_b200
    a = vo_b200.matchFeatures(np.asfortranarray(f1), np.asfortranarray(f2), ctx=ctx)
    b, _ = oracle.match(f1, f2)
    assert np.array_equal(a, b)


def test_large_linearity_property(ctx):
    """Full-size property (no oracle run): matching f2 against a row-permuted copy of itself returns
    the permutation with metric 0 for every row that has a unique nearest neighbour."""
    n = 16384
    f2 = sift_like_descriptors(n, 1234)
    perm = np.random.default_rng(1).permutation(n)
    f1 = f2[perm]
    import vo_b200.api as api
    j1, s1, s2 = api.match_top2(f1, f2, ctx=ctx)
    assert np.array_equal(j1, perm.astype(np.uint32))
    assert float(s1.max()) <= 1e-6          # 2 - 2c with c one or two ulps below 1
    assert float(s2.min()) > 0.1


def test_best2_records_row_sharded_equals_unsharded(ctx):
    """BASELINE config 5 in miniature: 16-byte best-2 records of row shards, concatenated (what the
    all-gather does), equal the unsharded records; kept rows equal matchFeatures' pairs and metric."""
    import ctypes as C
    import torch
    from vo_b200 import _lib, shard
    f1, f2 = correlated_pair(1500, 2300, seed=41)
    q = torch.from_numpy(f1).cuda(); l = torch.from_numpy(f2).cuda()
    full, _ = shard.relocalise_row_sharded_dev(ctx, q, l, 0, 1)
    parts = []
    for lo, hi in shard.row_chunks(len(f1), 3):
        rec, _ = shard.relocalise_row_sharded_dev(ctx, q[lo:hi].contiguous(), l, 0, 1)
        parts.append(rec.clone())
    full, cat = full.cpu().numpy(), torch.cat(parts, 0).cpu().numpy()
    keep = full[:, 3] == 1
    assert np.array_equal(keep, cat[:, 3] == 1)
    assert np.array_equal(full[keep], cat[keep])
    opairs, ometric = oracle.match(f1, f2)
    assert np.array_equal(np.nonzero(keep)[0].astype(np.uint32), opairs[:, 0])
    assert np.array_equal(full[keep, 0].view(np.uint32), opairs[:, 1])
    assert np.array_equal(full[keep, 1].view(np.uint32), ometric.view(np.uint32))


def test_peer_gather_records_equal_the_allgather_form(ctx):
    """vo_match_best2_gather_dev (the all-gather fused into the epilogue: records stored into every rank's CUDA-IPC
    buffer) on one rank, with the slices of a 3-way sharding written one after the other into the same gathered buffer,
    equals the plain records of the unsharded match; bench.py --gpus N checks the multi-process form on every rank."""
    import ctypes as C
    import torch
    from vo_b200 import _lib, shard
    f1, f2 = correlated_pair(1500, 2300, seed=43)
    q = torch.from_numpy(f1).cuda(); l = torch.from_numpy(f2).cuda()
    full, _ = shard.relocalise_row_sharded_dev(ctx, q, l, 0, 1)
    pg = shard.PeerGather(ctx, len(f1), 0, 1, None)
    got = pg.run(q, l)
    ctx.sync()
    assert torch.equal(got, full)
    L = _lib.lib()
    got.zero_()
    for lo, hi in shard.row_chunks(len(f1), 3):          # what three ranks would each store into this rank's buffer
        qs = q[lo:hi].contiguous()
        _lib.check(L.vo_match_best2_gather_dev(ctx.handle, C.c_void_p(qs.data_ptr()), hi - lo, C.c_void_p(l.data_ptr()), len(f2), 128,
                                               None, pg._table, 1, C.c_size_t(lo), C.c_void_p(ctx.stream)))
    ctx.sync()
    a, b = got.cpu().numpy(), full.cpu().numpy()
    keep = b[:, 3] == 1
    assert np.array_equal(a[:, 3] == 1, keep) and np.array_equal(a[keep], b[keep]) and keep.sum() > 100
    # a prepared landmark set (converted once, float rows no longer needed) gives the same records
    n2 = shard.prepare_landmarks(ctx, l)
    del l
    got.zero_()
    again = pg.run(q, None, n_landmarks=n2)
    ctx.sync()
    assert torch.equal(again, full)
    import vo_b200
    with pytest.raises(vo_b200.VoError, match="prepared"):
        pg.run(q, None, n_landmarks=n2 + 1)
    with pytest.raises(vo_b200.VoError, match="not integers"):
        shard.prepare_landmarks(ctx, q * 0.5)
    pg.close()


def test_cta_pair_kernel_bit_exact():
    """The experimental variants of the exact-integer kernel -- cta_group::2 CTA pairs (VO_MATCH_PAIRS=1) and
    A operand in TMEM (VO_MATCH_TS=1); switches are read once per process, hence the subprocess -- must
    return the same bit-exact results as the default kernel."""
    import os, subprocess, sys
    code = r'''
import sys, os, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import vo_b200, vo_b200.api as api
from conftest import correlated_pair
from oracle import oracle
ctx = vo_b200.Context(0)
for n1, n2, seed in [(1, 1, 1), (129, 257, 2), (1000, 777, 3), (3000, 3100, 4)]:
    f1, f2 = correlated_pair(n1, n2, seed=seed)
    j1, s1, s2 = api.match_top2(f1, f2, ctx=ctx)
    oj1, os1, os2 = oracle.match_top2(f1, f2)
    assert np.array_equal(j1, oj1) and np.array_equal(s1.view(np.uint32), os1.view(np.uint32)) and np.array_equal(s2.view(np.uint32), os2.view(np.uint32))
    p, m = vo_b200.matchFeatures(f1, f2, return_metric=True, ctx=ctx)
    op, om = oracle.match(f1, f2)
    assert np.array_equal(p, op) and np.array_equal(m.view(np.uint32), om.view(np.uint32))
print("pairs ok")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    variants = [dict(VO_MATCH_PAIRS="1", VO_MATCH_PAIR_TILE="256"), dict(VO_MATCH_PAIRS="1", VO_MATCH_PAIR_TILE="128"),
                dict(VO_MATCH_TS="1")]            # CTA pairs (two tile widths) and the A-in-TMEM kernel
    for v in variants:
        env = dict(os.environ, **v)
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "pairs ok" in r.stdout, str(v) + r.stdout + r.stderr


def test_match_large_property(ctx):
    """Above 16384 query rows matchFeatures takes the count/scan/scatter compaction path.  No oracle run
    at this size: queries are noisy copies of a permutation of the landmarks (plus unrelated rows), so
    the expected pairs are known by construction; cross-checked against the exact top-2 of the same data."""
    import vo_b200
    import vo_b200.api as api
    n = 20000
    rng = np.random.default_rng(8)
    f2 = sift_like_descriptors(n, 4321)
    perm = rng.permutation(n)
    f1 = sift_like_descriptors(n, 999)
    copied = rng.random(n) < 0.6
    noisy = np.clip(np.rint(f2[perm] + rng.normal(0, 4.0, (n, 128))), 0, 255).astype(np.float32)
    f1[copied] = noisy[copied]
    pairs, metric = vo_b200.matchFeatures(f1, f2, return_metric=True, ctx=ctx)
    assert np.all(np.diff(pairs[:, 0].astype(np.int64)) > 0)                 # ascending, no duplicates
    assert np.array_equal(pairs[:, 0], np.nonzero(copied)[0].astype(np.uint32))
    assert np.array_equal(pairs[:, 1], perm[copied].astype(np.uint32))
    j1, s1, s2 = api.match_top2(f1, f2, ctx=ctx)
    keep = (s1 <= np.float32(0.04)) & (s1 / np.maximum(s2, np.float32(1e-6)) <= np.float32(0.6))
    assert np.array_equal(np.nonzero(keep)[0].astype(np.uint32), pairs[:, 0])
    assert np.array_equal(j1[keep], pairs[:, 1]) and np.array_equal(s1[keep].view(np.uint32), metric.view(np.uint32))
    # Unique at this size (forward-backward): copies are mutual nearest neighbours, so nothing is lost
    upairs = vo_b200.matchFeatures(f1, f2, Unique=True, ctx=ctx)
    assert np.array_equal(upairs, pairs)


def test_best2_records_empty_sides(ctx):
    """Relocalisation records with an empty landmark set (every row: keep = 0, j1 = UINT32_MAX) and with
    no queries at all."""
    import torch
    from vo_b200 import shard
    q = torch.from_numpy(sift_like_descriptors(300, 5)).cuda()
    none = torch.empty((0, 128), dtype=torch.float32, device="cuda")
    rec, counts = shard.relocalise_row_sharded_dev(ctx, q, none, 0, 1)
    r = rec.cpu().numpy()
    assert r.shape == (300, 4) and (r[:, 3] == 0).all() and (r[:, 0].view(np.uint32) == 0xFFFFFFFF).all()
    rec, counts = shard.relocalise_row_sharded_dev(ctx, none, q, 0, 1)
    assert rec.shape[0] == 0 and counts == [0]


@pytest.mark.parametrize("n1,n2", [(64, 40000), (300, 70000)])
def test_skinny_problems_use_every_sm_and_stay_exact(ctx, n1, n2):
    """Few query rows against a long landmark list: the persistent form cuts the single row panel into many
    column segments (one per SM); candidates of all segments must merge to the oracle's answer."""
    import vo_b200.api as api
    f1, f2 = correlated_pair(n1, n2, seed=n1 + 7)
    j1, s1, s2 = api.match_top2(f1, f2, ctx=ctx)
    oj1, os1, os2 = oracle.match_top2(f1, f2)
    assert np.array_equal(j1, oj1)
    assert np.array_equal(s1.view(np.uint32), os1.view(np.uint32)) and np.array_equal(s2.view(np.uint32), os2.view(np.uint32))
    _check(ctx, f1, f2)
