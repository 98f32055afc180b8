"""First-light diagnostics for the tcgen05 match GEMM (run on the GPU box)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import vo_b200, vo_b200.api as api
from conftest import sift_like_descriptors, correlated_pair
from oracle import oracle

ctx = vo_b200.Context(0)
for (n1, n2) in [(128, 256), (300, 700), (1000, 3000)]:
    f1 = sift_like_descriptors(n1, 11); f2 = sift_like_descriptors(n2, 12)
    c = api.match_debug_gemm(f1, f2, ctx=ctx)
    ref = f1.astype(np.float64) @ f2.astype(np.float64).T
    bad = c.astype(np.float64) != ref
    print(f"gemm {n1}x{n2}: mismatches {bad.sum()} / {bad.size}; max|diff| {np.nanmax(np.abs(c-ref)):.3g}; nan {np.isnan(c).sum()}")
    if bad.any():
        r, cc = np.nonzero(bad)
        print("  bad rows (first 20 uniq):", np.unique(r)[:20], " bad cols (first 20 uniq):", np.unique(cc)[:20])
        print("  sample got/ref:", c[r[0], cc[0]], ref[r[0], cc[0]], " c[0,:4]", c[0, :4], " ref[0,:4]", ref[0, :4])
        # is it a permutation of k? compare against partial-k dots
        for kk in (16, 32, 64, 128):
            part = f1[:, :kk].astype(np.float64) @ f2[:, :kk].astype(np.float64).T
            print(f"   matches first-{kk} partial dot: {(c == part).mean():.3f}")
for (n1, n2) in [(300, 700), (3000, 3000)]:
    f1, f2 = correlated_pair(n1, n2, 5)
    t = time.time(); j1, s1, s2 = api.match_top2(f1, f2, ctx=ctx); dt = time.time() - t
    oj1, os1, os2 = oracle.match_top2(f1, f2)
    print(f"top2 {n1}x{n2}: j1 equal {np.mean(j1 == oj1):.4f} s1 equal {np.mean(s1 == os1):.4f} s2 equal {np.mean(s2 == os2):.4f} stats {ctx.match_stats()} {dt*1e3:.1f} ms")
    t = time.time(); p = vo_b200.matchFeatures(f1, f2, ctx=ctx); dt = time.time() - t
    op, _ = oracle.match(f1, f2)
    print(f"match: pairs {p.shape} oracle {op.shape} equal {p.shape == op.shape and np.array_equal(p, op)} {dt*1e3:.1f} ms")
print("done")
