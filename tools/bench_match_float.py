"""General-float path of the match (unit-norm float descriptors, as MATLAB may hand over): timing of the
split-bf16 kernel + exact re-rank.  Usage: python tools/bench_match_float.py [n ...]"""
import os, sys, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
import vo_b200
from vo_b200 import _lib
from conftest import correlated_pair

ctx = vo_b200.Context(0)
L = _lib.lib()
for n in [int(x) for x in sys.argv[1:]] or [8192, 32768]:
    a, b = correlated_pair(n, n, seed=1234)
    a /= np.linalg.norm(a, axis=1, keepdims=True); b /= np.linalg.norm(b, axis=1, keepdims=True)
    f1 = torch.from_numpy(a.astype(np.float32)).cuda(); f2 = torch.from_numpy(b.astype(np.float32)).cuda()
    j1 = torch.empty(n, dtype=torch.int32, device="cuda"); i2 = torch.empty(n, dtype=torch.int32, device="cuda")
    s1 = torch.empty(n, device="cuda"); npairs = torch.zeros(1, dtype=torch.int32, device="cuda")
    def call():
        _lib.check(L.vo_match_dev(ctx.handle, C.c_void_p(f1.data_ptr()), n, C.c_void_p(f2.data_ptr()), n, 128, None,
                                  C.c_void_p(j1.data_ptr()), C.c_void_p(i2.data_ptr()), C.c_void_p(s1.data_ptr()),
                                  C.c_void_p(npairs.data_ptr()), C.c_void_p(ctx.stream)))
    for _ in range(3): call()
    ctx.sync(); ctx.profile_enable(True)
    for _ in range(5): call()
    ctx.sync(); prof = ctx.profile(); ctx.profile_enable(False)
    g = prof["match_gemm_topk"]; t = g["ms"] / g["launches"]
    print(n, "general-float match: gemm %.3f ms (%.0f TF on 2*n*n*128), pairs %d, prep %.3f ms, stages %s" % (
        t, 2.0 * n * n * 128 / t / 1e9, int(npairs.item()), prof["match_prep"]["ms"] / prof["match_prep"]["launches"],
        {k: round(v["ms"] / v["launches"], 3) for k, v in prof.items()}))
