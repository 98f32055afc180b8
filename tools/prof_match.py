"""Profiling target: matchFeatures-mode (and optionally exact top-2) match at n x n x 128 on
device-resident descriptors.  Used under ncu; numbers printed here are never bench values."""
import os, sys, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import vo_b200
from vo_b200 import _lib
from conftest import correlated_pair

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
mode = sys.argv[2] if len(sys.argv) > 2 else "match"
ctx = vo_b200.Context(0)
a, b = correlated_pair(n, n, seed=1234)
f1 = torch.from_numpy(a).cuda(); f2 = torch.from_numpy(b).cuda()
j1 = torch.empty(n, dtype=torch.int32, device="cuda"); i2 = torch.empty(n, dtype=torch.int32, device="cuda")
s1 = torch.empty(n, device="cuda"); s2 = torch.empty(n, device="cuda"); npairs = torch.zeros(1, dtype=torch.int32, device="cuda")
L = _lib.lib()
for _ in range(3):
    if mode == "match":
        _lib.check(L.vo_match_dev(ctx.handle, C.c_void_p(f1.data_ptr()), n, C.c_void_p(f2.data_ptr()), n, 128, None,
                                  C.c_void_p(j1.data_ptr()), C.c_void_p(i2.data_ptr()), C.c_void_p(s1.data_ptr()),
                                  C.c_void_p(npairs.data_ptr()), C.c_void_p(ctx.stream)))
    else:
        _lib.check(L.vo_match_top2_dev(ctx.handle, C.c_void_p(f1.data_ptr()), n, C.c_void_p(f2.data_ptr()), n, 128,
                                       C.c_void_p(j1.data_ptr()), C.c_void_p(s1.data_ptr()), C.c_void_p(s2.data_ptr()),
                                       C.c_void_p(ctx.stream)))
ctx.sync()
print("ok", mode, n, int(npairs.item()))
