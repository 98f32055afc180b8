"""Profiling target: two matchFeatures-mode calls on unit-norm float descriptors (general-float path), n x n.
Used under ncu; numbers printed here are never bench values."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import ctypes as C
import torch
import vo_b200
from vo_b200 import _lib, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ctx = vo_b200.Context(0); L = _lib.lib()
a, b = synth.descriptor_sets("float", n, n, seed=1234)
f1 = torch.from_numpy(a).cuda(); f2 = torch.from_numpy(b).cuda()
i1 = torch.empty(n, dtype=torch.int32, device="cuda"); i2 = torch.empty_like(i1); mt = torch.empty(n, device="cuda")
npairs = torch.zeros(1, dtype=torch.int32, device="cuda")
for _ in range(2):
    _lib.check(L.vo_match_dev(ctx.handle, C.c_void_p(f1.data_ptr()), n, C.c_void_p(f2.data_ptr()), n, 128, None,
                              C.c_void_p(i1.data_ptr()), C.c_void_p(i2.data_ptr()), C.c_void_p(mt.data_ptr()),
                              C.c_void_p(npairs.data_ptr()), C.c_void_p(ctx.stream)))
ctx.sync()
print("pairs", int(npairs.item()))
