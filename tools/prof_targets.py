"""Short profiling target: 2 steps of vo_frames_dev (8+1 frames) and 2 calls of the 32768^2 match
GEMM.  Used under ncu (see profiles/README.md); numbers printed here are never bench values."""
import os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import vo_b200
from vo_b200 import synth, vo, _lib
from conftest import sift_like_descriptors
import ctypes as C

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
ctx = vo_b200.Context(0)
import bench
l, r, _ = bench.street_frames(B + 1)              # the benchmark workload (rendered street sequence)
dl, dr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
for i in range(2):
    rel, st, cnt = vo.run_frames(None, None, synth.KITTI_P0, synth.KITTI_P1, seed=1, ctx=ctx,
                                 device_ptrs=(dl.data_ptr(), dr.data_ptr(), B + 1, 376, 1241))
print("frames ok", cnt[:, :3].mean(0), ctx.kernel_launches())
f1 = torch.from_numpy(sift_like_descriptors(N, 1234)).cuda(); f2 = torch.from_numpy(sift_like_descriptors(N, 5678)).cuda()
j1 = torch.empty(N, dtype=torch.int32, device="cuda"); s1 = torch.empty(N, device="cuda"); s2 = torch.empty(N, device="cuda")
for i in range(2):
    _lib.check(_lib.lib().vo_match_top2_dev(ctx.handle, C.c_void_p(f1.data_ptr()), N, C.c_void_p(f2.data_ptr()), N, 128,
                                            C.c_void_p(j1.data_ptr()), C.c_void_p(s1.data_ptr()), C.c_void_p(s2.data_ptr()), C.c_void_p(ctx.stream)))
ctx.sync()
print("match ok", ctx.kernel_launches())
