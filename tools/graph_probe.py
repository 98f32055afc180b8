"""vo_frames_dev, one batch of 32 + 1 street frames at a time on one stream: wall time and host CPU time per call with
plain launches and with the captured CUDA graph (vo_frames_use_graph).  Numbers printed here are not bench values."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
import vo_b200
from vo_b200 import synth, vo
import bench
B = 32
l, r, _ = bench.street_frames(2 * B + 1)
dl, dr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
H, W = l.shape[1:]
for graph in (False, True, False, True):
    ctx = vo_b200.Context(0)
    ctx.use_frames_graph(graph)
    def call(k):
        lo = (k % 2) * B
        return vo.run_frames(None, None, synth.KITTI_P0, synth.KITTI_P1, seed=1, first_frame=lo, ctx=ctx,
                             device_ptrs=(dl[lo:].data_ptr(), dr[lo:].data_ptr(), B + 1, H, W))
    for k in range(4): out = call(k)
    n = 20
    t0, c0 = time.perf_counter(), time.process_time()
    for k in range(n): out = call(k)
    t1, c1 = time.perf_counter(), time.process_time()
    print(f"graph={int(graph)} state={ctx.frames_graph_state()}: {1e3 * (t1 - t0) / n:.3f} ms per call of {B} frames, host CPU {1e3 * (c1 - c0) / n:.3f} ms per call, "
          f"launches per call {ctx.kernel_launches() // (n + 4)}", flush=True)
    del ctx
