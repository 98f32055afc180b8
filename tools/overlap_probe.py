"""Probe: throughput of the frame loop with N batches in flight (N contexts/streams driven by N host threads)."""
import os, sys, threading, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import vo_b200
from vo_b200 import synth, vo
import bench

B, steps, warm = 32, 6, 3
left, right = bench.make_frames(4, B, seed=20260)
dl, dr = left.cuda(), right.cuda()
for nctx in (1, 2, 3):
    ctxs = [vo_b200.Context(0) for _ in range(nctx)]
    def run(ci, n):
        for i in range(n):
            b = (i * nctx + ci) % 4
            vo.run_frames(None, None, synth.KITTI_P0, synth.KITTI_P1, seed=1, first_frame=0, ctx=ctxs[ci],
                          device_ptrs=(dl[b].data_ptr(), dr[b].data_ptr(), B + 1, 376, 1241))
    for ci in range(nctx):
        run(ci, warm)
    torch.cuda.synchronize()
    t0 = time.time()
    th = [threading.Thread(target=run, args=(ci, steps)) for ci in range(nctx)]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"{nctx} batches in flight: {nctx * steps * B / dt:.0f} frames/s ({1e3 * dt / (nctx * steps):.2f} ms per step)")
    for c in ctxs: c.close()
