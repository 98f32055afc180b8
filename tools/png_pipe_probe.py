"""Where does the device-decode pipeline spend its time?  Decode-only and decode+frames rates for 1..8 host threads."""
import os, sys, tempfile, threading, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import cv2, torch, numpy as np
import vo_b200
from vo_b200 import io, synth, vo
import bench
B = 32
l, r, _ = bench.street_frames(B + 1)
d = tempfile.mkdtemp()
paths = []
for k, a in enumerate((l, r)):
    for i in range(B + 1):
        p = os.path.join(d, f"{k}_{i:04d}.png"); cv2.imwrite(p, a[i]); paths.append(p)
files = [open(p, "rb").read() for p in paths]
H, W = 376, 1241

def run(nthr, reps, mode):
    ctxs = [vo_b200.Context(0) for _ in range(nthr)]
    bufs = [torch.empty((2 * (B + 1), H, W), dtype=torch.uint8, device="cuda") for _ in range(nthr)]
    def work(k, n):
        for _ in range(n):
            if mode in ("read", "read+frames"):
                io.read_batch_dev(paths, H, W, bufs[k], ctxs[k])
            if mode in ("mem", "mem+frames"):
                io.decode_batch_dev(files, H, W, bufs[k], ctxs[k])
            if mode.endswith("frames"):
                vo.run_frames(None, None, synth.KITTI_P0, synth.KITTI_P1, seed=1, ctx=ctxs[k],
                              device_ptrs=(bufs[k][0].data_ptr(), bufs[k][B + 1].data_ptr(), B + 1, H, W))
    for k in range(nthr):
        work(k, 1)
    th = [threading.Thread(target=work, args=(k, reps)) for k in range(nthr)]
    t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    for c in ctxs: c.close()
    return nthr * reps * B / dt

for mode in ("mem", "read", "frames", "mem+frames"):
    print(mode, {n: round(run(n, 6, mode)) for n in (1, 2, 4, 8)}, flush=True)
