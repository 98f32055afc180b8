"""vo_frames on a 33-frame stack held in pageable host memory (what a MATLAB session or NumPy hands over), in pinned
memory and on the device: ms per call for VO_UPLOAD_THREADS = 1 (plain cudaMemcpy2DAsync), 2, 4, 8.  Not a bench value."""
import os, subprocess, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
    import numpy as np, torch
    import vo_b200
    from vo_b200 import synth, vo
    import bench
    l, r, _ = bench.street_frames(33)
    ctx = vo_b200.Context(0)
    pl, pr = torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory()
    dl, dr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
    def t(f, n=10):
        f(); f()
        t0 = time.perf_counter()
        for _ in range(n): out = f()
        return 1e3 * (time.perf_counter() - t0) / n, out
    a, ra = t(lambda: vo.run_frames(l, r, synth.KITTI_P0, synth.KITTI_P1, seed=1, ctx=ctx))
    b, rb = t(lambda: vo.run_frames(pl.numpy(), pr.numpy(), synth.KITTI_P0, synth.KITTI_P1, seed=1, ctx=ctx))
    c, rc = t(lambda: vo.run_frames(None, None, synth.KITTI_P0, synth.KITTI_P1, seed=1, ctx=ctx, device_ptrs=(dl.data_ptr(), dr.data_ptr(), 33, l.shape[1], l.shape[2])))
    same = all(np.array_equal(x, y) for x, y in zip(ra, rc)) and all(np.array_equal(x, y) for x, y in zip(rb, rc))
    print(f"threads={os.environ.get('VO_UPLOAD_THREADS')}: pageable {a:.2f} ms, pinned {b:.2f} ms, device-resident {c:.2f} ms per call of 32 frames; equal={same}", flush=True)
else:
    for thr in ("1", "2", "4", "8"):
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=dict(os.environ, VO_UPLOAD_THREADS=thr))
