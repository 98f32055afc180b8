"""Run the B200 stereo VO on a KITTI odometry sequence laid out like the reference expects
(kitti/<seq>/image_0, image_1, calib.txt; optional ground truth kitti/poses/<seq>.txt) and write the
trajectory in KITTI format.  Mirrors what VO.m does end to end (VO.m:13-48 setup, 64-232 loop, 238-250
evaluation) with the PNG decode overlapped with the GPU.

    python tools/run_kitti.py kitti/00 --poses kitti/poses/00.txt --frames 500 --out poses_b200.txt
"""
import argparse, glob, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sequence", help="directory with image_0/, image_1/ and calib.txt")
    ap.add_argument("--poses", help="ground-truth poses (KITTI format) for the error report")
    ap.add_argument("--frames", type=int, default=0, help="use only the first N frames")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default="poses_b200.txt")
    ap.add_argument("--decode", default="host", choices=["host", "device"],
                    help="host: the library's PNG decoder on every core; device: inflate + un-filter kernels on the GPU")
    a = ap.parse_args()
    import vo_b200
    from vo_b200 import io, vo, kitti_eval
    left = sorted(glob.glob(os.path.join(a.sequence, "image_0", "*.png")))
    right = sorted(glob.glob(os.path.join(a.sequence, "image_1", "*.png")))
    if not left or len(left) != len(right):
        raise SystemExit(f"no stereo PNG pairs under {a.sequence}/image_0 and image_1")
    if a.frames:
        left, right = left[:a.frames], right[:a.frames]
    P = {}
    for line in open(os.path.join(a.sequence, "calib.txt")):            # VO.m:24-33: P0 and P1
        k, v = line.split(":", 1)
        P[k.strip()] = np.array(v.split(), dtype=np.float64).reshape(3, 4)
    t0 = time.time()
    if a.decode == "device":
        rel, status, counts = io.run_sequence_device(left, right, P["P0"], P["P1"], batch=a.batch, seed=a.seed)
    else:
        rel, status, counts = io.run_sequence(left, right, P["P0"], P["P1"], batch=a.batch, seed=a.seed)
    dt = time.time() - t0
    bad = np.nonzero(status[1:] != 0)[0] + 1
    for i in bad:
        rel[i] = np.eye(4)                                               # estworldpose failed: hold the pose
    poses = np.array([np.eye(4)] + vo.chain_poses(rel[1:]))              # VO.m:130
    kitti_eval.save_poses(a.out, poses)
    print(f"{len(left)} frames in {dt:.2f} s ({len(left) / dt:.0f} frames/s incl. PNG decode), "
          f"{len(bad)} pose failures, mean tracked points {counts[1:, 6].mean():.0f}; wrote {a.out}")
    if a.poses:
        gt = kitti_eval.load_poses(a.poses)[:len(poses)]
        t_err, r_err, n = kitti_eval.kitti_errors(poses, gt)
        xz = kitti_eval.xz_error(poses, gt)
        kitti = (f"KITTI t_err {100 * t_err:.3f} %  r_err {np.degrees(r_err):.5f} deg/m over {n} segments" if n
                 else "KITTI t_err / r_err: n/a (sequence shorter than 100 m)")
        print(f"{kitti}; xz error (PlotOnMap.m:20) max {xz.max():.2f} m, last {xz[-1]:.2f} m")


if __name__ == "__main__":
    main()
