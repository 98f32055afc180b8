"""The VO.m loop mirror with OpenCV operators (tests/cv2_ops.py) over the first n rendered street frames: KITTI t_err /
r_err beside the oracle's golden trajectory (tests/golden/street_oracle_225.npz).  CPU only; informational."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import bench
from vo_b200 import vo, synth
from cv2_ops import Cv2Ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 225
g = np.load(os.path.join(R, "tests", "golden", "street_oracle_225.npz"))
left, right, gt = bench.street_frames(n)
v = vo.VisualOdometry(synth.KITTI_P0, synth.KITTI_P1, Cv2Ops(seed=1))
t0 = time.time()
rel = [np.eye(4)]
for i in range(n):
    a = v.step(left[i], right[i])
    if i: rel.append(a)
dt = time.time() - t0
rel = np.array(rel)
e_cv = bench.trajectory_errors(rel, gt, n)
e_or = bench.trajectory_errors(g["rel"], gt, n)
d = [np.linalg.norm(rel[i][:3, 3] - g["rel"][i][:3, 3]) for i in range(1, n)]
print(f"{n} frames, OpenCV-operator loop {dt:.1f} s ({n / dt:.2f} frames/s on one thread)")
print("OpenCV loop:", {k: e_cv[k] for k in ("t_err_pct", "r_err_deg_per_m", "xz_err_final_m")})
print("oracle     :", {k: e_or[k] for k in ("t_err_pct", "r_err_deg_per_m", "xz_err_final_m")})
print(f"per-frame |t_cv - t_oracle|: median {np.median(d) * 100:.2f} cm, max {np.max(d) * 100:.2f} cm")
