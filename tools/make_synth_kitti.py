"""Write a small KITTI-odometry-shaped dataset (image_0/, image_1/, calib.txt, poses.txt) rendered along the
first poses of the reference's ground truth (tests/golden/kitti00_reference_data.npz) -- for trying
tools/run_kitti.py where the real KITTI images are not available.  usage: make_synth_kitti.py OUT_DIR [N]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, cv2
from vo_b200 import synth
out, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 12
d = np.load(os.path.join(R, "tests", "golden", "kitti00_reference_data.npz"))
gt = np.tile(np.eye(4), (n, 1, 1)); gt[:, :3, :] = d["poses"][:n]
left, right = synth.plane_world(gt, seed=7)
for name, arr in (("image_0", left), ("image_1", right)):
    os.makedirs(os.path.join(out, name), exist_ok=True)
    for i in range(n):
        cv2.imwrite(os.path.join(out, name, f"{i:06d}.png"), arr[i])
with open(os.path.join(out, "calib.txt"), "w") as f:
    for k, P in (("P0", d["calib"][0]), ("P1", d["calib"][1])):
        f.write(f"{k}: " + " ".join(f"{v:.12e}" for v in P.reshape(-1)) + "\n")
np.savetxt(os.path.join(out, "poses.txt"), gt[:, :3, :].reshape(n, 12), fmt="%.9e")
print("wrote", n, "stereo frames to", out)
