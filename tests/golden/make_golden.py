"""Generates the committed golden vectors under tests/golden/ (run in the build container, where
cv2 4.13 and /root/reference exist; the GPU box only reads the .npz files).

The reference (MATLAB + closed-source toolbox) ships no tests or golden vectors, so the oracle is
pinned against independent implementations: OpenCV 4.13 SIFT / GaussianBlur / triangulatePoints / solveP3P,
NumPy float64 brute-force matching, the published Philox4x32-10 known answer, and the reference's
own data files (kitti/00/calib.txt, kitti/poses/00.txt).

    python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from vo_b200 import synth  # noqa: E402

cv2.setNumThreads(1)

# ---- SIFT: OpenCV keypoints + descriptors on two seeded textures
for tag, shape, seed in (("small", (120, 160), 1), ("wide", (188, 620), 2)):
    img = synth.texture(shape[0], shape[1], seed=seed)
    kps, desc = cv2.SIFT_create().detectAndCompute(img, None)
    k = np.array([[p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave] for p in kps], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, f"sift_cv2_{tag}.npz"), image=img, kps=k.astype(np.float32),
                        octave=np.array([p.octave for p in kps], dtype=np.int32), desc=desc.astype(np.uint8))
    print(tag, len(kps), "keypoints")

# ---- SIFT at the BASELINE image size: a 376 x 1241 frame of the benchmark's street sequence (bench.street_frames)
if os.environ.get("VO_GOLDEN_FULLSIZE", "1") == "1":
    sys.path.insert(0, ROOT)
    import bench  # noqa: E402
    img = bench.street_frames(225)[0][40]
    kps, desc = cv2.SIFT_create().detectAndCompute(img, None)
    k = np.array([[p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave] for p in kps], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "sift_cv2_kitti_size.npz"), image=img, kps=k.astype(np.float32),
                        octave=np.array([p.octave for p in kps], dtype=np.int32), desc=desc.astype(np.uint8))
    print("kitti_size", len(kps), "keypoints")

# ---- SIFT with non-default options (the MATLAB name-value pairs NumLayersInOctave, Sigma,
# ContrastThreshold, EdgeThreshold map onto these OpenCV constructor arguments)
img = synth.texture(150, 210, seed=9)
opt_sets = {"layers2": dict(nOctaveLayers=2), "layers4": dict(nOctaveLayers=4), "sigma1p2": dict(sigma=1.2),
            "contrast_edge": dict(contrastThreshold=0.08, edgeThreshold=5)}
store = {"image": img}
for tag, kw in opt_sets.items():
    kps, desc = cv2.SIFT_create(**kw).detectAndCompute(img, None)
    store[f"kps_{tag}"] = np.array([[p.pt[0], p.pt[1], p.size, p.angle] for p in kps], dtype=np.float32)
    store[f"desc_{tag}"] = desc.astype(np.uint8)
    print("options", tag, len(kps), "keypoints")
np.savez_compressed(os.path.join(HERE, "sift_cv2_options.npz"), **store)

# ---- Gaussian blur + base image (cv::resize + GaussianBlur)
img = synth.texture(96, 130, seed=5)
f = img.astype(np.float32)
blurs = {f"{s:.4f}": cv2.GaussianBlur(f, (0, 0), s) for s in (1.2263, 1.9466, 3.09)}
dbl = cv2.resize(f, (f.shape[1] * 2, f.shape[0] * 2), interpolation=cv2.INTER_LINEAR)
base = cv2.GaussianBlur(dbl, (0, 0), float(np.sqrt(1.6 ** 2 - 1.0)))
np.savez_compressed(os.path.join(HERE, "blur_cv2.npz"), image=img, base=base.astype(np.float32),
                    **{f"blur_{k}": v for k, v in blurs.items()})

# ---- matching: float64 brute force (unit-normalised SSD, nearest + second nearest)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import correlated_pair  # noqa: E402
f1, f2 = correlated_pair(300, 400, seed=42)
a = f1.astype(np.float64); a /= np.linalg.norm(a, axis=1, keepdims=True)
b = f2.astype(np.float64); b /= np.linalg.norm(b, axis=1, keepdims=True)
S = 2.0 - 2.0 * a @ b.T
order = np.argsort(S, axis=1, kind="stable")
j1 = order[:, 0]; s1 = S[np.arange(300), j1]; s2 = S[np.arange(300), order[:, 1]]
keep = (s1 <= 0.04) & (s1 / np.maximum(s2, 1e-30) <= 0.6)
np.savez_compressed(os.path.join(HERE, "match_f64.npz"), f1=f1.astype(np.uint8), f2=f2.astype(np.uint8),
                    j1=j1.astype(np.uint32), s1=s1, s2=s2, keep=keep)
print("match: kept", keep.sum())

# ---- triangulation: cv2.triangulatePoints with the reference's calibration (kitti/00/calib.txt)
rng = np.random.default_rng(3)
X = np.c_[rng.uniform(-20, 20, 64), rng.uniform(-3, 2, 64), rng.uniform(4, 80, 64)]
P0, P1 = synth.KITTI_P0, synth.KITTI_P1
proj = lambda P: (lambda h: h[:, :2] / h[:, 2:])(np.c_[X, np.ones(64)] @ P.T)
x1 = proj(P0) + rng.normal(0, 0.3, (64, 2)); x2 = proj(P1) + rng.normal(0, 0.3, (64, 2))
c = cv2.triangulatePoints(P0, P1, x1.T, x2.T)
np.savez_compressed(os.path.join(HERE, "triangulate_cv2.npz"), x1=x1, x2=x2, xyz=(c[:3] / c[3]).T, truth=X)

# ---- P3P: every solution cv2.solveP3P finds for 200 random three-point problems (KITTI intrinsics)
rng = np.random.default_rng(12)
K = np.array([[718.856, 0, 607.1928], [0, 718.856, 185.2157], [0, 0, 1.0]])
p3p_X, p3p_uv, p3p_n, p3p_R, p3p_t = [], [], [], [], []
for _ in range(200):
    R = cv2.Rodrigues(rng.normal(0, 0.3, 3))[0]; t = rng.normal(0, 1, 3)
    Xw = np.c_[rng.uniform(-10, 10, 3), rng.uniform(-3, 3, 3), rng.uniform(5, 40, 3)]
    Xc = Xw @ R.T + t
    uv = np.c_[K[0, 0] * Xc[:, 0] / Xc[:, 2] + K[0, 2], K[1, 1] * Xc[:, 1] / Xc[:, 2] + K[1, 2]]
    n, rv, tv = cv2.solveP3P(Xw.reshape(-1, 1, 3), uv.reshape(-1, 1, 2), K, None, flags=cv2.SOLVEPNP_P3P)
    Rs = np.zeros((4, 3, 3)); ts = np.zeros((4, 3))
    for i in range(n):
        Rs[i] = cv2.Rodrigues(rv[i])[0]; ts[i] = tv[i].ravel()
    p3p_X.append(Xw); p3p_uv.append(uv); p3p_n.append(n); p3p_R.append(Rs); p3p_t.append(ts)
np.savez_compressed(os.path.join(HERE, "p3p_cv2.npz"), X=np.array(p3p_X), uv=np.array(p3p_uv), n=np.array(p3p_n),
                    R=np.array(p3p_R), t=np.array(p3p_t), K=K)
print("p3p:", int(np.sum(p3p_n)), "solutions")

# ---- reference data files: calibration and the head of the ground-truth poses of sequence 00
ref = "/root/reference/kitti"
if os.path.isdir(ref):
    calib = [l.split()[1:] for l in open(os.path.join(ref, "00", "calib.txt")).read().strip().splitlines()]
    np.savez_compressed(os.path.join(HERE, "kitti00_reference_data.npz"),
                        calib=np.array(calib[:2], dtype=np.float64).reshape(2, 3, 4),
                        poses=np.loadtxt(os.path.join(ref, "poses", "00.txt"))[:400].reshape(-1, 3, 4),
                        times=np.loadtxt(os.path.join(ref, "00", "times.txt"))[:400])
    print("kitti: calib + 400 poses")
print("done")
