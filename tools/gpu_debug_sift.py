"""First-light diagnostics for the SIFT kernels (run on the GPU box)."""
import sys, os, time
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import vo_b200
from vo_b200 import synth
from oracle import oracle
ctx = vo_b200.Context(0)
for shape, seed in [((120, 160), 1), ((376, 1241), 3)]:
    img = synth.texture(shape[0], shape[1], seed=seed)
    t = time.time(); pts = vo_b200.detectSIFTFeatures(img, capacity=16384, ctx=ctx); dt = time.time() - t
    t = time.time(); pts = vo_b200.detectSIFTFeatures(img, capacity=16384, ctx=ctx); dt2 = time.time() - t
    t = time.time(); okp, odesc = oracle.sift(img); dto = time.time() - t
    print(f"{shape}: gpu {len(pts)} kps ({dt*1e3:.1f} ms first, {dt2*1e3:.1f} ms second), oracle {len(okp)} ({dto*1e3:.0f} ms)")
    n = min(len(pts), len(okp))
    if n:
        same = sum(all(okp[f][i] == pts.kps[f][i] for f in ("x", "y", "size", "angle", "response", "octave")) for i in range(n))
        print(f"  positional exact keypoint matches {same}/{n}; exact desc rows {sum(np.array_equal(odesc[i], pts._features[i]) for i in range(n))}/{n}")
        print("  gpu first 3:", pts.kps[:3]); print("  ora first 3:", okp[:3])
imgs = np.stack([synth.texture(376, 1241, seed=s) for s in range(8)])
for rep in range(3):
    t = time.time(); r = vo_b200.sift_batch(imgs, capacity=8192, ctx=ctx); dt = time.time() - t
    print(f"batch 8 x 1241x376: {dt*1e3:.1f} ms ({[len(x) for x in r]})")
print("done")
