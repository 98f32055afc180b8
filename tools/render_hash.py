"""Hashes of the street-world renderer's stages (is the rendering identical on another machine?)."""
import hashlib, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
from vo_b200 import synth
h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
big = synth._big_texture(1024, 1024, 701)
print("texture", h(big))
d = np.load(os.path.join(R, "tests", "golden", "kitti00_reference_data.npz"))
gt = np.tile(np.eye(4), (4, 1, 1)); gt[:, :3, :] = d["poses"][:4]
w = synth.StreetWorld(gt, seed=7)
print("big", h(w._big), "mip", h(w.planes[0][3][2]))
l, r = w.render(1)
print("frame", h(l), h(r))
