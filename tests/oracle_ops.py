"""CPU-oracle operator set for vo.VisualOdometry (test infrastructure: the same loop, every toolbox
call answered by oracle/ instead of the GPU)."""
import numpy as np

from oracle import oracle


class OracleOps:
    def __init__(self, seed=0, unique=False):
        self.seed = seed
        self.unique = unique

    def detect_and_extract(self, img):
        kps, desc = oracle.sift(img)
        loc = np.stack([kps["x"], kps["y"]], axis=1).astype(np.float32) + np.float32(1.0)
        return desc, loc

    def matchFeatures(self, f1, f2):
        return oracle.match(f1, f2, unique=self.unique)[0]

    def triangulate(self, p1, p2, P1, P2):
        return oracle.triangulate(np.asarray(p1, np.float64), np.asarray(p2, np.float64), P1, P2)[0]

    def estworldpose(self, image_points, world_points, K4, frame_index):
        seed = (self.seed + frame_index * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        return oracle.p3p(image_points, world_points, K4, seed=seed)
