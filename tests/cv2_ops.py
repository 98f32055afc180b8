"""OpenCV operator set for vo.VisualOdometry (test infrastructure): the reference's loop with every toolbox call answered
by an independent implementation -- cv2.SIFT, cv2.BFMatcher with the matchFeatures rules applied to its distances,
cv2.triangulatePoints, cv2.solvePnPRansac (P3P).  Used to pin the ORACLE's trajectory, not the product."""
import numpy as np


class Cv2Ops:
    def __init__(self, seed=0):
        import cv2
        self.cv2 = cv2
        self.sift = cv2.SIFT_create()
        self.bf = cv2.BFMatcher(cv2.NORM_L2)
        cv2.setRNGSeed(int(seed) & 0x7FFFFFFF)

    def detect_and_extract(self, img):
        kps, desc = self.sift.detectAndCompute(np.ascontiguousarray(img), None)
        loc = np.array([k.pt for k in kps], dtype=np.float32).reshape(-1, 2) + np.float32(1.0)    # MATLAB's 1-based [x y]
        return (desc if desc is not None else np.zeros((0, 128), np.float32)), loc

    def matchFeatures(self, f1, f2):
        if len(f1) == 0 or len(f2) == 0:
            return np.zeros((0, 2), np.uint32)
        u1 = (f1 / np.maximum(np.linalg.norm(f1, axis=1, keepdims=True), 1e-12)).astype(np.float32)
        u2 = (f2 / np.maximum(np.linalg.norm(f2, axis=1, keepdims=True), 1e-12)).astype(np.float32)
        out = []
        for i, m in enumerate(self.bf.knnMatch(u1, u2, k=min(2, len(u2)))):
            s1 = m[0].distance ** 2
            s2 = m[1].distance ** 2 if len(m) > 1 else np.inf
            if s1 <= 0.04 and (len(m) < 2 or (1.0 if s2 < 1e-6 else s1 / s2) <= 0.6):
                out.append((i, m[0].trainIdx))
        return np.array(out, dtype=np.uint32).reshape(-1, 2)

    def triangulate(self, p1, p2, P1, P2):
        h = self.cv2.triangulatePoints(np.asarray(P1, np.float64), np.asarray(P2, np.float64),
                                       np.asarray(p1, np.float64).T, np.asarray(p2, np.float64).T)
        return (h[:3] / h[3]).T

    def estworldpose(self, image_points, world_points, K4, frame_index):
        cv2 = self.cv2
        n = len(image_points)
        if n < 4:
            return dict(A=np.eye(4), status=1, n_inliers=0, inliers=np.zeros(n, bool))
        K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.0]])
        ok, rv, tv, inl = cv2.solvePnPRansac(np.asarray(world_points, np.float64), np.asarray(image_points, np.float64), K, None,
                                             iterationsCount=1000, reprojectionError=1.0, confidence=0.99, flags=cv2.SOLVEPNP_P3P)
        if not ok or inl is None or len(inl) < 4:
            return dict(A=np.eye(4), status=2, n_inliers=0, inliers=np.zeros(n, bool))
        R = cv2.Rodrigues(rv)[0]
        A = np.eye(4); A[:3, :3] = R.T; A[:3, 3] = (-R.T @ tv).ravel()        # camera pose in the world frame (estworldpose)
        mask = np.zeros(n, bool); mask[inl.ravel()] = True
        return dict(A=A, status=0, n_inliers=int(mask.sum()), inliers=mask)
