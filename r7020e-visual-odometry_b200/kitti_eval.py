"""Trajectory evaluation: the reference's xz-plane error (PlotOnMap.m:9-10, 20) and the KITTI
odometry devkit's t_err / r_err, plus KITTI-format pose I/O.

Poses are 4x4 camera-to-world matrices; ground truth rows are 3x4 row-major (kitti/poses/00.txt).
"""
import numpy as np

LENGTHS = (100, 200, 300, 400, 500, 600, 700, 800)


def load_poses(path):
    a = np.loadtxt(path).reshape(-1, 3, 4)
    out = np.tile(np.eye(4), (len(a), 1, 1))
    out[:, :3, :] = a
    return out


def save_poses(path, poses):
    np.savetxt(path, np.asarray(poses)[:, :3, :].reshape(len(poses), 12), fmt="%.9e")


def xz_error(est, truth):
    """Error(i) = ||[x z]_truth(i) - [x z]_est(i)|| (PlotOnMap.m:20).  The reference compares pose k
    (frame k+1, the first pose is appended at i = 2, VO.m:90,133) with ground-truth row k; pass
    est with the identity prepended to evaluate without that off-by-one."""
    est = np.asarray(est); truth = np.asarray(truth)
    n = min(len(est), len(truth))
    d = truth[:n, [0, 2], 3] - est[:n, [0, 2], 3]
    return np.linalg.norm(d, axis=1)


def _distances(poses):
    step = np.linalg.norm(np.diff(poses[:, :3, 3], axis=0), axis=1)
    return np.concatenate([[0.0], np.cumsum(step)])


def kitti_errors(est, truth, lengths=LENGTHS, step=10):
    """Returns (t_err [fraction], r_err [rad/m], n_segments) averaged over all (first, length)."""
    est = np.asarray(est); truth = np.asarray(truth)
    n = min(len(est), len(truth))
    dist = _distances(truth[:n])
    t_errs, r_errs = [], []
    for first in range(0, n, step):
        for L in lengths:
            last = np.searchsorted(dist, dist[first] + L, side="left")
            if last >= n:
                continue
            dg = np.linalg.inv(truth[first]) @ truth[last]
            de = np.linalg.inv(est[first]) @ est[last]
            E = np.linalg.inv(de) @ dg
            c = np.clip((np.trace(E[:3, :3]) - 1.0) * 0.5, -1.0, 1.0)
            r_errs.append(np.arccos(c) / L)
            t_errs.append(np.linalg.norm(E[:3, 3]) / L)
    if not t_errs:
        return float("nan"), float("nan"), 0
    return float(np.mean(t_errs)), float(np.mean(r_errs)), len(t_errs)
