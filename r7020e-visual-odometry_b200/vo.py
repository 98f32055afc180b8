"""The reference's frame loop (VO.m:64-232) on top of the B200 operators.

Two forms:

* ``VisualOdometry`` -- a line-by-line mirror of VO.m's loop body and of ``find_remaining_points``
  (VO.m:280-334) written against an operator set (``ops``) with the MATLAB call names.  The
  default operator set is the CUDA API (api.py); the parity tests plug the CPU oracle into the
  same loop, so the two runs differ only in who computes each toolbox call.
* ``run_frames`` -- the batched device-resident loop (C ABI ``vo_frames``): same arithmetic, no
  host round trips inside a batch; this is what the benchmark times.

Positions follow MATLAB: ``Location`` is 1-based (index_base=1) exactly as VO.m feeds it to
``triangulate`` and ``estworldpose``.
"""
import ctypes as C
import numpy as np

from . import _lib, api
from ._lib import check


class CudaOps:
    """The toolbox calls of VO.m bound to libvo_b200 (no CPU fallback)."""

    def __init__(self, ctx=None, seed=0, capacity=16384, Unique=False):
        self.ctx = ctx or api.default_context()
        self.seed = seed
        self.capacity = capacity
        self.Unique = Unique

    def detect_and_extract(self, img):
        pts = api.detectSIFTFeatures(img, index_base=1, capacity=self.capacity, ctx=self.ctx)
        feats, pts = api.extractFeatures(img, pts, "SIFT")
        return feats, pts.Location

    def matchFeatures(self, f1, f2):
        return api.matchFeatures(f1, f2, Unique=self.Unique, ctx=self.ctx)

    def triangulate(self, p1, p2, P1, P2):
        return api.triangulate(np.asarray(p1, np.float64), np.asarray(p2, np.float64), P1, P2, ctx=self.ctx)

    def estworldpose(self, image_points, world_points, K4, frame_index):
        seed = (self.seed + frame_index * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        return api.estworldpose(image_points, world_points, K4, Seed=seed, full=True, ctx=self.ctx)


def find_remaining_points(ops, old, cur):
    """VO.m:280-334.  ``old``: stereo-matched, row-aligned dict(l_desc, r_desc, l_pos, r_pos);
    ``cur``: raw sets of the current frame.  Returns (cur, old, lm, rm, counts)."""
    old = dict(old)
    cur = dict(cur)
    lm = ops.matchFeatures(cur["l_desc"], old["l_desc"])                       # VO.m:283
    for k in ("l_desc", "l_pos", "r_desc", "r_pos"):                           # VO.m:287-290
        old[k] = old[k][lm[:, 1]]
    rm = ops.matchFeatures(cur["r_desc"], old["r_desc"])                       # VO.m:293
    for k in ("l_desc", "l_pos", "r_desc", "r_pos"):                           # VO.m:297-300
        old[k] = old[k][rm[:, 1]]
    cur["l_desc"] = cur["l_desc"][lm[:, 0]]; cur["l_pos"] = cur["l_pos"][lm[:, 0]]   # VO.m:305-306
    cur["r_desc"] = cur["r_desc"][rm[:, 0]]; cur["r_pos"] = cur["r_pos"][rm[:, 0]]   # VO.m:307-308
    m3 = ops.matchFeatures(cur["l_desc"], cur["r_desc"])                       # VO.m:311
    cur["l_desc"] = cur["l_desc"][m3[:, 0]]; cur["l_pos"] = cur["l_pos"][m3[:, 0]]   # VO.m:314-315
    cur["r_desc"] = cur["r_desc"][m3[:, 1]]; cur["r_pos"] = cur["r_pos"][m3[:, 1]]   # VO.m:316-317
    m4 = ops.matchFeatures(cur["l_desc"], old["l_desc"])                       # VO.m:323
    for k in ("l_desc", "l_pos", "r_desc", "r_pos"):                           # VO.m:326-333
        old[k] = old[k][m4[:, 1]]
        cur[k] = cur[k][m4[:, 0]]
    return cur, old, lm, rm, (len(lm), len(rm), len(m3), len(m4))


def new_landmark_indices(l_loc, r_loc, old_l_loc, old_r_loc):
    """VO.m:147-154: indices (0-based) of stereo-matched features that "did not exist in the previous
    frame".  The reference tests ``isempty(find(old.Location == loc(index,:), 1))`` -- an N x 2 array
    compared with a 1 x 2 row, so a feature is old as soon as ANY tracked point shares its x OR its y
    coordinate, in the left or in the right image.  Mirrored literally (vectorised)."""
    l_loc = np.asarray(l_loc); r_loc = np.asarray(r_loc)
    old_l = np.asarray(old_l_loc).reshape(-1, 2); old_r = np.asarray(old_r_loc).reshape(-1, 2)
    seen_l = np.isin(l_loc[:, 0], old_l[:, 0]) | np.isin(l_loc[:, 1], old_l[:, 1])
    seen_r = np.isin(r_loc[:, 0], old_r[:, 0]) | np.isin(r_loc[:, 1], old_r[:, 1])
    return np.nonzero(~seen_l & ~seen_r)[0]


def create_landmarks_from_features(ops, features_l, features_r, P1, P2, pose, current_landmarks):
    """CreateLandmarksFromFeatures.m:1-21 with the per-point ``triangulate`` calls of its loop
    batched into one GPU call.  Every second feature (i = 1:2:n) is triangulated, kept when
    0 <= z <= 80 and moved to world coordinates with ``pose`` (4 x 4, premultiply convention);
    skipped rows stay zero exactly like the reference's pre-sized ``zeros`` array
    (``zeros(size(features_l, 2), 3)`` = 2 rows, grown by indexing)."""
    features_l = np.asarray(features_l).reshape(-1, 2); features_r = np.asarray(features_r).reshape(-1, 2)
    n = len(features_l)
    picks = np.arange(0, n, 2)                                   # i = 1:2:n (1-based)
    landmarks = np.zeros((2, 3))
    if n:
        coords = ops.triangulate(features_l[picks], features_r[picks], P1, P2)    # CreateLandmarksFromFeatures.m:7
        keep = ~((coords[:, 2] < 0) | (coords[:, 2] > 80))                        # :9-15
        if keep.any():
            A = np.asarray(pose, dtype=np.float64)
            # MATLAB grows the array (zero-filled) up to the last row actually written
            landmarks = np.zeros((max(2, int(picks[keep][-1]) + 1), 3))
            landmarks[picks[keep]] = coords[keep] @ A[:3, :3].T + A[:3, 3]        # :17 transformPointsForward
    cur = np.asarray(current_landmarks, dtype=np.float64).reshape(-1, 3)
    return np.vstack([cur, landmarks])                            # :20


class VisualOdometry:
    """State of the VO.m script: ``features`` (previous stereo-matched set), ``pose``, ``all_poses``;
    with ``view_3D`` also the ``landmarks`` map of VO.m:145-161."""

    def __init__(self, P1, P2, ops=None, view_3D=False):
        self.view_3D = view_3D
        self.landmarks = np.zeros((0, 3))
        self.p1 = np.asarray(P1, dtype=np.float64).reshape(3, 4)
        self.p2 = np.asarray(P2, dtype=np.float64).reshape(3, 4)
        # VO.m:35-38: intrinsics of the left camera from p1
        self.K4 = np.array([self.p1[0, 0], self.p1[1, 1], self.p1[0, 2], self.p1[1, 2]])
        self.ops = ops or CudaOps()
        self.features = None
        self.pose = np.eye(4)                     # VO.m:58
        self.all_poses = []                       # VO.m:59
        self.frame_index = 0
        self.log = []

    def step(self, lf, rf):
        """One iteration of `for i = 1:n_frames` (VO.m:64-231).  Returns rel_pose.A or None (i = 1)."""
        ops = self.ops
        l_desc, l_pos = ops.detect_and_extract(lf)                             # VO.m:79,83
        r_desc, r_pos = ops.detect_and_extract(rf)                             # VO.m:80,84
        matched = ops.matchFeatures(l_desc, r_desc)                            # VO.m:87
        rel = None
        rec = dict(n_l=len(l_desc), n_r=len(r_desc), k0=len(matched))
        if self.features is not None:                                          # VO.m:90
            cur = dict(l_desc=l_desc, l_pos=l_pos, r_desc=r_desc, r_pos=r_pos)
            cur, old, _, _, ks = find_remaining_points(ops, self.features, cur)  # VO.m:106
            old_pos = ops.triangulate(old["l_pos"], old["r_pos"], self.p1, self.p2)   # VO.m:114
            r = ops.estworldpose(cur["l_pos"].astype(np.float64), old_pos, self.K4, self.frame_index)  # VO.m:123
            rec.update(k1=ks[0], k2=ks[1], k3=ks[2], k4=ks[3], status=r["status"], inliers=r["n_inliers"])
            if r["status"] != 0:
                raise api.VoError(f"estworldpose failed at frame {self.frame_index} (status {r['status']})")
            rel = r["A"]
            self.pose = self.pose @ rel                                        # VO.m:130
            self.all_poses.append(self.pose.copy())                            # VO.m:133
            if self.view_3D:                                                   # VO.m:145-161
                ml, mr = l_pos[matched[:, 0]], r_pos[matched[:, 1]]
                idx = new_landmark_indices(ml, mr, old["l_pos"], old["r_pos"])
                self.landmarks = create_landmarks_from_features(ops, ml[idx], mr[idx], self.p1, self.p2, self.pose, self.landmarks)
        self.features = dict(l_desc=l_desc[matched[:, 0]], r_desc=r_desc[matched[:, 1]],   # VO.m:141-144 / 207-210
                             l_pos=l_pos[matched[:, 0]], r_pos=r_pos[matched[:, 1]])
        self.log.append(rec)
        self.frame_index += 1
        return rel


def run_frames(left, right, P1, P2, seed=0, first_frame=0, max_keypoints=8192, ctx=None, device_ptrs=None, col_major=False,
               Unique=False):
    """Batched device-resident loop body (vo_frames).  left/right: [n, rows, cols] uint8 host
    arrays (NumPy, or pinned torch CPU tensors via .numpy()).  With ``device_ptrs=(lptr, rptr,
    n, rows, cols)`` the images are already in HBM (vo_frames_dev) and left/right are ignored.
    ``col_major=True``: left/right are [n, cols, rows] arrays, i.e. the memory of MATLAB H x W x N stacks.
    Returns (rel_pose [n,4,4], status [n], counts [n,8])."""
    ctx = ctx or api.default_context()
    if device_ptrs is None:
        left = np.ascontiguousarray(left, dtype=np.uint8)
        right = np.ascontiguousarray(right, dtype=np.uint8)
        n, rows, cols = left.shape
        if col_major:
            rows, cols = cols, rows
    else:
        n, rows, cols = device_ptrs[2:]
    P1 = np.ascontiguousarray(P1, dtype=np.float64).reshape(12)
    P2 = np.ascontiguousarray(P2, dtype=np.float64).reshape(12)
    o = _lib.FramesOpts()
    o.sift.index_base = 1
    o.p3p.seed = seed
    o.p3p.adaptive = -1
    o.max_keypoints = max_keypoints
    o.first_frame = first_frame
    o.col_major = 1 if col_major else 0
    o.match.unique = 1 if Unique else 0          # matchFeatures(..., "Unique", true) in all five calls of the loop
    rel = np.zeros((n, 4, 4))
    status = np.zeros(n, dtype=np.int32)
    counts = np.zeros((n, 8), dtype=np.int32)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    if device_ptrs is None:
        check(_lib.lib().vo_frames(ctx.handle, p(left, C.c_uint8), p(right, C.c_uint8), n, rows, cols,
                                   p(P1, C.c_double), p(P2, C.c_double), C.byref(o), p(rel, C.c_double),
                                   p(status, C.c_int), p(counts, C.c_int)))
    else:
        check(_lib.lib().vo_frames_dev(ctx.handle, C.c_void_p(device_ptrs[0]), C.c_void_p(device_ptrs[1]),
                                       n, rows, cols, p(P1, C.c_double), p(P2, C.c_double), C.byref(o),
                                       p(rel, C.c_double), p(status, C.c_int), p(counts, C.c_int)))
    return rel, status, counts


def frames_landmarks(poses, cap=4096, ctx=None):
    """The landmark rows VO.m:145-161 appends for every frame of the batch that ``run_frames`` just processed on
    ``ctx`` (device-side selection, triangulation, depth filter and world transform: vo_frames_landmarks).
    poses: [n, 4, 4] world poses of the batch's frames (after the pose chain).  Returns a list of [rows_i, 3] arrays."""
    ctx = ctx or api.default_context()
    poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 16)
    n = len(poses)
    out = np.zeros((n, cap, 3))
    rows = np.zeros(n, dtype=np.int32)
    check(_lib.lib().vo_frames_landmarks(ctx.handle, poses.ctypes.data_as(C.POINTER(C.c_double)), n, cap,
                                         out.ctypes.data_as(C.POINTER(C.c_double)), rows.ctypes.data_as(C.POINTER(C.c_int))))
    return [out[i, :rows[i]].copy() for i in range(n)]


def chain_poses(rel_poses, start=None):
    """pose = pose * rel_pose (VO.m:130); returns the list of world poses (one per rel pose)."""
    pose = np.eye(4) if start is None else np.asarray(start, dtype=np.float64)
    out = []
    for a in rel_poses:
        pose = pose @ a
        out.append(pose.copy())
    return out


class FramePipeline:
    """Keeps ``depth`` batches of the frame loop in flight on one GPU: one context (own stream and device
    buffers) per slot, each driven by its own host thread (the C calls release the GIL).  While one
    batch is in its latency-bound matching / pose tail or waiting for its H2D copy, another fills the
    SMs; on B200 three slots give about +16 % frames/s over one batch at a time (bench.py --inflight).
    Batches are independent (each carries its one-frame halo), so results do not depend on ``depth``."""

    def __init__(self, depth=3, device=0):
        self.ctxs = [api.Context(device) for _ in range(max(1, depth))]

    def map(self, batches, P1, P2, seed=0):
        """batches: iterable of (left, right, first_frame) with [n, rows, cols] uint8 host arrays.
        Returns the list of (rel_pose, status, counts) in input order."""
        import threading
        batches = list(batches)
        out = [None] * len(batches)
        nxt = iter(range(len(batches)))
        lock = threading.Lock()
        err = []

        def worker(ctx):
            while not err:
                with lock:
                    i = next(nxt, None)
                if i is None:
                    return
                try:
                    l, r, first = batches[i]
                    out[i] = run_frames(l, r, P1, P2, seed=seed, first_frame=first, ctx=ctx)
                except Exception as e:           # surface the first failure in the caller's thread
                    err.append(e)
        th = [threading.Thread(target=worker, args=(c,)) for c in self.ctxs]
        [t.start() for t in th]
        [t.join() for t in th]
        if err:
            raise err[0]
        return out

    def close(self):
        for c in self.ctxs:
            c.close()
