"""Seeded synthetic KITTI-shaped stereo frames.

The reference reads kitti/00/image_0|image_1 PNGs (VO.m:16-17, 71-72); those images are not
shipped with the reference and there is no network, so the benchmarks and tests run on this
generator.  Texture recipe: band-limited noise (sum of Gaussian-blurred N(0,1) fields at sigma
{1,2,4,8,16} px with weights {0.2,0.3,0.4,0.5,0.6}) plus 150 random rectangles, rescaled to
uint8 0..255.  Two renderers:

* ``shift_stream``  -- throughput variant: one big texture, constant horizontal disparity and a
  constant per-frame shift (a fronto-parallel plane seen by a sideways-translating rig).
* ``plane_world``   -- geometric variant: textured planes in 3-D rendered through KITTI P0/P1
  (kitti/00/calib.txt) along a pose list, for trajectory tests.
"""
import numpy as np

KITTI_W, KITTI_H = 1241, 376
# kitti/00/calib.txt:1-2 of the reference (P0, P1), row-major 3x4
KITTI_P0 = np.array([[718.856, 0.0, 607.1928, 0.0],
                     [0.0, 718.856, 185.2157, 0.0],
                     [0.0, 0.0, 1.0, 0.0]])
KITTI_P1 = np.array([[718.856, 0.0, 607.1928, -386.1448],
                     [0.0, 718.856, 185.2157, 0.0],
                     [0.0, 0.0, 1.0, 0.0]])
KITTI_K4 = np.array([718.856, 718.856, 607.1928, 185.2157])
KITTI_BASELINE = 386.1448 / 718.856


def _gauss_blur_fft(field, sigma):
    h, w = field.shape
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.rfftfreq(w)[None, :]
    g = np.exp(-2.0 * (np.pi * sigma) ** 2 * (fx * fx + fy * fy))
    return np.fft.irfft2(np.fft.rfft2(field) * g, s=field.shape)


def texture(h, w, seed=0, n_rect=150):
    """Band-limited noise + rectangles, uint8."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((h, w))
    for sigma, wt in zip((1, 2, 4, 8, 16), (0.2, 0.3, 0.4, 0.5, 0.6)):
        f = _gauss_blur_fft(rng.standard_normal((h, w)), sigma)
        acc += wt * f / (f.std() + 1e-12)
    for _ in range(n_rect):
        rh, rw = rng.integers(6, max(7, h // 4)), rng.integers(6, max(7, w // 8))
        y0, x0 = rng.integers(0, max(1, h - rh)), rng.integers(0, max(1, w - rw))
        acc[y0:y0 + rh, x0:x0 + rw] += rng.normal(0.0, 1.0)
    lo, hi = np.percentile(acc, 0.5), np.percentile(acc, 99.5)
    out = np.clip((acc - lo) / (hi - lo), 0.0, 1.0) * 255.0
    return np.round(out).astype(np.uint8)


def shift_stream(n_frames, seed=20260, h=KITTI_H, w=KITTI_W, disparity=12, shift=3):
    """Returns (left, right) uint8 arrays of shape (n_frames, h, w).

    left_i(x) = T(x + i*shift), right_i(x) = T(x + i*shift + disparity), i.e. x_left - x_right =
    +disparity: a plane at depth Z = fx*b/disparity (32.2 m for KITTI calibration) seen by a rig
    moving +x by shift*Z/fx per frame.
    """
    big = texture(h, w + disparity + shift * n_frames + 8, seed)
    left = np.empty((n_frames, h, w), dtype=np.uint8)
    right = np.empty((n_frames, h, w), dtype=np.uint8)
    for i in range(n_frames):
        o = i * shift
        left[i] = big[:, o:o + w]
        right[i] = big[:, o + disparity:o + disparity + w]
    return left, right


def shift_stream_truth(disparity=12, shift=3):
    """Per-frame relative pose (4x4, camera i in camera i-1 coordinates) of shift_stream."""
    z = KITTI_P0[0, 0] * KITTI_BASELINE / disparity
    a = np.eye(4)
    a[0, 3] = shift * z / KITTI_P0[0, 0]
    return a, z


def plane_world(poses, seed=7, h=KITTI_H, w=KITTI_W, n_planes=9):
    """Render left/right views of textured planes for each 4x4 camera-to-world pose.

    Planes: a ground plane (y = 1.65 m below the camera), two side walls and fronto-parallel
    billboards ahead of the trajectory.  Rendering is a per-plane homography warp (cv2) with a
    painter's ordering by depth; occlusion is approximate, which is fine for odometry tests.
    """
    import cv2
    rng = np.random.default_rng(seed)
    poses = [np.asarray(p, dtype=np.float64).reshape(4, 4) for p in poses]
    centers = np.array([p[:3, 3] for p in poses])
    span = centers.max(axis=0) - centers.min(axis=0)
    zmax = centers[:, 2].max() + 60.0
    tex_px = 1024
    planes = []  # (origin, u_axis, v_axis, extent_u, extent_v, texture)

    def add(origin, ua, va, eu, ev, s):
        planes.append((np.asarray(origin, float), np.asarray(ua, float), np.asarray(va, float),
                       eu, ev, texture(tex_px, tex_px, s, n_rect=400)))

    xlo, xhi = centers[:, 0].min() - 12.0, centers[:, 0].max() + 12.0
    add([xlo, 1.65, -5.0], [1, 0, 0], [0, 0, 1], xhi - xlo, zmax + 5.0, seed * 100 + 1)      # ground
    add([xlo, -6.0, -5.0], [0, 0, 1], [0, 1, 0], zmax + 5.0, 7.65, seed * 100 + 2)           # left wall
    add([xhi, -6.0, -5.0], [0, 0, 1], [0, 1, 0], zmax + 5.0, 7.65, seed * 100 + 3)           # right wall
    add([xlo, -6.0, zmax], [1, 0, 0], [0, 1, 0], xhi - xlo, 7.65, seed * 100 + 4)            # far wall
    for k in range(max(0, n_planes - 4)):
        z = rng.uniform(8.0, zmax - 5.0)
        x = rng.uniform(xlo + 1.0, xhi - 5.0)
        add([x, rng.uniform(-3.0, 0.0), z], [1, 0, 0], [0, 1, 0], rng.uniform(2.0, 5.0),
            rng.uniform(1.5, 3.0), seed * 100 + 10 + k)
    _ = span
    K = KITTI_P0[:, :3]
    lefts, rights = [], []
    for T in poses:
        Rcw = T[:3, :3].T
        for cam, store in ((0, lefts), (1, rights)):
            tc = -Rcw @ T[:3, 3] - (np.array([KITTI_BASELINE, 0, 0]) if cam else 0.0)
            img = np.zeros((h, w), dtype=np.uint8)
            order = []
            for pl in planes:
                o, ua, va, eu, ev, tex = pl
                mid = Rcw @ (o + 0.5 * eu * ua + 0.5 * ev * va) + tc
                order.append((mid[2], pl))
            for _, (o, ua, va, eu, ev, tex) in sorted(order, key=lambda t: -t[0]):
                # texture pixel (s,t) -> world o + (s/tex_px*eu) ua + (t/tex_px*ev) va -> image
                M = np.stack([ua * eu / tex_px, va * ev / tex_px, o], axis=1)     # 3x3 world
                H = K @ (Rcw @ M + np.outer(tc, [0, 0, 1]))
                corners = np.array([[0, 0, 1], [tex_px, 0, 1], [tex_px, tex_px, 1], [0, tex_px, 1]], float).T
                cw = Rcw @ (M @ corners) + tc[:, None]
                if (cw[2] <= 0.3).any():
                    # clip the plane to the part in front of the camera by shrinking along v/u
                    # (cheap: skip planes that straddle the camera unless it is the ground/walls)
                    front = cw[2] > 0.3
                    if not front.any():
                        continue
                warped = cv2.warpPerspective(tex, H, (w, h), flags=cv2.INTER_LINEAR,
                                             borderMode=cv2.BORDER_CONSTANT, borderValue=0)
                mask = cv2.warpPerspective(np.full_like(tex, 255), H, (w, h), flags=cv2.INTER_NEAREST,
                                           borderMode=cv2.BORDER_CONSTANT, borderValue=0)
                # reject pixels whose pre-image lies behind the camera (homography sign flip)
                Hi = np.linalg.inv(H)
                ys, xs = np.mgrid[0:h, 0:w]
                wden = Hi[2, 0] * xs + Hi[2, 1] * ys + Hi[2, 2]
                s_ = (Hi[0, 0] * xs + Hi[0, 1] * ys + Hi[0, 2]) / wden
                t_ = (Hi[1, 0] * xs + Hi[1, 1] * ys + Hi[1, 2]) / wden
                pw = (M[:, 0:1] * s_.ravel() + M[:, 1:2] * t_.ravel() + M[:, 2:3])
                zc = (Rcw[2] @ pw + tc[2]).reshape(h, w)
                ok = (mask > 0) & (zc > 0.3)
                img[ok] = warped[ok]
            store.append(img)
    return np.stack(lefts), np.stack(rights)


# ------------------------------------------------------------------------------------ street world
# A geometrically exact stand-in for a KITTI sequence of any length (BASELINE.json configs[1]/[2], SURVEY
# 8(d) config 2/3): textured ground tiles, wall segments on both sides and billboards behind them are laid
# out along the given camera trajectory (the reference's own ground truth kitti/poses/00.txt), and every
# frame is rendered through P0/P1 (kitti/00/calib.txt) with a per-pixel depth buffer, so the stereo pair
# and consecutive frames are exact projections of one static 3-D world (real parallax, real occlusion,
# features entering and leaving the view, scale change -- what shift_stream does not have).
def _big_texture(h, w, seed):
    """The texture() recipe at a size where the FFT would be slow: Gaussian fields by separable blur."""
    import cv2
    rng = np.random.default_rng(seed)
    acc = np.zeros((h, w), dtype=np.float32)
    for sigma, wt in zip((1, 2, 4, 8, 16), (0.2, 0.3, 0.4, 0.5, 0.6)):
        f = cv2.GaussianBlur(rng.standard_normal((h, w), dtype=np.float32), (0, 0), sigma, borderType=cv2.BORDER_REFLECT)
        acc += np.float32(wt) * f / (f.std() + 1e-12)
    n_rect = 150 * (h * w) // (512 * 512)
    ys = rng.integers(0, h, n_rect); xs = rng.integers(0, w, n_rect)
    hs = rng.integers(6, 128, n_rect); ws = rng.integers(6, 96, n_rect)
    vs = rng.normal(0.0, 1.0, n_rect).astype(np.float32)
    for y0, x0, rh, rw, v in zip(ys, xs, hs, ws, vs):
        acc[y0:y0 + rh, x0:x0 + rw] += v
    lo, hi = np.percentile(acc[::4, ::4], 0.5), np.percentile(acc[::4, ::4], 99.5)
    return np.round(np.clip((acc - lo) / (hi - lo), 0.0, 1.0) * 255.0).astype(np.uint8)


class StreetWorld:
    """Static textured world along a trajectory.  ``poses``: [n, 4, 4] camera-to-world (KITTI convention:
    x right, y down, z forward).  ``render(i)`` -> (left, right) uint8 images of frame i."""
    TEX = 512

    def __init__(self, poses, seed=7, h=KITTI_H, w=KITTI_W, view_range=70.0):
        import cv2
        self.cv2 = cv2
        self.h, self.w, self.view_range = h, w, view_range
        self.poses = np.asarray(poses, dtype=np.float64).reshape(-1, 4, 4)
        rng = np.random.default_rng(seed)
        big = _big_texture(4096, 4096, seed * 100 + 1)
        self.planes = []    # (origin, u_axis*extent, v_axis*extent, mip pyramid)
        self._big, self._rng = big, rng
        c = self.poses[:, :3, 3]
        T = self.TEX

        def add(origin, eu, ev):
            y0 = int(rng.integers(0, big.shape[0] - T)); x0 = int(rng.integers(0, big.shape[1] - T))
            tex = np.ascontiguousarray(big[y0:y0 + T, x0:x0 + T])
            if rng.random() < 0.5:
                tex = np.ascontiguousarray(tex.T)
            mips = [tex]
            while mips[-1].shape[0] > 16:
                mips.append(cv2.pyrDown(mips[-1]))
            self.planes.append((np.asarray(origin, float), np.asarray(eu, float), np.asarray(ev, float), mips))

        # resample the path every `seg` metres
        seg = 8.0
        d = np.concatenate([[0.0], np.cumsum(np.linalg.norm(np.diff(c, axis=0), axis=1))])
        n_seg = max(2, int(np.ceil((d[-1] + 2 * view_range) / seg)))
        s = np.arange(n_seg + 1) * seg - 8.0
        ext = np.concatenate([[c[0] - (c[1] - c[0]) / max(d[1], 1e-9) * 16.0], c,
                              [c[-1] + (c[-1] - c[-2]) / max(d[-1] - d[-2], 1e-9) * (2 * view_range)]])
        dext = np.concatenate([[-16.0], d, [d[-1] + 2 * view_range]])
        pts = np.stack([np.interp(s, dext, ext[:, k]) for k in range(3)], axis=1)
        down = np.array([0.0, 1.0, 0.0])

        def clear_of_path(p, margin):
            return np.min(np.linalg.norm((c - p)[:, [0, 2]], axis=1)) > margin

        # ground: tiles of a global 12.5 m grid within 22 m of the path, at the height of the nearest path point
        cell = 12.5
        seen = set()
        for p in pts:
            for ix in range(int(np.floor((p[0] - 22) / cell)), int(np.floor((p[0] + 22) / cell)) + 1):
                for iz in range(int(np.floor((p[2] - 22) / cell)), int(np.floor((p[2] + 22) / cell)) + 1):
                    if (ix, iz) in seen:
                        continue
                    ctr = np.array([(ix + 0.5) * cell, 0.0, (iz + 0.5) * cell])
                    k = int(np.argmin(np.linalg.norm((pts - ctr)[:, [0, 2]], axis=1)))
                    if np.linalg.norm((pts[k] - ctr)[[0, 2]]) > 22 + cell:
                        continue
                    seen.add((ix, iz))
                    add([ix * cell, pts[k][1] + 1.65, iz * cell], [cell, 0, 0], [0, 0, cell])
        # walls and billboards along the path
        for a, b in zip(pts[:-1], pts[1:]):
            t = b - a
            L = np.linalg.norm(t)
            if L < 1e-6:
                continue
            t = t / L
            side = np.cross(down, t)
            side /= np.linalg.norm(side)
            for sgn in (-1.0, 1.0):
                off = rng.uniform(7.0, 9.0)
                o = a + sgn * off * side
                if clear_of_path(o + 0.5 * L * t, 6.0):
                    hgt = rng.uniform(4.5, 7.5)
                    add(o + down * (1.65 - hgt), t * L, down * hgt)                # wall segment, top-left origin
                if rng.random() < 0.6:                                              # billboard behind the wall line
                    off2 = rng.uniform(11.0, 18.0)
                    o2 = a + sgn * off2 * side + t * rng.uniform(0.0, L)
                    if clear_of_path(o2, 9.0):
                        hgt, wid = rng.uniform(6.0, 12.0), rng.uniform(4.0, 9.0)
                        yaw = rng.uniform(-0.6, 0.6)
                        u = np.cos(yaw) * t + np.sin(yaw) * side
                        add(o2 + down * (1.65 - hgt), u * wid, down * hgt)
        self.centers = np.array([o + 0.5 * eu + 0.5 * ev for o, eu, ev, _ in self.planes])
        self.K = KITTI_P0[:, :3]

    def render(self, i, depth=False):
        """(left, right) of frame i; with ``depth`` also the two depth buffers (z in metres, inf = nothing)."""
        cv2 = self.cv2
        T = self.poses[i]
        Rcw = T[:3, :3].T
        h, w, K = self.h, self.w, self.K
        near = np.nonzero(np.linalg.norm(self.centers - T[:3, 3], axis=1) < self.view_range)[0]
        out, zs = [], []
        xs = np.arange(w, dtype=np.float32)[None, :]
        ys = np.arange(h, dtype=np.float32)[:, None]
        for cam in (0, 1):
            tc = -Rcw @ T[:3, 3] - (np.array([KITTI_BASELINE, 0, 0]) if cam else 0.0)
            img = np.zeros((h, w), dtype=np.uint8)
            zbuf = np.full((h, w), np.inf, dtype=np.float32)
            for k in near:
                o, eu, ev, mips = self.planes[k]
                corners = np.stack([o, o + eu, o + eu + ev, o + ev], axis=1)       # 3 x 4 world
                cc = Rcw @ corners + tc[:, None]
                if (cc[2] <= 0.3).all():
                    continue
                # mip level: texels per pixel at the plane centre
                zc0 = max(float(cc[2].mean()), 0.5)
                texel = np.linalg.norm(eu) / self.TEX
                lvl = int(np.clip(np.floor(np.log2(max(zc0 / K[0, 0] / texel, 1e-9))) + 1, 0, len(mips) - 1))
                tex = mips[lvl]
                n = tex.shape[0]
                # texture pixel centre (s + 0.5, t + 0.5) / n  <->  world o + a*eu + b*ev
                M = np.stack([eu / n, ev / n, o + 0.5 * eu / n + 0.5 * ev / n], axis=1)
                Hm = K @ (Rcw @ M + np.outer(tc, [0, 0, 1]))                      # (s, t, 1) -> zc * (x, y, 1)
                if (cc[2] > 0.3).all():
                    pr = (K @ cc) / cc[2]
                    x0 = int(max(0, np.floor(pr[0].min()))); x1 = int(min(w, np.ceil(pr[0].max()) + 1))
                    y0 = int(max(0, np.floor(pr[1].min()))); y1 = int(min(h, np.ceil(pr[1].max()) + 1))
                    if x0 >= x1 or y0 >= y1:
                        continue
                else:
                    x0, x1, y0, y1 = 0, w, 0, h
                Hi = np.linalg.inv(Hm)
                Hi = Hi / np.abs(Hi).max()
                sx, sy = xs[:, x0:x1], ys[y0:y1]
                wd = (Hi[2, 0] * sx + Hi[2, 1] * sy + Hi[2, 2]).astype(np.float32)
                # Hi (x, y, 1) = (s, t, 1) / zc up to the common scale of Hi: recover zc from the forward map
                ss = (Hi[0, 0] * sx + Hi[0, 1] * sy + Hi[0, 2]).astype(np.float32)
                tt = (Hi[1, 0] * sx + Hi[1, 1] * sy + Hi[1, 2]).astype(np.float32)
                with np.errstate(divide="ignore", invalid="ignore"):
                    s_ = ss / wd; t_ = tt / wd
                    zc = (Hm[2, 0] * s_ + Hm[2, 1] * t_ + Hm[2, 2]).astype(np.float32)
                ok = (s_ >= 0) & (s_ <= n - 1) & (t_ >= 0) & (t_ <= n - 1) & (zc > 0.3) & (zc < zbuf[y0:y1, x0:x1])
                if not ok.any():
                    continue
                val = cv2.remap(tex, s_, t_, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
                img[y0:y1, x0:x1][ok] = val[ok]
                zbuf[y0:y1, x0:x1][ok] = zc[ok]
            out.append(img)
            zs.append(zbuf)
        if depth:
            return out[0], out[1], zs[0], zs[1]
        return out[0], out[1]


_WORLD = None


def _street_render(i):
    return _WORLD.render(i)


def street_sequence(poses, seed=7, h=KITTI_H, w=KITTI_W, workers=None, frames=None):
    """(left, right) uint8 [n, h, w] of a StreetWorld built on ``poses``, rendered on ``workers`` processes;
    ``frames``: render only the first so many poses (the world still spans all of them)."""
    global _WORLD
    import multiprocessing as mp
    import os
    _WORLD = StreetWorld(poses, seed=seed, h=h, w=w)
    n = len(_WORLD.poses) if frames is None else min(frames, len(_WORLD.poses))
    workers = workers or min(os.cpu_count() or 1, 32)
    if workers > 1 and n > 4:
        with mp.get_context("fork").Pool(workers) as pool:
            fr = pool.map(_street_render, range(n), chunksize=max(1, n // (4 * workers)))
    else:
        fr = [_WORLD.render(i) for i in range(n)]
    left = np.stack([f[0] for f in fr]); right = np.stack([f[1] for f in fr])
    return left, right


# ------------------------------------------------------------------------------------ descriptor sets
# inputs of the match sweep.
def sift_like_descriptors(n, seed, dim=128):
    """Integer-valued descriptors with OpenCV-SIFT statistics: g ~ |N(0,1)|^1.5, normalise to 512, clip at
    0.2*512, renormalise, round to 0..255 (float32)."""
    rng = np.random.default_rng(seed)
    g = np.abs(rng.standard_normal((n, dim), dtype=np.float32)) ** np.float32(1.5)
    g *= 512.0 / np.linalg.norm(g, axis=1, keepdims=True)
    g = np.minimum(g, np.float32(0.2 * 512.0))
    g *= 512.0 / np.linalg.norm(g, axis=1, keepdims=True)
    return np.clip(np.rint(g), 0, 255).astype(np.float32)


def descriptor_sets(kind, n1, n2, seed=1234):
    """(f1 [n1,128], f2 [n2,128]) float32.
    "integer": SIFT-like integer rows, half of f1 noisy copies of f2 rows (so matches exist);
    "ties":    on top of that 5 % exactly duplicated landmark rows, rows differing by +-1 in one bin and
               query rows that are exact copies (ties must resolve to the lowest index);
    "float":   general float32 unit vectors (what a caller holding unit-norm descriptors passes)."""
    rng = np.random.default_rng(seed)
    f2 = sift_like_descriptors(n2, seed + 1)
    f1 = sift_like_descriptors(n1, seed + 2)
    k = min(n1, n2) // 2
    src = rng.permutation(n2)[:k]
    dst = rng.permutation(n1)[:k]
    f1[dst] = np.clip(np.rint(f2[src] + rng.normal(0, 6.0, (k, 128)).astype(np.float32)), 0, 255)
    if kind == "integer":
        return f1, f2
    if kind == "ties":
        m = max(2, n2 // 20) // 2 * 2
        dup = rng.permutation(n2)[:m]
        f2[dup[: m // 2]] = f2[dup[m // 2:]]                          # exact duplicates
        near = rng.permutation(n2)[: m // 2]
        col = rng.integers(0, 128, len(near))
        f2[near, col] = np.clip(f2[near, col] + 1.0, 0, 255)          # +-1 in one bin
        q = rng.permutation(n1)[: max(1, n1 // 4)]
        f1[q] = f2[rng.integers(0, n2, len(q))]                       # exact copies of landmark rows
        f1[q[::3], 5] = np.clip(f1[q[::3], 5] + 1.0, 0, 255)
        return f1, f2
    if kind == "float":
        g2 = f2 + rng.random((n2, 128), dtype=np.float32)
        g2 /= np.linalg.norm(g2, axis=1, keepdims=True)
        g1 = f1 + rng.random((n1, 128), dtype=np.float32)
        g1 /= np.linalg.norm(g1, axis=1, keepdims=True)
        return g1.astype(np.float32), g2.astype(np.float32)
    raise ValueError(kind)
