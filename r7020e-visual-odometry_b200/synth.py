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
