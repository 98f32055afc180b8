#!/usr/bin/env python
"""bench.py -- KITTI-shaped stereo frames/s of the VO hot path on B200, with roofline evidence.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

Workload (BASELINE.json configs[1]): the per-frame body of VO.m (SIFT x2 -> stereo match ->
find_remaining_points (4 matches) -> triangulate -> P3P-MSAC) over 1241x376 stereo frames.  KITTI
images are not shipped with the reference and there is no network, so the frames are rendered: a static
textured 3-D street world seen through kitti/00/calib.txt's P0/P1 along the reference's own ground truth
kitti/poses/00.txt (synth.StreetWorld; exact stereo / temporal geometry, occlusion, features entering and
leaving the view -- data: "synthetic").  A *step* = one vo_frames call over B+1 consecutive frames (one
halo frame + B new frames -> B relative poses).

One JSON line on stdout (rank 0).  All five BASELINE.json configs appear in it, each with a parity flag:
  value / e2e    configs[1]/[2]: frames/s device-resident (vo_frames_dev) and through the C ABI with HOST
                 pinned frames (vo_frames: H2D of every frame, D2H of the poses inside the timed region)
  trajectory     configs[0]/[1]: KITTI devkit t_err / r_err and the reference's xz error (PlotOnMap.m:20) of
                 the GPU poses against the ground truth, and against the CPU oracle run on the same frames
  throughput_variant   the constant-shift stream of round 1 (synth.shift_stream), same timing
  roofline       dominant kernel of the step (live CUDA-event timers inside the library)
  match_gemm     configs[3]: tcgen05 match GEMM sweep 2048..65536 squared, integer / adversarial-tie /
                 general-float inputs, sampled rows checked bit-exact against the oracle at every size
  reloc          configs[4]: 131072 query rows per rank x 1M landmarks, row-sharded, all-gather of the
                 16-byte records, 4096-hypothesis P3P; sampled rows checked against the oracle
  e2e_dropin     the literal drop-in: the MEX gateways driven through the stand-in MATLAB host
  e2e_from_png   PNG files -> poses (decode on host threads overlapped with the GPU)
  cpu_baseline   the C oracle on ALL host cores over the first frames of the same sequence (rank 0, N=1)

--impl reference times the reference's CPU implementation of the path.  MATLAB is not available
offline, so this is the oracle port run with one process per host core.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# The driver parses stdout for ONE JSON line.  Libraries (NCCL prints its version banner) write to
# file descriptor 1 behind Python's back, so fd 1 is pointed at stderr for the whole run and the JSON
# line goes to a private duplicate of the original stdout.
_JSON_OUT = None


def _claim_stdout():
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


H, W = 376, 1241
METRIC = "kitti_stereo_frames_per_s_end_to_end"
UNIT = "frames/s"
SEQ_SEED = 7
# identical in both arms (the driver compares the arms' configs)
CONFIG = dict(workload="VO.m loop body (VO.m:64-232) on 1241x376 stereo frames of a rendered street world along "
                       "kitti/poses/00.txt through kitti/00/calib.txt P0/P1 (BASELINE.json configs[1]; KITTI images "
                       "absent offline)",
              rows=H, cols=W, halo_frames=1, sequence="synth.StreetWorld(seed=7), frames 0..n of kitti/poses/00.txt")


# Measured on B200 with tools/ubench_tmem.cu (profiles/r1_ubench_tmem.txt): back-to-back tcgen05.mma.kind::i8,
# cta_group::1, M = 128, N = 256, K = 32 with both operands in shared memory take 171.2 cycles (ideal 128), i.e.
# 2*128*256*32 / 171.2 op per cycle and SM x 148 SMs x 1.965 GHz.  The only measured peak of the pipe the match uses.
I8_MMA_RATE_TOPS = 2 * 128 * 256 * 32 / 171.2 * 148 * 1.965e9 / 1e12


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled in-process through NVML every 10 ms
    (no nvidia-smi child process competing for the host cores)."""
    BITS = dict(hw_slowdown=0x8, sw_power_cap=0x4, sw_thermal_slowdown=0x20, hw_thermal_slowdown=0x40)

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.mx, self.err = index, [], set(), None, None
        self._stop = threading.Event()
        self.t = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.replace(",", "").isdigit() else self.index
            self.hd = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.hd, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:   # noqa: BLE001
            self.err = str(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.hd, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.hd))
                for k, bit in self.BITS.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception as e:   # noqa: BLE001
                self.err = str(e)
                return
            self._stop.wait(0.01)

    def mark(self):
        return len(self.sm)

    def stop(self, lo=0, hi=None):
        if self.t is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvml unavailable: %s" % self.err])
        self._stop.set()
        self.t.join()
        sm = self.sm[lo:hi] or self.sm
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=self.mx, reasons=sorted(self.reasons),
                    samples=len(sm), sampler="nvml in-process, 10 ms")


# ------------------------------------------------------------------------------------ frames
def gt_poses(n):
    d = np.load(os.path.join(ROOT, "tests", "golden", "kitti00_reference_data.npz"))
    n = min(n, len(d["poses"]))
    gt = np.tile(np.eye(4), (n, 1, 1))
    gt[:, :3, :] = d["poses"][:n]
    return gt


def street_frames(n, rank=0, barrier=None):
    """(left, right, gt) of the first n frames of the street sequence.  Rendered once per box (rank 0, all host
    cores) and cached in the temp dir; the other ranks and the reference arm load the cache."""
    import tempfile
    world_poses = gt_poses(10 ** 9)          # the world is laid out along ALL poses of the fixture, whatever n is rendered
    gt = world_poses[:n]
    n = len(gt)
    path = os.path.join(tempfile.gettempdir(), f"vo_b200_street_s{SEQ_SEED}_w{len(world_poses)}_n{n}_{H}x{W}.npy")
    if rank == 0 and not os.path.exists(path):
        from vo_b200 import synth
        t0 = time.time()
        left, right = synth.street_sequence(world_poses, seed=SEQ_SEED, h=H, w=W, frames=n)
        np.save(path + ".tmp.npy", np.stack([left, right]))
        os.replace(path + ".tmp.npy", path)
        log(f"rendered {n} street frames in {time.time() - t0:.1f} s")
    if barrier is not None:
        barrier()
    a = np.load(path)
    return a[0], a[1], gt


# ------------------------------------------------------------------------------------ CPU legs
def _oracle_frame(args):
    """SIFT x2 + stereo match of one frame with the oracle (worker process)."""
    from oracle_ops import OracleOps
    l, r = args
    ops = OracleOps()
    ld, lp = ops.detect_and_extract(l)
    rd, rp = ops.detect_and_extract(r)
    m = ops.matchFeatures(ld, rd)
    return ld, lp, rd, rp, m


def _oracle_pair(args):
    """find_remaining_points + triangulate + P3P for one frame pair with the oracle -> (status, A, K4)."""
    from oracle_ops import OracleOps
    from vo_b200 import vo, synth
    prev, cur, idx = args
    ops = OracleOps(seed=1)
    ld, lp, rd, rp, m = prev
    old = dict(l_desc=ld[m[:, 0]], r_desc=rd[m[:, 1]], l_pos=lp[m[:, 0]], r_pos=rp[m[:, 1]])
    c = dict(l_desc=cur[0], l_pos=cur[1], r_desc=cur[2], r_pos=cur[3])
    c, o, _, _, ks = vo.find_remaining_points(ops, old, c)
    xyz = ops.triangulate(o["l_pos"], o["r_pos"], synth.KITTI_P0, synth.KITTI_P1)
    r = ops.estworldpose(c["l_pos"].astype(np.float64), xyz, synth.KITTI_K4, idx)
    return r["status"], r["A"], ks[3]


def _oracle_top2_chunk(args):
    """Oracle best-2 of a few query rows against one column chunk of the landmark set."""
    from oracle import oracle
    f1, f2, off = args
    j, s1, s2 = oracle.match_top2(f1, np.asarray(f2, dtype=np.float32))
    return j.astype(np.int64) + off, s1, s2


_POOL = None


def pool():
    """Worker processes for the oracle legs.  "spawn": the parent holds a CUDA context."""
    global _POOL
    if _POOL is None:
        import multiprocessing as mp
        _POOL = mp.get_context("spawn").Pool(os.cpu_count() or 1)
    return _POOL


def oracle_top2_rows(f1_rows, f2):
    """oracle.match_top2(f1_rows, f2) with the columns of f2 spread over the worker pool.  The score of a pair
    does not depend on the chunking, and the merge keeps the oracle's order (score, then lowest column)."""
    n2 = len(f2)
    nchunk = max(1, min(os.cpu_count() or 1, n2 // 4096))
    edges = np.linspace(0, n2, nchunk + 1).astype(int)
    parts = pool().map(_oracle_top2_chunk, [(f1_rows, f2[a:b], int(a)) for a, b in zip(edges[:-1], edges[1:])])
    j = np.stack([p[0] for p in parts] * 2, axis=1)                       # candidate columns: each chunk's best twice
    s = np.stack([p[1] for p in parts] + [p[2] for p in parts], axis=1)   # best of each chunk, then its runner-up
    n = len(f1_rows)
    j1 = np.zeros(n, dtype=np.uint32); s1 = np.zeros(n, dtype=np.float32); s2 = np.zeros(n, dtype=np.float32)
    for i in range(n):
        best = s[i, :nchunk]
        k = int(np.argmin(best))                    # first minimum = lowest chunk = lowest column on ties
        j1[i] = j[i, k]; s1[i] = best[k]
        rest = np.concatenate([best[:k], best[k + 1:], s[i, nchunk + k: nchunk + k + 1]])
        s2[i] = rest.min() if len(rest) else np.float32(np.inf)
    return j1, s1, s2


def oracle_keep(s1, s2, n2, match_threshold=1.0, max_ratio=0.6):
    """The acceptance tests of oracle/match.c (vo_oracle_match) applied to best-2 scores, in float32."""
    s1 = s1.astype(np.float32); s2 = s2.astype(np.float32)
    keep = s1 <= np.float32(match_threshold) * np.float32(0.04)
    if n2 >= 2:
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.where(s2 < np.float32(1e-6), np.float32(1.0), s1 / s2).astype(np.float32)
        keep &= ratio <= np.float32(max_ratio)
    return keep


def oracle_sequence(left, right, n):
    """The oracle over frames 0..n-1 on all host cores: (rel [n,4,4], status [n], seconds, tracked)."""
    t0 = time.perf_counter()
    fr = pool().map(_oracle_frame, [(left[i], right[i]) for i in range(n)], chunksize=1)
    res = pool().map(_oracle_pair, [(fr[i - 1], fr[i], i) for i in range(1, n)], chunksize=1)
    dt = time.perf_counter() - t0
    rel = np.tile(np.eye(4), (n, 1, 1))
    status = np.zeros(n, dtype=np.int32)
    for i, (st, A, _) in enumerate(res, start=1):
        status[i] = st
        if st == 0:
            rel[i] = A
    return rel, status, dt, [r[2] for r in res]


def cv2_info(left, right):
    """Informational: OpenCV's own SIFT / brute-force matcher on the same frames (SURVEY 8d "CPU path timed
    beside it (3)") -- the closest stand-in for what the toolbox (believed to wrap OpenCV) costs."""
    try:
        import cv2
    except Exception as e:   # noqa: BLE001
        return dict(error=str(e))
    out = dict(version=cv2.__version__)
    cores = os.cpu_count() or 1
    for nt, tag in ((1, "1_thread"), (cores, "all_threads")):
        cv2.setNumThreads(nt)
        s = cv2.SIFT_create()
        t0 = time.perf_counter()
        feats = [s.detectAndCompute(im, None) for im in (left[0], right[0], left[1], right[1])]
        t_sift = (time.perf_counter() - t0) / 4
        bf = cv2.BFMatcher(cv2.NORM_L2)
        t0 = time.perf_counter()
        bf.knnMatch(feats[0][1], feats[1][1], k=2)
        t_match = time.perf_counter() - t0
        # VO.m body: 2 SIFT + 5 matches (the 4 temporal ones are smaller: counted as 2 full ones)
        out[tag] = dict(threads=nt, sift_ms_per_image=1e3 * t_sift, bf_knn_match_ms=1e3 * t_match,
                        keypoints=len(feats[0][0]), est_frames_per_s=1.0 / (2 * t_sift + 3 * t_match))
    cv2.setNumThreads(cores)
    return out


def trajectory_errors(rel, gt, n):
    """KITTI devkit errors (Appendix A.5) and the reference's xz error (PlotOnMap.m:20) of chained poses."""
    from vo_b200 import vo, kitti_eval
    est = np.array([np.eye(4)] + vo.chain_poses(rel[1:n]))
    est = gt[0] @ est
    t, r, k = kitti_eval.kitti_errors(est, gt[:n])
    xz = kitti_eval.xz_error(est, gt[:n])
    return dict(frames=int(n), t_err_pct=100.0 * t, r_err_deg_per_m=float(np.degrees(r)), segments=int(k),
                lengths_m=[L for L in kitti_eval.LENGTHS if L < kitti_eval._distances(gt[:n])[-1]],
                xz_err_final_m=float(xz[-1]), xz_err_max_m=float(xz.max()))


def run_reference(args, rank, world):
    """--impl reference: the oracle port on all host cores (one process per core), rank 0 only.  Every step
    processes at least one frame per core, so the pool is fully subscribed."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # frames per step: a multiple of the core count (incl. the halo frame) sized so that the whole run stays
    # within a few minutes (one round of `cores` frames takes about 2.2 s of wall time on the GPU box's host)
    rounds = max(1, int(150.0 / (2.2 * max(args.steps + args.warmup, 1))))
    n = min(cores * rounds, 385) - 1
    left, right, gt = street_frames(n + 1)
    log(f"reference arm: {n} frames (+1 halo) per step on {cores} processes")

    def step():
        return oracle_sequence(left, right, n + 1)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rel, status, _, _ = step()
    dt = time.perf_counter() - t0
    v = args.steps * n / dt
    line = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32 (SIFT, match) / f64 (triangulate, P3P)", data="synthetic", config=CONFIG,
                frames_per_step=n, status_ok=int((status[1:] == 0).sum()),
                note="reference = CPU oracle port (oracle/*.c), one process per core; MATLAB + Computer Vision Toolbox "
                     "not installed (BASELINE.md)",
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind="port",
                                  sample=f"{n} frames (+1 halo) per step, one process per core, {cores} cores"),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


# ------------------------------------------------------------------------------------ GPU legs
def match_gemm_leg(ctx, torch, pk, check=True):
    """BASELINE.json configs[3]: n x n x 128 for n = 2048..65536 on descriptors resident in HBM.
    Three input kinds (synth.descriptor_sets): "integer" at every size, "ties" and "float" at three sizes.
    Two consumers of the same kernel are timed: "match" = matchFeatures (vo_match_dev, default options) and
    "top2" = exact best-2 of every row (vo_match_top2_dev).  time = the GEMM + top-k kernel (CUDA events inside
    the library).  Parity: 256 sampled query rows of each run are compared bit for bit (index, score bits, kept
    pairs) with the oracle."""
    import ctypes as C
    from vo_b200 import _lib, synth
    L = _lib.lib()
    stream = torch.cuda.ExternalStream(ctx.stream)
    out = []
    plan = [("integer", (2048, 4096, 8192, 16384, 32768, 65536)), ("ties", (2048, 16384, 65536)),
            ("float", (2048, 16384, 32768))]
    all_ok = True
    for kind, sizes in plan:
        for n in sizes:
            a, b = synth.descriptor_sets(kind, n, n, seed=1234)
            f1 = torch.from_numpy(a).cuda(); f2 = torch.from_numpy(b).cuda()
            j1 = torch.empty(n, dtype=torch.int32, device="cuda"); s1 = torch.empty(n, device="cuda"); s2 = torch.empty(n, device="cuda")
            i1 = torch.empty(n, dtype=torch.int32, device="cuda"); i2 = torch.empty(n, dtype=torch.int32, device="cuda")
            mt = torch.empty(n, device="cuda"); npairs = torch.zeros(1, dtype=torch.int32, device="cuda")

            def call_top2():
                _lib.check(L.vo_match_top2_dev(ctx.handle, C.c_void_p(f1.data_ptr()), n, C.c_void_p(f2.data_ptr()), n, 128,
                                               C.c_void_p(j1.data_ptr()), C.c_void_p(s1.data_ptr()), C.c_void_p(s2.data_ptr()),
                                               C.c_void_p(ctx.stream)))

            def call_match():
                _lib.check(L.vo_match_dev(ctx.handle, C.c_void_p(f1.data_ptr()), n, C.c_void_p(f2.data_ptr()), n, 128, None,
                                          C.c_void_p(i1.data_ptr()), C.c_void_p(i2.data_ptr()), C.c_void_p(mt.data_ptr()),
                                          C.c_void_p(npairs.data_ptr()), C.c_void_p(ctx.stream)))
            rec = dict(kind=kind, n1=n, n2=n, dim=128)
            for mode, call in (("match", call_match), ("top2", call_top2)):
                torch.cuda.synchronize()
                for _ in range(3):
                    call()
                ctx.sync()
                ctx.profile_enable(True)
                reps = 10
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(reps):
                    call()
                e1.record(stream)
                ctx.sync()
                prof = ctx.profile()
                ctx.profile_enable(False)
                g = prof["match_gemm_topk"]
                flops = 2.0 * n * n * 128
                t_kernel = g["ms"] / g["launches"] * 1e-3
                tf = flops / t_kernel / 1e12
                t_call = e0.elapsed_time(e1) / reps
                d = dict(kernel_ms=1e3 * t_kernel, call_ms=t_call, tflops=tf, call_tflops=flops / (t_call * 1e-3) / 1e12,
                         frac_of_burst_peak=tf / pk["tf_burst"], frac_of_sustained_peak=tf / pk["tf_sust"],
                         frac_of_2x_burst_peak=tf / (2 * pk["tf_burst"]), frac_of_measured_i8_mma_rate=tf / I8_MMA_RATE_TOPS)
                if mode == "match":
                    d["pairs"] = int(npairs.item())
                    rec.update(d)
                rec[mode] = d
            if check:
                rows = np.sort(np.random.default_rng(n).permutation(n)[:256])
                oj, os1, os2 = oracle_top2_rows(a[rows], b)
                gj, gs1, gs2 = j1.cpu().numpy().view(np.uint32)[rows], s1.cpu().numpy()[rows], s2.cpu().numpy()[rows]
                top2_ok = bool(np.array_equal(gj, oj) and np.array_equal(gs1.view(np.uint32), os1.view(np.uint32))
                               and np.array_equal(gs2.view(np.uint32), os2.view(np.uint32)))
                keep = oracle_keep(os1, os2, n)
                p = int(npairs.item())
                gi1 = i1.cpu().numpy().view(np.uint32)[:p]; gi2 = i2.cpu().numpy().view(np.uint32)[:p]; gm = mt.cpu().numpy()[:p]
                sel = np.isin(gi1, rows)
                match_ok = bool(np.array_equal(gi1[sel], rows[keep].astype(np.uint32)) and np.array_equal(gi2[sel], oj[keep])
                                and np.array_equal(gm[sel].view(np.uint32), os1[keep].view(np.uint32)))
                rec["parity"] = dict(rows_checked=len(rows), top2_bit_exact=top2_ok, match_bit_exact=match_ok,
                                     kept_in_sample=int(keep.sum()))
                all_ok = all_ok and top2_ok and match_ok
            out.append(rec)
            del f1, f2
    ints = [d for d in out if d["kind"] == "integer"]
    best = max(ints, key=lambda d: d["tflops"])
    head = [d for d in ints if d["n1"] == 32768][0]
    return dict(head, sweep=out, parity_all_bit_exact=all_ok if check else None, best_tflops=best["tflops"],
                best_frac_of_burst_peak=best["frac_of_burst_peak"], peak_tflops_burst=pk["tf_burst"],
                peak_tflops_sustained=pk["tf_sust"], peak_source=pk["src"], measured_i8_mma_rate_tops=I8_MMA_RATE_TOPS,
                path="integer / ties: u8 x u8 -> s32 tcgen05.mma.kind::i8 (exact integer dot, K = 128), fused integer-prefilter "
                     "top-3 epilogue, no C written; float: split-bf16 kind::f16 GEMM + exact FP32 re-rank; ops counted as "
                     "2*N1*N2*128; peaks are the measured bf16 cuBLAS figures (the int8 pipe's nominal peak is 2x bf16: "
                     "frac_of_2x_burst_peak)")


def run_ours(args, rank, world, local_rank):
    import torch
    import vo_b200
    from vo_b200 import synth, vo
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libvo_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # `--inflight` batches are kept in flight: one context (own stream, own device buffers) per batch slot,
    # each driven by its own host thread (the C calls release the GIL).  While one batch is in its
    # latency-bound matching / pose tail, or waiting for its H2D copy, another one fills the SMs.
    n_ctx = max(1, args.inflight)
    ctxs = [vo_b200.Context(local_rank) for _ in range(n_ctx)]
    ctx = ctxs[0]
    pk = peaks()
    B = args.batch
    P0, P1 = synth.KITTI_P0, synth.KITTI_P1

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n_batches = max(1, min(12, 399 // B))
    n_seq = n_batches * B + 1
    sleft, sright, gt = street_frames(n_seq, rank, barrier if dist is not None else None)

    def batches_of(left, right, nb):
        """[nb, B+1, H, W] pinned views: batch b = frames b*B .. b*B+B (one-frame halo)."""
        L = torch.empty((nb, B + 1, H, W), dtype=torch.uint8).pin_memory()
        R = torch.empty((nb, B + 1, H, W), dtype=torch.uint8).pin_memory()
        for b in range(nb):
            L[b] = torch.from_numpy(left[b * B: b * B + B + 1]); R[b] = torch.from_numpy(right[b * B: b * B + B + 1])
        return L, R
    left, right = batches_of(sleft, sright, n_batches)
    dleft, dright = left.cuda(), right.cuda()
    streams = [torch.cuda.ExternalStream(c.stream) for c in ctxs]
    rot = rank % n_batches      # ranks start at different batches of the sequence

    def make_steps(left, right, dleft, dright, nb):
        def step_dev(i, c):
            b = (i + rot) % nb
            return vo.run_frames(None, None, P0, P1, seed=1, first_frame=b * B, ctx=c,
                                 device_ptrs=(dleft[b].data_ptr(), dright[b].data_ptr(), B + 1, H, W))

        def step_host(i, c):
            b = (i + rot) % nb
            return vo.run_frames(left[b].numpy(), right[b].numpy(), P0, P1, seed=1, first_frame=b * B, ctx=c)
        return step_dev, step_host
    step_dev, step_host = make_steps(left, right, dleft, dright, n_batches)

    def timed(step_fn, profile, use, sampler=None):
        """Exactly args.steps steps after args.warmup warm-up steps, spread round-robin over the first
        `use` contexts; device time from CUDA events on the launching streams, max over streams and ranks."""
        for i in range(args.warmup * use):
            step_fn(i, ctxs[i % use])
        barrier()
        if profile:
            ctx.profile_enable(True)
        launches0 = sum(c.kernel_launches() for c in ctxs)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(use)]
        cpu0 = time.process_time(); wall0 = time.perf_counter()
        m0 = sampler.mark() if sampler else 0
        e0.record(streams[0])
        results = [None] * args.steps

        nxt = iter(range(args.steps))
        lock = threading.Lock()

        def worker(k):   # batch slots pull the next step from a shared queue
            while True:
                with lock:
                    i = next(nxt, None)
                if i is None:
                    break
                results[i] = step_fn(args.warmup * use + i, ctxs[k])
            e1[k].record(streams[k])
        if use == 1:
            worker(0)
        else:
            th = [threading.Thread(target=worker, args=(k,)) for k in range(use)]
            [t.start() for t in th]
            [t.join() for t in th]
        barrier()
        m1 = sampler.mark() if sampler else 0
        host = dict(cpu_s=time.process_time() - cpu0, wall_s=time.perf_counter() - wall0)
        ms = max(e0.elapsed_time(e) for e in e1)
        prof = ctx.profile() if profile else None
        if profile:
            ctx.profile_enable(False)
        per_rank = [ms]
        if dist is not None:
            t = torch.tensor([ms, host["cpu_s"]], device="cuda", dtype=torch.float64)
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            per_rank = [float(x[0]) for x in allt]
            host["cpu_s_per_rank"] = [round(float(x[1]), 4) for x in allt]
            ms = max(per_rank)
        counts = [r[2] for r in results]
        return dict(ms=ms, prof=prof, launches=sum(c.kernel_launches() for c in ctxs) - launches0, counts=counts,
                    per_rank_ms=per_rank, host=host, marks=(m0, m1))

    sampler = ClockSampler(local_rank)
    sampler.start()
    t_dev = timed(step_dev, profile=False, use=n_ctx, sampler=sampler)
    clocks = sampler.stop(*t_dev["marks"])
    t_host = timed(step_host, profile=False, use=n_ctx)
    graph_states = [c.frames_graph_state() for c in ctxs]   # 2 = the timed passes replayed a captured CUDA graph (VO_FRAMES_GRAPH=1)
    # roofline pass: the same steps, one batch at a time on one stream, with the library's per-stage CUDA
    # events switched on (they cost a few % of a step, and overlapping batches would smear the stages)
    t_prof = timed(step_dev, profile=True, use=1)
    prof, counts = t_prof["prof"], t_prof["counts"]
    frames = args.steps * B * world
    value = frames / (t_dev["ms"] * 1e-3)
    e2e = frames / (t_host["ms"] * 1e-3)

    # throughput variant of round 1: constant-shift stream (every keypoint tracks, ~100 % inliers), same timing
    tv = None
    if not args.quick:
        nb2 = 4
        l2 = np.empty((nb2 * B + 1, H, W), dtype=np.uint8); r2 = np.empty_like(l2)
        for b in range(nb2):
            l, r = synth.shift_stream(B + 1, seed=20260 + 7919 * rank + 1000 * b, h=H, w=W)
            l2[b * B: b * B + B + 1] = l; r2[b * B: b * B + B + 1] = r
        left2, right2 = batches_of(l2, r2, nb2)
        dl2, dr2 = left2.cuda(), right2.cuda()
        sd2, sh2 = make_steps(left2, right2, dl2, dr2, nb2)
        a = timed(sd2, profile=False, use=n_ctx)
        b_ = timed(sh2, profile=False, use=n_ctx)
        cnt2 = np.concatenate(a["counts"], axis=0).astype(np.float64)
        tv = dict(workload="synth.shift_stream: one texture, constant 12 px disparity, 3 px/frame shift (round-1 headline)",
                  value=frames / (a["ms"] * 1e-3), e2e=frames / (b_["ms"] * 1e-3), unit=UNIT, ms_per_step=a["ms"] / args.steps,
                  keypoints_per_image=float(cnt2[:, :2].mean()), tracked_per_frame=float(cnt2[:, 6].mean()))
        del dl2, dr2, left2, right2

    # trajectory of the whole rendered sequence through the batched loop (rank 0 reports it)
    rel_all = np.tile(np.eye(4), (n_seq, 1, 1)); status_all = np.zeros(n_seq, dtype=np.int32)
    cnt_all = np.zeros((n_seq, 8), dtype=np.int32)
    for b in range(n_batches):
        r_, s_, c_ = vo.run_frames(None, None, P0, P1, seed=1, first_frame=b * B, ctx=ctx,
                                   device_ptrs=(dleft[b].data_ptr(), dright[b].data_ptr(), B + 1, H, W))
        rel_all[b * B + 1: b * B + B + 1] = r_[1:]; status_all[b * B + 1: b * B + B + 1] = s_[1:]
        cnt_all[b * B + 1: b * B + B + 1] = c_[1:]
    if rank != 0:
        if dist is not None:
            if not args.quick:
                reloc_leg(args, torch, dist, ctx, rank, world, pk)
                png_leg_all_ranks(torch, dist, sleft, sright, B, rank, world)
            dist.destroy_process_group()
        return
    trajectory = dict(gpu=trajectory_errors(rel_all, gt, n_seq), status_ok=int((status_all[1:] == 0).sum()),
                      tracked_min=int(cnt_all[1:, 6].min()), tracked_median=float(np.median(cnt_all[1:, 6])),
                      inliers_median=float(np.median(cnt_all[1:, 7])))

    # dominant kernel of the step and its roofline.  Launches of the same kernel are merged (the TMA blur
    # kernel runs once per octave and layer), so "dominant" is by kernel, not by launch site.
    cnt = np.concatenate(counts, axis=0).astype(np.float64)
    KERNEL_OF = {"sift_blur_dog_tma_oct0": "sift_blur_tma_kernel", "sift_blur_dog_tma_oct1+": "sift_blur_tma_kernel"}
    total_ms = sum(v["ms"] for v in prof.values())
    shares = {k: round(v["ms"] / total_ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    kern = {}
    for k, v in prof.items():
        kk = kern.setdefault(KERNEL_OF.get(k, k), dict(ms=0.0, launches=0, bytes=0.0, flops=0.0))
        for f in ("ms", "launches", "bytes", "flops"):
            kk[f] += v[f]
    # algorithmic flops of the five matches per step from the observed sizes: 2*N1*N2*128
    # (M0: N_L x N_R, M1: N_L x K0', M2: N_R x K1, M3: K1 x K2, M4: K3 x K2; primes = previous frame)
    k0_prev = np.concatenate([[0.0], cnt[:-1, 2]])
    pair = (cnt[:, 3] + cnt[:, 4] + cnt[:, 5] + cnt[:, 6]) > 0
    mm = (cnt[:, 0] * cnt[:, 1]).sum() + (pair * (cnt[:, 0] * k0_prev + cnt[:, 1] * cnt[:, 3] + cnt[:, 3] * cnt[:, 4]
                                                  + cnt[:, 5] * cnt[:, 4])).sum()
    if "match_gemm_topk" in kern:
        kern["match_gemm_topk"]["flops"] = 2.0 * mm * 128
    traffic = {}
    for name in ("r2_traffic.json", "r1_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_image", {})
            break

    def roof(name):
        d = kern[name]
        r = dict(kernel=name, launches=int(d["launches"]), avg_launch_ms=d["ms"] / max(d["launches"], 1),
                 share_of_step=round(d["ms"] / total_ms, 4), traffic=None)
        if d["bytes"] > 0:
            ach = d["bytes"] / (d["ms"] * 1e-3) / 1e9
            r.update(bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                     algorithmic_bytes_per_launch=d["bytes"] / d["launches"], peak_source=pk["src"])
            if name in traffic:   # ncu dram__bytes_read+write per image of 1241x376, scaled to this launch mix
                r["traffic"] = traffic[name] * (B + 1) * 2 * args.steps / d["launches"]
        elif d["flops"] > 0:
            ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
            r.update(bound="tensor", achieved=ach, peak=pk["tf_sust"], unit="TFLOP/s", frac=ach / pk["tf_sust"],
                     algorithmic_flops_per_launch=d["flops"] / d["launches"], peak_source=pk["src"] + " (sustained)")
        else:
            r.update(bound="latency", achieved=None, peak=None, unit=None, frac=None)
        return r
    dom = max(kern, key=lambda k: kern[k]["ms"])
    roofline = roof(dom)
    roofline_all = [roof(k) for k in sorted(kern, key=lambda k: -kern[k]["ms"])]
    # the pyramid stage as a whole against its COMPULSORY traffic (109.95 MB per 1241x376 image)
    pyr = [k for k in kern if k.startswith(("sift_base", "sift_blur", "sift_downsample", "sift_extrema", "sift_pyramid"))]
    pyr_ms = sum(kern[k]["ms"] for k in pyr)
    if pyr_ms > 0:
        S = sum((2 * H >> o) * (2 * W >> o) for o in range(9))
        comp = (H * W + 11 * S * 4) * (B + 1) * 2 * args.steps
        pyramid = dict(kernels=sorted(pyr), ms_per_step=pyr_ms / args.steps, compulsory_bytes_per_image=H * W + 11 * S * 4,
                       achieved_gbs=comp / (pyr_ms * 1e-3) / 1e9, frac_of_hbm_peak=comp / (pyr_ms * 1e-3) / 1e9 / pk["hbm"])
    else:
        pyramid = None
    gemm_ms = prof.get("match_gemm_topk", dict(ms=0))["ms"]
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=t_dev["ms"] / args.steps, ms_per_step_profiled_serial=t_prof["ms"] / args.steps,
                higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32 (SIFT) / u8->s32 exact-integer tensor-core GEMM (match) / f64 (triangulate, P3P)",
                data="synthetic", config=CONFIG,
                run=dict(frames_per_step_per_gpu=B, distinct_batches=n_batches, sequence_frames=n_seq,
                         l2="per-step working set ~%.1f GB of pyramids >> 126 MB L2; %d distinct input batches"
                            % ((B + 1) * 2 * 110e6 / 1e9, n_batches),
                         batches_in_flight=n_ctx, frames_graph_state=graph_states,
                         timing="value / e2e: exactly `steps` steps with `batches_in_flight` batches in flight (one stream "
                                "and host thread per batch slot), CUDA events on the launching streams, max over ranks; "
                                "roofline / stage_share / ms_per_step_profiled_serial: the same steps run one batch at a time "
                                "with per-stage CUDA events on",
                         parallelism=f"frame-sharded x{world}, no data-path collective"),
                e2e=dict(value=e2e, unit=UNIT, h2d_bytes_per_step=int((B + 1) * 2 * H * W),
                         d2h_bytes_per_step=int((B + 1) * (16 * 8 + 4 + 8 * 4)), ms_per_step=t_host["ms"] / args.steps),
                gpu_launches=int(t_dev["launches"]), clocks=clocks, roofline=roofline, roofline_all=roofline_all,
                pyramid_stage=pyramid, stage_share=shares,
                per_rank_ms=dict(value=[round(x, 3) for x in t_dev["per_rank_ms"]], e2e=[round(x, 3) for x in t_host["per_rank_ms"]]),
                host=dict(value=t_dev["host"], e2e=t_host["host"], cores=os.cpu_count()),
                keypoints_per_image=float(cnt[:, :2].mean()), tracked_per_frame=float(cnt[:, 6].mean()),
                match_gflop_per_step=2.0 * mm * 128 / 1e9 / args.steps, match_gemm_ms_per_step=gemm_ms / args.steps,
                trajectory=trajectory, throughput_variant=tv)
    side = []
    if world == 1 and not args.quick:
        side += [("match_gemm", lambda: match_gemm_leg(ctx, torch, pk, check=not args.no_cpu)),
                 ("e2e_dropin", lambda: dropin_leg(sleft, sright, gt)),
                 ("e2e_from_png", lambda: png_leg(sleft, sright, B))]
    if not args.quick:
        side += [("reloc", lambda: reloc_leg(args, torch, dist, ctx, rank, world, pk))]
        if world > 1:
            side += [("e2e_from_png", lambda: png_leg_all_ranks(torch, dist, sleft, sright, B, rank, world))]
    if world == 1 and not args.no_cpu and not args.quick:
        side += [("cpu_baseline", lambda: cpu_leg(sleft, sright, gt, rel_all, status_all, trajectory))]
    for name, fn in side:
        t0 = time.time()
        try:
            line[name] = fn()
        except Exception as e:  # keep the headline line even if a side leg fails  # noqa: BLE001
            import traceback
            traceback.print_exc()
            line[name] = dict(error=f"{type(e).__name__}: {e}")
        log(f"{name}: {time.time() - t0:.1f} s")
    pz = [line["trajectory"].get("parity_vs_oracle"), (line.get("match_gemm") or {}).get("parity_all_bit_exact"),
          (line.get("reloc") or {}).get("parity_rows_bit_exact"), (line.get("e2e_dropin") or {}).get("equals_vo_frames")]
    line["parity"] = dict(trajectory_vs_oracle=pz[0], match_sweep_bit_exact=pz[1], reloc_rows_bit_exact=pz[2], dropin_equals_batched=pz[3],
                          png_paths_equal_vo_frames=(line.get("e2e_from_png") or {}).get("equals_vo_frames"))
    emit(line)
    if _POOL is not None:
        _POOL.terminate()
    if dist is not None:
        dist.destroy_process_group()


def cpu_leg(sleft, sright, gt, rel_gpu, status_gpu, trajectory):
    """The oracle on every host core over the first frames of the sequence: cpu_baseline, and the oracle's
    trajectory next to the GPU's on the same frames (north star: t_err / r_err within 2 %)."""
    cores = os.cpu_count() or 1
    n = min(len(sleft), 225)
    rel_o, status_o, dt, tracked = oracle_sequence(sleft, sright, n)
    to = trajectory_errors(rel_o, gt, n)
    tg = trajectory_errors(rel_gpu, gt, n)
    dmax = float(np.abs(rel_o[:n] - rel_gpu[:n]).max())
    same_status = bool(np.array_equal(status_o[:n], status_gpu[:n]))
    rt = abs(tg["t_err_pct"] - to["t_err_pct"]) / max(to["t_err_pct"], 1e-12)
    rr = abs(tg["r_err_deg_per_m"] - to["r_err_deg_per_m"]) / max(to["r_err_deg_per_m"], 1e-12)
    trajectory.update(oracle=to, gpu_on_oracle_frames=tg, rel_pose_max_abs_diff=dmax, status_equal=same_status,
                      t_err_rel_diff=rt, r_err_rel_diff=rr, parity_vs_oracle=bool(rt <= 0.02 and rr <= 0.02 and same_status))
    out = dict(value=(n - 1) / dt, unit=UNIT, cores=cores, kind="port",
               sample=f"frames 0..{n - 1} of the same sequence ({2 * n} SIFT, {n} stereo matches, {n - 1} tracked pairs + P3P) "
                      f"in {dt:.1f} s on {cores} processes (one per core), C oracle (MATLAB + Computer Vision Toolbox not "
                      "installed: BASELINE.md)",
               per_core=(n - 1) / dt / cores, tracked_median=float(np.median(tracked)))
    out["cv2"] = cv2_info(sleft, sright)
    return out


def dropin_leg(sleft, sright, gt):
    """The literal drop-in: MATLAB-side calls through the .mexa64 gateways, driven by the stand-in
    MATLAB host (csrc/mex/mexshim.cpp).  Two forms: the six-call loop of VO.m (one gateway call per toolbox call),
    and the batched gateway vo_frames_mex (image stacks in, poses out)."""
    from vo_b200 import mexhost
    return mexhost.bench_dropin(sleft, sright, n_percall=24, n_batched=129, batch=32)


def png_leg(sleft, sright, B):
    """PNG files -> poses.  The frames are written as 8-bit gray PNGs (OpenCV / libpng: adaptive row
    filters, dynamic Huffman blocks), then read back two ways: "host" = io.run_sequence (the library's host decoder on
    every core decodes batch k+1 while vo_frames runs batch k, one batch at a time) and "device" =
    io.run_sequence_device (the files go to the GPU as they are; inflate + un-filter kernels, one warp per image;
    four / eight batches in flight on their own contexts, the host only reads files)."""
    import shutil
    import tempfile
    import cv2
    import torch
    import vo_b200
    from vo_b200 import io, synth, vo
    n = min(len(sleft), 12 * B + 1)
    d = tempfile.mkdtemp(prefix="vo_b200_png_")
    lf, rf = [], []
    nbytes = 0
    for i in range(n):
        for name, arr, lst in (("image_0", sleft, lf), ("image_1", sright, rf)):
            os.makedirs(os.path.join(d, name), exist_ok=True)
            p = os.path.join(d, name, f"{i:06d}.png")
            cv2.imwrite(p, arr[i]); lst.append(p); nbytes += os.path.getsize(p)
    out = dict(unit=UNIT, frames=n - 1, png_bytes_per_image=nbytes / (2 * n), host_cores=os.cpu_count())
    try:
        ref = vo.run_frames(sleft[:B + 1], sright[:B + 1], synth.KITTI_P0, synth.KITTI_P1, seed=1)
        io.run_sequence(lf[:B + 1], rf[:B + 1], synth.KITTI_P0, synth.KITTI_P1, batch=B, seed=1)       # warm-up
        t0 = time.perf_counter()
        rel, status, counts = io.run_sequence(lf, rf, synth.KITTI_P0, synth.KITTI_P1, batch=B, seed=1)
        dt = time.perf_counter() - t0
        out["host"] = dict(value=(n - 1) / dt, equals_vo_frames=bool(np.array_equal(rel[:B + 1], ref[0])),
                           note="host decode on every core, decode of batch k+1 overlaps vo_frames of batch k, one batch at a time")
        for depth in (8,):
            pipe = io.DevicePngPipeline(H, W, batch=B, depth=3, decoders=10, device=torch.cuda.current_device())
            pipe.run(lf, rf, synth.KITTI_P0, synth.KITTI_P1, seed=1)      # warm-up: plans, buffers of every slot
            t0 = time.perf_counter()
            rel, status, counts = pipe.run(lf, rf, synth.KITTI_P0, synth.KITTI_P1, seed=1)
            dt = time.perf_counter() - t0
            pipe.close()
            out[f"device_depth{depth}"] = dict(value=(n - 1) / dt, equals_vo_frames=bool(np.array_equal(rel[:B + 1], ref[0])))
        # the decode kernels alone: one batch of 2 * (B + 1) images
        c = vo_b200.Context(torch.cuda.current_device())
        buf = torch.empty((2 * (B + 1), H, W), dtype=torch.uint8, device="cuda")
        io.read_batch_dev(lf[:B + 1] + rf[:B + 1], H, W, buf, c)
        c.profile_enable(True)
        io.read_batch_dev(lf[:B + 1] + rf[:B + 1], H, W, buf, c)
        pr = c.profile(); c.profile_enable(False); c.close()
        out["decode_kernels_ms_per_batch"] = dict(images=2 * (B + 1), inflate=pr["png_inflate"]["ms"], unfilter=pr["png_unfilter"]["ms"])
        out["device"] = out.pop("device_depth8")
        out["device"]["note"] = ("files go to the GPU as they are, inflate + un-filter kernels (one warp per image): ten batches in decode, "
                                 "three in the frame loop, each on its own context; the host reads files and gathers IDAT chunks")
        out["value"] = max(out["host"]["value"], out["device"]["value"])
        out["equals_vo_frames"] = bool(out["host"]["equals_vo_frames"] and out["device"]["equals_vo_frames"])
        out["note"] = ("wall clock, files in the page cache; value = the better of the two forms on this box: the host form scales with the "
                       "cores per GPU (400-700 images/s per core), the device form does not need them")
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return out


def png_leg_all_ranks(torch, dist, sleft, sright, B, rank, world):
    """N > 1: every rank runs the PNG leg at the same time (they share the host's cores, as a real multi-GPU job does);
    rank 0 reports its own line plus the slowest rank's rates."""
    dist.barrier()
    try:
        r = png_leg(sleft, sright, B)
        v = [r["host"]["value"], r["device"]["value"]]
    except Exception as e:   # noqa: BLE001
        r = dict(error=f"{type(e).__name__}: {e}")
        v = [0.0, 0.0]
    t = torch.tensor(v, device="cuda", dtype=torch.float64)
    allv = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allv, t)
    if rank == 0 and "error" not in r:
        r["all_ranks"] = dict(host_per_rank=[round(float(x[0])) for x in allv], device_per_rank=[round(float(x[1])) for x in allv],
                              host_total=float(sum(x[0] for x in allv)), device_total=float(sum(x[1] for x in allv)),
                              cores_per_gpu=(os.cpu_count() or 1) / world)
    return r


def reloc_leg(args, torch, dist, ctx, rank, world, pk):
    """BASELINE.json configs[4] (every rank calls this)."""
    from vo_b200 import reloc
    r = reloc.run(ctx, rank, world, dist, queries_per_rank=131072, landmarks=1048576, hyps=4096, reps=3,
                  oracle_rows=None if args.no_cpu else oracle_top2_rows, oracle_keep=oracle_keep)
    if r is not None:
        r["frac_of_2x_burst_peak_per_gpu"] = r["aggregate_tops_gemm"] / world / (2 * pk["tf_burst"])
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="new frames per step per GPU")
    ap.add_argument("--inflight", type=int, default=3, help="batches kept in flight per GPU (contexts / streams / host threads)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip every oracle leg (cpu_baseline, parity checks)")
    ap.add_argument("--quick", action="store_true", help="headline + trajectory only (used for the ncu launch list)")
    args = ap.parse_args()
    _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        if _POOL is not None:
            _POOL.terminate()
        return
    if args.warmup < 3:
        args.warmup = 3
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
