"""Multi-GPU sharding of the hot path (one process per GPU, torch.distributed for the plumbing).

Only what shards naturally:

* ``frame_chunks`` / ``run_sequence_sharded`` -- an offline sequence is cut into contiguous chunks,
  one per rank, each with a one-frame halo (frame i's tracker state is a pure function of frame
  i-1's images, VO.m:207-210, 225-230).  No data-path collective; the per-frame relative poses
  (16 doubles) are gathered to every rank and the pose chain pose = pose * rel (VO.m:130) is
  multiplied sequentially -- it does not shard.
* ``match_top2_row_sharded`` -- map relocalisation: query rows are sharded across ranks, the
  landmark descriptors are replicated, each rank's per-row best-2 is final for its rows, and one
  all-gather (16 B per row: j1, s1, s2, pad) assembles the result.

The sharding arithmetic is pure host logic (tested with the gloo backend on CPU); the compute
callbacks are the CUDA operators.
"""
import numpy as np


def frame_chunks(n_frames, world):
    """[(lo, hi)] per rank over frames 1..n_frames-1 (frame 0 only seeds); rank r processes images
    lo-1 .. hi-1 and emits relative poses for frames lo .. hi-1.  Contiguous, balanced, disjoint."""
    n = max(n_frames - 1, 0)
    base, rem = divmod(n, world)
    out, lo = [], 1
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def row_chunks(n_rows, world):
    base, rem = divmod(n_rows, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def run_sequence_sharded(load_frames, n_frames, run_frames_fn, rank, world, dist=None, batch=32, device=None):
    """load_frames(lo, hi) -> (left, right) uint8 arrays for frames lo..hi-1.
    run_frames_fn(left, right, first_frame) -> (rel [n,4,4], status [n], counts [n,8]).
    Returns (rel_all [n_frames,4,4], status_all [n_frames]) on every rank (frame 0: identity)."""
    import torch
    lo, hi = frame_chunks(n_frames, world)[rank]
    rel = np.tile(np.eye(4), (n_frames, 1, 1))
    status = np.zeros(n_frames, dtype=np.int64)
    for b0 in range(lo, hi, batch):
        b1 = min(b0 + batch, hi)
        left, right = load_frames(b0 - 1, b1)            # one-frame halo
        r, s, _ = run_frames_fn(left, right, b0 - 1)
        rel[b0:b1] = r[1:]
        status[b0:b1] = s[1:]
    if world > 1:
        # every rank owns a disjoint frame range and the rest is identity/zero: a SUM of the
        # deviations from identity assembles the full array (tiny: 128 B per frame)
        t = torch.from_numpy(rel - np.eye(4)).to(device or "cpu")
        st = torch.from_numpy(status).to(device or "cpu")
        dist.all_reduce(t)
        dist.all_reduce(st)
        rel = t.cpu().numpy() + np.eye(4)
        status = st.cpu().numpy()
    return rel, status


def match_top2_row_sharded(queries, landmarks, match_top2_fn, rank, world, dist=None, device=None):
    """queries [n1, d] (every rank holds all, or at least its slice), landmarks [n2, d] replicated.
    match_top2_fn(q, l) -> (j1 u32, s1 f32, s2 f32).  Returns the gathered (j1, s1, s2) for all n1 rows."""
    import torch
    n1 = len(queries)
    chunks = row_chunks(n1, world)
    lo, hi = chunks[rank]
    j1, s1, s2 = match_top2_fn(queries[lo:hi], landmarks)
    if world == 1:
        return j1, s1, s2
    width = max(h - l for l, h in chunks)
    rec = torch.zeros((width, 4), dtype=torch.float32)       # 16-byte record per row
    rec[: hi - lo, 0] = torch.from_numpy(j1.astype(np.uint32).view(np.float32))   # bit-cast, not converted
    rec[: hi - lo, 1] = torch.from_numpy(s1)
    rec[: hi - lo, 2] = torch.from_numpy(s2)
    rec = rec.to(device or "cpu")
    out = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(out, rec)
    J, S1, S2 = [], [], []
    for (l, h), t in zip(chunks, out):
        a = t[: h - l].cpu().numpy()
        J.append(a[:, 0].copy().view(np.uint32)); S1.append(a[:, 1].copy()); S2.append(a[:, 2].copy())
    return np.concatenate(J), np.concatenate(S1), np.concatenate(S2)


def relocalise_row_sharded_dev(ctx, queries_dev, landmarks_dev, rank, world, dist=None, opts=None):
    """Device-resident relocalisation match (BASELINE config 5).  queries_dev: this rank's slice
    [n1_local, 128] float32 CUDA tensor (row_chunks(n1, world)[rank] of the full query set),
    landmarks_dev: [n2, 128] float32 CUDA tensor, replicated.  Each rank runs the tcgen05 match on its
    rows (vo_match_best2_dev: one 16-byte record {j1, s1, s2, keep} per row); one NCCL all-gather of the
    records assembles every row on every rank.  Returns (records [n1_total, 4] int32 CUDA tensor --
    columns 1, 2 are float32 bit patterns -- and the per-rank row counts)."""
    import ctypes as C
    import torch
    from . import _lib
    n1, n2 = int(queries_dev.shape[0]), int(landmarks_dev.shape[0])
    rec = torch.empty((max(n1, 1), 4), dtype=torch.int32, device=queries_dev.device)
    mo = None
    if opts is not None:
        mo = _lib.MatchOpts(*opts)
    _lib.check(_lib.lib().vo_match_best2_dev(ctx.handle, C.c_void_p(queries_dev.data_ptr()), n1,
                                             C.c_void_p(landmarks_dev.data_ptr()), n2, int(queries_dev.shape[1]),
                                             C.byref(mo) if mo is not None else None, C.c_void_p(rec.data_ptr()),
                                             C.c_void_p(ctx.stream)))
    if world == 1:
        ctx.sync()
        return rec[:n1], [n1]
    # ranks may hold different row counts: gather the counts, pad to the widest slice
    cnt = torch.tensor([n1], dtype=torch.int64, device=queries_dev.device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    ctx.sync()                              # the match ran on the library's stream
    dist.all_gather(cnts, cnt)
    counts = [int(c.item()) for c in cnts]
    width = max(counts)
    if rec.shape[0] < width:
        pad = torch.zeros((width, 4), dtype=torch.int32, device=rec.device); pad[:n1] = rec[:n1]; rec = pad
    out = torch.empty((world * width, 4), dtype=torch.int32, device=rec.device)
    dist.all_gather_into_tensor(out, rec[:width].contiguous())
    parts = [out[r * width: r * width + counts[r]] for r in range(world)]
    return torch.cat(parts, 0), counts
