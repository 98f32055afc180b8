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


def prepare_landmarks(ctx, landmarks_dev):
    """vo_landmarks_prepare: convert a replicated landmark set ([n2, 128] float32 CUDA tensor of integer-valued SIFT rows)
    once into the match operand form held by ``ctx``; afterwards the float tensor may be freed and the shard calls take
    ``landmarks_dev=None``.  Returns n2."""
    import ctypes as C
    from . import _lib
    n2 = int(landmarks_dev.shape[0])
    _lib.check(_lib.lib().vo_landmarks_prepare(ctx.handle, C.c_void_p(landmarks_dev.data_ptr()), n2, int(landmarks_dev.shape[1]),
                                              C.c_void_p(ctx.stream)))
    return n2


class PeerGather:
    """Relocalisation shard with the all-gather fused into the match epilogue (vo_match_best2_gather_dev).

    Every rank owns a device buffer for the gathered records of ALL query rows ([n_total, 4] int32 = 16 bytes per row),
    allocated by the library and exported through CUDA IPC; every rank maps every peer's buffer.  ``run`` then launches
    the match on this rank's rows: the records kernel stores each 16-byte record into all ``world`` buffers directly
    (NVLink / NVSwitch peer stores, issued while the kernel runs), so no collective moves data afterwards -- one
    barrier on the launching stream tells the ranks that every slice has landed.  Slices are the equal ``row_chunks``
    of the total, so no counts are exchanged and nothing synchronises with the host."""

    def __init__(self, ctx, n_total, rank, world, dist):
        import ctypes as C
        import torch
        from . import _lib
        self.ctx, self.rank, self.world, self.dist, self.n_total = ctx, rank, world, dist, n_total
        L = _lib.lib()
        self._L = L
        own = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        _lib.check(L.vo_peer_alloc(ctx.handle, C.c_size_t(n_total * 16), C.byref(own), handle))
        self.own = own.value
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, bytes(handle))
        else:
            handles[0] = bytes(handle)
        self.ptrs = []
        for r in range(world):
            if r == rank:
                self.ptrs.append(self.own)
            else:
                p = C.c_void_p()
                hb = (C.c_uint8 * 64).from_buffer_copy(handles[r])
                _lib.check(L.vo_peer_open(ctx.handle, hb, C.byref(p)))
                self.ptrs.append(p.value)
        self._table = (C.c_void_p * world)(*self.ptrs)
        self.chunks = row_chunks(n_total, world)
        self._flag = torch.zeros(1, device=torch.device("cuda", torch.cuda.current_device()))
        self._stream = torch.cuda.ExternalStream(ctx.stream)

    def records(self):
        """This rank's gathered buffer as an [n_total, 4] int32 CUDA tensor (columns 1, 2 are float32 bit patterns)."""
        import torch

        class _Arr:
            pass
        a = _Arr()
        a.__cuda_array_interface__ = dict(shape=(self.n_total, 4), typestr="<i4", data=(self.own, False), version=3)
        return torch.as_tensor(a, device=torch.device("cuda", torch.cuda.current_device()))

    def run(self, queries_dev, landmarks_dev, opts=None, n_landmarks=None):
        """queries_dev: this rank's slice (row_chunks(n_total, world)[rank]) [n1, 128] float32 CUDA; landmarks replicated
        (``landmarks_dev=None`` with ``n_landmarks`` = the count given to prepare_landmarks: the prepared set is used).
        Asynchronous: the records are complete on every rank once the launching stream (ctx.stream) has passed the
        barrier enqueued here."""
        import ctypes as C
        import torch
        from . import _lib
        lo, hi = self.chunks[self.rank]
        n1 = int(queries_dev.shape[0])
        n2 = int(landmarks_dev.shape[0]) if landmarks_dev is not None else int(n_landmarks)
        assert n1 == hi - lo, "the query slice must be this rank's row_chunks share"
        mo = _lib.MatchOpts(*opts) if opts is not None else None
        _lib.check(self._L.vo_match_best2_gather_dev(self.ctx.handle, C.c_void_p(queries_dev.data_ptr()), n1,
                                                     C.c_void_p(landmarks_dev.data_ptr()) if landmarks_dev is not None else None, n2,
                                                     int(queries_dev.shape[1]),
                                                     C.byref(mo) if mo is not None else None, self._table, self.world,
                                                     C.c_size_t(lo), C.c_void_p(self.ctx.stream)))
        if self.world > 1:
            with torch.cuda.stream(self._stream):          # the barrier follows the kernel in stream order
                self.dist.all_reduce(self._flag)
        return self.records()

    def close(self):
        import ctypes as C
        from . import _lib
        if self.world > 1:
            import torch
            torch.cuda.synchronize()
            self.dist.barrier()                            # nobody unmaps while a peer may still be writing
        for r, p in enumerate(self.ptrs):
            if r != self.rank and p:
                self._L.vo_peer_close(self.ctx.handle, C.c_void_p(p))
        if self.world > 1:
            self.dist.barrier()
        if self.own:
            self._L.vo_peer_free(self.ctx.handle, C.c_void_p(self.own))
            self.own = None
        self.ptrs = []
