"""Input staging for the frame loop: the reference's ``imageDatastore`` /
``readimage`` (VO.m:16-17, 71-72) for KITTI odometry frames, i.e. 8-bit grayscale PNG files.

``ImageDatastore`` mirrors the two members VO.m uses (``Files``, ``readimage``); ``read_batch`` decodes
many files with native worker threads (libvo_b200: its own inflate + PNG row filters) into one
contiguous -- optionally pinned -- batch buffer, and ``run_sequence`` overlaps decoding batch k+1 on
the host with vo_frames on batch k."""
import ctypes as C
import glob
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib
from ._lib import check


def png_info(data):
    """(rows, cols, bit_depth, colour_type) of a PNG held in memory (bytes)."""
    buf = np.frombuffer(data, dtype=np.uint8)
    r, c, d, t = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    check(_lib.lib().vo_png_info(buf.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(len(buf)), C.byref(r), C.byref(c),
                                 C.byref(d), C.byref(t)))
    return r.value, c.value, d.value, t.value


def png_decode(data):
    """8-bit grayscale PNG (bytes) -> uint8 [rows, cols]."""
    rows, cols, _, _ = png_info(data)
    buf = np.frombuffer(data, dtype=np.uint8)
    out = np.empty((rows, cols), dtype=np.uint8)
    check(_lib.lib().vo_png_decode_gray8(buf.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(len(buf)),
                                         out.ctypes.data_as(C.POINTER(C.c_uint8)), cols, rows, cols))
    return out


def inflate_zlib(data, n_out):
    """The reader's inflate stage alone: zlib stream (bytes) -> exactly ``n_out`` bytes (Adler-32 checked)."""
    buf = np.frombuffer(data, dtype=np.uint8)
    out = np.empty(max(n_out, 1), dtype=np.uint8)
    check(_lib.lib().vo_inflate_zlib(buf.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(len(buf)),
                                     out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(n_out)))
    return out[:n_out].tobytes()


def read_batch(paths, rows=None, cols=None, out=None, threads=0):
    """Decode ``paths`` into out[n, rows, cols] uint8 (allocated when None; pass a pinned buffer's
    NumPy view to decode straight into page-locked memory).  The GIL is released while decoding."""
    paths = [os.fspath(p) for p in paths]
    if rows is None or cols is None:
        with open(paths[0], "rb") as f:
            rows, cols, _, _ = png_info(f.read(64))
    n = len(paths)
    if out is None:
        out = np.empty((n, rows, cols), dtype=np.uint8)
    assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"] and out.size >= n * rows * cols
    arr = (C.c_char_p * n)(*[p.encode() for p in paths])
    check(_lib.lib().vo_png_read_batch(arr, n, rows, cols, out.ctypes.data_as(C.POINTER(C.c_uint8)), threads))
    return out[:n] if out.ndim == 3 else out


def decode_batch_dev(files, rows, cols, out_dev, ctx=None):
    """Device-side decode (vo_png_decode_batch_dev): ``files`` = encoded PNG files as bytes objects, ``out_dev`` a
    torch uint8 CUDA tensor with room for [n, rows, cols] (or an int device pointer).  The DEFLATE streams are
    inflated and un-filtered by CUDA kernels; the host only gathers the IDAT chunks."""
    from . import api
    ctx = ctx or api.default_context()
    n = len(files)
    bufs = [np.frombuffer(f, dtype=np.uint8) for f in files]
    ptrs = (C.POINTER(C.c_uint8) * n)(*[b.ctypes.data_as(C.POINTER(C.c_uint8)) for b in bufs])
    sizes = (C.c_size_t * n)(*[len(b) for b in bufs])
    dptr = out_dev if isinstance(out_dev, int) else out_dev.data_ptr()
    check(_lib.lib().vo_png_decode_batch_dev(ctx.handle, ptrs, sizes, n, rows, cols, C.c_void_p(dptr)))
    return out_dev


def read_batch_dev(paths, rows, cols, out_dev, ctx=None, threads=0):
    """Files on disk -> device batch buffer (vo_png_read_batch_dev: a few host threads read, the GPU decodes)."""
    from . import api
    ctx = ctx or api.default_context()
    paths = [os.fspath(p) for p in paths]
    n = len(paths)
    arr = (C.c_char_p * n)(*[p.encode() for p in paths])
    dptr = out_dev if isinstance(out_dev, int) else out_dev.data_ptr()
    check(_lib.lib().vo_png_read_batch_dev(ctx.handle, arr, n, rows, cols, C.c_void_p(dptr), threads))
    return out_dev


class ImageDatastore:
    """imageDatastore(folder) as VO.m uses it: sorted ``Files`` and ``readimage(i)`` (1-based)."""

    def __init__(self, location, pattern="*.png"):
        self.Files = sorted(glob.glob(os.path.join(os.fspath(location), pattern)))

    def __len__(self):
        return len(self.Files)

    def readimage(self, i):
        with open(self.Files[i - 1], "rb") as f:
            return png_decode(f.read())


def run_sequence(left_files, right_files, P1, P2, batch=32, seed=0, threads=0, ctx=None, pinned=True):
    """The VO.m loop over a stereo PNG sequence: frames are decoded ``batch`` at a time (plus the
    one-frame halo) into two alternating batch buffers by a background thread while the GPU runs
    vo_frames on the other one.  Returns (rel_pose [n,4,4], status [n], counts [n,8])."""
    from . import vo
    n = len(left_files)
    assert len(right_files) == n and n > 0
    with open(left_files[0], "rb") as f:
        rows, cols, _, _ = png_info(f.read(64))
    bufs = []
    for _ in range(2):
        if pinned:
            import torch
            t = torch.empty((2, batch + 1, rows, cols), dtype=torch.uint8).pin_memory()
            bufs.append((t, t.numpy()))
        else:
            a = np.empty((2, batch + 1, rows, cols), dtype=np.uint8)
            bufs.append((a, a))
    chunks = [(max(b0 - 1, 0), min(b0 + batch, n)) for b0 in range(0, n, batch)]   # [lo, hi) with halo

    def load(k, slot):
        lo, hi = chunks[k]
        read_batch(left_files[lo:hi], rows, cols, bufs[slot][1][0], threads)
        read_batch(right_files[lo:hi], rows, cols, bufs[slot][1][1], threads)

    rel = np.tile(np.eye(4), (n, 1, 1)); status = np.zeros(n, dtype=np.int32); counts = np.zeros((n, 8), dtype=np.int32)
    load(0, 0)
    # the decode of batch k+1 runs as a future: a decode error (bad / missing PNG, size mismatch) is re-raised
    # by .result() BEFORE that buffer is handed to vo_frames, never swallowed by a background thread
    with ThreadPoolExecutor(max_workers=1) as pool:
        pending = None
        for k, (lo, hi) in enumerate(chunks):
            if pending is not None:
                pending.result()
            pending = pool.submit(load, k + 1, (k + 1) & 1) if k + 1 < len(chunks) else None
            m = hi - lo
            a = bufs[k & 1][1]
            try:
                r, s, c = vo.run_frames(a[0, :m], a[1, :m], P1, P2, seed=seed, first_frame=lo, ctx=ctx)
            except BaseException:
                if pending is not None:
                    pending.cancel()
                raise
            first = 0 if lo == 0 else 1          # the halo frame's outputs belong to the previous chunk
            rel[lo + first:hi] = r[first:]; status[lo + first:hi] = s[first:]; counts[lo + first:hi] = c[first:]
    return rel, status, counts


class DevicePngPipeline:
    """PNG files -> poses with the decode on the GPU (vo_png.cu).  Two pools of host threads, each thread with its
    own context (stream): ``decoders`` slots read the files of a batch (left and right together: one decode launch, one
    warp per image) into a device batch buffer; ``depth`` slots run vo_frames_dev on decoded buffers.  The serial
    inflate of a batch is tens of milliseconds of latency on a few per cent of the SMs, so many batches are kept in
    decode at once (a decode slot costs one 31 MB buffer), while three frame-loop slots (7.5 GB of pyramids each)
    are enough to saturate the SMs.  The host only reads files and gathers IDAT chunks."""

    def __init__(self, rows, cols, batch=32, depth=3, decoders=10, device=0, threads=2):
        import torch
        from . import api
        self.rows, self.cols, self.batch, self.device, self.threads = rows, cols, batch, device, threads
        self.frame_ctxs = [api.Context(device) for _ in range(max(1, depth))]
        self.decode_ctxs = [api.Context(device) for _ in range(max(1, decoders))]
        n_buf = len(self.frame_ctxs) + len(self.decode_ctxs) + 2
        self.bufs = [torch.empty((2 * (batch + 1), rows, cols), dtype=torch.uint8, device=torch.device("cuda", device))
                     for _ in range(n_buf)]

    def run(self, left_files, right_files, P1, P2, seed=0):
        """Returns (rel_pose [n,4,4], status [n], counts [n,8])."""
        import queue
        import threading
        from . import vo
        n, batch, rows, cols = len(left_files), self.batch, self.rows, self.cols
        assert len(right_files) == n and n > 0
        chunks = [(max(b0 - 1, 0), min(b0 + batch, n)) for b0 in range(0, n, batch)]   # [lo, hi) with halo
        rel = np.tile(np.eye(4), (n, 1, 1)); status = np.zeros(n, dtype=np.int32); counts = np.zeros((n, 8), dtype=np.int32)
        free = queue.Queue()
        for b in self.bufs:
            free.put(b)
        ready = queue.Queue()
        nxt = iter(range(len(chunks)))
        lock = threading.Lock()
        err = []

        live = [len(self.decode_ctxs)]

        def decoder(ctx):
            try:
                while not err:
                    with lock:
                        k = next(nxt, None)
                    if k is None:
                        break
                    lo, hi = chunks[k]
                    buf = free.get()
                    try:
                        read_batch_dev(list(left_files[lo:hi]) + list(right_files[lo:hi]), rows, cols, buf, ctx, self.threads)
                    except BaseException:
                        free.put(buf)
                        raise
                    ready.put((k, buf))
            except BaseException as e:   # noqa: BLE001  (surfaced in the caller's thread)
                err.append(e)
            finally:
                with lock:
                    live[0] -= 1
                    last = live[0] == 0
                if last:                      # every batch has been queued: one end marker per frame thread
                    for _ in self.frame_ctxs:
                        ready.put(None)

        def framer(ctx):
            while True:
                item = ready.get()
                if item is None:
                    return
                k, buf = item
                try:
                    if not err:               # after a failure the queue is only drained (the decoders must not block)
                        lo, hi = chunks[k]
                        m = hi - lo
                        r, s, c = vo.run_frames(None, None, P1, P2, seed=seed, first_frame=lo, ctx=ctx,
                                                device_ptrs=(buf[0].data_ptr(), buf[m].data_ptr(), m, rows, cols))
                        first = 0 if lo == 0 else 1
                        rel[lo + first:hi] = r[first:]; status[lo + first:hi] = s[first:]; counts[lo + first:hi] = c[first:]
                except BaseException as e:   # noqa: BLE001
                    err.append(e)
                finally:
                    free.put(buf)
        th = [threading.Thread(target=decoder, args=(c,)) for c in self.decode_ctxs] + \
             [threading.Thread(target=framer, args=(c,)) for c in self.frame_ctxs]
        [t.start() for t in th]
        [t.join() for t in th]
        if err:
            raise err[0]
        return rel, status, counts

    def close(self):
        for c in self.frame_ctxs + self.decode_ctxs:
            c.close()
        self.bufs = []


def run_sequence_device(left_files, right_files, P1, P2, batch=32, seed=0, depth=3, decoders=10, device=0, threads=2):
    """run_sequence with the PNG decode on the GPU (DevicePngPipeline for one call)."""
    with open(left_files[0], "rb") as f:
        rows, cols, _, _ = png_info(f.read(64))
    n_chunks = max(1, (len(left_files) + batch - 1) // batch)
    pipe = DevicePngPipeline(rows, cols, batch=batch, depth=min(depth, n_chunks), decoders=min(decoders, n_chunks), device=device,
                             threads=threads)
    try:
        return pipe.run(left_files, right_files, P1, P2, seed=seed)
    finally:
        pipe.close()
