"""GPU suite: the device-side PNG decoder (vo_png.cu: DEFLATE + row filters as CUDA kernels) against the images
that were encoded -- every PNG filter type, stored / fixed / dynamic DEFLATE blocks, long codes, matches of every
kind, several IDAT chunks, OpenCV's encoder at KITTI size -- and its error behaviour on malformed files."""
import os
import struct
import zlib

import numpy as np
import pytest

from test_io_cpu import _chunk, encode_png

pytestmark = pytest.mark.gpu


def _png_from_raw(raw, rows, cols, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=15, idat_split=1, sync=0):
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, 9, strategy)
    if sync:
        z = b"".join(co.compress(raw[i:i + sync]) + co.flush(zlib.Z_SYNC_FLUSH) for i in range(0, len(raw), sync)) + co.flush()
    else:
        z = co.compress(raw) + co.flush()
    parts = [z[i * len(z) // idat_split:(i + 1) * len(z) // idat_split] for i in range(idat_split)]
    ihdr = struct.pack(">IIBBBBB", cols, rows, 8, 0, 0, 0, 0)
    return b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", ihdr) + b"".join(_chunk(b"IDAT", p) for p in parts) + _chunk(b"IEND", b"")


def _decode(files, rows, cols, ctx):
    import torch
    from vo_b200 import io
    out = torch.empty((len(files), rows, cols), dtype=torch.uint8, device="cuda")
    io.decode_batch_dev(files, rows, cols, out, ctx)
    return out.cpu().numpy()


def test_device_decode_every_filter_type(ctx):
    rng = np.random.default_rng(0)
    img = (rng.integers(0, 256, (77, 153)) * (rng.random((77, 153)) < 0.7)).astype(np.uint8)
    img[5:40, 10:120] = np.arange(110, dtype=np.uint8)[None, :] * 7
    files = [encode_png(img, f, idat_split=k) for f, k in (([0], 1), ([1], 2), ([2], 3), ([3], 1), ([4], 2), ([0, 1, 2, 3, 4], 3), ([4, 3, 4, 1], 1))]
    got = _decode(files, 77, 153, ctx)
    for g in got:
        assert np.array_equal(g, img)
    one = encode_png(img[:1, :1], [4])
    assert np.array_equal(_decode([one], 1, 1, ctx)[0], img[:1, :1])


def test_device_decode_every_block_type(ctx):
    """The filtered scanlines are compressed with every zlib level / strategy / window: stored, fixed and dynamic
    blocks, second-level tables, long and overlapping matches, empty stored blocks from sync flushes."""
    from vo_b200 import io
    rng = np.random.default_rng(3)
    rows, cols = 96, 311
    imgs = {
        "noise": rng.integers(0, 256, (rows, cols), dtype=np.uint8),
        "flat": np.full((rows, cols), 17, dtype=np.uint8),
        "smooth": (np.cumsum(rng.integers(-2, 3, rows * cols)) % 256).astype(np.uint8).reshape(rows, cols),
        "skewed": (rng.geometric(0.08, rows * cols) % 256).astype(np.uint8).reshape(rows, cols),
        "period": np.tile(rng.integers(0, 256, 5, dtype=np.uint8), rows * cols // 5 + 1)[:rows * cols].reshape(rows, cols),
    }
    files, want = [], []
    for name, img in imgs.items():
        raw = b"".join(bytes([0]) + img[y].tobytes() for y in range(rows))          # filter type 0: the image is the payload
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
                files.append(_png_from_raw(raw, rows, cols, level, strategy)); want.append(img)
        files.append(_png_from_raw(raw, rows, cols, 6, wbits=9, idat_split=4)); want.append(img)
        files.append(_png_from_raw(raw, rows, cols, 6, sync=5003)); want.append(img)
    got = _decode(files, rows, cols, ctx)
    for k, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g, w), k
    host = np.stack([io.png_decode(f) for f in files[:8]])
    assert np.array_equal(host, got[:8])


def test_device_decode_opencv_files_at_kitti_size(ctx, tmp_path):
    """OpenCV-written 1241 x 376 frames (adaptive filters, dynamic blocks): device decode == host decode == the frames;
    and the device-decoding sequence runner gives the poses of vo_frames on the decoded frames."""
    cv2 = pytest.importorskip("cv2")
    import torch
    from vo_b200 import io, synth, vo
    left, right = synth.shift_stream(5, seed=4)
    lf, rf = [], []
    for i in range(5):
        for name, arr, lst in (("image_0", left, lf), ("image_1", right, rf)):
            os.makedirs(os.path.join(tmp_path, name), exist_ok=True)
            p = os.path.join(tmp_path, name, f"{i:06d}.png")
            assert cv2.imwrite(p, arr[i]); lst.append(p)
    out = torch.empty((10, 376, 1241), dtype=torch.uint8, device="cuda")
    io.read_batch_dev(lf + rf, 376, 1241, out, ctx)
    got = out.cpu().numpy()
    assert np.array_equal(got[:5], left) and np.array_equal(got[5:], right)
    rel, status, counts = io.run_sequence_device(lf, rf, synth.KITTI_P0, synth.KITTI_P1, batch=2, seed=5, depth=2, decoders=3)
    rel0, status0, counts0 = vo.run_frames(left, right, synth.KITTI_P0, synth.KITTI_P1, seed=5, ctx=ctx)
    assert np.array_equal(rel, rel0) and np.array_equal(status, status0) and np.array_equal(counts, counts0)


def test_device_pipeline_surfaces_a_bad_file(ctx, tmp_path):
    """A corrupt file in the middle of a sequence ends the pipelined run with the decoder's error (no hang, no poses
    for a buffer that was not decoded)."""
    cv2 = pytest.importorskip("cv2")
    from vo_b200 import io, synth, VoError
    left, right = synth.shift_stream(6, seed=4, h=94, w=311)
    lf, rf = [], []
    for i in range(6):
        for name, arr, lst in (("l", left, lf), ("r", right, rf)):
            p = os.path.join(tmp_path, f"{name}{i:06d}.png")
            assert cv2.imwrite(p, arr[i]); lst.append(p)
    data = bytearray(open(lf[3], "rb").read())
    for k in range(200, 260):
        data[k] ^= 0x5A
    open(lf[3], "wb").write(bytes(data))
    with pytest.raises(VoError, match="png"):
        io.run_sequence_device(lf, rf, synth.KITTI_P0, synth.KITTI_P1, batch=2, seed=5, depth=2, decoders=3)


def test_device_decode_rejects_malformed_files(ctx):
    """Corrupted, truncated and mis-sized streams end in an error naming the image, never in a crash or a hang."""
    from vo_b200 import VoError
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (40, 64), dtype=np.uint8)
    good = encode_png(img, [0, 1, 2, 3, 4])
    assert np.array_equal(_decode([good], 40, 64, ctx)[0], img)
    with pytest.raises(VoError, match="expected"):
        _decode([good], 41, 64, ctx)
    with pytest.raises(VoError, match="signature"):
        _decode([b"definitely not a png file, not even close....."], 40, 64, ctx)
    raw = b"".join(bytes([0]) + img[y].tobytes() for y in range(40))
    z = bytearray(zlib.compress(raw, 6))
    ihdr = struct.pack(">IIBBBBB", 64, 40, 8, 0, 0, 0, 0)
    wrap = lambda zz: b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", bytes(zz)) + _chunk(b"IEND", b"")
    n_bad = 0
    for trial in range(60):                         # bit flips anywhere in the stream: an error or (rarely) the same image
        zz = bytearray(z)
        pos = int(rng.integers(0, len(zz)))
        zz[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            g = _decode([good, wrap(zz)], 40, 64, ctx)
            assert np.array_equal(g[0], img)
            assert np.array_equal(g[1], img)        # accepted only if the Adler-32 still matches: the pixels are right
        except VoError as e:
            assert "png 1:" in str(e)
            n_bad += 1
    assert n_bad >= 55
    for cut in (2, 7, len(z) // 2, len(z) - 5, len(z) - 1):
        with pytest.raises(VoError, match="png 0:"):
            _decode([wrap(z[:cut])], 40, 64, ctx)
    with pytest.raises(VoError, match="png 0:"):    # a valid stream of the wrong length
        _decode([wrap(zlib.compress(raw + b"abc", 6))], 40, 64, ctx)
    with pytest.raises(VoError, match="filter"):
        _decode([wrap(zlib.compress(bytes([9]) + raw[1:], 6))], 40, 64, ctx)
    # a paired-literal entry that straddles the end of the image (odd raw size, two zero literals per table entry), then
    # end of block and a 60000-byte stored block: the overrun has to be caught at the block end, not after the copy
    raw0 = bytes(39 * 65)
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 9, zlib.Z_HUFFMAN_ONLY)
    zz = co.compress(raw0 + b"\0") + co.flush(zlib.Z_SYNC_FLUSH)
    zz += bytes([1]) + struct.pack("<HH", 60000, 60000 ^ 0xFFFF) + bytes(60000) + struct.pack(">I", zlib.adler32(raw0))
    ihdr2 = struct.pack(">IIBBBBB", 64, 39, 8, 0, 0, 0, 0)
    bad = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", ihdr2) + _chunk(b"IDAT", zz) + _chunk(b"IEND", b"")
    with pytest.raises(VoError, match="png 0:"):
        _decode([bad], 39, 64, ctx)
    assert np.array_equal(_decode([good], 40, 64, ctx)[0], img)     # the context is still healthy
