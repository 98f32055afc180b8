"""CPU suite: input staging.  The native PNG decoder (the library's inflate + the five
PNG row filters) against an independent encoder written here (every filter type, per row) and against
OpenCV's encoder; the inflate stage alone against zlib on every block type and on malformed streams;
error behaviour for the PNG flavours outside KITTI's (8-bit gray) scope."""
import os
import struct
import zlib

import numpy as np
import pytest


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def encode_png(img, filters, depth=8, color=0, idat_split=1):
    """Minimal PNG encoder: row y uses filter type filters[y % len(filters)]."""
    rows, cols = img.shape
    raw = bytearray()
    for y in range(rows):
        ft = filters[y % len(filters)]
        cur = img[y].astype(np.int32)
        up = img[y - 1].astype(np.int32) if y else np.zeros(cols, np.int32)
        left = np.concatenate([[0], cur[:-1]])
        ul = np.concatenate([[0], up[:-1]])
        if ft == 0: f = cur
        elif ft == 1: f = cur - left
        elif ft == 2: f = cur - up
        elif ft == 3: f = cur - ((left + up) >> 1)
        else: f = cur - np.array([_paeth(int(a), int(b), int(c)) for a, b, c in zip(left, up, ul)])
        raw.append(ft); raw += bytes((f & 255).astype(np.uint8))
    z = zlib.compress(bytes(raw), 6)
    parts = [z[i * len(z) // idat_split:(i + 1) * len(z) // idat_split] for i in range(idat_split)]
    ihdr = struct.pack(">IIBBBBB", cols, rows, depth, color, 0, 0, 0)
    return b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", ihdr) + _chunk(b"tEXt", b"k\0v") + b"".join(_chunk(b"IDAT", p) for p in parts) + _chunk(b"IEND", b"")


def test_png_decode_every_filter_type():
    from vo_b200 import io
    rng = np.random.default_rng(0)
    img = (rng.integers(0, 256, (37, 53)) * (rng.random((37, 53)) < 0.7)).astype(np.uint8)
    img[5:20, 10:40] = np.arange(30, dtype=np.uint8)[None, :] * 7       # smooth region: filters actually predict
    for filters in ([0], [1], [2], [3], [4], [0, 1, 2, 3, 4], [4, 3, 4, 1]):
        png = encode_png(img, filters, idat_split=3)
        assert io.png_info(png) == (37, 53, 8, 0)
        assert np.array_equal(io.png_decode(png), img), filters
    one = encode_png(img[:1, :1], [4])
    assert np.array_equal(io.png_decode(one), img[:1, :1])


def test_png_decode_matches_opencv_encoder(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from vo_b200 import io, synth
    left, right = synth.shift_stream(3, seed=2, h=94, w=311)
    lf, rf = [], []
    for i in range(3):
        for name, arr, lst in (("l", left, lf), ("r", right, rf)):
            p = os.path.join(tmp_path, f"{name}{i:06d}.png")
            assert cv2.imwrite(p, arr[i])
            lst.append(p)
    out = io.read_batch(lf + rf, threads=3)
    assert np.array_equal(out[:3], left) and np.array_equal(out[3:], right)
    ds = io.ImageDatastore(tmp_path, "l*.png")
    assert len(ds) == 3 and np.array_equal(ds.readimage(2), left[1])


def test_png_rejects_what_it_does_not_support(tmp_path):
    from vo_b200 import io, VoError
    img = np.zeros((4, 6), dtype=np.uint8)
    with pytest.raises(VoError, match="8-bit grayscale"):
        io.png_decode(encode_png(img, [0], depth=8, color=3))
    with pytest.raises(VoError, match="signature"):
        io.png_decode(b"not a png at all" * 4)
    bad = bytearray(encode_png(img, [0]))
    with pytest.raises(VoError, match="cannot open"):
        io.read_batch([os.path.join(tmp_path, "missing.png")], 4, 6)
    p = os.path.join(tmp_path, "a.png")
    open(p, "wb").write(bytes(bad))
    with pytest.raises(VoError, match="expected"):
        io.read_batch([p], 5, 6)


def _streams():
    """zlib streams that exercise every block type and table shape: stored, fixed and dynamic blocks,
    long codes (second-level tables), long matches, distance-1 runs, multi-block inputs."""
    rng = np.random.default_rng(7)
    datas = {
        "empty": b"",
        "one": b"x",
        "zeros": bytes(70000),
        "random": rng.integers(0, 256, 200000, dtype=np.uint8).tobytes(),
        "text": (b"the quick brown fox jumps over the lazy dog; " * 3000)[:100003],
        "skewed": (rng.geometric(0.08, 300000) % 256).astype(np.uint8).tobytes(),      # long Huffman codes for rare bytes
        "smooth": (np.cumsum(rng.integers(-2, 3, 250000)) % 256).astype(np.uint8).tobytes(),
        "short_period": bytes(rng.integers(0, 256, 5, dtype=np.uint8)) * 20000,        # overlapping matches, distance < 8
        "image_rows": np.concatenate([np.r_[np.uint8(4), (rng.integers(0, 40, 1241) * (rng.random(1241) < .5)).astype(np.uint8)]
                                      for _ in range(120)]).tobytes(),
    }
    for name, d in datas.items():
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
                co = zlib.compressobj(level, zlib.DEFLATED, 15, 9, strategy)
                yield f"{name}/l{level}/s{strategy}", d, co.compress(d) + co.flush()
        co = zlib.compressobj(6, zlib.DEFLATED, 9, 1, zlib.Z_DEFAULT_STRATEGY)      # 512-byte window, small memLevel: many blocks
        yield f"{name}/w9", d, co.compress(d) + co.flush()
        co = zlib.compressobj(6)                                                      # sync flushes: empty stored blocks between
        z = b"".join(co.compress(d[i:i + 20011]) + co.flush(zlib.Z_SYNC_FLUSH) for i in range(0, len(d), 20011)) + co.flush()
        yield f"{name}/syncflush", d, z


def test_inflate_matches_zlib_on_every_block_type():
    from vo_b200 import io
    n = 0
    for name, d, z in _streams():
        assert zlib.decompress(z) == d
        assert io.inflate_zlib(z, len(d)) == d, name
        n += 1
    assert n > 150


def test_inflate_rejects_malformed_streams_without_overrunning():
    from vo_b200 import io, VoError
    rng = np.random.default_rng(11)
    d = (np.cumsum(rng.integers(-3, 4, 60000)) % 256).astype(np.uint8).tobytes()
    z = zlib.compress(d, 6)
    for bad, n_out in ((z[:-1], len(d)), (z[:len(z) // 2], len(d)), (z, len(d) - 1), (z, len(d) + 1),
                       (z[:-4] + bytes(4), len(d)), (b"\x78\x9c", 0), (b"", 0), (bytes(64), 10), (b"\x78\x9c\x07" + bytes(20), 10)):
        with pytest.raises(VoError):
            io.inflate_zlib(bad, n_out)
    # bit flips anywhere: either rejected, or (if the flip is harmless) the exact original -- never a crash
    for _ in range(300):
        b = bytearray(z)
        i = int(rng.integers(2, len(b)))
        b[i] ^= 1 << int(rng.integers(0, 8))
        try:
            assert io.inflate_zlib(bytes(b), len(d)) == d
        except VoError:
            pass
    # random garbage behind a valid header
    for _ in range(200):
        g = b"\x78\x9c" + rng.integers(0, 256, int(rng.integers(1, 400)), dtype=np.uint8).tobytes()
        try:
            io.inflate_zlib(g, int(rng.integers(0, 3000)))
        except VoError:
            pass


def test_png_paeth_runs_use_the_wavefront_path():
    """Runs of 1..9 Paeth rows between other filter types, widths around the wavefront's prologue."""
    from vo_b200 import io
    rng = np.random.default_rng(3)
    for cols in (1, 2, 3, 4, 5, 17, 64):
        img = (np.cumsum(rng.integers(-6, 7, (23, cols)), axis=1) % 256).astype(np.uint8)
        for filters in ([4], [4, 4, 1], [2, 4, 4, 4, 4, 4, 3], [4, 4, 4, 0, 4, 4, 4, 4, 4, 4, 4, 4, 4, 1], [1, 4, 2, 4, 4, 3]):
            assert np.array_equal(io.png_decode(encode_png(img, filters)), img), (cols, filters)
