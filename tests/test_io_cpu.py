"""CPU suite: input staging.  The native PNG decoder (zlib inflate + the five PNG row
filters) against an independent encoder written here (every filter type, per row) and against
OpenCV's encoder; error behaviour for the PNG flavours outside KITTI's (8-bit gray) scope."""
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
