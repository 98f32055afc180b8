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


class _Bits:
    """LSB-first DEFLATE bit writer for hand-built streams."""
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def put(self, v, nbits):
        self.acc |= v << self.n
        self.n += nbits
        while self.n >= 8:
            self.out.append(self.acc & 255); self.acc >>= 8; self.n -= 8

    def align(self):
        if self.n:
            self.out.append(self.acc & 255); self.acc, self.n = 0, 0


def _dyn_block_AA(bits, last, n_lit=2):
    """A dynamic-Huffman block whose alphabet is {'A': 1 bit, end-of-block: 1 bit} (so two literals pair up in
    the decoder's two-literal table entries), emitting n_lit 'A's."""
    bits.put(last, 1); bits.put(2, 2)
    bits.put(0, 5); bits.put(1, 5); bits.put(14, 4)           # 257 literal/length codes, 2 distance codes, 18 pre-code lengths
    order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
    for s in order[:18]:
        bits.put(1 if s in (1, 18) else 0, 3)                 # pre-code: symbol 1 -> '0', symbol 18 -> '1'

    def zeros(k):
        bits.put(1, 1); bits.put(k - 11, 7)
    zeros(65); bits.put(0, 1)                                  # lengths 0..64 = 0, 'A' = 1
    zeros(138); zeros(52); bits.put(0, 1)                      # 66..255 = 0, end-of-block = 1
    bits.put(0, 1); bits.put(0, 1)                             # two distance codes of length 1
    for _ in range(n_lit):
        bits.put(0, 1)
    bits.put(1, 1)


def _stored_block(bits, last, payload):
    bits.put(last, 1); bits.put(0, 2); bits.align()
    bits.out += len(payload).to_bytes(2, "little") + (len(payload) ^ 0xFFFF).to_bytes(2, "little") + payload


def _zlib_wrap(bits, data):
    bits.align()
    return b"\x78\x9c" + bytes(bits.out) + zlib.adler32(data).to_bytes(4, "big")


def test_inflate_hand_built_multi_block_streams():
    """Streams no encoder emits: literal pairs at the very end of the output followed by end-of-block and a stored
    block (the pair's second byte may land one past the end -- every later bound must notice)."""
    from vo_b200 import io, VoError
    rng = np.random.default_rng(5)
    # valid: dynamic 'A'*k + stored tail, for even and odd k
    for k in (0, 1, 2, 3, 7, 8):
        tail = rng.integers(0, 256, 1000, dtype=np.uint8).tobytes()
        b = _Bits(); _dyn_block_AA(b, 0, k); _stored_block(b, 1, tail)
        data = b"A" * k + tail
        z = _zlib_wrap(b, data)
        assert zlib.decompress(z) == data
        assert io.inflate_zlib(z, len(data)) == data
        # dynamic block last, stored first
        b = _Bits(); _stored_block(b, 0, tail); _dyn_block_AA(b, 1, k)
        z = _zlib_wrap(b, tail + b"A" * k)
        assert zlib.decompress(z) == tail + b"A" * k
        assert io.inflate_zlib(z, len(tail) + k) == tail + b"A" * k
    # the ADVICE proof of concept: 'A','A',end-of-block into a ONE-byte output, then a 60000-byte stored block
    for n_out in (0, 1):
        for k in (n_out + 1, n_out + 2):
            b = _Bits(); _dyn_block_AA(b, 0, k); _stored_block(b, 1, bytes(60000))
            with pytest.raises(VoError):
                io.inflate_zlib(_zlib_wrap(b, b""), n_out)
            b = _Bits(); _dyn_block_AA(b, 0, k); _dyn_block_AA(b, 0, 2); _stored_block(b, 0, b"xy"); _stored_block(b, 1, bytes(65535))
            with pytest.raises(VoError):
                io.inflate_zlib(_zlib_wrap(b, b""), n_out)
