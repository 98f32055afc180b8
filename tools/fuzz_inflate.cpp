// Differential fuzz of the library's inflate (csrc/vo_inflate.cu) against zlib, meant to run under
// AddressSanitizer / UBSan on the CPU box:
//   cp r7020e-visual-odometry_b200/csrc/vo_inflate.cu /tmp/inf.cpp
//   g++ -O1 -g -fsanitize=address,undefined -Ir7020e-visual-odometry_b200/csrc -o /tmp/fuzz tools/fuzz_inflate.cpp /tmp/inf.cpp -lz
//   /tmp/fuzz <seed> <iterations>
// Streams come from zlib at random levels / strategies / window sizes over four kinds of data, then get
// random bit flips, byte replacements, truncation and wrong expected sizes.  Buffers are exact-size
// heap blocks (payload + the documented slack), so any access beyond the contract trips ASan.  The
// decoder must agree with zlib whenever it accepts, and must accept every unmodified stream.
#include "vo_inflate.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include <zlib.h>
int main(int argc, char** argv) {
  std::mt19937_64 rng(argc > 1 ? atoi(argv[1]) : 1);
  const int iters = argc > 2 ? atoi(argv[2]) : 20000;
  long ok = 0, bad = 0, agree = 0;
  for (int it = 0; it < iters; ++it) {
    // make a valid stream from structured data
    size_t n = rng() % 5000;
    std::vector<uint8_t> d(n);
    int mode = rng() % 4;
    for (size_t i = 0; i < n; ++i) d[i] = mode == 0 ? rng() & 255 : mode == 1 ? (rng() % 7 == 0 ? rng() & 255 : 0) : mode == 2 ? (uint8_t)(i * 3 + (rng() & 3)) : (uint8_t)((rng() % 100 < 90) ? d[i ? i - 1 : 0] : rng());
    uLongf zn = compressBound(n);
    std::vector<uint8_t> z(zn);
    int level = rng() % 10;
    z_stream s; memset(&s, 0, sizeof(s));
    int strat[5] = {Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED};
    deflateInit2(&s, level, Z_DEFLATED, 9 + rng() % 7, 1 + rng() % 9, strat[rng() % 5]);
    s.next_in = d.data(); s.avail_in = n; s.next_out = z.data(); s.avail_out = zn;
    deflate(&s, Z_FINISH); zn = s.total_out; deflateEnd(&s);
    // mutate
    int nm = rng() % 4;   // 0: none
    for (int m = 0; m < nm && zn; ++m) { size_t p = rng() % zn; if (rng() & 1) z[p] ^= 1 << (rng() & 7); else z[p] = rng(); }
    size_t zcut = (rng() % 8 == 0 && zn) ? rng() % zn : zn;
    size_t want = (rng() % 8 == 0) ? rng() % (n + 10) : n;
    // exact-size heap buffers so ASan sees any access past the documented slack
    uint8_t* in = (uint8_t*)malloc(zcut + vo::INFLATE_IN_SLACK); memcpy(in, z.data(), zcut); memset(in + zcut, 0, vo::INFLATE_IN_SLACK);
    uint8_t* out = (uint8_t*)malloc(want + vo::INFLATE_OUT_SLACK);
    bool r = vo::inflate_zlib_fast(in, zcut, out, want);
    // zlib's verdict
    std::vector<uint8_t> ref(want + 1);
    uLongf got = want + 1;
    int zr = uncompress(ref.data(), &got, z.data(), zcut);
    bool zok = zr == Z_OK && got == want;
    if (r) { ++ok; if (!zok || memcmp(out, ref.data(), want)) { printf("MISMATCH it=%d r=%d zr=%d\n", it, r, zr); return 1; } }
    else { ++bad; if (zok && nm == 0 && zcut == zn) { printf("FALSE REJECT it=%d n=%zu level=%d\n", it, n, level); return 1; } }
    if (r == zok) ++agree;
    free(in); free(out);
  }
  printf("ok %ld rejected %ld agree-with-zlib %ld of %d\n", ok, bad, agree, iters);
}
