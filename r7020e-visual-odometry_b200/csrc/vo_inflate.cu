// vo_inflate.cu -- DEFLATE (RFC 1951) / zlib-wrapper (RFC 1950) decoder for the PNG input stage
// (at >= 1000 frames/s the host-side inflate of the KITTI PNGs bounds the
// end-to-end rate).  Host code only.  Written for literal-heavy image data: a 64-bit bit buffer that
// is refilled once per symbol group with one unaligned 8-byte load, an 11-bit first-level
// literal/length table whose entries carry base value, extra-bit count and code length, up to three
// literals per refill, and 8-byte match copies.  The decoder is bounds-safe on arbitrary input; on any
// error it only reports failure -- the caller (vo_io.cu) then asks zlib for the authoritative
// verdict, so error behaviour is zlib's.
#include "vo_inflate.h"
#include <cstring>
#include <zlib.h>   // adler32 (fallback when the CPU has no AVX2)
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace vo {

namespace {

constexpr int LL_BITS = 11, D_BITS = 8, PRE_BITS = 7;
constexpr int LL_CAP = (1 << LL_BITS) + 288 * 16, D_CAP = (1 << D_BITS) + 32 * 128;
// table entry: bit 31 literal, 30 end of block, 29 sub-table link, 28 invalid;
// bits 8..27 value (literal byte | base length | base distance | sub-table offset);
// bits 4..7 extra-bit count (or sub-table index bits); bits 0..3 code bits to consume
// literal entries may carry a second literal (bit 27, value in bits 16..23, bits 0..3 = both code lengths)
constexpr uint32_t E_LIT = 1u << 31, E_EOB = 1u << 30, E_SUB = 1u << 29, E_BAD = 1u << 28, E_LIT2 = 1u << 27;

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t PRE_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint32_t ll_payload(int s) {
  if (s < 256) return E_LIT | ((uint32_t)s << 8);
  if (s == 256) return E_EOB;
  if (s <= 285) return ((uint32_t)LEN_BASE[s - 257] << 8) | ((uint32_t)LEN_EXTRA[s - 257] << 4);
  return E_BAD;
}
inline uint32_t d_payload(int s) {
  if (s < 30) return ((uint32_t)DIST_BASE[s] << 8) | ((uint32_t)DIST_EXTRA[s] << 4);
  return E_BAD;
}
inline uint32_t pre_payload(int s) { return (uint32_t)s << 8; }

struct Rev8 { uint8_t t[256]; constexpr Rev8() : t() { for (int i = 0; i < 256; ++i) { int r = 0; for (int b = 0; b < 8; ++b) r |= ((i >> b) & 1) << (7 - b); t[i] = (uint8_t)r; } } };
constexpr Rev8 REV8;
inline uint32_t bit_reverse(uint32_t c, int len) {   // len <= 15
  return (((uint32_t)REV8.t[c & 255u] << 8) | REV8.t[(c >> 8) & 255u]) >> (16 - len);
}

// Adler-32 of the inflated bytes, 32 bytes per step.  With A_s = a at the start of step s, a step adds
// sum(d) to a and 32 * A_s + sum((32 - j) * d_j) to b: psadbw gives the byte sums, pmaddubsw against
// the weights 32..1 the weighted sums, and the running total of the byte sums before each step gives
// sum_s (A_s - a0).  The modulo is taken every 5536 bytes, long before a 32-bit lane can overflow.
#if defined(__x86_64__)
__attribute__((target("avx2"))) inline uint64_t hsum(__m256i v) {
  __m128i x = _mm_add_epi32(_mm256_castsi256_si128(v), _mm256_extracti128_si256(v, 1));
  x = _mm_add_epi32(x, _mm_shuffle_epi32(x, 0x4E));
  x = _mm_add_epi32(x, _mm_shuffle_epi32(x, 0xB1));
  return (uint64_t)(uint32_t)_mm_cvtsi128_si32(x);
}
__attribute__((target("avx2"))) uint32_t adler32_avx2(uint32_t adler, const uint8_t* p, size_t n) {
  uint64_t a = adler & 0xFFFFu, b = adler >> 16;
  const __m256i weights = _mm256_setr_epi8(32, 31, 30, 29, 28, 27, 26, 25, 24, 23, 22, 21, 20, 19, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1);
  const __m256i ones16 = _mm256_set1_epi16(1), zero = _mm256_setzero_si256();
  while (n >= 32) {
    const size_t blk = n < 5536 ? (n & ~(size_t)31) : 5536;
    n -= blk;
    __m256i va = zero, vw = zero, vprev = zero;
    for (size_t k = 0; k < blk; k += 32, p += 32) {
      const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
      vprev = _mm256_add_epi32(vprev, va);
      va = _mm256_add_epi32(va, _mm256_sad_epu8(d, zero));
      vw = _mm256_add_epi32(vw, _mm256_madd_epi16(_mm256_maddubs_epi16(d, weights), ones16));
    }
    b = (b + (uint64_t)blk * a + 32u * hsum(vprev) + hsum(vw)) % 65521u;
    a = (a + hsum(va)) % 65521u;
  }
  for (; n; --n, ++p) { a += *p; b += a; }
  return (uint32_t)(((b % 65521u) << 16) | (a % 65521u));
}
#endif

// Canonical Huffman code lengths -> two-level decode table indexed by the next bits of the
// (LSB-first) stream.  Returns false for an over-subscribed code or a table overflow; an incomplete
// code leaves its unused slots E_BAD.
template <typename Payload>
bool build_table(const uint8_t* lens, int n, int tb, uint32_t* table, int cap, Payload payload) {
  int count[16] = {0};
  for (int s = 0; s < n; ++s) ++count[lens[s]];
  count[0] = 0;
  int left = 1;
  for (int l = 1; l <= 15; ++l) { left = (left << 1) - count[l]; if (left < 0) return false; }
  uint32_t next[16]; uint32_t code = 0;
  for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
  const int primary = 1 << tb;
  for (int i = 0; i < primary; ++i) table[i] = E_BAD | 1u;
  uint8_t sub_bits[1 << LL_BITS];
  bool any_long = false;
  uint32_t nx[16]; memcpy(nx, next, sizeof(nx));
  for (int l = tb + 1; l <= 15; ++l) any_long = any_long || count[l];
  if (any_long) {
    memset(sub_bits, 0, (size_t)primary);
    for (int s = 0; s < n; ++s) {
      const int l = lens[s];
      if (!l) continue;
      const uint32_t rev = bit_reverse(nx[l]++, l);
      if (l > tb) { uint8_t& sb = sub_bits[rev & (uint32_t)(primary - 1)]; if (l - tb > sb) sb = (uint8_t)(l - tb); }
    }
  }
  int free_at = primary;
  for (int s = 0; s < n; ++s) {
    const int l = lens[s];
    if (!l) continue;
    const uint32_t rev = bit_reverse(next[l]++, l);
    if (l <= tb) {
      const uint32_t e = payload(s) | (uint32_t)l;
      for (uint32_t k = rev; k < (uint32_t)primary; k += 1u << l) table[k] = e;
    } else {
      const uint32_t pfx = rev & (uint32_t)(primary - 1);
      uint32_t link = table[pfx];
      if (!(link & E_SUB)) {
        const int sb = sub_bits[pfx];
        if (free_at + (1 << sb) > cap) return false;
        link = E_SUB | ((uint32_t)free_at << 8) | ((uint32_t)sb << 4) | (uint32_t)tb;
        table[pfx] = link;
        for (int k = 0; k < (1 << sb); ++k) table[free_at + k] = E_BAD | 1u;
        free_at += 1 << sb;
      }
      const int sb = (int)((link >> 4) & 15u);
      uint32_t* sub = table + ((link >> 8) & 0xFFFFFu);
      const uint32_t e = payload(s) | (uint32_t)(l - tb);
      for (uint32_t k = rev >> tb; k < (1u << sb); k += 1u << (l - tb)) sub[k] = e;
    }
  }
  return true;
}

struct Tables { uint32_t ll[LL_CAP]; uint32_t d[D_CAP]; };

// Image data is mostly literals with short codes: where two consecutive literal codes fit into the
// LL_BITS index, the first-level entry decodes both at once (the second code is fully determined by
// the index bits that remain after the first, because its length does not exceed them).
void pair_literals(uint32_t* ll) {
  uint32_t one[1 << LL_BITS];
  memcpy(one, ll, sizeof(one));
  for (uint32_t i = 0; i < (1u << LL_BITS); ++i) {
    const uint32_t e1 = one[i];
    if (!(e1 & E_LIT)) continue;
    const uint32_t l1 = e1 & 15u;
    if (l1 >= (uint32_t)LL_BITS) continue;
    const uint32_t e2 = one[i >> l1];
    const uint32_t l2 = e2 & 15u;
    if ((e2 & E_LIT) && l1 + l2 <= (uint32_t)LL_BITS)
      ll[i] = E_LIT | E_LIT2 | (e1 & 0xFF00u) | ((e2 & 0xFF00u) << 8) | (l1 + l2);
  }
}

const Tables& fixed_tables() {
  static const Tables* t = [] {
    Tables* x = new Tables;
    uint8_t lens[288];
    for (int i = 0; i < 144; ++i) lens[i] = 8;
    for (int i = 144; i < 256; ++i) lens[i] = 9;
    for (int i = 256; i < 280; ++i) lens[i] = 7;
    for (int i = 280; i < 288; ++i) lens[i] = 8;
    build_table(lens, 288, LL_BITS, x->ll, LL_CAP, ll_payload);
    pair_literals(x->ll);
    uint8_t dl[32];
    for (int i = 0; i < 32; ++i) dl[i] = 5;
    build_table(dl, 32, D_BITS, x->d, D_CAP, d_payload);
    return x;
  }();
  return *t;
}

}  // namespace

// in: n_in bytes of a zlib stream followed by INFLATE_IN_SLACK readable bytes; out: capacity
// n_out + INFLATE_OUT_SLACK.  Succeeds only if the stream inflates to exactly n_out bytes and the
// Adler-32 trailer matches.  The bit reader runs up to 8 bytes ahead of the bits it has consumed, so
// `in` may legitimately pass the end of the stream by a few bytes; in_lim bounds how far.
bool inflate_zlib_fast(const uint8_t* in0, size_t n_in, uint8_t* out0, size_t n_out) {
  if (n_in < 6) return false;
  const unsigned cmf = in0[0], flg = in0[1];
  if ((cmf & 15u) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u != 0u || (flg & 32u)) return false;
  const uint8_t* in = in0 + 2;
  const uint8_t* const in_end = in0 + n_in;
  const uint8_t* const in_lim = in_end + 8;
  uint8_t* out = out0;
  uint8_t* const out_end = out0 + n_out;
  uint64_t bitbuf = 0;
  unsigned bitcnt = 0;
  Tables dyn;   // ~44 KB on the stack of the decoding thread

#define VO_REFILL() do { uint64_t w_; memcpy(&w_, in, 8); bitbuf |= w_ << bitcnt; in += (63u - bitcnt) >> 3; bitcnt |= 56u; } while (0)
#define VO_DROP(n_) do { bitbuf >>= (n_); bitcnt -= (unsigned)(n_); } while (0)
#define VO_BITS(n_) ((uint32_t)bitbuf & ((1u << (n_)) - 1u))
// one or two literals; the second byte is written unconditionally (it lands in the output slack at
// worst and is overwritten by whatever comes next)
#define VO_EMIT(e_) do { VO_DROP((e_) & 15u); out[0] = (uint8_t)((e_) >> 8); out[1] = (uint8_t)((e_) >> 16); out += 1 + (((e_) >> 27) & 1u); } while (0)

  for (;;) {
    if (in > in_lim) return false;
    VO_REFILL();
    const unsigned last = VO_BITS(1), type = ((uint32_t)bitbuf >> 1) & 3u;
    VO_DROP(3);
    const Tables* tb = nullptr;
    if (type == 0) {   // stored: back to byte granularity
      VO_DROP(bitcnt & 7u);
      const uint8_t* p = in - (bitcnt >> 3);
      if (p + 4 > in_end) return false;
      const unsigned len = p[0] | (p[1] << 8), nlen = p[2] | (p[3] << 8);
      if ((len ^ 0xFFFFu) != nlen) return false;
      p += 4;
      // a paired literal may have left `out` one byte past out_end: reject before the unsigned difference wraps
      if (out > out_end || (size_t)(in_end - p) < len || (size_t)(out_end - out) < len) return false;
      memcpy(out, p, len);
      out += len; in = p + len; bitbuf = 0; bitcnt = 0;
      if (last) break;
      continue;
    } else if (type == 1) {
      tb = &fixed_tables();
    } else if (type == 2) {
      const unsigned hlit = VO_BITS(5) + 257; VO_DROP(5);
      const unsigned hdist = VO_BITS(5) + 1; VO_DROP(5);
      const unsigned hclen = VO_BITS(4) + 4; VO_DROP(4);
      if (hlit > 286 || hdist > 30) return false;
      uint8_t pre[19] = {0};
      for (unsigned i = 0; i < hclen; ++i) {
        if (bitcnt < 3) { if (in > in_lim) return false; VO_REFILL(); }
        pre[PRE_ORDER[i]] = (uint8_t)VO_BITS(3); VO_DROP(3);
      }
      uint32_t pt[1 << PRE_BITS];
      if (!build_table(pre, 19, PRE_BITS, pt, 1 << PRE_BITS, pre_payload)) return false;
      uint8_t lens[286 + 30 + 138];
      unsigned i = 0;
      while (i < hlit + hdist) {
        if (in > in_lim) return false;
        VO_REFILL();
        const uint32_t e = pt[VO_BITS(PRE_BITS)];
        if (e & E_BAD) return false;
        VO_DROP(e & 15u);
        const unsigned sym = (e >> 8) & 31u;
        if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
        unsigned rep; uint8_t v = 0;
        if (sym == 16) { if (!i) return false; v = lens[i - 1]; rep = 3 + VO_BITS(2); VO_DROP(2); }
        else if (sym == 17) { rep = 3 + VO_BITS(3); VO_DROP(3); }
        else { rep = 11 + VO_BITS(7); VO_DROP(7); }
        if (i + rep > hlit + hdist) return false;
        memset(lens + i, v, rep); i += rep;
      }
      if (lens[256] == 0) return false;
      if (!build_table(lens, (int)hlit, LL_BITS, dyn.ll, LL_CAP, ll_payload)) return false;
      if (!build_table(lens + hlit, (int)hdist, D_BITS, dyn.d, D_CAP, d_payload)) return false;
      pair_literals(dyn.ll);
      tb = &dyn;
    } else {
      return false;
    }
    const uint32_t* const ll = tb->ll;
    const uint32_t* const dt = tb->d;
    for (;;) {
      if (in > in_lim) return false;
      VO_REFILL();   // >= 56 bits: one length/distance pair needs at most 48
      uint32_t e = ll[VO_BITS(LL_BITS)];
      if (e & E_LIT) {   // up to three literals per refill (3 x 15 bits)
        if (out >= out_end) return false;
        VO_EMIT(e);
        e = ll[VO_BITS(LL_BITS)];
        if (e & E_LIT) {
          if (out >= out_end) return false;
          VO_EMIT(e);
          e = ll[VO_BITS(LL_BITS)];
          if (e & E_LIT) {
            if (out >= out_end) return false;
            VO_EMIT(e);
            continue;
          }
        }
        VO_REFILL();   // adds high bits only: e is still the entry of the next symbol
      }
      if (e & E_SUB) {
        VO_DROP(e & 15u);
        e = ll[((e >> 8) & 0xFFFFFu) + VO_BITS((e >> 4) & 15u)];
        if (e & E_LIT) {
          if (out >= out_end) return false;
          VO_EMIT(e);
          continue;
        }
      }
      if (e & (E_EOB | E_BAD)) {
        if (e & E_BAD) return false;
        if (out > out_end) return false;   // end of block after a literal pair that overran the output
        VO_DROP(e & 15u);
        break;
      }
      VO_DROP(e & 15u);
      const unsigned xb = (e >> 4) & 15u;
      const size_t length = ((e >> 8) & 0xFFFFu) + VO_BITS(xb);
      VO_DROP(xb);
      uint32_t d = dt[VO_BITS(D_BITS)];
      if (d & E_SUB) {
        VO_DROP(d & 15u);
        d = dt[((d >> 8) & 0xFFFFFu) + VO_BITS((d >> 4) & 15u)];
      }
      if (d & E_BAD) return false;
      VO_DROP(d & 15u);
      const unsigned db = (d >> 4) & 15u;
      const size_t dist = ((d >> 8) & 0xFFFFu) + VO_BITS(db);
      VO_DROP(db);
      if (out > out_end || dist > (size_t)(out - out0) || length > (size_t)(out_end - out)) return false;   // (a literal pair may end one byte past out_end)
      const uint8_t* src = out - dist;
      uint8_t* dst = out;
      out += length;
      if (dist >= 8) {           // 8 bytes at a time; may write up to 7 bytes past the match (slack)
        memcpy(dst, src, 8);
        while ((dst += 8) < out) { src += 8; memcpy(dst, src, 8); }
      } else if (dist == 1) {    // a run of one byte (flat image regions after filtering)
        const uint64_t v = 0x0101010101010101ull * *src;
        memcpy(dst, &v, 8);
        if (length > 8) memset(dst + 8, (int)*src, length - 8);
      } else {
        do { *dst++ = *src++; } while (dst < out);
      }
    }
    if (last) break;
  }
#undef VO_REFILL
#undef VO_DROP
#undef VO_BITS
#undef VO_EMIT
  if (out != out_end) return false;
  const uint8_t* p = in - (bitcnt >> 3);   // first byte not touched by the bit reader = the Adler-32 trailer
  if (p + 4 > in_end) return false;
  const uint32_t want = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
  uint32_t a = 1;
#if defined(__x86_64__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) return adler32_avx2(a, out0, n_out) == want;
#endif
  for (size_t off = 0; off < n_out; off += (1u << 30)) {
    const size_t m = n_out - off < (1u << 30) ? n_out - off : (1u << 30);
    a = (uint32_t)adler32(a, out0 + off, (uInt)m);
  }
  return a == want;
}

}  // namespace vo
