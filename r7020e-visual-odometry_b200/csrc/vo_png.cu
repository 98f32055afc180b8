// vo_png.cu -- device-side PNG decode for the input stage (VO.m:16-17, 71-72).
//
// On a host with few cores per GPU the PNG inflate bounds the end-to-end rate (a core decodes 400-700 KITTI frames
// per second, the GPU consumes 9 k images per second).  Here the compressed files go to the device as they are
// (about half the bytes of the decoded frames) and two kernels turn them into the [n][rows][cols] batch buffer:
//
//   png_inflate_kernel    one warp per image.  DEFLATE is a serial bit stream, so lane 0 walks the symbols with the
//                         same two-level tables as the host decoder (vo_inflate.cu: 11-bit first level, entries that
//                         decode two literals at once, 8-bit distance table), held in shared memory; the other lanes
//                         only help to clear / pair the tables.  The parallelism is across the images of a batch (66)
//                         and across the batches in flight; a warp-sized block costs next to nothing beside the SIFT
//                         kernels of another batch running on the same SMs.
//   png_unfilter_kernel   one warp per image: the five PNG row filters as a 32-row skewed wavefront (lane k works on
//                         row y0 + k at column t - k, its "up" and "up-left" pixels come from lane k - 1 by shuffle),
//                         plus the Adler-32 of the inflated bytes (per-row sums, combined in row order) checked
//                         against the stream's trailer.
//
// Scope: what KITTI odometry ships -- 8-bit grayscale, non-interlaced.  Malformed input ends in a per-image error
// status, never in an out-of-bounds access (every read and write is bounds-checked, every loop consumes input).
#include "vo_internal.h"
#include "vo_ptx.cuh"
#include <thread>

namespace vo {

namespace png {
constexpr int LL_BITS = 11, D_BITS = 8, PRE_BITS = 7;
// second-level capacity (entries).  FULL covers every legal code; SMALL covers what image data produces (a handful of
// codes longer than the first level) in a third of the shared memory, so that a decode block does not take the room of
// two blur blocks on its SM while it walks its stream.  A stream that does not fit SMALL is decoded again with FULL.
constexpr int LL_SUB_FULL = 288 * 16, D_SUB_FULL = 32 * 128, LL_SUB_SMALL = 1024, D_SUB_SMALL = 256;
constexpr uint32_t E_LIT = 1u << 31, E_EOB = 1u << 30, E_SUB = 1u << 29, E_BAD = 1u << 28, E_LIT2 = 1u << 27;

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_pre_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

enum { ST_OK = 0, ST_HEADER = 1, ST_BLOCK = 2, ST_CODE = 3, ST_OVERRUN = 4, ST_DIST = 5, ST_LENGTH = 6, ST_ADLER = 7, ST_FILTER = 8, ST_TRUNC = 9,
       ST_BIG = 10 /* the code tables need the full-capacity kernel */ };

struct Job { uint32_t in_off, in_len; };   // zlib stream inside the staging buffer (16-byte aligned offset)

__device__ __forceinline__ uint32_t ll_payload(int s) {
  if (s < 256) return E_LIT | ((uint32_t)s << 8);
  if (s == 256) return E_EOB;
  if (s <= 285) return ((uint32_t)c_len_base[s - 257] << 8) | ((uint32_t)c_len_extra[s - 257] << 4);
  return E_BAD;
}
__device__ __forceinline__ uint32_t d_payload(int s) {
  if (s < 30) return ((uint32_t)c_dist_base[s] << 8) | ((uint32_t)c_dist_extra[s] << 4);
  return E_BAD;
}
__device__ __forceinline__ uint32_t bit_reverse(uint32_t c, int len) { return __brev(c) >> (32 - len); }

// Canonical code lengths -> two-level table (same layout as vo_inflate.cu).  kind: 0 literal/length, 1 distance,
// 2 code-length code.  Called by lane 0; `table` was cleared to E_BAD | 1 by the whole warp.
// returns ST_OK, ST_CODE (over-subscribed code) or ST_BIG (second level does not fit `cap`)
__device__ int build_table(const uint8_t* lens, int n, int tb, uint32_t* table, int cap, int kind, uint8_t* sub_bits) {
  int count[16];
  for (int l = 0; l < 16; ++l) count[l] = 0;
  for (int s = 0; s < n; ++s) ++count[lens[s]];
  count[0] = 0;
  int left = 1;
  for (int l = 1; l <= 15; ++l) { left = (left << 1) - count[l]; if (left < 0) return ST_CODE; }
  uint32_t next[16], nx[16]; uint32_t code = 0;
  next[0] = nx[0] = 0;
  for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = nx[l] = code; }
  const int primary = 1 << tb;
  bool any_long = false;
  for (int l = tb + 1; l <= 15; ++l) any_long = any_long || count[l];
  if (any_long) {
    for (int i = 0; i < primary; ++i) sub_bits[i] = 0;
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
    const uint32_t pay = kind == 0 ? ll_payload(s) : (kind == 1 ? d_payload(s) : ((uint32_t)s << 8));
    if (l <= tb) {
      const uint32_t e = pay | (uint32_t)l;
      for (uint32_t k = rev; k < (uint32_t)primary; k += 1u << l) table[k] = e;
    } else {
      const uint32_t pfx = rev & (uint32_t)(primary - 1);
      uint32_t link = table[pfx];
      if (!(link & E_SUB)) {
        const int sb = sub_bits[pfx];
        if (free_at + (1 << sb) > cap) return ST_BIG;
        link = E_SUB | ((uint32_t)free_at << 8) | ((uint32_t)sb << 4) | (uint32_t)tb;
        table[pfx] = link;
        for (int k = 0; k < (1 << sb); ++k) table[free_at + k] = E_BAD | 1u;
        free_at += 1 << sb;
      }
      const int sb = (int)((link >> 4) & 15u);
      uint32_t* sub = table + ((link >> 8) & 0xFFFFFu);
      const uint32_t e = pay | (uint32_t)(l - tb);
      for (uint32_t k = rev >> tb; k < (1u << sb); k += 1u << (l - tb)) sub[k] = e;
    }
  }
  return ST_OK;
}

template <int LL_SUB, int D_SUB>
struct Smem {
  static constexpr int LL_CAP = (1 << LL_BITS) + LL_SUB, D_CAP = (1 << D_BITS) + D_SUB;
  uint32_t ll[LL_CAP];            // literal/length table (first level with paired literals)
  uint32_t d[D_CAP];
  uint32_t pre[1 << PRE_BITS];
  uint8_t lens[320 + 140];
  uint8_t sub_bits[1 << LL_BITS];
};

// The bit reader.  A 64-bit shift register costs three or four instructions per shift on this machine, and a serial
// decoder is bound by exactly that dependent chain, so the reader keeps two 32-bit words of the stream (w0 = the word
// holding the next bit, w1 = the one after) and a third, already requested, behind them: the next 32 bits are ONE
// funnel shift, consuming n bits is an add, and the load of the following word is issued a whole word (two or three
// symbols) before its value is needed.  Streams start on a 16-byte boundary of the staging buffer.
struct Bits {
  const uint32_t* words; uint32_t wi, wlim;   // wi = index of w0; wlim = last word index that may be loaded
  uint32_t w0, w1, w2, w3, bp;                // bp = position of the next bit inside w0 (0..31); w3 = the word in flight
  bool over;
  __device__ __forceinline__ uint32_t load(uint32_t i) { if (i > wlim) { over = true; return 0u; } return words[i]; }
  __device__ __forceinline__ void seek(uint32_t byte_pos) {
    wi = byte_pos >> 2; bp = (byte_pos & 3u) * 8u;
    w0 = load(wi); w1 = load(wi + 1); w2 = load(wi + 2); w3 = load(wi + 3);
  }
  __device__ __forceinline__ uint32_t peek32() const { return __funnelshift_r(w0, w1, bp); }   // bp < 32
  __device__ __forceinline__ void consume(uint32_t n) {                                        // n <= 32
    bp += n;
    if (bp >= 32u) { bp -= 32u; w0 = w1; w1 = w2; w2 = w3; ++wi; w3 = load(wi + 3); }
  }
  __device__ __forceinline__ uint32_t byte_pos() const { return wi * 4u + ((bp + 7u) >> 3); }   // first byte not touched
};

// grid = images (or the images listed in `only`), block = one warp.  raw: [n][raw_stride] filtered scanlines
// (rows * (cols + 1) bytes used).
template <int LL_SUB, int D_SUB>
__global__ void __launch_bounds__(32)
png_inflate_kernel(const uint8_t* __restrict__ staging, const Job* __restrict__ jobs, const int* __restrict__ only,
                   uint8_t* __restrict__ raw_base, size_t raw_stride, uint32_t n_raw, int* __restrict__ status,
                   uint32_t* __restrict__ adler_want) {
  using SM = Smem<LL_SUB, D_SUB>;
  extern __shared__ __align__(16) uint8_t png_smem[];
  SM& sm = *reinterpret_cast<SM*>(png_smem);
  constexpr int LL_CAP = SM::LL_CAP, D_CAP = SM::D_CAP;
  const int img = only ? only[blockIdx.x] : blockIdx.x, lane = threadIdx.x;
  const Job job = jobs[img];
  const uint8_t* in = staging + job.in_off;
  const uint32_t n_in = job.in_len;
  uint8_t* out0 = raw_base + (size_t)img * raw_stride;
  int err = ST_OK;
  Bits br; br.words = reinterpret_cast<const uint32_t*>(in); br.wlim = (n_in + 8u) / 4u + 3u; br.over = false;
  br.wi = 0; br.bp = 0; br.w0 = br.w1 = br.w2 = br.w3 = 0;
  uint32_t out = 0;
  bool last = false;
  if (lane == 0) {
    if (n_in < 6u) err = ST_HEADER;
    else {
      const unsigned cmf = in[0], flg = in[1];
      if ((cmf & 15u) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u != 0u || (flg & 32u)) err = ST_HEADER;
      br.seek(2);
    }
  }
  for (;;) {
    // ---- block header (lane 0), then the tables
    int type = -1;
    if (lane == 0 && err == ST_OK) {
      if (br.over) err = ST_TRUNC;
      else {
        const uint32_t hb = br.peek32();
        last = (hb & 1u) != 0; type = (int)((hb >> 1) & 3u); br.consume(3);
        if (type == 0) {   // stored block: byte granularity
          uint32_t p = br.byte_pos();
          if (p + 4 > n_in) err = ST_TRUNC;
          else {
            const uint32_t len = in[p] | ((uint32_t)in[p + 1] << 8), nlen = in[p + 2] | ((uint32_t)in[p + 3] << 8);
            p += 4;
            if ((len ^ 0xFFFFu) != nlen) err = ST_BLOCK;
            else if (p + len > n_in || out + len > n_raw) err = ST_OVERRUN;
            else {
              for (uint32_t k = 0; k < len; ++k) out0[out + k] = in[p + k];
              out += len;
              br.seek(p + len);
            }
          }
        } else if (type == 3) err = ST_BLOCK;
      }
    }
    type = __shfl_sync(0xffffffffu, type, 0);
    if (__shfl_sync(0xffffffffu, err, 0) != ST_OK) break;
    if (type == 0) {
      if (__shfl_sync(0xffffffffu, (int)last, 0)) break;
      continue;
    }
    // clear the first levels (whole warp)
    for (int i = lane; i < (1 << LL_BITS); i += 32) sm.ll[i] = E_BAD | 1u;
    for (int i = lane; i < (1 << D_BITS); i += 32) sm.d[i] = E_BAD | 1u;
    for (int i = lane; i < (1 << PRE_BITS); i += 32) sm.pre[i] = E_BAD | 1u;
    __syncwarp();
    if (lane == 0) {
      if (type == 1) {
        for (int i = 0; i < 144; ++i) sm.lens[i] = 8;
        for (int i = 144; i < 256; ++i) sm.lens[i] = 9;
        for (int i = 256; i < 280; ++i) sm.lens[i] = 7;
        for (int i = 280; i < 288; ++i) sm.lens[i] = 8;
        for (int i = 0; i < 32; ++i) sm.lens[288 + i] = 5;
        err = build_table(sm.lens, 288, LL_BITS, sm.ll, LL_CAP, 0, sm.sub_bits);
        if (err == ST_OK) err = build_table(sm.lens + 288, 32, D_BITS, sm.d, D_CAP, 1, sm.sub_bits);
      } else {
        uint32_t hb = br.peek32();
        const unsigned hlit = (hb & 31u) + 257, hdist = ((hb >> 5) & 31u) + 1, hclen = ((hb >> 10) & 15u) + 4;
        br.consume(14);
        if (hlit > 286 || hdist > 30) err = ST_CODE;
        for (int i = 0; i < 19; ++i) sm.lens[320 + i] = 0;
        for (unsigned i = 0; i < hclen && err == ST_OK; ++i) {
          sm.lens[320 + c_pre_order[i]] = (uint8_t)(br.peek32() & 7u); br.consume(3);
        }
        if (br.over) err = ST_TRUNC;
        if (err == ST_OK) err = build_table(sm.lens + 320, 19, PRE_BITS, sm.pre, 1 << PRE_BITS, 2, sm.sub_bits);
        unsigned i = 0;
        while (err == ST_OK && i < hlit + hdist) {
          if (br.over) { err = ST_TRUNC; break; }
          const uint32_t bits = br.peek32();
          const uint32_t e = sm.pre[bits & ((1u << PRE_BITS) - 1u)];
          if (e & E_BAD) { err = ST_CODE; break; }
          const uint32_t l = e & 15u;
          const unsigned sym = (e >> 8) & 31u;
          if (sym < 16) { br.consume(l); sm.lens[i++] = (uint8_t)sym; continue; }
          unsigned rep; uint8_t v = 0;
          if (sym == 16) { if (!i) { err = ST_CODE; break; } v = sm.lens[i - 1]; rep = 3 + ((bits >> l) & 3u); br.consume(l + 2); }
          else if (sym == 17) { rep = 3 + ((bits >> l) & 7u); br.consume(l + 3); }
          else { rep = 11 + ((bits >> l) & 127u); br.consume(l + 7); }
          if (i + rep > hlit + hdist) { err = ST_CODE; break; }
          for (unsigned k = 0; k < rep; ++k) sm.lens[i + k] = v;
          i += rep;
        }
        if (err == ST_OK && sm.lens[256] == 0) err = ST_CODE;
        if (err == ST_OK) {
          // the distance lengths follow the literal/length lengths: move them so both tables see their own array
          for (unsigned k = 0; k < hdist; ++k) sm.lens[320 + 32 + k] = sm.lens[hlit + k];
          err = build_table(sm.lens, (int)hlit, LL_BITS, sm.ll, LL_CAP, 0, sm.sub_bits);
          if (err == ST_OK) err = build_table(sm.lens + 320 + 32, (int)hdist, D_BITS, sm.d, D_CAP, 1, sm.sub_bits);
        }
      }
    }
    if (__shfl_sync(0xffffffffu, err, 0) != ST_OK) break;
    __syncwarp();
    // pair literals, in place: where two consecutive literal codes fit into the first-level index, the entry decodes
    // both.  Entry i reads entry i >> l1 (l1 >= 1), which lies in a lower 32-entry chunk for every i >= 64 and in the
    // chunk itself or below otherwise: chunks are visited from the top by the whole warp, the two lowest by lane 0.
    for (int base = (1 << LL_BITS) - 32; base >= 0; base -= 32) {
      const bool serial = base < 64;
      for (int k = serial ? 31 : lane; k >= 0 && (serial ? lane == 0 : k == lane); k = serial ? k - 1 : -1) {
        const uint32_t i = (uint32_t)(base + k);
        const uint32_t e1 = sm.ll[i];
        if (!(e1 & E_LIT) || (e1 & E_LIT2)) continue;
        const uint32_t l1 = e1 & 15u;
        if (l1 >= (uint32_t)LL_BITS) continue;
        const uint32_t e2 = sm.ll[i >> l1];
        const uint32_t l2 = e2 & 15u;
        if ((e2 & E_LIT) && !(e2 & E_LIT2) && l1 + l2 <= (uint32_t)LL_BITS)
          sm.ll[i] = E_LIT | E_LIT2 | (e1 & 0xFF00u) | ((e2 & 0xFF00u) << 8) | (l1 + l2);
      }
      __syncwarp();
    }
    // ---- the symbols of this block (lane 0).  A serial decoder runs at the latency of its dependent chain, and on this
    // machine a conditional branch of a lone warp costs about as much as a shared-memory load (ncu: the `wait` samples
    // sit on the branches), so the literal path is ONE loop-back branch: funnel shift, table look-up, two byte stores
    // (the second lands in the next symbol's place, or in the slack, when the entry holds one literal) and a branch-free
    // advance of the reader (selects and a predicated load; a read past the stream re-reads its last word and is
    // caught at the next block boundary).
    if (lane == 0) {
      const uint32_t* ll = sm.ll;
      const uint32_t* dt = sm.d;
      const uint32_t* const wp = br.words;
      const uint32_t wlim = br.wlim, lim = n_raw;
      uint32_t w0 = br.w0, w1 = br.w1, w2 = br.w2, w3 = br.w3, bp = br.bp, wi = br.wi;
      uint32_t o = out;
      const uint32_t ll_s = smem_u32(sm.ll);
      uint8_t* obase = out0;
      asm volatile("" : "+l"(obase));                      // one register pair: every address is a single wide add
// consume n_ bits and form the next 32: both candidate funnel shifts (the shift count wraps mod 32) are independent of
// whether a word boundary was crossed, so the only serial steps are add -> shift -> select; the rotation of the word
// registers and the load of the word after next happen off that chain, and the loaded value is first used two
// boundaries later
#define VO_PNG_ADV(n_) do { bp += (n_); const bool adv_ = bp >= 32u; \
                            const uint32_t ba_ = __funnelshift_r(w0, w1, bp), bb_ = __funnelshift_r(w1, w2, bp); \
                            bits = adv_ ? bb_ : ba_; \
                            if (adv_) { w0 = w1; w1 = w2; w2 = w3; w3 = wp[min(wi + 4u, wlim)]; ++wi; bp -= 32u; } } while (0)
      uint32_t bits;
      bits = __funnelshift_r(w0, w1, bp);
      uint32_t e = ll[bits & ((1u << LL_BITS) - 1u)];
      for (;;) {
        while ((int)e < 0 && o < lim) {                    // E_LIT: one or two literals
          // A lone warp issues in order, so whatever stalls delays everything behind it: the look-up of the NEXT symbol
          // is issued first (volatile asm keeps the order), and the stores of this one, the rotation of the word
          // registers and the refill (predicated, straight into w3, whose value is not read before the boundary after
          // next) run in the shadow of that shared-memory load.
          const uint32_t ecur = e;
          bp += ecur & 15u;
          const uint32_t ba = __funnelshift_r(w0, w1, bp), bb = __funnelshift_r(w1, w2, bp);
          const uint32_t adv = bp >> 5;                      // 0 or 1
          bits = adv ? bb : ba;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(ll_s + ((bits << 2) & (((1u << LL_BITS) - 1u) << 2))));
          uint8_t* p = obase + o;
          p[0] = (uint8_t)(ecur >> 8); p[1] = (uint8_t)(ecur >> 16);
          o += 1u + ((ecur >> 27) & 1u);
          if (adv) { w0 = w1; w1 = w2; w2 = w3; ++wi; bp -= 32u; }
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.global.nc.u32 %0, [%1];\n\t}"
                       : "+r"(w3) : "l"(wp + min(wi + 3u, wlim)), "r"(adv));
        }
        if ((int)e < 0) { err = ST_OVERRUN; break; }       // a literal with no room left
        uint32_t used = 0;
        if (e & E_SUB) {
          used = e & 15u;                                  // the first-level bits
          e = ll[((e >> 8) & 0xFFFFFu) + ((bits >> used) & ((1u << ((e >> 4) & 15u)) - 1u))];
          if (e & E_LIT) {
            if (o >= lim) { err = ST_OVERRUN; break; }
            obase[o++] = (uint8_t)(e >> 8);
            VO_PNG_ADV(used + (e & 15u));
            e = ll[bits & ((1u << LL_BITS) - 1u)];
            continue;
          }
        }
        if (e & (E_EOB | E_BAD)) {
          if (e & E_BAD) { err = ST_CODE; break; }
          if (o > lim) { err = ST_OVERRUN; break; }
          VO_PNG_ADV(used + (e & 15u));
          break;
        }
        used += e & 15u;
        const uint32_t xb = (e >> 4) & 15u;                 // code (<= 15) + extra bits (<= 5) sit in the same 32 bits
        const uint32_t length = ((e >> 8) & 0xFFFFu) + ((bits >> used) & ((1u << xb) - 1u));
        VO_PNG_ADV(used + xb);                             // bits = distance code (<= 15) + extra bits (<= 13)
        uint32_t d = dt[bits & ((1u << D_BITS) - 1u)];
        used = 0;
        if (d & E_SUB) {
          used = d & 15u;
          d = dt[((d >> 8) & 0xFFFFFu) + ((bits >> used) & ((1u << ((d >> 4) & 15u)) - 1u))];
        }
        if (d & E_BAD) { err = ST_CODE; break; }
        used += d & 15u;
        const uint32_t db = (d >> 4) & 15u;
        const uint32_t dist = ((d >> 8) & 0xFFFFu) + ((bits >> used) & ((1u << db) - 1u));
        VO_PNG_ADV(used + db);
        if (o > lim || dist > o) { err = ST_DIST; break; }
        if (length > lim - o) { err = ST_LENGTH; break; }
        const uint8_t* src = obase + (o - dist);           // (a 32 KB window in shared memory was measured: no faster)
        uint8_t* dst = obase + o;
        if (dist >= length) {                              // no overlap: all loads first, then the stores
          uint32_t k = 0;
          for (; k + 4 <= length; k += 4) {
            const uint8_t a0 = src[k], a1 = src[k + 1], a2 = src[k + 2], a3 = src[k + 3];
            dst[k] = a0; dst[k + 1] = a1; dst[k + 2] = a2; dst[k + 3] = a3;
          }
          for (; k < length; ++k) dst[k] = src[k];
        } else {
          for (uint32_t k = 0; k < length; ++k) dst[k] = src[k];
        }
        o += length;
        e = ll[bits & ((1u << LL_BITS) - 1u)];
      }
#undef VO_PNG_ADV
      br.w0 = w0; br.w1 = w1; br.w2 = w2; br.w3 = w3; br.bp = bp; br.wi = wi;
      if (wi * 4u > n_in + 8u) br.over = true;             // the reader ran past the stream (it then re-read its last word)
      if (br.over && err == ST_OK) err = ST_TRUNC;
      out = o;
    }
    if (__shfl_sync(0xffffffffu, err, 0) != ST_OK) break;
    if (__shfl_sync(0xffffffffu, (int)last, 0)) break;
  }
  if (lane == 0) {
    if (err == ST_OK && out != n_raw) err = ST_LENGTH;
    if (err == ST_OK) {
      const uint32_t p = br.byte_pos();                   // first byte after the last block = the Adler-32 trailer
      if (p + 4 > n_in) err = ST_TRUNC;
      else adler_want[img] = ((uint32_t)in[p] << 24) | ((uint32_t)in[p + 1] << 16) | ((uint32_t)in[p + 2] << 8) | in[p + 3];
    }
    status[img] = err;
  }
}

__device__ __forceinline__ int paeth(int a, int b, int c) {
  const int p = b - c, q = a - c;
  const int pa = abs(p), pb = abs(q), pc = abs(p + q);
  const int bc = pb <= pc ? b : c;
  return (pa <= pb && pa <= pc) ? a : bc;
}

// grid = images, block = one warp.  out: [n][rows][cols] (image i at out + i * out_stride).
__global__ void __launch_bounds__(32)
png_unfilter_kernel(const uint8_t* __restrict__ raw_base, size_t raw_stride, int rows, int cols, uint8_t* __restrict__ out_base,
                    size_t out_stride, int* __restrict__ status, const uint32_t* __restrict__ adler_want,
                    const int* __restrict__ only) {
  extern __shared__ __align__(16) uint8_t unf_smem[];
  uint8_t* lastrow = unf_smem;                                                   // [cols]: the finished row above the group
  uint32_t* row_a = reinterpret_cast<uint32_t*>(unf_smem + ((cols + 15) & ~15));   // [rows] per-row byte sums
  uint32_t* row_b = row_a + rows;                                                // [rows] per-row position-weighted sums
  const int img = only ? only[blockIdx.x] : blockIdx.x, lane = threadIdx.x;
  if (status[img] != ST_OK) return;
  const uint8_t* raw = raw_base + (size_t)img * raw_stride;
  uint8_t* out = out_base + (size_t)img * out_stride;
  const size_t stride = (size_t)cols + 1;
  int bad = 0;
  for (int y0 = 0; y0 < rows; y0 += 32) {
    const int y = y0 + lane;
    const bool active = y < rows;
    const uint8_t* src = raw + stride * (active ? y : 0);
    const int ft = active ? src[0] : 0;
    if (active && ft > 4) bad = 1;
    int a = 0, c = 0, cur = 0;
    uint32_t sa = (uint32_t)ft, sb = (uint32_t)ft;
    for (int t = 0; t < cols + 31; ++t) {
      const int x = t - lane;
      int b = __shfl_up_sync(0xffffffffu, cur, 1);        // lane k - 1 finished column x one step ago
      const bool in_row = active && x >= 0 && x < cols;
      if (lane == 0) b = (y0 > 0 && in_row) ? lastrow[x] : 0;
      if (in_row) {
        const int s = src[1 + x];
        int v = s;
        if (ft == 1) v = s + a;
        else if (ft == 2) v = s + b;
        else if (ft == 3) v = s + ((a + b) >> 1);
        else if (ft == 4) v = s + paeth(a, b, c);
        v &= 255;
        out[(size_t)y * cols + x] = (uint8_t)v;
        if (lane == 31) lastrow[x] = (uint8_t)v;          // 31 columns behind lane 0's reads of the same array
        c = b; a = v; cur = v;
        sa += (uint32_t)s; sb += sa;
      }
    }
    if (active) { row_a[y] = sa; row_b[y] = sb; }
    __syncwarp();
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    if (bad) { status[img] = ST_FILTER; return; }
    unsigned long long A = 1, B = 0;
    const unsigned long long L = (unsigned long long)stride;
    for (int y = 0; y < rows; ++y) {
      B = (B + L * A + row_b[y]) % 65521ull;
      A = (A + row_a[y]) % 65521ull;
    }
    if ((uint32_t)((B << 16) | A) != adler_want[img]) status[img] = ST_ADLER;
  }
}

static inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

static const char* status_text(int s) {
  switch (s) {
    case ST_HEADER: return "bad zlib header";
    case ST_BLOCK: return "bad DEFLATE block";
    case ST_CODE: return "bad Huffman code";
    case ST_OVERRUN: return "more data than the image holds";
    case ST_DIST: return "match distance beyond the start of the data";
    case ST_LENGTH: return "the stream does not inflate to rows * (cols + 1) bytes";
    case ST_ADLER: return "Adler-32 mismatch";
    case ST_FILTER: return "bad row filter type";
    case ST_TRUNC: return "truncated stream";
    case ST_BIG: return "code tables exceed the decoder's capacity";
    default: return "unknown";
  }
}
}  // namespace png

// n PNG files held in host memory -> out_dev[n][rows][cols] on the context's stream.  Returns after the decode has
// finished (one synchronisation), with the first failing image reported through vo_last_error().
int png_decode_batch_device(vo_ctx* ctx, const uint8_t* const* files, const size_t* sizes, int n, int rows, int cols,
                            uint8_t* out_dev) {
  using namespace png;
  cudaStream_t st = ctx->stream;
  // 1. host: check the headers, gather the IDAT payloads of every file into one pinned staging buffer
  size_t total = 0;
  for (int i = 0; i < n; ++i) total += ((sizes[i] + 15) & ~(size_t)15) + 32;
  uint8_t* hstage; VO_TRY(pin_buf(ctx, "png_stage", total + 64, &hstage));
  Job* hjobs; VO_TRY(pin_buf(ctx, "png_jobs", (size_t)n, &hjobs));
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    const uint8_t* f = files[i]; const size_t sz = sizes[i];
    if (sz < 33 || memcmp(f, sig, 8) != 0 || be32(f + 8) != 13 || memcmp(f + 12, "IHDR", 4) != 0) { set_error("png %d: bad signature / IHDR", i); return VO_ERR_ARG; }
    if ((int)be32(f + 16) != cols || (int)be32(f + 20) != rows) { set_error("png %d: image is %u x %u, expected %d x %d", i, be32(f + 20), be32(f + 16), rows, cols); return VO_ERR_ARG; }
    if (f[24] != 8 || f[25] != 0 || f[28] != 0) { set_error("png %d: only 8-bit grayscale non-interlaced files are supported (depth %d, colour type %d, interlace %d)", i, f[24], f[25], f[28]); return VO_ERR_ARG; }
    size_t pos = 8, len_idat = 0;
    bool end = false;
    while (pos + 12 <= sz && !end) {
      const uint32_t len = be32(f + pos);
      if (pos + 12 + (size_t)len > sz) { set_error("png %d: truncated chunk", i); return VO_ERR_ARG; }
      if (memcmp(f + pos + 4, "IDAT", 4) == 0) { memcpy(hstage + off + len_idat, f + pos + 8, len); len_idat += len; }
      else if (memcmp(f + pos + 4, "IEND", 4) == 0) end = true;
      pos += 12 + (size_t)len;
    }
    memset(hstage + off + len_idat, 0, 32);                      // the bit reader may run a few bytes ahead
    hjobs[i].in_off = (uint32_t)off; hjobs[i].in_len = (uint32_t)len_idat;
    off += ((len_idat + 15) & ~(size_t)15) + 32;
  }
  // 2. device
  const uint32_t n_raw = (uint32_t)rows * (uint32_t)(cols + 1);
  const size_t raw_stride = ((size_t)n_raw + 8 + 15) & ~(size_t)15;
  uint8_t* dstage; VO_TRY(dev_buf(ctx, "png_dstage", total + 64, &dstage));
  Job* djobs; VO_TRY(dev_buf(ctx, "png_djobs", (size_t)n, &djobs));
  uint8_t* draw; VO_TRY(dev_buf(ctx, "png_raw", (size_t)n * raw_stride, &draw));
  int* dstat; VO_TRY(dev_buf(ctx, "png_stat", (size_t)2 * n, &dstat));          // status[n], adler_want[n]
  int* hstat; VO_TRY(pin_buf(ctx, "png_hstat", (size_t)n, &hstat));
  VO_CUDA(cudaMemcpyAsync(dstage, hstage, off, cudaMemcpyHostToDevice, st));
  VO_CUDA(cudaMemcpyAsync(djobs, hjobs, (size_t)n * sizeof(Job), cudaMemcpyHostToDevice, st));
  using SmSmall = Smem<LL_SUB_SMALL, D_SUB_SMALL>;
  using SmFull = Smem<LL_SUB_FULL, D_SUB_FULL>;
  VO_TRY(ensure_dyn_smem_of(png_inflate_kernel<LL_SUB_FULL, D_SUB_FULL>, sizeof(SmFull)));
  const size_t unf_smem = (size_t)((cols + 15) & ~15) + (size_t)2 * rows * sizeof(uint32_t);
  VO_TRY(ensure_dyn_smem_of(png_unfilter_kernel, unf_smem));
  uint32_t* dwant = reinterpret_cast<uint32_t*>(dstat + n);
  {
    ProfScope ps(ctx, st, "png_inflate", (double)off + (double)n * n_raw);
    png_inflate_kernel<LL_SUB_SMALL, D_SUB_SMALL><<<n, 32, sizeof(SmSmall), st>>>(dstage, djobs, nullptr, draw, raw_stride, n_raw, dstat, dwant);
  }
  {
    ProfScope ps(ctx, st, "png_unfilter", (double)n * (n_raw + (double)rows * cols));
    png_unfilter_kernel<<<n, 32, unf_smem, st>>>(draw, raw_stride, rows, cols, out_dev, (size_t)rows * cols, dstat, dwant, nullptr);
  }
  VO_CUDA(cudaMemcpyAsync(hstat, dstat, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
  VO_CUDA(cudaStreamSynchronize(st));
  // streams whose code tables did not fit the small second level: once more with the full capacity
  int n_big = 0;
  for (int i = 0; i < n; ++i) n_big += hstat[i] == ST_BIG;
  if (n_big) {
    int* hbig; VO_TRY(pin_buf(ctx, "png_hbig", (size_t)n, &hbig));
    int* dbig; VO_TRY(dev_buf(ctx, "png_dbig", (size_t)n, &dbig));
    for (int i = 0, k = 0; i < n; ++i) if (hstat[i] == ST_BIG) hbig[k++] = i;
    VO_CUDA(cudaMemcpyAsync(dbig, hbig, (size_t)n_big * sizeof(int), cudaMemcpyHostToDevice, st));
    png_inflate_kernel<LL_SUB_FULL, D_SUB_FULL><<<n_big, 32, sizeof(SmFull), st>>>(dstage, djobs, dbig, draw, raw_stride, n_raw, dstat, dwant);
    png_unfilter_kernel<<<n_big, 32, unf_smem, st>>>(draw, raw_stride, rows, cols, out_dev, (size_t)rows * cols, dstat, dwant, dbig);
    ctx->kernel_launches += 2;
    VO_CUDA(cudaMemcpyAsync(hstat, dstat, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    VO_CUDA(cudaStreamSynchronize(st));
  }
  for (int i = 0; i < n; ++i)
    if (hstat[i] != ST_OK) { set_error("png %d: %s", i, status_text(hstat[i])); return VO_ERR_ARG; }
  return VO_OK;
}

}  // namespace vo

using namespace vo;

extern "C" {

int vo_png_decode_batch_dev(vo_ctx* ctx, const uint8_t* const* files, const size_t* sizes, int n, int rows, int cols,
                            uint8_t* out_dev) {
  VO_CHECK_ARG(ctx && files && sizes && out_dev, "null argument");
  VO_CHECK_ARG(n >= 0 && rows > 0 && cols > 0 && rows <= 2047 && cols <= 4095, "bad size");
  if (n == 0) return VO_OK;
  VO_CUDA(cudaSetDevice(ctx->device));
  ctx->kernel_launches += 2;
  return png_decode_batch_device(ctx, files, sizes, n, rows, cols, out_dev);
}

// Reads n files with n_threads host threads (<= 0: one per hardware thread, at most 8: reading is cheap) and decodes
// them on the device.
int vo_png_read_batch_dev(vo_ctx* ctx, const char* const* paths, int n, int rows, int cols, uint8_t* out_dev, int n_threads) {
  VO_CHECK_ARG(ctx && paths && out_dev && n >= 0, "bad argument");
  if (n == 0) return VO_OK;
  std::vector<std::vector<uint8_t>> bufs((size_t)n);
  std::vector<int> rc((size_t)n, VO_OK);
  if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
  if (n_threads > 8) n_threads = 8;
  if (n_threads > n) n_threads = n;
  if (n_threads < 1) n_threads = 1;
  auto work = [&](int t) {
    for (int i = t; i < n; i += n_threads) {
      FILE* fp = fopen(paths[i], "rb");
      if (!fp) { rc[i] = VO_ERR_ARG; continue; }
      fseek(fp, 0, SEEK_END);
      const long sz = ftell(fp);
      fseek(fp, 0, SEEK_SET);
      bufs[i].resize(sz > 0 ? (size_t)sz : 0);
      const size_t got = sz > 0 ? fread(bufs[i].data(), 1, (size_t)sz, fp) : 0;
      fclose(fp);
      if (got != bufs[i].size()) rc[i] = VO_ERR_ARG;
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  std::vector<const uint8_t*> ptrs((size_t)n); std::vector<size_t> sizes((size_t)n);
  for (int i = 0; i < n; ++i) {
    if (rc[i] != VO_OK) { set_error("cannot read %s", paths[i]); return VO_ERR_ARG; }
    ptrs[i] = bufs[i].data(); sizes[i] = bufs[i].size();
  }
  return vo_png_decode_batch_dev(ctx, ptrs.data(), sizes.data(), n, rows, cols, out_dev);
}

}  // extern "C"
