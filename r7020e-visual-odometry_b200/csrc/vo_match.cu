// vo_match.cu -- matchFeatures(f1, f2) on B200 (replaces VO.m:87, 283, 293, 311, 323).
//
// Pipeline (all on one stream, no host round trip between kernels):
//   1. match_prep_rows128_kernel   rows -> 1/||row|| (oracle fmaf order), the row as 128 u8 (one 128-byte
//                              row = one 128B-swizzled K block), max_j 1/||b_j|| per problem and a device
//                              flag "some value is not an integer in 0..255" (match_prep_kernel is the
//                              dim < 128 / gathered / column-major form of the same thing).
//      match_prep_split_kernel only when that flag is set: bf16 hi/lo split operands
//                              A' = [hi | hi | lo], B' = [hi | lo | hi] for the general-float path.
//   2. match_topk_u8_kernel    THE hot kernel.  tcgen05.mma.kind::i8 (u8 x u8 -> s32: the exact integer
//                              dot product), one CTA per SM, 18 warps: warp 0 = TMA producer (two resident
//                              128-row A panels, B tiles through a 7-stage mbarrier ring), warp 1 = MMA
//                              issuer (elect.sync, operands in uniform registers) + TMEM allocator
//                              (2 accumulator stages x 2 panels x 128 columns), warps 2-17 = epilogue:
//                              tcgen05.ld 32x32b.x32, integer max-tree prefilter against the row's bound,
//                              exact key = dot * 1/||b_j|| only for chunks that can matter, per-row top-3
//                              (key, column) in registers across all column tiles.  C is never stored.
//                              Persistent form (one CTA per SM over a contiguous share of panel x tile
//                              space) for single large problems; variants behind environment switches:
//                              match_topk_u8x2_kernel (cta_group::2 pairs), match_topk_u8ts_kernel (A in TMEM).
//      match_topk_kernel       the same structure with kind::f16 on the split-bf16 operands; both kernels
//                              are launched and one returns at once on the device flag (no host sync).
//                              With the caller's score bound (matchFeatures) it contracts the hi x hi term
//                              alone (K = 128 instead of 384, six B stages instead of four) and the certificate
//                              carries the one-term margin; the exact top-2 mode contracts all three terms.
//   3. match_finalize_kernel   merges the per-segment candidates, recomputes the oracle's exact FP32 score
//                              of the candidates and certifies the row from the monotone key -> score map
//                              (nearest neighbour with lowest index on ties, s2 or a proof that the
//                              threshold / ratio outcome cannot change).
//   4. match_rowscan_kernel    rows that could not be certified (3-way ties, near-ties on the split path)
//                              are re-evaluated by an exact scan of all columns (integer dp4a / FP32).
//   5. match_select_* / match_records_kernel   threshold + ratio (+ Unique) tests, ordered compaction, or
//                              the 16-byte best-2 records of the relocalisation shard.
//
// Score arithmetic is the contract in oracle/match.c (DESIGN.md "match arithmetic").
#include "vo_internal.h"
#include "vo_match.h"
#include "vo_ptx.cuh"
#include <cuda_bf16.h>
#include <cfloat>
#include <algorithm>

namespace vo {

constexpr int BM = 128;                       // rows per CTA == UMMA M
constexpr int BN = 256;                       // columns per tile == UMMA N
constexpr int BK = 64;                        // bf16 per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_KBLOCK_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;    // 32 KB
constexpr int MAX_KBLOCKS = 6;                // K' <= 384  (dim <= 128 split, or dim <= 384 exact)
constexpr int B_STAGES = 4;                   // B ring stages beside the 6 A K blocks of the three-term form ...
constexpr int B_STAGES_MAX = 6;               // ... and what the same shared memory holds when A is shorter (one term: 2 K blocks)
constexpr int NUM_EPI_WARPS = 16;              // 4 TMEM lane quarters x 4 column quarters of 64
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;
constexpr int NCAND = 3;
constexpr int SMEM_BYTES = MAX_KBLOCKS * A_KBLOCK_BYTES + B_STAGES * B_STAGE_BYTES + 256 + 2 * BN * 4 + BM * 4;   // 232,192 <= 232,448
constexpr float SPLIT_EPS = 1.0f / 8192.0f;   // three-term split: |approx key - oracle key| <= 2^-13 * ||a||
// One-term form (hi x hi only, used when the caller's score bound decides the rows): bf16 keeps 8 significant bits, so
// |x - hi(x)| <= 2^-8 |x| and |sum a_k b_k - sum hi(a_k) hi(b_k)| <= (2^-7 + 2^-16) sum |a_k b_k| <= 2^-7 (1 + 2^-9) ||a|| ||b||
// (Cauchy-Schwarz); FP32 accumulation and the sequential oracle dot add < 2^-16 each.  1/128 + 1/2048 leaves room to spare.
constexpr float ONE_TERM_EPS = 1.0f / 128.0f + 1.0f / 2048.0f;
// General-float path with a score bound: a column can only matter if its oracle key exceeds key_floor / inva, i.e. if
// its approximate key exceeds that minus eps * ||a||.  The same expression seeds the epilogue's bound and the
// certificate in match_finalize_kernel (1.0625: the subtraction itself is rounded).
__device__ __forceinline__ float general_key_floor(float key_floor, float ia, float eps) {
  const float na = __fdiv_rn(1.0f, ia);
  return __fdiv_rn(key_floor, ia) - 1.0625f * eps * na;
}

// generic mbarrier / TMA wrappers live in vo_ptx.cuh
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile, 128-byte swizzle: 8-row groups are 1024 B apart (SBO), LBO unused (=1).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// tcgen05.ld is asynchronous: its destination registers are only valid after tcgen05.wait::ld.  The
// registers are passed through the wait as read-write operands so the compiler cannot copy or
// repack them (e.g. to form f32x2 pairs) before the data has landed.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}

__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

struct Top3 {
  float k1, k2, k3;
  uint32_t i1, i2, i3;
  __device__ __forceinline__ void init() {
    k1 = k2 = k3 = -INFINITY;
    i1 = i2 = i3 = 0xFFFFFFFFu;
  }
  // strict '>' keeps the lowest column among equal keys when columns arrive in ascending order
  __device__ __forceinline__ void insert(float v, uint32_t j) {
    if (v > k1) { k3 = k2; i3 = i2; k2 = k1; i2 = i1; k1 = v; i1 = j; }
    else if (v > k2) { k3 = k2; i3 = i2; k2 = v; i2 = j; }
    else { k3 = v; i3 = j; }
  }
};

// Handles the (rare) "this chunk holds a new top-3 candidate" events of one thread.  g[q] is the max of
// v[8q..8q+7], m the chunk max.  Each iteration takes the current chunk maximum (lowest column on
// ties), inserts it, masks it out and re-reduces only its 8-group; ends when nothing beats thr.
template <int Q>
__device__ __forceinline__ void top3_take_from_group(float (&v)[32], float (&g)[4], float m, Top3& top, uint32_t j0) {
  int e = 7;
#pragma unroll
  for (int c = 6; c >= 0; --c) e = (v[8 * Q + c] == m) ? c : e;
  top.insert(m, j0 + 8 * Q + e);
#pragma unroll
  for (int c = 0; c < 8; ++c) v[8 * Q + c] = (c == e) ? -INFINITY : v[8 * Q + c];
  const float m01 = fmaxf(fmaxf(v[8 * Q], v[8 * Q + 1]), v[8 * Q + 2]);
  const float m23 = fmaxf(fmaxf(v[8 * Q + 3], v[8 * Q + 4]), v[8 * Q + 5]);
  g[Q] = fmaxf(fmaxf(m01, m23), fmaxf(v[8 * Q + 6], v[8 * Q + 7]));
}
__device__ __forceinline__ void top3_events(float (&v)[32], float (&g)[4], float m, float& thr, Top3& top, uint32_t j0) {
  while (m > thr) {
    if (g[0] == m) top3_take_from_group<0>(v, g, m, top, j0);
    else if (g[1] == m) top3_take_from_group<1>(v, g, m, top, j0);
    else if (g[2] == m) top3_take_from_group<2>(v, g, m, top, j0);
    else top3_take_from_group<3>(v, g, m, top, j0);
    thr = fmaxf(thr, top.k3);
    m = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
  }
}

// ------------------------------------------------------------------------- the GEMM kernel
// grid (row panels, column splits, problems).  Counts are read from device memory, so the launch
// shape depends only on capacities; CTAs outside the live problem exit (or publish empty slots).
__global__ void __launch_bounds__(NUM_THREADS, 1)
match_topk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const float* __restrict__ invb_base, int invb_stride, const int* __restrict__ n1p,
                  int n1_stride, int cap1, const int* __restrict__ n2p, int n2_stride, int cap2,
                  const int* __restrict__ nonint_flag, int kp_blocks, int n_splits,
                  uint2* __restrict__ cand_base, size_t cand_stride, float* __restrict__ dbg_c, int dbg_ld,
                  const float* __restrict__ inva_base, int inva_stride, float key_floor, int terms, float eps,
                  const int* __restrict__ invb_max_bits) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // 128B-swizzled operand tiles need 1024-byte alignment
  if (*nonint_flag == 0) return;   // exact-integer inputs are handled by match_topk_u8_kernel
  const int prob = blockIdx.z;
  const int n1 = min(n1p[prob * n1_stride], cap1), n2 = min(n2p[prob * n2_stride], cap2);
  const int m0 = blockIdx.x * BM;
  if (m0 >= n1) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_total = (n2 + BN - 1) / BN;
  const int split = blockIdx.y;
  const int t_begin = (int)((long long)split * tiles_total / n_splits);
  const int t_end = (int)((long long)(split + 1) * tiles_total / n_splits);
  const int n_slots = n_splits * 4;
  uint2* cand = cand_base + (size_t)prob * cand_stride;
  const float* invb = invb_base + (size_t)prob * invb_stride;

  if (t_begin >= t_end) {  // nothing to contract: publish empty candidate slots
    if (warp >= 2) {
      const int e = warp - 2, quarter = warp & 3, cq = e >> 2;
      const int row = m0 + quarter * 32 + lane;
      if (row < n1) {
        uint2* out = cand + ((size_t)row * n_slots + split * 4 + cq) * NCAND;
        for (int c = 0; c < NCAND; ++c) out[c] = make_uint2(__float_as_uint(-INFINITY), 0xFFFFFFFFu);
      }
    }
    return;
  }
  // operands are A' = [hi | hi | lo], B' = [hi | lo | hi]: the first kp_blocks K blocks are the hi x hi term alone
  const int kblocks = terms * kp_blocks;

  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  const uint32_t sA = base;
  // the B ring starts right after the A K blocks this launch uses: 4 stages of 32 KB with three terms, 6 with one
  const uint32_t sB = sA + kblocks * A_KBLOCK_BYTES;
  const int n_stages = min(B_STAGES_MAX, (MAX_KBLOCKS * A_KBLOCK_BYTES + B_STAGES * B_STAGE_BYTES - kblocks * A_KBLOCK_BYTES) / B_STAGE_BYTES);
  const uint32_t bars = sA + MAX_KBLOCKS * A_KBLOCK_BYTES + B_STAGES * B_STAGE_BYTES;
  const uint32_t bar_a_full = bars;
  const uint32_t bar_b_full = bars + 8;                        // [B_STAGES_MAX]
  const uint32_t bar_b_empty = bar_b_full + 8 * B_STAGES_MAX;  // [B_STAGES_MAX]
  const uint32_t bar_t_full = bar_b_empty + 8 * B_STAGES_MAX;  // [2]
  const uint32_t bar_t_empty = bar_t_full + 16;            // [2]
  const uint32_t bar_i_full = bar_t_empty + 16;            // [2]  1/||b|| slice of the tile has landed
  const uint32_t tmem_slot = bar_i_full + 16;
  const uint32_t s_invb_addr = bars + 256;                 // [2][BN] floats, filled by bulk TMA copies
  float* s_invb = reinterpret_cast<float*>(smem_raw + (s_invb_addr - smem_u32(smem_raw)));
  float* s_thr = s_invb + 2 * BN;                          // [BM] per-row lower bound of the third-best key
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    mbar_init(bar_a_full, 1);
    for (int s = 0; s < B_STAGES_MAX; ++s) { mbar_init(bar_b_full + 8 * s, 1); mbar_init(bar_b_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_t_full + 8 * s, 1); mbar_init(bar_t_empty + 8 * s, NUM_EPI_WARPS); mbar_init(bar_i_full + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + BM) {
    float t0 = -INFINITY;
    const int r = m0 + (int)threadIdx.x - 64;
    if (key_floor > 0.f && r < n1) {   // score bound (matchFeatures mode): keys at or below it cannot change the row's outcome
      const float ia = inva_base[(size_t)prob * inva_stride + r];
      if (ia > 0.f) t0 = general_key_floor(key_floor, ia, eps);
    }
    s_thr[threadIdx.x - 64] = t0;
  }
  if (warp == 1) {  // whole warp: allocate all 512 TMEM columns (2 accumulator stages x 256)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      mbar_expect_tx(bar_a_full, kblocks * A_KBLOCK_BYTES);
      for (int kb = 0; kb < kblocks; ++kb) tma_load_3d(sA + kb * A_KBLOCK_BYTES, &tmA, bar_a_full, kb * BK, m0, prob);
      int stage = 0; uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t)
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(bar_b_empty + 8 * stage, phase ^ 1);
          mbar_expect_tx(bar_b_full + 8 * stage, B_STAGE_BYTES);
          tma_load_3d(sB + stage * B_STAGE_BYTES, &tmB, bar_b_full + 8 * stage, kb * BK, t * BN, prob);
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1) {
    {  // ===== MMA issuer (warp-converged; one elected lane issues, operands stay warp-uniform) =====
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=256, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      mbar_wait(bar_a_full, 0);
      tc_fence_after();
      const uint64_t ad0 = umma_desc_sw128(sA), bd0 = umma_desc_sw128(sB);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(bar_t_empty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        uint32_t elected;
        asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\tselp.u32 %0, 1, 0, pe;\n\t}" : "=r"(elected));
        if (elected) {   // the epilogue of tile t-2 has drained this stage: its 1/||b|| slot can be refilled
          mbar_expect_tx(bar_i_full + 8 * acc, BN * 4);
          bulk_load_1d(s_invb_addr + acc * BN * 4, invb + (size_t)t * BN, BN * 4, bar_i_full + 8 * acc);
        }
        __syncwarp();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(bar_b_full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t ads = ad0 + (uint64_t)(kb * (A_KBLOCK_BYTES >> 4)), bds = bd0 + (uint64_t)(stage * (B_STAGE_BYTES >> 4));
          asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\tselp.u32 %0, 1, 0, pe;\n\t}" : "=r"(elected));
          if (elected) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              tc_mma_bf16(d_tmem, ads + (uint64_t)(2 * k), bds + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            tc_commit(bar_b_empty + 8 * stage);
            if (kb + 1 == kblocks) tc_commit(bar_t_full + 8 * acc);
          }
          __syncwarp();
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===== epilogue: thread = (row, 64-column quarter); running top-3 over all tiles of the split.
    // The four threads that share a row (one per column quarter) exchange a lower bound of the row's
    // third-best key through s_thr: an element at or below that bound can never enter the merged
    // top-3, so it is skipped (every skipped element still satisfies key <= final k3, which is all
    // the certification in match_finalize_kernel relies on).  The TMEM load of chunk c+1 is in
    // flight while chunk c is reduced; warps are decoupled from each other (no CTA barrier).
    const int e = warp - 2, quarter = warp & 3, cq = e >> 2;
    const int row_in_cta = quarter * 32 + lane;
    const int row = m0 + row_in_cta;
    Top3 top; top.init();
    int acc = 0; uint32_t acc_phase = 0;
    // Prefilter on the raw accumulators (the float counterpart of the u8 kernel's integer test): with a positive bound,
    //   dot <= thr_raw = RD(thr / max_j 1/||b_j||) * (1 - 2^-20)   implies   fl(dot * 1/||b_j||) <= thr   for every column,
    // so a chunk whose largest accumulator stays at or below thr_raw needs no key at all: one max tree and one compare.
    const float ibm = __int_as_float(invb_max_bits[prob]);
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(bar_i_full + 8 * acc, acc_phase);
      mbar_wait(bar_t_full + 8 * acc, acc_phase);
      tc_fence_after();
      float thr = fmaxf(top.k3, s_thr[row_in_cta]);
      const float thr_raw = (thr > 0.f && ibm > 0.f) ? __fdiv_rd(thr, ibm) * 0.99999905f : -INFINITY;
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + cq * 64);
      uint32_t r0[32], r1[32];
      tmem_ld32(tbase, r0);
#pragma unroll
      for (int chunk = 0; chunk < 2; ++chunk) {
        const int col_in_tile = cq * 64 + chunk * 32;
        const int j0 = t * BN + col_in_tile;
        uint32_t (&r)[32] = chunk == 0 ? r0 : r1;
        tmem_ld_wait(r);
        if (chunk == 0) tmem_ld32(tbase + 32, r1);
        if (j0 < n2) {
          if (dbg_c != nullptr && row < n1) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (j0 + c < n2) dbg_c[(size_t)row * dbg_ld + j0 + c] = __uint_as_float(r[c]);
          }
          float rm[8];
#pragma unroll
          for (int q = 0; q < 8; ++q)
            rm[q] = fmaxf(fmaxf(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1])), fmaxf(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
          const float rmax = fmaxf(fmaxf(fmaxf(rm[0], rm[1]), fmaxf(rm[2], rm[3])), fmaxf(fmaxf(rm[4], rm[5]), fmaxf(rm[6], rm[7])));
          if (!(rmax > thr_raw)) continue;            // nothing in this chunk can beat the row's bound
          float v[32];
          const float4* ibp = reinterpret_cast<const float4*>(s_invb + acc * BN + col_in_tile);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 i4 = ibp[q];
            const unsigned long long a0 = ((unsigned long long)r[4 * q + 1] << 32) | r[4 * q];
            const unsigned long long a1 = ((unsigned long long)r[4 * q + 3] << 32) | r[4 * q + 2];
            const unsigned long long b0 = ((unsigned long long)__float_as_uint(i4.y) << 32) | __float_as_uint(i4.x);
            const unsigned long long b1 = ((unsigned long long)__float_as_uint(i4.w) << 32) | __float_as_uint(i4.z);
            const unsigned long long p0 = f2_mul(a0, b0), p1 = f2_mul(a1, b1);
            v[4 * q] = __uint_as_float((uint32_t)p0); v[4 * q + 1] = __uint_as_float((uint32_t)(p0 >> 32));
            v[4 * q + 2] = __uint_as_float((uint32_t)p1); v[4 * q + 3] = __uint_as_float((uint32_t)(p1 >> 32));
          }
          if (j0 + 32 <= n2) {
            float g[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float m01 = fmaxf(fmaxf(v[8 * q], v[8 * q + 1]), v[8 * q + 2]);
              const float m23 = fmaxf(fmaxf(v[8 * q + 3], v[8 * q + 4]), v[8 * q + 5]);
              g[q] = fmaxf(fmaxf(m01, m23), fmaxf(v[8 * q + 6], v[8 * q + 7]));
            }
            const float m = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
            if (m > thr) top3_events(v, g, m, thr, top, (uint32_t)j0);
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (j0 + c < n2 && v[c] > thr) { top.insert(v[c], (uint32_t)(j0 + c)); thr = fmaxf(thr, top.k3); }
          }
        }
      }
      tc_fence_before();
      if (top.k3 > s_thr[row_in_cta]) s_thr[row_in_cta] = top.k3;   // racy max: any stored value is a valid bound
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_t_empty + 8 * acc);
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
    if (row < n1) {
      uint2* out = cand + ((size_t)row * n_slots + split * 4 + cq) * NCAND;
      out[0] = make_uint2(__float_as_uint(top.k1), top.i1);
      out[1] = make_uint2(__float_as_uint(top.k2), top.i2);
      out[2] = make_uint2(__float_as_uint(top.k3), top.i3);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------ the GEMM kernel, exact-integer path
// SIFT descriptors are integers 0..255 and K = 128, so the operands are stored as u8 and contracted by
// tcgen05.mma.kind::i8 (u8 x u8 -> s32, K = 32 per instruction): half the shared-memory operand bytes
// per MMA of the bf16 form and twice the MACs per issue slot, and the s32 accumulator IS the oracle's
// dot product (sum < 2^23).  The CTA keeps TWO 128-row A panels resident (M = 256) and streams
// 128-column B tiles (16 KB, one 128B-swizzled K block) through a 7-stage ring.
// TMEM: 2 accumulator stages x (2 panels x 128 columns).
//
// Epilogue thread = (panel, TMEM lane quarter, 64-column half); candidate slot = split*2 + half.
// The epilogue never forms key = dot * 1/||b_j|| on the fast path.  A row only needs columns whose
// key can beat its current bound thr (third-best key so far, or the caller's score bound, below), and
//     dot <= thr_raw = floor(thr / max_j 1/||b_j||) - 2   implies   fl(dot * 1/||b_j||) <= thr,
// so a chunk of 32 accumulators costs one integer max tree (16 VIMNMX3) and one compare.  Only chunks
// that pass go through the exact path (convert, scale by 1/||b_j||, strict '>' insert into the top-3).
//
// Score bound ("thresholded" mode, used by matchFeatures): a row is only kept when s1 <= T and
// s1/s2 <= R, so columns with score above T/R can never change the row's outcome.  thr starts at the
// key of that score (key_floor / inva_i) instead of -inf; match_finalize_kernel proves, per row, that
// the decision and (j1, s1) are exactly the oracle's, or sends the row to the exact scan.
constexpr int UBM = 256, UBN = 128;
constexpr int U_TILE_BYTES = 128 * 128;         // 128 rows x 128 u8
constexpr int U_STAGES = 7;
constexpr int U_SCRATCH_INTS = 32 * 33;         // per epilogue warp: one 32 x 32 chunk, padded (transpose)
constexpr int U_SMEM_BYTES = 2 * U_TILE_BYTES + U_STAGES * U_TILE_BYTES + 256 + UBM * 4 + NUM_EPI_WARPS * U_SCRATCH_INTS * 4;

__device__ __forceinline__ int imax3(int a, int b, int c) { return max(max(a, b), c); }
__device__ __forceinline__ int raw_bound(float thr, float bnorm) {
  // largest integer dot that provably cannot reach a key above thr (bnorm = 1 / max_j invb_j, 0 if none)
  if (!(thr > 0.f) || !(bnorm > 0.f)) return -1;
  const float q = fminf(__fmul_rn(thr, bnorm), 1.0e9f);
  return __float2int_rd(q) - 2;
}

// One work segment of a CTA: rows [m0, m0 + 256), column tiles [t0, t1), candidate slot pair `slot`.
struct USeg { int m0, t0, t1, slot; };

// Two launch forms share this kernel.  Grid form (batched problems with device-side counts): grid =
// (row panels, column splits, problems), one segment per CTA.  Persistent form (sched_L > 0: one problem
// whose sizes the host knows): grid = number of SMs; the panel x tile space is cut into equal contiguous
// ranges of sched_L tiles in panel-major order, so every SM gets the same work (128 row panels would
// occupy 128 of 148 SMs) and a CTA re-loads its A panels only when its range crosses a panel boundary.
// Segment k of a panel writes candidate slots 2k, 2k+1; match_finalize_kernel merges them.
__global__ void __launch_bounds__(NUM_THREADS, 1)
match_topk_u8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const float* __restrict__ invb_base, int invb_stride, const float* __restrict__ inva_base,
                     int inva_stride, const int* __restrict__ invb_max_bits, const int* __restrict__ n1p,
                     int n1_stride, int cap1, const int* __restrict__ n2p, int n2_stride, int cap2,
                     const int* __restrict__ nonint_flag, int n_splits, uint2* __restrict__ cand_base,
                     size_t cand_stride, int slots_per_row, float key_floor, float* __restrict__ dbg_c, int dbg_ld,
                     int sched_L, int sched_T, long long sched_total) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if (*nonint_flag != 0) return;   // general floats: match_topk_kernel (split-bf16) handles them
  const bool persistent = sched_L > 0;
  const int prob = persistent ? 0 : blockIdx.z;
  const int n1 = min(n1p[prob * n1_stride], cap1), n2 = min(n2p[prob * n2_stride], cap2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint2* cand = cand_base + (size_t)prob * cand_stride;
  const float* invb = invb_base + (size_t)prob * invb_stride;
  // ---- this CTA's segments
  const long long g_begin = persistent ? (long long)blockIdx.x * sched_L : 0;
  const long long g_end = persistent ? min(sched_total, g_begin + sched_L) : 1;
  USeg grid_seg = {0, 0, 0, 0};
  if (persistent) {
    if (g_begin >= g_end) return;
  } else {
    const int tiles_total = (n2 + UBN - 1) / UBN;
    grid_seg.m0 = blockIdx.x * UBM;
    if (grid_seg.m0 >= n1) return;
    grid_seg.t0 = (int)((long long)blockIdx.y * tiles_total / n_splits);
    grid_seg.t1 = (int)((long long)(blockIdx.y + 1) * tiles_total / n_splits);
    grid_seg.slot = blockIdx.y * 2;
    if (grid_seg.t0 >= grid_seg.t1) {   // nothing to contract: publish empty candidate slots
      if (warp >= 2) {
        const int e = warp - 2, quarter = warp & 3, panel = e >> 3, half = (e >> 2) & 1;
        const int row = grid_seg.m0 + panel * 128 + quarter * 32 + lane;
        if (row < n1) {
          uint2* out = cand + ((size_t)row * slots_per_row + grid_seg.slot + half) * NCAND;
          for (int c = 0; c < NCAND; ++c) out[c] = make_uint2(__float_as_uint(-INFINITY), 0xFFFFFFFFu);
        }
      }
      return;
    }
  }
  // cursor g -> next segment (every role walks the same list with its own cursor)
  auto next_seg = [&](long long& g, USeg& sg) -> bool {
    if (g >= g_end) return false;
    if (!persistent) { sg = grid_seg; g = g_end; return true; }
    const int panel = (int)(g / sched_T), t0 = (int)(g - (long long)panel * sched_T);
    const int len = (int)min((long long)(sched_T - t0), g_end - g);
    sg.m0 = panel * UBM; sg.t0 = t0; sg.t1 = t0 + len;
    sg.slot = 2 * (int)(blockIdx.x - ((long long)panel * sched_T) / sched_L);
    g += len;
    return true;
  };

  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  const uint32_t sA = base;                              // [2 panels] x 16 KB
  const uint32_t sB = sA + 2 * U_TILE_BYTES;             // [U_STAGES] x 16 KB
  const uint32_t bars = sB + U_STAGES * U_TILE_BYTES;
  const uint32_t bar_a_full = bars;
  const uint32_t bar_b_full = bars + 8;
  const uint32_t bar_b_empty = bar_b_full + 8 * U_STAGES;
  const uint32_t bar_t_full = bar_b_empty + 8 * U_STAGES;
  const uint32_t bar_t_empty = bar_t_full + 16;
  const uint32_t bar_a_empty = bar_t_empty + 16;
  const uint32_t tmem_slot = bar_a_empty + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - base));
  float* s_thr = reinterpret_cast<float*>(smem_raw + (bars + 256 - base));   // [UBM]

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    mbar_init(bar_a_full, 1); mbar_init(bar_a_empty, 1);
    for (int s = 0; s < U_STAGES; ++s) { mbar_init(bar_b_full + 8 * s, 1); mbar_init(bar_b_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_t_full + 8 * s, 1); mbar_init(bar_t_empty + 8 * s, NUM_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      long long g = g_begin; USeg sg; int seg_i = 0;
      int stage = 0; uint32_t phase = 0;
      while (next_seg(g, sg)) {
        mbar_wait(bar_a_empty, (uint32_t)((seg_i & 1) ^ 1));   // the previous segment's MMAs are done with the A panels
        mbar_expect_tx(bar_a_full, 2 * U_TILE_BYTES);
        for (int pn = 0; pn < 2; ++pn) tma_load_3d(sA + pn * U_TILE_BYTES, &tmA, bar_a_full, 0, sg.m0 + pn * 128, prob);
        for (int t = sg.t0; t < sg.t1; ++t) {
          mbar_wait(bar_b_empty + 8 * stage, phase ^ 1);
          mbar_expect_tx(bar_b_full + 8 * stage, U_TILE_BYTES);
          tma_load_3d(sB + stage * U_TILE_BYTES, &tmB, bar_b_full + 8 * stage, 0, t * UBN, prob);
          if (++stage == U_STAGES) { stage = 0; phase ^= 1; }
        }
        ++seg_i;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: D=s32, A=B=u8, K-major, N=128, M=128 (two panels), K=32 per MMA =====
    // The whole warp runs the loop converged and ONE elected lane issues: every value below is
    // warp-uniform, so the descriptors live in uniform registers (inside an `if (lane == 0)` region the
    // compiler rebuilt and re-broadcast them for every MMA, and the issuing thread's own instruction
    // stream -- about 100 cycles per MMA -- paced the tile loop, not the tensor pipe).
    {
      const uint32_t idesc = (2u << 4) | ((uint32_t)(UBN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t ad0 = umma_desc_sw128(sA), ad1 = umma_desc_sw128(sA + U_TILE_BYTES), bd0 = umma_desc_sw128(sB);
      long long g = g_begin; USeg sg; int seg_i = 0;
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      while (next_seg(g, sg)) {
        mbar_wait(bar_a_full, (uint32_t)(seg_i & 1));
        tc_fence_after();
        uint32_t elected;
        for (int t = sg.t0; t < sg.t1; ++t) {
          mbar_wait(bar_t_empty + 8 * acc, acc_phase ^ 1);
          mbar_wait(bar_b_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t d0 = tmem_base + acc * 256, d1 = d0 + UBN;
          const uint64_t bds = bd0 + (uint64_t)(stage * (U_TILE_BYTES >> 4));
          asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\tselp.u32 %0, 1, 0, pe;\n\t}" : "=r"(elected));
          if (elected) {
            asm volatile(
                "{\n\t.reg .pred p0, p1;\n\tsetp.ne.b32 p0, 0, 0;\n\tsetp.eq.b32 p1, 0, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %2, %4, %5, p0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%1], %3, %4, %5, p0;\n\t}"
                ::"r"(d0), "r"(d1), "l"(ad0), "l"(ad1), "l"(bds), "r"(idesc) : "memory");
#pragma unroll
            for (int k = 1; k < 4; ++k)
              asm volatile(
                  "{\n\t.reg .pred p1;\n\tsetp.eq.b32 p1, 0, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::i8 [%0], %2, %4, %5, p1;\n\t"
                  "tcgen05.mma.cta_group::1.kind::i8 [%1], %3, %4, %5, p1;\n\t}"
                  ::"r"(d0), "r"(d1), "l"(ad0 + (uint64_t)(2 * k)), "l"(ad1 + (uint64_t)(2 * k)), "l"(bds + (uint64_t)(2 * k)), "r"(idesc)
                  : "memory");
            tc_commit(bar_b_empty + 8 * stage);
            tc_commit(bar_t_full + 8 * acc);
            if (t + 1 == sg.t1) tc_commit(bar_a_empty);   // last tile of the segment: A panels may be replaced
          }
          __syncwarp();
          if (++stage == U_STAGES) { stage = 0; phase ^= 1; }
          acc ^= 1; if (acc == 0) acc_phase ^= 1;
        }
        ++seg_i;
      }
    }
  } else {
    // ===== epilogue =====
    const int e = warp - 2, quarter = warp & 3, panel = e >> 3, half = (e >> 2) & 1;
    const int row_in_cta = panel * 128 + quarter * 32 + lane;
    const float bmax = __int_as_float(invb_max_bits[prob]);
    const float bnorm = bmax > 0.f ? __fdiv_rn(1.0f, bmax) : 0.f;
    int* scr = reinterpret_cast<int*>(s_thr + UBM) + e * U_SCRATCH_INTS;
    long long g = g_begin; USeg sg;
    int acc = 0; uint32_t acc_phase = 0;
    while (next_seg(g, sg)) {
    const int row = sg.m0 + row_in_cta;
    // per-row key bound of this segment's panel: the caller's score bound in key units, or -inf.  The two
    // named barriers keep the 16 epilogue warps from re-initialising s_thr while a slower warp still uses it.
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (half == 0) {
      float t0 = -INFINITY;
      if (key_floor > 0.f && row < n1) {
        const float ia = inva_base[(size_t)prob * inva_stride + row];
        if (ia > 0.f) t0 = __fdiv_rn(key_floor, ia);
      }
      s_thr[row_in_cta] = t0;
    }
    asm volatile("bar.sync 1, 512;" ::: "memory");
    Top3 top; top.init();
    float thr = s_thr[row_in_cta];
    int thr_raw = raw_bound(thr, bnorm);
    if (row >= n1) { thr = INFINITY; thr_raw = 0x7fffffff; }   // padding rows never take the exact path
    for (int t = sg.t0; t < sg.t1; ++t) {
      mbar_wait(bar_t_full + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256 + panel * UBN + half * 64);
      uint32_t r0[32], r1[32];
      tmem_ld32(tbase, r0);
      tmem_ld32(tbase + 32, r1);
      const float shared_thr = s_thr[row_in_cta];
      if (shared_thr > thr) { thr = shared_thr; thr_raw = raw_bound(thr, bnorm); }
      tmem_ld_wait(r0);
      tmem_ld_wait(r1);
      // the accumulators are in registers: hand the TMEM stage back to the MMA warp at once
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_t_empty + 8 * acc);
#pragma unroll
      for (int chunk = 0; chunk < 2; ++chunk) {
        uint32_t (&r)[32] = chunk == 0 ? r0 : r1;
        const int j0 = t * UBN + half * 64 + chunk * 32;
        if (dbg_c != nullptr && row < n1) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (j0 + c < n2) dbg_c[(size_t)row * dbg_ld + j0 + c] = (float)(int)r[c];
        }
        int l1[12];
#pragma unroll
        for (int q = 0; q < 10; ++q) l1[q] = imax3((int)r[3 * q], (int)r[3 * q + 1], (int)r[3 * q + 2]);
        l1[10] = (int)r[30]; l1[11] = (int)r[31];
        const int a0 = imax3(l1[0], l1[1], l1[2]), a1 = imax3(l1[3], l1[4], l1[5]);
        const int a2 = imax3(l1[6], l1[7], l1[8]), a3 = imax3(l1[9], l1[10], l1[11]);
        const int m = max(imax3(a0, a1, a2), a3);
        const unsigned flagged = __ballot_sync(0xffffffffu, m > thr_raw);
        if (flagged) {
          // Exact path, warp-cooperative and warp-uniform.  The chunk is transposed through a padded
          // shared-memory scratch so that, for each flagged row L in turn, lane c holds column c:
          // one convert + multiply gives all 32 exact keys of the row, a ballot picks the columns above
          // the row's bound (ascending), and lane L alone updates its top-3.
#pragma unroll
          for (int c = 0; c < 32; ++c) scr[c * 33 + lane] = (int)r[c];
          const int jc = j0 + lane;
          const float ib = (jc < n2) ? invb[jc] : 0.f;
          __syncwarp();
          unsigned f = flagged;
          while (f) {
            const int L = __ffs(f) - 1;
            f &= f - 1;
            const float key = __fmul_rn((float)scr[lane * 33 + L], ib);
            float thr_l = __shfl_sync(0xffffffffu, thr, L);
            unsigned cm = __ballot_sync(0xffffffffu, jc < n2 && key > thr_l);
            while (cm) {
              const int c = __ffs(cm) - 1;
              const float kk = __shfl_sync(0xffffffffu, key, c);
              if (lane == L && kk > thr) {
                top.insert(kk, (uint32_t)(j0 + c));
                if (top.k3 > thr) { thr = top.k3; thr_raw = raw_bound(thr, bnorm); }
              }
              thr_l = __shfl_sync(0xffffffffu, thr, L);
              cm &= __ballot_sync(0xffffffffu, key > thr_l) & ~((2u << c) - 1u);
            }
          }
          __syncwarp();
        }
      }
      if (top.k3 > shared_thr) s_thr[row_in_cta] = top.k3;   // racy max: any stored value is a valid bound
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
    if (row < n1) {
      uint2* out = cand + ((size_t)row * slots_per_row + sg.slot + half) * NCAND;
      out[0] = make_uint2(__float_as_uint(top.k1), top.i1);
      out[1] = make_uint2(__float_as_uint(top.k2), top.i2);
      out[2] = make_uint2(__float_as_uint(top.k3), top.i3);
    }
    }   // segments
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// --------------------------------------------------- exact-integer path, A operand in TMEM
// tools/ubench_tmem.cu: a cta_group::1 MMA with both operands in shared memory is bound by the rate at
// which the tensor pipe reads them (about 75 B/cycle/SM).  The A panels of a CTA never change, so here
// they live in TMEM (written once with tcgen05.st: lane = row, 32 columns = the row's 128 u8) and the
// MMA takes A from TMEM ("TS" form); shared memory then only feeds B.  TMEM budget: 64 columns for the
// two A panels leave 448, i.e. two accumulator stages of 2 panels x 96 columns, so tiles are 96 columns
// wide and there are 24 epilogue warps (panel, TMEM lane quarter, 32-column third), one chunk each.
constexpr int TN = 96;                           // tile width (columns)
constexpr int T_TILE_BYTES = TN * 128;           // 12 KB of u8 B rows
constexpr int T_STAGES = 10;
constexpr int T_EPI_WARPS = 24;
constexpr int T_THREADS = (2 + T_EPI_WARPS) * 32;
constexpr int T_ACOL = 2 * 2 * TN;               // TMEM column of A panel 0 (after the accumulators)
constexpr int T_SMEM_BYTES = T_STAGES * T_TILE_BYTES + 256 + UBM * 4 + T_EPI_WARPS * U_SCRATCH_INTS * 4;

__global__ void __launch_bounds__(T_THREADS, 1)
match_topk_u8ts_kernel(const __grid_constant__ CUtensorMap tmB, const uint8_t* __restrict__ u8a, int a_alloc,
                       const float* __restrict__ invb_base, int invb_stride, const float* __restrict__ inva_base,
                       int inva_stride, const int* __restrict__ invb_max_bits, const int* __restrict__ n1p,
                       int n1_stride, int cap1, const int* __restrict__ n2p, int n2_stride, int cap2,
                       const int* __restrict__ nonint_flag, int n_splits, uint2* __restrict__ cand_base,
                       size_t cand_stride, int slots_per_row, float key_floor, float* __restrict__ dbg_c, int dbg_ld) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if (*nonint_flag != 0) return;
  const int prob = blockIdx.z;
  const int n1 = min(n1p[prob * n1_stride], cap1), n2 = min(n2p[prob * n2_stride], cap2);
  const int m0 = blockIdx.x * UBM;
  if (m0 >= n1) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_total = (n2 + TN - 1) / TN;
  const int split = blockIdx.y;
  const int t_begin = (int)((long long)split * tiles_total / n_splits);
  const int t_end = (int)((long long)(split + 1) * tiles_total / n_splits);
  uint2* cand = cand_base + (size_t)prob * cand_stride;
  const float* invb = invb_base + (size_t)prob * invb_stride;
  // epilogue warp -> (panel, lane quarter, column third); a warp may only touch TMEM lanes 32*(warp%4)..
  const int e = warp - 2, quarter = warp & 3, group = e >> 2, panel = group / 3, third = group - 3 * panel;

  if (t_begin >= t_end) {
    if (warp >= 2) {
      const int row = m0 + panel * 128 + quarter * 32 + lane;
      if (row < n1) {
        uint2* out = cand + ((size_t)row * slots_per_row + split * 4 + third) * NCAND;
        for (int c = 0; c < NCAND; ++c) out[c] = make_uint2(__float_as_uint(-INFINITY), 0xFFFFFFFFu);
        if (third == 0) for (int c = 0; c < NCAND; ++c) out[3 * NCAND + c] = make_uint2(__float_as_uint(-INFINITY), 0xFFFFFFFFu);
      }
    }
    return;
  }
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  const uint32_t sB = base;                                   // [T_STAGES] x 12 KB
  const uint32_t bars = sB + T_STAGES * T_TILE_BYTES;
  const uint32_t bar_a_full = bars;
  const uint32_t bar_b_full = bars + 8;
  const uint32_t bar_b_empty = bar_b_full + 8 * T_STAGES;
  const uint32_t bar_t_full = bar_b_empty + 8 * T_STAGES;
  const uint32_t bar_t_empty = bar_t_full + 16;
  const uint32_t tmem_slot = bar_t_empty + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - base));
  float* s_thr = reinterpret_cast<float*>(smem_raw + (bars + 256 - base));   // [UBM]

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    mbar_init(bar_a_full, 8);                                 // the 8 warps that write the A panels
    for (int s = 0; s < T_STAGES; ++s) { mbar_init(bar_b_full + 8 * s, 1); mbar_init(bar_b_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_t_full + 8 * s, 1); mbar_init(bar_t_empty + 8 * s, T_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + UBM) {
    const int r = threadIdx.x - 64;
    float t0 = -INFINITY;
    if (key_floor > 0.f && m0 + r < n1) {
      const float ia = inva_base[(size_t)prob * inva_stride + m0 + r];
      if (ia > 0.f) t0 = __fdiv_rn(key_floor, ia);
    }
    s_thr[r] = t0;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer: B tiles only =====
      int stage = 0; uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(bar_b_empty + 8 * stage, phase ^ 1);
        mbar_expect_tx(bar_b_full + 8 * stage, T_TILE_BYTES);
        tma_load_3d(sB + stage * T_TILE_BYTES, &tmB, bar_b_full + 8 * stage, 0, t * TN, prob);
        if (++stage == T_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {  // ===== MMA issuer (warp-converged, one elected lane): D=s32 [tmem], A=u8 [tmem], B=u8 smem, M=128, N=96, K=32 =====
      const uint32_t idesc = (2u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(bar_a_full, 0);
      tc_fence_after();
      const uint64_t bd0 = umma_desc_sw128(sB);
      const uint32_t a0 = tmem_base + T_ACOL, a1 = a0 + 32;
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(bar_t_empty + 8 * acc, acc_phase ^ 1);
        mbar_wait(bar_b_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * (2 * TN), d1 = d0 + TN;
        const uint64_t bds = bd0 + (uint64_t)(stage * (T_TILE_BYTES >> 4));
        uint32_t elected;
        asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\tselp.u32 %0, 1, 0, pe;\n\t}" : "=r"(elected));
        if (elected) {
          asm volatile(
              "{\n\t.reg .pred p0;\n\tsetp.ne.b32 p0, 0, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], [%2], %4, %5, p0;\n\t"
              "tcgen05.mma.cta_group::1.kind::i8 [%1], [%3], %4, %5, p0;\n\t}"
              ::"r"(d0), "r"(d1), "r"(a0), "r"(a1), "l"(bds), "r"(idesc) : "memory");
#pragma unroll
          for (int k = 1; k < 4; ++k)
            asm volatile(
                "{\n\t.reg .pred p1;\n\tsetp.eq.b32 p1, 0, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], [%2], %4, %5, p1;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%1], [%3], %4, %5, p1;\n\t}"
                ::"r"(d0), "r"(d1), "r"(a0 + 8 * k), "r"(a1 + 8 * k), "l"(bds + (uint64_t)(2 * k)), "r"(idesc) : "memory");
          tc_commit(bar_b_empty + 8 * stage);
          tc_commit(bar_t_full + 8 * acc);
        }
        __syncwarp();
        if (++stage == T_STAGES) { stage = 0; phase ^= 1; }
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===== epilogue =====
    const int row_in_cta = panel * 128 + quarter * 32 + lane;
    const int row = m0 + row_in_cta;
    if (third == 0) {   // these 8 warps first put the A panels into TMEM: lane = row, 32 columns = 128 u8
      uint32_t a[32];
      const uint4* src = reinterpret_cast<const uint4*>(u8a + ((size_t)prob * a_alloc + row) * 128);
#pragma unroll
      for (int q = 0; q < 8; ++q) { const uint4 v = __ldg(src + q); a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w; }
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(T_ACOL + panel * 32);
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
          "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
          "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
          ::"r"(taddr), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
            "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]),
            "r"(a[16]), "r"(a[17]), "r"(a[18]), "r"(a[19]), "r"(a[20]), "r"(a[21]), "r"(a[22]), "r"(a[23]),
            "r"(a[24]), "r"(a[25]), "r"(a[26]), "r"(a[27]), "r"(a[28]), "r"(a[29]), "r"(a[30]), "r"(a[31])
          : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a_full);
    }
    const float bmax = __int_as_float(invb_max_bits[prob]);
    const float bnorm = bmax > 0.f ? __fdiv_rn(1.0f, bmax) : 0.f;
    int* scr = reinterpret_cast<int*>(s_thr + UBM) + e * U_SCRATCH_INTS;
    Top3 top; top.init();
    float thr = s_thr[row_in_cta];
    int thr_raw = raw_bound(thr, bnorm);
    if (row >= n1) { thr = INFINITY; thr_raw = 0x7fffffff; }
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(bar_t_full + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * (2 * TN) + panel * TN + third * 32);
      uint32_t r[32];
      tmem_ld32(tbase, r);
      const float shared_thr = s_thr[row_in_cta];
      if (shared_thr > thr) { thr = shared_thr; thr_raw = raw_bound(thr, bnorm); }
      tmem_ld_wait(r);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_t_empty + 8 * acc);
      const int j0 = t * TN + third * 32;
      if (dbg_c != nullptr && row < n1) {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (j0 + c < n2) dbg_c[(size_t)row * dbg_ld + j0 + c] = (float)(int)r[c];
      }
      int l1[12];
#pragma unroll
      for (int q = 0; q < 10; ++q) l1[q] = imax3((int)r[3 * q], (int)r[3 * q + 1], (int)r[3 * q + 2]);
      l1[10] = (int)r[30]; l1[11] = (int)r[31];
      const int a0 = imax3(l1[0], l1[1], l1[2]), a1 = imax3(l1[3], l1[4], l1[5]);
      const int a2 = imax3(l1[6], l1[7], l1[8]), a3 = imax3(l1[9], l1[10], l1[11]);
      const int m = max(imax3(a0, a1, a2), a3);
      const unsigned flagged = __ballot_sync(0xffffffffu, m > thr_raw);
      if (flagged) {
#pragma unroll
        for (int c = 0; c < 32; ++c) scr[c * 33 + lane] = (int)r[c];
        const int jc = j0 + lane;
        const float ib = (jc < n2) ? invb[jc] : 0.f;
        __syncwarp();
        unsigned f = flagged;
        while (f) {
          const int L = __ffs(f) - 1;
          f &= f - 1;
          const float key = __fmul_rn((float)scr[lane * 33 + L], ib);
          float thr_l = __shfl_sync(0xffffffffu, thr, L);
          unsigned cm = __ballot_sync(0xffffffffu, jc < n2 && key > thr_l);
          while (cm) {
            const int c = __ffs(cm) - 1;
            const float kk = __shfl_sync(0xffffffffu, key, c);
            if (lane == L && kk > thr) {
              top.insert(kk, (uint32_t)(j0 + c));
              if (top.k3 > thr) { thr = top.k3; thr_raw = raw_bound(thr, bnorm); }
            }
            thr_l = __shfl_sync(0xffffffffu, thr, L);
            cm &= __ballot_sync(0xffffffffu, key > thr_l) & ~((2u << c) - 1u);
          }
        }
        __syncwarp();
      }
      if (top.k3 > shared_thr) s_thr[row_in_cta] = top.k3;
      acc ^= 1; if (acc == 0) acc_phase ^= 1;
    }
    if (row < n1) {
      uint2* out = cand + ((size_t)row * slots_per_row + split * 4 + third) * NCAND;
      out[0] = make_uint2(__float_as_uint(top.k1), top.i1);
      out[1] = make_uint2(__float_as_uint(top.k2), top.i2);
      out[2] = make_uint2(__float_as_uint(top.k3), top.i3);
      if (third == 0)   // the fourth slot of this split is not used by this kernel
        for (int c = 0; c < NCAND; ++c) out[3 * NCAND + c] = make_uint2(__float_as_uint(-INFINITY), 0xFFFFFFFFu);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------- exact-integer path on CTA PAIRS (cta_group::2)
// tools/ubench_tmem.cu shows that cta_group::1 MMAs are bound by the rate at which the tensor pipe
// reads its operands from shared memory (108 cycles per M128 x N128 x K32 MMA, 64 ideal).  In
// cta_group::2 form two SMs of a cluster share one M = 256, N = 256 MMA: each SM holds its own 128 A rows
// and HALF of the B tile (128 of the 256 columns), so per MMA an SM reads 8 KB for 128 ideal cycles.
// Layout per CTA: A panel 16 KB (its 128 rows, resident), B ring X_STAGES x 16 KB (its 128 B rows of each
// 256-column tile), TMEM 2 accumulator stages x 256 columns (its 128 rows x 256 columns per tile).
// Protocol (the CUTLASS 2-SM pattern): both CTAs issue their TMA loads with .cta_group::2 and signal
// the LEADER's (cluster rank 0) "full" barrier; the leader's MMA thread issues the MMAs and
// multicast-commits to the "stage empty" and "accumulator full" barriers of BOTH CTAs; the epilogue
// warps of both CTAs arrive (remotely for rank 1) on the leader's "accumulator empty" barrier.
// Epilogue thread = (TMEM lane quarter, 64-column quarter); candidate slot = split * 4 + quarter.
// XN = tile width in columns (256: two 32-column chunks per epilogue warp and 2 TMEM stages; 128: one chunk
// and 4 TMEM stages -- the deeper accumulator ring hides the MMA -> epilogue -> MMA signalling latency).
constexpr int X_RING_BYTES = 8 * U_TILE_BYTES;     // B ring: 128 KB per CTA whatever the tile width
constexpr int X_SMEM_BYTES = U_TILE_BYTES + X_RING_BYTES + 512 + 128 * 4 + NUM_EPI_WARPS * U_SCRATCH_INTS * 4;
// shared::cluster address of the leader's (cluster rank 0) copy of a local shared-memory object
__device__ __forceinline__ uint32_t leader_addr(uint32_t local) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(local)); return r;
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_addr(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm_mc(uint32_t bar) {   // arrive on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {   // arrive on the leader CTA's copy of `bar`
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(bar)) : "memory");
}

template <int XN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
match_topk_u8x2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const float* __restrict__ invb_base, int invb_stride, const float* __restrict__ inva_base,
                       int inva_stride, const int* __restrict__ invb_max_bits, const int* __restrict__ n1p,
                       int n1_stride, int cap1, const int* __restrict__ n2p, int n2_stride, int cap2,
                       const int* __restrict__ nonint_flag, int n_splits, uint2* __restrict__ cand_base,
                       size_t cand_stride, int slots_per_row, float key_floor, float* __restrict__ dbg_c, int dbg_ld,
                       long long* __restrict__ clk) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // every exit before the cluster barriers below is taken by BOTH CTAs of a pair (same pair index, split, problem)
  if (*nonint_flag != 0) return;
  const int prob = blockIdx.z;
  const int n1 = min(n1p[prob * n1_stride], cap1), n2 = min(n2p[prob * n2_stride], cap2);
  const uint32_t rank = cluster_ctarank();            // == blockIdx.x & 1 for cluster dims (2,1,1)
  const int m_pair = (blockIdx.x >> 1) * 256;
  if (m_pair >= n1) return;
  const int m0 = m_pair + (int)rank * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NS = 512 / XN;                       // TMEM accumulator stages
  constexpr int CH = XN / 128;                       // 32-column chunks per epilogue warp and tile
  constexpr int XB_BYTES = (XN / 2) * 128;           // this CTA's half of a B tile
  constexpr int X_STAGES = X_RING_BYTES / XB_BYTES;
  const int tiles_total = (n2 + XN - 1) / XN;
  const int split = blockIdx.y;
  const int t_begin = (int)((long long)split * tiles_total / n_splits);
  const int t_end = (int)((long long)(split + 1) * tiles_total / n_splits);
  uint2* cand = cand_base + (size_t)prob * cand_stride;
  const float* invb = invb_base + (size_t)prob * invb_stride;

  if (t_begin >= t_end) {
    if (warp >= 2) {
      const int e = warp - 2, quarter = warp & 3, cq = e >> 2;
      const int row = m0 + quarter * 32 + lane;
      if (row < n1) {
        uint2* out = cand + ((size_t)row * slots_per_row + split * 4 + cq) * NCAND;
        for (int c = 0; c < NCAND; ++c) out[c] = make_uint2(__float_as_uint(-INFINITY), 0xFFFFFFFFu);
      }
    }
    return;
  }
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  const uint32_t sA = base;                              // 16 KB
  const uint32_t sB = sA + U_TILE_BYTES;                 // [X_STAGES] x XB_BYTES
  const uint32_t bars = sB + X_RING_BYTES;
  const uint32_t bar_a_full = bars;                      // leader only
  const uint32_t bar_b_full = bars + 8;                  // leader only [X_STAGES]
  const uint32_t bar_b_empty = bar_b_full + 8 * X_STAGES;   // both CTAs [X_STAGES]
  const uint32_t bar_t_full = bar_b_empty + 8 * X_STAGES;   // both CTAs [NS]
  const uint32_t bar_t_empty = bar_t_full + 8 * NS;         // leader only [NS]
  const uint32_t tmem_slot = bar_t_empty + 8 * NS;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - base));
  float* s_thr = reinterpret_cast<float*>(smem_raw + (bars + 512 - base));   // [128]

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    mbar_init(bar_a_full, 1);
    for (int s = 0; s < X_STAGES; ++s) { mbar_init(bar_b_full + 8 * s, 1); mbar_init(bar_b_empty + 8 * s, 1); }
    for (int s = 0; s < NS; ++s) { mbar_init(bar_t_full + 8 * s, 1); mbar_init(bar_t_empty + 8 * s, 2 * NUM_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 128) {
    const int r = threadIdx.x - 64;
    float t0 = -INFINITY;
    if (key_floor > 0.f && m0 + r < n1) {
      const float ia = inva_base[(size_t)prob * inva_stride + m0 + r];
      if (ia > 0.f) t0 = __fdiv_rn(key_floor, ia);
    }
    s_thr[r] = t0;
  }
  if (warp == 1) {   // the same warp of both CTAs allocates all 512 columns for the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // barriers of both CTAs are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer (both CTAs; all transaction bytes land on the leader's barriers) =====
      if (rank == 0) mbar_expect_tx(bar_a_full, 2 * U_TILE_BYTES);
      tma_load_3d_2sm(sA, &tmA, bar_a_full, 0, m0, prob);
      int stage = 0; uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(bar_b_empty + 8 * stage, phase ^ 1);
        if (rank == 0) mbar_expect_tx(bar_b_full + 8 * stage, 2 * XB_BYTES);
        tma_load_3d_2sm(sB + stage * XB_BYTES, &tmB, bar_b_full + 8 * stage, 0, t * XN + (int)rank * (XN / 2), prob);
        if (++stage == X_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {  // ===== MMA issuer (leader CTA, warp-converged, one elected lane): M=256 (pair), N=XN, K=32 per MMA =====
      const uint32_t idesc = (2u << 4) | ((uint32_t)(XN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      mbar_wait(bar_a_full, 0);
      tc_fence_after();
      const uint64_t ad0 = umma_desc_sw128(sA), bd0 = umma_desc_sw128(sB);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      long long c_te = 0, c_bf = 0; const long long c_start = clock64();
      for (int t = t_begin; t < t_end; ++t) {
        if (clk) {   // debug: where does the issuing warp wait?
          const long long c0 = clock64();
          mbar_wait(bar_t_empty + 8 * acc, acc_phase ^ 1);
          const long long c1 = clock64();
          mbar_wait(bar_b_full + 8 * stage, phase);
          c_te += c1 - c0; c_bf += clock64() - c1;
        } else {
          mbar_wait(bar_t_empty + 8 * acc, acc_phase ^ 1);
          mbar_wait(bar_b_full + 8 * stage, phase);
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * XN;
        const uint64_t bds = bd0 + (uint64_t)(stage * (XB_BYTES >> 4));
        uint32_t elected;
        asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\tselp.u32 %0, 1, 0, pe;\n\t}" : "=r"(elected));
        if (elected) {
          asm volatile(
              "{\n\t.reg .pred p0, p1;\n\tsetp.ne.b32 p0, 0, 0;\n\tsetp.eq.b32 p1, 0, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %5, %9, p0;\n\t"
              "tcgen05.mma.cta_group::2.kind::i8 [%0], %2, %6, %9, p1;\n\t"
              "tcgen05.mma.cta_group::2.kind::i8 [%0], %3, %7, %9, p1;\n\t"
              "tcgen05.mma.cta_group::2.kind::i8 [%0], %4, %8, %9, p1;\n\t}"
              ::"r"(d_tmem), "l"(ad0), "l"(ad0 + 2), "l"(ad0 + 4), "l"(ad0 + 6), "l"(bds), "l"(bds + 2), "l"(bds + 4), "l"(bds + 6),
                "r"(idesc)
              : "memory");
          tc_commit_2sm_mc(bar_b_empty + 8 * stage);
          tc_commit_2sm_mc(bar_t_full + 8 * acc);
        }
        __syncwarp();
        if (++stage == X_STAGES) { stage = 0; phase ^= 1; }
        if (++acc == NS) { acc = 0; acc_phase ^= 1; }
      }
      if (clk && lane == 0) {
        long long* o = clk + 4 * ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
        o[0] = clock64() - c_start; o[1] = c_te; o[2] = c_bf; o[3] = t_end - t_begin;
      }
    }
  } else {
    // ===== epilogue (both CTAs): same fast / exact paths as match_topk_u8_kernel =====
    const int e = warp - 2, quarter = warp & 3, cq = e >> 2;
    const int row_in_cta = quarter * 32 + lane;
    const int row = m0 + row_in_cta;
    const float bmax = __int_as_float(invb_max_bits[prob]);
    const float bnorm = bmax > 0.f ? __fdiv_rn(1.0f, bmax) : 0.f;
    int* scr = reinterpret_cast<int*>(s_thr + 128) + e * U_SCRATCH_INTS;
    Top3 top; top.init();
    float thr = s_thr[row_in_cta];
    int thr_raw = raw_bound(thr, bnorm);
    if (row >= n1) { thr = INFINITY; thr_raw = 0x7fffffff; }
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(bar_t_full + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * XN + cq * (XN / 4));
      uint32_t r0[32], r1[32];
      tmem_ld32(tbase, r0);
      if (CH == 2) tmem_ld32(tbase + 32, r1);
      const float shared_thr = s_thr[row_in_cta];
      if (shared_thr > thr) { thr = shared_thr; thr_raw = raw_bound(thr, bnorm); }
      tmem_ld_wait(r0);
      if (CH == 2) tmem_ld_wait(r1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bar_t_empty + 8 * acc);
#pragma unroll
      for (int chunk = 0; chunk < CH; ++chunk) {
        uint32_t (&r)[32] = chunk == 0 ? r0 : r1;
        const int j0 = t * XN + cq * (XN / 4) + chunk * 32;
        if (dbg_c != nullptr && row < n1) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (j0 + c < n2) dbg_c[(size_t)row * dbg_ld + j0 + c] = (float)(int)r[c];
        }
        int l1[12];
#pragma unroll
        for (int q = 0; q < 10; ++q) l1[q] = imax3((int)r[3 * q], (int)r[3 * q + 1], (int)r[3 * q + 2]);
        l1[10] = (int)r[30]; l1[11] = (int)r[31];
        const int a0 = imax3(l1[0], l1[1], l1[2]), a1 = imax3(l1[3], l1[4], l1[5]);
        const int a2 = imax3(l1[6], l1[7], l1[8]), a3 = imax3(l1[9], l1[10], l1[11]);
        const int m = max(imax3(a0, a1, a2), a3);
        const unsigned flagged = __ballot_sync(0xffffffffu, m > thr_raw);
        if (flagged) {
#pragma unroll
          for (int c = 0; c < 32; ++c) scr[c * 33 + lane] = (int)r[c];
          const int jc = j0 + lane;
          const float ib = (jc < n2) ? invb[jc] : 0.f;
          __syncwarp();
          unsigned f = flagged;
          while (f) {
            const int L = __ffs(f) - 1;
            f &= f - 1;
            const float key = __fmul_rn((float)scr[lane * 33 + L], ib);
            float thr_l = __shfl_sync(0xffffffffu, thr, L);
            unsigned cm = __ballot_sync(0xffffffffu, jc < n2 && key > thr_l);
            while (cm) {
              const int c = __ffs(cm) - 1;
              const float kk = __shfl_sync(0xffffffffu, key, c);
              if (lane == L && kk > thr) {
                top.insert(kk, (uint32_t)(j0 + c));
                if (top.k3 > thr) { thr = top.k3; thr_raw = raw_bound(thr, bnorm); }
              }
              thr_l = __shfl_sync(0xffffffffu, thr, L);
              cm &= __ballot_sync(0xffffffffu, key > thr_l) & ~((2u << c) - 1u);
            }
          }
          __syncwarp();
        }
      }
      if (top.k3 > shared_thr) s_thr[row_in_cta] = top.k3;
      if (++acc == NS) { acc = 0; acc_phase ^= 1; }
    }
    if (row < n1) {
      uint2* out = cand + ((size_t)row * slots_per_row + split * 4 + cq) * NCAND;
      out[0] = make_uint2(__float_as_uint(top.k1), top.i1);
      out[1] = make_uint2(__float_as_uint(top.k2), top.i2);
      out[2] = make_uint2(__float_as_uint(top.k3), top.i3);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // no CTA of the pair exits (or frees TMEM) while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

__global__ void match_set_counts_kernel(int* counts, int n1, int n2) {
  if (threadIdx.x == 0) { counts[0] = n1; counts[1] = n2; }
}

// ------------------------------------------------------------------------------ prep kernel
// grid (cap_pad / 32, problems).  Writes out[p][row][3*kp] bf16 (A: hi|hi|lo, B: hi|lo|hi),
// inv[p][row] (0 for rows >= n), an optional compact row-major fp32 copy (needed when rows are
// gathered or column-major), and ORs the "not an integer in 0..255" flag.
constexpr int PREP_ROWS = 32;

// rows [row0, row0 + PREP_ROWS) of problem `prob` -> tile[r][dim + 1] (zero beyond the live count)
__device__ __forceinline__ void prep_load_tile(const MatchOperand& op, int prob, int row0, int n, int dim, float* tile) {
  const int ldt = dim + 1, tid = threadIdx.x;
  const float* f = op.base + (size_t)prob * op.prob_stride;
  const uint32_t* g = op.gather ? op.gather + (size_t)prob * op.gather_stride : nullptr;
  if (!op.col_major) {
    for (int idx = tid; idx < PREP_ROWS * dim; idx += 256) {
      const int r = idx / dim, k = idx - r * dim;
      float v = 0.f;
      if (row0 + r < n) {
        const size_t src_row = g ? (size_t)g[row0 + r] : (size_t)(row0 + r);
        v = f[src_row * dim + k];
      }
      tile[r * ldt + k] = v;
    }
  } else {
    for (int idx = tid; idx < PREP_ROWS * dim; idx += 256) {
      const int k = idx / PREP_ROWS, r = idx - k * PREP_ROWS;
      tile[r * ldt + k] = (row0 + r < n) ? f[(size_t)k * op.ld + row0 + r] : 0.f;
    }
  }
}

// Always runs: 1/||row|| in the oracle's arithmetic, the u8 operands of the exact-integer kernel, the
// "some value is not an integer in 0..255" flag and max_j 1/||b_j||.  Eight threads share a row.  When the
// row is all integers 0..255 every partial sum of squares is an integer below 2^24, hence exact in
// fp32 in ANY order and equal to the oracle's sequential fmaf fold; otherwise one thread redoes the
// fold sequentially.
__global__ void __launch_bounds__(256)
match_prep_kernel(const MatchOperand op, int rows_alloc, int dim, uint8_t* __restrict__ out_u8,
                  float* __restrict__ inv, int* __restrict__ nonint_flag, int* __restrict__ inv_max_bits) {
  extern __shared__ float tile[];  // [PREP_ROWS][dim + 1]
  const int prob = blockIdx.y;
  const int n = min(op.count[prob * op.count_stride], op.cap);
  const int row0 = blockIdx.x * PREP_ROWS;
  const int ldt = dim + 1;
  const int tid = threadIdx.x;
  float* inv_p = inv + (size_t)prob * rows_alloc;
  uint32_t* u8_p = reinterpret_cast<uint32_t*>(out_u8 + ((size_t)prob * rows_alloc + row0) * 128);
  if (row0 >= n) {   // dead rows still feed the MMA as zero operands
    for (int r = tid; r < PREP_ROWS; r += 256)
      if (row0 + r < rows_alloc) inv_p[row0 + r] = 0.f;
    for (int idx = tid; idx < PREP_ROWS * 32; idx += 256)
      if (row0 + idx / 32 < rows_alloc) u8_p[idx] = 0u;
    return;
  }
  prep_load_tile(op, prob, row0, n, dim, tile);
  __syncthreads();
  {
    const int r = tid >> 3, part = tid & 7;
    const float* x = tile + r * ldt;
    float acc = 0.f;
    bool isint = true;
    for (int k = part; k < dim; k += 8) {
      const float v = x[k];
      acc = fmaf(v, v, acc);
      isint = isint && (v == truncf(v)) && (v >= 0.f) && (v <= 255.f);
    }
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
      acc += __shfl_xor_sync(0xffffffffu, acc, off);
      isint = __shfl_xor_sync(0xffffffffu, (int)isint, off) && isint;
    }
    if (part == 0) {
      if (!isint) {
        acc = 0.f;
        for (int k = 0; k < dim; ++k) acc = fmaf(x[k], x[k], acc);
        atomicOr(nonint_flag, 1);
      }
      const float iv = (acc == 0.f || row0 + r >= n) ? 0.f : __fdiv_rn(1.0f, __fsqrt_rn(acc));
      if (row0 + r < rows_alloc) inv_p[row0 + r] = iv;
      if (inv_max_bits != nullptr && iv > 0.f) atomicMax(inv_max_bits + prob, __float_as_int(iv));   // positive floats order like their bits
    }
  }
  // u8 operands (garbage, and unused, when some value is not 0..255)
  for (int idx = tid; idx < PREP_ROWS * 32; idx += 256) {
    const int r = idx >> 5, k4 = (idx & 31) * 4;
    if (row0 + r >= rows_alloc) continue;
    uint32_t w = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const float v = (k4 + b < dim) ? tile[r * ldt + k4 + b] : 0.f;
      w |= (uint32_t)(__float2int_rz(fminf(fmaxf(v, 0.f), 255.f)) & 255) << (8 * b);
    }
    u8_p[idx] = w;
  }
}

// The common layout (row-major, dim = 128, optionally gathered rows) without shared memory: eight
// threads own a row, each loads four float4 (coalesced 128-byte segments), reduces its 16 squares and
// packs its 16 bytes.  Same results as match_prep_kernel.
__global__ void __launch_bounds__(256)
match_prep_rows128_kernel(const MatchOperand op, int rows_alloc, uint8_t* __restrict__ out_u8,
                          float* __restrict__ inv, int* __restrict__ nonint_flag, int* __restrict__ inv_max_bits) {
  const int prob = blockIdx.y;
  const int n = min(op.count[prob * op.count_stride], op.cap);
  const int row = blockIdx.x * PREP_ROWS + (threadIdx.x >> 3), part = threadIdx.x & 7;
  if (row >= rows_alloc) return;
  uint32_t* u8_row = reinterpret_cast<uint32_t*>(out_u8 + ((size_t)prob * rows_alloc + row) * 128);
  if (row >= n) {   // dead rows feed the MMA as zero operands
    if (part == 0) inv[(size_t)prob * rows_alloc + row] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) u8_row[part + 8 * q] = 0u;
    return;
  }
  const size_t src_row = op.gather ? (size_t)op.gather[(size_t)prob * op.gather_stride + row] : (size_t)row;
  const float* x = op.base + (size_t)prob * op.prob_stride + src_row * 128;
  float acc = 0.f;
  bool isint = true;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + part + 8 * q);
    const float e[4] = {v.x, v.y, v.z, v.w};
    uint32_t w = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      acc = fmaf(e[b], e[b], acc);
      isint = isint && (e[b] == truncf(e[b])) && (e[b] >= 0.f) && (e[b] <= 255.f);
      w |= (uint32_t)(__float2int_rz(fminf(fmaxf(e[b], 0.f), 255.f)) & 255) << (8 * b);
    }
    u8_row[part + 8 * q] = w;
  }
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, off);
    isint = __shfl_xor_sync(0xffffffffu, (int)isint, off) && isint;
  }
  if (part == 0) {
    if (!isint) {   // general floats: the oracle's sequential fold
      acc = 0.f;
      for (int k = 0; k < 128; ++k) acc = fmaf(x[k], x[k], acc);
      atomicOr(nonint_flag, 1);
    }
    const float iv = (acc == 0.f) ? 0.f : __fdiv_rn(1.0f, __fsqrt_rn(acc));
    inv[(size_t)prob * rows_alloc + row] = iv;
    if (inv_max_bits != nullptr && iv > 0.f) atomicMax(inv_max_bits + prob, __float_as_int(iv));
  }
}

// Runs only when the non-integer flag is set (general float descriptors): bf16 hi/lo split operands
// out[p][row][3*kp] (A: hi|hi|lo, B: hi|lo|hi) for match_topk_kernel and, for gathered or column-major
// inputs, the compact row-major fp32 copy the exact re-scoring reads.
__global__ void __launch_bounds__(256)
match_prep_split_kernel(const MatchOperand op, int rows_alloc, int dim, int kp, int is_b,
                        __nv_bfloat16* __restrict__ out, float* __restrict__ raw_copy, const int* __restrict__ nonint_flag) {
  extern __shared__ float tile[];
  if (*nonint_flag == 0) return;
  const int prob = blockIdx.y;
  const int n = min(op.count[prob * op.count_stride], op.cap);
  const int row0 = blockIdx.x * PREP_ROWS;
  if (row0 >= n) return;   // dead rows are masked by inv = 0
  const int ldt = dim + 1, tid = threadIdx.x;
  prep_load_tile(op, prob, row0, n, dim, tile);
  __syncthreads();
  const int ld_out = 3 * kp;
  __nv_bfloat16* out_p = out + (size_t)prob * rows_alloc * ld_out;
  float* raw_p = raw_copy ? raw_copy + (size_t)prob * rows_alloc * dim : nullptr;
  for (int idx = tid; idx < PREP_ROWS * kp; idx += 256) {
    const int r = idx / kp, k = idx - r * kp;
    if (row0 + r >= rows_alloc) continue;
    const float v = (k < dim) ? tile[r * ldt + k] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* o = out_p + (size_t)(row0 + r) * ld_out + k;
    o[0] = hi;
    o[kp] = is_b ? lo : hi;
    o[2 * kp] = is_b ? hi : lo;
    if (raw_p != nullptr && k < dim && row0 + r < n) raw_p[(size_t)(row0 + r) * dim + k] = v;
  }
}

// ------------------------------------------------------------------------- exact score math
__device__ __forceinline__ float exact_key(const float* __restrict__ a, const float* __restrict__ b, int dim, float invb) {
  float acc = 0.f;
  for (int k = 0; k < dim; ++k) acc = fmaf(a[k], b[k], acc);
  return __fmul_rn(acc, invb);
}
__device__ __forceinline__ float score_from_key(float key, float inva) {
  const float c = __fmul_rn(key, inva);
  const float s = fmaf(-2.0f, c, 2.0f);
  return s < 0.f ? 0.f : s;
}

__device__ __forceinline__ bool keep_row(float s1, float s2, int n2, float thr, float max_ratio) {
  if (!(s1 <= thr)) return false;
  if (n2 >= 2) {
    const float ratio = (s2 < 1e-6f) ? 1.0f : __fdiv_rn(s1, s2);
    if (!(ratio <= max_ratio)) return false;
  }
  return true;
}

struct RawRows { const float* base; size_t prob_stride; };   // compact row-major fp32 rows per problem

// thread per row: merge candidate slots, exact re-rank, certify or flag for the row scan
__global__ void __launch_bounds__(128)
match_finalize_kernel(const uint2* __restrict__ cand_base, size_t cand_stride, int n_slots, RawRows ra, RawRows rb,
                      int dim, const float* __restrict__ inva_base, int inva_stride,
                      const float* __restrict__ invb_base, int invb_stride, const int* __restrict__ n1p,
                      int n1_stride, int cap1, const int* __restrict__ n2p, int n2_stride, int cap2,
                      const int* __restrict__ nonint_flag, int all_slots, int sched_L, int sched_T, MatchFilter flt,
                      float gen_eps, int general_skipped, uint32_t* __restrict__ j1_out,
                      float* __restrict__ s1_out, float* __restrict__ s2_out, int row_stride,
                      int* __restrict__ scan_list, int* __restrict__ scan_count) {
  const int prob = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n1 = min(n1p[prob * n1_stride], cap1), n2 = min(n2p[prob * n2_stride], cap2);
  if (i >= n1) return;
  const size_t orow = (size_t)prob * row_stride + i;
  if (n2 <= 0) { j1_out[orow] = 0xFFFFFFFFu; s1_out[orow] = INFINITY; s2_out[orow] = INFINITY; return; }
  float k[3] = {-INFINITY, -INFINITY, -INFINITY};
  uint32_t j[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
  // row stride is n_slots (thin kernel: 4 slots per split); the fat kernel fills the first half only
  const uint2* c = cand_base + (size_t)prob * cand_stride + (size_t)i * n_slots * NCAND;
  int used = ((*nonint_flag) || all_slots) ? n_slots * NCAND : (n_slots / 2) * NCAND;
  if (sched_L > 0 && !(*nonint_flag)) {   // persistent u8 kernel: the row's panel was cut into n_seg segments, 2 slots each
    const long long p0 = (long long)(i / UBM) * sched_T;
    used = (int)((p0 + sched_T - 1) / sched_L - p0 / sched_L + 1) * 2 * NCAND;
  }
  for (int s = 0; s < used; ++s) {
    const uint2 e = c[s];
    if (e.y >= (uint32_t)n2) continue;
    const float key = __uint_as_float(e.x);
    int pos = 3;
    for (int q = 2; q >= 0; --q)
      if (key > k[q] || (key == k[q] && e.y < j[q])) pos = q;
    if (pos < 3) {
      for (int q = 2; q > pos; --q) { k[q] = k[q - 1]; j[q] = j[q - 1]; }
      k[pos] = key; j[pos] = e.y;
    }
  }
  const float ia = inva_base[(size_t)prob * inva_stride + i];
  const float* invb = invb_base + (size_t)prob * invb_stride;
  const float* a = ra.base + (size_t)prob * ra.prob_stride + (size_t)i * dim;
  const float* b_rows = rb.base + (size_t)prob * rb.prob_stride;
  // exact-integer path: the tensor-core dot is the oracle's dot, so the epilogue key is already exact;
  // split path: re-score the candidates with the oracle's sequential FP32 dot
  const bool general = (*nonint_flag) != 0;
  if (general && (rb.base == nullptr || general_skipped)) {   // non-integer rows where the caller promised integers (general kernels not launched), or
                                                              // non-integer queries against a prepared (u8-only) landmark set: nothing to re-score with
    j1_out[orow] = 0xFFFFFFFFu; s1_out[orow] = INFINITY; s2_out[orow] = INFINITY;
    return;
  }
  float s[3]; int nc = 0;
  for (int q = 0; q < 3; ++q) {
    if (j[q] == 0xFFFFFFFFu) { s[q] = INFINITY; continue; }
    const float key = general ? exact_key(a, b_rows + (size_t)j[q] * dim, dim, invb[j[q]]) : k[q];
    s[q] = score_from_key(key, ia);
    ++nc;
  }
  uint32_t jj[3] = {j[0], j[1], j[2]};
  for (int p = 0; p < 2; ++p)
    for (int q = 0; q < 2 - p; ++q)
      if (s[q + 1] < s[q] || (s[q + 1] == s[q] && jj[q + 1] < jj[q])) {
        const float ts = s[q]; s[q] = s[q + 1]; s[q + 1] = ts;
        const uint32_t tj = jj[q]; jj[q] = jj[q + 1]; jj[q + 1] = tj;
      }
  // Every column that is not a candidate has key <= kb: the third-best candidate key (each epilogue
  // thread only skipped keys at or below a lower bound of it) or, on the exact-integer path with a
  // score bound, the row's initial key bound.  The score is a monotone function of the key, so every
  // such column has score >= sb.
  // General floats: the candidate keys are approximate (|approx - oracle| <= gen_eps * ||a||), the candidates' scores
  // above are exact; a skipped column's oracle key is at most the bound plus that margin.
  float kb = (nc >= 3) ? k[2] : -INFINITY;
  const bool bounded = flt.key_floor > 0.f && ia > 0.f;
  if (bounded) kb = fmaxf(kb, general ? general_key_floor(flt.key_floor, ia, gen_eps) : __fdiv_rn(flt.key_floor, ia));
  bool certified;
  float s2v = (n2 >= 2) ? s[1] : INFINITY;
  if (kb == -INFINITY) {
    certified = (nc == n2);            // nothing was skipped: the candidates are all the columns
  } else {
    if (general) kb = kb + gen_eps * ((ia > 0.f) ? __fdiv_rn(1.0f, ia) : 0.f);
    const float sb = score_from_key(kb, ia);
    if (!bounded) {
      certified = (nc >= 2) && (sb > s[0]) && (sb >= s[1]);     // (j1, s1, s2) are exactly the oracle's
    } else if (fminf(s[0], sb) > flt.thr_score) {
      certified = true;                  // every score is above the match threshold: the row is rejected
    } else if (!(sb > s[0])) {
      certified = false;
    } else if (nc >= 2 && sb >= s[1]) {
      certified = true;                  // exact top-2
    } else {
      // j1 and s1 are exact; the true s2 lies in [sb, s[1]].  The row's outcome is already decided when
      // it fails the threshold, or passes the ratio test even against the smallest possible s2.
      s2v = (n2 >= 2) ? sb : INFINITY;
      certified = !(s[0] <= flt.thr_score) || keep_row(s[0], s2v, n2, flt.thr_score, flt.max_ratio);
    }
  }
  j1_out[orow] = jj[0]; s1_out[orow] = s[0]; s2_out[orow] = s2v;
  if (!certified) {
    const int slot = atomicAdd(scan_count, 1);
    scan_list[slot] = prob * row_stride + i;
  }
}

// exact scan of every column for the rows in scan_list (block per row, grid-stride).  General floats:
// the oracle's sequential FP32 dot on the raw rows.  Exact-integer inputs: the u8 operands and an
// integer dot product, which IS the oracle's dot (every partial sum < 2^24).
__global__ void __launch_bounds__(256)
match_rowscan_kernel(const int* __restrict__ scan_list, const int* __restrict__ scan_count, RawRows ra, RawRows rb,
                     const uint8_t* __restrict__ u8a, const uint8_t* __restrict__ u8b, int a_alloc, int b_alloc,
                     const int* __restrict__ nonint_flag, int dim, const float* __restrict__ inva_base, int inva_stride,
                     const float* __restrict__ invb_base, int invb_stride, const int* __restrict__ n2p,
                     int n2_stride, int cap2, uint32_t* __restrict__ j1_out, float* __restrict__ s1_out,
                     float* __restrict__ s2_out, int row_stride) {
  __shared__ float sa[256];
  __shared__ uint32_t sa8[32];
  __shared__ float rs1[256], rs2[256];
  __shared__ uint32_t rj1[256];
  const int count = *scan_count;
  const bool general = (*nonint_flag) != 0;
  for (int f = blockIdx.x; f < count; f += gridDim.x) {
    const int code = scan_list[f];
    const int prob = code / row_stride, i = code - prob * row_stride;
    const int n2 = min(n2p[prob * n2_stride], cap2);
    const float* invb = invb_base + (size_t)prob * invb_stride;
    __syncthreads();
    if (general) {
      for (int k = threadIdx.x; k < dim; k += blockDim.x) sa[k] = ra.base[(size_t)prob * ra.prob_stride + (size_t)i * dim + k];
    } else if (threadIdx.x < 32) {
      sa8[threadIdx.x] = reinterpret_cast<const uint32_t*>(u8a + ((size_t)prob * a_alloc + i) * 128)[threadIdx.x];
    }
    __syncthreads();
    const float ia = inva_base[(size_t)prob * inva_stride + i];
    float b1 = INFINITY, b2 = INFINITY; uint32_t bj = 0xFFFFFFFFu;
    for (int jx = threadIdx.x; jx < n2; jx += blockDim.x) {
      float key;
      if (general) {
        key = exact_key(sa, rb.base + (size_t)prob * rb.prob_stride + (size_t)jx * dim, dim, invb[jx]);
      } else {
        const uint4* bw = reinterpret_cast<const uint4*>(u8b + ((size_t)prob * b_alloc + jx) * 128);
        unsigned dot = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint4 w = bw[q];
          dot = __dp4a(sa8[4 * q], w.x, dot); dot = __dp4a(sa8[4 * q + 1], w.y, dot);
          dot = __dp4a(sa8[4 * q + 2], w.z, dot); dot = __dp4a(sa8[4 * q + 3], w.w, dot);
        }
        key = __fmul_rn((float)(int)dot, invb[jx]);
      }
      const float sc = score_from_key(key, ia);
      if (sc < b1) { b2 = b1; b1 = sc; bj = (uint32_t)jx; }
      else if (sc < b2) b2 = sc;
    }
    rs1[threadIdx.x] = b1; rs2[threadIdx.x] = b2; rj1[threadIdx.x] = bj;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if (threadIdx.x < off) {
        const float o1 = rs1[threadIdx.x + off], o2 = rs2[threadIdx.x + off];
        const uint32_t oj = rj1[threadIdx.x + off];
        float m1 = rs1[threadIdx.x], m2 = rs2[threadIdx.x]; uint32_t mj = rj1[threadIdx.x];
        if (o1 < m1 || (o1 == m1 && oj < mj)) { m2 = fminf(m1, o2); m1 = o1; mj = oj; }
        else { m2 = fminf(m2, o1); }
        rs1[threadIdx.x] = m1; rs2[threadIdx.x] = m2; rj1[threadIdx.x] = mj;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const size_t orow = (size_t)prob * row_stride + i;
      j1_out[orow] = rj1[0]; s1_out[orow] = rs1[0]; s2_out[orow] = (n2 >= 2) ? rs2[0] : INFINITY;
    }
  }
}

// ---------------------------------------------------------------------- select + compaction
// one block per problem: ordered compaction of the surviving rows (batched path, n1 <= ~16k)
__global__ void __launch_bounds__(1024)
match_select_block_kernel(const uint32_t* __restrict__ j1, const float* __restrict__ s1, const float* __restrict__ s2,
                          int row_stride, const int* __restrict__ n1p, int n1_stride, int cap1,
                          const int* __restrict__ n2p, int n2_stride, int cap2, float thr, float max_ratio,
                          int index_base, uint32_t* __restrict__ idx1, uint32_t* __restrict__ idx2,
                          float* __restrict__ metric, int out_stride, int* __restrict__ n_pairs, int np_stride,
                          const uint32_t* __restrict__ back_j1, int back_stride) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int prob = blockIdx.x;
  const int n1 = min(n1p[prob * n1_stride], cap1), n2 = min(n2p[prob * n2_stride], cap2);
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < n1 && n2 > 0; base += 1024) {
    const int i = base + threadIdx.x;
    const size_t row = (size_t)prob * row_stride + i;
    bool keep = (i < n1) && keep_row(s1[row], s2[row], n2, thr, max_ratio);
    // Unique: the landmark's own nearest query row must be this row (forward-backward consistency)
    if (keep && back_j1 != nullptr) keep = back_j1[(size_t)prob * back_stride + j1[row]] == (uint32_t)i;
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_sums[w] = __popc(ballot);
    __syncthreads();
    if (w == 0) {
      int x = warp_sums[lane];
      for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
      }
      warp_sums[lane] = x;
    }
    __syncthreads();
    if (keep) {
      const int pos = carry + (w ? warp_sums[w - 1] : 0) + __popc(ballot & ((1u << lane) - 1u));
      const size_t o = (size_t)prob * out_stride + pos;
      idx1[o] = (uint32_t)i + (uint32_t)index_base;
      idx2[o] = j1[row] + (uint32_t)index_base;
      if (metric) metric[o] = s1[row];
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_sums[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) n_pairs[prob * np_stride] = carry;
}

// large single problem: count / scan / scatter over 1024-row blocks
constexpr int SEL_BLOCK = 1024;
__global__ void __launch_bounds__(SEL_BLOCK)
match_count_kernel(const uint32_t* __restrict__ j1, const float* __restrict__ s1, const float* __restrict__ s2,
                   const uint32_t* __restrict__ back_j1, const int* __restrict__ n1p,
                   const int* __restrict__ n2p, float thr, float max_ratio, int* __restrict__ block_counts) {
  const int i = blockIdx.x * SEL_BLOCK + threadIdx.x;
  const int n1 = *n1p, n2 = *n2p;
  bool keep = false;
  if (i < n1 && n2 > 0) {
    keep = keep_row(s1[i], s2[i], n2, thr, max_ratio);
    if (keep && back_j1 != nullptr) keep = (back_j1[j1[i]] == (uint32_t)i);
  }
  const int cnt = __syncthreads_count(keep);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(1024)
match_scan_kernel(int* __restrict__ block_counts, int n_blocks, int* __restrict__ total) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    const int idx = base + threadIdx.x;
    const int v = idx < n_blocks ? block_counts[idx] : 0;
    int x = v;
    for (int off = 1; off < 32; off <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, off);
      if ((threadIdx.x & 31) >= off) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_sums[threadIdx.x];
      for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, off);
        if (threadIdx.x >= off) w += y;
      }
      warp_sums[threadIdx.x] = w;
    }
    __syncthreads();
    const int warp_off = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0;
    const int incl = x + warp_off + carry;
    if (idx < n_blocks) block_counts[idx] = incl - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SEL_BLOCK)
match_scatter_kernel(const uint32_t* __restrict__ j1, const float* __restrict__ s1, const float* __restrict__ s2,
                     const uint32_t* __restrict__ back_j1, const int* __restrict__ n1p,
                     const int* __restrict__ n2p, float thr, float max_ratio, int index_base,
                     const int* __restrict__ block_offsets, uint32_t* __restrict__ idx1,
                     uint32_t* __restrict__ idx2, float* __restrict__ metric) {
  __shared__ int warp_sums[32];
  const int i = blockIdx.x * SEL_BLOCK + threadIdx.x;
  const int n1 = *n1p, n2 = *n2p;
  bool keep = false;
  if (i < n1 && n2 > 0) {
    keep = keep_row(s1[i], s2[i], n2, thr, max_ratio);
    if (keep && back_j1 != nullptr) keep = (back_j1[j1[i]] == (uint32_t)i);
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) warp_sums[w] = __popc(ballot);
  __syncthreads();
  if (w == 0) {
    int x = warp_sums[lane];
    for (int off = 1; off < 32; off <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += y;
    }
    warp_sums[lane] = x;
  }
  __syncthreads();
  if (keep) {
    const int pos = block_offsets[blockIdx.x] + (w ? warp_sums[w - 1] : 0) + __popc(ballot & ((1u << lane) - 1u));
    idx1[pos] = (uint32_t)i + (uint32_t)index_base;
    idx2[pos] = j1[i] + (uint32_t)index_base;
    if (metric) metric[pos] = s1[i];
  }
}

// 16-byte best-2 record per query row (the unit the relocalisation shards all-gather):
// {j1, s1, s2, keep} with keep = 1 iff the row passes the threshold and ratio tests.
__global__ void __launch_bounds__(256)
match_records_kernel(const uint32_t* __restrict__ j1, const float* __restrict__ s1, const float* __restrict__ s2,
                     int n1, int n2, float thr, float max_ratio, int index_base, uint4* __restrict__ rec) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n1) return;
  const bool keep = n2 > 0 && keep_row(s1[i], s2[i], n2, thr, max_ratio);
  rec[i] = make_uint4(keep ? j1[i] + (uint32_t)index_base : 0xFFFFFFFFu, __float_as_uint(s1[i]), __float_as_uint(s2[i]), keep ? 1u : 0u);
}

// The same record written straight into the gathered array of EVERY rank: peers.p[r] is rank r's buffer (this rank's
// own one included), mapped into this process with CUDA IPC, so the stores to the other GPUs travel over NVLink /
// NVSwitch while the kernel runs -- the all-gather of the relocalisation shard is the epilogue's own stores and no
// collective moves data afterwards (one barrier tells the ranks that every slice has landed).
struct PeerTable { uint4* p[16]; int n; };
__global__ void __launch_bounds__(256)
match_records_gather_kernel(const uint32_t* __restrict__ j1, const float* __restrict__ s1, const float* __restrict__ s2,
                            int n1, int n2, float thr, float max_ratio, int index_base, const PeerTable peers, size_t row0) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n1) return;
  const bool keep = n2 > 0 && keep_row(s1[i], s2[i], n2, thr, max_ratio);
  const uint4 rec = make_uint4(keep ? j1[i] + (uint32_t)index_base : 0xFFFFFFFFu, __float_as_uint(s1[i]), __float_as_uint(s2[i]), keep ? 1u : 0u);
#pragma unroll 4
  for (int r = 0; r < peers.n; ++r) peers.p[r][row0 + i] = rec;
}

// --------------------------------------------------------------------------- host plumbing
static int make_operand_map(CUtensorMap* map, const __nv_bfloat16* base, int rows_alloc, int kp, int n_prob, int box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled driver entry point not available"); return VO_ERR_CUDA; }
  cuuint64_t gdim[3] = {(cuuint64_t)(3 * kp), (cuuint64_t)rows_alloc, (cuuint64_t)n_prob};
  cuuint64_t gstride[2] = {(cuuint64_t)(3 * kp) * sizeof(__nv_bfloat16),
                           (cuuint64_t)rows_alloc * (3 * kp) * sizeof(__nv_bfloat16)};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return VO_ERR_CUDA; }
  return VO_OK;
}

// u8 operands [n_prob][rows_alloc][128]: one 128-row x 128-byte box = one 128B-swizzled K block
static int make_u8_map(CUtensorMap* map, const uint8_t* base, int rows_alloc, int n_prob, int box_rows = 128) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled driver entry point not available"); return VO_ERR_CUDA; }
  cuuint64_t gdim[3] = {128, (cuuint64_t)rows_alloc, (cuuint64_t)n_prob};
  cuuint64_t gstride[2] = {128, (cuuint64_t)rows_alloc * 128};
  cuuint32_t box[3] = {128, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (u8) failed (%d)", (int)r); return VO_ERR_CUDA; }
  return VO_OK;
}

constexpr bool MATCH_PAIRS_DEFAULT = false;
constexpr bool MATCH_TS_DEFAULT = false;

MatchFilter make_match_filter(const vo_match_opts& o) {
  MatchFilter f{0.f, 0.f, 0.f};
  f.thr_score = o.match_threshold * 0.04f;
  f.max_ratio = o.max_ratio;
  // scores above thr_score / max_ratio cannot change a row's outcome; as a key bound (c = 1 - s/2),
  // loosened by 2^-10 so that match_finalize_kernel's exact check passes with room to spare
  const double s_bound = (o.max_ratio > 0.f) ? (double)f.thr_score / (double)o.max_ratio : 4.0;
  if (s_bound < 1.9) f.key_floor = (float)((1.0 - 0.5 * s_bound) * (1.0 - 1.0 / 1024.0));
  return f;
}

// Converts the landmark rows of ONE problem once (u8 operand rows, 1/||row||, their maximum) into the scratch buffers
// that match_batch_top2 uses under `tag`, and keeps the maximum aside: later calls under the same tag with
// B.prepared = 1 skip the conversion (the float rows are then not read at all).  Integer-valued rows only.
int match_prepare_landmarks(vo_ctx* ctx, const MatchOperand& B, int dim, const char* tag, cudaStream_t st) {
  auto nm = [&](const char* base) { return std::string(base) + "_" + tag; };
  if (dim != 128) { set_error("vo_landmarks_prepare: dim must be 128"); return VO_ERR_ARG; }
  const int b_alloc = div_up(B.cap > 0 ? B.cap : 1, BN) * BN;
  int* ctl; VO_TRY(dev_buf(ctx, nm("m_ctl").c_str(), 8 + (size_t)1, &ctl));
  VO_CUDA(cudaMemsetAsync(ctl, 0, 9 * sizeof(int), st));
  int* keep_b; VO_TRY(dev_buf(ctx, nm("m_keepB").c_str(), (size_t)1, &keep_b));
  uint8_t* u8B; VO_TRY(dev_buf(ctx, nm("m_u8B").c_str(), (size_t)b_alloc * 128, &u8B));
  float* invB; VO_TRY(dev_buf(ctx, nm("m_invB").c_str(), (size_t)b_alloc, &invB));
  if ((reinterpret_cast<uintptr_t>(B.base) & 15) == 0)
    match_prep_rows128_kernel<<<dim3(b_alloc / PREP_ROWS, 1), 256, 0, st>>>(B, b_alloc, u8B, invB, ctl, ctl + 8);
  else {
    const size_t prep_smem = (size_t)PREP_ROWS * (dim + 1) * sizeof(float);
    match_prep_kernel<<<dim3(b_alloc / PREP_ROWS, 1), 256, prep_smem, st>>>(B, b_alloc, dim, u8B, invB, ctl, ctl + 8);
  }
  ctx->kernel_launches += 1;
  VO_CUDA(cudaMemcpyAsync(keep_b, ctl + 8, sizeof(int), cudaMemcpyDeviceToDevice, st));
  int* h; VO_TRY(pin_buf(ctx, "m_prep_flag", 4, &h));
  VO_CUDA(cudaMemcpyAsync(h, ctl, sizeof(int), cudaMemcpyDeviceToHost, st));
  VO_CUDA(cudaStreamSynchronize(st));
  if (h[0] != 0) { set_error("vo_landmarks_prepare: the rows are not integers in 0..255 (prepared landmark sets use the exact-integer path)"); return VO_ERR_ARG; }
  return VO_OK;
}

int match_batch_top2(vo_ctx* ctx, const MatchOperand& A, const MatchOperand& B, int n_prob, int dim,
                     const char* tag, float* dbg_c, cudaStream_t st, MatchTop2* out, const MatchFilter* filter) {
  const MatchFilter flt = filter ? *filter : MatchFilter{0.f, 0.f, 0.f};
  auto nm = [&](const char* base) { return std::string(base) + "_" + tag; };
  const int kp = div_up(dim, BK) * BK;
  if (dim > 128) { set_error("vo_match: dim %d > 128 is not supported by the tensor-core path", dim); return VO_ERR_ARG; }
  const int a_alloc = div_up(A.cap > 0 ? A.cap : 1, UBM) * UBM;   // whole 256-row panels (u8 kernel)
  const int b_tiles = div_up(B.cap > 0 ? B.cap : 1, BN);
  const int b_alloc = b_tiles * BN;   // multiple of 256, so the epilogue's 32-wide invb loads stay in bounds

  int* ctl; VO_TRY(dev_buf(ctx, nm("m_ctl").c_str(), 8 + (size_t)(n_prob > 0 ? n_prob : 1), &ctl));
  VO_CUDA(cudaMemsetAsync(ctl, 0, (8 + (size_t)(n_prob > 0 ? n_prob : 1)) * sizeof(int), st));
  int* invb_max = ctl + 8;   // [n_prob] bit pattern of max_j 1/||b_j||
  int* keep_b; VO_TRY(dev_buf(ctx, nm("m_keepB").c_str(), (size_t)(n_prob > 0 ? n_prob : 1), &keep_b));   // invb_max of a prepared B
  if (B.prepared) VO_CUDA(cudaMemcpyAsync(invb_max, keep_b, (size_t)n_prob * sizeof(int), cudaMemcpyDeviceToDevice, st));
  uint8_t *u8A, *u8B;
  VO_TRY(dev_buf(ctx, nm("m_u8A").c_str(), (size_t)(n_prob > 0 ? n_prob : 1) * a_alloc * 128, &u8A));
  __nv_bfloat16 *opA, *opB; float *invA, *invB, *rawA = nullptr, *rawB = nullptr;
  VO_TRY(dev_buf(ctx, nm("m_opA").c_str(), (size_t)n_prob * a_alloc * 3 * kp, &opA));
  VO_TRY(dev_buf(ctx, nm("m_opB").c_str(), (size_t)n_prob * b_alloc * 3 * kp, &opB));
  VO_TRY(dev_buf(ctx, nm("m_u8B").c_str(), (size_t)(n_prob > 0 ? n_prob : 1) * b_alloc * 128, &u8B));
  VO_TRY(dev_buf(ctx, nm("m_invA").c_str(), (size_t)n_prob * a_alloc, &invA));
  VO_TRY(dev_buf(ctx, nm("m_invB").c_str(), (size_t)n_prob * b_alloc, &invB));
  RawRows ra{A.base, A.prob_stride}, rb{B.base, B.prob_stride};
  if (A.gather || A.col_major) {
    VO_TRY(dev_buf(ctx, nm("m_rawA").c_str(), (size_t)n_prob * a_alloc * dim, &rawA));
    ra = RawRows{rawA, (size_t)a_alloc * dim};
  }
  if (B.gather || B.col_major) {
    VO_TRY(dev_buf(ctx, nm("m_rawB").c_str(), (size_t)n_prob * b_alloc * dim, &rawB));
    rb = RawRows{rawB, (size_t)b_alloc * dim};
  }
  VO_TRY(dev_buf(ctx, nm("m_j1").c_str(), (size_t)n_prob * a_alloc, &out->j1));
  VO_TRY(dev_buf(ctx, nm("m_s1").c_str(), (size_t)n_prob * a_alloc, &out->s1));
  VO_TRY(dev_buf(ctx, nm("m_s2").c_str(), (size_t)n_prob * a_alloc, &out->s2));
  out->row_stride = a_alloc; out->ctl = ctl;
  ctx->match_stats[2] = 0; ctx->match_stats[3] = kp;
  if (A.cap <= 0 || n_prob <= 0) return VO_OK;

  const size_t prep_smem = (size_t)PREP_ROWS * (dim + 1) * sizeof(float);
  const bool skip_general = A.integer_rows && B.integer_rows && dbg_c == nullptr;
  const bool exact_sizes = (n_prob == 1 && !A.gather && !B.gather);   // caps are the live sizes
  {
    ProfScope ps(ctx, st, "match_prep", exact_sizes ? ((double)A.cap + B.cap) * (dim * 4.0 + 132.0) : 0.0, 0.0, (skip_general ? 2 : 4) - (B.prepared ? (skip_general ? 1 : 2) : 0));
    if (dim == 128 && !A.col_major && (reinterpret_cast<uintptr_t>(A.base) & 15) == 0 && A.prob_stride % 4 == 0)
      match_prep_rows128_kernel<<<dim3(a_alloc / PREP_ROWS, n_prob), 256, 0, st>>>(A, a_alloc, u8A, invA, ctl, nullptr);
    else
      match_prep_kernel<<<dim3(a_alloc / PREP_ROWS, n_prob), 256, prep_smem, st>>>(A, a_alloc, dim, u8A, invA, ctl, nullptr);
    if (B.prepared) {
      // the landmark rows were converted once (vo_landmarks_prepare): nothing to read, nothing to write
    } else if (dim == 128 && !B.col_major && (reinterpret_cast<uintptr_t>(B.base) & 15) == 0 && B.prob_stride % 4 == 0)
      match_prep_rows128_kernel<<<dim3(b_alloc / PREP_ROWS, n_prob), 256, 0, st>>>(B, b_alloc, u8B, invB, ctl, invb_max);
    else
      match_prep_kernel<<<dim3(b_alloc / PREP_ROWS, n_prob), 256, prep_smem, st>>>(B, b_alloc, dim, u8B, invB, ctl, invb_max);
    // general float descriptors only (both return at once when every value is an integer 0..255; not launched at all
    // when the caller vouches for integer rows: the frame loop matches descriptors its own SIFT kernel has just written)
    if (!skip_general) {
      match_prep_split_kernel<<<dim3(a_alloc / PREP_ROWS, n_prob), 256, prep_smem, st>>>(A, a_alloc, dim, kp, 0, opA, rawA, ctl);
      if (!B.prepared) match_prep_split_kernel<<<dim3(b_alloc / PREP_ROWS, n_prob), 256, prep_smem, st>>>(B, b_alloc, dim, kp, 1, opB, rawB, ctl);
    }
  }
  VO_CUDA(cudaGetLastError());

  // column splits: equal-sized CTAs run in waves of num_sms, so pick the split count (<= 8) that
  // minimises waves x (tiles per CTA + a fixed per-CTA prologue of ~6 tile times)
  const int m_blocks = a_alloc / UBM;
  const int u_tiles = div_up(B.cap > 0 ? B.cap : 1, UBN);
  int n_splits = 1;
  {
    long long best = -1;
    for (int sp = 1; sp <= 8 && sp <= u_tiles; ++sp) {
      const long long ctas = (long long)m_blocks * n_prob * sp;
      const long long cost = ((ctas + ctx->num_sms - 1) / ctx->num_sms) * (div_up(u_tiles, sp) + 6);
      if (best < 0 || cost < best) { best = cost; n_splits = sp; }
    }
    // many short problems (the frame loop): waves are plentiful, fewer and longer CTAs amortise best
    if ((long long)m_blocks * n_prob >= 4LL * ctx->num_sms) n_splits = 1;
    if (n_splits > b_tiles) n_splits = b_tiles;
  }
  // CTA-pair (cta_group::2) kernel: VO_MATCH_PAIRS=0/1 overrides the default
  static const int pairs_env = [] { const char* e = getenv("VO_MATCH_PAIRS"); return e ? atoi(e) : -1; }();
  const bool use_pairs = pairs_env >= 0 ? pairs_env != 0 : MATCH_PAIRS_DEFAULT;
  static const int ts_env = [] { const char* e = getenv("VO_MATCH_TS"); return e ? atoi(e) : -1; }();
  const bool use_ts = !use_pairs && (ts_env >= 0 ? ts_env != 0 : MATCH_TS_DEFAULT);
  if (use_ts) {
    const int t_tiles = div_up(B.cap > 0 ? B.cap : 1, TN);
    if (n_splits > t_tiles) n_splits = t_tiles;
  }
  static const int pair_tile = [] { const char* e = getenv("VO_MATCH_PAIR_TILE"); return (e && atoi(e) == 256) ? 256 : 128; }();
  if (use_pairs) {
    const int x_tiles = div_up(B.cap > 0 ? B.cap : 1, pair_tile);
    if (n_splits > x_tiles) n_splits = x_tiles;
  }
  // persistent form of the u8 kernel: one problem whose sizes the host knows and enough tiles to share out
  int sched_L = 0, sched_T = 0; long long sched_total = 0;
  if (exact_sizes && !use_pairs && !use_ts && B.cap > 0) {
    const int T = div_up(B.cap, UBN);
    const long long total = (long long)div_up(A.cap, UBM) * T;
    // measured: the equal shares pay off once a CTA's share is long enough to amortise the extra segment
    // start (A reload, threshold reset; without a score bound also a second "everything is a candidate" phase)
    const long long share = (total + ctx->num_sms - 1) / ctx->num_sms;
    // ... or when the grid form could not even occupy half the SMs (few query rows against a long landmark list)
    const bool skinny = (long long)m_blocks * 8 < ctx->num_sms / 2 && total >= ctx->num_sms;
    if (share >= (flt.key_floor > 0.f ? 128 : 512) || skinny) {
      sched_T = T; sched_total = total;
      sched_L = (int)share;
      const int s_max = (T + sched_L - 1) / sched_L + 1;          // segments a panel can be cut into
      n_splits = std::max(n_splits, div_up(2 * s_max, 4));        // candidate row stride (4 slots per split)
      // the persistent kernel does not use the split count beyond that stride, so it is free to suit the grid of the
      // general-float kernel (128-row panels x splits, waves of num_sms; about three tile times of prologue per CTA)
      const int panels = a_alloc / BM;
      long long best = -1; int best_sp = n_splits;
      for (int sp = n_splits; sp <= 8 && sp <= b_tiles; ++sp) {
        const long long ctas = (long long)panels * sp;
        const long long cost = ((ctas + ctx->num_sms - 1) / ctx->num_sms) * (div_up(b_tiles, sp) + 3);
        if (best < 0 || cost < best) { best = cost; best_sp = sp; }
      }
      n_splits = best_sp;
    }
  }
  // general-float path: with a score bound (matchFeatures) the hi x hi term alone decides almost every row and the
  // certificate carries the larger margin; without one (exact top-2 of every row) all three terms are contracted.
  // The debug GEMM always contracts the three terms.  VO_MATCH_FLOAT_TERMS=3 forces the three-term form.
  static const int terms_env = [] { const char* e = getenv("VO_MATCH_FLOAT_TERMS"); return e ? atoi(e) : 0; }();
  const int gen_terms = (flt.key_floor > 0.f && dbg_c == nullptr && terms_env != 3) ? 1 : 3;
  const float gen_eps = gen_terms == 1 ? ONE_TERM_EPS : SPLIT_EPS;
  ctx->match_float_terms = gen_terms;
  const int n_slots = n_splits * 4;
  const size_t cand_stride = (size_t)a_alloc * n_slots * NCAND;
  uint2* cand; VO_TRY(dev_buf(ctx, nm("m_cand").c_str(), (size_t)n_prob * cand_stride, &cand));
  int* scan_list; VO_TRY(dev_buf(ctx, nm("m_scanlist").c_str(), (size_t)n_prob * a_alloc, &scan_list));

  if (B.cap > 0) {
    CUtensorMap tmA, tmB, tmA8, tmB8;
    VO_TRY(make_operand_map(&tmA, opA, a_alloc, kp, n_prob, BM));
    VO_TRY(make_operand_map(&tmB, opB, b_alloc, kp, n_prob, BN));
    VO_TRY(make_u8_map(&tmA8, u8A, a_alloc, n_prob));
    VO_TRY(make_u8_map(&tmB8, u8B, b_alloc, n_prob));
    VO_TRY(ensure_dyn_smem_of(match_topk_u8ts_kernel, T_SMEM_BYTES));
    VO_TRY(ensure_dyn_smem_of(match_topk_u8x2_kernel<128>, X_SMEM_BYTES));
    VO_TRY(ensure_dyn_smem_of(match_topk_u8x2_kernel<256>, X_SMEM_BYTES));
    VO_TRY(ensure_dyn_smem_of(match_topk_kernel, SMEM_BYTES));
    VO_TRY(ensure_dyn_smem_of(match_topk_u8_kernel, U_SMEM_BYTES));
    // both variants are launched; each reads the device-side "non-integer input" flag and one of them
    // returns at once (no host synchronisation to pick the path)
    ProfScope ps(ctx, st, "match_gemm_topk", 0.0, exact_sizes ? 2.0 * A.cap * B.cap * dim : 0.0, skip_general ? 1 : 2);
    if (!skip_general) match_topk_kernel<<<dim3(a_alloc / BM, n_splits, n_prob), NUM_THREADS, SMEM_BYTES, st>>>(
        tmA, tmB, invB, b_alloc, A.count, A.count_stride, A.cap, B.count, B.count_stride, B.cap, ctl, kp / BK, n_splits, cand,
        cand_stride, dbg_c, B.cap, invA, a_alloc, flt.key_floor, gen_terms, gen_eps, invb_max);
    if (use_ts) {
      CUtensorMap tmB96;
      VO_TRY(make_u8_map(&tmB96, u8B, b_alloc, n_prob, TN));
      match_topk_u8ts_kernel<<<dim3(m_blocks, n_splits, n_prob), T_THREADS, T_SMEM_BYTES, st>>>(
          tmB96, u8A, a_alloc, invB, b_alloc, invA, a_alloc, invb_max, A.count, A.count_stride, A.cap, B.count, B.count_stride,
          B.cap, ctl, n_splits, cand, cand_stride, n_slots, flt.key_floor, dbg_c, B.cap);
    } else if (!use_pairs) {
      const dim3 ugrid = sched_L > 0 ? dim3((unsigned)div_up((int)((sched_total + sched_L - 1) / sched_L), 1), 1, 1) : dim3(m_blocks, n_splits, n_prob);
      match_topk_u8_kernel<<<ugrid, NUM_THREADS, U_SMEM_BYTES, st>>>(
          tmA8, tmB8, invB, b_alloc, invA, a_alloc, invb_max, A.count, A.count_stride, A.cap, B.count, B.count_stride, B.cap,
          ctl, n_splits, cand, cand_stride, n_slots, flt.key_floor, dbg_c, B.cap, sched_L, sched_T, sched_total);
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * m_blocks, n_splits, n_prob);   // one CTA pair (cluster of 2) per 256-row panel
      cfg.blockDim = dim3(NUM_THREADS);
      cfg.dynamicSmemBytes = X_SMEM_BYTES;
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      static const bool clk_on = getenv("VO_MATCH_CLK") != nullptr;
      long long* clk = nullptr;
      const size_t n_cta = (size_t)cfg.gridDim.x * cfg.gridDim.y * cfg.gridDim.z;
      if (clk_on) { VO_TRY(dev_buf(ctx, "m_clk", 4 * n_cta, &clk)); VO_CUDA(cudaMemsetAsync(clk, 0, 32 * n_cta, st)); }
      if (pair_tile == 128) {
        CUtensorMap tmB64;
        VO_TRY(make_u8_map(&tmB64, u8B, b_alloc, n_prob, 64));
        VO_CUDA(cudaLaunchKernelEx(&cfg, match_topk_u8x2_kernel<128>, tmA8, tmB64, (const float*)invB, b_alloc, (const float*)invA, a_alloc,
                                   (const int*)invb_max, A.count, A.count_stride, A.cap, B.count, B.count_stride, B.cap,
                                   (const int*)ctl, n_splits, cand, cand_stride, n_slots, flt.key_floor, dbg_c, B.cap, clk));
      } else {
        VO_CUDA(cudaLaunchKernelEx(&cfg, match_topk_u8x2_kernel<256>, tmA8, tmB8, (const float*)invB, b_alloc, (const float*)invA, a_alloc,
                                   (const int*)invb_max, A.count, A.count_stride, A.cap, B.count, B.count_stride, B.cap,
                                   (const int*)ctl, n_splits, cand, cand_stride, n_slots, flt.key_floor, dbg_c, B.cap, clk));
      }
      if (clk_on) {   // debug only: synchronises
        std::vector<long long> h(4 * n_cta);
        VO_CUDA(cudaMemcpyAsync(h.data(), clk, 32 * n_cta, cudaMemcpyDeviceToHost, st));
        VO_CUDA(cudaStreamSynchronize(st));
        double tot = 0, te = 0, bf = 0, tiles = 0; int n = 0;
        for (size_t i = 0; i < n_cta; ++i) if (h[4 * i + 3] > 0) { tot += h[4 * i]; te += h[4 * i + 1]; bf += h[4 * i + 2]; tiles += h[4 * i + 3]; ++n; }
        if (n) fprintf(stderr, "[match clk] %d leader CTAs, %.0f tiles each: %.0f cycles/tile, wait t_empty %.0f, wait b_full %.0f\n", n, tiles / n, tot / tiles, te / tiles, bf / tiles);
      }
    }
    VO_CUDA(cudaGetLastError());
    ctx->match_stats[2] = 1;
  }
  match_finalize_kernel<<<dim3(div_up(A.cap, 128), n_prob), 128, 0, st>>>(
      cand, cand_stride, n_slots, ra, rb, dim, invA, a_alloc, invB, b_alloc, A.count, A.count_stride, A.cap, B.count,
      B.count_stride, B.cap, ctl, (use_pairs || use_ts) ? 1 : 0, sched_L, sched_T, flt, gen_eps, skip_general ? 1 : 0, out->j1, out->s1, out->s2, a_alloc, scan_list, ctl + 1);
  if (B.cap > 0) {
    int scan_grid = ctx->num_sms * 2;
    match_rowscan_kernel<<<scan_grid, 256, 0, st>>>(scan_list, ctl + 1, ra, rb, u8A, u8B, a_alloc, b_alloc, ctl, dim, invA, a_alloc, invB, b_alloc, B.count,
                                                    B.count_stride, B.cap, out->j1, out->s1, out->s2, a_alloc);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

void fill_match_opts(const vo_match_opts* in, vo_match_opts* o) {
  o->match_threshold = 1.0f; o->max_ratio = 0.6f; o->unique = 0; o->index_base = 0;
  if (in) {
    if (in->match_threshold > 0) o->match_threshold = in->match_threshold;
    if (in->max_ratio > 0) o->max_ratio = in->max_ratio;
    o->unique = in->unique != 0;
    o->index_base = in->index_base;
  }
}

int match_batch_select(vo_ctx* ctx, const MatchTop2& t, const MatchOperand& A, const MatchOperand& B, int n_prob,
                       const vo_match_opts& o, uint32_t* idx1, uint32_t* idx2, float* metric, int out_stride,
                       int* n_pairs, int np_stride, cudaStream_t st, const MatchTop2* back) {
  (void)ctx;
  if (n_prob <= 0) return VO_OK;
  const float thr = o.match_threshold * 0.04f;
  ProfScope ps(ctx, st, "match_select");
  match_select_block_kernel<<<n_prob, 1024, 0, st>>>(t.j1, t.s1, t.s2, t.row_stride, A.count, A.count_stride, A.cap, B.count,
                                                     B.count_stride, B.cap, thr, o.max_ratio, o.index_base, idx1, idx2, metric,
                                                     out_stride, n_pairs, np_stride, back ? back->j1 : nullptr,
                                                     back ? back->row_stride : 0);
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// ------------------------------------------------------------- single-problem entry points
struct Single {
  MatchOperand A, B;
  int* counts;   // device {n1, n2}
};

static int single_operands(vo_ctx* ctx, const float* f1, int n1, const float* f2, int n2, int col_major, const char* tag,
                           cudaStream_t st, Single* s) {
  VO_TRY(dev_buf(ctx, (std::string("m_counts_") + tag).c_str(), 8, &s->counts));
  match_set_counts_kernel<<<1, 32, 0, st>>>(s->counts, n1, n2);
  ctx->kernel_launches += 1;
  s->A = MatchOperand(); s->B = MatchOperand();
  s->A.base = f1; s->A.count = s->counts; s->A.cap = n1; s->A.col_major = col_major; s->A.ld = n1;
  s->B.base = f2; s->B.count = s->counts + 1; s->B.cap = n2; s->B.col_major = col_major; s->B.ld = n2;
  return VO_OK;
}

// device-resident single match (inputs on device, row- or col-major; outputs on device)
static int match_device(vo_ctx* ctx, const float* f1, int n1, const float* f2, int n2, int dim, int col_major,
                        const vo_match_opts* opts, uint32_t* idx1, uint32_t* idx2, float* metric, int* n_pairs_dev,
                        cudaStream_t st) {
  vo_match_opts o; fill_match_opts(opts, &o);
  if (n1 == 0 || n2 == 0) {
    VO_CUDA(cudaMemsetAsync(n_pairs_dev, 0, sizeof(int), st));
    return VO_OK;
  }
  Single s; VO_TRY(single_operands(ctx, f1, n1, f2, n2, col_major, "f", st, &s));
  MatchTop2 t;
  const MatchFilter flt = make_match_filter(o);
  VO_TRY(match_batch_top2(ctx, s.A, s.B, 1, dim, "f", nullptr, st, &t, &flt));
  const uint32_t* bj1 = nullptr;
  if (o.unique) {
    Single sb; VO_TRY(single_operands(ctx, f2, n2, f1, n1, col_major, "b", st, &sb));
    MatchTop2 tb;
    VO_TRY(match_batch_top2(ctx, sb.A, sb.B, 1, dim, "b", nullptr, st, &tb, &flt));
    bj1 = tb.j1;
    ctx->match_stats[2] = 2;
  }
  if (!o.unique && n1 <= 16384)
    return match_batch_select(ctx, t, s.A, s.B, 1, o, idx1, idx2, metric, n1, n_pairs_dev, 1, st);
  const int nb = div_up(n1, SEL_BLOCK);
  int* blk; VO_TRY(dev_buf(ctx, "m_blk", (size_t)nb + 1, &blk));
  const float thr = o.match_threshold * 0.04f;
  ctx->kernel_launches += 3;
  match_count_kernel<<<nb, SEL_BLOCK, 0, st>>>(t.j1, t.s1, t.s2, bj1, s.counts, s.counts + 1, thr, o.max_ratio, blk);
  match_scan_kernel<<<1, 1024, 0, st>>>(blk, nb, n_pairs_dev);
  match_scatter_kernel<<<nb, SEL_BLOCK, 0, st>>>(t.j1, t.s1, t.s2, bj1, s.counts, s.counts + 1, thr, o.max_ratio, o.index_base, blk,
                                                 idx1, idx2, metric);
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

static int upload_features(vo_ctx* ctx, const char* name, const float* f, int n, int dim, float** out, cudaStream_t st) {
  VO_TRY(dev_buf(ctx, name, (size_t)(n > 0 ? n : 1) * dim, out));
  if (n > 0) VO_CUDA(cudaMemcpyAsync(*out, f, (size_t)n * dim * sizeof(float), cudaMemcpyHostToDevice, st));
  return VO_OK;
}

static int read_stats(vo_ctx* ctx, cudaStream_t st) {
  int* ctl; VO_TRY(dev_buf(ctx, "m_ctl_f", 8, &ctl));
  int* h; VO_TRY(pin_buf(ctx, "m_ctl", 8, &h));
  VO_CUDA(cudaMemcpyAsync(h, ctl, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  VO_CUDA(cudaStreamSynchronize(st));
  ctx->match_stats[0] = h[0] ? 0 : 1;
  ctx->match_stats[1] = h[1];
  if (h[0]) ctx->match_stats[3] *= ctx->match_float_terms;
  return VO_OK;
}

}  // namespace vo

using namespace vo;

extern "C" {

int vo_match(vo_ctx* ctx, const float* f1, int n1, const float* f2, int n2, int dim, int col_major,
             const vo_match_opts* opts, uint32_t* idx1, uint32_t* idx2, float* metric, int* n_pairs) {
  VO_CHECK_ARG(ctx && n_pairs, "ctx/n_pairs is null");
  VO_CHECK_ARG(n1 >= 0 && n2 >= 0, "negative size");
  VO_CHECK_ARG(dim > 0 && dim <= 128, "dim must be in 1..128 (SIFT descriptors are 128-d)");
  VO_CHECK_ARG((n1 == 0 || f1) && (n2 == 0 || f2), "feature pointer is null");
  VO_CHECK_ARG(n1 == 0 || (idx1 && idx2), "output pointer is null");
  VO_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  *n_pairs = 0;
  if (n1 == 0 || n2 == 0) { ctx->match_stats[0] = 1; ctx->match_stats[1] = 0; ctx->match_stats[2] = 0; return VO_OK; }
  float *d1, *d2;
  VO_TRY(upload_features(ctx, "m_in1", f1, n1, dim, &d1, st));
  VO_TRY(upload_features(ctx, "m_in2", f2, n2, dim, &d2, st));
  uint32_t *o1, *o2; float* om; int* np;
  VO_TRY(dev_buf(ctx, "m_out1", (size_t)n1, &o1));
  VO_TRY(dev_buf(ctx, "m_out2", (size_t)n1, &o2));
  VO_TRY(dev_buf(ctx, "m_outm", (size_t)n1, &om));
  VO_TRY(dev_buf(ctx, "m_np", 4, &np));
  VO_TRY(match_device(ctx, d1, n1, d2, n2, dim, col_major, opts, o1, o2, om, np, st));
  int* hres; VO_TRY(pin_buf(ctx, "m_res", 8, &hres));
  VO_CUDA(cudaMemcpyAsync(hres, np, sizeof(int), cudaMemcpyDeviceToHost, st));
  VO_TRY(read_stats(ctx, st));
  const int p = hres[0];
  if (p > 0) {
    VO_CUDA(cudaMemcpyAsync(idx1, o1, (size_t)p * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    VO_CUDA(cudaMemcpyAsync(idx2, o2, (size_t)p * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (metric) VO_CUDA(cudaMemcpyAsync(metric, om, (size_t)p * sizeof(float), cudaMemcpyDeviceToHost, st));
    VO_CUDA(cudaStreamSynchronize(st));
  }
  *n_pairs = p;
  return VO_OK;
}

int vo_match_dev(vo_ctx* ctx, const float* f1_dev, int n1, const float* f2_dev, int n2, int dim,
                 const vo_match_opts* opts, uint32_t* idx1_dev, uint32_t* idx2_dev, float* metric_dev,
                 int* n_pairs_dev, void* stream) {
  VO_CHECK_ARG(ctx && n_pairs_dev, "ctx/n_pairs_dev is null");
  VO_CHECK_ARG(n1 >= 0 && n2 >= 0, "negative size");
  VO_CHECK_ARG(dim > 0 && dim <= 128, "dim must be in 1..128 (SIFT descriptors are 128-d)");
  VO_CUDA(cudaSetDevice(ctx->device));
  return match_device(ctx, f1_dev, n1, f2_dev, n2, dim, 0, opts, idx1_dev, idx2_dev, metric_dev, n_pairs_dev,
                      (cudaStream_t)stream);
}

int vo_match_top2_dev(vo_ctx* ctx, const float* f1_dev, int n1, const float* f2_dev, int n2, int dim,
                      uint32_t* j1_dev, float* s1_dev, float* s2_dev, void* stream) {
  VO_CHECK_ARG(ctx, "ctx is null");
  VO_CHECK_ARG(n1 >= 0 && n2 >= 0, "negative size");
  VO_CHECK_ARG(dim > 0 && dim <= 128, "dim must be in 1..128 (SIFT descriptors are 128-d)");
  VO_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (n1 == 0) return VO_OK;
  Single s; VO_TRY(single_operands(ctx, f1_dev, n1, f2_dev, n2, 0, "f", st, &s));
  MatchTop2 t;
  VO_TRY(match_batch_top2(ctx, s.A, s.B, 1, dim, "f", nullptr, st, &t));
  VO_CUDA(cudaMemcpyAsync(j1_dev, t.j1, (size_t)n1 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  VO_CUDA(cudaMemcpyAsync(s1_dev, t.s1, (size_t)n1 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  VO_CUDA(cudaMemcpyAsync(s2_dev, t.s2, (size_t)n1 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return VO_OK;
}

int vo_match_best2_dev(vo_ctx* ctx, const float* f1_dev, int n1, const float* f2_dev, int n2, int dim,
                       const vo_match_opts* opts, void* records_dev, void* stream) {
  VO_CHECK_ARG(ctx && records_dev, "ctx/records_dev is null");
  VO_CHECK_ARG(n1 >= 0 && n2 >= 0, "negative size");
  VO_CHECK_ARG(dim > 0 && dim <= 128, "dim must be in 1..128 (SIFT descriptors are 128-d)");
  VO_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (n1 == 0) return VO_OK;
  vo_match_opts o; fill_match_opts(opts, &o);
  const bool prepared = f2_dev == nullptr && n2 > 0;    // an empty landmark set has a null pointer too
  VO_CHECK_ARG(!prepared || (ctx->landmarks_prepared == n2 && dim == 128), "f2_dev is null but no landmark set of this size was prepared (vo_landmarks_prepare)");
  const char* tag = prepared ? "lm" : "f";
  Single s; VO_TRY(single_operands(ctx, f1_dev, n1, f2_dev, n2, 0, tag, st, &s));
  s.B.prepared = prepared ? 1 : 0;
  MatchTop2 t;
  const MatchFilter flt = make_match_filter(o);
  VO_TRY(match_batch_top2(ctx, s.A, s.B, 1, dim, tag, nullptr, st, &t, &flt));
  match_records_kernel<<<div_up(n1, 256), 256, 0, st>>>(t.j1, t.s1, t.s2, n1, n2, o.match_threshold * 0.04f, o.max_ratio,
                                                        o.index_base, static_cast<uint4*>(records_dev));
  ctx->kernel_launches += 1;
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

int vo_landmarks_prepare(vo_ctx* ctx, const float* f2_dev, int n2, int dim, void* stream) {
  VO_CHECK_ARG(ctx && f2_dev && n2 > 0, "null or empty landmark set");
  VO_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  Single s; VO_TRY(single_operands(ctx, f2_dev, 0, f2_dev, n2, 0, "lm", st, &s));
  VO_TRY(match_prepare_landmarks(ctx, s.B, dim, "lm", st));
  ctx->landmarks_prepared = n2;
  return VO_OK;
}

int vo_match_best2_gather_dev(vo_ctx* ctx, const float* f1_dev, int n1, const float* f2_dev, int n2, int dim,
                              const vo_match_opts* opts, void* const* peer_bufs, int n_peers, size_t first_row, void* stream) {
  VO_CHECK_ARG(ctx && peer_bufs, "ctx/peer_bufs is null");
  VO_CHECK_ARG(n1 >= 0 && n2 >= 0, "negative size");
  VO_CHECK_ARG(n_peers >= 1 && n_peers <= 16, "n_peers must be in 1..16");
  VO_CHECK_ARG(dim > 0 && dim <= 128, "dim must be in 1..128 (SIFT descriptors are 128-d)");
  VO_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (n1 == 0) return VO_OK;
  vo_match_opts o; fill_match_opts(opts, &o);
  const bool prepared = f2_dev == nullptr && n2 > 0;    // an empty landmark set has a null pointer too
  VO_CHECK_ARG(!prepared || (ctx->landmarks_prepared == n2 && dim == 128), "f2_dev is null but no landmark set of this size was prepared (vo_landmarks_prepare)");
  const char* tag = prepared ? "lm" : "f";
  Single s; VO_TRY(single_operands(ctx, f1_dev, n1, f2_dev, n2, 0, tag, st, &s));
  s.B.prepared = prepared ? 1 : 0;
  MatchTop2 t;
  const MatchFilter flt = make_match_filter(o);
  VO_TRY(match_batch_top2(ctx, s.A, s.B, 1, dim, tag, nullptr, st, &t, &flt));
  PeerTable pt; pt.n = n_peers;
  for (int r = 0; r < 16; ++r) pt.p[r] = r < n_peers ? static_cast<uint4*>(peer_bufs[r]) : nullptr;
  for (int r = 0; r < n_peers; ++r) VO_CHECK_ARG(pt.p[r] != nullptr, "a peer buffer is null");
  match_records_gather_kernel<<<div_up(n1, 256), 256, 0, st>>>(t.j1, t.s1, t.s2, n1, n2, o.match_threshold * 0.04f, o.max_ratio,
                                                               o.index_base, pt, first_row);
  ctx->kernel_launches += 1;
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// Buffers that other processes (one per GPU) can write into: cudaMalloc + CUDA IPC.
int vo_peer_alloc(vo_ctx* ctx, size_t bytes, void** dev_ptr, uint8_t handle[64]) {
  VO_CHECK_ARG(ctx && dev_ptr && handle && bytes > 0, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  VO_CUDA(cudaSetDevice(ctx->device));
  VO_CUDA(cudaMalloc(dev_ptr, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *dev_ptr);
  if (e != cudaSuccess) { cudaFree(*dev_ptr); *dev_ptr = nullptr; set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return VO_ERR_CUDA; }
  memcpy(handle, &h, 64);
  return VO_OK;
}
int vo_peer_open(vo_ctx* ctx, const uint8_t handle[64], void** dev_ptr) {
  VO_CHECK_ARG(ctx && dev_ptr && handle, "bad argument");
  VO_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h; memcpy(&h, handle, 64);
  VO_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return VO_OK;
}
int vo_peer_close(vo_ctx* ctx, void* dev_ptr) {
  VO_CHECK_ARG(ctx && dev_ptr, "bad argument");
  VO_CUDA(cudaSetDevice(ctx->device));
  VO_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return VO_OK;
}
int vo_peer_free(vo_ctx* ctx, void* dev_ptr) {
  VO_CHECK_ARG(ctx && dev_ptr, "bad argument");
  VO_CUDA(cudaSetDevice(ctx->device));
  VO_CUDA(cudaFree(dev_ptr));
  return VO_OK;
}

int vo_match_top2(vo_ctx* ctx, const float* f1, int n1, const float* f2, int n2, int dim, int col_major,
                  uint32_t* j1_out, float* s1_out, float* s2_out) {
  VO_CHECK_ARG(ctx, "ctx is null");
  VO_CHECK_ARG(n1 >= 0 && n2 >= 0, "negative size");
  VO_CHECK_ARG(dim > 0 && dim <= 128, "dim must be in 1..128 (SIFT descriptors are 128-d)");
  VO_CHECK_ARG((n1 == 0 || f1) && (n2 == 0 || f2), "feature pointer is null");
  VO_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (n1 == 0) return VO_OK;
  float *d1, *d2;
  VO_TRY(upload_features(ctx, "m_in1", f1, n1, dim, &d1, st));
  VO_TRY(upload_features(ctx, "m_in2", f2, n2, dim, &d2, st));
  Single s; VO_TRY(single_operands(ctx, d1, n1, d2, n2, col_major, "f", st, &s));
  MatchTop2 t;
  VO_TRY(match_batch_top2(ctx, s.A, s.B, 1, dim, "f", nullptr, st, &t));
  VO_CUDA(cudaMemcpyAsync(j1_out, t.j1, (size_t)n1 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  VO_CUDA(cudaMemcpyAsync(s1_out, t.s1, (size_t)n1 * sizeof(float), cudaMemcpyDeviceToHost, st));
  VO_CUDA(cudaMemcpyAsync(s2_out, t.s2, (size_t)n1 * sizeof(float), cudaMemcpyDeviceToHost, st));
  VO_TRY(read_stats(ctx, st));
  return VO_OK;
}

int vo_match_stats(vo_ctx* ctx, int stats[4]) {
  VO_CHECK_ARG(ctx && stats, "null argument");
  for (int i = 0; i < 4; ++i) stats[i] = ctx->match_stats[i];
  return VO_OK;
}

int vo_match_debug_gemm(vo_ctx* ctx, const float* f1, int n1, const float* f2, int n2, int dim, float* c_out) {
  VO_CHECK_ARG(ctx && f1 && f2 && c_out, "null argument");
  VO_CHECK_ARG(n1 > 0 && n2 > 0, "empty input");
  VO_CHECK_ARG(dim > 0 && dim <= 128, "dim must be in 1..128 (SIFT descriptors are 128-d)");
  VO_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  float *d1, *d2, *dc;
  VO_TRY(upload_features(ctx, "m_in1", f1, n1, dim, &d1, st));
  VO_TRY(upload_features(ctx, "m_in2", f2, n2, dim, &d2, st));
  VO_TRY(dev_buf(ctx, "m_dbgc", (size_t)n1 * n2, &dc));
  VO_CUDA(cudaMemsetAsync(dc, 0xFF, (size_t)n1 * n2 * sizeof(float), st));
  Single s; VO_TRY(single_operands(ctx, d1, n1, d2, n2, 0, "f", st, &s));
  MatchTop2 t;
  VO_TRY(match_batch_top2(ctx, s.A, s.B, 1, dim, "f", dc, st, &t));
  VO_CUDA(cudaMemcpyAsync(c_out, dc, (size_t)n1 * n2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  VO_CUDA(cudaStreamSynchronize(st));
  return VO_OK;
}

}  // extern "C"
