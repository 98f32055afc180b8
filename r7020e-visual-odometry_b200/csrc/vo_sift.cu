// vo_sift.cu -- detectSIFTFeatures + extractFeatures("Method","SIFT") on B200 (VO.m:79-84).
//
// Batch-first: every kernel covers one (octave, layer) of ALL images of the batch, so a 1241x376
// frame (far too small to fill 148 SMs alone) is processed B images at a time.
//
// HBM layout (per plan): gauss[o][l][b][h_o][pitch_o] and dog[o][l][b][h_o][pitch_o], fp32, pitch
// rounded up to 32 floats so rows are 128-byte aligned (vector loads/stores, TMA-legal strides).
//
//   sift_base_kernel      u8 -> 2x bilinear upsample (cv::resize convention) -> blur(sig_diff)
//   sift_blur_dog_kernel  G[l] -> G[l+1] (separable Gaussian through shared memory, reflect-101)
//                         fused with D[l] = G[l+1] - G[l]; each layer is read once, written once
//   sift_downsample_kernel  G[o+1][0] = G[o][nl](2y, 2x)
//   sift_extrema_kernel   3x3x3 DoG extrema, warp-ballot compaction into a candidate list
//   sift_refine_orient_kernel  warp per candidate: quadratic sub-pixel fit, contrast/edge tests,
//                         36-bin orientation histogram (shared-memory integer atomics), peaks
//   sift_rank_kernel / sift_dedupe_kernel  OpenCV keypoint order (x, y, -size, angle, ...) by
//                         rank sort, duplicate removal, compaction
//   sift_descriptor_kernel  warp per keypoint: 4x4x8 trilinear histogram (shared-memory integer
//                         atomics), clip 0.2, scale 512, round to 0..255
//
// Arithmetic follows the contract in oracle/sift.c (DESIGN.md "SIFT arithmetic"): explicit fmaf,
// no implicit contraction (-fmad=false), polynomial exp/atan2, histogram contributions rounded to
// 1/4096 and summed as integers, so results do not depend on thread scheduling.
#include "vo_internal.h"
#include "vo_ptx.cuh"
#include <cfloat>
#include <cmath>

namespace vo {

constexpr int MAX_OCT = 16;
constexpr int MAX_R = 32;   // generic-radius kernels: NumLayersInOctave >= 2 at Sigma 1.6 needs r = 18
constexpr int SIFT_BORDER = 5;
constexpr int ORI_BINS = 36;
constexpr float SIFT_FIX = 4096.0f;
constexpr float SIFT_INV_FIX = 1.0f / 4096.0f;
constexpr int TILE_W = 128, TILE_H = 32;

struct Taps { float k[MAX_R + 1]; int r; };
struct OctInfo { int h[MAX_OCT], w[MAX_OCT], pitch[MAX_OCT]; size_t goff[MAX_OCT], doff[MAX_OCT]; };

struct RefinedKp;
struct SiftPlan {
  int rows = 0, cols = 0, batch = 0, nl = 0;
  float sigma = 0, contrast = 0, edge = 0;
  int n_oct = 0;
  int h[MAX_OCT], w[MAX_OCT], pitch[MAX_OCT];
  size_t goff[MAX_OCT], doff[MAX_OCT];   // float offsets of octave o inside gauss/dog
  float* gauss = nullptr; float* dog = nullptr;
  uint8_t* img = nullptr;       // [batch][rows][cols]
  uint8_t* img_t = nullptr;     // staging for column-major input
  Taps taps[8]; Taps base_taps;
  CUtensorMap tm_gauss[MAX_OCT];   // {w, h, layer*batch} view of every octave's Gaussian stack (blur input via TMA)
  CUtensorMap tm_dog[MAX_OCT];     // the same view of the DoG stack, box {256, 4, 1} (extrema input via TMA)
  int cand_cap = 0, kp_cap = 0;
  uint32_t* cand = nullptr; int* counters = nullptr;   // counters[b*4 + {0:cand,1:raw kp,2:final}]
  int* work = nullptr;                                  // work[b*2 + {0: next refined candidate, 1: next keypoint}] (dynamic scheduling)
  vo_keypoint* raw = nullptr; vo_keypoint* sorted = nullptr; vo_keypoint* final_kp = nullptr;
  float* desc = nullptr; float2* trig = nullptr; struct RefinedKp* refined = nullptr;
  size_t layer_elems(int o) const { return (size_t)batch * h[o] * pitch[o]; }
  float* G(int o, int l) const { return gauss + goff[o] + (size_t)l * layer_elems(o); }
  float* D(int o, int l) const { return dog + doff[o] + (size_t)l * layer_elems(o); }
};

void sift_plan_destroy(SiftPlan* p) {
  if (!p) return;
  cudaFree(p->gauss); cudaFree(p->dog); cudaFree(p->img); cudaFree(p->img_t); cudaFree(p->cand);
  cudaFree(p->counters); cudaFree(p->work); cudaFree(p->raw); cudaFree(p->sorted); cudaFree(p->final_kp); cudaFree(p->desc); cudaFree(p->trig); cudaFree(p->refined);
  delete p;
}

// ----------------------------------------------------------------------------- primitives
__device__ __forceinline__ float vo_expf(float x) {
  const float t = x * 1.44269504088896341f;
  float n = rintf(t);
  const bool under = n < -126.f;     // result is 0; evaluated branch-free with n clamped
  if (n > 127.f) n = 127.f;
  if (under) n = -126.f;
  float r = fmaf(n, -0.693145751953125f, x);
  r = fmaf(n, -1.42860682030941723e-6f, r);
  float p = 1.0f / 720.0f;
  p = fmaf(p, r, 1.0f / 120.0f);
  p = fmaf(p, r, 1.0f / 24.0f);
  p = fmaf(p, r, 1.0f / 6.0f);
  p = fmaf(p, r, 0.5f);
  p = fmaf(p, r, 1.0f);
  p = fmaf(p, r, 1.0f);
  return under ? 0.f : __int_as_float(__float_as_int(p) + (((int)n) << 23));
}

// vo_expf for arguments whose rounded exponent can neither overflow nor underflow (here: Gaussian
// window weights, x in [-40, 0]): the same operations without the range handling, same bits.
__device__ __forceinline__ float vo_expf_window(float x) {
  const float n = rintf(x * 1.44269504088896341f);
  float r = fmaf(n, -0.693145751953125f, x);
  r = fmaf(n, -1.42860682030941723e-6f, r);
  float p = 1.0f / 720.0f;
  p = fmaf(p, r, 1.0f / 120.0f);
  p = fmaf(p, r, 1.0f / 24.0f);
  p = fmaf(p, r, 1.0f / 6.0f);
  p = fmaf(p, r, 0.5f);
  p = fmaf(p, r, 1.0f);
  p = fmaf(p, r, 1.0f);
  return __int_as_float(__float_as_int(p) + (((int)n) << 23));
}

// (uint32_t)__float2int_rn(v * 4096) for 0 <= v * 4096 < 2^22 without the conversion pipe: v * 4096 is
// exact, so one fused multiply-add onto 1.5 * 2^23 rounds it to nearest-even into the mantissa.
__device__ __forceinline__ uint32_t sift_fix(float v) {
  return (uint32_t)__float_as_int(fmaf(v, SIFT_FIX, 12582912.f)) & 0x3FFFFFu;
}

__device__ __forceinline__ float vo_atan2deg(float y, float x) {
  const float p1 = 0.9997878412794807f * 57.29577951308232f;
  const float p3 = -0.3258083974640975f * 57.29577951308232f;
  const float p5 = 0.1555786518463281f * 57.29577951308232f;
  const float p7 = -0.04432655554792128f * 57.29577951308232f;
  const float ax = fabsf(x), ay = fabsf(y);
  // branch-free form of: ax >= ay ? poly(ay / (ax + eps)) : 90 - poly(ax / (ay + eps)) (same operations)
  const bool xbig = ax >= ay;
  const float c = __fdiv_rn(xbig ? ay : ax, (xbig ? ax : ay) + 2.220446049250313e-16f);
  const float c2 = c * c;
  const float pa = fmaf(fmaf(fmaf(p7, c2, p5), c2, p3), c2, p1) * c;
  float a = xbig ? pa : 90.f - pa;
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  }
  return p;
}

// value of the 2x-upsampled image (cv::resize INTER_LINEAR convention) at (x, y)
__device__ __forceinline__ float dbl_at(const uint8_t* __restrict__ img, int rows, int cols, int x, int y) {
  int ya = (y & 1) ? (y >> 1) : (y >> 1) - 1, yb = ya + 1;
  const float wyb = (y & 1) ? 0.25f : 0.75f, wya = 1.0f - wyb;
  if (ya < 0) ya = 0;
  if (yb > rows - 1) yb = rows - 1;
  int xa = (x & 1) ? (x >> 1) : (x >> 1) - 1, xb = xa + 1;
  const float wxb = (x & 1) ? 0.25f : 0.75f, wxa = 1.0f - wxb;
  if (xa < 0) xa = 0;
  if (xb > cols - 1) xb = cols - 1;
  const float a = img[(size_t)ya * cols + xa], b = img[(size_t)ya * cols + xb];
  const float c = img[(size_t)yb * cols + xa], d = img[(size_t)yb * cols + xb];
  return wya * (wxa * a + wxb * b) + wyb * (wxa * c + wxb * d);
}

__global__ void sift_transpose_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int rows, int cols, int ld,
                                         size_t src_step, size_t dst_step) {
  __shared__ uint8_t t[32][33];
  src += blockIdx.z * src_step; dst += blockIdx.z * dst_step;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;   // bx over cols, by over rows
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {    // read src (col-major): element (r,c) at src[c*ld + r]
    const int c = bx + i, r = by + threadIdx.x;
    if (c < cols && r < rows) t[i][threadIdx.x] = src[(size_t)c * ld + r];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = by + i, c = bx + threadIdx.x;
    if (r < rows && c < cols) dst[(size_t)r * cols + c] = t[threadIdx.x][i];
  }
}

// ----------------------------------------------------------------- separable blur (+ DoG)
// One block = TILE_W x TILE_H outputs of one image.  RT > 0: compile-time radius (fully unrolled),
// RT == 0: runtime radius taps.r (generic options).  FROM_U8: the source is the uint8 input and the
// tile is filled with the 2x-upsampled values on the fly (base image), no DoG.
template <int RT, bool FROM_U8>
__global__ void __launch_bounds__(256)
sift_blur_dog_kernel(const float* __restrict__ src, const uint8_t* __restrict__ src8, float* __restrict__ dst,
                     float* __restrict__ dog, int h, int w, int pitch, int rows8, int cols8, const Taps taps) {
  extern __shared__ float sm[];
  const int R = RT > 0 ? RT : taps.r;
  const int in_w = ((TILE_W + 2 * R + 3) & ~3) + 4;   // padded row length of the input tile
  const int in_h = TILE_H + 2 * R;
  float* s_in = sm;                       // [in_h][in_w]
  float* s_tmp = sm + in_h * in_w;        // [in_h][TILE_W]
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
  const size_t img_off = (size_t)b * h * pitch;
  const int tid = threadIdx.x;

  // ---- load tile + halo (reflect-101 at the image border)
  const int load_w = TILE_W + 2 * R;
  for (int idx = tid; idx < in_h * load_w; idx += 256) {
    const int iy = idx / load_w, ix = idx - iy * load_w;
    const int gy = reflect101(y0 - R + iy, h), gx = reflect101(x0 - R + ix, w);
    float v;
    if (FROM_U8) v = dbl_at(src8 + (size_t)b * rows8 * cols8, rows8, cols8, gx, gy);
    else v = __ldg(src + img_off + (size_t)gy * pitch + gx);
    s_in[iy * in_w + ix] = v;
  }
  __syncthreads();

  // ---- row pass: warp per row, 4 consecutive outputs per lane
  {
    const int lane = tid & 31, wrp = tid >> 5;
    for (int row = wrp; row < in_h; row += 8) {
      const float* p = s_in + row * in_w + 4 * lane;
      float win[4 + 2 * (RT > 0 ? RT : MAX_R)];
      const int nwin = 4 + 2 * R;
#pragma unroll
      for (int q = 0; q < (4 + 2 * (RT > 0 ? RT : MAX_R) + 3) / 4; ++q) {
        if (4 * q < nwin) {
          const float4 v4 = *reinterpret_cast<const float4*>(p + 4 * q);
          win[4 * q] = v4.x;
          if (4 * q + 1 < 4 + 2 * (RT > 0 ? RT : MAX_R)) win[4 * q + 1] = v4.y;
          if (4 * q + 2 < 4 + 2 * (RT > 0 ? RT : MAX_R)) win[4 * q + 2] = v4.z;
          if (4 * q + 3 < 4 + 2 * (RT > 0 ? RT : MAX_R)) win[4 * q + 3] = v4.w;
        }
      }
      float acc[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = taps.k[0] * win[R + o];
      if (RT > 0) {
#pragma unroll
        for (int i = 1; i <= (RT > 0 ? RT : 1); ++i) {
#pragma unroll
          for (int o = 0; o < 4; ++o) acc[o] = fmaf(taps.k[i], win[RT + o - i] + win[RT + o + i], acc[o]);
        }
      } else {
        for (int i = 1; i <= R; ++i)
          for (int o = 0; o < 4; ++o) acc[o] = fmaf(taps.k[i], p[R + o - i] + p[R + o + i], acc[o]);
      }
      *reinterpret_cast<float4*>(s_tmp + row * TILE_W + 4 * lane) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
  }
  __syncthreads();

  // ---- column pass: thread = (column, 16-row half), 4 outputs at a time
  {
    const int x = tid & (TILE_W - 1), half = tid >> 7;
    const int gx = x0 + x;
#pragma unroll 1
    for (int chunk = 0; chunk < 4; ++chunk) {
      const int ty = half * 16 + chunk * 4;     // first output row of this chunk inside the tile
      if (y0 + ty >= h) break;
      const float* p = s_tmp + ty * TILE_W + x;   // window row 0 <-> tile row ty - R  (+R offset in s_tmp)
      float acc[4];
      if (RT > 0) {
        float win[4 + 2 * (RT > 0 ? RT : 1)];
#pragma unroll
        for (int q = 0; q < 4 + 2 * (RT > 0 ? RT : 1); ++q) win[q] = p[q * TILE_W];
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[o] = taps.k[0] * win[RT + o];
#pragma unroll
        for (int i = 1; i <= (RT > 0 ? RT : 1); ++i) {
#pragma unroll
          for (int o = 0; o < 4; ++o) acc[o] = fmaf(taps.k[i], win[RT + o - i] + win[RT + o + i], acc[o]);
        }
      } else {
        for (int o = 0; o < 4; ++o) acc[o] = taps.k[0] * p[(R + o) * TILE_W];
        for (int i = 1; i <= R; ++i)
          for (int o = 0; o < 4; ++o) acc[o] = fmaf(taps.k[i], p[(R + o - i) * TILE_W] + p[(R + o + i) * TILE_W], acc[o]);
      }
      if (gx < w) {
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int gy = y0 + ty + o;
          if (gy < h) {
            const size_t g = img_off + (size_t)gy * pitch + gx;
            dst[g] = acc[o];
            if (!FROM_U8) dog[g] = acc[o] - s_in[(R + ty + o) * in_w + R + x];
          }
        }
      }
    }
  }
}

// ------------------------------------------------------ streaming separable blur (+ DoG)
// The large octaves use a streaming formulation instead of square tiles: a block owns a strip of
// ST_W columns and walks down the image ST_CH rows at a time.  Row-filtered rows live in a ring
// buffer in shared memory (ST_CH + 2R rows), so no row is ever filtered twice (no vertical halo
// recompute) and every input row is read once from HBM (neighbouring lanes' overlapping windows
// hit L1).  Row pass: warp per row, 16-byte vector loads, 4 outputs per lane.  Column pass: thread
// per column PAIR, packed f32x2 math (FFMA2/FADD2: two columns per instruction), fused
// D[l] = G[l+1] - G[l].  Each f32x2 lane is an IEEE fmaf, so results equal the scalar contract.
typedef unsigned long long u64;
__device__ __forceinline__ u64 f2_fma(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 f2_add(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 f2_mul(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 f2_bcast(float k) { const float2 v = make_float2(k, k); return *reinterpret_cast<const u64*>(&v); }

// TMA-staged variant (the one used for the large octaves).  Input rows arrive in 16-row groups through
// cp.async.bulk.tensor.3d (tensor map {w, h, layer*batch}; out-of-image columns are zero-filled and
// then patched with the reflect-101 values, which are already inside the box) into a double-buffered
// stage; the load of group g+2 is issued right after group g has been row-filtered, i.e. a whole
// chunk of compute ahead of its use, so no warp waits on HBM latency.  The ring holds 4 groups
// (64 rows): the column pass of chunk c reads groups c..c+2 while group c+3 is being written, which
// leaves a single __syncthreads per chunk.
// Box rows hold 164 floats (x0 - 16 .. x0 + 147) and ring rows 132: both strides are 4 (mod 32) banks, so eight lanes
// that read or write 16 bytes in eight different rows / row-column combinations touch 32 different banks (MODE 2).
// next octave's seed, written by the blur that produces layer nl: G[o+1][0](y, x) = G[o][nl](2y, 2x)
struct DsOut { float* dst; int pitch, h, w; };
constexpr int TS_W = 128, TS_G = 16, TS_HALO = 16, TS_BOXW = TS_W + 2 * TS_HALO + 4, TS_RS = TS_W + 4;
__host__ __device__ constexpr int TS_MIRROR(int r) { return 3 + 2 * r; }
// PACK: the row pass runs on packed f32x2 too.  A lane's four outputs are the pairs (c, c+1) and (c+2, c+3); the tap
// pairs (in[c+d], in[c+d+1]) are 8-byte aligned in the stage for even d only, so each warp first writes a copy of its
// row shifted by one column (three shared-memory instructions per lane), where the odd-d pairs are aligned.  (A second
// TMA box from x0 - 15 cannot do that: a tiled TMA load whose innermost start is not 16-byte aligned faults.)  2R + 1
// packed operations per output pair instead of 2(2R + 1) scalar ones; each f32x2 lane is an IEEE fmaf / add, so the
// bits do not change.
// MODE 2: a lane computes EIGHT consecutive outputs of one row, and a warp covers two rows (lane = segment * 2 + row):
// 8 + 2R taps are loaded for 8 outputs instead of 4 + 2R for 4, i.e. 20 instead of 36 bytes of shared memory per pixel at
// R = 13 -- the kernel is bound by shared-memory traffic, not by its arithmetic.
template <int R, int MODE, bool DS>
__global__ void __launch_bounds__(256, MODE == 1 ? 2 : 3)
sift_blur_tma_kernel(const __grid_constant__ CUtensorMap tm, int z_base, float* __restrict__ dst,
                     float* __restrict__ dog, const float* __restrict__ src, int h, int w, int pitch,
                     int seg_rows, const Taps taps, const DsOut ds) {
  static_assert(R <= TS_HALO, "halo too small");
  // ring slots 0..MIRROR-1 are kept a second time at RING + slot, so the 4 + 2R consecutive rows a
  // column window reads are always contiguous in shared memory (one base address, immediate offsets)
  constexpr int RING = 64, MIRROR = TS_MIRROR(R) + (MODE == 3 ? 4 : 0);   // MODE 3 reads 8 + 2R consecutive rows
  extern __shared__ __align__(128) uint8_t tsm[];      // > 48 KB: dynamic shared memory (opt-in)
  constexpr bool PACK = MODE == 1;
  constexpr int SROW_BYTES = PACK ? 8 * TS_BOXW * 4 : 0;   // PACK: one shifted row per warp
  float (*stage)[TS_G][TS_BOXW] = reinterpret_cast<float (*)[TS_G][TS_BOXW]>(tsm);                        // [2]
  float (*srow)[TS_BOXW] = reinterpret_cast<float (*)[TS_BOXW]>(tsm + 2 * TS_G * TS_BOXW * 4);            // [8]
  float (*ring)[TS_RS] = reinterpret_cast<float (*)[TS_RS]>(tsm + 2 * TS_G * TS_BOXW * 4 + SROW_BYTES);    // [RING + MIRROR]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(tsm + 2 * TS_G * TS_BOXW * 4 + SROW_BYTES + (RING + MIRROR) * TS_RS * 4);
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * TS_W;
  const int ys = blockIdx.y * seg_rows, ye = min(ys + seg_rows, h);
  const size_t img_off = (size_t)b * h * pitch;
  const float* img = src + img_off;
  float* dst_i = dst + img_off;
  float* dog_i = dog + img_off;
  // the three image pointers stay in registers as 64-bit values; pixel addresses are then one wide
  // multiply-add of an unsigned 32-bit element index each
  asm volatile("" : "+l"(img), "+l"(dst_i), "+l"(dog_i));
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const bool edge = (x0 - R < 0) || (x0 + TS_W + R > w);
  const uint32_t bar0 = smem_u32(&bars[0]);
  const uint32_t stage0 = smem_u32(&stage[0][0][0]);
  constexpr uint32_t STAGE_BYTES = TS_G * TS_BOXW * 4;
  if (tid == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // group g = rows [ys + 16(g-1), ys + 16g); chunk c needs groups c .. c+2
  const int n_chunks = (ye - ys + TS_G - 1) / TS_G;
  const int n_groups = n_chunks + 2;
  auto issue = [&](int g) {   // one thread
    const uint32_t bar = bar0 + 8 * (g & 1);
    mbar_expect_tx(bar, STAGE_BYTES);
    tma_load_3d(stage0 + (g & 1) * STAGE_BYTES, &tm, bar, x0 - TS_HALO, ys + TS_G * (g - 1), z_base + b);
  };
  if (tid == 0) { issue(0); if (n_groups > 1) issue(1); }

  auto row_pass_group = [&](int g) {
    mbar_wait(bar0 + 8 * (g & 1), (uint32_t)((g >> 1) & 1));
    float (*sg)[TS_BOXW] = stage[g & 1];
    const int row0 = ys + TS_G * (g - 1);
    if (edge) {   // block-uniform: patch the zero-filled out-of-image halo columns with reflect-101 values
      for (int idx = tid; idx < TS_G * R; idx += 256) {
        const int r = idx / R, k = idx - r * R;
        if (x0 == 0) sg[r][TS_HALO - 1 - k] = sg[r][TS_HALO + 1 + k];          // x = -1-k  <-  x = 1+k
        if (x0 + TS_W + R > w) {
          const int cr = w + k - x0 + TS_HALO, sr = w - 2 - k - x0 + TS_HALO;  // x = w+k   <-  x = w-2-k
          if (cr < TS_BOXW && sr >= 0) sg[r][cr] = sg[r][sr];
        }
      }
      __syncthreads();
    }
    if constexpr (MODE >= 2) {
      const int seg = lane >> 1, rr = 2 * wrp + (lane & 1);
      const int row = row0 + rr;
      if (!(row < 0 || row >= h || row < ys - R)) {
        const float* p = &sg[rr][8 * seg];                         // box col 0 <-> x0 - 16
        constexpr int Q0 = (TS_HALO - R) / 4, Q1 = (TS_HALO + 8 + R + 3) / 4;   // float4s that hold needed taps
        float win[4 * (Q1 - Q0)];
#pragma unroll
        for (int q = Q0; q < Q1; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
          win[4 * (q - Q0)] = v.x; win[4 * (q - Q0) + 1] = v.y; win[4 * (q - Q0) + 2] = v.z; win[4 * (q - Q0) + 3] = v.w;
        }
        constexpr int C = TS_HALO - 4 * Q0;                        // index of output 0's centre tap in win
        float acc[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] = taps.k[0] * win[C + o];
#pragma unroll
        for (int i = 1; i <= R; ++i) {
#pragma unroll
          for (int o = 0; o < 8; ++o) acc[o] = fmaf(taps.k[i], win[C + o - i] + win[C + o + i], acc[o]);
        }
        const int slot = row & (RING - 1);
        const float4 oa = make_float4(acc[0], acc[1], acc[2], acc[3]), ob = make_float4(acc[4], acc[5], acc[6], acc[7]);
        *reinterpret_cast<float4*>(&ring[slot][8 * seg]) = oa;
        *reinterpret_cast<float4*>(&ring[slot][8 * seg + 4]) = ob;
        if (slot < MIRROR) {
          *reinterpret_cast<float4*>(&ring[RING + slot][8 * seg]) = oa;
          *reinterpret_cast<float4*>(&ring[RING + slot][8 * seg + 4]) = ob;
        }
      }
    } else {
    for (int rr = wrp; rr < TS_G; rr += 8) {
      const int row = row0 + rr;
      if (row < 0 || row >= h || row < ys - R) continue;        // warp-uniform
      const int slot = row & (RING - 1);
      if constexpr (PACK) {
        // tap pair T(d) = (in[c+d], in[c+d+1]), c = 4*lane: even d from the plain stage, odd d from the shifted one
        __syncwarp();                                            // the previous row's reads of srow are done
#pragma unroll
        for (int q = lane; q < TS_BOXW / 4; q += 32) {           // srow[x] = row[x + 1]
          const float4 v = *reinterpret_cast<const float4*>(&sg[rr][4 * q]);
          const float e = 4 * q + 4 < TS_BOXW ? sg[rr][4 * q + 4] : 0.f;
          *reinterpret_cast<float4*>(&srow[wrp][4 * q]) = make_float4(v.y, v.z, v.w, e);
        }
        __syncwarp();
        const float* pe = &sg[rr][TS_HALO + 4 * lane];
        const float* po = &srow[wrp][TS_HALO + 4 * lane];
        constexpr int QE0 = -((R - (R & 1) + 3) / 4), QE1 = (R + 2) / 4;          // float4 at 4q holds T(4q), T(4q+2):   d in [-R, R+2]
        constexpr int QO0 = -((R - 1 + (R & 1) + 4) / 4), QO1 = (R + 1) / 4;      // float4 at 4q holds T(4q+1), T(4q+3)
        u64 te[2 * (QE1 - QE0 + 1)], to[2 * (QO1 - QO0 + 1)];
#pragma unroll
        for (int q = QE0; q <= QE1; ++q) {
          const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(pe + 4 * q);
          te[2 * (q - QE0)] = v.x; te[2 * (q - QE0) + 1] = v.y;
        }
#pragma unroll
        for (int q = QO0; q <= QO1; ++q) {
          const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(po + 4 * q);
          to[2 * (q - QO0)] = v.x; to[2 * (q - QO0) + 1] = v.y;
        }
        // T(d): d even -> te[(d - 4*QE0) / 2], d odd -> to[(d - 1 - 4*QO0) / 2]
        auto T = [&](int d) -> u64 { return (d & 1) ? to[(d - 1 - 4 * QO0) / 2] : te[(d - 4 * QE0) / 2]; };
        const u64 k0 = f2_bcast(taps.k[0]);
        u64 a0 = f2_mul(k0, T(0)), a1 = f2_mul(k0, T(2));
#pragma unroll
        for (int i = 1; i <= R; ++i) {
          const u64 ki = f2_bcast(taps.k[i]);
          a0 = f2_fma(ki, f2_add(T(-i), T(i)), a0);
          a1 = f2_fma(ki, f2_add(T(2 - i), T(2 + i)), a1);
        }
        const ulonglong2 o2 = make_ulonglong2(a0, a1);
        *reinterpret_cast<ulonglong2*>(&ring[slot][4 * lane]) = o2;
        if (slot < MIRROR) *reinterpret_cast<ulonglong2*>(&ring[RING + slot][4 * lane]) = o2;   // warp-uniform
      } else {
      const float* p = &sg[rr][4 * lane];                       // box col 0 <-> x0 - 16
      constexpr int Q0 = (TS_HALO - R) / 4, Q1 = (TS_HALO + 4 + R + 3) / 4;   // float4s that hold needed taps
      float win[4 * (Q1 - Q0)];
#pragma unroll
      for (int q = Q0; q < Q1; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
        win[4 * (q - Q0)] = v.x; win[4 * (q - Q0) + 1] = v.y; win[4 * (q - Q0) + 2] = v.z; win[4 * (q - Q0) + 3] = v.w;
      }
      constexpr int C = TS_HALO - 4 * Q0;                        // index of output 0's centre tap in win
      float acc[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = taps.k[0] * win[C + o];
#pragma unroll
      for (int i = 1; i <= R; ++i) {
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[o] = fmaf(taps.k[i], win[C + o - i] + win[C + o + i], acc[o]);
      }
      const float4 o4 = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(&ring[slot][4 * lane]) = o4;
      if (slot < MIRROR) *reinterpret_cast<float4*>(&ring[RING + slot][4 * lane]) = o4;   // warp-uniform
      }
    }
    }
  };

  const int cp = tid & 63, rg = tid >> 6;
  const int x = x0 + 2 * cp;
  row_pass_group(0);
  __syncthreads();
  if (tid == 0 && 2 < n_groups) issue(2);
  if (n_groups > 1) row_pass_group(1);
  __syncthreads();
  if (tid == 0 && 3 < n_groups) issue(3);
  for (int c = 0; c < n_chunks; ++c) {
    row_pass_group(c + 2);
    __syncthreads();
    if (tid == 0 && c + 4 < n_groups) issue(c + 4);
    if constexpr (MODE == 3) {
      // column pass with EIGHT rows and ONE column per thread: 8 + 2R ring values are loaded for 8 outputs (17 bytes of
      // shared memory per pixel at R = 13 instead of 30), scalar arithmetic (the same IEEE operations as the packed form)
      const int col = tid & 127, rg8 = tid >> 7;
      const int xs = x0 + col;
      const int y8 = ys + c * TS_G + rg8 * 8;
      if (y8 < ye && xs < w) {
        float cc1[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) cc1[o] = (y8 + o < ye) ? __ldg(img + (unsigned)((y8 + o) * pitch + xs)) : 0.f;
        float win1[8 + 2 * R];
        const int s8 = (y8 - R) & (RING - 1);
        if (y8 - R >= 0 && y8 + 7 + R < h) {
          const float* base = &ring[s8][col];       // s8 + q <= RING - 1 + MIRROR
#pragma unroll
          for (int q = 0; q < 8 + 2 * R; ++q) win1[q] = base[q * TS_RS];
        } else {
#pragma unroll
          for (int q = 0; q < 8 + 2 * R; ++q) win1[q] = ring[reflect101(y8 - R + q, h) & (RING - 1)][col];
        }
        float a8[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) a8[o] = taps.k[0] * win1[R + o];
#pragma unroll
        for (int i = 1; i <= R; ++i) {
#pragma unroll
          for (int o = 0; o < 8; ++o) a8[o] = fmaf(taps.k[i], win1[R + o - i] + win1[R + o + i], a8[o]);
        }
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const int y = y8 + o;
          if (y < ye) {
            const unsigned g = (unsigned)(y * pitch + xs);
            dst_i[g] = a8[o];
            dog_i[g] = a8[o] - cc1[o];
            if constexpr (DS) {
              if (!((y | xs) & 1) && (y >> 1) < ds.h && (xs >> 1) < ds.w) ds.dst[((size_t)b * ds.h + (y >> 1)) * ds.pitch + (xs >> 1)] = a8[o];
            }
          }
        }
      }
      continue;
    }
    const int yf = ys + c * TS_G + rg * 4;
    if (yf < ye && x < w) {
      // centre pixels of G[l] for the fused DoG: issued first so their latency hides behind the filter
      float2 cc[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) cc[o] = (yf + o < ye) ? __ldg(reinterpret_cast<const float2*>(img + (unsigned)((yf + o) * pitch + x))) : make_float2(0.f, 0.f);
      u64 win[4 + 2 * R];
      const int s0 = (yf - R) & (RING - 1);
      if (yf - R >= 0 && yf + 3 + R < h) {
        const float* base = &ring[s0][2 * cp];   // s0 + q <= RING - 1 + MIRROR
#pragma unroll
        for (int q = 0; q < 4 + 2 * R; ++q) win[q] = *reinterpret_cast<const u64*>(base + q * TS_RS);
      } else {
#pragma unroll
        for (int q = 0; q < 4 + 2 * R; ++q) {
          const int ry = reflect101(yf - R + q, h);
          win[q] = *reinterpret_cast<const u64*>(&ring[ry & (RING - 1)][2 * cp]);
        }
      }
      u64 acc[4];
      const u64 k0 = f2_bcast(taps.k[0]);
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = f2_mul(k0, win[R + o]);
#pragma unroll
      for (int i = 1; i <= R; ++i) {
        const u64 ki = f2_bcast(taps.k[i]);
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[o] = f2_fma(ki, f2_add(win[R + o - i], win[R + o + i]), acc[o]);
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int y = yf + o;
        if (y < ye) {
          const unsigned g = (unsigned)(y * pitch + x);
          const float2 v = *reinterpret_cast<const float2*>(&acc[o]);
          *reinterpret_cast<float2*>(dst_i + g) = v;
          *reinterpret_cast<float2*>(dog_i + g) = make_float2(v.x - cc[o].x, v.y - cc[o].y);
          if constexpr (DS) {                                                         // x is even: v.x is column 2 * (x / 2)
            if (!(y & 1) && (y >> 1) < ds.h && (x >> 1) < ds.w) ds.dst[((size_t)b * ds.h + (y >> 1)) * ds.pitch + (x >> 1)] = v.x;
          }
        }
      }
    }
  }
}

constexpr int ST_W = 128, ST_CH = 16;   // strip width / rows per chunk of the streaming base-image kernel

// Base image, streaming: u8 -> 2x bilinear upsample (cv::resize convention) -> blur(sig_diff), same
// strip-walking structure as sift_blur_tma_kernel (without TMA: the source is the small u8 image).  The upsample is separable and the contract's
// value  wya*(wxa*a + wxb*b) + wyb*(wxa*c + wxb*d)  only mixes two SOURCE rows, so the horizontal
// blends h(r, x) = wxa*img[r][xa] + wxb*img[r][xb] are formed once per source row into a small ring
// and every upsampled sample is 2 multiplies + 1 add on top of them (identical operations, identical
// order, bit-identical result).
template <int R>
__global__ void __launch_bounds__(256, 3)
sift_base_stream_kernel(const uint8_t* __restrict__ img8, float* __restrict__ dst, int h, int w, int pitch,
                        int rows8, int cols8, int seg_rows, const Taps taps) {
  constexpr int RP = (R + 3) & ~3;
  constexpr int NWIN = 4 + 2 * RP;
  constexpr int RING = 64, HRING = 16, HW = ST_W + 2 * RP;
  static_assert(ST_CH + 2 * R <= RING, "ring too small");
  __shared__ __align__(16) float ring[RING][ST_W];
  __shared__ __align__(16) float hring[HRING][HW];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * ST_W;
  const int ys = blockIdx.y * seg_rows, ye = min(ys + seg_rows, h);
  const uint8_t* src = img8 + (size_t)b * rows8 * cols8;
  float* out = dst + (size_t)b * h * pitch;
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;

  auto src_rows = [&](int y, int& ya, int& yb) {
    ya = (y & 1) ? (y >> 1) : (y >> 1) - 1; yb = ya + 1;
    if (ya < 0) ya = 0;
    if (yb > rows8 - 1) yb = rows8 - 1;
  };
  int h_done = -1;   // source rows <= h_done are (or were) in the ring; block-uniform
  // each lane owns columns q = lane + 32k of the blend ring: their source columns and weights never change
  constexpr int HQ = (HW + 31) / 32;
  int qxa[HQ], qxb[HQ]; float qwb[HQ];
#pragma unroll
  for (int k = 0; k < HQ; ++k) {
    const int x = reflect101(x0 - RP + lane + 32 * k, w);
    int xa = (x & 1) ? (x >> 1) : (x >> 1) - 1, xb = xa + 1;
    qwb[k] = (x & 1) ? 0.25f : 0.75f;
    qxa[k] = xa < 0 ? 0 : xa;
    qxb[k] = xb > cols8 - 1 ? cols8 - 1 : xb;
  }
  auto ensure_h = [&](int y_lo, int y_hi) {   // h rows needed by upsampled rows [y_lo, y_hi]; warp per source row
    int lo, hi, t;
    src_rows(y_lo, lo, t); src_rows(y_hi, t, hi);
    if (lo <= h_done) lo = h_done + 1;
    for (int r = lo + wrp; r <= hi; r += 8) {
      const uint8_t* rowp = src + (size_t)r * cols8;
      float a[HQ], c[HQ];
#pragma unroll
      for (int k = 0; k < HQ; ++k) { a[k] = rowp[qxa[k]]; c[k] = rowp[qxb[k]]; }
#pragma unroll
      for (int k = 0; k < HQ; ++k)
        if (lane + 32 * k < HW) hring[r & (HRING - 1)][lane + 32 * k] = (1.0f - qwb[k]) * a[k] + qwb[k] * c[k];
    }
    if (hi > h_done) h_done = hi;
  };
  auto row_pass = [&](int row) {
    int ya, yb;
    src_rows(row, ya, yb);
    const float wyb = (row & 1) ? 0.25f : 0.75f, wya = 1.0f - wyb;
    const float4* pa = reinterpret_cast<const float4*>(&hring[ya & (HRING - 1)][4 * lane]);
    const float4* pb = reinterpret_cast<const float4*>(&hring[yb & (HRING - 1)][4 * lane]);
    float win[NWIN];
#pragma unroll
    for (int q = 0; q < NWIN / 4; ++q) {
      const float4 va = pa[q], vb = pb[q];
      const float ea[4] = {va.x, va.y, va.z, va.w}, eb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)   // only the taps the filter reads: [RP - R, RP + 4 + R)
        if (4 * q + j >= RP - R && 4 * q + j < RP + 4 + R) win[4 * q + j] = wya * ea[j] + wyb * eb[j];
    }
    float acc[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[o] = taps.k[0] * win[RP + o];
#pragma unroll
    for (int i = 1; i <= R; ++i) {
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = fmaf(taps.k[i], win[RP + o - i] + win[RP + o + i], acc[o]);
    }
    *reinterpret_cast<float4*>(&ring[row & (RING - 1)][4 * lane]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  };

  {
    const int lo = max(ys - R, 0), hi = min(ys + R, h);
    ensure_h(lo, hi - 1);
    __syncthreads();
    for (int row = lo + wrp; row < hi; row += 8) row_pass(row);
  }
  const int cp = tid & 63, rg = tid >> 6;
  const int x = x0 + 2 * cp;
  for (int yc = ys; yc < ye; yc += ST_CH) {
    const int lo = yc + R, hi = min(yc + ST_CH + R, h);
    __syncthreads();                       // previous chunk's column pass is done with the rings
    if (lo < hi) ensure_h(lo, hi - 1);
    __syncthreads();
    for (int row = lo + wrp; row < hi; row += 8) row_pass(row);
    __syncthreads();
    const int yf = yc + rg * 4;
    if (yf < ye && x < w) {
      u64 win[4 + 2 * R];
      if (yf - R >= 0 && yf + 3 + R < h) {
#pragma unroll
        for (int q = 0; q < 4 + 2 * R; ++q)
          win[q] = *reinterpret_cast<const u64*>(&ring[(yf - R + q) & (RING - 1)][2 * cp]);
      } else {
#pragma unroll
        for (int q = 0; q < 4 + 2 * R; ++q) {
          const int ry = reflect101(yf - R + q, h);
          win[q] = *reinterpret_cast<const u64*>(&ring[ry & (RING - 1)][2 * cp]);
        }
      }
      u64 acc[4];
      const u64 k0 = f2_bcast(taps.k[0]);
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = f2_mul(k0, win[R + o]);
#pragma unroll
      for (int i = 1; i <= R; ++i) {
        const u64 ki = f2_bcast(taps.k[i]);
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[o] = f2_fma(ki, f2_add(win[R + o - i], win[R + o + i]), acc[o]);
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int y = yf + o;
        if (y < ye) *reinterpret_cast<float2*>(out + (size_t)y * pitch + x) = *reinterpret_cast<const float2*>(&acc[o]);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
sift_downsample_kernel(const float* __restrict__ src, float* __restrict__ dst, int sh, int spitch, int dh, int dw, int dpitch) {
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x < dw) dst[(size_t)b * dh * dpitch + (size_t)y * dpitch + x] = __ldg(src + (size_t)b * sh * spitch + (size_t)(2 * y) * spitch + 2 * x);
}

// ------------------------------------------------------ small octaves: one launch for all of them
// Octaves too small for the streaming kernels (a few thousand pixels and less) used to cost one
// launch per layer.  Here one block per image builds ALL of them: layer 0 of each octave is the 2x
// decimation of the previous octave's layer nl, then each layer is blurred from the previous one
// entirely in shared memory (row pass, column pass, reflect-101), written to the Gaussian stack and
// differenced into the DoG stack.  Same operations in the same order as sift_blur_dog_kernel.
struct TapsAll { Taps t[8]; };
__global__ void __launch_bounds__(1024)
sift_small_octaves_kernel(float* __restrict__ gauss, float* __restrict__ dog, const OctInfo oi, int first_oct,
                          int n_oct, int nl, int batch_stride, int cap, const TapsAll taps) {
  extern __shared__ float ssm[];
  float* A = ssm; float* B = ssm + cap; float* T = ssm + 2 * cap;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  for (int o = first_oct; o < n_oct; ++o) {
    const int h = oi.h[o], w = oi.w[o], pitch = oi.pitch[o];
    const size_t lstride = (size_t)batch_stride * h * pitch;
    float* g0 = gauss + oi.goff[o] + (size_t)b * h * pitch;
    float* d0 = dog + oi.doff[o] + (size_t)b * h * pitch;
    const int n = h * w;
    if (o == 0) {
      for (int idx = tid; idx < n; idx += nt) { const int y = idx / w, x = idx - y * w; A[idx] = g0[(size_t)y * pitch + x]; }
    } else {
      const int hp = oi.h[o - 1], pp = oi.pitch[o - 1];
      const float* src = gauss + oi.goff[o - 1] + (size_t)nl * batch_stride * hp * pp + (size_t)b * hp * pp;
      for (int idx = tid; idx < n; idx += nt) {
        const int y = idx / w, x = idx - y * w;
        const float v = src[(size_t)(2 * y) * pp + 2 * x];
        A[idx] = v;
        g0[(size_t)y * pitch + x] = v;
      }
    }
    __syncthreads();
    for (int i = 1; i < nl + 3; ++i) {
      const Taps& tp = taps.t[i];
      const int R = tp.r;
      for (int idx = tid; idx < n; idx += nt) {
        const int y = idx / w, x = idx - y * w;
        const float* row = A + y * w;
        float acc = tp.k[0] * row[x];
        if (x >= R && x + R < w) {
          for (int j = 1; j <= R; ++j) acc = fmaf(tp.k[j], row[x - j] + row[x + j], acc);
        } else {
          for (int j = 1; j <= R; ++j) acc = fmaf(tp.k[j], row[reflect101(x - j, w)] + row[reflect101(x + j, w)], acc);
        }
        T[idx] = acc;
      }
      __syncthreads();
      float* gi = g0 + (size_t)i * lstride;
      float* di = d0 + (size_t)(i - 1) * lstride;
      for (int idx = tid; idx < n; idx += nt) {
        const int y = idx / w, x = idx - y * w;
        float acc = tp.k[0] * T[idx];
        if (y >= R && y + R < h) {
          for (int j = 1; j <= R; ++j) acc = fmaf(tp.k[j], T[idx - j * w] + T[idx + j * w], acc);
        } else {
          for (int j = 1; j <= R; ++j) acc = fmaf(tp.k[j], T[reflect101(y - j, h) * w + x] + T[reflect101(y + j, h) * w + x], acc);
        }
        B[idx] = acc;
        gi[(size_t)y * pitch + x] = acc;
        di[(size_t)y * pitch + x] = acc - A[idx];
      }
      __syncthreads();
      float* t2 = A; A = B; B = t2;
    }
  }
}

// ----------------------------------------------------------------------- extrema detection
// 3x3x3 extrema of the DoG stack, separable and barrier-free.  A pixel of layer l is a maximum iff
// val >= max3x3(l-1), max3x3(l), max3x3(l+1) (its own 3x3 max contains val itself), symmetrically
// for minima -- the same predicate as comparing with all 26 neighbours.  One warp sweeps a strip
// of 30 columns (lanes 1..30; lanes 0 and 31 are halo) down EX_ROWS rows: the horizontal 3-max/min
// comes from two warp shuffles, the vertical one from a rolling 3-row register window, all NL+2
// layers are carried together, so every DoG value is loaded from HBM once (coalesced 128-byte rows)
// and there is no shared memory and no barrier.  Candidates are compacted with warp ballots.
// candidate word = oct<<28 | layer<<25 | y<<13 | x
constexpr int EX_COLS = 30, EX_ROWS = 32;
template <int NL>
__global__ void __launch_bounds__(128)
sift_extrema_kernel(const float* __restrict__ dog_oct, int oct, int batch, int h, int w, int pitch,
                    float threshold, uint32_t* __restrict__ cand, int cand_cap, int* __restrict__ counters) {
  constexpr int L = NL + 2;
  const int lane = threadIdx.x & 31;
  const int strip = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int x = strip * EX_COLS - 1 + lane;
  if (strip * EX_COLS >= w) return;                       // warp-uniform
  const int xc = min(max(x, 0), w - 1);
  const int y0 = blockIdx.y * EX_ROWS;
  const int b = blockIdx.z;
  const size_t layer_stride = (size_t)batch * h * pitch;
  // one 64-bit base per layer stays in registers; a pixel address is then ONE wide multiply-add of the 32-bit element
  // offset row * pitch + column (the row part is warp-uniform), instead of a 64-bit add chain per layer and row
  const float* lp[L];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    lp[l] = dog_oct + (size_t)b * h * pitch + (size_t)l * layer_stride;
    asm volatile("" : "+l"(lp[l]));
  }
  float hxA[L], hxB[L], hxC[L], hnA[L], hnB[L], hnC[L], cB[L], cC[L];
#pragma unroll
  for (int l = 0; l < L; ++l) { hxA[l] = hxB[l] = hxC[l] = hnA[l] = hnB[l] = hnC[l] = cB[l] = cC[l] = 0.f; }
  const bool col_ok = lane >= 1 && lane <= EX_COLS && x >= SIFT_BORDER && x < w - SIFT_BORDER;
  const int y_end = min(y0 + EX_ROWS, h - SIFT_BORDER);    // last row (exclusive) that can hold a keypoint
  // the next row's values are loaded one iteration ahead, so the shuffles never wait on HBM (two rows
  // ahead and L2 prefetches further down were measured slower).  Unrolling by 3 lets the compiler rename the
  // rolling 3-row window instead of moving registers every row.
  float nxt[L];
  {
    const unsigned off = (unsigned)(min(max(y0 - 1, 0), h - 1) * pitch) + (unsigned)xc;
#pragma unroll
    for (int l = 0; l < L; ++l) nxt[l] = __ldg(lp[l] + off);
  }
#pragma unroll 3
  for (int r = y0 - 1; r <= y_end; ++r) {
    float v[L];
#pragma unroll
    for (int l = 0; l < L; ++l) v[l] = nxt[l];
    if (r < y_end) {
      const unsigned off = (unsigned)(min(max(r + 1, 0), h - 1) * pitch) + (unsigned)xc;
#pragma unroll
      for (int l = 0; l < L; ++l) nxt[l] = __ldg(lp[l] + off);
    }
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const float lf = __shfl_up_sync(0xffffffffu, v[l], 1), rt = __shfl_down_sync(0xffffffffu, v[l], 1);
      hxA[l] = hxB[l]; hxB[l] = hxC[l]; hnA[l] = hnB[l]; hnB[l] = hnC[l]; cB[l] = cC[l];
      hxC[l] = fmaxf(fmaxf(lf, v[l]), rt);
      hnC[l] = fminf(fminf(lf, v[l]), rt);
      cC[l] = v[l];
    }
    const int y = r - 1;                                   // the row whose 3x3 windows are now complete
    if (y < y0 || y < SIFT_BORDER || y >= y_end) continue;  // warp-uniform
    float mx[L], mn[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
      mx[l] = fmaxf(fmaxf(hxA[l], hxB[l]), hxC[l]);
      mn[l] = fminf(fminf(hnA[l], hnB[l]), hnC[l]);
    }
    // branch-free tests (a maximum needs val > threshold > 0, a minimum val < -threshold); candidates are rare, so one
    // vote over the three layers decides whether the compaction below runs at all
    bool ext_l[NL + 1];
    bool any_ext = false;
#pragma unroll
    for (int layer = 1; layer <= NL; ++layer) {
      const float val = cB[layer];
      const float m27 = fmaxf(fmaxf(mx[layer - 1], mx[layer]), mx[layer + 1]);
      const float n27 = fminf(fminf(mn[layer - 1], mn[layer]), mn[layer + 1]);
      ext_l[layer] = col_ok & (((val > threshold) & (val >= m27)) | ((val < -threshold) & (val <= n27)));
      any_ext |= ext_l[layer];
    }
    if (!__any_sync(0xffffffffu, any_ext)) continue;       // warp-uniform
#pragma unroll
    for (int layer = 1; layer <= NL; ++layer) {
      const bool ext = ext_l[layer];
      const unsigned mask = __ballot_sync(0xffffffffu, ext);
      if (mask) {
        const int leader = __ffs(mask) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&counters[b * 4 + 0], __popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (ext) {
          const int slot = base + __popc(mask & ((1u << lane) - 1u));
          if (slot < cand_cap)
            cand[(size_t)b * cand_cap + slot] = ((uint32_t)oct << 28) | ((uint32_t)layer << 25) | ((uint32_t)y << 13) | (uint32_t)x;
        }
      }
    }
  }
}

// (Two columns per lane -- one 8-byte load and two shuffles per layer and row for two pixels -- was built and measured in
// round 2: 96-103 registers, 16-20 warps per SM instead of 36, 0.96 ms per step against 0.92: removed again.)
// TMA-fed form for the wide octaves (an experiment kept behind VO_EXT_TMA=1, bit-identical results).  The register
// form above keeps only one row per layer in flight per warp (about 20 KB per SM) and runs at 48 % of the HBM
// roofline, so the question was whether it starves for bytes in flight.  Here a block of 8 warps walks a strip of 240
// columns down a row segment and the DoG rows arrive through cp.async.bulk.tensor in groups of 4 rows x (NL + 2)
// layers x 256 columns (20 KB), three groups deep: 40 KB per block and 120 KB per SM in flight without holding a
// register.  Measured: 1.26 ms per step against 1.03 ms (ncu: same 1.5e8 warp instructions, issue-active 63 % against
// 71 %, top stalls wait + barrier instead of long_scoreboard): the test is bound by instruction issue, not by memory
// parallelism -- as were a second row in registers, per-lane cp.async rings (round 1) and L2 prefetches (round 2, removed again).
// Out-of-image halo columns / rows are zero-filled by TMA and only ever feed windows of border pixels, which cannot
// be keypoints.
constexpr int XT_WARPS = 8, XT_COLS = XT_WARPS * EX_COLS, XT_BOXW = 256, XT_G = 4, XT_NST = 3;
template <int NL>
__global__ void __launch_bounds__(XT_WARPS * 32, 3)
sift_extrema_tma_kernel(const __grid_constant__ CUtensorMap tm, int oct, int batch_stride, int h, int w, int seg_rows,
                        float threshold, uint32_t* __restrict__ cand, int cand_cap, int* __restrict__ counters) {
  constexpr int L = NL + 2;
  extern __shared__ __align__(128) uint8_t xsm[];
  float (*stage)[L][XT_G][XT_BOXW] = reinterpret_cast<float (*)[L][XT_G][XT_BOXW]>(xsm);     // [XT_NST]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(xsm + XT_NST * L * XT_G * XT_BOXW * 4);
  constexpr uint32_t GROUP_BYTES = L * XT_G * XT_BOXW * 4, LAYER_BYTES = XT_G * XT_BOXW * 4;
  const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
  const int b = blockIdx.z;
  const int xs = blockIdx.x * XT_COLS;
  const int ys = max(blockIdx.y * seg_rows, SIFT_BORDER), ye = min((blockIdx.y + 1) * seg_rows, h - SIFT_BORDER);
  if (ys >= ye) return;                                   // block-uniform
  const uint32_t bar0 = smem_u32(&bars[0]);
  const uint32_t stage0 = smem_u32(&stage[0][0][0][0]);
  if (tid == 0) {
    for (int k = 0; k < XT_NST; ++k) mbar_init(bar0 + 8 * k, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // group g = rows [ys - 1 + XT_G * g, ys - 1 + XT_G * (g + 1)); rows ys - 1 .. ye are needed
  const int n_groups = (ye - ys + 2 + XT_G - 1) / XT_G;
  auto issue = [&](int g) {   // one thread
    const int slot = g % XT_NST;
    const uint32_t bar = bar0 + 8 * slot;
    mbar_expect_tx(bar, GROUP_BYTES);
#pragma unroll
    for (int l = 0; l < L; ++l)
      tma_load_3d(stage0 + slot * GROUP_BYTES + l * LAYER_BYTES, &tm, bar, xs - 4, ys - 1 + XT_G * g, l * batch_stride + b);
  };
  if (tid == 0)
    for (int g = 0; g < XT_NST && g < n_groups; ++g) issue(g);
  const int x = xs - 1 + EX_COLS * wq + lane;             // global column of this lane (lanes 0 and 31 are halo)
  const int bc = 3 + EX_COLS * wq + lane;                 // its column inside the box (box column 0 <-> xs - 4)
  const bool col_ok = lane >= 1 && lane <= EX_COLS && x >= SIFT_BORDER && x < w - SIFT_BORDER;
  const bool warp_on = xs + EX_COLS * wq < w;             // warp-uniform: this warp's strip holds image columns
  float hxA[L], hxB[L], hxC[L], hnA[L], hnB[L], hnC[L], cB[L], cC[L];
#pragma unroll
  for (int l = 0; l < L; ++l) { hxA[l] = hxB[l] = hxC[l] = hnA[l] = hnB[l] = hnC[l] = cB[l] = cC[l] = 0.f; }
  for (int g = 0; g < n_groups; ++g) {
    const int slot = g % XT_NST;
    mbar_wait(bar0 + 8 * slot, (uint32_t)((g / XT_NST) & 1));
    if (warp_on) {
#pragma unroll
      for (int rr = 0; rr < XT_G; ++rr) {
        const int r = ys - 1 + XT_G * g + rr;             // the row now entering the window
        if (r > ye) break;                                // block-uniform
        float v[L];
#pragma unroll
        for (int l = 0; l < L; ++l) v[l] = stage[slot][l][rr][bc];
#pragma unroll
        for (int l = 0; l < L; ++l) {
          const float lf = __shfl_up_sync(0xffffffffu, v[l], 1), rt = __shfl_down_sync(0xffffffffu, v[l], 1);
          hxA[l] = hxB[l]; hxB[l] = hxC[l]; hnA[l] = hnB[l]; hnB[l] = hnC[l]; cB[l] = cC[l];
          hxC[l] = fmaxf(fmaxf(lf, v[l]), rt);
          hnC[l] = fminf(fminf(lf, v[l]), rt);
          cC[l] = v[l];
        }
        const int y = r - 1;                              // the row whose 3x3 windows are now complete
        if (y < ys) continue;                             // block-uniform (window warm-up)
        float mx[L], mn[L];
#pragma unroll
        for (int l = 0; l < L; ++l) {
          mx[l] = fmaxf(fmaxf(hxA[l], hxB[l]), hxC[l]);
          mn[l] = fminf(fminf(hnA[l], hnB[l]), hnC[l]);
        }
#pragma unroll
        for (int layer = 1; layer <= NL; ++layer) {
          const float val = cB[layer];
          bool ext = false;
          if (col_ok && fabsf(val) > threshold) {
            if (val > 0) ext = val >= mx[layer - 1] && val >= mx[layer] && val >= mx[layer + 1];
            else ext = val <= mn[layer - 1] && val <= mn[layer] && val <= mn[layer + 1];
          }
          const unsigned mask = __ballot_sync(0xffffffffu, ext);
          if (mask) {
            const int leader = __ffs(mask) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&counters[b * 4 + 0], __popc(mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (ext) {
              const int slot_c = base + __popc(mask & ((1u << lane) - 1u));
              if (slot_c < cand_cap)
                cand[(size_t)b * cand_cap + slot_c] = ((uint32_t)oct << 28) | ((uint32_t)layer << 25) | ((uint32_t)y << 13) | (uint32_t)x;
            }
          }
        }
      }
    }
    __syncthreads();                                      // every warp is done with this slot
    if (tid == 0 && g + XT_NST < n_groups) issue(g + XT_NST);
  }
}

// --------------------------------------------------------- refinement + orientation (warp/cand)

// Refinement: THREAD per candidate (the sub-pixel fit is scalar work; a warp per candidate executed
// it 32 times over).  Survivors are appended, warp-aggregated, to a compact list of refined records
// that the orientation kernel then processes with a warp each.
struct RefinedKp { float x, y, size, resp; int oct_word; uint32_t where; };   // where = o<<28 | layer<<25 | r<<13 | c
__global__ void __launch_bounds__(128)
sift_refine_kernel(const float* __restrict__ dog, const OctInfo oi, int batch, int nl, float contrast_thr, float edge_thr,
                   float sigma, const uint32_t* __restrict__ cand, int cand_cap, int* __restrict__ counters,
                   RefinedKp* __restrict__ refined) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int n_cand = min(counters[b * 4 + 0], cand_cap);
  const float img_scale = 1.f / 255.f, deriv_scale = img_scale * 0.5f, second_deriv_scale = img_scale,
              cross_deriv_scale = img_scale * 0.25f;
  for (int base = blockIdx.x * 128; base < n_cand; base += gridDim.x * 128) {
    const int ci = base + threadIdx.x;
    bool ok = ci < n_cand;
    float kx = 0, ky = 0, ksize = 0, kresp = 0; int koct = 0;
    uint32_t where = 0;
    if (ok) {
      do {
        const uint32_t word = cand[(size_t)b * cand_cap + ci];
        const int o = word >> 28;
        int layer = (word >> 25) & 7, r = (word >> 13) & 4095, c = word & 8191;
        const int rows = oi.h[o], cols = oi.w[o], pitch = oi.pitch[o];
        const size_t lstride = (size_t)batch * rows * pitch;
        const float* dbase = dog + oi.doff[o] + (size_t)b * rows * pitch;
        float xi = 0, xr = 0, xc = 0;
        int it = 0;
        for (; it < 5; ++it) {
          const float* img = dbase + (size_t)layer * lstride;
          const float* prv = img - lstride;
          const float* nxt = img + lstride;
          const size_t q = (size_t)r * pitch + c;
          const float dD0 = (img[q + 1] - img[q - 1]) * deriv_scale;
          const float dD1 = (img[q + pitch] - img[q - pitch]) * deriv_scale;
          const float dD2 = (nxt[q] - prv[q]) * deriv_scale;
          const float v2 = img[q] * 2.f;
          const float dxx = (img[q + 1] + img[q - 1] - v2) * second_deriv_scale;
          const float dyy = (img[q + pitch] + img[q - pitch] - v2) * second_deriv_scale;
          const float dss = (nxt[q] + prv[q] - v2) * second_deriv_scale;
          const float dxy = (img[q + pitch + 1] - img[q + pitch - 1] - img[q - pitch + 1] + img[q - pitch - 1]) * cross_deriv_scale;
          const float dxs = (nxt[q + 1] - nxt[q - 1] - prv[q + 1] + prv[q - 1]) * cross_deriv_scale;
          const float dys = (nxt[q + pitch] - nxt[q - pitch] - prv[q + pitch] + prv[q - pitch]) * cross_deriv_scale;
          const float a00 = dxx, a01 = dxy, a02 = dxs, a11 = dyy, a12 = dys, a22 = dss;
          const float m0 = a11 * a22 - a12 * a12, m1 = a01 * a22 - a12 * a02, m2 = a01 * a12 - a11 * a02;
          const float det = a00 * m0 - a01 * m1 + a02 * m2;
          float X0 = 0, X1 = 0, X2 = 0;
          if (det != 0.f) {
            const float d = __fdiv_rn(1.f, det);
            X0 = d * (dD0 * m0 - a01 * (dD1 * a22 - a12 * dD2) + a02 * (dD1 * a12 - a11 * dD2));
            X1 = d * (a00 * (dD1 * a22 - a12 * dD2) - dD0 * m1 + a02 * (a01 * dD2 - dD1 * a02));
            X2 = d * (a00 * (a11 * dD2 - dD1 * a12) - a01 * (a01 * dD2 - dD1 * a02) + dD0 * m2);
          }
          xi = -X2; xr = -X1; xc = -X0;
          if (fabsf(xi) < 0.5f && fabsf(xr) < 0.5f && fabsf(xc) < 0.5f) break;
          const float big = (float)(INT32_MAX / 3);
          if (fabsf(xi) > big || fabsf(xr) > big || fabsf(xc) > big) { ok = false; break; }
          c += __float2int_rn(xc); r += __float2int_rn(xr); layer += __float2int_rn(xi);
          if (layer < 1 || layer > nl || c < SIFT_BORDER || c >= cols - SIFT_BORDER || r < SIFT_BORDER || r >= rows - SIFT_BORDER) {
            ok = false; break;
          }
        }
        if (it >= 5) ok = false;
        if (ok) {
          const float* img = dbase + (size_t)layer * lstride;
          const float* prv = img - lstride;
          const float* nxt = img + lstride;
          const size_t q = (size_t)r * pitch + c;
          const float dD0 = (img[q + 1] - img[q - 1]) * deriv_scale;
          const float dD1 = (img[q + pitch] - img[q - pitch]) * deriv_scale;
          const float dD2 = (nxt[q] - prv[q]) * deriv_scale;
          const float t = dD0 * xc + dD1 * xr + dD2 * xi;
          const float contr = img[q] * img_scale + t * 0.5f;
          if (fabsf(contr) * nl < contrast_thr) ok = false;
          const float v2 = img[q] * 2.f;
          const float dxx = (img[q + 1] + img[q - 1] - v2) * second_deriv_scale;
          const float dyy = (img[q + pitch] + img[q - pitch] - v2) * second_deriv_scale;
          const float dxy = (img[q + pitch + 1] - img[q + pitch - 1] - img[q - pitch + 1] + img[q - pitch - 1]) * cross_deriv_scale;
          const float tr = dxx + dyy, det = dxx * dyy - dxy * dxy;
          if (det <= 0 || tr * tr * edge_thr >= (edge_thr + 1) * (edge_thr + 1) * det) ok = false;
          kx = (c + xc) * (float)(1 << o);
          ky = (r + xr) * (float)(1 << o);
          koct = o + (layer << 8) + (__float2int_rn((xi + 0.5f) * 255.f) << 16);
          ksize = sigma * vo_expf(__fdiv_rn((float)layer + xi, (float)nl) * 0.693147180559945f) * (float)(1 << o) * 2.f;
          kresp = fabsf(contr);
        }
        where = ((uint32_t)o << 28) | ((uint32_t)layer << 25) | ((uint32_t)r << 13) | (uint32_t)c;
      } while (false);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, ok);
    if (mask) {
      const int leader = __ffs(mask) - 1;
      int slot0 = 0;
      if (lane == leader) slot0 = atomicAdd(&counters[b * 4 + 3], __popc(mask));
      slot0 = __shfl_sync(0xffffffffu, slot0, leader);
      if (ok) {
        RefinedKp rk; rk.x = kx; rk.y = ky; rk.size = ksize; rk.resp = kresp; rk.oct_word = koct; rk.where = where;
        refined[(size_t)b * cand_cap + slot0 + __popc(mask & ((1u << lane) - 1u))] = rk;
      }
    }
  }
}

// Orientation: WARP per refined keypoint.
__global__ void __launch_bounds__(128)
sift_orient_kernel(const float* __restrict__ gauss, const OctInfo oi, int batch, const RefinedKp* __restrict__ refined,
                   int cand_cap, int* __restrict__ counters, vo_keypoint* __restrict__ raw, int kp_cap, int* __restrict__ work) {
  __shared__ uint32_t s_hist[4][ORI_BINS];
  __shared__ float s_sm[4][ORI_BINS + 4];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int n_ref = min(counters[b * 4 + 3], cand_cap);
  // candidates are handed out one at a time (their windows differ 4x in size: a fixed assignment leaves warps idle)
  for (;;) {
    int ci = 0;
    if (lane == 0) ci = atomicAdd(&work[b * 2 + 0], 1);
    ci = __shfl_sync(0xffffffffu, ci, 0);
    if (ci >= n_ref) break;
    const RefinedKp rk = refined[(size_t)b * cand_cap + ci];
    const int o = rk.where >> 28, layer = (rk.where >> 25) & 7, r = (rk.where >> 13) & 4095, c = rk.where & 8191;
    const int rows = oi.h[o], cols = oi.w[o], pitch = oi.pitch[o];
    const size_t lstride = (size_t)batch * rows * pitch;
    const float kx = rk.x, ky = rk.y, ksize = rk.size, kresp = rk.resp; const int koct = rk.oct_word;
    // ---- orientation histogram over the (2*radius+1)^2 window of gauss[o][layer]
    const float scl_octv = ksize * 0.5f / (float)(1 << o);
    const int radius = __float2int_rn(4.5f * scl_octv);
    const float osig = 1.5f * scl_octv;
    const float expf_scale = __fdiv_rn(-1.f, 2.f * osig * osig);
    const float* gimg = gauss + oi.goff[o] + (size_t)layer * lstride + (size_t)b * rows * pitch;
    asm volatile("" : "+l"(gimg));   // one 64-bit register pair: each address below is a single IMAD.WIDE
    for (int k = lane; k < ORI_BINS; k += 32) s_hist[wib][k] = 0;
    __syncwarp();
    const int side = 2 * radius + 1, total = side * side;
    // two samples per lane in flight (loads first, arithmetic after); (i, j) advance incrementally
    struct OSamp { float xp, xm, yu, yd; int i, j; bool ok; };
    auto ofetch = [&](int i, int j, OSamp& sm) {
      const int y = r + i, x = c + j;
      sm.i = i; sm.j = j;
      sm.ok = !(y <= 0 || y >= rows - 1 || x <= 0 || x >= cols - 1);
      if (sm.ok) {
        const unsigned ctr = (unsigned)(y * pitch + x);
        const float* q = gimg + ctr;
        sm.xp = __ldg(q + 1); sm.xm = __ldg(q - 1);
        sm.yu = __ldg(gimg + (ctr - (unsigned)pitch)); sm.yd = __ldg(gimg + (ctr + (unsigned)pitch));
      }
    };
    auto oaccum = [&](const OSamp& sm) {
      if (!sm.ok) return;
      const float dx = sm.xp - sm.xm;
      const float dy = sm.yu - sm.yd;
      // |i|, |j| <= radius = rn(3 * osig): the argument stays above -36 for every scale
      const float wgt = vo_expf_window((float)(sm.i * sm.i + sm.j * sm.j) * expf_scale);
      const float ang = vo_atan2deg(dy, dx);
      const float mag = __fsqrt_rn(fmaf(dx, dx, dy * dy));
      int bin = __float2int_rn((ORI_BINS / 360.f) * ang);
      if (bin >= ORI_BINS) bin -= ORI_BINS;
      if (bin < 0) bin += ORI_BINS;
      atomicAdd(&s_hist[wib][bin], sift_fix(wgt * mag));
    };
    int wi = lane / side - radius, wj = lane % side - radius;
    for (int idx = lane; idx < total; idx += 64) {
      OSamp s0, s1;
      ofetch(wi, wj, s0);
      wj += 32; while (wj > radius) { wj -= side; ++wi; }
      s1.ok = false;
      if (idx + 32 < total) ofetch(wi, wj, s1);
      wj += 32; while (wj > radius) { wj -= side; ++wi; }
      oaccum(s0);
      oaccum(s1);
    }
    __syncwarp();
    float* th = s_sm[wib];   // th[i + 2] = raw bin i, with 2 wrapped entries on each side
    for (int k = lane; k < ORI_BINS; k += 32) th[k + 2] = (float)s_hist[wib][k] * SIFT_INV_FIX;
    __syncwarp();
    if (lane == 0) { th[1] = th[ORI_BINS + 1]; th[0] = th[ORI_BINS]; th[ORI_BINS + 2] = th[2]; th[ORI_BINS + 3] = th[3]; }
    __syncwarp();
    float hv[2];
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      const int k = lane + 32 * rep;
      hv[rep] = -1.f;
      if (k < ORI_BINS)
        hv[rep] = (th[k] + th[k + 4]) * (1.f / 16.f) + (th[k + 1] + th[k + 3]) * (4.f / 16.f) + th[k + 2] * (6.f / 16.f);
    }
    float mx = fmaxf(hv[0], hv[1]);
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    __syncwarp();
    // smoothed histogram back to shared memory for neighbour access
    float* hs = s_sm[wib];
    __syncwarp();
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) { const int k = lane + 32 * rep; if (k < ORI_BINS) hs[k] = hv[rep]; }
    __syncwarp();
    const float mag_thr = mx * 0.8f;
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      const int j = lane + 32 * rep;
      bool peak = false; float angle = 0.f;
      if (j < ORI_BINS) {
        const int l = j > 0 ? j - 1 : ORI_BINS - 1, r2 = j < ORI_BINS - 1 ? j + 1 : 0;
        const float hj = hs[j], hl = hs[l], hr = hs[r2];
        if (hj > hl && hj > hr && hj >= mag_thr) {
          float bin = j + __fdiv_rn(0.5f * (hl - hr), hl - 2 * hj + hr);
          bin = bin < 0 ? ORI_BINS + bin : bin >= ORI_BINS ? bin - ORI_BINS : bin;
          angle = 360.f - (360.f / ORI_BINS) * bin;
          if (fabsf(angle - 360.f) < FLT_EPSILON) angle = 0.f;
          peak = true;
        }
      }
      const unsigned mask = __ballot_sync(0xffffffffu, peak);
      if (mask) {
        const int leader = __ffs(mask) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&counters[b * 4 + 1], __popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (peak) {
          const int slot = base + __popc(mask & ((1u << lane) - 1u));
          if (slot < kp_cap) {
            vo_keypoint kp; kp.x = kx; kp.y = ky; kp.size = ksize; kp.angle = angle; kp.response = kresp; kp.octave = koct;
            raw[(size_t)b * kp_cap + slot] = kp;
          }
        }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ ordering + duplicates
__device__ __forceinline__ bool kp_before(const vo_keypoint& a, int ia, const vo_keypoint& b, int ib) {
  if (a.x != b.x) return a.x < b.x;
  if (a.y != b.y) return a.y < b.y;
  if (a.size != b.size) return a.size > b.size;
  if (a.angle != b.angle) return a.angle < b.angle;
  if (a.response != b.response) return a.response > b.response;
  if (a.octave != b.octave) return a.octave > b.octave;
  return ia < ib;
}

// rank sort: grid (kp_cap/256, batch)
__global__ void __launch_bounds__(256)
sift_rank_kernel(const vo_keypoint* __restrict__ raw, int kp_cap, const int* __restrict__ counters,
                 vo_keypoint* __restrict__ sorted) {
  __shared__ vo_keypoint tile[256];
  const int b = blockIdx.y;
  const int n = min(counters[b * 4 + 1], kp_cap);
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (blockIdx.x * 256 >= n) return;
  const vo_keypoint* src = raw + (size_t)b * kp_cap;
  vo_keypoint me;
  if (i < n) me = src[i];
  int rank = 0;
  for (int base = 0; base < n; base += 256) {
    __syncthreads();
    if (base + threadIdx.x < n) tile[threadIdx.x] = src[base + threadIdx.x];
    __syncthreads();
    const int m = min(256, n - base);
    if (i < n)
      for (int j = 0; j < m; ++j) rank += kp_before(tile[j], base + j, me, i) ? 1 : 0;
  }
  if (i < n) sorted[(size_t)b * kp_cap + rank] = me;
}

// Bucketed rank sort, one block per image.  The order is dominated by x, so keypoints are first
// counted into SORT_BUCKETS buckets of x (monotone in x, hence consistent with the full order), a block
// scan gives each bucket's first rank, and an element's rank inside its bucket is the number of
// bucket mates that precede it under kp_before (a handful of comparisons instead of n).
constexpr int SORT_BUCKETS = 2048;
__global__ void __launch_bounds__(1024)
sift_rank_bucket_kernel(const vo_keypoint* __restrict__ raw, int kp_cap, const int* __restrict__ counters,
                        vo_keypoint* __restrict__ sorted, float x_scale) {
  extern __shared__ int srt[];                 // start[SORT_BUCKETS + 1], fill[SORT_BUCKETS], perm[kp_cap] (u16 pairs)
  int* start = srt;
  int* fill = srt + SORT_BUCKETS + 1;
  unsigned short* perm = reinterpret_cast<unsigned short*>(fill + SORT_BUCKETS);
  __shared__ int warp_sums[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = min(counters[b * 4 + 1], kp_cap);
  const vo_keypoint* src = raw + (size_t)b * kp_cap;
  for (int k = tid; k <= SORT_BUCKETS; k += 1024) { start[k] = 0; if (k < SORT_BUCKETS) fill[k] = 0; }
  __syncthreads();
  auto bucket_of = [&](float x) { return min(SORT_BUCKETS - 1, max(0, (int)(x * x_scale))); };
  for (int i = tid; i < n; i += 1024) atomicAdd(&start[bucket_of(src[i].x) + 1], 1);
  __syncthreads();
  // inclusive scan of start[1..SORT_BUCKETS] (2 entries per thread)
  {
    const int a0 = start[2 * tid + 1], a1 = start[2 * tid + 2];
    int x = a0 + a1;
    const int lane = tid & 31, w = tid >> 5;
    for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
    if (lane == 31) warp_sums[w] = x;
    __syncthreads();
    if (w == 0) {
      int v = warp_sums[lane];
      for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, v, off); if (lane >= off) v += y; }
      warp_sums[lane] = v;
    }
    __syncthreads();
    const int incl = x + (w ? warp_sums[w - 1] : 0);
    start[2 * tid + 1] = incl - a1;
    start[2 * tid + 2] = incl;
  }
  __syncthreads();
  for (int i = tid; i < n; i += 1024) {
    const int bk = bucket_of(src[i].x);
    perm[start[bk] + atomicAdd(&fill[bk], 1)] = (unsigned short)i;
  }
  __syncthreads();
  for (int i = tid; i < n; i += 1024) {
    const vo_keypoint me = src[i];
    const int bk = bucket_of(me.x);
    int rank = start[bk];
    for (int q = start[bk]; q < start[bk + 1]; ++q) {
      const int j = perm[q];
      if (j != i) rank += kp_before(src[j], j, me, i) ? 1 : 0;
    }
    sorted[(size_t)b * kp_cap + rank] = me;
  }
}

// one block per image: drop exact duplicates (x, y, size, angle), compact, apply the first-octave
// (-1) adjustment: pt *= 0.5, size *= 0.5, octave word - 1
__global__ void __launch_bounds__(1024)
sift_dedupe_kernel(const vo_keypoint* __restrict__ sorted, int kp_cap, int* __restrict__ counters,
                   vo_keypoint* __restrict__ out, float loc_offset) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int b = blockIdx.x;
  const int n = min(counters[b * 4 + 1], kp_cap);
  const vo_keypoint* src = sorted + (size_t)b * kp_cap;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    bool keep = false; vo_keypoint me;
    if (i < n) {
      me = src[i];
      keep = true;
      if (i > 0) {
        const vo_keypoint p = src[i - 1];
        keep = (p.x != me.x || p.y != me.y || p.size != me.size || p.angle != me.angle);
      }
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
      const int pos = carry + (w ? warp_sums[w - 1] : 0) + __popc(ballot & ((1u << lane) - 1u));
      me.octave = (me.octave & ~255) | ((me.octave - 1) & 255);
      me.x = me.x * 0.5f + loc_offset; me.y = me.y * 0.5f + loc_offset; me.size *= 0.5f;
      out[(size_t)b * kp_cap + pos] = me;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_sums[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) counters[b * 4 + 2] = carry;
}

// ------------------------------------------------------------------------------ descriptors
// Per-keypoint rotation (float)cos/sin of the descriptor orientation, evaluated in FP64 like the
// oracle; kept out of the descriptor kernel so that kernel stays at a low register count.
__global__ void __launch_bounds__(256)
sift_trig_kernel(const vo_keypoint* __restrict__ kps, int kp_cap, const int* __restrict__ counters, float2* __restrict__ trig) {
  const int b = blockIdx.y;
  const int n = min(counters[b * 4 + 2], kp_cap);
  for (int k = blockIdx.x * 256 + threadIdx.x; k < n; k += gridDim.x * 256) {
    float ori = 360.f - kps[(size_t)b * kp_cap + k].angle;
    if (fabsf(ori - 360.f) < FLT_EPSILON) ori = 0.f;
    const double ang = (double)ori * (3.14159265358979323846 / 180.0);
    trig[(size_t)b * kp_cap + k] = make_float2((float)cos(ang), (float)sin(ang));
  }
}

// Warp per keypoint.  The (2r+1)^2 window is scanned 32 samples at a time with a cheap accept test
// (inside the rotated 4x4 bin grid and the image); accepted samples are compacted through a small
// per-warp queue so the expensive part (gradient, exp, atan2, trilinear scatter) always runs on a
// full warp.  The scatter uses integer shared-memory atomics into 4 privatised copies of the
// 6x6x10 histogram (copy = lane & 3) to cut same-address serialisation; integer sums make the
// result independent of the order and of the copy assignment.
constexpr int DESC_COPIES = 2;
constexpr int DESC_MAX_ROWS = 160;   // window rows handled by the interval scan (radius <= 79)
__global__ void __launch_bounds__(128, 10)
sift_descriptor_kernel(const float* __restrict__ gauss, const OctInfo oi, int batch, int nl,
                       const vo_keypoint* __restrict__ kps, const float2* __restrict__ trig, int kp_cap,
                       const int* __restrict__ counters, float loc_offset, float* __restrict__ desc,
                       unsigned long long* __restrict__ algo_bytes, int* __restrict__ work) {
  constexpr int D = 4, N = 8, HLEN = (D + 2) * (D + 2) * (N + 2);
  __shared__ uint32_t s_hist[4][DESC_COPIES * HLEN];
  __shared__ unsigned s_row[4][DESC_MAX_ROWS + 3];
  __shared__ __align__(16) float s_vec[4][128];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int n = min(counters[b * 4 + 2], kp_cap);
  uint32_t* hist = s_hist[wib] + (lane & (DESC_COPIES - 1)) * HLEN;
  unsigned long long my_bytes = 0;
  // keypoints are handed out one at a time (dynamic scheduling: window sizes differ 4x within an octave)
  for (;;) {
    int ki = 0;
    if (lane == 0) ki = atomicAdd(&work[b * 2 + 1], 1);
    ki = __shfl_sync(0xffffffffu, ki, 0);
    if (ki >= n) break;
    const vo_keypoint kp = kps[(size_t)b * kp_cap + ki];
    int oc = kp.octave & 255; const int layer = (kp.octave >> 8) & 255;
    oc = oc < 128 ? oc : (-128 | oc);
    const float scale = oc >= 0 ? __fdiv_rn(1.f, (float)(1 << oc)) : (float)(1 << -oc);
    const float size = kp.size * scale;
    float ori = 360.f - kp.angle;
    if (fabsf(ori - 360.f) < FLT_EPSILON) ori = 0.f;
    const int po = oc + 1;
    const int rows = oi.h[po], cols = oi.w[po], pitch = oi.pitch[po];
    const float* img = gauss + oi.goff[po] + (size_t)layer * batch * rows * pitch + (size_t)b * rows * pitch;
    asm volatile("" : "+l"(img));   // keep the layer pointer as one 64-bit register pair (addresses = one IMAD.WIDE each)
    const float ptx = (kp.x - loc_offset) * scale, pty = (kp.y - loc_offset) * scale;
    const float scl = size * 0.5f;
    const int px = __float2int_rn(ptx), py = __float2int_rn(pty);
    const float2 cs = trig[(size_t)b * kp_cap + ki];
    const float bins_per_rad = N / 360.f, exp_scale = -1.f / (D * D * 0.5f);
    const float hist_width = 3.0f * scl;
    int radius = __float2int_rn(hist_width * 1.4142135623730951f * (D + 1) * 0.5f);
    const int diag = (int)sqrt((double)cols * cols + (double)rows * rows);
    if (radius > diag) radius = diag;
    const float cos_t = __fdiv_rn(cs.x, hist_width), sin_t = __fdiv_rn(cs.y, hist_width);
    for (int k = lane; k < DESC_COPIES * HLEN; k += 32) s_hist[wib][k] = 0;
    __syncwarp();

    // fetch() applies the exact acceptance test of the contract (the row intervals below are only a
    // superset) and issues the four gradient loads; accumulate() does the arithmetic and the trilinear
    // scatter.  (Keeping two samples per lane in flight was measured slower: the kernel is issue-bound,
    // and the extra registers cost occupancy.)
    struct Samp { float c_rot, r_rot, rbin, cbin, xp, xm, yu, yd; bool ok; };
    auto fetch = [&](int i, int j, Samp& sm, auto clipped) {
      sm.c_rot = j * cos_t - i * sin_t;
      sm.r_rot = j * sin_t + i * cos_t;
      sm.rbin = sm.r_rot + D / 2 - 0.5f; sm.cbin = sm.c_rot + D / 2 - 0.5f;
      sm.ok = sm.rbin > -1 && sm.rbin < D && sm.cbin > -1 && sm.cbin < D;
      if constexpr (!decltype(clipped)::value) {   // (the row intervals are already clipped to the image)
        const int r = py + i, c = px + j;
        sm.ok = sm.ok && r > 0 && r < rows - 1 && c > 0 && c < cols - 1;
      }
      if (sm.ok) {   // unsigned 32-bit element indices into the layer image: one wide multiply-add per address
        const unsigned ctr = (unsigned)((py + i) * pitch + (px + j));
        const float* pc = img + ctr;
        sm.xp = __ldg(pc + 1); sm.xm = __ldg(pc - 1);
        sm.yu = __ldg(img + (ctr - (unsigned)pitch)); sm.yd = __ldg(img + (ctr + (unsigned)pitch));
      }
    };
    auto accumulate = [&](const Samp& sm) {
      if (!sm.ok) return;
      float rbin = sm.rbin, cbin = sm.cbin;
      const float dx = sm.xp - sm.xm;
      const float dy = sm.yu - sm.yd;
      // accepted samples have |c_rot|, |r_rot| < 2.5: the argument lies in (-1.6, 0]
      const float wgt = vo_expf_window((sm.c_rot * sm.c_rot + sm.r_rot * sm.r_rot) * exp_scale);
      const float a = vo_atan2deg(dy, dx);
      const float mag = __fsqrt_rn(fmaf(dx, dx, dy * dy)) * wgt;
      float obin = (a - ori) * bins_per_rad;
      const int r0 = (int)floorf(rbin), c0 = (int)floorf(cbin);
      int o0 = (int)floorf(obin);
      rbin -= r0; cbin -= c0; obin -= o0;
      if (o0 < 0) o0 += N;
      if (o0 >= N) o0 -= N;
      const float v_r1 = mag * rbin, v_r0 = mag - v_r1;
      const float v_rc11 = v_r1 * cbin, v_rc10 = v_r1 - v_rc11;
      const float v_rc01 = v_r0 * cbin, v_rc00 = v_r0 - v_rc01;
      const float v111 = v_rc11 * obin, v110 = v_rc11 - v111;
      const float v101 = v_rc10 * obin, v100 = v_rc10 - v101;
      const float v011 = v_rc01 * obin, v010 = v_rc01 - v011;
      const float v001 = v_rc00 * obin, v000 = v_rc00 - v001;
      uint32_t* hp = hist + ((r0 + 1) * (D + 2) + c0 + 1) * (N + 2) + o0;
      atomicAdd(hp, sift_fix(v000));
      atomicAdd(hp + 1, sift_fix(v001));
      atomicAdd(hp + (N + 2), sift_fix(v010));
      atomicAdd(hp + (N + 3), sift_fix(v011));
      atomicAdd(hp + (D + 2) * (N + 2), sift_fix(v100));
      atomicAdd(hp + (D + 2) * (N + 2) + 1, sift_fix(v101));
      atomicAdd(hp + (D + 3) * (N + 2), sift_fix(v110));
      atomicAdd(hp + (D + 3) * (N + 2) + 1, sift_fix(v111));
    };
    auto process = [&](int i, int j) { Samp sm; fetch(i, j, sm, std::false_type{}); accumulate(sm); };
    auto process_clipped = [&](int i, int j) { Samp sm; fetch(i, j, sm, std::true_type{}); accumulate(sm); };

    const int side = 2 * radius + 1, total = side * side;
    my_bytes += (unsigned long long)total * 4ull + 512ull;   // patch read + descriptor written
    if (side <= DESC_MAX_ROWS) {
      // Per window row i the accepted samples lie in one j-interval: |j*cos_t - i*sin_t| < 2.5 and
      // |j*sin_t + i*cos_t| < 2.5 (bin units), clipped to the window and the image.  Lanes compute a
      // conservative integer superset per row (floor/ceil of the real bounds), a warp scan turns the
      // row counts into offsets, then the warp sweeps the packed candidate list 32 at a time.
      int carry = 0;
      for (int rb = 0; rb < side; rb += 32) {
        const int ri = rb + lane;
        int jlo = 0, cntr = 0;
        if (ri < side) {
          const int i = ri - radius;
          float lo = (float)-radius, hi = (float)radius;
          const float ic = i * cos_t, is = i * sin_t;
          if (fabsf(cos_t) > 1e-6f) {       // |j*cos_t - is| < 2.5
            const float a0 = __fdividef(is - 2.5f, cos_t), a1 = __fdividef(is + 2.5f, cos_t);
            lo = fmaxf(lo, fminf(a0, a1)); hi = fminf(hi, fmaxf(a0, a1));
          } else if (!(fabsf(is) < 2.6f)) hi = lo - 1.f;
          if (fabsf(sin_t) > 1e-6f) {       // |j*sin_t + ic| < 2.5
            const float b0 = __fdividef(-ic - 2.5f, sin_t), b1 = __fdividef(-ic + 2.5f, sin_t);
            lo = fmaxf(lo, fminf(b0, b1)); hi = fminf(hi, fmaxf(b0, b1));
          } else if (!(fabsf(ic) < 2.6f)) hi = lo - 1.f;
          int l = (int)floorf(lo), h2 = (int)ceilf(hi);
          l = max(l, max(-radius, 1 - px)); h2 = min(h2, min(radius, cols - 2 - px));
          const int r = py + i;
          if (r > 0 && r < rows - 1 && h2 >= l) { jlo = l; cntr = h2 - l + 1; }
        }
        int incl = cntr;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += y;
        }
        // one word per row: (packed index of its first candidate << 8) | (jlo + 128)
        if (ri < side) s_row[wib][ri] = ((unsigned)(carry + incl - cntr) << 8) | (unsigned)(jlo + 128);
        carry += __shfl_sync(0xffffffffu, incl, 31);
      }
      const int ncand = carry;
      if (lane < 3) s_row[wib][side + lane] = lane == 0 ? ((unsigned)ncand << 8) : 0xffffffffu;
      __syncwarp();
      // The row of packed index k is found by walking forward from the previous one; the next two row
      // words stay in registers so the walk never waits on the shared-memory load it has just issued.
      int row = 0;
      unsigned cur = s_row[wib][0], nxt = s_row[wib][1], nx2 = s_row[wib][2];
      for (int k = lane; k < ncand; k += 32) {
        const unsigned kk = ((unsigned)k << 8) | 255u;
        while (kk >= nxt) { cur = nxt; nxt = nx2; ++row; nx2 = s_row[wib][row + 2]; }
        process_clipped(row - radius, (int)(cur & 255u) - 128 + (k - (int)(cur >> 8)));
      }
    } else {
      // very large windows (non-default options): plain scan of the whole window
      int i = lane / side - radius, j = lane % side - radius;
      for (int base = 0; base < total; base += 32) {
        if (base + lane < total) process(i, j);
        j += 32;
        while (j > radius) { j -= side; ++i; }
      }
    }
    __syncwarp();
    // fold the copies and the circular orientation bins, flatten to 128 floats
    for (int e = lane; e < 128; e += 32) {
      const int k = e & 7, cell = e >> 3, ci = cell >> 2, cj = cell & 3;
      const int idx = ((ci + 1) * (D + 2) + (cj + 1)) * (N + 2);
      uint32_t v = 0;
#pragma unroll
      for (int cp = 0; cp < DESC_COPIES; ++cp) {
        v += s_hist[wib][cp * HLEN + idx + k];
        if (k < 2) v += s_hist[wib][cp * HLEN + idx + N + k];
      }
      s_vec[wib][e] = (float)v * SIFT_INV_FIX;
    }
    __syncwarp();
    // the two norms are sequential folds in the oracle's order (k ascending); the values come four per load
    float nrm2 = 0.f;
    const float4* sv4 = reinterpret_cast<const float4*>(s_vec[wib]);
#pragma unroll 4
    for (int q = 0; q < 32; ++q) {
      const float4 v = sv4[q];
      nrm2 = fmaf(v.x, v.x, nrm2); nrm2 = fmaf(v.y, v.y, nrm2); nrm2 = fmaf(v.z, v.z, nrm2); nrm2 = fmaf(v.w, v.w, nrm2);
    }
    const float thr = __fsqrt_rn(nrm2) * 0.2f;
    nrm2 = 0.f;
#pragma unroll 4
    for (int q = 0; q < 32; ++q) {
      const float4 v = sv4[q];
      const float a = fminf(v.x, thr), b2 = fminf(v.y, thr), c2 = fminf(v.z, thr), d2 = fminf(v.w, thr);
      nrm2 = fmaf(a, a, nrm2); nrm2 = fmaf(b2, b2, nrm2); nrm2 = fmaf(c2, c2, nrm2); nrm2 = fmaf(d2, d2, nrm2);
    }
    const float sn = __fsqrt_rn(nrm2);
    const float sc = __fdiv_rn(512.f, fmaxf(sn, FLT_EPSILON));
    float4 o4;
    float* ov = reinterpret_cast<float*>(&o4);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float v = rintf(fminf(s_vec[wib][4 * lane + q], thr) * sc);
      ov[q] = v > 255.f ? 255.f : (v < 0.f ? 0.f : v);
    }
    reinterpret_cast<float4*>(desc + ((size_t)b * kp_cap + ki) * 128)[lane] = o4;
    __syncwarp();
  }
  if (algo_bytes != nullptr && lane == 0 && my_bytes) atomicAdd(algo_bytes + ((blockIdx.x * 4 + wib) & 63), my_bytes);
}

// --------------------------------------------------------------------------- host plumbing
static void gauss_taps(float sigma, Taps* t) {
  const int ksize = ((int)lrint((double)sigma * 8.0 + 1.0)) | 1;
  const int r = ksize / 2;
  double sum = 0, v[64];
  const double s2 = -0.5 / ((double)sigma * (double)sigma);
  for (int i = 0; i < ksize; ++i) {
    const double x = i - (ksize - 1) * 0.5;
    v[i] = std::exp(s2 * x * x);
    sum += v[i];
  }
  sum = 1.0 / sum;
  memset(t, 0, sizeof(*t));
  t->r = r;
  for (int i = 0; i <= r && i <= MAX_R; ++i) t->k[i] = (float)(v[r + i] * sum);
}

static size_t blur_smem(int R) {
  const int in_w = ((TILE_W + 2 * R + 3) & ~3) + 4, in_h = TILE_H + 2 * R;
  return (size_t)(in_h * in_w + in_h * TILE_W) * sizeof(float);
}

template <int RT, bool U8>
static int launch_blur_t(const float* src, const uint8_t* src8, float* dst, float* dog, int h, int w, int pitch,
                         int rows8, int cols8, int batch, const Taps& t, cudaStream_t st) {
  const size_t smem = blur_smem(t.r);
  VO_TRY(ensure_dyn_smem_of(sift_blur_dog_kernel<RT, U8>, blur_smem(RT > 0 ? RT : MAX_R)));
  dim3 grid(div_up(w, TILE_W), div_up(h, TILE_H), batch);
  sift_blur_dog_kernel<RT, U8><<<grid, 256, smem, st>>>(src, src8, dst, dog, h, w, pitch, rows8, cols8, t);
  return VO_OK;
}

template <int R, int MODE>
static int launch_tma_t2(const CUtensorMap& tm, int z_base, const float* src, float* dst, float* dog, int h, int w, int pitch,
                         int batch, const Taps& t, int num_sms, cudaStream_t st, const DsOut& ds) {
  const int strips = div_up(w, TS_W);
  int n_seg = 1;
  const int target = num_sms * 6;
  if (strips * batch < target) n_seg = div_up(target, strips * batch);
  int seg_rows = div_up(div_up(h, n_seg), TS_G) * TS_G;
  if (seg_rows < 4 * TS_G) seg_rows = 4 * TS_G;
  n_seg = div_up(h, seg_rows);
  dim3 grid(strips, n_seg, batch);
  constexpr int smem = 2 * TS_G * TS_BOXW * 4 + (MODE == 1 ? 8 * TS_BOXW * 4 : 0) + (64 + TS_MIRROR(R) + (MODE == 3 ? 4 : 0)) * TS_RS * 4 + 64;
  if (ds.dst != nullptr) {
    VO_TRY(ensure_dyn_smem_of(sift_blur_tma_kernel<R, MODE, true>, smem));
    sift_blur_tma_kernel<R, MODE, true><<<grid, 256, smem, st>>>(tm, z_base, dst, dog, src, h, w, pitch, seg_rows, t, ds);
  } else {
    VO_TRY(ensure_dyn_smem_of(sift_blur_tma_kernel<R, MODE, false>, smem));
    sift_blur_tma_kernel<R, MODE, false><<<grid, 256, smem, st>>>(tm, z_base, dst, dog, src, h, w, pitch, seg_rows, t, ds);
  }
  return VO_OK;
}
// VO_BLUR_ROW selects the row pass: 2 (default) = eight outputs per lane, 0 = four outputs per lane (round 1),
// 1 = packed f32x2 (measured slower: 3.27 vs 2.45 ms per step, it reads more shared memory per pixel; kept for A/B runs)
// 3 = as 2 with a column pass of eight rows x one column per thread (scalar; fewer shared bytes, more issue slots: 2.47 vs 2.36)
template <int R>
static int launch_tma_t(const CUtensorMap& tm, int z_base, const float* src, float* dst, float* dog, int h, int w, int pitch,
                        int batch, const Taps& t, int num_sms, cudaStream_t st, const DsOut& ds) {
  static const int mode = [] { const char* e = getenv("VO_BLUR_ROW"); return e ? atoi(e) : 2; }();
  if (mode == 1) return launch_tma_t2<R, 1>(tm, z_base, src, dst, dog, h, w, pitch, batch, t, num_sms, st, ds);
  if (mode == 0) return launch_tma_t2<R, 0>(tm, z_base, src, dst, dog, h, w, pitch, batch, t, num_sms, st, ds);
  if (mode == 3) return launch_tma_t2<R, 3>(tm, z_base, src, dst, dog, h, w, pitch, batch, t, num_sms, st, ds);
  return launch_tma_t2<R, 2>(tm, z_base, src, dst, dog, h, w, pitch, batch, t, num_sms, st, ds);
}

// tm/z_base: tensor map of the octave's Gaussian stack and the z index of image 0 of the source layer
// *ds_done is set when the launched kernel also wrote the next octave's seed (ds.dst != nullptr on the TMA path)
static int launch_blur(const CUtensorMap& tm, int z_base, const float* src, float* dst, float* dog, int h, int w, int pitch,
                       int batch, const Taps& t, int num_sms, cudaStream_t st, const DsOut& ds = DsOut{nullptr, 0, 0, 0},
                       bool* ds_done = nullptr) {
  if (h >= 64 && w >= 96) {
    switch (t.r) {
      case 5: if (ds_done) *ds_done = ds.dst != nullptr; return launch_tma_t<5>(tm, z_base, src, dst, dog, h, w, pitch, batch, t, num_sms, st, ds);
      case 6: if (ds_done) *ds_done = ds.dst != nullptr; return launch_tma_t<6>(tm, z_base, src, dst, dog, h, w, pitch, batch, t, num_sms, st, ds);
      case 8: if (ds_done) *ds_done = ds.dst != nullptr; return launch_tma_t<8>(tm, z_base, src, dst, dog, h, w, pitch, batch, t, num_sms, st, ds);
      case 10: if (ds_done) *ds_done = ds.dst != nullptr; return launch_tma_t<10>(tm, z_base, src, dst, dog, h, w, pitch, batch, t, num_sms, st, ds);
      case 13: if (ds_done) *ds_done = ds.dst != nullptr; return launch_tma_t<13>(tm, z_base, src, dst, dog, h, w, pitch, batch, t, num_sms, st, ds);
      default: break;
    }
  }
  switch (t.r) {
    case 5: return launch_blur_t<5, false>(src, nullptr, dst, dog, h, w, pitch, 0, 0, batch, t, st);
    case 6: return launch_blur_t<6, false>(src, nullptr, dst, dog, h, w, pitch, 0, 0, batch, t, st);
    case 8: return launch_blur_t<8, false>(src, nullptr, dst, dog, h, w, pitch, 0, 0, batch, t, st);
    case 10: return launch_blur_t<10, false>(src, nullptr, dst, dog, h, w, pitch, 0, 0, batch, t, st);
    case 13: return launch_blur_t<13, false>(src, nullptr, dst, dog, h, w, pitch, 0, 0, batch, t, st);
    default: return launch_blur_t<0, false>(src, nullptr, dst, dog, h, w, pitch, 0, 0, batch, t, st);
  }
}

static void fill_sift_opts(const vo_sift_opts* in, vo_sift_opts* o) {
  o->contrast_threshold = 0.04f / 3.0f; o->edge_threshold = 10.f; o->num_layers_in_octave = 3; o->sigma = 1.6f; o->index_base = 0;
  if (in) {
    if (in->contrast_threshold > 0) o->contrast_threshold = in->contrast_threshold;
    if (in->edge_threshold > 0) o->edge_threshold = in->edge_threshold;
    if (in->num_layers_in_octave > 0) o->num_layers_in_octave = in->num_layers_in_octave;
    if (in->sigma > 0) o->sigma = in->sigma;
    o->index_base = in->index_base;
  }
}

static int get_plan(vo_ctx* ctx, int rows, int cols, int batch, const vo_sift_opts& o, int capacity, SiftPlan** out) {
  SiftPlan* p = ctx->sift_plan;
  const int nl = o.num_layers_in_octave;
  // candidate words pack y in 12 bits and x in 13 bits of the 2x base image: every entry point (vo_sift,
  // vo_sift_batch, vo_frames*) comes through here, so the limit is enforced once
  VO_CHECK_ARG(rows > 0 && cols > 0 && rows <= 4095 / 2 && cols <= 8191 / 2, "image too large (max 2047 x 4095)");
  int kp_cap = ((capacity > 4096 ? capacity : 4096) + 1023) / 1024 * 1024;
  if (p && p->rows == rows && p->cols == cols && p->batch >= batch && p->nl == nl && p->sigma == o.sigma && p->kp_cap >= kp_cap) {
    *out = p; return VO_OK;
  }
  if (p) { VO_CUDA(cudaStreamSynchronize(ctx->stream)); sift_plan_destroy(p); ctx->sift_plan = nullptr; }
  ++ctx->alloc_generation;   // a new plan: new buffer addresses and tensor maps
  if (nl > 5) { set_error("vo_sift: NumLayersInOctave > 5 is not supported"); return VO_ERR_ARG; }
  p = new SiftPlan();
  p->rows = rows; p->cols = cols; p->batch = batch; p->nl = nl; p->sigma = o.sigma;
  const int R2 = rows * 2, C2 = cols * 2;
  int n_oct = (int)lrint(std::log((double)(R2 < C2 ? R2 : C2)) / std::log(2.0) - 2.0) + 1;
  if (n_oct < 1) n_oct = 1;
  if (n_oct > MAX_OCT) n_oct = MAX_OCT;
  size_t gtot = 0, dtot = 0;
  int used = 0;
  for (int oc = 0; oc < n_oct; ++oc) {
    p->h[oc] = oc == 0 ? R2 : p->h[oc - 1] / 2;
    p->w[oc] = oc == 0 ? C2 : p->w[oc - 1] / 2;
    if (p->h[oc] < 1 || p->w[oc] < 1) break;
    p->pitch[oc] = (p->w[oc] + 31) / 32 * 32;
    p->goff[oc] = gtot; p->doff[oc] = dtot;
    gtot += (size_t)(nl + 3) * p->layer_elems(oc);
    dtot += (size_t)(nl + 2) * p->layer_elems(oc);
    used = oc + 1;
  }
  p->n_oct = used;
  const double k = std::pow(2.0, 1.0 / nl);
  for (int i = 1; i < nl + 3; ++i) {
    const double sp = std::pow(k, (double)(i - 1)) * o.sigma, stt = sp * k;
    gauss_taps((float)std::sqrt(stt * stt - sp * sp), &p->taps[i]);
    if (p->taps[i].r > MAX_R) { set_error("vo_sift: Sigma too large for the blur kernels"); sift_plan_destroy(p); return VO_ERR_ARG; }
  }
  const float sd2 = o.sigma * o.sigma - 0.5f * 0.5f * 4.0f;
  gauss_taps(sqrtf(sd2 > 0.01f ? sd2 : 0.01f), &p->base_taps);
  if (p->base_taps.r > MAX_R) { set_error("vo_sift: Sigma too large"); sift_plan_destroy(p); return VO_ERR_ARG; }
  p->kp_cap = kp_cap; p->cand_cap = kp_cap * 8;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes); };
  A((void**)&p->gauss, (gtot + 256) * sizeof(float)); A((void**)&p->dog, dtot * sizeof(float));
  A((void**)&p->img, (size_t)batch * rows * cols); A((void**)&p->img_t, (size_t)batch * rows * cols);
  A((void**)&p->cand, (size_t)batch * p->cand_cap * sizeof(uint32_t)); A((void**)&p->counters, (size_t)batch * 4 * sizeof(int));
  A((void**)&p->work, (size_t)batch * 2 * sizeof(int));
  A((void**)&p->raw, (size_t)batch * kp_cap * sizeof(vo_keypoint)); A((void**)&p->sorted, (size_t)batch * kp_cap * sizeof(vo_keypoint));
  A((void**)&p->final_kp, (size_t)batch * kp_cap * sizeof(vo_keypoint)); A((void**)&p->desc, (size_t)batch * kp_cap * 128 * sizeof(float));
  A((void**)&p->trig, (size_t)batch * kp_cap * sizeof(float2));
  A((void**)&p->refined, (size_t)batch * p->cand_cap * 24);
  if (e != cudaSuccess) { set_error("vo_sift: device allocation failed: %s", cudaGetErrorString(e)); sift_plan_destroy(p); return VO_ERR_CUDA; }
  {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) { set_error("cuTensorMapEncodeTiled driver entry point not available"); sift_plan_destroy(p); return VO_ERR_CUDA; }
    for (int oc = 0; oc < p->n_oct; ++oc) {
      cuuint64_t gdim[3] = {(cuuint64_t)p->w[oc], (cuuint64_t)p->h[oc], (cuuint64_t)(nl + 3) * batch};
      cuuint64_t gstr[2] = {(cuuint64_t)p->pitch[oc] * 4, (cuuint64_t)p->h[oc] * p->pitch[oc] * 4};
      cuuint32_t box[3] = {(cuuint32_t)TS_BOXW, (cuuint32_t)TS_G, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = enc(&p->tm_gauss[oc], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)(p->gauss + p->goff[oc]), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("vo_sift: cuTensorMapEncodeTiled failed (%d) for octave %d", (int)r, oc); sift_plan_destroy(p); return VO_ERR_CUDA; }
      cuuint64_t ddim[3] = {(cuuint64_t)p->w[oc], (cuuint64_t)p->h[oc], (cuuint64_t)(nl + 2) * batch};
      cuuint32_t dbox[3] = {(cuuint32_t)XT_BOXW, (cuuint32_t)XT_G, 1};
      r = enc(&p->tm_dog[oc], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)(p->dog + p->doff[oc]), ddim, gstr, dbox, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("vo_sift: cuTensorMapEncodeTiled failed (%d) for the DoG stack of octave %d", (int)r, oc); sift_plan_destroy(p); return VO_ERR_CUDA; }
    }
  }
  ctx->sift_plan = p; *out = p;
  return VO_OK;
}

// Runs the whole detector+descriptor on p->img (device, batch images) -> p->final_kp / p->desc / counters.
int sift_run_device(vo_ctx* ctx, SiftPlan* p, int batch, const vo_sift_opts& o, cudaStream_t st) {
  const int nl = p->nl;
  VO_CUDA(cudaMemsetAsync(p->counters, 0, (size_t)batch * 4 * sizeof(int), st));
  VO_CUDA(cudaMemsetAsync(p->work, 0, (size_t)batch * 2 * sizeof(int), st));
  // pyramid
  ProfScope* ps_base = new ProfScope(ctx, st, "sift_base_upsample_blur", (double)batch * ((double)p->rows * p->cols + (double)p->h[0] * p->w[0] * 4.0));
  if (p->base_taps.r == 5 && p->h[0] >= 64) {
    const int strips = div_up(p->w[0], ST_W);
    int n_seg = 1;
    const int target = ctx->num_sms * 8;
    if (strips * batch < target) n_seg = div_up(target, strips * batch);
    int seg_rows = div_up(div_up(p->h[0], n_seg), ST_CH) * ST_CH;
    if (seg_rows < 4 * ST_CH) seg_rows = 4 * ST_CH;
    n_seg = div_up(p->h[0], seg_rows);
    sift_base_stream_kernel<5><<<dim3(strips, n_seg, batch), 256, 0, st>>>(p->img, p->G(0, 0), p->h[0], p->w[0], p->pitch[0], p->rows, p->cols, seg_rows, p->base_taps);
  } else if (p->base_taps.r == 5)
    VO_TRY((launch_blur_t<5, true>(nullptr, p->img, p->G(0, 0), nullptr, p->h[0], p->w[0], p->pitch[0], p->rows, p->cols, batch, p->base_taps, st)));
  else
    VO_TRY((launch_blur_t<0, true>(nullptr, p->img, p->G(0, 0), nullptr, p->h[0], p->w[0], p->pitch[0], p->rows, p->cols, batch, p->base_taps, st)));
  delete ps_base;
  // note: layer_elems uses p->batch; kernels index images with the plan's batch stride
  OctInfo oi;
  memset(&oi, 0, sizeof(oi));
  for (int oc = 0; oc < p->n_oct; ++oc) { oi.h[oc] = p->h[oc]; oi.w[oc] = p->w[oc]; oi.pitch[oc] = p->pitch[oc]; oi.goff[oc] = p->goff[oc]; oi.doff[oc] = p->doff[oc]; }
  // octaves from first_small on go to the fused shared-memory kernel (when one octave fits: 3 buffers)
  int first_small = p->n_oct;
  for (int oc = 0; oc < p->n_oct; ++oc)
    if (!(p->h[oc] >= 64 && p->w[oc] >= 96)) { first_small = oc; break; }
  const int small_cap = first_small < p->n_oct ? p->h[first_small] * p->w[first_small] : 0;
  const bool fuse_small = first_small < p->n_oct && (size_t)small_cap * 12 <= 200 * 1024;
  bool seeded = false;   // layer 0 of the octave about to be processed was written by the previous octave's blur
  for (int oc = 0; oc < (fuse_small ? first_small : p->n_oct); ++oc) {
    const double px = (double)batch * p->h[oc] * p->w[oc];
    if (oc > 0 && !seeded) {
      ProfScope ps(ctx, st, "sift_downsample", px * 8.0);
      dim3 g(div_up(p->w[oc], 256), p->h[oc], batch);
      sift_downsample_kernel<<<g, 256, 0, st>>>(p->G(oc - 1, nl), p->G(oc, 0), p->h[oc - 1], p->pitch[oc - 1], p->h[oc], p->w[oc], p->pitch[oc]);
    }
    seeded = false;
    const bool tma = p->h[oc] >= 64 && p->w[oc] >= 96;   // same test as launch_blur
    const char* nm = tma ? (oc == 0 ? "sift_blur_dog_tma_oct0" : "sift_blur_dog_tma_oct1+") : "sift_blur_dog_small";
    // the next octave's layer 0 is the 2x decimation of this octave's layer nl: the blur that produces that layer writes it
    // too (unless the next octave belongs to the shared-memory kernel below, which decimates for itself)
    const bool next_here = oc + 1 < (fuse_small ? first_small : p->n_oct);
    for (int i = 1; i < nl + 3; ++i) {
      ProfScope ps(ctx, st, nm, px * 12.0);   // read G[l], write G[l+1], write D[l]
      DsOut ds{nullptr, 0, 0, 0};
      if (i == nl && next_here) ds = DsOut{p->G(oc + 1, 0), p->pitch[oc + 1], p->h[oc + 1], p->w[oc + 1]};
      bool done = false;
      VO_TRY(launch_blur(p->tm_gauss[oc], (i - 1) * p->batch, p->G(oc, i - 1), p->G(oc, i), p->D(oc, i - 1), p->h[oc], p->w[oc], p->pitch[oc], batch, p->taps[i], ctx->num_sms, st, ds, &done));
      if (i == nl) seeded = done;
    }
  }
  if (fuse_small) {
    double px = 0;
    for (int oc = first_small; oc < p->n_oct; ++oc) px += (double)batch * p->h[oc] * p->w[oc];
    TapsAll ta;
    memset(&ta, 0, sizeof(ta));
    for (int i = 1; i < nl + 3; ++i) ta.t[i] = p->taps[i];
    const size_t smem = (size_t)small_cap * 12;
    VO_TRY(ensure_dyn_smem_of(sift_small_octaves_kernel, smem));
    ProfScope ps(ctx, st, "sift_blur_dog_small", px * (8.0 + 12.0 * (nl + 2)));
    sift_small_octaves_kernel<<<batch, 1024, smem, st>>>(p->gauss, p->dog, oi, first_small, p->n_oct, nl, p->batch, small_cap, ta);
  }
  VO_CUDA(cudaGetLastError());
  // extrema
  const float contrast_cv = o.contrast_threshold * nl;   // OpenCV-style threshold (0.04)
  const float threshold = (float)(int)std::floor(0.5 * contrast_cv / nl * 255.0);
  for (int oc = 0; oc < p->n_oct; ++oc) {
    if (p->h[oc] <= 2 * SIFT_BORDER || p->w[oc] <= 2 * SIFT_BORDER) continue;
    dim3 g(div_up(div_up(p->w[oc], EX_COLS), 4), div_up(p->h[oc], EX_ROWS), batch);
    ProfScope ps(ctx, st, "sift_extrema", (double)batch * p->h[oc] * p->w[oc] * 4.0 * (nl + 2));
    // VO_EXT_TMA=1: the TMA-fed form (measured slower, 1.26 vs 1.03 ms per step: the test is bound by instruction issue,
    // 71 % issue-active under ncu, not by bytes in flight; the block barriers of the staged form cost more than its loads save)
    static const bool ext_tma = [] { const char* e = getenv("VO_EXT_TMA"); return e ? atoi(e) != 0 : false; }();
    if (ext_tma && nl == 3 && p->w[oc] >= XT_COLS && p->h[oc] >= 64) {
      // row segments: enough blocks for a few waves of 3 per SM, at least 64 rows each
      const int strips = div_up(p->w[oc], XT_COLS);
      int n_seg = div_up(ctx->num_sms * 3 * 3, strips * batch);
      if (n_seg < 1) n_seg = 1;
      int seg_rows = div_up(p->h[oc], n_seg);
      if (seg_rows < 64) seg_rows = 64;
      n_seg = div_up(p->h[oc], seg_rows);
      constexpr int smem = XT_NST * 5 * XT_G * XT_BOXW * 4 + 64;
      VO_TRY(ensure_dyn_smem_of(sift_extrema_tma_kernel<3>, smem));
      sift_extrema_tma_kernel<3><<<dim3(strips, n_seg, batch), XT_WARPS * 32, smem, st>>>(p->tm_dog[oc], oc, p->batch, p->h[oc], p->w[oc], seg_rows,
                                                                                       threshold, p->cand, p->cand_cap, p->counters);
      continue;
    }
#define VO_EXTREMA(NLV) sift_extrema_kernel<NLV><<<g, 128, 0, st>>>(p->D(oc, 0), oc, p->batch, p->h[oc], p->w[oc], p->pitch[oc], threshold, p->cand, p->cand_cap, p->counters)
    switch (nl) {
      case 1: VO_EXTREMA(1); break;
      case 2: VO_EXTREMA(2); break;
      case 3: VO_EXTREMA(3); break;
      case 4: VO_EXTREMA(4); break;
      default: VO_EXTREMA(5); break;
    }
#undef VO_EXTREMA
  }
  {
    dim3 g(ctx->num_sms * 4 / (batch > 4 ? 4 : 1), batch);
    ProfScope ps(ctx, st, "sift_refine_orient", 0.0, 0.0, 2);
    sift_refine_kernel<<<dim3(div_up(p->cand_cap / 4, 128), batch), 128, 0, st>>>(p->dog, oi, p->batch, nl, contrast_cv, o.edge_threshold, o.sigma, p->cand, p->cand_cap, p->counters, p->refined);
    sift_orient_kernel<<<g, 128, 0, st>>>(p->gauss, oi, p->batch, p->refined, p->cand_cap, p->counters, p->raw, p->kp_cap, p->work);
  }
  {
    dim3 g(p->kp_cap / 256, batch);
    ProfScope ps(ctx, st, "sift_sort_dedupe", 0.0, 0.0, 2);
    if (p->kp_cap <= 65536 && p->w[0] > 0) {
      const size_t smem = (size_t)(2 * SORT_BUCKETS + 1) * sizeof(int) + (size_t)p->kp_cap * sizeof(unsigned short);
      VO_TRY(ensure_dyn_smem_of(sift_rank_bucket_kernel, smem));
      // raw keypoint x is in base-image (2x) pixels: [0, w[0])
      sift_rank_bucket_kernel<<<batch, 1024, smem, st>>>(p->raw, p->kp_cap, p->counters, p->sorted, (float)SORT_BUCKETS / (float)p->w[0]);
    } else {
      sift_rank_kernel<<<g, 256, 0, st>>>(p->raw, p->kp_cap, p->counters, p->sorted);
    }
    sift_dedupe_kernel<<<batch, 1024, 0, st>>>(p->sorted, p->kp_cap, p->counters, p->final_kp, (float)o.index_base);
  }
  {
    dim3 g(ctx->num_sms * 4 / (batch > 4 ? 4 : 1), batch);
    // when profiling, the kernel also accumulates its data-dependent algorithmic bytes (SURVEY 8d:
    // (2r+1)^2*4 + 512 per keypoint) into 64 spread counters owned by the context
    unsigned long long* ab = nullptr;
    if (ctx->prof_enabled) {
      VO_TRY(dev_buf(ctx, "prof_desc_bytes", 64, &ab));
      ctx->prof_late_stage = ctx->prof_stage_id("sift_descriptor");
    }
    ProfScope ps(ctx, st, "sift_descriptor", 0.0, 0.0, 2);
    sift_trig_kernel<<<dim3(8, batch), 256, 0, st>>>(p->final_kp, p->kp_cap, p->counters, p->trig);
    sift_descriptor_kernel<<<g, 128, 0, st>>>(p->gauss, oi, p->batch, nl, p->final_kp, p->trig, p->kp_cap, p->counters, (float)o.index_base, p->desc, ab, p->work);
  }
  VO_CUDA(cudaGetLastError());
  return VO_OK;
}

// accessors used by the frame pipeline
int sift_prepare(vo_ctx* ctx, int rows, int cols, int batch, const vo_sift_opts* opts, int capacity, SiftPlan** plan, vo_sift_opts* filled) {
  fill_sift_opts(opts, filled);
  return get_plan(ctx, rows, cols, batch, *filled, capacity, plan);
}
uint8_t* sift_plan_images(SiftPlan* p) { return p->img; }
// n MATLAB-ordered images (contiguous, rows*cols bytes each) -> plan images first_img, first_img + img_step, ...
int sift_load_col_major(vo_ctx* ctx, SiftPlan* p, int first_img, int img_step, const uint8_t* src, int n, bool on_device, cudaStream_t st) {
  const size_t ib = (size_t)p->rows * p->cols;
  const uint8_t* stage = src;
  if (!on_device) {
    VO_TRY(upload_2d(ctx, p->img_t + (size_t)first_img * ib, (size_t)img_step * ib, src, ib, ib, n, st));
    stage = p->img_t + (size_t)first_img * ib;
  }
  const size_t sstep = on_device ? ib : (size_t)img_step * ib;
  dim3 g(div_up(p->cols, 32), div_up(p->rows, 32), n);
  sift_transpose_u8_kernel<<<g, dim3(32, 8), 0, st>>>(stage, p->img + (size_t)first_img * ib, p->rows, p->cols, p->rows, sstep, (size_t)img_step * ib);
  return VO_OK;
}
vo_keypoint* sift_plan_keypoints(SiftPlan* p) { return p->final_kp; }
float* sift_plan_desc(SiftPlan* p) { return p->desc; }
int* sift_plan_counters(SiftPlan* p) { return p->counters; }
int sift_plan_kp_cap(SiftPlan* p) { return p->kp_cap; }

}  // namespace vo

using namespace vo;

// dst[k*n + i] = src[i*128 + k]: one image's descriptors as a MATLAB M x 128 matrix
__global__ void __launch_bounds__(256)
sift_desc_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  __shared__ float t[32][33];
  const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) if (i0 + r < n) t[r][threadIdx.x] = src[(size_t)(i0 + r) * 128 + k0 + threadIdx.x];
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) if (i0 + threadIdx.x < n) dst[(size_t)(k0 + r) * n + i0 + threadIdx.x] = t[threadIdx.x][r];
}

static int sift_host(vo_ctx* ctx, const uint8_t* imgs, int n_img, int rows, int cols, int ld, int col_major,
                     const vo_sift_opts* opts, int capacity, vo_keypoint* kps, float* desc, int* n_out, int desc_col_major = 0) {
  VO_CHECK_ARG(ctx && imgs && n_out, "null argument");
  VO_CHECK_ARG(n_img > 0 && rows > 0 && cols > 0 && capacity >= 0, "bad size");
  VO_CHECK_ARG(capacity == 0 || (kps && desc), "null output");
  VO_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  vo_sift_opts o; SiftPlan* p;
  VO_TRY(sift_prepare(ctx, rows, cols, n_img, opts, capacity, &p, &o));
  if (col_major) {
    for (int b = 0; b < n_img; ++b)
      VO_CUDA(cudaMemcpy2DAsync(p->img_t + (size_t)b * rows * cols, rows, imgs + (size_t)b * ld * cols, ld, rows, cols, cudaMemcpyHostToDevice, st));
    dim3 g(div_up(cols, 32), div_up(rows, 32), n_img);
    sift_transpose_u8_kernel<<<g, dim3(32, 8), 0, st>>>(p->img_t, p->img, rows, cols, rows, (size_t)rows * cols, (size_t)rows * cols);
  } else if (ld == cols) {
    VO_CUDA(cudaMemcpyAsync(p->img, imgs, (size_t)n_img * rows * cols, cudaMemcpyHostToDevice, st));
  } else {
    for (int b = 0; b < n_img; ++b)
      VO_CUDA(cudaMemcpy2DAsync(p->img + (size_t)b * rows * cols, cols, imgs + (size_t)b * ld * rows, ld, cols, rows, cudaMemcpyHostToDevice, st));
  }
  VO_TRY(sift_run_device(ctx, p, n_img, o, st));
  int* hc; VO_TRY(pin_buf(ctx, "sift_counts", (size_t)n_img * 4, &hc));
  VO_CUDA(cudaMemcpyAsync(hc, p->counters, (size_t)n_img * 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  VO_CUDA(cudaStreamSynchronize(st));
  int rc = VO_OK;
  for (int b = 0; b < n_img; ++b) {
    int n = hc[b * 4 + 2];
    if (hc[b * 4 + 0] > p->cand_cap || hc[b * 4 + 1] > p->kp_cap) {
      set_error("vo_sift: internal capacity exceeded (candidates %d/%d, keypoints %d/%d); raise `capacity`", hc[b * 4 + 0], p->cand_cap, hc[b * 4 + 1], p->kp_cap);
      rc = VO_ERR_CAPACITY;
    }
    n_out[b] = n;
    if (n > capacity) { if (rc == VO_OK) set_error("vo_sift: %d keypoints exceed capacity %d", n, capacity); rc = VO_ERR_CAPACITY; n = capacity; }
    if (n > 0) {
      VO_CUDA(cudaMemcpyAsync(kps + (size_t)b * capacity, p->final_kp + (size_t)b * p->kp_cap, (size_t)n * sizeof(vo_keypoint), cudaMemcpyDeviceToHost, st));
      const float* dsrc = p->desc + (size_t)b * p->kp_cap * 128;
      if (desc_col_major) {
        float* dt; VO_TRY(dev_buf(ctx, "sift_desc_t", (size_t)n_img * p->kp_cap * 128, &dt));
        dt += (size_t)b * p->kp_cap * 128;
        sift_desc_transpose_kernel<<<dim3(div_up(n, 32), 4), dim3(32, 8), 0, st>>>(dsrc, dt, n);
        ++ctx->kernel_launches;
        dsrc = dt;
      }
      VO_CUDA(cudaMemcpyAsync(desc + (size_t)b * capacity * 128, dsrc, (size_t)n * 128 * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
  }
  VO_CUDA(cudaStreamSynchronize(st));
  return rc;
}

extern "C" {

int vo_sift(vo_ctx* ctx, const uint8_t* img, int rows, int cols, int ld, int col_major, const vo_sift_opts* opts,
            int capacity, vo_keypoint* kps, float* desc, int* n_out) {
  return sift_host(ctx, img, 1, rows, cols, ld, col_major, opts, capacity, kps, desc, n_out);
}

int vo_sift_batch(vo_ctx* ctx, const uint8_t* imgs, int n_img, int rows, int cols, const vo_sift_opts* opts,
                  int capacity, vo_keypoint* kps, float* desc, int* n_out) {
  return sift_host(ctx, imgs, n_img, rows, cols, cols, 0, opts, capacity, kps, desc, n_out);
}

int vo_sift_stack(vo_ctx* ctx, const uint8_t* imgs, int n_img, int rows, int cols, int col_major, const vo_sift_opts* opts,
                  int capacity, vo_keypoint* kps, float* desc, int desc_col_major, int* n_out) {
  return sift_host(ctx, imgs, n_img, rows, cols, col_major ? rows : cols, col_major, opts, capacity, kps, desc, n_out, desc_col_major);
}

}  // extern "C"
