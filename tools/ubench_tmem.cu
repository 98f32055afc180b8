// ubench_tmem.cu -- microbenchmarks that bound the match GEMM epilogue on sm_100a:
//   (1) tcgen05.ld throughput per SM as a function of shape and warp count,
//   (2) tcgen05.mma issue rate for M=128 with N=128 / N=256 operands in shared memory (SS mode),
//   (3) both at once (MMA writing one accumulator stage while warps drain the other).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_tmem tools/ubench_tmem.cu
// Numbers printed are cycles (clock64) per SM; one CTA per SM, 148 CTAs.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <algorithm>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc, int i8) {
  if (i8)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

#define LD_X32(shape_str)                                                                                   \
  asm volatile("tcgen05.ld.sync.aligned." shape_str ".b32 "                                                 \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                                    \
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                   \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),        \
                 "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),    \
                 "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), \
                 "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), \
                 "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                         \
               : "r"(taddr))

// mode 0: 32x32b.x32 (32 lanes x 32 columns), mode 1: 16x256b.x4 (16 lanes x 64 columns... 32 regs),
// mode 2: 32x32b.x32 issued twice before one wait (two loads in flight)
template <int MODE>
__device__ __forceinline__ uint32_t ld_loop(uint32_t tmem_base, int warp, int iters) {
  uint32_t r[32];
  uint32_t sink = 0;
  const uint32_t lane_q = (uint32_t)((warp & 3) * 32) << 16;
  for (int it = 0; it < iters; ++it) {
    const uint32_t taddr = tmem_base + lane_q + (uint32_t)(((it + (warp >> 2)) * 32) & 511 & ~31);
    if (MODE == 0) { LD_X32("32x32b.x32"); }
    else if (MODE == 1) { LD_X32("16x256b.x8"); }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < 32; ++c) sink ^= r[c];
  }
  return sink;
}

// kernel 1: LDTM only.  nwarps warps loop over the 512 columns.
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_ldtm(int nwarps, int iters, long long* cycles, uint32_t* sinkp) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_slot;
  __syncthreads();
  const long long t0 = clock64();
  uint32_t sink = 0;
  if (warp < nwarps) sink = ld_loop<MODE>(tb, warp, iters);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (sink == 0x12345678u) sinkp[0] = sink;
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
  }
}

// kernel 2: MMA (warp 0, one lane) with optional concurrent LDTM by warps 4..4+nld-1.
// n_mma back-to-back groups of 8 MMAs (K = 128 bf16, or K = 256 i8) of shape M=128 x N=bn.
__global__ void __launch_bounds__(640, 1) k_mma(int bn, int groups, int nld, int ld_iters, int i8, long long* cycles, long long* ld_cycles, uint32_t* sinkp) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) unsigned long long bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 * 2 + 32768 * 2) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    if (lane == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_slot;
  const uint32_t sA = smem_u32(smem), sB = sA + 32768;
  const long long t0 = clock64();
  uint32_t sink = 0;
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t idesc = i8 ? ((2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))
                                : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24));
      for (int g = 0; g < groups; ++g) {
        const uint32_t d = tb + (uint32_t)((g & 1) * 256);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma(d, desc_sw128(sA + kb * 16384 + k * 32), desc_sw128(sB + kb * 32768 + k * 32), idesc, (kb | k) != 0, i8);
      }
      tc_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 4 + nld) {
    const long long l0 = clock64();
    sink = ld_loop<0>(tb, warp - 4, ld_iters);
    const long long l1 = clock64();
    if (lane == 0 && warp == 4) ld_cycles[blockIdx.x] = l1 - l0;
  }
  const long long t1w = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1w - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (sink == 0x12345678u) sinkp[0] = sink;
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
  }
}

static double median(std::vector<long long> v) {
  std::sort(v.begin(), v.end());
  return (double)v[v.size() / 2];
}
#include <algorithm>

int main() {
  const int G = 148;
  long long *cyc, *lcyc; uint32_t* sink;
  cudaMalloc(&cyc, G * 8); cudaMalloc(&lcyc, G * 8); cudaMalloc(&sink, 4);
  std::vector<long long> h(G), hl(G);
  const int iters = 2000;
  for (int mode = 0; mode < 1; ++mode)
    for (int nw : {1, 4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k_ldtm<0><<<G, 512>>>(nw, iters, cyc, sink); else k_ldtm<1><<<G, 512>>>(nw, iters, cyc, sink);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h.data(), cyc, G * 8, cudaMemcpyDeviceToHost);
      const double c = median(h);
      printf("ldtm mode=%d (%s) warps=%2d: %.0f cyc for %d x 4 KB loads/warp -> %.1f B/cyc/SM, %.1f cyc/load/warp  [%s]\n", mode,
             mode == 0 ? "32x32b.x32" : "16x256b.x8", nw, c, iters, (double)nw * iters * 4096.0 / c, c / iters,
             cudaGetErrorString(cudaGetLastError()));
    }
  cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int groups = 400;
  for (int i8 = 0; i8 < 2; ++i8)
    for (int bn : {128, 256})
      for (int nld : {0, 4, 8, 16}) {
        // ld_iters sized so the drains roughly cover the MMA time at 64 B/cyc/SM
        const int ld_iters = 1200;
        for (int rep = 0; rep < 2; ++rep) {
          k_mma<<<G, 640, 98 * 1024>>>(bn, groups, nld, ld_iters, i8, cyc, lcyc, sink);
          cudaDeviceSynchronize();
        }
        cudaMemcpy(h.data(), cyc, G * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(hl.data(), lcyc, G * 8, cudaMemcpyDeviceToHost);
        const double c = median(h), cl = nld ? median(hl) : 0.0;
        printf("mma kind=%s M=128 N=%3d: %d groups x 8 MMAs in %.0f cyc -> %.1f cyc/MMA (ideal %d)", i8 ? "i8 " : "f16", bn, groups, c,
               c / (groups * 8.0), i8 ? bn / 2 * 1 : bn / 2);
        if (nld) printf("; concurrent ldtm warps=%2d: %.0f cyc for %d loads/warp -> %.1f B/cyc/SM", nld, cl, ld_iters, nld * ld_iters * 4096.0 / cl);
        printf("  [%s]\n", cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
