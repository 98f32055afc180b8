// vo_internal.h -- shared internals of libvo_b200.so (context, error handling, scratch buffers).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <vector>
#include <map>
#include <string>
#include "../../include/vo_b200.h"

namespace vo {

void set_error(const char* fmt, ...);

#define VO_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      vo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return VO_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define VO_CHECK_ARG(cond, msg)                      \
  do {                                               \
    if (!(cond)) {                                   \
      vo::set_error("bad argument: %s", msg);        \
      return VO_ERR_ARG;                             \
    }                                                \
  } while (0)

#define VO_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != VO_OK) return _r; \
  } while (0)

// A named, grow-only device (or pinned-host) scratch buffer owned by the context.
struct Scratch {
  void* ptr = nullptr;
  size_t bytes = 0;
  bool host = false;
};

// Optional stage profiler: when enabled, every stage launch is bracketed by CUDA events on the
// launching stream; totals (ms, launches, algorithmic bytes / flops) are read back per stage name.
struct ProfRec { int stage; cudaEvent_t e0, e1; };
struct ProfStage { std::string name; double ms = 0; long long launches = 0; double bytes = 0; double flops = 0; };


struct SiftPlan;   // vo_sift.cu
struct FramePlan;  // vo_frames.cu

}  // namespace vo

struct vo_ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  std::map<std::string, vo::Scratch> scratch;
  int match_stats[4] = {0, 0, 0, 0};
  int match_float_terms = 3;   // bf16 terms the general-float GEMM of the last call contracted (1 with a score bound)
  long long kernel_launches = 0;   // kernels launched by this context (bench.py's gpu_launches)
  int landmarks_prepared = 0;      // rows of the landmark set converted by vo_landmarks_prepare (0: none)
  long long alloc_generation = 0;  // bumped whenever a scratch buffer or the SIFT plan is (re)allocated: captured graphs hold the old addresses
  std::vector<cudaEvent_t> upload_events;   // upload_2d: two staging slots per worker thread, one event each ...
  std::vector<char> upload_pending;         // ... and whether a DMA out of the slot has been queued (it must finish before the slot is refilled)
  int frames_graph = 0;            // vo_frames_use_graph: replay the frame loop's launch sequence as a CUDA graph (1: on; -1: capture failed, off)
  bool prof_enabled = false;
  std::vector<vo::ProfStage> prof_stages;
  std::vector<vo::ProfRec> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  int prof_late_stage = -1;   // stage whose bytes are accumulated on the device (scratch "prof_desc_bytes")
  int prof_stage_id(const char* name);
  void prof_collect();
  vo::SiftPlan* sift_plan = nullptr;
  vo::FramePlan* frame_plan = nullptr;

  // returns a device buffer of at least `bytes` (contents undefined), grown geometrically
  int dev(const char* name, size_t bytes, void** out);
  int pinned(const char* name, size_t bytes, void** out);
};

namespace vo {
template <typename T>
inline int dev_buf(vo_ctx* c, const char* name, size_t count, T** out) {
  void* p = nullptr;
  int r = c->dev(name, count * sizeof(T), &p);
  *out = static_cast<T*>(p);
  return r;
}
template <typename T>
inline int pin_buf(vo_ctx* c, const char* name, size_t count, T** out) {
  void* p = nullptr;
  int r = c->pinned(name, count * sizeof(T), &p);
  *out = static_cast<T*>(p);
  return r;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

void sift_plan_destroy(SiftPlan*);
void frame_plan_destroy(FramePlan*);

static inline int div_up(int a, int b) { return (a + b - 1) / b; }

// Opt-in dynamic shared memory of a kernel, raised (never lowered) once per (device, kernel).  Contexts on
// different host threads and different devices share kernels, so the record is process-wide and mutex-guarded.
int ensure_dyn_smem(const void* func, size_t bytes);
template <typename F>
inline int ensure_dyn_smem_of(F* func, size_t bytes) { return ensure_dyn_smem(reinterpret_cast<const void*>(func), bytes); }

// cudaMemcpy2DAsync(host -> device) for big inputs in PAGEABLE host memory (a MATLAB array, a NumPy array): a few host
// threads copy row groups into pinned staging slots and queue the DMA of each slot behind them, so the copy runs at the
// speed of several cores' memcpy instead of one staged stream.  Pinned or small sources go straight to
// cudaMemcpy2DAsync.  On return the source has been read completely (as with a pageable cudaMemcpyAsync).
int upload_2d(vo_ctx* ctx, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaStream_t st);

// RAII bracket around one or more launches of a stage
struct ProfScope {
  vo_ctx* c; cudaStream_t st; ProfRec rec; bool on;
  ProfScope(vo_ctx* ctx, cudaStream_t s, const char* name, double bytes = 0, double flops = 0, int kernels = 1);
  ~ProfScope();
};
}  // namespace vo
