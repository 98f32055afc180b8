// vo_ctx.cu -- context lifetime, error string, scratch buffers, driver entry points.
#include "vo_internal.h"
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>
#include <algorithm>
#include <cstring>
#include <utility>

namespace vo {

static thread_local char g_err[1024] = "";

int ensure_dyn_smem(const void* func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> done;
  int dev = 0;
  VO_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> g(mu);
  size_t& cur = done[std::make_pair(dev, func)];
  if (bytes > cur) {
    VO_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cur = bytes;
  }
  return VO_OK;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

ProfScope::ProfScope(vo_ctx* ctx, cudaStream_t s, const char* name, double bytes, double flops, int kernels) : c(ctx), st(s), on(false) {
  if (ctx) ctx->kernel_launches += kernels;
  if (!ctx || !ctx->prof_enabled) return;
  rec.stage = ctx->prof_stage_id(name);
  ProfStage& ps = ctx->prof_stages[rec.stage];
  ps.bytes += bytes; ps.flops += flops; ps.launches += 1;
  auto get = [&]() { cudaEvent_t e; if (!ctx->prof_pool.empty()) { e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); } else cudaEventCreate(&e); return e; };
  rec.e0 = get(); rec.e1 = get();
  cudaEventRecord(rec.e0, st);
  on = true;
}
ProfScope::~ProfScope() {
  if (!on) return;
  cudaEventRecord(rec.e1, st);
  c->prof_pending.push_back(rec);
}


int upload_2d(vo_ctx* ctx, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaStream_t st) {
  if (width == 0 || height == 0) return VO_OK;
  static const int n_thr_env = [] { const char* e = getenv("VO_UPLOAD_THREADS"); return e ? atoi(e) : 4; }();
  constexpr size_t SLOT = 2u << 20;
  cudaPointerAttributes at;
  bool pageable = false;
  if (cudaPointerGetAttributes(&at, src) == cudaSuccess) pageable = at.type == cudaMemoryTypeUnregistered;
  else (void)cudaGetLastError();
  int n_thr = n_thr_env;
  const unsigned hw = std::thread::hardware_concurrency();
  if (hw > 0 && (unsigned)n_thr > hw) n_thr = (int)hw;
  if ((size_t)n_thr > height) n_thr = (int)height;
  if (!pageable || n_thr < 2 || width * height < (8u << 20) || width > SLOT) {
    VO_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, st));
    return VO_OK;
  }
  uint8_t* stage; VO_TRY(pin_buf(ctx, "upload_stage", (size_t)n_thr * 2 * SLOT, &stage));
  while (ctx->upload_events.size() < (size_t)n_thr * 2) {
    cudaEvent_t e; VO_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->upload_events.push_back(e);
    ctx->upload_pending.push_back(0);
  }
  const size_t rows_per_slot = SLOT / width;       // >= 1
  std::vector<int> rc(n_thr, (int)cudaSuccess);
  std::vector<std::thread> pool;
  const int device = ctx->device;
  bool spawn_failed = false;
  for (int t = 0; t < n_thr && !spawn_failed; ++t) {
    try {
    pool.emplace_back([&, t]() {
      cudaSetDevice(device);
      const size_t r0 = height * t / n_thr, r1 = height * (t + 1) / n_thr;
      int k = 0;
      for (size_t r = r0; r < r1; r += rows_per_slot, k ^= 1) {
        const size_t nr = std::min(rows_per_slot, r1 - r);
        uint8_t* slot = stage + ((size_t)t * 2 + k) * SLOT;
        cudaEvent_t ev = ctx->upload_events[(size_t)t * 2 + k];
        char& pending = ctx->upload_pending[(size_t)t * 2 + k];   // (also set by an earlier call: left, then right images)
        if (pending) { const cudaError_t e = cudaEventSynchronize(ev); if (e != cudaSuccess) { rc[t] = (int)e; return; } }
        for (size_t q = 0; q < nr; ++q) memcpy(slot + q * width, static_cast<const uint8_t*>(src) + (r + q) * spitch, width);
        cudaError_t e = cudaMemcpy2DAsync(static_cast<uint8_t*>(dst) + r * dpitch, dpitch, slot, width, width, nr, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(ev, st);
        if (e != cudaSuccess) { rc[t] = (int)e; return; }
        pending = 1;
      }
    });
    } catch (...) { spawn_failed = true; }   // no thread to be had (resource limits): the rows are copied the plain way below
  }
  for (auto& th : pool) th.join();
  for (int t = 0; t < n_thr; ++t) VO_CUDA((cudaError_t)rc[t]);
  if (spawn_failed) VO_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, st));
  return VO_OK;
}

}  // namespace vo

int vo_ctx::prof_stage_id(const char* name) {
  for (size_t i = 0; i < prof_stages.size(); ++i)
    if (prof_stages[i].name == name) return (int)i;
  vo::ProfStage ps; ps.name = name;
  prof_stages.push_back(ps);
  return (int)prof_stages.size() - 1;
}

void vo_ctx::prof_collect() {
  if (prof_late_stage >= 0 && prof_late_stage < (int)prof_stages.size()) {
    auto it = scratch.find("prof_desc_bytes");
    if (it != scratch.end() && it->second.ptr) {
      unsigned long long h[64];
      cudaStreamSynchronize(stream);
      if (cudaMemcpy(h, it->second.ptr, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
        double tot = 0;
        for (int i = 0; i < 64; ++i) tot += (double)h[i];
        prof_stages[prof_late_stage].bytes = tot;
      }
    }
  }
  for (auto& r : prof_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess)
      prof_stages[r.stage].ms += ms;
    prof_pool.push_back(r.e0); prof_pool.push_back(r.e1);
  }
  prof_pending.clear();
}

int vo_ctx::dev(const char* name, size_t bytes, void** out) {
  vo::Scratch& s = scratch[name];
  if (s.bytes < bytes || s.ptr == nullptr) {
    if (s.ptr) {
      VO_CUDA(cudaStreamSynchronize(stream));
      VO_CUDA(cudaFree(s.ptr));
      s.ptr = nullptr; s.bytes = 0;
    }
    size_t want = bytes < 256 ? 256 : bytes + bytes / 4;
    VO_CUDA(cudaMalloc(&s.ptr, want));
    s.bytes = want;
    s.host = false;
    ++alloc_generation;
  }
  *out = s.ptr;
  return VO_OK;
}

int vo_ctx::pinned(const char* name, size_t bytes, void** out) {
  vo::Scratch& s = scratch[std::string("pin:") + name];
  if (s.bytes < bytes || s.ptr == nullptr) {
    if (s.ptr) {
      VO_CUDA(cudaStreamSynchronize(stream));
      VO_CUDA(cudaFreeHost(s.ptr));
      s.ptr = nullptr; s.bytes = 0;
    }
    size_t want = bytes < 256 ? 256 : bytes + bytes / 4;
    VO_CUDA(cudaMallocHost(&s.ptr, want));
    s.bytes = want;
    s.host = true;
    ++alloc_generation;
  }
  *out = s.ptr;
  return VO_OK;
}

extern "C" {

int vo_version(void) { return 100; }

const char* vo_last_error(void) { return vo::g_err; }

int vo_ctx_create(int device, vo_ctx** out) {
  VO_CHECK_ARG(out != nullptr, "ctx out pointer is null");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    vo::set_error("no CUDA device available (%s); libvo_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return VO_ERR_CUDA;
  }
  VO_CHECK_ARG(device >= 0 && device < count, "device index out of range");
  VO_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  VO_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    vo::set_error("device %d is sm_%d%d; libvo_b200 is built for sm_100a only", device, prop.major,
                  prop.minor);
    return VO_ERR_CUDA;
  }
  vo_ctx* c = new vo_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    vo::set_error("cudaStreamCreate failed");
    return VO_ERR_CUDA;
  }
  { const char* e = getenv("VO_FRAMES_GRAPH"); c->frames_graph = (e && atoi(e) != 0) ? 1 : 0; }   // see vo_frames_use_graph
  *out = c;
  return VO_OK;
}

void vo_ctx_destroy(vo_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->sift_plan) vo::sift_plan_destroy(c->sift_plan);
  if (c->frame_plan) vo::frame_plan_destroy(c->frame_plan);
  c->prof_collect();
  for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : c->upload_events) cudaEventDestroy(e);
  for (auto& kv : c->scratch) {
    if (!kv.second.ptr) continue;
    if (kv.second.host) cudaFreeHost(kv.second.ptr);
    else cudaFree(kv.second.ptr);
  }
  cudaStreamDestroy(c->stream);
  delete c;
}

int vo_ctx_sync(vo_ctx* c) {
  VO_CHECK_ARG(c != nullptr, "ctx is null");
  VO_CUDA(cudaStreamSynchronize(c->stream));
  return VO_OK;
}

void* vo_ctx_stream(vo_ctx* c) { return c ? (void*)c->stream : nullptr; }

int vo_profile_enable(vo_ctx* c, int on) {
  VO_CHECK_ARG(c != nullptr, "ctx is null");
  c->prof_collect();
  c->prof_enabled = on != 0;
  c->prof_stages.clear();
  c->prof_late_stage = -1;
  {
    void* p = nullptr;
    if (c->dev("prof_desc_bytes", 64 * sizeof(unsigned long long), &p) == VO_OK) cudaMemsetAsync(p, 0, 64 * sizeof(unsigned long long), c->stream);
  }
  return VO_OK;
}

long long vo_kernel_launches(vo_ctx* c) { return c ? c->kernel_launches : 0; }

int vo_profile_count(vo_ctx* c) {
  if (!c) return 0;
  c->prof_collect();
  return (int)c->prof_stages.size();
}

int vo_profile_get(vo_ctx* c, int i, char* name, int name_cap, double* ms, long long* launches, double* bytes, double* flops) {
  VO_CHECK_ARG(c != nullptr && i >= 0 && i < (int)c->prof_stages.size(), "bad stage index");
  const vo::ProfStage& ps = c->prof_stages[i];
  if (name && name_cap > 0) { strncpy(name, ps.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (ms) *ms = ps.ms;
  if (launches) *launches = ps.launches;
  if (bytes) *bytes = ps.bytes;
  if (flops) *flops = ps.flops;
  return VO_OK;
}

}  // extern "C"
