// vo_ctx.cu -- context lifetime, error string, scratch buffers, driver entry points.
#include "vo_internal.h"
#include <atomic>
#include <mutex>
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
