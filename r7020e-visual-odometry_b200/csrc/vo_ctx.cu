// vo_ctx.cu -- context lifetime, error string, scratch buffers, driver entry points.
#include "vo_internal.h"

namespace vo {

static thread_local char g_err[1024] = "";

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

}  // namespace vo

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
  *out = c;
  return VO_OK;
}

void vo_ctx_destroy(vo_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->sift_plan) vo::sift_plan_destroy(c->sift_plan);
  if (c->frame_plan) vo::frame_plan_destroy(c->frame_plan);
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

}  // extern "C"
