// vo_frames.cu -- one pass of the VO.m loop body (VO.m:64-232) over a batch of stereo frames,
// entirely on the device: no host round trip between SIFT, the five matchFeatures calls,
// find_remaining_points' index chain (VO.m:280-334), triangulation and P3P-MSAC.
//
// Frame i of the batch uses images 2i (left) and 2i+1 (right) of the SIFT plan.  Problem p of the
// tracking stage is the frame pair (p, p+1).  Every data-dependent size (keypoints per image,
// K0..K4) lives in device memory; launch shapes depend on capacities only.
#include "vo_internal.h"
#include "vo_match.h"
#include "vo_stages.h"

namespace vo {

// What the landmark pass (vo_frames_landmarks) needs from the most recent vo_frames call on the context: the device
// arrays live in the context's scratch buffers and stay valid until the next vo_frames call.
struct FramePlan {
  int n = 0, kc = 0;
  const vo_keypoint* kps = nullptr;
  const uint32_t *l0 = nullptr, *r0 = nullptr;
  const int* K = nullptr;                // K[s*n + p]
  const double *old_l = nullptr, *old_r = nullptr, *dP = nullptr;
  const int* dstat = nullptr;
  // vo_frames_use_graph: the launch sequence of one call (SIFT ... P3P and the result copies), captured once per
  // (shape, options, buffer generation) and replayed; the image upload and the per-call parameters stay outside
  std::vector<unsigned char> g_key;
  long long g_gen = -1;
  cudaGraphExec_t g_exec = nullptr;
  long long g_launches = 0;
};
void frame_plan_destroy(FramePlan* p) {
  if (p && p->g_exec) cudaGraphExecDestroy(p->g_exec);
  delete p;
}

// out_k[p][k] = src_k[p][idx[p][k]] for k < cnt[p]   (up to two arrays share one index list)
__global__ void __launch_bounds__(256)
compose_kernel(uint32_t* __restrict__ out1, const uint32_t* __restrict__ src1, uint32_t* __restrict__ out2,
               const uint32_t* __restrict__ src2, const uint32_t* __restrict__ idx, const int* __restrict__ cnt,
               int stride) {
  const int p = blockIdx.y;
  const int n = min(cnt[p], stride);
  for (int k = blockIdx.x * 256 + threadIdx.x; k < n; k += gridDim.x * 256) {
    const uint32_t j = idx[(size_t)p * stride + k];
    if (out1) out1[(size_t)p * stride + k] = src1[(size_t)p * stride + j];
    if (out2) out2[(size_t)p * stride + k] = src2[(size_t)p * stride + j];
  }
}

// final gather of find_remaining_points + the Location reads of VO.m:114,124:
//   old L/R pixel positions (previous frame, for triangulation), current L positions (for P3P)
__global__ void __launch_bounds__(256)
gather_points_kernel(const vo_keypoint* __restrict__ kps, int kp_cap, const uint32_t* __restrict__ oL2,
                     const uint32_t* __restrict__ oR2, const uint32_t* __restrict__ cL3,
                     const uint32_t* __restrict__ a4, const uint32_t* __restrict__ b4, const int* __restrict__ k4,
                     double* __restrict__ old_l, double* __restrict__ old_r, double* __restrict__ cur_l) {
  const int p = blockIdx.y;
  const int n = min(k4[p], kp_cap);
  const vo_keypoint* kl0 = kps + (size_t)(2 * p) * kp_cap;       // previous left
  const vo_keypoint* kr0 = kps + (size_t)(2 * p + 1) * kp_cap;   // previous right
  const vo_keypoint* kl1 = kps + (size_t)(2 * p + 2) * kp_cap;   // current left
  for (int k = blockIdx.x * 256 + threadIdx.x; k < n; k += gridDim.x * 256) {
    const size_t o = (size_t)p * kp_cap + k;
    const uint32_t ia = a4[o], ib = b4[o];
    const vo_keypoint ol = kl0[oL2[(size_t)p * kp_cap + ib]];
    const vo_keypoint orr = kr0[oR2[(size_t)p * kp_cap + ib]];
    const vo_keypoint cl = kl1[cL3[(size_t)p * kp_cap + ia]];
    old_l[2 * o] = (double)ol.x; old_l[2 * o + 1] = (double)ol.y;
    old_r[2 * o] = (double)orr.x; old_r[2 * o + 1] = (double)orr.y;
    cur_l[2 * o] = (double)cl.x; cur_l[2 * o + 1] = (double)cl.y;
  }
}

}  // namespace vo

using namespace vo;

static int frames_core(vo_ctx* ctx, const uint8_t* left, const uint8_t* right, int on_device, int n, int rows, int cols,
                       const double P1[12], const double P2[12], const vo_frames_opts* opts, double* rel_pose,
                       int* status, int* counts) {
  VO_CHECK_ARG(ctx && left && right && P1 && P2 && rel_pose && status, "null argument");
  VO_CHECK_ARG(n >= 1 && rows > 0 && cols > 0, "bad size");
  VO_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  vo_match_opts mo; memset(&mo, 0, sizeof(mo)); fill_match_opts(opts ? &opts->match : nullptr, &mo);   // (zeroed: the option structs are part of the graph key)
  mo.index_base = 0;
  vo_p3p_opts po; memset(&po, 0, sizeof(po)); fill_p3p_opts(opts ? &opts->p3p : nullptr, &po);
  const int want_cap = (opts && opts->max_keypoints > 0) ? opts->max_keypoints : 8192;
  const int first_frame = opts ? opts->first_frame : 0;
  SiftPlan* plan; vo_sift_opts so; memset(&so, 0, sizeof(so));
  VO_TRY(sift_prepare(ctx, rows, cols, 2 * n, opts ? &opts->sift : nullptr, want_cap, &plan, &so));
  const int kc = sift_plan_kp_cap(plan);
  const size_t img_bytes = (size_t)rows * cols;
  uint8_t* dimg = sift_plan_images(plan);
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if (opts && opts->col_major) {   // MATLAB H x W x N stacks: staged as they are, transposed on the device
    VO_TRY(sift_load_col_major(ctx, plan, 0, 2, left, n, on_device != 0, st));
    VO_TRY(sift_load_col_major(ctx, plan, 1, 2, right, n, on_device != 0, st));
    ctx->kernel_launches += 2;
  } else {
    if (on_device) {
      VO_CUDA(cudaMemcpy2DAsync(dimg, 2 * img_bytes, left, img_bytes, img_bytes, n, kind, st));
      VO_CUDA(cudaMemcpy2DAsync(dimg + img_bytes, 2 * img_bytes, right, img_bytes, img_bytes, n, kind, st));
    } else {   // pageable stacks go through threaded pinned staging
      VO_TRY(upload_2d(ctx, dimg, 2 * img_bytes, left, img_bytes, img_bytes, n, st));
      VO_TRY(upload_2d(ctx, dimg + img_bytes, 2 * img_bytes, right, img_bytes, img_bytes, n, st));
    }
  }
  const float* desc = sift_plan_desc(plan);
  const vo_keypoint* kps = sift_plan_keypoints(plan);
  const int* cnt = sift_plan_counters(plan);
  const size_t img_stride = (size_t)kc * 128;   // floats per image's descriptor block
  const int np = n - 1;

  // ---- buffers (no launches here: a replayed graph needs the same addresses and nothing else)
  // index / count arrays, [n][kc] each
  uint32_t *l0, *r0, *a1, *b1, *oL1, *oR1, *a2, *b2, *oL2, *oR2, *a3, *b3, *cL3, *cR3, *a4, *b4;
  uint32_t** arrs[] = {&l0, &r0, &a1, &b1, &oL1, &oR1, &a2, &b2, &oL2, &oR2, &a3, &b3, &cL3, &cR3, &a4, &b4};
  uint32_t* pool; VO_TRY(dev_buf(ctx, "fr_idx", (size_t)16 * n * kc, &pool));
  for (int i = 0; i < 16; ++i) *arrs[i] = pool + (size_t)i * n * kc;
  int* K; VO_TRY(dev_buf(ctx, "fr_K", (size_t)5 * n + 8, &K));   // K[s*n + p], s = 0..4
  double *old_l, *old_r, *cur_l, *world, *dA, *dP;
  VO_TRY(dev_buf(ctx, "fr_oldl", (size_t)n * kc * 2, &old_l));
  VO_TRY(dev_buf(ctx, "fr_oldr", (size_t)n * kc * 2, &old_r));
  VO_TRY(dev_buf(ctx, "fr_curl", (size_t)n * kc * 2, &cur_l));
  VO_TRY(dev_buf(ctx, "fr_world", (size_t)n * kc * 3, &world));
  VO_TRY(dev_buf(ctx, "fr_A", (size_t)n * 16, &dA));
  VO_TRY(dev_buf(ctx, "fr_P", 32, &dP));
  int* dstat; VO_TRY(dev_buf(ctx, "fr_stat", (size_t)n * 4, &dstat));
  double* hA; VO_TRY(pin_buf(ctx, "fr_hA", (size_t)n * 16, &hA));
  int* hI; VO_TRY(pin_buf(ctx, "fr_hI", (size_t)n * 4 + (size_t)5 * n + (size_t)2 * n * 4 + 16, &hI));
  int* hK = hI + (size_t)n * 4;
  int* hC = hK + (size_t)5 * n;

  // ---- per-call parameters (outside the replayed part): projection matrices, intrinsics and the call's share of the
  // P3P seed, which the hypothesis kernel adds on the device
  double* hP; VO_TRY(pin_buf(ctx, "fr_hP", 32, &hP));
  for (int k = 0; k < 12; ++k) { hP[k] = P1[k]; hP[12 + k] = P2[k]; }
  hP[24] = P1[0]; hP[25] = P1[5]; hP[26] = P1[2]; hP[27] = P1[6];   // fx fy cx cy (VO.m:35-38)
  const uint64_t seed_add = (uint64_t)(first_frame + 1) * 0x9E3779B97F4A7C15ull;
  memcpy(hP + 28, &seed_add, sizeof(seed_add));
  VO_CUDA(cudaMemcpyAsync(dP, hP, 29 * sizeof(double), cudaMemcpyHostToDevice, st));
  const uint64_t* seed_add_dev = reinterpret_cast<const uint64_t*>(dP + 28);

  const MatchFilter mflt = make_match_filter(mo);
  // ---- the launch sequence of one call
  auto enqueue = [&]() -> int {
  VO_TRY(sift_run_device(ctx, plan, 2 * n, so, st));
  VO_CUDA(cudaMemsetAsync(K, 0, ((size_t)5 * n + 8) * sizeof(int), st));

  auto raw_op = [&](int first_img) {   // raw descriptor set of image (first_img + 2p)
    MatchOperand o; o.base = desc + (size_t)first_img * img_stride; o.prob_stride = 2 * img_stride;
    o.integer_rows = 1;   // written by sift_descriptor_kernel: saturated, rounded to 0..255
    o.count = cnt + first_img * 4 + 2; o.count_stride = 8; o.cap = kc; return o;
  };
  auto gath_op = [&](int first_img, const uint32_t* g, const int* c) {
    MatchOperand o = raw_op(first_img); o.gather = g; o.gather_stride = kc; o.count = c; o.count_stride = 1; return o;
  };
  MatchTop2 t, tb;
  // matchFeatures(A, B) of VO.m for a batch of problems; with Unique also the reversed problems (B, A)
  auto match = [&](const MatchOperand& A, const MatchOperand& B, int nprob, uint32_t* i1, uint32_t* i2, int* npairs) -> int {
    VO_TRY(match_batch_top2(ctx, A, B, nprob, 128, "fr", nullptr, st, &t, &mflt));
    if (mo.unique) VO_TRY(match_batch_top2(ctx, B, A, nprob, 128, "frb", nullptr, st, &tb, &mflt));
    return match_batch_select(ctx, t, A, B, nprob, mo, i1, i2, nullptr, kc, npairs, 1, st, mo.unique ? &tb : nullptr);
  };
  // matched = matchFeatures(l_desc, r_desc)                                      VO.m:87
  {
    MatchOperand A = raw_op(0), B = raw_op(1);
    VO_TRY(match(A, B, n, l0, r0, K));
  }
  const dim3 cg(8, np > 0 ? np : 1);
  if (np > 0) {
    // M1 = matchFeatures(cur.l_desc, old.l_desc)                                 VO.m:283
    {
      MatchOperand A = raw_op(2), B = gath_op(0, l0, K);
      VO_TRY(match(A, B, np, a1, b1, K + n));
      compose_kernel<<<cg, 256, 0, st>>>(oL1, l0, oR1, r0, b1, K + n, kc);            // VO.m:287-290
      ctx->kernel_launches += 6;   // 5 compose launches + gather_points below
    }
    // M2 = matchFeatures(cur.r_desc, old.r_desc)                                 VO.m:293
    {
      MatchOperand A = raw_op(3), B = gath_op(1, oR1, K + n);
      VO_TRY(match(A, B, np, a2, b2, K + 2 * n));
      compose_kernel<<<cg, 256, 0, st>>>(oL2, oL1, oR2, oR1, b2, K + 2 * n, kc);      // VO.m:297-300
    }
    // M3 = matchFeatures(cur.l_desc(M1(:,1)), cur.r_desc(M2(:,1)))               VO.m:305-311
    {
      MatchOperand A = gath_op(2, a1, K + n), B = gath_op(3, a2, K + 2 * n);
      VO_TRY(match(A, B, np, a3, b3, K + 3 * n));
      compose_kernel<<<cg, 256, 0, st>>>(cL3, a1, nullptr, nullptr, a3, K + 3 * n, kc);  // VO.m:314-315
      compose_kernel<<<cg, 256, 0, st>>>(cR3, a2, nullptr, nullptr, b3, K + 3 * n, kc);  // VO.m:316-317
    }
    // M4 = matchFeatures(cur.l_desc, old.l_desc)                                 VO.m:323
    {
      MatchOperand A = gath_op(2, cL3, K + 3 * n), B = gath_op(0, oL2, K + 2 * n);
      VO_TRY(match(A, B, np, a4, b4, K + 4 * n));
    }
    // triangulate the old pair (VO.m:114) and estimate the pose (VO.m:123-127)
    gather_points_kernel<<<cg, 256, 0, st>>>(kps, kc, oL2, oR2, cL3, a4, b4, K + 4 * n, old_l, old_r, cur_l);
    {
      ProfScope ps(ctx, st, "triangulate");
      VO_TRY(triangulate_batch_device(old_l, old_r, K + 4 * n, 1, kc, np, dP, world, st));
    }
    ProfScope ps(ctx, st, "p3p_msac", 0.0, 0.0, 4);
    VO_TRY(p3p_batch_device(ctx, cur_l, world, K + 4 * n, kc, np, dP + 24, po, dA, nullptr, dstat, dstat + n, st, seed_add_dev));
  }
  // results
  if (np > 0) {
    VO_CUDA(cudaMemcpyAsync(hA, dA, (size_t)np * 16 * sizeof(double), cudaMemcpyDeviceToHost, st));
    VO_CUDA(cudaMemcpyAsync(hI, dstat, (size_t)n * 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  VO_CUDA(cudaMemcpyAsync(hK, K, (size_t)5 * n * sizeof(int), cudaMemcpyDeviceToHost, st));
  VO_CUDA(cudaMemcpyAsync(hC, cnt, (size_t)2 * n * 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  return VO_OK;
  };

  if (!ctx->frame_plan) ctx->frame_plan = new FramePlan();
  FramePlan* fp = ctx->frame_plan;
  fp->n = n; fp->kc = kc; fp->kps = kps; fp->l0 = l0; fp->r0 = r0; fp->K = K; fp->old_l = old_l; fp->old_r = old_r; fp->dP = dP;
  fp->dstat = dstat;

  // CUDA graph (vo_frames_use_graph): the first call with a given shape and option set runs as it is (it allocates);
  // the second one is captured, every later one only uploads its inputs and replays
  std::vector<unsigned char> key;
  auto put = [&](const void* q, size_t bytes) { const unsigned char* c = static_cast<const unsigned char*>(q); key.insert(key.end(), c, c + bytes); };
  const bool want_graph = ctx->frames_graph > 0 && !ctx->prof_enabled;
  if (want_graph) {
    const int dims[5] = {n, rows, cols, kc, want_cap};
    put(dims, sizeof(dims)); put(&so, sizeof(so)); put(&mo, sizeof(mo)); put(&po, sizeof(po)); put(&plan, sizeof(plan));
  }
  const bool same = want_graph && fp->g_key == key && fp->g_gen == ctx->alloc_generation;
  if (same && fp->g_exec != nullptr) {
    VO_CUDA(cudaGraphLaunch(fp->g_exec, st));
    ctx->kernel_launches += fp->g_launches;
  } else if (same) {
    const long long l0c = ctx->kernel_launches;
    VO_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    const int rc_cap = enqueue();
    cudaGraph_t graph = nullptr;
    const cudaError_t e_end = cudaStreamEndCapture(st, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc_cap == VO_OK && e_end == cudaSuccess && graph != nullptr && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
      fp->g_exec = exec;
      fp->g_launches = ctx->kernel_launches - l0c;
      cudaGraphDestroy(graph);
      VO_CUDA(cudaGraphLaunch(fp->g_exec, st));
    } else {   // something in the sequence cannot be captured on this driver: run it the plain way from now on
      if (graph) cudaGraphDestroy(graph);
      (void)cudaGetLastError();
      ctx->frames_graph = -1;
      ctx->kernel_launches = l0c;
      VO_TRY(enqueue());
    }
  } else {
    VO_TRY(enqueue());
    if (fp->g_exec) { cudaGraphExecDestroy(fp->g_exec); fp->g_exec = nullptr; }
    fp->g_key = key;                       // empty when graphs are off
    fp->g_gen = ctx->alloc_generation;     // after the allocations of this call
  }
  VO_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < 16; ++k) rel_pose[k] = (k % 5 == 0) ? 1.0 : 0.0;
  status[0] = 0;
  int rc = VO_OK;
  for (int i = 0; i < n; ++i) {
    if (i >= 1) {
      memcpy(rel_pose + 16 * i, hA + 16 * (i - 1), 16 * sizeof(double));
      status[i] = hI[i - 1];
    }
    const int nl = hC[(2 * i) * 4 + 2], nr = hC[(2 * i + 1) * 4 + 2];
    if (nl > kc || nr > kc || hC[(2 * i) * 4 + 1] > kc || hC[(2 * i + 1) * 4 + 1] > kc) {
      set_error("vo_frames: frame %d has more keypoints (%d / %d) than max_keypoints capacity %d", i, nl, nr, kc);
      rc = VO_ERR_CAPACITY;
    } else if (hC[(2 * i) * 4 + 0] > 8 * kc || hC[(2 * i + 1) * 4 + 0] > 8 * kc) {   // the plan keeps 8 candidates per keypoint slot
      set_error("vo_frames: frame %d has more extrema candidates (%d / %d) than the plan holds (%d); raise max_keypoints",
                i, hC[(2 * i) * 4 + 0], hC[(2 * i + 1) * 4 + 0], 8 * kc);
      rc = VO_ERR_CAPACITY;
    }
    if (counts) {
      int* c = counts + 8 * i;
      c[0] = nl; c[1] = nr; c[2] = hK[i];
      c[3] = i >= 1 ? hK[n + i - 1] : 0; c[4] = i >= 1 ? hK[2 * n + i - 1] : 0;
      c[5] = i >= 1 ? hK[3 * n + i - 1] : 0; c[6] = i >= 1 ? hK[4 * n + i - 1] : 0;
      c[7] = i >= 1 ? hI[n + 3 * (i - 1)] : 0;
    }
  }
  return rc;
}

// VO.m:145-161 for every frame of the last batch, after the caller has multiplied the pose chain (VO.m:130).
static int frames_landmarks(vo_ctx* ctx, const double* poses, int n, int cap, double* landmarks, int* rows) {
  VO_CHECK_ARG(ctx && poses && landmarks && rows, "null argument");
  FramePlan* fp = ctx->frame_plan;
  if (!fp || fp->n < 1) { set_error("vo_frames_landmarks: no vo_frames call has run on this context"); return VO_ERR_STATE; }
  VO_CHECK_ARG(n == fp->n, "vo_frames_landmarks: n_frames differs from the last vo_frames call");
  VO_CHECK_ARG(cap >= 2, "vo_frames_landmarks: cap must be at least 2");
  VO_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int kc = fp->kc;
  const int dcap = cap < kc ? cap : kc;      // a frame has at most kc stereo matches
  uint32_t* newidx; VO_TRY(dev_buf(ctx, "lm_idx", (size_t)n * kc, &newidx));
  int* cnt; VO_TRY(dev_buf(ctx, "lm_cnt", (size_t)2 * n, &cnt));        // n_new[n], rows[n]
  double* dposes; VO_TRY(dev_buf(ctx, "lm_pose", (size_t)n * 16, &dposes));
  double* dout; VO_TRY(dev_buf(ctx, "lm_out", (size_t)n * dcap * 3, &dout));
  VO_CUDA(cudaMemcpyAsync(dposes, poses, (size_t)n * 16 * sizeof(double), cudaMemcpyHostToDevice, st));
  VO_CUDA(cudaMemsetAsync(cnt, 0, (size_t)2 * n * sizeof(int), st));
  VO_CUDA(cudaMemsetAsync(dout, 0, (size_t)n * dcap * 3 * sizeof(double), st));
  {
    ProfScope ps(ctx, st, "landmarks", 0.0, 0.0, 2);
    VO_TRY(landmarks_device(fp->kps, kc, fp->l0, fp->r0, fp->K, fp->old_l, fp->old_r, fp->K + 4 * n, fp->dstat, fp->dP, dposes, n,
                            newidx, cnt, dout, dcap, cnt + n, st));
  }
  int* h; VO_TRY(pin_buf(ctx, "lm_h", (size_t)3 * n, &h));             // n_new[p], rows[i], status[p]
  VO_CUDA(cudaMemcpyAsync(h, cnt, (size_t)2 * n * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (n > 1) VO_CUDA(cudaMemcpyAsync(h + 2 * n, fp->dstat, (size_t)(n - 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  VO_CUDA(cudaStreamSynchronize(st));
  int rc = VO_OK;
  rows[0] = 0;
  for (int i = 1; i < n; ++i) {
    // the reference's array starts as zeros(2, 3) and grows to the last row written (CreateLandmarksFromFeatures.m:2, 17)
    const int r = h[2 * n + i - 1] != 0 ? 0 : (h[n + i] > 2 ? h[n + i] : 2);
    if (h[i - 1] > dcap) { set_error("vo_frames_landmarks: frame %d has %d new features, cap is %d", i, h[i - 1], cap); rc = VO_ERR_CAPACITY; }
    rows[i] = r;
    if (r > 0)
      VO_CUDA(cudaMemcpyAsync(landmarks + (size_t)i * cap * 3, dout + (size_t)i * dcap * 3, (size_t)r * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  VO_CUDA(cudaStreamSynchronize(st));
  return rc;
}

extern "C" {
int vo_frames_use_graph(vo_ctx* ctx, int enable) {
  VO_CHECK_ARG(ctx, "ctx is null");
  ctx->frames_graph = enable ? 1 : 0;
  return VO_OK;
}
int vo_frames_graph_state(vo_ctx* ctx) {
  if (!ctx) return 0;
  if (ctx->frames_graph <= 0) return ctx->frames_graph;
  return (ctx->frame_plan && ctx->frame_plan->g_exec) ? 2 : 1;
}
int vo_frames_landmarks(vo_ctx* ctx, const double* poses, int n_frames, int cap, double* landmarks, int* rows) {
  return frames_landmarks(ctx, poses, n_frames, cap, landmarks, rows);
}
int vo_frames(vo_ctx* ctx, const uint8_t* left, const uint8_t* right, int n, int rows, int cols, const double P1[12],
              const double P2[12], const vo_frames_opts* opts, double* rel_pose, int* status, int* counts) {
  return frames_core(ctx, left, right, 0, n, rows, cols, P1, P2, opts, rel_pose, status, counts);
}
int vo_frames_dev(vo_ctx* ctx, const uint8_t* left, const uint8_t* right, int n, int rows, int cols, const double P1[12],
                  const double P2[12], const vo_frames_opts* opts, double* rel_pose, int* status, int* counts) {
  return frames_core(ctx, left, right, 1, n, rows, cols, P1, P2, opts, rel_pose, status, counts);
}
}
