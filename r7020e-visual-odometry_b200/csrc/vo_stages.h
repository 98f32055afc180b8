// vo_stages.h -- device-resident stage launchers shared with the frame loop (vo_frames.cu).
#pragma once
#include "vo_internal.h"

namespace vo {
// vo_sift.cu
int sift_prepare(vo_ctx* ctx, int rows, int cols, int batch, const vo_sift_opts* opts, int capacity, SiftPlan** plan,
                 vo_sift_opts* filled);
int sift_run_device(vo_ctx* ctx, SiftPlan* p, int batch, const vo_sift_opts& o, cudaStream_t st);
uint8_t* sift_plan_images(SiftPlan* p);
// MATLAB-ordered images (element (r,c) at src[c*rows + r]) staged in the plan -> the plan's row-major image buffer
int sift_load_col_major(vo_ctx* ctx, SiftPlan* p, int first_img, int img_step, const uint8_t* src, int n, bool on_device, cudaStream_t st);
vo_keypoint* sift_plan_keypoints(SiftPlan* p);
float* sift_plan_desc(SiftPlan* p);
int* sift_plan_counters(SiftPlan* p);   // counters[img*4 + 2] = keypoints of image img
int sift_plan_kp_cap(SiftPlan* p);
// vo_geom.cu
void fill_p3p_opts(const vo_p3p_opts* in, vo_p3p_opts* o);
int p3p_batch_device(vo_ctx* ctx, const double* img, const double* world, const int* n_dev, int cap, int n_prob,
                     const double* K4_dev, const vo_p3p_opts& o, double* A_dev, uint8_t* inliers_dev, int* status_dev,
                     int* info_dev, cudaStream_t st, const uint64_t* seed_add_dev = nullptr);   // seed_add_dev: added to o.seed on the device
int triangulate_batch_device(const double* pts1, const double* pts2, const int* n_dev, int n_stride, int cap, int n_prob,
                             const double* P_dev, double* xyz, cudaStream_t st);
int landmarks_device(const vo_keypoint* kps, int kc, const uint32_t* l0, const uint32_t* r0, const int* K0, const double* old_l,
                     const double* old_r, const int* K4, const int* status, const double* P_dev, const double* poses_dev,
                     int n_frames, uint32_t* newidx, int* n_new, double* out, int cap, int* rows, cudaStream_t st);
}  // namespace vo
