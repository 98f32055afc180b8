// [desc, loc, scale, orient, metric, octave, layer, count] = vo_sift_mex(I, 'ContrastThreshold',0.0133,
//        'EdgeThreshold',10,'NumLayersInOctave',3,'Sigma',1.6)
// Drop-in for detectSIFTFeatures(I) + extractFeatures(I, pts, "Method","SIFT") at VO.m:79-84.
// I: HxW uint8 (or single/double in [0,1]), column major.  desc Mx128 single, loc Mx2 single
// (1-based [x y] = SIFTPoints.Location), scale = size/2, orient in radians, octave/layer int32.
// I may also be an H x W x N uint8 stack (e.g. cat(3, lf, rf): both images of VO.m:79-84 in one call): the
// outputs of the N images are concatenated along the rows and count (N x 1 int32) gives the rows of each image.
// The descriptors are transposed to MATLAB's layout on the device and copied straight into the output array.
#include "mex_common.h"
#include <math.h>

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 1) mexErrMsgIdAndTxt("vo:sift:nargin", "vo_sift_mex(I, ...) needs an image");
  if (nlhs > 8) mexErrMsgIdAndTxt("vo:sift:nargout", "too many outputs");
  const mwSize nd = mxGetNumberOfDimensions(prhs[0]);
  const mwSize* dims = mxGetDimensions(prhs[0]);
  if (nd > 3) mexErrMsgIdAndTxt("vo:sift:size", "image must be H x W or H x W x N");
  const int rows = (int)dims[0], cols = (int)dims[1], n_img = nd == 3 ? (int)dims[2] : 1;
  if (rows < 1 || cols < 1 || n_img < 1) mexErrMsgIdAndTxt("vo:sift:empty", "image is empty");
  std::vector<uint8_t> conv;
  const uint8_t* img = nullptr;
  const mxClassID cls = mxGetClassID(prhs[0]);
  if (cls == mxUINT8_CLASS) {
    img = (const uint8_t*)mxGetData(prhs[0]);
  } else if (cls == mxSINGLE_CLASS || cls == mxDOUBLE_CLASS) {
    conv.resize((size_t)rows * cols * n_img);
    for (size_t i = 0; i < conv.size(); ++i) {
      const double v = cls == mxSINGLE_CLASS ? ((const float*)mxGetData(prhs[0]))[i] : mxGetPr(prhs[0])[i];
      const double s = nearbyint(v * 255.0);
      conv[i] = (uint8_t)(s < 0 ? 0 : (s > 255 ? 255 : s));
    }
    img = conv.data();
  } else {
    mexErrMsgIdAndTxt("vo:sift:class", "image must be uint8, single or double");
  }
  vo_mex_check_pairs(nrhs, 1);
  vo_sift_opts o; memset(&o, 0, sizeof(o));
  double v;
  if (vo_mex_opt(nrhs, prhs, 1, "ContrastThreshold", &v)) o.contrast_threshold = (float)v;
  if (vo_mex_opt(nrhs, prhs, 1, "EdgeThreshold", &v)) o.edge_threshold = (float)v;
  if (vo_mex_opt(nrhs, prhs, 1, "NumLayersInOctave", &v)) o.num_layers_in_octave = (int)v;
  if (vo_mex_opt(nrhs, prhs, 1, "Sigma", &v)) o.sigma = (float)v;
  o.index_base = 1;
  static int cap = 8192;             // grown on demand, remembered across calls
  std::vector<int> n(n_img);
  static std::vector<vo_keypoint> kps; static std::vector<float> desc;
  for (;;) {
    kps.resize((size_t)cap * n_img); desc.resize((size_t)cap * 128 * n_img);
    const int rc = vo_sift_stack(vo_mex_ctx("vo_sift_mex"), img, n_img, rows, cols, /*col_major=*/1, &o, cap, kps.data(), desc.data(),
                                 /*desc_col_major=*/1, n.data());
    if (rc == VO_ERR_CAPACITY && cap < (1 << 20)) { cap *= 4; continue; }
    vo_mex_check(rc, "vo:sift:cuda");
    break;
  }
  size_t total = 0;
  for (int b = 0; b < n_img; ++b) total += (size_t)n[b];
  plhs[0] = mxCreateNumericMatrix(total, 128, mxSINGLE_CLASS, mxREAL);
  mxArray* loc = mxCreateNumericMatrix(total, 2, mxSINGLE_CLASS, mxREAL);
  mxArray* scale = mxCreateNumericMatrix(total, 1, mxSINGLE_CLASS, mxREAL);
  mxArray* orient = mxCreateNumericMatrix(total, 1, mxSINGLE_CLASS, mxREAL);
  mxArray* metric = mxCreateNumericMatrix(total, 1, mxSINGLE_CLASS, mxREAL);
  mxArray* octave = mxCreateNumericMatrix(total, 1, mxINT32_CLASS, mxREAL);
  mxArray* layer = mxCreateNumericMatrix(total, 1, mxINT32_CLASS, mxREAL);
  mxArray* count = mxCreateNumericMatrix(n_img, 1, mxINT32_CLASS, mxREAL);
  float* d = (float*)mxGetData(plhs[0]);
  size_t r0 = 0;
  for (int b = 0; b < n_img; ++b) {
    const size_t m = (size_t)n[b];
    const float* src = desc.data() + (size_t)b * cap * 128;       // m x 128, column-major, leading dimension m
    for (int k = 0; k < 128; ++k) memcpy(d + (size_t)k * total + r0, src + (size_t)k * m, m * sizeof(float));
    const vo_keypoint* kp = kps.data() + (size_t)b * cap;
    for (size_t i = 0; i < m; ++i) {
      ((float*)mxGetData(loc))[r0 + i] = kp[i].x; ((float*)mxGetData(loc))[total + r0 + i] = kp[i].y;
      ((float*)mxGetData(scale))[r0 + i] = kp[i].size * 0.5f;
      ((float*)mxGetData(orient))[r0 + i] = kp[i].angle * 0.017453292519943295f;
      ((float*)mxGetData(metric))[r0 + i] = kp[i].response;
      int oc = kp[i].octave & 255; oc = oc < 128 ? oc : (-128 | oc);
      ((int32_t*)mxGetData(octave))[r0 + i] = oc;
      ((int32_t*)mxGetData(layer))[r0 + i] = (kp[i].octave >> 8) & 255;
    }
    ((int32_t*)mxGetData(count))[b] = (int32_t)m;
    r0 += m;
  }
  mxArray* outs[7] = {loc, scale, orient, metric, octave, layer, count};
  for (int k = 0; k < 7; ++k) {
    if (nlhs > k + 1) plhs[k + 1] = outs[k];
    else mxDestroyArray(outs[k]);
  }
}
