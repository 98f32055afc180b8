// [desc, loc, scale, orient, metric, octave, layer] = vo_sift_mex(I, 'ContrastThreshold',0.0133,
//        'EdgeThreshold',10,'NumLayersInOctave',3,'Sigma',1.6)
// Drop-in for detectSIFTFeatures(I) + extractFeatures(I, pts, "Method","SIFT") at VO.m:79-84.
// I: HxW uint8 (or single/double in [0,1]), column major.  desc Mx128 single, loc Mx2 single
// (1-based [x y] = SIFTPoints.Location), scale = size/2, orient in radians, octave/layer int32.
#include "mex_common.h"
#include <math.h>

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 1) mexErrMsgIdAndTxt("vo:sift:nargin", "vo_sift_mex(I, ...) needs an image");
  if (nlhs > 7) mexErrMsgIdAndTxt("vo:sift:nargout", "too many outputs");
  const int rows = (int)mxGetM(prhs[0]), cols = (int)mxGetN(prhs[0]);
  if (rows < 1 || cols < 1) mexErrMsgIdAndTxt("vo:sift:empty", "image is empty");
  std::vector<uint8_t> conv;
  const uint8_t* img = nullptr;
  const mxClassID cls = mxGetClassID(prhs[0]);
  if (cls == mxUINT8_CLASS) {
    img = (const uint8_t*)mxGetData(prhs[0]);
  } else if (cls == mxSINGLE_CLASS || cls == mxDOUBLE_CLASS) {
    conv.resize((size_t)rows * cols);
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
  int cap = 16384, n = 0;
  std::vector<vo_keypoint> kps; std::vector<float> desc;
  for (;;) {
    kps.resize(cap); desc.resize((size_t)cap * 128);
    const int rc = vo_sift(vo_mex_ctx("vo_sift_mex"), img, rows, cols, /*ld=*/rows, /*col_major=*/1, &o, cap, kps.data(), desc.data(), &n);
    if (rc == VO_ERR_CAPACITY && cap < (1 << 20)) { cap *= 4; continue; }
    vo_mex_check(rc, "vo:sift:cuda");
    break;
  }
  plhs[0] = mxCreateNumericMatrix(n, 128, mxSINGLE_CLASS, mxREAL);
  float* d = (float*)mxGetData(plhs[0]);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 128; ++k) d[(size_t)k * n + i] = desc[(size_t)i * 128 + k];
  mxArray* loc = mxCreateNumericMatrix(n, 2, mxSINGLE_CLASS, mxREAL);
  mxArray* scale = mxCreateNumericMatrix(n, 1, mxSINGLE_CLASS, mxREAL);
  mxArray* orient = mxCreateNumericMatrix(n, 1, mxSINGLE_CLASS, mxREAL);
  mxArray* metric = mxCreateNumericMatrix(n, 1, mxSINGLE_CLASS, mxREAL);
  mxArray* octave = mxCreateNumericMatrix(n, 1, mxINT32_CLASS, mxREAL);
  mxArray* layer = mxCreateNumericMatrix(n, 1, mxINT32_CLASS, mxREAL);
  for (int i = 0; i < n; ++i) {
    ((float*)mxGetData(loc))[i] = kps[i].x; ((float*)mxGetData(loc))[n + i] = kps[i].y;
    ((float*)mxGetData(scale))[i] = kps[i].size * 0.5f;
    ((float*)mxGetData(orient))[i] = kps[i].angle * 0.017453292519943295f;
    ((float*)mxGetData(metric))[i] = kps[i].response;
    int oc = kps[i].octave & 255; oc = oc < 128 ? oc : (-128 | oc);
    ((int32_t*)mxGetData(octave))[i] = oc;
    ((int32_t*)mxGetData(layer))[i] = (kps[i].octave >> 8) & 255;
  }
  mxArray* outs[6] = {loc, scale, orient, metric, octave, layer};
  for (int k = 0; k < 6; ++k) {
    if (nlhs > k + 1) plhs[k + 1] = outs[k];
    else mxDestroyArray(outs[k]);
  }
}
