// [indexPairs, matchMetric] = vo_match_mex(f1, f2, 'MatchThreshold',1.0,'MaxRatio',0.6,'Unique',false)
// Drop-in for matchFeatures(f1, f2) at VO.m:87, 283, 293, 311, 323: N1xD / N2xD single, column
// major; indexPairs Px2 uint32, 1-based, ascending in column 1; matchMetric Px1 single.
#include "mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 2) mexErrMsgIdAndTxt("vo:match:nargin", "vo_match_mex(f1, f2, ...) needs two feature matrices");
  if (nlhs > 2) mexErrMsgIdAndTxt("vo:match:nargout", "too many outputs");
  for (int k = 0; k < 2; ++k)
    if (mxGetClassID(prhs[k]) != mxSINGLE_CLASS) mexErrMsgIdAndTxt("vo:match:class", "features must be single");
  const int n1 = (int)mxGetM(prhs[0]), n2 = (int)mxGetM(prhs[1]);
  const int d1 = (int)mxGetN(prhs[0]), d2 = (int)mxGetN(prhs[1]);
  if (n1 > 0 && n2 > 0 && d1 != d2) mexErrMsgIdAndTxt("vo:match:dim", "feature lengths differ (%d vs %d)", d1, d2);
  vo_mex_check_pairs(nrhs, 2);
  vo_match_opts o; memset(&o, 0, sizeof(o));
  double v;
  if (vo_mex_opt(nrhs, prhs, 2, "MatchThreshold", &v)) o.match_threshold = (float)v;
  if (vo_mex_opt(nrhs, prhs, 2, "MaxRatio", &v)) o.max_ratio = (float)v;
  if (vo_mex_opt(nrhs, prhs, 2, "Unique", &v)) o.unique = v != 0;
  o.index_base = 1;
  std::vector<uint32_t> i1(n1 > 0 ? n1 : 1), i2(n1 > 0 ? n1 : 1);
  std::vector<float> m(n1 > 0 ? n1 : 1);
  int p = 0;
  if (n1 > 0 && n2 > 0)
    vo_mex_check(vo_match(vo_mex_ctx("vo_match_mex"), (const float*)mxGetData(prhs[0]), n1, (const float*)mxGetData(prhs[1]), n2,
                          d1, /*col_major=*/1, &o, i1.data(), i2.data(), m.data(), &p), "vo:match:cuda");
  plhs[0] = mxCreateNumericMatrix(p, 2, mxUINT32_CLASS, mxREAL);
  uint32_t* out = (uint32_t*)mxGetData(plhs[0]);
  for (int k = 0; k < p; ++k) { out[k] = i1[k]; out[p + k] = i2[k]; }
  if (nlhs > 1) {
    plhs[1] = mxCreateNumericMatrix(p, 1, mxSINGLE_CLASS, mxREAL);
    memcpy(mxGetData(plhs[1]), m.data(), sizeof(float) * p);
  }
}
