// [A, inlierIdx, status] = vo_p3p_mex(imagePoints, worldPoints, K, 'MaxNumTrials',1000,
//        'Confidence',99,'MaxReprojectionError',1,'Seed',0)
// Drop-in for estworldpose(imagePoints, worldPoints, intrinsics) at VO.m:123-127:
// rel_pose = rigidtform3d(A) feeds VO.m:130 unchanged.  K: [fx fy cx cy] or a 3x3 intrinsic matrix.
// Like estworldpose, errors when the estimate fails and no status output is requested.
#include "mex_common.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 3) mexErrMsgIdAndTxt("vo:p3p:nargin", "vo_p3p_mex(imagePoints, worldPoints, K, ...)");
  if (nlhs > 3) mexErrMsgIdAndTxt("vo:p3p:nargout", "too many outputs");
  if (mxGetClassID(prhs[0]) != mxDOUBLE_CLASS || mxGetClassID(prhs[1]) != mxDOUBLE_CLASS || mxGetClassID(prhs[2]) != mxDOUBLE_CLASS)
    mexErrMsgIdAndTxt("vo:p3p:class", "imagePoints, worldPoints and K must be double");
  const int n = (int)mxGetM(prhs[0]);
  if (mxGetN(prhs[0]) != 2 || mxGetN(prhs[1]) != 3 || (int)mxGetM(prhs[1]) != n)
    mexErrMsgIdAndTxt("vo:p3p:size", "imagePoints must be Nx2 and worldPoints Nx3");
  double K[4];
  const double* k = mxGetPr(prhs[2]);
  if (mxGetNumberOfElements(prhs[2]) == 4) { for (int i = 0; i < 4; ++i) K[i] = k[i]; }
  else if (mxGetM(prhs[2]) == 3 && mxGetN(prhs[2]) == 3) { K[0] = k[0]; K[1] = k[4]; K[2] = k[6]; K[3] = k[7]; }
  else mexErrMsgIdAndTxt("vo:p3p:K", "K must be [fx fy cx cy] or 3x3");
  vo_mex_check_pairs(nrhs, 3);
  vo_p3p_opts o; memset(&o, 0, sizeof(o)); o.adaptive = -1;
  double v;
  if (vo_mex_opt(nrhs, prhs, 3, "MaxNumTrials", &v)) o.max_num_trials = (int)v;
  if (vo_mex_opt(nrhs, prhs, 3, "Confidence", &v)) o.confidence = v;
  if (vo_mex_opt(nrhs, prhs, 3, "MaxReprojectionError", &v)) o.max_reproj_error = v;
  vo_mex_opt_u64(nrhs, prhs, 3, "Seed", &o.seed);
  plhs[0] = mxCreateDoubleMatrix(4, 4, mxREAL);
  mxArray* inl = mxCreateLogicalMatrix(n, 1);
  int status = 0;
  vo_mex_check(vo_p3p(vo_mex_ctx("vo_p3p_mex"), mxGetPr(prhs[0]), mxGetPr(prhs[1]), n, /*col_major=*/1, K, &o, mxGetPr(plhs[0]),
                      (uint8_t*)mxGetData(inl), &status, nullptr), "vo:p3p:cuda");
  if (nlhs < 3 && status != 0)
    mexErrMsgIdAndTxt(status == 1 ? "vo:p3p:notEnoughPts" : "vo:p3p:notEnoughInliers",
                      status == 1 ? "at least 4 points are required" : "not enough inliers");
  if (nlhs > 1) plhs[1] = inl; else mxDestroyArray(inl);
  if (nlhs > 2) { plhs[2] = mxCreateNumericMatrix(1, 1, mxINT32_CLASS, mxREAL); *(int32_t*)mxGetData(plhs[2]) = status; }
}
