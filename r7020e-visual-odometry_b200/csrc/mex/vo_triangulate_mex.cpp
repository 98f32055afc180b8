// [xyz, reprojErr, valid] = vo_triangulate_mex(pts1, pts2, P1, P2)
// Drop-in for triangulate(p_l, p_r, p1, p2) at VO.m:114-115 and CreateLandmarksFromFeatures.m:7.
// pts: Nx2 single or double (1x2 in the reference's loop; the batched Nx2 form replaces the loop
// VO.m:113-116 by two calls).  P: 3x4 double camProjection (or the legacy 4x3 camMatrix).
#include "mex_common.h"

static void get_P(const mxArray* a, double P[12]) {
  if (mxGetClassID(a) != mxDOUBLE_CLASS) mexErrMsgIdAndTxt("vo:triangulate:class", "projection matrices must be double");
  const double* p = mxGetPr(a);
  if (mxGetM(a) == 3 && mxGetN(a) == 4) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) P[4 * r + c] = p[c * 3 + r]; }
  else if (mxGetM(a) == 4 && mxGetN(a) == 3) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) P[4 * r + c] = p[r * 4 + c]; }
  else mexErrMsgIdAndTxt("vo:triangulate:size", "projection matrix must be 3x4 (or 4x3 camMatrix)");
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs != 4) mexErrMsgIdAndTxt("vo:triangulate:nargin", "vo_triangulate_mex(pts1, pts2, P1, P2)");
  if (nlhs > 3) mexErrMsgIdAndTxt("vo:triangulate:nargout", "too many outputs");
  const mxClassID cls = mxGetClassID(prhs[0]);
  if ((cls != mxSINGLE_CLASS && cls != mxDOUBLE_CLASS) || mxGetClassID(prhs[1]) != cls)
    mexErrMsgIdAndTxt("vo:triangulate:class", "points must both be single or both be double");
  const int n = (int)mxGetM(prhs[0]);
  if (mxGetN(prhs[0]) != 2 || mxGetN(prhs[1]) != 2 || (int)mxGetM(prhs[1]) != n)
    mexErrMsgIdAndTxt("vo:triangulate:size", "points must be Nx2 with equal N");
  double P1[12], P2[12];
  get_P(prhs[2], P1); get_P(prhs[3], P2);
  plhs[0] = mxCreateNumericMatrix(n, 3, cls, mxREAL);
  mxArray* err = mxCreateNumericMatrix(n, 1, cls, mxREAL);
  mxArray* valid = mxCreateLogicalMatrix(n, 1);
  if (n > 0)
    vo_mex_check(vo_triangulate(vo_mex_ctx("vo_triangulate_mex"), mxGetData(prhs[0]), mxGetData(prhs[1]), n, cls == mxDOUBLE_CLASS,
                                /*col_major=*/1, P1, P2, mxGetData(plhs[0]), mxGetData(err), (uint8_t*)mxGetData(valid)),
                 "vo:triangulate:cuda");
  if (nlhs > 1) plhs[1] = err; else mxDestroyArray(err);
  if (nlhs > 2) plhs[2] = valid; else mxDestroyArray(valid);
}
