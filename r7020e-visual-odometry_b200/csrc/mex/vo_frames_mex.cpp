// [relA, status, counts] = vo_frames_mex(L, R, P1, P2, 'Seed',0, 'FirstFrame',0, 'MaxKeypoints',8192,
//        'MaxNumTrials',1000, 'Confidence',99, 'MaxReprojectionError',1, 'MaxRatio',0.6, 'MatchThreshold',1)
// The batched drop-in for the body of `for i = 1:n_frames` (VO.m:64-232): L and R are H x W x N uint8 stacks of N
// consecutive left / right frames (as readimage returns them, concatenated along the third dimension), P1 / P2 the
// 3x4 projection matrices of VO.m:30-33 (4x3 legacy camMatrix accepted).  Frame 1 of the stack only seeds the tracker
// (VO.m:207-210); for every later frame k the gateway returns what VO.m:123-127 computes:
//   relA(:,:,k)  4x4 double, rel_pose = rigidtform3d(relA(:,:,k)) feeds VO.m:130 unchanged (identity for k = 1)
//   status(k)    int32 estworldpose status (0 ok, 1 fewer than 4 points, 2 not enough inliers)
//   counts(:,k)  int32 [N_L N_R K0 K1 K2 K3 K4 inliers]: keypoints, stereo matches, the four chained matches
// To stream a sequence call with overlapping stacks: frames [i0-1, i0+B) and 'FirstFrame', i0-1.
// One H2D copy of the stacks (transposed on the device), one synchronisation, no CPU fallback.
#include "mex_common.h"
#include <stdlib.h>

static void read_P(const mxArray* a, double P[12], const char* nm) {
  if (mxGetClassID(a) != mxDOUBLE_CLASS) mexErrMsgIdAndTxt("vo:frames:class", "%s must be double", nm);
  const double* p = mxGetPr(a);
  if (mxGetM(a) == 3 && mxGetN(a) == 4) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) P[r * 4 + c] = p[c * 3 + r]; }
  else if (mxGetM(a) == 4 && mxGetN(a) == 3) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) P[r * 4 + c] = p[r * 4 + c]; }
  else mexErrMsgIdAndTxt("vo:frames:size", "%s must be 3x4 (or 4x3 camMatrix)", nm);
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 4) mexErrMsgIdAndTxt("vo:frames:nargin", "vo_frames_mex(L, R, P1, P2, ...)");
  if (nlhs > 3) mexErrMsgIdAndTxt("vo:frames:nargout", "too many outputs");
  if (mxGetClassID(prhs[0]) != mxUINT8_CLASS || mxGetClassID(prhs[1]) != mxUINT8_CLASS)
    mexErrMsgIdAndTxt("vo:frames:class", "L and R must be uint8 (H x W x N)");
  const mwSize nd = mxGetNumberOfDimensions(prhs[0]);
  const mwSize* d = mxGetDimensions(prhs[0]);
  const mwSize* dr = mxGetDimensions(prhs[1]);
  if (nd < 2 || nd > 3 || mxGetNumberOfDimensions(prhs[1]) != nd) mexErrMsgIdAndTxt("vo:frames:size", "L and R must be H x W x N");
  const int rows = (int)d[0], cols = (int)d[1], n = nd == 3 ? (int)d[2] : 1;
  if (rows < 1 || cols < 1 || n < 1) mexErrMsgIdAndTxt("vo:frames:empty", "empty image stack");
  for (mwSize k = 0; k < nd; ++k)
    if (d[k] != dr[k]) mexErrMsgIdAndTxt("vo:frames:size", "L and R differ in size");
  double P1[12], P2[12];
  read_P(prhs[2], P1, "P1"); read_P(prhs[3], P2, "P2");
  vo_mex_check_pairs(nrhs, 4);
  vo_frames_opts o; memset(&o, 0, sizeof(o));
  o.sift.index_base = 1;     // MATLAB Location
  o.p3p.adaptive = -1;
  o.col_major = 1;
  double v;
  vo_mex_opt_u64(nrhs, prhs, 4, "Seed", &o.p3p.seed);
  if (vo_mex_opt(nrhs, prhs, 4, "FirstFrame", &v)) o.first_frame = (int)v;
  if (vo_mex_opt(nrhs, prhs, 4, "MaxKeypoints", &v)) o.max_keypoints = (int)v;
  if (vo_mex_opt(nrhs, prhs, 4, "MaxNumTrials", &v)) o.p3p.max_num_trials = (int)v;
  if (vo_mex_opt(nrhs, prhs, 4, "Confidence", &v)) o.p3p.confidence = v;
  if (vo_mex_opt(nrhs, prhs, 4, "MaxReprojectionError", &v)) o.p3p.max_reproj_error = v;
  if (vo_mex_opt(nrhs, prhs, 4, "MaxRatio", &v)) o.match.max_ratio = (float)v;
  if (vo_mex_opt(nrhs, prhs, 4, "MatchThreshold", &v)) o.match.match_threshold = (float)v;
  if (vo_mex_opt(nrhs, prhs, 4, "ContrastThreshold", &v)) o.sift.contrast_threshold = (float)v;
  if (vo_mex_opt(nrhs, prhs, 4, "EdgeThreshold", &v)) o.sift.edge_threshold = (float)v;
  if (vo_mex_opt(nrhs, prhs, 4, "NumLayersInOctave", &v)) o.sift.num_layers_in_octave = (int)v;
  if (vo_mex_opt(nrhs, prhs, 4, "Sigma", &v)) o.sift.sigma = (float)v;
  const mwSize da[3] = {4, 4, (mwSize)n};
  plhs[0] = mxCreateNumericArray(3, da, mxDOUBLE_CLASS, mxREAL);
  mxArray* st = mxCreateNumericMatrix(n, 1, mxINT32_CLASS, mxREAL);
  mxArray* cn = mxCreateNumericMatrix(8, n, mxINT32_CLASS, mxREAL);
  double* A = mxGetPr(plhs[0]);
  vo_ctx* ctx = vo_mex_ctx("vo_frames_mex");
  {   // a MATLAB loop feeds one stack at a time on one stream: replay the launch sequence as a CUDA graph (bit-identical;
      // VO_FRAMES_GRAPH=0 in the environment keeps plain launches)
    static const bool graph = [] { const char* e = getenv("VO_FRAMES_GRAPH"); return !(e && atoi(e) == 0); }();
    if (graph && vo_frames_graph_state(ctx) == 0) vo_frames_use_graph(ctx, 1);
  }
  vo_mex_check(vo_frames(ctx, (const uint8_t*)mxGetData(prhs[0]), (const uint8_t*)mxGetData(prhs[1]), n,
                         rows, cols, P1, P2, &o, A, (int*)mxGetData(st), (int*)mxGetData(cn)), "vo:frames:cuda");
  for (int k = 0; k < n; ++k) {           // row-major 4x4 -> MATLAB column-major
    double* a = A + 16 * k;
    for (int r = 0; r < 4; ++r)
      for (int c = r + 1; c < 4; ++c) { const double t = a[r * 4 + c]; a[r * 4 + c] = a[c * 4 + r]; a[c * 4 + r] = t; }
  }
  if (nlhs > 1) plhs[1] = st; else mxDestroyArray(st);
  if (nlhs > 2) plhs[2] = cn; else mxDestroyArray(cn);
}
