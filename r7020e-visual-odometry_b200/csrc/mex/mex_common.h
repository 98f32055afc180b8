// mex_common.h -- shared by the four gateways: the process-static vo_ctx (created on first call,
// mexLock()ed, released by a mexAtExit handler), name-value parsing, error forwarding.
#pragma once
#include "mex.h"
#include "../../../include/vo_b200.h"
#include <string.h>
#include <vector>

static vo_ctx* g_ctx = nullptr;
static void vo_mex_atexit(void) { if (g_ctx) { vo_ctx_destroy(g_ctx); g_ctx = nullptr; } }

static vo_ctx* vo_mex_ctx(const char* fn) {
  if (!g_ctx) {
    if (vo_ctx_create(0, &g_ctx) != VO_OK) mexErrMsgIdAndTxt("vo:ctx:create", "%s: %s", fn, vo_last_error());
    mexLock();
    mexAtExit(vo_mex_atexit);
  }
  return g_ctx;
}

static void vo_mex_check(int rc, const char* id) {
  if (rc != VO_OK) mexErrMsgIdAndTxt(id, "%s", vo_last_error());
}

// trailing 'Name', value pairs starting at prhs[first]; returns 1 and the value if `name` is present
static int vo_mex_opt(int nrhs, const mxArray* prhs[], int first, const char* name, double* value) {
  for (int i = first; i + 1 < nrhs; i += 2) {
    char key[64];
    if (!mxIsChar(prhs[i]) || mxGetString(prhs[i], key, sizeof(key)) != 0)
      mexErrMsgIdAndTxt("vo:args:name", "expected a parameter name at argument %d", i + 1);
    if (strcmp(key, name) == 0) { *value = mxGetScalar(prhs[i + 1]); return 1; }
  }
  return 0;
}
// an integer option that may exceed 2^53 (the RNG seed): uint64 / int64 scalars are read exactly
static int vo_mex_opt_u64(int nrhs, const mxArray* prhs[], int first, const char* name, uint64_t* value) {
  for (int i = first; i + 1 < nrhs; i += 2) {
    char key[64];
    if (!mxIsChar(prhs[i]) || mxGetString(prhs[i], key, sizeof(key)) != 0) continue;
    if (strcmp(key, name) != 0) continue;
    const mxClassID c = mxGetClassID(prhs[i + 1]);
    if (c == mxUINT64_CLASS || c == mxINT64_CLASS) *value = *(const uint64_t*)mxGetData(prhs[i + 1]);
    else *value = (uint64_t)mxGetScalar(prhs[i + 1]);
    return 1;
  }
  return 0;
}
static void vo_mex_check_pairs(int nrhs, int first) {
  if ((nrhs - first) % 2 != 0) mexErrMsgIdAndTxt("vo:args:pairs", "name-value arguments must come in pairs");
}
