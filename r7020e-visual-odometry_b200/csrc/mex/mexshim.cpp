// mexshim.cpp -- a tiny fake MATLAB host: implements the mx*/mex* functions of mex.h and a C entry
// point that calls a gateway's mexFunction with error trapping.  Tests drive it through ctypes.
#include "mex.h"
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

// m = dims[0], n = product of the remaining dimensions (what MATLAB's mxGetN returns for N-D arrays)
struct mxArray_tag { mxClassID cls; size_t m, n; void* data; size_t ndim; size_t dims[4]; };

static size_t elem_size(mxClassID c) {
  switch (c) {
    case mxDOUBLE_CLASS: case mxINT64_CLASS: case mxUINT64_CLASS: return 8;
    case mxSINGLE_CLASS: case mxINT32_CLASS: case mxUINT32_CLASS: return 4;
    case mxINT16_CLASS: case mxUINT16_CLASS: case mxCHAR_CLASS: return 2;
    default: return 1;
  }
}
static jmp_buf g_jmp; static int g_jmp_set = 0;
static char g_err_id[128], g_err_msg[1024];
static void (*g_atexit)(void) = nullptr;
static int g_locked = 0;

extern "C" {
size_t mxGetM(const mxArray* a) { return a->m; }
size_t mxGetN(const mxArray* a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray* a) { return a->m * a->n; }
mwSize mxGetNumberOfDimensions(const mxArray* a) { return a->ndim; }
const mwSize* mxGetDimensions(const mxArray* a) { return a->dims; }
void* mxGetData(const mxArray* a) { return a->data; }
double* mxGetPr(const mxArray* a) { return (double*)a->data; }
mxClassID mxGetClassID(const mxArray* a) { return a->cls; }
int mxIsChar(const mxArray* a) { return a->cls == mxCHAR_CLASS; }
int mxIsEmpty(const mxArray* a) { return a->m * a->n == 0; }
double mxGetScalar(const mxArray* a) {
  switch (a->cls) {
    case mxDOUBLE_CLASS: return *(double*)a->data;
    case mxSINGLE_CLASS: return *(float*)a->data;
    case mxINT32_CLASS: return *(int32_t*)a->data;
    case mxUINT32_CLASS: return *(uint32_t*)a->data;
    case mxINT64_CLASS: return (double)*(int64_t*)a->data;
    case mxUINT64_CLASS: return (double)*(uint64_t*)a->data;
    case mxLOGICAL_CLASS: case mxUINT8_CLASS: return *(uint8_t*)a->data;
    default: return 0;
  }
}
int mxGetString(const mxArray* a, char* buf, mwSize buflen) {
  const size_t n = a->m * a->n;
  if (a->cls != mxCHAR_CLASS || n + 1 > buflen) return 1;
  for (size_t i = 0; i < n; ++i) buf[i] = (char)((uint16_t*)a->data)[i];
  buf[n] = 0;
  return 0;
}
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity) {
  mxArray* a = (mxArray*)malloc(sizeof(mxArray));
  a->cls = cls; a->m = m; a->n = n; a->ndim = 2; a->dims[0] = m; a->dims[1] = n; a->dims[2] = a->dims[3] = 1;
  a->data = calloc((m * n) != 0 ? m * n : 1, elem_size(cls));
  return a;
}
mxArray* mxCreateNumericArray(mwSize ndim, const mwSize* dims, mxClassID cls, mxComplexity c) {
  size_t rest = 1;
  for (size_t k = 1; k < ndim; ++k) rest *= dims[k];
  mxArray* a = mxCreateNumericMatrix(ndim ? dims[0] : 0, ndim > 1 ? rest : 1, cls, c);
  a->ndim = ndim < 2 ? 2 : (ndim > 4 ? 4 : ndim);
  for (size_t k = 0; k < 4; ++k) a->dims[k] = k < ndim ? dims[k] : 1;
  return a;
}
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) { return mxCreateNumericMatrix(m, n, mxDOUBLE_CLASS, c); }
mxArray* mxCreateLogicalMatrix(mwSize m, mwSize n) { return mxCreateNumericMatrix(m, n, mxLOGICAL_CLASS, mxREAL); }
mxArray* mxCreateString(const char* s) {
  const size_t n = strlen(s);
  mxArray* a = mxCreateNumericMatrix(1, n, mxCHAR_CLASS, mxREAL);
  for (size_t i = 0; i < n; ++i) ((uint16_t*)a->data)[i] = (uint16_t)s[i];
  return a;
}
void mxDestroyArray(mxArray* a) { if (a) { free(a->data); free(a); } }
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_err_msg, sizeof(g_err_msg), fmt, ap);
  va_end(ap);
  snprintf(g_err_id, sizeof(g_err_id), "%s", id);
  if (g_jmp_set) longjmp(g_jmp, 1);
  fprintf(stderr, "%s: %s\n", g_err_id, g_err_msg);
  abort();
}
void mexLock(void) { g_locked = 1; }
int mexAtExit(void (*fn)(void)) { g_atexit = fn; return 0; }

// ---- host-side helpers for the test driver
mxArray* shim_from_buffer(int cls, size_t m, size_t n, const void* src) {
  mxArray* a = mxCreateNumericMatrix(m, n, (mxClassID)cls, mxREAL);
  if (src && m * n != 0) memcpy(a->data, src, m * n * elem_size((mxClassID)cls));
  return a;
}
// N-D array (up to 4 dimensions, column-major like MATLAB)
mxArray* shim_from_buffer_nd(int cls, size_t ndim, const size_t* dims, const void* src) {
  mxArray* a = mxCreateNumericArray(ndim, dims, (mxClassID)cls, mxREAL);
  if (src && a->m * a->n != 0) memcpy(a->data, src, a->m * a->n * elem_size((mxClassID)cls));
  return a;
}
typedef void (*mexfn_t)(int, mxArray**, int, const mxArray**);
// returns 0 on success, 1 if the gateway raised (id/message via shim_last_error)
int shim_call(mexfn_t fn, int nlhs, mxArray** plhs, int nrhs, const mxArray** prhs) {
  g_err_id[0] = g_err_msg[0] = 0;
  g_jmp_set = 1;
  int rc = 0;
  if (setjmp(g_jmp) == 0) fn(nlhs, plhs, nrhs, prhs);
  else rc = 1;
  g_jmp_set = 0;
  return rc;
}
const char* shim_last_error_id(void) { return g_err_id; }
const char* shim_last_error_msg(void) { return g_err_msg; }
void shim_run_atexit(void) { if (g_atexit) { g_atexit(); g_atexit = nullptr; } }
int shim_is_locked(void) { return g_locked; }
}
