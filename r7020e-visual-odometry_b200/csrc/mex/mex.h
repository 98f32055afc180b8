/*
 * mex.h -- stand-in for MATLAB's <mex.h>/<matrix.h>, only so the gateways compile and can be driven
 * by a fake host in this MATLAB-less container.  It declares the subset of the
 * published MEX C API the four gateways use, with MATLAB's own names, signatures and class IDs.
 * With a real MATLAB, build with `mex` and MathWorks' header instead; this file is then unused.
 */
#ifndef VO_MEX_SHIM_H
#define VO_MEX_SHIM_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef unsigned char mxLogical;
typedef enum {
  mxUNKNOWN_CLASS = 0, mxCELL_CLASS, mxSTRUCT_CLASS, mxLOGICAL_CLASS, mxCHAR_CLASS, mxVOID_CLASS,
  mxDOUBLE_CLASS, mxSINGLE_CLASS, mxINT8_CLASS, mxUINT8_CLASS, mxINT16_CLASS, mxUINT16_CLASS,
  mxINT32_CLASS, mxUINT32_CLASS, mxINT64_CLASS, mxUINT64_CLASS
} mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX } mxComplexity;

size_t mxGetM(const mxArray* a);
size_t mxGetN(const mxArray* a);
size_t mxGetNumberOfElements(const mxArray* a);
mwSize mxGetNumberOfDimensions(const mxArray* a);
const mwSize* mxGetDimensions(const mxArray* a);
void* mxGetData(const mxArray* a);
double* mxGetPr(const mxArray* a);
double mxGetScalar(const mxArray* a);
mxClassID mxGetClassID(const mxArray* a);
int mxIsChar(const mxArray* a);
int mxIsEmpty(const mxArray* a);
int mxGetString(const mxArray* a, char* buf, mwSize buflen);
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
mxArray* mxCreateNumericArray(mwSize ndim, const mwSize* dims, mxClassID cls, mxComplexity c);
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray* mxCreateLogicalMatrix(mwSize m, mwSize n);
mxArray* mxCreateString(const char* s);
void mxDestroyArray(mxArray* a);
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);   /* does not return */
void mexLock(void);
int mexAtExit(void (*fn)(void));

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
