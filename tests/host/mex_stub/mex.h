/* Minimal declarations of the MATLAB mex API, so that the gateway sources can be compiled in a container
 * without MATLAB: syntax check in tests/test_abi.py, and the builds of tests/host/Makefile that run against mex_mock/.  Not a MATLAB header and never shipped. */
#pragma once
#include <stddef.h>
typedef struct mxArray_tag mxArray;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
extern "C" {
bool mxIsChar(const mxArray*); int mxGetString(const mxArray*, char*, size_t);
bool mxIsStruct(const mxArray*); bool mxIsDouble(const mxArray*); bool mxIsComplex(const mxArray*); bool mxIsEmpty(const mxArray*);
mxArray* mxGetField(const mxArray*, size_t, const char*); size_t mxGetNumberOfElements(const mxArray*);
double* mxGetPr(const mxArray*); double mxGetScalar(const mxArray*);
mxArray* mxCreateDoubleMatrix(size_t, size_t, mxComplexity); mxArray* mxCreateDoubleScalar(double);
mxArray* mxCreateStructMatrix(size_t, size_t, int, const char**); void mxSetFieldByNumber(mxArray*, size_t, int, mxArray*);
void mexErrMsgIdAndTxt(const char*, const char*, ...); int mexAtExit(void (*)(void));
/* like MATLAB's mex.h: the entry point has C linkage */
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
}
