/* matlab/stub/mex.h -- TEST INFRASTRUCTURE: prototypes of the handful of MEX / mxArray API
 * functions that matlab/cpk_b200_mex.cpp uses (MATLAB R2018a "interleaved complex" API names:
 * mxGetDoubles etc.), so that the gateway can be compiled, linked and DRIVEN without MATLAB.
 * The matching toy runtime is mex_stub.cpp.  With a real MATLAB the gateway is built against
 * MATLAB's own mex.h instead; nothing here ships to users. */
#ifndef CPK_STUB_MEX_H
#define CPK_STUB_MEX_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef size_t mwSize;
typedef size_t mwIndex;
typedef bool mxLogical;
typedef struct mxArray_tag mxArray;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

bool mxIsSparse(const mxArray *a);
bool mxIsDouble(const mxArray *a);
bool mxIsComplex(const mxArray *a);
bool mxIsChar(const mxArray *a);
bool mxIsNaN(double v);
mwIndex *mxGetJc(const mxArray *a);
mwIndex *mxGetIr(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
double *mxGetDoubles(const mxArray *a);
double mxGetScalar(const mxArray *a);
int mxGetString(const mxArray *a, char *buf, mwSize buflen);
mxArray *mxCreateDoubleScalar(double v);
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray *mxCreateLogicalScalar(bool v);
void mxDestroyArray(mxArray *a);
mxArray *mxDuplicateArray(const mxArray *a);
void mexMakeArrayPersistent(mxArray *a);
bool mxIsClass(const mxArray *a, const char *name);
int mexCallMATLAB(int nlhs, mxArray *plhs[], int nrhs, mxArray *prhs[], const char *fn);
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
int mexAtExit(void (*fn)(void));
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#ifdef __cplusplus
}
#endif
#endif
