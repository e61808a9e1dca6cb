// static_check.cu -- compiles rule_kernels.cuh with nvcc for one configuration so that register
// counts (-Xptxas -v) and SASS (cuobjdump) can be inspected without a GPU.  The production path
// compiles the same header with NVRTC at run time (jit.cpp); this file is not linked into the library.
#ifndef CUB_DIM
#define CUB_DIM 6
#endif
#ifndef CUB_FDIM
#define CUB_FDIM 1
#endif
#ifndef CUB_NPARAMS
#define CUB_NPARAMS 12
#endif
#ifndef CUB_HAS_TRANSFORM
#define CUB_HAS_TRANSFORM 0
#endif
#ifndef CUB_INTEGRAND
#define CUB_INTEGRAND GenzCornerPeak<CUB_DIM>
#endif
#ifndef CUB_FAST_RCP
#define CUB_FAST_RCP 1
#endif
#include "rule_kernels.cuh"
