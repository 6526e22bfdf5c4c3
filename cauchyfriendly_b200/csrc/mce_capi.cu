// mce_capi.cu -- libmce_b200.so: the C ABI of include/mce_b200.h over the CUDA backend (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -lineinfo -shared ... (see build.py).
// There is deliberately no other backend in this translation unit: without a CUDA device mce_create() fails.
#include "backend_cuda.cuh"
#define MCE_BACKEND mce::CudaBackend
#define MCE_VERSION_STRING "mce-b200 0.1 (sm_100a)"
#include "mce_capi_impl.h"
