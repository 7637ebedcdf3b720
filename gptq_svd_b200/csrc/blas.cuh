// cuBLAS handle for the plain library GEMMs (fp64 BLAS-3 inside the solver, the fp32
// diagnostic GEMMs of the error metric).  One handle per host thread.
#pragma once
#include <cublas_v2.h>

#include "common.cuh"

namespace tq {

int get_cublas(cublasHandle_t* out, cudaStream_t stream);

#define TQ_CUBLAS_CHECK(expr)                                                     \
  do {                                                                            \
    cublasStatus_t _s = (expr);                                                   \
    if (_s != CUBLAS_STATUS_SUCCESS) {                                            \
      tq::set_error("%s:%d: %s -> cublas status %d", __FILE__, __LINE__, #expr, int(_s)); \
      return TQ_ERR_CUDA;                                                         \
    }                                                                             \
  } while (0)

}  // namespace tq
