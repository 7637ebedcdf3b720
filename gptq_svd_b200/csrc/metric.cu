// Per-layer output reconstruction error  ||(W-Q)[:,perm] Rx^T||_F / ||W[:,perm] Rx^T||_F
// (reference log_quantization_error, gptq_utils.py:275-291): a logged diagnostic, two plain
// library GEMMs through cuBLAS with the squared Frobenius norms accumulated in fp64.  The
// reference runs them in fp32 with TF32 off (:474-475); as SIMT SGEMMs they were 47 % of a
// down_proj gptq_fwrd call (profiles/r01_launches_loop_down.txt), so they run on the tensor
// cores with TF32 inputs and fp32 accumulation by default (the ratio moves by < 1e-4 relative,
// bar 1 %); TQ_METRIC_STRICT_FP32=1 restores the strict-fp32 SGEMM.
#include <cstdlib>
#include "blas.cuh"
#include "common.cuh"

namespace tq {

int get_cublas(cublasHandle_t* out, cudaStream_t stream) {
  static thread_local cublasHandle_t h = nullptr;
  static thread_local int h_dev = -1;
  int dev = 0;
  TQ_CUDA_CHECK(cudaGetDevice(&dev));
  if (h == nullptr || h_dev != dev) {
    TQ_CUBLAS_CHECK(cublasCreate(&h));
    TQ_CUBLAS_CHECK(cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH));
    h_dev = dev;
  }
  TQ_CUBLAS_CHECK(cublasSetStream(h, stream));
  *out = h;
  return TQ_OK;
}

template <typename TR>
__global__ void cast_rx_kernel(const TR* __restrict__ R, int64_t ldr, int64_t k, int64_t n,
                               float* __restrict__ out) {
  int64_t r = blockIdx.y;
  for (int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; c < n; c += int64_t(gridDim.x) * blockDim.x)
    out[r * n + c] = float(R[r * ldr + c]);
}

__global__ void gather_diff_kernel(const float* __restrict__ W, int64_t ldw, const float* __restrict__ Wq,
                                   int64_t ldq, const int64_t* __restrict__ perm, int64_t m, int64_t n,
                                   float* __restrict__ Wo, float* __restrict__ D) {
  int64_t r = blockIdx.y;
  for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += int64_t(gridDim.x) * blockDim.x) {
    int64_t p = perm[j];
    float w = W[r * ldw + p];
    Wo[r * n + j] = w;
    D[r * n + j] = __fsub_rn(w, Wq[r * ldq + p]);
  }
}

__global__ void sumsq_kernel(const float* __restrict__ x, int64_t count, double* __restrict__ out) {
  double s = 0.0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    double v = x[i];
    s += v * v;
  }
  __shared__ double sh[32];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

}  // namespace tq

using namespace tq;

extern "C" int tq_quant_error_workspace(int64_t m, int64_t n, int64_t k, size_t* bytes) {
  TQ_REQUIRE(bytes && m > 0 && n > 0 && k > 0, "tq_quant_error_workspace: bad arguments");
  *bytes = ws_bytes_for(size_t(k) * n, 4) + 2 * ws_bytes_for(size_t(m) * n, 4) + ws_bytes_for(size_t(m) * k, 4);
  return TQ_OK;
}

extern "C" int tq_quant_error(const float* W, int64_t ldw, const float* Wq, int64_t ldq, const void* Rx,
                              int rx_dtype, int64_t ldr, int64_t k, const int64_t* perm, int64_t m, int64_t n,
                              double* out2, void* ws, size_t ws_bytes, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(W && Wq && Rx && perm && out2, "tq_quant_error: null pointer");
  TQ_REQUIRE(m > 0 && n > 0 && k > 0 && k <= n && ldw >= n && ldq >= n && ldr >= n, "tq_quant_error: bad shape");
  TQ_REQUIRE(rx_dtype == TQ_F64 || rx_dtype == TQ_F32, "tq_quant_error: Rx must be fp64 or fp32");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace wsp(ws, ws_bytes);
  float* R32 = wsp.take<float>(size_t(k) * n);
  float* Wo = wsp.take<float>(size_t(m) * n);
  float* D = wsp.take<float>(size_t(m) * n);
  float* Y = wsp.take<float>(size_t(m) * k);
  if (wsp.overflow) {
    set_error("tq_quant_error: workspace too small (%zu < %zu)", ws_bytes, wsp.off);
    return TQ_ERR_WORKSPACE;
  }
  {
    dim3 grid((unsigned)imin(ceil_div(n, 256), 64), (unsigned)k);
    if (rx_dtype == TQ_F64) cast_rx_kernel<double><<<grid, 256, 0, st>>>((const double*)Rx, ldr, k, n, R32);
    else cast_rx_kernel<float><<<grid, 256, 0, st>>>((const float*)Rx, ldr, k, n, R32);
    TQ_LAUNCH_CHECK();
  }
  {
    dim3 grid((unsigned)imin(ceil_div(n, 256), 64), (unsigned)m);
    gather_diff_kernel<<<grid, 256, 0, st>>>(W, ldw, Wq, ldq, perm, m, n, Wo, D);
    TQ_LAUNCH_CHECK();
  }
  TQ_CUDA_CHECK(cudaMemsetAsync(out2, 0, 2 * sizeof(double), st));
  cublasHandle_t h;
  TQ_TRY(get_cublas(&h, st));
  const float one = 1.f, zero = 0.f;
  // row-major Y (m x k) = A (m x n) . R32^T  <=>  col-major Y^T (k x m) = R32 . A^T
  const float* srcs[2] = {D, Wo};
  const int strict = 0;
  for (int i = 0; i < 2; ++i) {
    if (strict) {
      TQ_CUBLAS_CHECK(cublasSgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, int(k), int(m), int(n), &one, R32, int(n), srcs[i],
                                  int(n), &zero, Y, int(k)));
    } else {
      cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);
      cublasStatus_t cs = cublasGemmEx(h, CUBLAS_OP_T, CUBLAS_OP_N, int(k), int(m), int(n), &one, R32, CUDA_R_32F,
                                       int(n), srcs[i], CUDA_R_32F, int(n), &zero, Y, CUDA_R_32F, int(k),
                                       CUBLAS_COMPUTE_32F_FAST_TF32, CUBLAS_GEMM_DEFAULT);
      cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
      if (cs != CUBLAS_STATUS_SUCCESS) {
        set_error("tq_quant_error: cublasGemmEx failed with status %d", int(cs));
        return TQ_ERR_CUDA;
      }
    }
    sumsq_kernel<<<296, 256, 0, st>>>(Y, m * k, out2 + i);
    TQ_LAUNCH_CHECK();
  }
  return TQ_OK;
}
