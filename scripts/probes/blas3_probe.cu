// cuBLAS fp64 rates for the shapes of the band reduction (sy2sb): rank-64 / rank-128 DSYR2K, DSYMM with 64 columns,
// and the DGEMM forms they could be replaced by.  Build (shared cudart, so no runtime symbol table is embedded):
//   nvcc -O2 --cudart shared -gencode arch=compute_100a,code=sm_100a scripts/probes/blas3_probe.cu -lcublas -o /tmp/blas3_probe
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <functional>

static float time_ms(cudaStream_t st, const std::function<void()>& f, int it = 10) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  for (int i = 0; i < it; ++i) f();
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / it;
}

int main() {
  cublasHandle_t h;
  cublasCreate(&h);
  cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
  cudaStream_t st;
  cudaStreamCreate(&st);
  cublasSetStream(h, st);
  const int n = 12288;
  double *A, *V, *W, *X;
  cudaMalloc(&A, sizeof(double) * n * n);
  cudaMalloc(&V, sizeof(double) * n * 256);
  cudaMalloc(&W, sizeof(double) * n * 256);
  cudaMalloc(&X, sizeof(double) * n * 256);
  cudaMemset(A, 0, sizeof(double) * n * n);
  cudaMemset(V, 0, sizeof(double) * n * 256);
  cudaMemset(W, 0, sizeof(double) * n * 256);
  const double one = 1.0, mone = -1.0, zero = 0.0;
  for (int s : {3072, 6144, 12288}) {
    for (int K : {64, 128, 256}) {
      float ms = time_ms(st, [&] { cublasDsyr2k(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, s, K, &mone, V, s, W, s, &one, A, n); });
      printf("s=%5d DSYR2K K=%3d            %7.3f ms  %5.1f TF/s (2 s^2 K flop)\n", s, K, ms, 2.0 * s * s * K / ms / 1e9);
      ms = time_ms(st, [&] { cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, s, K, &mone, V, s, &one, A, n); });
      printf("s=%5d DSYRK  K=%3d            %7.3f ms  %5.1f TF/s (s^2 K flop)\n", s, K, ms, 1.0 * s * s * K / ms / 1e9);
      ms = time_ms(st, [&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, s, s, K, &mone, V, s, W, s, &one, A, n); });
      printf("s=%5d DGEMM  full square K=%3d %7.3f ms  %5.1f TF/s (2 s^2 K flop)\n", s, K, ms, 2.0 * s * s * K / ms / 1e9);
      ms = time_ms(st, [&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, s / 2, s / 2, K, &mone, V + s / 2, s, W, s, &one, A + s / 2, n); });
      printf("s=%5d DGEMM  s/2 x s/2   K=%3d %7.3f ms  %5.1f TF/s\n", s, K, ms, 2.0 * (s / 2) * (s / 2) * K / ms / 1e9);
    }
    for (int N : {64, 128}) {
      float ms = time_ms(st, [&] { cublasDsymm(h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, s, N, &one, A, n, V, s, &zero, X, s); });
      printf("s=%5d DSYMM  N=%3d            %7.3f ms  %5.1f TF/s (2 s^2 N flop)\n", s, N, ms, 2.0 * s * s * N / ms / 1e9);
      ms = time_ms(st, [&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, s, N, s, &one, A, n, V, s, &zero, X, s); });
      printf("s=%5d DGEMM  A(sxs) V(sx%3d)  %7.3f ms  %5.1f TF/s\n", s, N, ms, 2.0 * s * s * N / ms / 1e9);
    }
  }
  return 0;
}
