// Reference-GPTQ factor (the `--mode gptq` solver, gptq_utils.py:129-165): damped Cholesky of the
// Hessian, its inverse, and the UPPER Cholesky factor of the inverse,
//     L L^T = H + damp * mean(diag H) * I,   H^-1 = L^-T L^-1,   U^T U = H^-1,
// with the reference's damping ladder (damp = 10^e * damp_percent, e = 0..4, first success wins)
// and its identity fallback.  fp64 throughout, column-major inside (a symmetric row-major H is
// its own column-major image, and the lower factor of H^-1 in column-major IS the upper factor in
// row-major, so no transposes are needed).
//   * blocked right-looking Cholesky: 128 x 128 diagonal blocks factored by ONE CTA in shared
//     memory (a non-positive pivot raises a flag: the reference catches torch's RuntimeError),
//     panel by DTRSM, trailing update by DSYRK (lower triangle only), one block column of look-ahead;
//   * L^-1 by DTRSM against the identity, H^-1 = L^-T L^-1 by DSYRK.
#include <cmath>

#include "solver_kernels.cuh"

namespace tq {

constexpr int kChNb = 128;

// Hd (n x n col-major, ld n) = Hsrc[perm, perm] (or Hsrc) ; lower triangle + diagonal are what matter
__global__ void chol_gather_kernel(const double* __restrict__ H, int64_t ldh, int64_t n,
                                   const int64_t* __restrict__ perm, double* __restrict__ out) {
  const int64_t c = blockIdx.y;
  const int64_t pc = perm ? perm[c] : c;
  for (int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < n; r += int64_t(gridDim.x) * blockDim.x) {
    const int64_t pr = perm ? perm[r] : r;
    out[r + c * n] = H[pr * ldh + pc];
  }
}

// scal[0] = mean of the diagonal (1.0 when it is exactly 0, gptq_utils.py:143-145); single CTA
__global__ void __launch_bounds__(1024) chol_mean_diag_kernel(const double* __restrict__ A, int64_t n, double* scal) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += A[i + i * n];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) {
    double m = s / double(n);
    scal[0] = (m == 0.0) ? 1.0 : m;
  }
}

// dst = src with damp * mean added to the diagonal
__global__ void chol_damped_copy_kernel(const double* __restrict__ src, int64_t n, double damp,
                                        const double* __restrict__ scal, double* __restrict__ dst) {
  const int64_t c = blockIdx.y;
  const double add = damp * scal[0];
  for (int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < n; r += int64_t(gridDim.x) * blockDim.x)
    dst[r + c * n] = src[r + c * n] + (r == c ? add : 0.0);
}

// Unblocked Cholesky of one jb x jb diagonal block (lower, in place, col-major ld lda): single CTA of 1024 threads,
// the block lives in shared memory.  Thread (r, g): row r = tid % 128, column group g = tid / 128.  Column c:
// scale it (group 0), then the eight groups sweep the trailing columns cc = c + 1 + g, + 8, ... - no index
// arithmetic beyond additions, rows contiguous in shared memory, two CTA barriers per column (~18 us per block;
// the first version walked the trailing block with a flat index and a division per element: ~200 us).
__global__ void __launch_bounds__(1024) chol_diag_block_kernel(double* __restrict__ A, int64_t lda, int jb, int* fail) {
  extern __shared__ double blk[];   // jb x jb, ld jb
  __shared__ double dg[kChNb];      // the diagonal of the factor (blk keeps the Schur diagonal until the end)
  const int tid = threadIdx.x;
  const int r = tid & (kChNb - 1), g = tid >> 7;
  for (int idx = tid; idx < jb * jb; idx += blockDim.x) blk[idx] = A[(idx % jb) + int64_t(idx / jb) * lda];
  __syncthreads();
  for (int c = 0; c < jb; ++c) {
    const double dcc = blk[c + c * jb];
    if (!(dcc > 0.0)) {              // not positive definite (also catches NaN): same in every thread
      if (tid == 0) *fail = 1;
      return;
    }
    const double rcc = sqrt(dcc);
    const double inv = 1.0 / rcc;
    if (g == 0 && r > c && r < jb) blk[r + c * jb] *= inv;
    if (tid == 0) dg[c] = rcc;
    __syncthreads();
    if (r < jb) {
      const double lrc = blk[r + c * jb];
      for (int cc = c + 1 + g; cc <= r; cc += 8) blk[r + cc * jb] = fma(-lrc, blk[cc + c * jb], blk[r + cc * jb]);
    }
    __syncthreads();
  }
  for (int idx = tid; idx < jb * jb; idx += blockDim.x) {
    const int rr = idx % jb, c = idx / jb;
    if (rr >= c) A[rr + int64_t(c) * lda] = (rr == c) ? dg[c] : blk[idx];
  }
}

// Side stream of the look-ahead (one per host thread and device, created on first use): its own cuBLAS handle and
// two events.  PEDANTIC math like the main handle.
struct CholSide {
  cudaStream_t s = nullptr;
  cublasHandle_t h = nullptr;
  cudaEvent_t ev_col = nullptr, ev_panel = nullptr;
  int dev = -1;
};
static int chol_side(CholSide** out) {
  static thread_local CholSide cs;
  int dev = 0;
  TQ_CUDA_CHECK(cudaGetDevice(&dev));
  if (cs.s == nullptr || cs.dev != dev) {
    // highest priority: the panel's single CTA and its DTRSM must get SMs ahead of the DSYRK that becomes ready at
    // the same moment on the main stream, or the look-ahead only starts when that DSYRK drains
    int prio_lo = 0, prio_hi = 0;
    TQ_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    TQ_CUDA_CHECK(cudaStreamCreateWithPriority(&cs.s, cudaStreamNonBlocking, prio_hi));
    TQ_CUDA_CHECK(cudaEventCreateWithFlags(&cs.ev_col, cudaEventDisableTiming));
    TQ_CUDA_CHECK(cudaEventCreateWithFlags(&cs.ev_panel, cudaEventDisableTiming));
    TQ_CUBLAS_CHECK(cublasCreate(&cs.h));
    TQ_CUBLAS_CHECK(cublasSetMathMode(cs.h, CUBLAS_PEDANTIC_MATH));
    TQ_CUBLAS_CHECK(cublasSetStream(cs.h, cs.s));
    cs.dev = dev;
  }
  *out = &cs;
  return TQ_OK;
}

// in place lower Cholesky of A (n x n col-major, ld n); returns TQ_ERR_NOCONV when not positive definite.
// Right-looking with a look-ahead of one block column: once panel j is there, the main stream first brings block
// column j + 1 up to date (a rem x 128 x 128 DGEMM), then updates the rest of the trailing matrix (DSYRK) while the
// side stream already factors the diagonal block j + 1 and solves for panel j + 1 - the two latency-bound steps
// (~125 + ~70 us) that used to sit between consecutive DSYRKs (~180 us on average at n = 10825).
int chol_lower(cublasHandle_t h, cudaStream_t st, double* A, int64_t n, int* fail) {
  TQ_CUDA_CHECK(cudaFuncSetAttribute(chol_diag_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kChNb * kChNb * 8));
  const double one = 1.0, mone = -1.0;
  TQ_CUDA_CHECK(cudaMemsetAsync(fail, 0, sizeof(int), st));
  const bool ahead = n > 4 * kChNb;
  CholSide* side = nullptr;
  if (ahead) TQ_TRY(chol_side(&side));
  // panel j on stream `ps` with handle `ph`: diagonal block, then A21 <- A21 L11^-T
  auto panel = [&](cublasHandle_t ph, cudaStream_t ps, int64_t j0) -> int {
    const int jb = int(imin(kChNb, n - j0));
    chol_diag_block_kernel<<<1, 1024, size_t(jb) * jb * 8, ps>>>(A + j0 + j0 * n, n, jb, fail);
    TQ_LAUNCH_CHECK();
    const int64_t rem = n - j0 - jb;
    if (rem > 0)
      TQ_CUBLAS_CHECK(cublasDtrsm(ph, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT,
                                  int(rem), jb, &one, A + j0 + j0 * n, int(n), A + (j0 + jb) + j0 * n, int(n)));
    return TQ_OK;
  };
  TQ_TRY(panel(h, st, 0));
  bool panel_on_side = false;
  for (int64_t j0 = 0; j0 < n; j0 += kChNb) {
    const int jb = int(imin(kChNb, n - j0));
    const int64_t j1 = j0 + jb, rem = n - j1;
    if (rem <= 0) break;
    if (panel_on_side) TQ_CUDA_CHECK(cudaStreamWaitEvent(st, side->ev_panel, 0));     // panel j0 is ready
    const double* A21 = A + j1 + j0 * n;                                               // rem x jb
    const int nb1 = int(imin(kChNb, rem));                                             // width of block column j1
    if (ahead && rem > nb1) {
      // block column j1 (its diagonal block and the panel below): A[j1:, j1:j1+nb1] -= A21 A21[0:nb1, :]^T
      TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, int(rem), nb1, jb, &mone, A21, int(n), A21, int(n), &one,
                                  A + j1 + j1 * n, int(n)));
      TQ_CUDA_CHECK(cudaEventRecord(side->ev_col, st));
      TQ_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->ev_col, 0));
      TQ_TRY(panel(side->h, side->s, j1));
      TQ_CUDA_CHECK(cudaEventRecord(side->ev_panel, side->s));
      panel_on_side = true;
      // the rest of the trailing matrix on the main stream: rows and columns from j1 + nb1 on
      const int64_t r2 = rem - nb1;
      TQ_CUBLAS_CHECK(cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, int(r2), jb, &mone, A21 + nb1, int(n), &one,
                                  A + (j1 + nb1) + (j1 + nb1) * n, int(n)));
    } else {
      TQ_CUBLAS_CHECK(cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, int(rem), jb, &mone, A21, int(n), &one,
                                  A + j1 + j1 * n, int(n)));
      panel_on_side = false;
      TQ_TRY(panel(h, st, j1));
    }
  }
  if (panel_on_side) TQ_CUDA_CHECK(cudaStreamWaitEvent(st, side->ev_panel, 0));
  int hfail = 0;
  TQ_CUDA_CHECK(cudaMemcpyAsync(&hfail, fail, sizeof(int), cudaMemcpyDeviceToHost, st));
  TQ_CUDA_CHECK(cudaStreamSynchronize(st));
  return hfail ? TQ_ERR_NOCONV : TQ_OK;
}

__global__ void chol_identity_kernel(double* __restrict__ X, int64_t n) {
  const int64_t c = blockIdx.y;
  for (int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < n; r += int64_t(gridDim.x) * blockDim.x)
    X[r + c * n] = (r == c) ? 1.0 : 0.0;
}

// out (row-major n x n, ld ldo) = upper factor U: U[j, i] = L2[i + j n] for i >= j, 0 below the diagonal
__global__ void chol_emit_upper_kernel(const double* __restrict__ L2, int64_t n, double* __restrict__ out, int64_t ldo) {
  const int64_t j = blockIdx.y;     // row of U = column of L2
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    out[j * ldo + i] = (i >= j) ? L2[i + j * n] : 0.0;
}

}  // namespace tq

using namespace tq;

extern "C" int tq_cholesky_workspace(int64_t n, size_t* bytes) {
  TQ_REQUIRE(bytes && n > 0, "tq_cholesky_workspace: bad arguments");
  *bytes = ws_bytes_for(size_t(n) * n, 8) * 3 + ws_bytes_for(16, 8) * 2 + 4096;
  return TQ_OK;
}

extern "C" int tq_cholesky_solve(const double* H, int64_t ldh, int64_t n, const int64_t* perm, double damp_percent,
                                 double* Hinv_chol, int64_t ldo, int* damp_exp_host, void* ws, size_t ws_bytes,
                                 void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(H && Hinv_chol && damp_exp_host && n > 0 && ldh >= n && ldo >= n && n < (1 << 30),
             "tq_cholesky_solve: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace wsp(ws, ws_bytes);
  double* Hd = wsp.take<double>(size_t(n) * n);     // permuted H
  double* A = wsp.take<double>(size_t(n) * n);      // damped copy -> L -> H^-1 -> L2
  double* X = wsp.take<double>(size_t(n) * n);      // L^-1
  double* scal = wsp.take<double>(16);
  int* fail = wsp.take<int>(4);
  if (wsp.overflow) {
    set_error("tq_cholesky_solve: workspace too small (see tq_cholesky_workspace)");
    return TQ_ERR_WORKSPACE;
  }
  cublasHandle_t h;
  TQ_TRY(get_cublas(&h, st));
  dim3 grid((unsigned)imin(ceil_div(n, 256), 64), (unsigned)n);
  chol_gather_kernel<<<grid, 256, 0, st>>>(H, ldh, n, perm, Hd);
  TQ_LAUNCH_CHECK();
  chol_mean_diag_kernel<<<1, 1024, 0, st>>>(Hd, n, scal);
  TQ_LAUNCH_CHECK();
  const double one = 1.0, zero = 0.0;
  *damp_exp_host = -1;
  for (int e = 0; e < 5; ++e) {
    const double damp = pow(10.0, double(e)) * damp_percent;
    chol_damped_copy_kernel<<<grid, 256, 0, st>>>(Hd, n, damp, scal, A);
    TQ_LAUNCH_CHECK();
    int rc = chol_lower(h, st, A, n, fail);
    if (rc == TQ_ERR_NOCONV) continue;
    TQ_TRY(rc);
    // X = L^-1 ; A = X^T X (lower) = H^-1
    chol_identity_kernel<<<grid, 256, 0, st>>>(X, n);
    TQ_LAUNCH_CHECK();
    TQ_CUBLAS_CHECK(cublasDtrsm(h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, int(n),
                                int(n), &one, A, int(n), X, int(n)));
    TQ_CUBLAS_CHECK(cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, int(n), int(n), &one, X, int(n), &zero, A,
                                int(n)));
    rc = chol_lower(h, st, A, n, fail);
    if (rc == TQ_ERR_NOCONV) continue;
    TQ_TRY(rc);
    chol_emit_upper_kernel<<<grid, 256, 0, st>>>(A, n, Hinv_chol, ldo);
    TQ_LAUNCH_CHECK();
    *damp_exp_host = e;
    return TQ_OK;
  }
  // "Hessian is singular. Using Identity fallback." (gptq_utils.py:161-163)
  for (int64_t r0 = 0; r0 < 1; ++r0) {
    chol_identity_kernel<<<grid, 256, 0, st>>>(X, n);
    TQ_LAUNCH_CHECK();
    chol_emit_upper_kernel<<<grid, 256, 0, st>>>(X, n, Hinv_chol, ldo);
    TQ_LAUNCH_CHECK();
  }
  return TQ_OK;
}
