// Host emulation of the WHOLE two-stage path of gptq_svd_b200/csrc/two_stage.cu (test infrastructure, CPU only):
// the host driver (panel loop of the band reduction, wavefront schedule of the Q2 back-transformation, Q1) is
// compiled unchanged with g++, its kernels run on the emulation runtime (emu_runtime.h), and the few cuBLAS / CUDA
// runtime entry points it calls are defined HERE as plain column-major reference loops on host memory - nothing of
// libcublas / libcudart is linked.  The cluster QR panel kernel (qr.cu) needs thread-block clusters and is replaced
// by a Householder QR with the same in-place LAPACK layout.  tests/test_two_stage_emu.py drives this.
#include "emu_runtime.h"

#include <cstdlib>
#include <cublas_v2.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../gptq_svd_b200/csrc/two_stage.cu"

// ------------------------------------------------------------------ what two_stage.cu expects from its siblings
namespace tq {
static char g_err_buf[512] = "";
static int g_emu_sms = 3;
thread_local int64_t g_launch_count = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err_buf, sizeof(g_err_buf), fmt, ap);
  va_end(ap);
}
int num_sms() { return g_emu_sms; }
bool trace_enabled() { return false; }
int two_stage_setting() { return 1; }
int prof_begin_launch(cudaStream_t, double, int) { return -1; }
void prof_end_launch(cudaStream_t, int) {}
bool stage_callback_set() { return false; }
void notify_stage(int) {}
StageTimer::StageTimer(cudaStream_t s, const char* n) : st(s), name(n), t0(0) {}
StageTimer::~StageTimer() {}

int check_device() { return TQ_OK; }
int get_cublas(cublasHandle_t* out, cudaStream_t) {
  *out = nullptr;
  return TQ_OK;
}
// eigh.cu: A (column-major, both triangles) = the symmetric matrix defined by the lower triangle of the row-major H
int copy_symmetric_lower(cudaStream_t, const double* H, int64_t ldh, int64_t n, double* A) {
  for (int64_t c = 0; c < n; ++c)
    for (int64_t r = 0; r < n; ++r) A[r + c * n] = (r >= c) ? H[r * ldh + c] : H[c * ldh + r];
  return TQ_OK;
}

// stands in for qr_r_colmajor_tau (qr.cu): unblocked Householder QR, R on / above the diagonal, reflector tails
// below it (unit diagonal implied), tau_out[j]
int qr_r_colmajor_tau(cublasHandle_t, cudaStream_t, double* A, int64_t lda, int64_t k, int64_t n, Workspace&,
                      double* tau_out) {
  const int64_t kk = k < n ? k : n;
  for (int64_t j = 0; j < kk; ++j) {
    double* x = A + j + j * lda;
    const int64_t len = k - j;
    double ss = 0.0;
    for (int64_t i = 1; i < len; ++i) ss += x[i] * x[i];
    double tau = 0.0, beta = x[0], scl = 0.0;
    if (len > 1 && ss != 0.0) {
      beta = -copysign(hypot(x[0], sqrt(ss)), x[0]);
      tau = (beta - x[0]) / beta;
      scl = 1.0 / (x[0] - beta);
    }
    for (int64_t i = 1; i < len; ++i) x[i] *= scl;
    x[0] = 1.0;
    for (int64_t c = j + 1; c < n; ++c) {
      double* y = A + j + c * lda;
      double w = 0.0;
      for (int64_t i = 0; i < len; ++i) w += x[i] * y[i];
      w *= tau;
      for (int64_t i = 0; i < len; ++i) y[i] -= w * x[i];
    }
    x[0] = beta;
    tau_out[j] = tau;
  }
  return TQ_OK;
}
}  // namespace tq

// ------------------------------------------------------------------ CUDA runtime on host memory
extern "C" {
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
  memset(p, v, n);
  return cudaSuccess;
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t) {
  memcpy(dst, src, n);
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned int) {
  *p = std::malloc(n);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void* p) {
  std::free(p);
  return cudaSuccess;
}
// streams and events: every operation of the emulation runs at once and in program order, so a side stream is only
// a tag here - the look-ahead's ARITHMETIC is checked, its two event dependencies are argued in two_stage.cu
cudaError_t cudaGetDevice(int* dev) {
  *dev = 0;
  return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned int) {
  *s = reinterpret_cast<cudaStream_t>(0x1);
  return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned int) {
  *e = reinterpret_cast<cudaEvent_t>(0x1);
  return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }
cublasStatus_t cublasCreate_v2(cublasHandle_t* h) {
  *h = nullptr;
  return CUBLAS_STATUS_SUCCESS;
}
cublasStatus_t cublasSetMathMode(cublasHandle_t, cublasMath_t) { return CUBLAS_STATUS_SUCCESS; }
cublasStatus_t cublasSetStream_v2(cublasHandle_t, cudaStream_t) { return CUBLAS_STATUS_SUCCESS; }
const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaLaunchCooperativeKernel(const void* func, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t) {
  tq::ChaseArgs a = *static_cast<tq::ChaseArgs*>(args[0]);
  if (func == (const void*)tq::sb2st_chase_kernel_t<false, false>)
    emu_run(grid, block, smem, /*concurrent=*/true, [&] { tq::sb2st_chase_kernel_t<false, false>(a); });
  else if (func == (const void*)tq::sb2st_chase_kernel_t<true, false>)
    emu_run(grid, block, smem, /*concurrent=*/true, [&] { tq::sb2st_chase_kernel_t<true, false>(a); });
  else if (func == (const void*)tq::sb2st_chase_kernel_t<false, true>)
    emu_run(grid, block, smem, /*concurrent=*/true, [&] { tq::sb2st_chase_kernel_t<false, true>(a); });
  else if (func == (const void*)tq::sb2st_chase_kernel_t<true, true>)
    emu_run(grid, block, smem, /*concurrent=*/true, [&] { tq::sb2st_chase_kernel_t<true, true>(a); });
  else
    return cudaErrorInvalidDeviceFunction;
  return cudaSuccess;
}

// ------------------------------------------------------------------ reference BLAS (column-major)
static inline double opel(const double* A, int lda, cublasOperation_t t, int i, int j) {
  return t == CUBLAS_OP_N ? A[i + size_t(j) * lda] : A[j + size_t(i) * lda];
}
cublasStatus_t cublasDgemm_v2(cublasHandle_t, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                              const double* alpha, const double* A, int lda, const double* B, int ldb,
                              const double* beta, double* C, int ldc) {
  if (lda < (ta == CUBLAS_OP_N ? m : k) || ldb < (tb == CUBLAS_OP_N ? k : n) || ldc < m) return CUBLAS_STATUS_INVALID_VALUE;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) {
      double s = 0.0;
      for (int p = 0; p < k; ++p) s += opel(A, lda, ta, i, p) * opel(B, ldb, tb, p, j);
      double& c = C[i + size_t(j) * ldc];
      c = (*beta == 0.0 ? 0.0 : *beta * c) + *alpha * s;
    }
  return CUBLAS_STATUS_SUCCESS;
}
cublasStatus_t cublasDgemmStridedBatched(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                                         const double* alpha, const double* A, int lda, long long sa, const double* B,
                                         int ldb, long long sb, const double* beta, double* C, int ldc, long long sc,
                                         int count) {
  for (int i = 0; i < count; ++i) {
    cublasStatus_t s = cublasDgemm_v2(h, ta, tb, m, n, k, alpha, A + i * sa, lda, B + i * sb, ldb, beta, C + i * sc, ldc);
    if (s != CUBLAS_STATUS_SUCCESS) return s;
  }
  return CUBLAS_STATUS_SUCCESS;
}
// C = alpha A B + beta C, A symmetric m x m (side left), only the `uplo` triangle of A is read
cublasStatus_t cublasDsymm_v2(cublasHandle_t, cublasSideMode_t side, cublasFillMode_t uplo, int m, int n,
                              const double* alpha, const double* A, int lda, const double* B, int ldb,
                              const double* beta, double* C, int ldc) {
  if (side != CUBLAS_SIDE_LEFT || lda < m || ldb < m || ldc < m) return CUBLAS_STATUS_INVALID_VALUE;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) {
      double s = 0.0;
      for (int p = 0; p < m; ++p) {
        const bool stored = (uplo == CUBLAS_FILL_MODE_LOWER) ? (i >= p) : (i <= p);
        s += (stored ? A[i + size_t(p) * lda] : A[p + size_t(i) * lda]) * B[p + size_t(j) * ldb];
      }
      double& c = C[i + size_t(j) * ldc];
      c = (*beta == 0.0 ? 0.0 : *beta * c) + *alpha * s;
    }
  return CUBLAS_STATUS_SUCCESS;
}
// C = alpha (A B^T + B A^T) + beta C on the `uplo` triangle (trans = N: A, B are n x k)
cublasStatus_t cublasDsyr2k_v2(cublasHandle_t, cublasFillMode_t uplo, cublasOperation_t trans, int n, int k,
                               const double* alpha, const double* A, int lda, const double* B, int ldb,
                               const double* beta, double* C, int ldc) {
  if (trans != CUBLAS_OP_N || lda < n || ldb < n || ldc < n) return CUBLAS_STATUS_INVALID_VALUE;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      if ((uplo == CUBLAS_FILL_MODE_LOWER) ? (i < j) : (i > j)) continue;
      double s = 0.0;
      for (int p = 0; p < k; ++p) s += A[i + size_t(p) * lda] * B[j + size_t(p) * ldb] + B[i + size_t(p) * ldb] * A[j + size_t(p) * lda];
      double& c = C[i + size_t(j) * ldc];
      c = *beta * c + *alpha * s;
    }
  return CUBLAS_STATUS_SUCCESS;
}

// ------------------------------------------------------------------ entry points for the test
static std::vector<char> g_ws;
static size_t g_ws_off = 0;

const char* emu_last_error(void) { return tq::g_err_buf; }

// the batches apply_q2 issues for an n x n problem, as (sb0, k0, count, hg) quadruples; returns their number
// (or a negative status).  No arithmetic: this is the schedule alone, cheap at any n.
int emu_q2_schedule(int64_t n, int* out, int cap) {
  int cnt = 0;
  const int st = tq::q2_for_each_batch(n, [&](int sb0, int k0, int count, int hg) -> int {
    if (cnt < cap) {
      out[4 * cnt + 0] = sb0, out[4 * cnt + 1] = k0, out[4 * cnt + 2] = count, out[4 * cnt + 3] = hg;
    }
    ++cnt;
    return TQ_OK;
  });
  return st == TQ_OK ? cnt : st;
}

// A: n x n column-major (both triangles valid on entry; only the lower one is used).  On exit A holds the band and
// the stage-1 reflectors, (d, e) the tridiagonal matrix.  `sms` = CTAs the bulge chase may use.
int emu_two_stage_reduce(double* A, int64_t n, double* d, double* e, int sms) {
  tq::g_emu_sms = sms;
  if (!tq::two_stage_usable(n)) return -100;
  g_ws.assign(tq::two_stage_ws_bytes(n) + size_t(n) * n * 8 * 3 + (1 << 20), 0);
  for (size_t i = 0; i + 8 <= g_ws.size(); i += 8) *reinterpret_cast<double*>(&g_ws[i]) = NAN;   // like fresh device memory
  tq::Workspace ws(g_ws.data(), g_ws.size());
  const int st = tq::two_stage_reduce(nullptr, nullptr, A, n, d, e, ws);
  g_ws_off = ws.off;
  return st;
}

// the C ABI's debug entry point itself (H row-major), on host memory
int emu_two_stage_debug(const double* H, int64_t n, double* band_out, double* d, double* e, int sms) {
  tq::g_emu_sms = sms;
  std::vector<char> ws(tq::two_stage_ws_bytes(n) + size_t(n) * n * 8 * 4 + (1 << 20), 0);
  return tq_two_stage_debug(H, n, n, band_out, d, e, ws.data(), ws.size(), nullptr);
}

// Z (n x ncols column-major, ld n) <- Q1 Q2 Z with the reflectors of the last emu_two_stage_reduce
int emu_two_stage_back(const double* A, int64_t n, double* Z, int64_t ncols) {
  tq::Workspace ws(g_ws.data(), g_ws.size());
  ws.off = g_ws_off;
  return tq::two_stage_back(nullptr, nullptr, A, n, Z, ncols, ws);
}
}
