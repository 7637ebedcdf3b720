// Spectral solver glue - replaces process_hessian_alt (reference gptq_utils.py:87-126):
//   eigh (eigh.cu) -> clamp / sqrt / flip and the retained-rank rule (block prefix scan
//   built from warp-level scans) -> S = Lambda^1/2 V_k^T -> pivot order + R_x (pivoted Cholesky of
//   S^T S, pchol.cu; or the Householder column-pivoted QR of S, qr.cu) ->
//   B = Lambda^-1/2 V_k^T[:, perm] -> unpivoted QR -> sign-normalised R_x, R (row-major).
#include <cuda.h>

#include "solver_kernels.cuh"

namespace tq {

int eigh_colmajor(cublasHandle_t h, cudaStream_t st, const double* H, int64_t ldh, int64_t n, double* w,
                  double* Zout, Workspace& ws, const EighColumnChooser& choose);
size_t rfactor_ws_bytes(int64_t n, int64_t k);
int r_from_rx(cublasHandle_t h, cudaStream_t st, const double* Hk, int64_t ldh, int64_t n, int64_t k, const double* Rx,
              int64_t ldrx, const int64_t* perm, double* R, int64_t ldr, Workspace ws);
size_t eigh_ws_bytes(int64_t n);
int copy_symmetric_lower(cudaStream_t st, const double* H, int64_t ldh, int64_t n, double* A);
int qr_r_colmajor(cublasHandle_t h, cudaStream_t st, double* A, int64_t lda, int64_t k, int64_t n, Workspace& ws);
int qrcp_colmajor(cublasHandle_t h, cudaStream_t st, double* A, int64_t lda, int64_t k, int64_t n, int64_t* perm,
                  Workspace& ws);
size_t qr_stage_ws_bytes(int64_t k, int64_t n);
size_t pchol_ws_bytes(int64_t n, int64_t k);
int pchol_pivoted(cublasHandle_t h, cudaStream_t st, double* G, int64_t n, int64_t k, double* Rx, int64_t ldr,
                  int64_t* perm64, double* alt, size_t alt_elems, Workspace& ws);
__global__ void emit_r_kernel(const double* __restrict__ A, int64_t lda, int64_t k, int64_t n,
                              double* __restrict__ R, int64_t ldr);

// ------------------------------------------------------------------ rank rule
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Single CTA (1024 threads).  eig_desc[i] = max(w[n-1-i], 1e-12); S = sqrt(eig_desc);
// energy = S*S (the reference squares the square root, gptq_utils.py:94,98).
//   energy:        k = #{ i : cumsum(energy)_i <= (1 - thr) * sum(energy) };  k += (k < n)
//   mean_trimmed:  k = #{ i : S_i > thr * mean(S[1:33]) }
//   otherwise:     k = n
__global__ void __launch_bounds__(1024)
rank_select_kernel(const double* __restrict__ w, int64_t n, double thr, int method, double* __restrict__ eig_desc,
                   long long* __restrict__ k_out, int64_t nvals, int64_t min_rank) {
  // nvals <= n: only the leading nvals values take part in the rule (process_sketch: the sketch has
  // min(rank, n) singular values, gptq_utils.py:49-64); the result is clamped to [min_rank, nvals]
  __shared__ double wsum[32];
  __shared__ double carry_s, total_s, ref_s;
  __shared__ unsigned long long count_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int64_t i = tid; i < n; i += blockDim.x) eig_desc[i] = fmax(w[n - 1 - i], 1e-12);
  if (tid == 0) {
    carry_s = 0.0;
    count_s = 0ull;
  }
  __syncthreads();
  if (method == TQ_RANK_FULL) {
    if (tid == 0) *k_out = (long long)nvals;
    return;
  }
  if (method == TQ_RANK_MEAN_TRIMMED) {
    if (tid == 0) {
      const int64_t ref_k = nvals < 33 ? nvals : 33;
      double s = 0.0;
      for (int64_t i = 1; i < ref_k; ++i) s += sqrt(eig_desc[i]);
      ref_s = nvals > 1 ? s / double(ref_k - 1) : sqrt(eig_desc[0]);
    }
    __syncthreads();
    unsigned long long c = 0;
    for (int64_t i = tid; i < nvals; i += blockDim.x) c += (sqrt(eig_desc[i]) > thr * ref_s) ? 1ull : 0ull;
    atomicAdd(&count_s, c);
    __syncthreads();
    if (tid == 0) {
      long long k = (long long)count_s;
      if (k > nvals) k = nvals;
      if (k < min_rank) k = min_rank;
      *k_out = k;
    }
    return;
  }
  // energy: pass 1 total (fixed order: chunked block scan), pass 2 count
  for (int pass = 0; pass < 2; ++pass) {
    if (tid == 0) carry_s = 0.0;
    __syncthreads();
    const double target = pass ? (1.0 - thr) * total_s : 0.0;
    unsigned long long c = 0;
    for (int64_t base = 0; base < nvals; base += blockDim.x) {
      const int64_t i = base + tid;
      double en = 0.0;
      if (i < nvals) {
        const double s = sqrt(eig_desc[i]);
        en = s * s;
      }
      double v = warp_incl_scan(en, lane);
      if (lane == 31) wsum[wid] = v;
      __syncthreads();
      if (wid == 0) {
        double t = wsum[lane];
        t = warp_incl_scan(t, lane);
        wsum[lane] = t;
      }
      __syncthreads();
      const double prefix = carry_s + (wid ? wsum[wid - 1] : 0.0) + v;
      if (pass && i < nvals && prefix <= target) ++c;
      __syncthreads();
      if (tid == blockDim.x - 1) carry_s = prefix;
      __syncthreads();
    }
    if (pass == 0) {
      if (tid == 0) total_s = carry_s;
    } else {
      atomicAdd(&count_s, c);
    }
    __syncthreads();
  }
  if (tid == 0) {
    long long k = (long long)count_s;
    if (k < nvals) k += 1;
    if (k < min_rank) k = min_rank;
    *k_out = k;
  }
}

// flags[0] = 1 when a retained eigenvalue was raised by the 1e-12 clamp (then H_k must be built
// from the clamped spectrum, not as H minus the discarded part)
__global__ void clamp_flag_kernel(const double* __restrict__ w, int64_t n, const long long* __restrict__ k,
                                  long long* __restrict__ flags) {
  const long long kk = *k;
  flags[0] = (kk > 0 && kk <= n && w[n - kk] < 1e-12) ? 1 : 0;
}

// Y[:, i] = Z[:, i] * w[i]  (n rows, t columns, column-major ld n)
__global__ void scale_cols_kernel(const double* __restrict__ Z, const double* __restrict__ w, int64_t n, int64_t t,
                                  double* __restrict__ Y) {
  const int64_t i = blockIdx.y;
  const double wi = w[i];
  for (int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < n; r += int64_t(gridDim.x) * blockDim.x)
    Y[r + i * n] = Z[r + i * n] * wi;
}

// ------------------------------------------------------------------ S and B builders
// V is row-major: row r = eigenvector of w[r] (ascending).  Descending index i <-> row n-1-i.
// S (col-major k x n): S[i + j k] = sqrt(e_i) * V[n-1-i][j]              (gptq_utils.py:112)
__global__ void build_s_kernel(const double* __restrict__ V, int64_t n, int64_t k, const double* __restrict__ eig,
                               double* __restrict__ S) {
  __shared__ double t[32][33];
  const int64_t i0 = int64_t(blockIdx.y) * 32, j0 = int64_t(blockIdx.x) * 32;
  for (int a = threadIdx.y; a < 32; a += blockDim.y) {
    int64_t i = i0 + a, j = j0 + threadIdx.x;
    t[a][threadIdx.x] = (i < k && j < n) ? sqrt(eig[i]) * V[(n - 1 - i) * n + j] : 0.0;
  }
  __syncthreads();
  for (int a = threadIdx.y; a < 32; a += blockDim.y) {
    int64_t j = j0 + a, i = i0 + threadIdx.x;
    if (i < k && j < n) S[i + j * k] = t[threadIdx.x][a];
  }
}

// B (col-major k x n): B[i + j k] = (1 / sqrt(e_i)) * V[n-1-i][perm[j]]   (gptq_utils.py:111,118-119)
__global__ void build_b_kernel(const double* __restrict__ V, int64_t n, int64_t k, const double* __restrict__ eig,
                               const int64_t* __restrict__ perm, double* __restrict__ B) {
  __shared__ double t[32][33];
  const int64_t i0 = int64_t(blockIdx.y) * 32, j0 = int64_t(blockIdx.x) * 32;
  const int64_t jj = j0 + threadIdx.x;
  const int64_t pj = jj < n ? perm[jj] : 0;
  for (int a = threadIdx.y; a < 32; a += blockDim.y) {
    int64_t i = i0 + a;
    t[a][threadIdx.x] = (i < k && jj < n) ? (1.0 / sqrt(eig[i])) * V[(n - 1 - i) * n + pj] : 0.0;
  }
  __syncthreads();
  for (int a = threadIdx.y; a < 32; a += blockDim.y) {
    int64_t j = j0 + a, i = i0 + threadIdx.x;
    if (i < k && j < n) B[i + j * k] = t[threadIdx.x][a];
  }
}

static size_t solver_ws_bytes(int64_t n) {
  // V (n^2) + eigh scratch, later overlaid by S/B (n^2) + QR scratch
  size_t a = ws_bytes_for(size_t(n) * n, 8) + ws_bytes_for(n, 8) * 2 + 1024;
  size_t e = eigh_ws_bytes(n);
  size_t q = ws_bytes_for(size_t(n) * n, 8) * 2 + qr_stage_ws_bytes(n, n) + pchol_ws_bytes(n, n);
  size_t r = rfactor_ws_bytes(n, n);
  for (int64_t t = n / 2; t > 0; t /= 2) r = r > rfactor_ws_bytes(n, n - t) ? r : rfactor_ws_bytes(n, n - t);
  if (r > q) q = r;
  return a + (e > q ? e : q) + (size_t(1) << 20);
}

}  // namespace tq

using namespace tq;

extern "C" int tq_solver_workspace(int64_t n, size_t* bytes) {
  TQ_REQUIRE(bytes && n > 0, "tq_solver_workspace: bad arguments");
  *bytes = solver_ws_bytes(n);
  return TQ_OK;
}

extern "C" int tq_rank_select(const double* w_asc, int64_t n, double threshold, int method, double* eig_desc,
                              int64_t* k_host, void* ws, size_t ws_bytes, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(w_asc && eig_desc && k_host && n > 0, "tq_rank_select: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace wsp(ws, ws_bytes);
  long long* kd = wsp.take<long long>(1);
  if (wsp.overflow) {
    set_error("tq_rank_select: workspace too small");
    return TQ_ERR_WORKSPACE;
  }
  rank_select_kernel<<<1, 1024, 0, st>>>(w_asc, n, threshold, method, eig_desc, kd, n, 0);
  TQ_LAUNCH_CHECK();
  long long kh = 0;
  TQ_CUDA_CHECK(cudaMemcpyAsync(&kh, kd, sizeof(long long), cudaMemcpyDeviceToHost, st));
  TQ_CUDA_CHECK(cudaStreamSynchronize(st));
  *k_host = kh;
  return TQ_OK;
}

// Which eigenvectors a solve needs (decided as soon as the eigenvalues exist, before any back-transformation):
//   kSubsetDropped  t = n - k < k and no retained eigenvalue was clamped: only the t DROPPED vectors -
//                   H_k = H - V_t diag(w_t) V_t^T for the pivoted Cholesky, and R follows from R_x (rfactor.cu);
//   kSubsetAll      t < k with a clamped eigenvalue (H_k must be built from the clamped spectrum): all n vectors,
//                   then B = Lambda^-1/2 V_k^T[:, perm] and its QR;
//   kSubsetKept     t >= k: only the k RETAINED vectors (H_k = S^T S, B and its QR).
enum { kSubsetDropped = 0, kSubsetAll = 1, kSubsetKept = 2 };

static int spectral_solve_impl(const double* H, int64_t ldh, int64_t n, double threshold, int method, int64_t nvals,
                               int64_t min_rank, double* R, double* Rx, int64_t* perm, double* eigvals,
                               int64_t* k_host, void* ws, size_t ws_bytes, void* stream, bool allow_dropped = true) {
  TQ_TRY(check_device());
  const int method_in = method;
  const bool force_householder = (method & TQ_SOLVE_HOUSEHOLDER_QRCP) != 0;
  method &= ~TQ_SOLVE_HOUSEHOLDER_QRCP;
  TQ_REQUIRE(H && R && Rx && perm && eigvals && k_host, "tq_spectral_solve: null pointer");
  TQ_REQUIRE(n > 0 && ldh >= n && n < (1 << 30), "tq_spectral_solve: bad shape n=%lld", (long long)n);
  cudaStream_t st = (cudaStream_t)stream;
  Workspace wsp(ws, ws_bytes);
  double* V = wsp.take<double>(size_t(n) * n);
  double* w = wsp.take<double>(n);
  long long* kd = wsp.take<long long>(2);
  if (wsp.overflow) {
    set_error("tq_spectral_solve: workspace too small (%zu bytes given, %zu needed)", ws_bytes, solver_ws_bytes(n));
    return TQ_ERR_WORKSPACE;
  }
  cublasHandle_t h;
  TQ_TRY(get_cublas(&h, st));
  long long kh2[2] = {0, 0};
  int subset = kSubsetAll;
  // the rank rule runs between the divide & conquer and the back-transformation (it needs eigenvalues only)
  auto choose = [&](const double* w_dev, int64_t* col0, int64_t* ncols) -> int {
    StageTimer tm(st, "rank");
    rank_select_kernel<<<1, 1024, 0, st>>>(w_dev, n, threshold, method, eigvals, kd, nvals, min_rank);
    TQ_LAUNCH_CHECK();
    clamp_flag_kernel<<<1, 1, 0, st>>>(w_dev, n, kd, kd + 1);
    TQ_LAUNCH_CHECK();
    TQ_CUDA_CHECK(cudaMemcpyAsync(kh2, kd, 2 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    TQ_CUDA_CHECK(cudaStreamSynchronize(st));
    const int64_t kk = kh2[0], tt = n - kk;
    if (kk <= 0 || kk > n) {
      *col0 = 0;
      *ncols = 0;
      return TQ_OK;
    }
    if (allow_dropped && !force_householder && kh2[1] == 0 && tt < kk) {
      subset = kSubsetDropped;
      *col0 = 0;
      *ncols = tt;
    } else if (tt < kk) {
      subset = kSubsetAll;
      *col0 = 0;
      *ncols = n;
    } else {
      subset = kSubsetKept;
      *col0 = n - kk;
      *ncols = kk;
    }
    return TQ_OK;
  };
  {
    Workspace sub = wsp;
    TQ_TRY(eigh_colmajor(h, st, H, ldh, n, w, V, sub, choose));
  }
  const long long kh = kh2[0];
  const bool clamped = kh2[1] != 0;
  *k_host = kh;
  const int64_t k = kh;
  if (k <= 0) return TQ_OK;   // nothing retained (mean_trimmed can return 0): R, Rx are empty
  TQ_REQUIRE(k <= n, "tq_spectral_solve: rank rule returned k=%lld > n", (long long)k);

  Workspace sub = wsp;
  double* SB = sub.take<double>(size_t(k) * n);
  if (sub.overflow) {
    set_error("tq_spectral_solve: workspace too small for S");
    return TQ_ERR_WORKSPACE;
  }
  dim3 tg((unsigned)ceil_div(n, 32), (unsigned)ceil_div(k, 32));
  // R_x and perm: diagonally pivoted Cholesky of H_k = S^T S (same pivots and factor as the
  // column-pivoted QR of S, see pchol.cu); Householder DLAQPS on S when asked for or when a
  // pivot is not positive.  H_k is formed with the cheaper of two DGEMMs:
  //   H_k = S^T S                         (2 n^2 k flop), or
  //   H_k = H - V_t diag(lambda_t) V_t^T  (2 n^2 (n-k) flop) when fewer pairs are discarded than
  //   kept and no retained eigenvalue was touched by the 1e-12 clamp.
  bool householder = force_householder;
  bool have_s = false;
  if (!householder) {
    Workspace s2 = sub;
    double* Gm = s2.take<double>(size_t(n) * n);
    if (s2.overflow) {
      set_error("tq_spectral_solve: workspace too small for the Gram matrix");
      return TQ_ERR_WORKSPACE;
    }
    StageTimer tm(st, "gram+pchol");
    const double one = 1.0, zero = 0.0, mone = -1.0;
    const int64_t t = n - k;
    if (subset != kSubsetKept && !clamped && t < k) {
      TQ_TRY(copy_symmetric_lower(st, H, ldh, n, Gm));
      if (t > 0) {
        dim3 sg((unsigned)imin(ceil_div(n, 256), 64), (unsigned)t);
        scale_cols_kernel<<<sg, 256, 0, st>>>(V, w, n, t, SB);          // SB is free until S / B are built
        TQ_LAUNCH_CHECK();
        TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, int(n), int(n), int(t), &mone, SB, int(n), V,
                                    int(n), &one, Gm, int(n)));
      }
    } else {
      build_s_kernel<<<tg, dim3(32, 8), 0, st>>>(V, n, k, eigvals, SB);
      TQ_LAUNCH_CHECK();
      have_s = true;
      TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, int(n), int(n), int(k), &one, SB, int(k), SB, int(k),
                                  &zero, Gm, int(n)));
    }
    // the R-from-R_x stage needs (P^T H_k P)[:k, :k] again: keep a copy in the R buffer, which is written last
    if (subset == kSubsetDropped)
      TQ_CUDA_CHECK(cudaMemcpyAsync(R, Gm, sizeof(double) * size_t(n) * n, cudaMemcpyDeviceToDevice, st));
    // SB (k x n) is free until S / B are built: the compacted Gram matrix ping-pongs into it
    const int rc = pchol_pivoted(h, st, Gm, n, k, Rx, n, perm, SB, size_t(k) * n, s2);
    have_s = false;
    if (rc == TQ_ERR_NOCONV) householder = true;
    else if (rc != TQ_OK) return rc;
  }
  if (subset == kSubsetDropped) {
    int rc = TQ_ERR_NOCONV;
    if (!householder) {
      StageTimer tm(st, "r_from_rx");
      rc = r_from_rx(h, st, R, n, n, k, Rx, n, perm, R, n, sub);
    }
    if (rc == TQ_OK) return TQ_OK;
    if (rc != TQ_ERR_NOCONV) return rc;
    // numerical rank below k (non-positive pivot): the Householder route needs the retained eigenvectors,
    // which this pass did not back-transform - solve again the long way (rare)
    return spectral_solve_impl(H, ldh, n, threshold, method_in, nvals, min_rank, R, Rx, perm, eigvals, k_host, ws,
                               ws_bytes, stream, false);
  }
  if (householder) {
    if (!have_s) {
      build_s_kernel<<<tg, dim3(32, 8), 0, st>>>(V, n, k, eigvals, SB);
      TQ_LAUNCH_CHECK();
    }
    {
      Workspace s2 = sub;
      StageTimer tm(st, "qrcp");
      TQ_TRY(qrcp_colmajor(h, st, SB, k, k, n, perm, s2));
    }
    emit_r_kernel<<<tg, dim3(32, 8), 0, st>>>(SB, k, k, n, Rx, n);
    TQ_LAUNCH_CHECK();
  }
  {
    StageTimer tm(st, "build_b");
    build_b_kernel<<<tg, dim3(32, 8), 0, st>>>(V, n, k, eigvals, perm, SB);
    TQ_LAUNCH_CHECK();
  }
  {
    Workspace s2 = sub;
    StageTimer tm(st, "qr_r");
    TQ_TRY(qr_r_colmajor(h, st, SB, k, k, n, s2));
  }
  {
    StageTimer tm(st, "emit_r");
    emit_r_kernel<<<tg, dim3(32, 8), 0, st>>>(SB, k, k, n, R, n);
    TQ_LAUNCH_CHECK();
  }
  return TQ_OK;
}

extern "C" int tq_spectral_solve(const double* H, int64_t ldh, int64_t n, double threshold, int method, double* R,
                                 double* Rx, int64_t* perm, double* eigvals, int64_t* k_host, void* ws,
                                 size_t ws_bytes, void* stream) {
  const int rc = spectral_solve_impl(H, ldh, n, threshold, method, n, 0, R, Rx, perm, eigvals, k_host, ws, ws_bytes,
                                     stream);
  if (n >= 8192) trace_flush((cudaStream_t)stream);
  return rc;
}

// ------------------------------------------------------------------ sketch path (gptq_utils.py:33-84, 171-211)
namespace tq {
__global__ void f32_to_f64_kernel(const float* __restrict__ src, int64_t lds, int64_t rows, int64_t n,
                                  double* __restrict__ dst) {
  const int64_t total = rows * n;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = idx / n, c = idx - r * n;
    dst[idx] = double(src[r * lds + c]);
  }
}
__global__ void any_to_f32_kernel(const void* __restrict__ src, int dtype, int64_t lds, int64_t rows, int64_t n,
                                  float* __restrict__ dst) {
  const int64_t total = rows * n;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = idx / n, c = idx - r * n;
    float v;
    if (dtype == TQ_F16) v = __half2float(reinterpret_cast<const __half*>(src)[r * lds + c]);
    else if (dtype == TQ_BF16) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[r * lds + c]);
    else if (dtype == TQ_F64) v = float(reinterpret_cast<const double*>(src)[r * lds + c]);
    else v = reinterpret_cast<const float*>(src)[r * lds + c];
    dst[idx] = v;
  }
}
}  // namespace tq

namespace tq {
struct TrailingTc {
  CUtensorMap ehi, elo, uhi, ulo;
  uint32_t idesc;
};
int trailing_tc_prepare_ex(TrailingTc* t, const float* A_hi, const float* A_lo, int64_t m, int64_t lda, int64_t kdim,
                           const float* BT_hi, const float* BT_lo, int64_t kpad, int64_t n);
int trailing_tc_launch_ex(const TrailingTc* t, float* C, int64_t ldc, int64_t m, int64_t N, int64_t e_col0,
                          int64_t u_row0, int64_t kcount, int64_t u_col0, float sign, cudaStream_t st);
__global__ void split_transpose_u_kernel(const float* __restrict__ U, int64_t k, int64_t n, int64_t kpad,
                                         float* __restrict__ UT_hi, float* __restrict__ UT_lo);

// hi / lo TF32 planes (rows x ldp, zero padded) of a row-major fp32 matrix
__global__ void split_planes_kernel(const float* __restrict__ A, int64_t lda, int64_t rows, int64_t cols, int64_t ldp,
                                    float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t r = blockIdx.y;
  for (int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; c < ldp; c += int64_t(gridDim.x) * blockDim.x) {
    float h = 0.f, l = 0.f;
    if (c < cols) {
      const float x = A[r * lda + c];
      uint32_t hb, lb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
      h = __uint_as_float(hb);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(__fsub_rn(x, h)));
      l = __uint_as_float(lb);
    }
    hi[r * ldp + c] = h;
    lo[r * ldp + c] = l;
  }
}
}  // namespace tq

extern "C" int tq_sketch_accum_workspace(int64_t rank, int64_t rows, int64_t n, size_t* bytes) {
  TQ_REQUIRE(bytes && rank > 0 && rows > 0 && n > 0, "tq_sketch_accum_workspace: bad arguments");
  const size_t rp = size_t((rows + 3) / 4 * 4);
  *bytes = ws_bytes_for(size_t(rows) * n, 4) + 2 * ws_bytes_for(size_t(n) * rp, 4) + 2 * ws_bytes_for(size_t(rank) * rp, 4) +
           4096;
  return TQ_OK;
}

/* Y (rank x n fp32) += Rb (rank x rows fp32) @ float32(X) (rows x n): Sketcher.hook_fn, gptq_utils.py:185-203 (an fp32
 * matmul with TF32 off in the reference).  Runs on the tensor cores with fp32-grade arithmetic: both operands split
 * into TF32 hi + lo planes, three tcgen05 MMAs per k-step, 128 k per TMEM accumulation, chunks added in fp32
 * registers, ONE rounding when the batch's product is added to Y - the kernel of the GPTQ trailing update
 * (trailing_tc.cu) with sign +1.  X is cast to fp32 and transposed first (the tensor core wants both operands
 * k-contiguous).  ws: tq_sketch_accum_workspace(rank, rows, n) bytes. */
extern "C" int tq_sketch_accum(float* Y, int64_t ldy, const float* Rb, int64_t ldr, const void* X, int x_dtype,
                               int64_t ldx, int64_t rank, int64_t rows, int64_t n, void* ws, size_t ws_bytes,
                               void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(Y && Rb && X && rank > 0 && rows > 0 && n > 0 && ldy >= n && ldr >= rows && ldx >= n,
             "tq_sketch_accum: bad arguments");
  TQ_REQUIRE(rows < (int64_t(1) << 31) && n < (int64_t(1) << 31), "tq_sketch_accum: shape too large");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rp = (rows + 3) / 4 * 4;
  Workspace wsp(ws, ws_bytes);
  float* Xf = wsp.take<float>(size_t(rows) * n);
  float* XT_hi = wsp.take<float>(size_t(n) * rp);
  float* XT_lo = wsp.take<float>(size_t(n) * rp);
  float* R_hi = wsp.take<float>(size_t(rank) * rp);
  float* R_lo = wsp.take<float>(size_t(rank) * rp);
  if (wsp.overflow) {
    set_error("tq_sketch_accum: workspace too small (see tq_sketch_accum_workspace)");
    return TQ_ERR_WORKSPACE;
  }
  any_to_f32_kernel<<<(unsigned)imin(ceil_div(rows * n, 256), 65535), 256, 0, st>>>(X, x_dtype, ldx, rows, n, Xf);
  TQ_LAUNCH_CHECK();
  {
    dim3 tg((unsigned)ceil_div(n, 32), (unsigned)ceil_div(rp, 32));
    split_transpose_u_kernel<<<tg, dim3(32, 8), 0, st>>>(Xf, rows, n, rp, XT_hi, XT_lo);
    TQ_LAUNCH_CHECK();
    dim3 sg((unsigned)imin(ceil_div(rp, 256), 64), (unsigned)rank);
    split_planes_kernel<<<sg, 256, 0, st>>>(Rb, ldr, rank, rows, rp, R_hi, R_lo);
    TQ_LAUNCH_CHECK();
  }
  TrailingTc tc;
  TQ_TRY(trailing_tc_prepare_ex(&tc, R_hi, R_lo, rank, rp, rp, XT_hi, XT_lo, rp, n));
  return trailing_tc_launch_ex(&tc, Y, ldy, rank, n, 0, 0, rows, 0, 1.0f, st);
}

extern "C" int tq_sketch_workspace(int64_t rank, int64_t n, size_t* bytes) {
  TQ_REQUIRE(bytes && n > 0 && rank > 0, "tq_sketch_workspace: bad arguments");
  *bytes = solver_ws_bytes(n) + ws_bytes_for(size_t(n) * n, 8) * 2 + ws_bytes_for(size_t(rank) * n, 8) +
           ws_bytes_for(n, 8) + 4096;
  return TQ_OK;
}

/* process_sketch (gptq_utils.py:33-84) on the scaled sketch Y (rank x n fp32): the reference takes
 * the singular values / right vectors of the sketch (geqrf + svd); they are the square roots of the
 * eigenvalues / the eigenvectors of Y^T Y, so the spectral solver runs on G = Y^T Y (fp64) with the
 * rule restricted to min(rank, n) values and a floor of 1 (:64).  Returns R (k x n) and perm. */
extern "C" int tq_sketch_solve(const float* Y, int64_t ldy, int64_t rank, int64_t n, double threshold, int method,
                               double* R, int64_t* perm, int64_t* k_host, void* ws, size_t ws_bytes, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(Y && R && perm && k_host && rank > 0 && n > 0 && ldy >= n, "tq_sketch_solve: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace wsp(ws, ws_bytes);
  double* G = wsp.take<double>(size_t(n) * n);
  double* Rx = wsp.take<double>(size_t(n) * n);
  double* eig = wsp.take<double>(n);
  double* Yd = wsp.take<double>(size_t(rank) * n);
  if (wsp.overflow) {
    set_error("tq_sketch_solve: workspace too small (see tq_sketch_workspace)");
    return TQ_ERR_WORKSPACE;
  }
  cublasHandle_t h;
  TQ_TRY(get_cublas(&h, st));
  f32_to_f64_kernel<<<(unsigned)imin(ceil_div(rank * n, 256), 65535), 256, 0, st>>>(Y, ldy, rank, n, Yd);
  TQ_LAUNCH_CHECK();
  const double one = 1.0, zero = 0.0;
  // row-major Yd (rank x n) is column-major n x rank: G = Yd^T Yd = (col-major Yd)(col-major Yd)^T
  TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, int(n), int(n), int(rank), &one, Yd, int(n), Yd, int(n),
                              &zero, G, int(n)));
  const int64_t nvals = rank < n ? rank : n;
  const size_t used = wsp.off;
  return spectral_solve_impl(G, n, n, threshold, method, nvals, 1, R, Rx, perm, eig, k_host,
                             static_cast<char*>(ws) + align_up(used, 256), ws_bytes - align_up(used, 256), stream);
}
