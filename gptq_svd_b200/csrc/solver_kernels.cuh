// fp64 building blocks shared by the spectral solver stages (column-major, LAPACK
// style indexing: element (r, c) of a matrix M with leading dimension ld is M[r + c*ld]).
//
// The BLAS-2 parts of the factorizations (tridiagonal reduction, pivoted-QR panels) are
// HBM-bound: their dominant operation is "dot every trailing column with one vector",
// implemented by dots3_kernel (one warp per column, coalesced down the column,
// deterministic shuffle-tree reduction).  BLAS-3 parts go to cuBLAS DGEMM.
#pragma once
#include "blas.cuh"
#include "common.cuh"

namespace tq {

struct DotSeg {
  const double* M;  // first element of the first column
  int64_t ld;
  int64_t ncols;
  double* out;      // out[j] = dot(M[:, j], x)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double block_sum(double v, double* sh /*32 doubles*/) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  if (threadIdx.x == 0) sh[0] = v;
  __syncthreads();
  v = sh[0];
  return v;
}

// Whole-CTA dot product of two length-len vectors (one trailing-matrix column against the
// reflector): every thread keeps 8 independent 8-byte loads in flight (64 KB per SM at
// 1024 resident threads, enough to cover HBM latency), then a fixed-order reduction
// (warp shuffles, then warp 0 over the per-warp partials) makes the result deterministic.
// Thread 0 of the CTA stores the sum to *out.  `shbuf` is 32 doubles; callers alternate
// between two buffers on consecutive calls so that one __syncthreads per call suffices.
__device__ __forceinline__ void cta_dot_store(const double* __restrict__ col, const double* __restrict__ v,
                                              int64_t len, double* shbuf, double* out) {
  const int64_t step = blockDim.x;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
  int64_t r = threadIdx.x;
  for (; r + 7 * step < len; r += 8 * step) {
    const double m0 = col[r], m1 = col[r + step], m2 = col[r + 2 * step], m3 = col[r + 3 * step];
    const double m4 = col[r + 4 * step], m5 = col[r + 5 * step], m6 = col[r + 6 * step], m7 = col[r + 7 * step];
    a0 = fma(m0, v[r], a0);
    a1 = fma(m1, v[r + step], a1);
    a2 = fma(m2, v[r + 2 * step], a2);
    a3 = fma(m3, v[r + 3 * step], a3);
    a4 = fma(m4, v[r + 4 * step], a4);
    a5 = fma(m5, v[r + 5 * step], a5);
    a6 = fma(m6, v[r + 6 * step], a6);
    a7 = fma(m7, v[r + 7 * step], a7);
  }
  for (; r + step < len; r += 2 * step) {
    const double m0 = col[r], m1 = col[r + step];
    a0 = fma(m0, v[r], a0);
    a1 = fma(m1, v[r + step], a1);
  }
  if (r < len) a2 = fma(col[r], v[r], a2);
  double sres = warp_sum(((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7)));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) shbuf[w] = sres;
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    double t = lane < nw ? shbuf[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) *out = t;
  }
}

// Warp-level partial dot over rows [r0, r1): 8 independent 8-byte loads per lane in flight,
// deterministic shuffle-tree reduction; every lane returns the sum.
__device__ __forceinline__ double warp_dot_range(const double* __restrict__ col, const double* __restrict__ v,
                                                 int64_t r0, int64_t r1, int lane) {
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
  int64_t r = r0 + lane;
  for (; r + 224 < r1; r += 256) {
    const double m0 = col[r], m1 = col[r + 32], m2 = col[r + 64], m3 = col[r + 96];
    const double m4 = col[r + 128], m5 = col[r + 160], m6 = col[r + 192], m7 = col[r + 224];
    a0 = fma(m0, v[r], a0);
    a1 = fma(m1, v[r + 32], a1);
    a2 = fma(m2, v[r + 64], a2);
    a3 = fma(m3, v[r + 96], a3);
    a4 = fma(m4, v[r + 128], a4);
    a5 = fma(m5, v[r + 160], a5);
    a6 = fma(m6, v[r + 192], a6);
    a7 = fma(m7, v[r + 224], a7);
  }
  for (; r + 32 < r1; r += 64) {
    const double m0 = col[r], m1 = col[r + 32];
    a0 = fma(m0, v[r], a0);
    a1 = fma(m1, v[r + 32], a1);
  }
  if (r < r1) a2 = fma(col[r], v[r], a2);
  return warp_sum(((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7)));
}

// Grid-wide barrier for the persistent (cooperatively launched, hence co-resident) panel
// kernels: one monotonically increasing counter in global memory, zeroed before the launch.
// Thread 0 of every CTA releases its writes, arrives and spins with acquire loads; the
// gpu-scope fence after the spin invalidates the SM's L1 so the CTA reads fresh data.
// Measured several times cheaper than cooperative_groups::grid_group::sync() on B200.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

// Partial dot of one column against v by ONE WARP of a CTA that streams the column
// cooperatively: the CTA covers rows contiguously (thread t takes rows t, t + blockDim, ...),
// 8 loads per thread in flight; the caller stores the per-warp partial and the consumer adds
// the blockDim/32 partials in fixed order - no block-level barrier in the streaming loop.
__device__ __forceinline__ double cta_strided_warp_dot(const double* __restrict__ col, const double* __restrict__ v,
                                                       int64_t len) {
  const int64_t step = blockDim.x;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
  int64_t r = threadIdx.x;
  for (; r + 7 * step < len; r += 8 * step) {
    const double m0 = col[r], m1 = col[r + step], m2 = col[r + 2 * step], m3 = col[r + 3 * step];
    const double m4 = col[r + 4 * step], m5 = col[r + 5 * step], m6 = col[r + 6 * step], m7 = col[r + 7 * step];
    a0 = fma(m0, v[r], a0);
    a1 = fma(m1, v[r + step], a1);
    a2 = fma(m2, v[r + 2 * step], a2);
    a3 = fma(m3, v[r + 3 * step], a3);
    a4 = fma(m4, v[r + 4 * step], a4);
    a5 = fma(m5, v[r + 5 * step], a5);
    a6 = fma(m6, v[r + 6 * step], a6);
    a7 = fma(m7, v[r + 7 * step], a7);
  }
  if (r + 3 * step < len) {
    const double m0 = col[r], m1 = col[r + step], m2 = col[r + 2 * step], m3 = col[r + 3 * step];
    a0 = fma(m0, v[r], a0);
    a1 = fma(m1, v[r + step], a1);
    a2 = fma(m2, v[r + 2 * step], a2);
    a3 = fma(m3, v[r + 3 * step], a3);
    r += 4 * step;
  }
  if (r + step < len) {
    const double m0 = col[r], m1 = col[r + step];
    a4 = fma(m0, v[r], a4);
    a5 = fma(m1, v[r + step], a5);
    r += 2 * step;
  }
  if (r < len) a6 = fma(col[r], v[r], a6);
  if (r + step < len) a7 = fma(col[r + step], v[r + step], a7);
  return warp_sum(((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7)));
}

// Work split of the streaming phase: `cols` columns of `len` rows are cut into row chunks so
// that there are about 8 equal units per warp (balanced tail) but no unit is shorter than 512
// rows.  Partial sums of the (at most kMaxChunks) chunks are added in fixed order by the
// consumer, so the result does not depend on the scheduling.
constexpr int kMaxChunks = 16;
__device__ __forceinline__ void chunk_plan(int64_t len, int64_t cols, int64_t nwarps, int& H, int64_t& L) {
  int64_t h = cols > 0 ? (8 * nwarps) / cols : 1;
  const int64_t hmax = (len + 511) / 512;
  if (h > hmax) h = hmax;
  if (h > kMaxChunks) h = kMaxChunks;
  if (h < 1) h = 1;
  L = ((len + h - 1) / h + 31) / 32 * 32;
  if (L < 32) L = 32;
  H = int((len + L - 1) / L);
  if (H < 1) H = 1;
}

// out[j] = dot(M[:, j], x) over `rows` rows for up to three column sets sharing x.
// `skip` (optional device flag): when *skip != 0 the kernel does nothing.
static __global__ void __launch_bounds__(256)
dots3_kernel(DotSeg s0, DotSeg s1, DotSeg s2, const double* __restrict__ x, int64_t rows,
             const int* __restrict__ skip) {
  if (skip && *skip) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int64_t total = s0.ncols + s1.ncols + s2.ncols;
  for (int64_t j = warp; j < total; j += nwarps) {
    const double* col;
    double* out;
    if (j < s0.ncols) {
      col = s0.M + j * s0.ld;
      out = s0.out + j;
    } else if (j < s0.ncols + s1.ncols) {
      int64_t jj = j - s0.ncols;
      col = s1.M + jj * s1.ld;
      out = s1.out + jj;
    } else {
      int64_t jj = j - s0.ncols - s1.ncols;
      col = s2.M + jj * s2.ld;
      out = s2.out + jj;
    }
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int64_t r = lane;
    for (; r + 96 < rows; r += 128) {
      double m0 = col[r], m1 = col[r + 32], m2 = col[r + 64], m3 = col[r + 96];
      a0 = fma(m0, x[r], a0);
      a1 = fma(m1, x[r + 32], a1);
      a2 = fma(m2, x[r + 64], a2);
      a3 = fma(m3, x[r + 96], a3);
    }
    for (; r < rows; r += 32) a0 = fma(col[r], x[r], a0);
    double s = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) *out = s;
  }
}

// Householder reflector (LAPACK DLARFG) for the vector [alpha; x] of length len stored
// contiguously at v: on exit v[0] = 1, v[1:] = x / (alpha - beta), *beta_out = beta,
// *tau_out = tau.  Single CTA.  `skip` as above.
static __global__ void __launch_bounds__(1024)
larfg_kernel(double* __restrict__ v, int64_t len, double* __restrict__ tau_out, double* __restrict__ beta_out,
             const int* __restrict__ skip) {
  if (skip && *skip) return;
  __shared__ double sh[32];
  double ss = 0.0;
  for (int64_t r = 1 + threadIdx.x; r < len; r += blockDim.x) ss = fma(v[r], v[r], ss);
  ss = block_sum(ss, sh);
  const double alpha = v[0];
  __syncthreads();
  double tau, beta, scal;
  if (len <= 1 || ss == 0.0) {
    tau = 0.0;
    beta = alpha;
    scal = 0.0;
  } else {
    const double xnorm = sqrt(ss);
    beta = -copysign(hypot(alpha, xnorm), alpha);
    tau = (beta - alpha) / beta;
    scal = 1.0 / (alpha - beta);
  }
  if (tau != 0.0)
    for (int64_t r = 1 + threadIdx.x; r < len; r += blockDim.x) v[r] *= scal;
  if (threadIdx.x == 0) {
    v[0] = 1.0;
    *tau_out = tau;
    *beta_out = beta;
  }
}

// T factor of a block reflector H = I - V T V^T (forward, columnwise; LAPACK DLARFT)
// from G = V^T V (jb x jb, ldg) and tau: T[t,t] = tau_t, T[0:t, t] = -tau_t T[0:t,0:t] G[0:t, t].
// Single CTA, jb <= 128.  T is jb x jb upper triangular, ldt.
static __global__ void __launch_bounds__(128)
larft_kernel(const double* __restrict__ G, int ldg, const double* __restrict__ tau, int jb, double* __restrict__ T,
             int ldt) {
  extern __shared__ double Ts[];  // jb x jb
  const int tid = threadIdx.x;
  for (int idx = tid; idx < jb * jb; idx += blockDim.x) Ts[idx] = 0.0;
  __syncthreads();
  for (int t = 0; t < jb; ++t) {
    const double tt = tau[t];
    // column t: rows r < t
    if (tid < t) {
      double s = 0.0;
      for (int q = tid; q < t; ++q) s = fma(Ts[tid + q * jb], G[q + t * ldg], s);  // T upper: T[tid, q], q >= tid
      Ts[tid + t * jb] = -tt * s;
    }
    if (tid == t) Ts[t + t * jb] = tt;
    __syncthreads();
  }
  for (int idx = tid; idx < jb * jb; idx += blockDim.x) T[(idx % jb) + (idx / jb) * ldt] = Ts[idx];
}

// Vc (s x jb, ld ldvc) = clean copy of the reflector block stored in A: unit diagonal,
// zeros above it, A's entries below it.  A points at the row of the first unit entry.
static __global__ void copy_reflectors_kernel(const double* __restrict__ A, int64_t lda, int64_t s, int jb,
                                       double* __restrict__ Vc, int64_t ldvc) {
  int t = blockIdx.y;
  for (int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < s; r += int64_t(gridDim.x) * blockDim.x) {
    double v = (r < t) ? 0.0 : (r == t ? 1.0 : A[r + int64_t(t) * lda]);
    Vc[r + int64_t(t) * ldvc] = v;
  }
}

// C (s x nc) <- (I - V op(T) V^T) C with V s x jb (clean), T jb x jb upper triangular
// (full storage, zeros below).  Work: jb x nc (x2).  trans_t: use T^T (H^T, as in QR).
inline int apply_block_reflector(cublasHandle_t h, const double* V, int64_t ldv, int64_t s, int jb,
                                 const double* T, int ldt, bool trans_t, double* C, int64_t ldc, int64_t nc,
                                 double* work1, double* work2) {
  if (s <= 0 || nc <= 0 || jb <= 0) return TQ_OK;
  const double one = 1.0, zero = 0.0, mone = -1.0;
  // work1 = V^T C  (jb x nc)
  TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, jb, int(nc), int(s), &one, V, int(ldv), C, int(ldc),
                              &zero, work1, jb));
  // work2 = op(T) work1
  TQ_CUBLAS_CHECK(cublasDgemm(h, trans_t ? CUBLAS_OP_T : CUBLAS_OP_N, CUBLAS_OP_N, jb, int(nc), jb, &one, T, ldt,
                              work1, jb, &zero, work2, jb));
  // C -= V work2
  TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, int(s), int(nc), jb, &mone, V, int(ldv), work2, jb,
                              &one, C, int(ldc)));
  return TQ_OK;
}

// G = V^T V and T = larft(G, tau)
inline int build_t_factor(cublasHandle_t h, cudaStream_t st, const double* V, int64_t ldv, int64_t s, int jb,
                          const double* tau, double* G, double* T) {
  const double one = 1.0, zero = 0.0;
  TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, jb, jb, int(s), &one, V, int(ldv), V, int(ldv), &zero,
                              G, jb));
  larft_kernel<<<1, 128, size_t(jb) * jb * sizeof(double), st>>>(G, jb, tau, jb, T, jb);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

inline unsigned dots_grid(int64_t ncols) {
  int64_t blocks = ceil_div(ncols, 8);  // 8 warps per 256-thread CTA
  int64_t cap = int64_t(num_sms()) * 8;
  return unsigned(imax(1, imin(blocks, cap)));
}

}  // namespace tq
