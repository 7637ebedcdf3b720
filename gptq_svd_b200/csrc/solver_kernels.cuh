// fp64 building blocks shared by the spectral solver stages (column-major, LAPACK
// style indexing: element (r, c) of a matrix M with leading dimension ld is M[r + c*ld]).
//
// The BLAS-2 parts of the factorizations (tridiagonal reduction, pivoted-QR panels) are
// HBM-bound: their dominant operation is "dot every trailing column with one vector",
// done inside the persistent panel kernels by cta_strided_warp_dot (one CTA per column,
// per-warp partials added in fixed order by the consumer).  BLAS-3 parts go to cuBLAS DGEMM.
#pragma once
#include <functional>

#include "blas.cuh"
#include "common.cuh"

namespace tq {

// see eigh_colmajor (eigh.cu): called with the device pointer of the ascending eigenvalues
using EighColumnChooser = std::function<int(const double* w_dev, int64_t* col0, int64_t* ncols)>;


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
// cooperatively.  The CTA covers the column contiguously; thread t takes the 16-byte pairs
// t, t + blockDim, ...  The column is staged global -> shared with per-thread cp.async
// (LDGSTS) 16-byte copies, kAsyncDepth deep, into slots owned by the issuing thread, so no
// registers are held while the data is in flight (128 KB per SM at 1024 resident threads -
// the register-staged version was latency-bound at ~47 % DRAM utilisation in ncu) and no
// block-level barrier is needed.  v comes from L1.  Every lane returns the warp's sum; the
// caller stores the per-warp partial and the consumer adds the blockDim/32 partials in fixed
// order.  Falls back to 8-byte loads when col and v do not share 16-byte parity (odd lda).
constexpr int kAsyncDepth = 8;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ double cta_strided_warp_dot(const double* __restrict__ col, const double* __restrict__ v,
                                                       int64_t len, double2* slots /*[kAsyncDepth][blockDim]*/) {
  const int64_t step = blockDim.x;
  const int tid = threadIdx.x;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  if ((((uintptr_t)col ^ (uintptr_t)v) & 8) == 0) {
    const int64_t head = (((uintptr_t)col & 8) && len > 0) ? 1 : 0;
    const int64_t npairs = (len - head) >> 1;
    if (tid == 0 && head) a2 = col[0] * v[0];
    if (tid == 1 && ((len - head) & 1)) a3 = col[len - 1] * v[len - 1];
    const double2* c2 = reinterpret_cast<const double2*>(col + head);
    const double2* v2 = reinterpret_cast<const double2*>(v + head);
    const int64_t K = npairs > tid ? (npairs - tid + step - 1) / step : 0;
#pragma unroll
    for (int k = 0; k < kAsyncDepth; ++k) {
      if (k < K) cp_async16(&slots[k * step + tid], &c2[tid + k * step]);
      cp_async_commit();
    }
    for (int64_t k = 0; k < K; ++k) {
      cp_async_wait<kAsyncDepth - 1>();
      const int slot = int(k % kAsyncDepth);
      const double2 m = slots[slot * step + tid];
      const double2 x = v2[tid + k * step];
      if (k & 1) {
        a2 = fma(m.x, x.x, a2);
        a3 = fma(m.y, x.y, a3);
      } else {
        a0 = fma(m.x, x.x, a0);
        a1 = fma(m.y, x.y, a1);
      }
      if (k + kAsyncDepth < K) cp_async16(&slots[slot * step + tid], &c2[tid + (k + kAsyncDepth) * step]);
      cp_async_commit();
    }
    cp_async_wait<0>();
  } else {
    int64_t r = tid;
    for (; r + 3 * step < len; r += 4 * step) {
      const double m0 = col[r], m1 = col[r + step], m2 = col[r + 2 * step], m3 = col[r + 3 * step];
      a0 = fma(m0, v[r], a0);
      a1 = fma(m1, v[r + step], a1);
      a2 = fma(m2, v[r + 2 * step], a2);
      a3 = fma(m3, v[r + 3 * step], a3);
    }
    for (; r < len; r += step) a0 = fma(col[r], v[r], a0);
  }
  return warp_sum((a0 + a1) + (a2 + a3));
}

// T factor of a block reflector H = I - V T V^T (forward, columnwise; LAPACK DLARFT)
// from G = V^T V (jb x jb, ldg) and tau, by RECURSIVE DOUBLING instead of DLARFT's jb serial
// triangular mat-vecs (measured 131 us per call at jb = 128, 18 % of the QR stage):
//   T([V1 V2]) = [ T1   -T1 (V1^T V2) T2 ]
//                [ 0          T2         ]
// Level b merges every pair of adjacent b-wide blocks at once (all pairs in parallel over the
// CTA): X = G12 T2, then T12 = -T1 X, in place over the strict upper triangle of G held in
// shared memory.  log2(jb) levels, two CTA barriers each.  Single CTA, jb <= 128.
// T is jb x jb upper triangular (zeros below the diagonal), ldt.
constexpr int kLarftThreads = 1024;
constexpr int kLarftMaxJb = 128;
constexpr size_t kLarftSmem = size_t(kLarftMaxJb) * kLarftMaxJb * 8 + size_t(kLarftMaxJb) * kLarftMaxJb / 4 * 8;

static __global__ void __launch_bounds__(kLarftThreads)
larft_kernel(const double* __restrict__ G, int ldg, const double* __restrict__ tau, int jb, double* __restrict__ T,
             int ldt, int64_t stride_g = 0, int64_t stride_tau = 0, int64_t stride_t = 0) {
  // batched use (one CTA per block reflector): CTA x works on G + x stride_g, tau + x stride_tau, T + x stride_t
  G += int64_t(blockIdx.x) * stride_g;
  tau += int64_t(blockIdx.x) * stride_tau;
  T += int64_t(blockIdx.x) * stride_t;
  TQ_DYN_SMEM(double, larft_sm);
  double* M = larft_sm;                 // jb x jb, column-major, ld jb
  double* X = larft_sm + jb * jb;       // <= jb * jb / 4 entries per level
  const int tid = threadIdx.x;
  for (int idx = tid; idx < jb * jb; idx += kLarftThreads) {
    const int r = idx % jb, c = idx / jb;
    M[idx] = (r < c) ? G[r + c * ldg] : (r == c ? tau[c] : 0.0);
  }
  __syncthreads();
  for (int b = 1; b < jb; b <<= 1) {
    const int npairs = (jb + 2 * b - 1) / (2 * b);
    const int nel = npairs * b * b;
    // X = G12 T2        (b x b2 per pair; T2 upper triangular)
    for (int e = tid; e < nel; e += kLarftThreads) {
      const int p = e / (b * b), rem = e - p * b * b;
      const int r = rem % b, c = rem / b;
      const int o = p * 2 * b, o2 = o + b;
      const int b2 = min(b, jb - o2);
      if (c >= b2) continue;
      const double* g = M + (o + r) + o2 * jb;          // G12[r, q] = g[q * jb]
      const double* t2 = M + o2 + (o2 + c) * jb;        // T2[q, c]  = t2[q]
      double s0 = 0.0, s1 = 0.0;
      int q = 0;
      for (; q + 1 <= c; q += 2) {
        s0 = fma(g[q * jb], t2[q], s0);
        s1 = fma(g[(q + 1) * jb], t2[q + 1], s1);
      }
      if (q <= c) s0 = fma(g[q * jb], t2[q], s0);
      X[e] = s0 + s1;
    }
    __syncthreads();
    // T12 = -T1 X       (T1 upper triangular)
    for (int e = tid; e < nel; e += kLarftThreads) {
      const int p = e / (b * b), rem = e - p * b * b;
      const int r = rem % b, c = rem / b;
      const int o = p * 2 * b, o2 = o + b;
      const int b2 = min(b, jb - o2);
      if (c >= b2) continue;
      const double* t1 = M + (o + r) + o * jb;          // T1[r, q] = t1[q * jb]
      const double* x = X + p * b * b + c * b;          // X[q, c]  = x[q]
      double s0 = 0.0, s1 = 0.0;
      int q = r;
      for (; q + 1 < b; q += 2) {
        s0 = fma(t1[q * jb], x[q], s0);
        s1 = fma(t1[(q + 1) * jb], x[q + 1], s1);
      }
      if (q < b) s0 = fma(t1[q * jb], x[q], s0);
      M[(o + r) + (o2 + c) * jb] = -(s0 + s1);
    }
    __syncthreads();
  }
  for (int idx = tid; idx < jb * jb; idx += kLarftThreads) T[(idx % jb) + (idx / jb) * ldt] = M[idx];
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
static inline int apply_block_reflector(cublasHandle_t h, const double* V, int64_t ldv, int64_t s, int jb,
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
static inline int build_t_factor(cublasHandle_t h, cudaStream_t st, const double* V, int64_t ldv, int64_t s, int jb,
                          const double* tau, double* G, double* T) {
  const double one = 1.0, zero = 0.0;
  TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, jb, jb, int(s), &one, V, int(ldv), V, int(ldv), &zero,
                              G, jb));
  static thread_local bool big_smem[kMaxDevices] = {};
  if (!big_smem[device_slot()]) {   // jb = 128 needs 160 KB of dynamic shared memory
    TQ_CUDA_CHECK(cudaFuncSetAttribute((const void*)larft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kLarftSmem)));
    big_smem[device_slot()] = true;
  }
  if (jb > kLarftMaxJb) {
    set_error("build_t_factor: jb = %d > %d", jb, kLarftMaxJb);
    return TQ_ERR_INVALID;
  }
  TQ_LAUNCH(larft_kernel, 1, kLarftThreads, size_t(jb) * jb * 10, st, G, jb, tau, jb, T, jb, int64_t(0), int64_t(0), int64_t(0));
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

constexpr int kMaxChunks = 16;  // per-warp partial slots of the panel kernels (512 threads)

inline unsigned dots_grid(int64_t ncols) {
  int64_t blocks = ceil_div(ncols, 8);  // 8 warps per 256-thread CTA
  int64_t cap = int64_t(num_sms()) * 8;
  return unsigned(imax(1, imin(blocks, cap)));
}

}  // namespace tq
