// Householder QR factorizations of the spectral solver (fp64, column-major inside).
//
//  * qrcp_colmajor: column-pivoted QR with LAPACK DGEQP3 / DLAQPS semantics (the
//    reference calls jax.scipy.linalg.qr(pivoting=True) -> MAGMA dgeqp3, gptq_utils.py:114):
//    pivot = first index of the largest downdated partial norm, norms downdated with the
//    DLAQPS cancellation safeguard (a failing column ends the panel and is recomputed
//    exactly), lazy panel update through F, trailing update by DGEMM.
//    The per-column work is BLAS-2 and HBM-bound: the panel kernel streams the whole
//    trailing matrix once per column (F(:,k) = tau A^T v).
//  * qr_r_colmajor: unpivoted blocked Householder QR that never forms Q (the reference
//    calls torch.linalg.qr and discards Q, gptq_utils.py:120): panel by BLAS-2 kernels,
//    trailing update by compact-WY DGEMMs.
#include <cooperative_groups.h>
#include <cstdlib>

#include "solver_kernels.cuh"

namespace cg = cooperative_groups;

namespace tq {

constexpr int kQrNb = 64;    // unpivoted QR panel width
constexpr int kQrcpNb = 32;  // pivoted QR panel width (DLAQPS)

// ------------------------------------------------------------------ layout helpers
// dst (col-major k x n, ld ldd) = src (row-major k x n, ld lds), 32x32 tiles through smem
__global__ void rowmajor_to_colmajor_kernel(const double* __restrict__ src, int64_t lds, int64_t k, int64_t n,
                                            double* __restrict__ dst, int64_t ldd) {
  __shared__ double t[32][33];
  int64_t i0 = int64_t(blockIdx.y) * 32, j0 = int64_t(blockIdx.x) * 32;
  for (int a = threadIdx.y; a < 32; a += blockDim.y) {
    int64_t i = i0 + a, j = j0 + threadIdx.x;
    t[a][threadIdx.x] = (i < k && j < n) ? src[i * lds + j] : 0.0;
  }
  __syncthreads();
  for (int a = threadIdx.y; a < 32; a += blockDim.y) {
    int64_t j = j0 + a, i = i0 + threadIdx.x;
    if (i < k && j < n) dst[i + j * ldd] = t[threadIdx.x][a];
  }
}

// R_out (row-major k x n, ld ldr) = sign(diag) * triu(A) with A col-major (gptq_utils.py:121-124)
__global__ void emit_r_kernel(const double* __restrict__ A, int64_t lda, int64_t k, int64_t n,
                              double* __restrict__ R, int64_t ldr) {
  __shared__ double t[32][33];
  int64_t i0 = int64_t(blockIdx.y) * 32, j0 = int64_t(blockIdx.x) * 32;
  for (int a = threadIdx.y; a < 32; a += blockDim.y) {
    int64_t j = j0 + a, i = i0 + threadIdx.x;
    t[a][threadIdx.x] = (i < k && j < n && i <= j) ? A[i + j * lda] : 0.0;
  }
  __syncthreads();
  for (int a = threadIdx.y; a < 32; a += blockDim.y) {
    int64_t i = i0 + a, j = j0 + threadIdx.x;
    if (i < k && j < n) {
      double d = A[i + i * lda];
      double sg = d > 0.0 ? 1.0 : (d < 0.0 ? -1.0 : 0.0);   // torch.sign
      R[i * ldr + j] = sg * t[threadIdx.x][a];
    }
  }
}


__global__ void set_diag_kernel(double* __restrict__ A, int64_t lda, int64_t j0, int jb,
                                const double* __restrict__ beta) {
  int t = threadIdx.x;
  if (t < jb) A[(j0 + t) + (j0 + t) * lda] = beta[j0 + t];
}

struct QrcpCtl {
  int stop;       // panel is over (a norm failed the safeguard in an earlier step)
  int stop_next;  // set during the step that flags a column
  int kb;         // columns factored in this panel
  int pad;
};

// ----------------------------------------------------------------------- persistent panels
// One cooperative launch factors a whole panel; phases are separated by grid_barrier.
// The streaming phases dot the trailing columns against the RAW reflector column
// u = [alpha; x] and fix the per-warp partial up to v = [1; scl x]:
//   p_v = scl * p_u + col[0] * (1 - scl * alpha)     (the warp owning row 0 adds the last term)
// so no barrier is needed between computing the Householder scalars and using v.
constexpr int kQrPanelThreads = 512;
constexpr int kQrPanelWarps = kQrPanelThreads / 32;
constexpr size_t kQrPanelSmem = size_t(kAsyncDepth) * kQrPanelThreads * sizeof(double2);

__device__ __forceinline__ int64_t imin_d(int64_t a, int64_t b) { return a < b ? a : b; }
__device__ __forceinline__ int64_t imax_d(int64_t a, int64_t b) { return a > b ? a : b; }

__device__ __forceinline__ double grid_total(const double* part, int nb, double* sh) {
  double v = (threadIdx.x < nb) ? part[threadIdx.x] : 0.0;
  for (int q = threadIdx.x + blockDim.x; q < nb; q += blockDim.x) v += part[q];
  return block_sum(v, sh);
}

__device__ __forceinline__ void householder_scalars(double alpha, double sumsq, int64_t len, double& tau,
                                                    double& beta, double& scl) {
  if (len <= 1 || sumsq == 0.0) {
    tau = 0.0;
    beta = alpha;
    scl = 0.0;
  } else {
    const double xnorm = sqrt(sumsq);
    beta = -copysign(hypot(alpha, xnorm), alpha);
    tau = (beta - alpha) / beta;
    scl = 1.0 / (alpha - beta);
  }
}

struct QrPanelArgs {
  double* A;
  int64_t lda, k;
  int64_t j0;
  int jb;
  double* tau;
  double* beta;
  double* wpart;   // kQrPanelWarps x kQrNb per-warp partials of A[c:, c+1:pend]^T v
  double* part;
  double* scal;
  unsigned int* bar;
};

// Unpivoted Householder panel (DGEQR2 on columns [j0, j0+jb), rows [j0, k)), 2 barriers per column:
//   A  finish the previous reflector (scale v, set the diagonal to beta) and apply it to the
//      remaining panel columns (own rows); partial sum of squares of column c    | barrier
//   B  scalars; w = A[c:, c+1:pend]^T v via raw dots (one CTA per column)         | barrier
__global__ void __launch_bounds__(kQrPanelThreads, 2) qr_panel_kernel(QrPanelArgs a) {
  extern __shared__ double2 dot_slots[];   // [kAsyncDepth][blockDim] cp.async staging of the streamed column
  __shared__ double sh[32];
  __shared__ double wd[kQrNb];
  double* const A = a.A;
  const int64_t lda = a.lda, k = a.k, j0 = a.j0;
  const int64_t gt = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthreads = int64_t(gridDim.x) * blockDim.x;
  const unsigned int nb = gridDim.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t pend = j0 + a.jb;
  unsigned int bar_target = 0;
  double tau_p = 0.0, beta_p = 0.0, scl_p = 0.0;      // scalars of the previous column
  for (int i = 0; i <= a.jb; ++i) {
    const int64_t c = j0 + i;
    const bool last = (i == a.jb);                    // epilogue pass: only finish reflector jb-1
    // ---------------- A
    if (i > 0) {
      const int nrem = int(pend - c);
      for (int t = threadIdx.x; t < nrem; t += blockDim.x) {
        double wsum = 0.0;
        for (int w = 0; w < kQrPanelWarps; ++w) wsum += a.wpart[w * kQrNb + t];
        wd[t] = wsum;
      }
      __syncthreads();
    }
    double ss = 0.0;
    for (int64_t r = gt; r < k; r += nthreads) {
      if (r < c - 1) continue;
      if (i > 0) {
        double vp;
        if (r == c - 1) {
          vp = 1.0;
          A[r + (c - 1) * lda] = beta_p;              // R's diagonal
        } else {
          vp = scl_p * A[r + (c - 1) * lda];
          A[r + (c - 1) * lda] = vp;                  // stored reflector tail
        }
        const double tv = tau_p * vp;
        for (int64_t cc = c; cc < pend; ++cc) A[r + cc * lda] = fma(-tv, wd[cc - c], A[r + cc * lda]);
      }
      if (!last) {
        const double v = A[r + c * lda];
        if (r == c) a.scal[0] = v;
        if (r > c) ss = fma(v, v, ss);
      }
    }
    if (last) break;
    ss = block_sum(ss, sh);
    if (threadIdx.x == 0) a.part[blockIdx.x] = ss;
    grid_barrier(a.bar, bar_target, nb);
    // ---------------- B
    const int64_t len = k - c;
    const double sumsq = grid_total(a.part, nb, sh);
    const double alpha = a.scal[0];
    double tau, beta, scl;
    householder_scalars(alpha, sumsq, len, tau, beta, scl);
    if (gt == 0) {
      a.tau[c] = tau;
      a.beta[c] = beta;
    }
    const double fix = 1.0 - scl * alpha;
    const double* u = A + c + c * lda;
    const int rem = int(pend - 1 - c);
    for (int64_t j = blockIdx.x; j < rem; j += gridDim.x) {
      const double* col = A + c + (c + 1 + j) * lda;
      double p = scl * cta_strided_warp_dot(col, u, len, dot_slots);
      if (lane == 0) {
        if (wid == 0) p = fma(col[0], fix, p);
        a.wpart[wid * kQrNb + j] = p;
      }
    }
    tau_p = tau;
    beta_p = beta;
    scl_p = scl;
    grid_barrier(a.bar, bar_target, nb);
  }
}


// ----------------------------------------------------------------------- cluster panel
// Unpivoted Householder panel held in DISTRIBUTED SHARED MEMORY.  The BLAS-2 panel of a
// tall matrix is pure latency: per column the grid-wide version above pays two grid barriers
// (~3 us each across 100+ CTAs) around a few microseconds of L2 traffic - 18.6 us per column
// measured at k = 11030.  Here ONE thread-block cluster (16 CTAs, non-portable size) owns the
// panel: CTA q keeps rows [q rpc, (q+1) rpc) x jb columns in its shared memory for the whole
// panel (<= 213 KB), partial dots are exchanged with DSMEM stores and ONE hardware cluster
// barrier per column (barrier.cluster, a few hundred ns) replaces the grid barriers.
// Per column i:
//   A  raw dots  d[cc] = sum_{g >= i} P[g, i] P[g, cc]  (cc > i), ss = sum_{g > i} P[g, i]^2:
//      warp w owns columns i + w, i + w + 16, ...; per-CTA partials are stored into EVERY CTA's
//      xch[parity][source][.]; the owner of panel row i broadcasts that row      | cluster barrier
//   B  every CTA adds the partials in rank order (bit-identical everywhere), forms the
//      Householder scalars and w = scl d + P[i, :] (1 - scl alpha), and updates its own rows
//      P[g, cc] -= tau v_g w[cc]; column i keeps the RAW x, scaled on the way out.
// On exit A holds R on and above the diagonal and the scaled reflector tails below it.
constexpr int kQcThreads = 512;
constexpr int kQcWarps = kQcThreads / 32;
constexpr int kQcMaxJb = 64;
constexpr int kQcMaxCluster = 16;
constexpr int kQcColsPerWarp = kQcMaxJb / kQcWarps;   // 4

struct QcArgs {
  double* A;
  int64_t lda, k, j0;
  int jb;
  int rpc;       // panel rows per CTA (leading dimension of the shared-memory slab)
  double* tau;
  double* beta;
};

__global__ void __launch_bounds__(kQcThreads, 1) qr_cluster_panel_kernel(QcArgs a) {
  extern __shared__ double slab[];                         // rpc x jb, column-major
  __shared__ double xch[2][kQcMaxCluster][kQcMaxJb];        // [parity][source CTA][cc - i]
  __shared__ double prow[2][kQcMaxJb];                      // panel row i, columns i.. (from its owner)
  __shared__ double dsum[kQcMaxJb];
  __shared__ double wv[kQcMaxJb];                           // tau * w, indexed by cc - i
  __shared__ double scl_s[kQcMaxJb], beta_s[kQcMaxJb];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = int(cluster.block_rank());
  const int csize = int(cluster.num_blocks());
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int jb = a.jb, ldp = a.rpc;
  const int64_t s = a.k - a.j0;                             // panel rows
  const int64_t row0 = int64_t(rank) * a.rpc;               // first panel row of this CTA
  const int nrows = int(imax_d(0, imin_d(int64_t(a.rpc), s - row0)));
  double* const Ap = a.A + a.j0 + a.j0 * a.lda;             // panel origin
  for (int cc = 0; cc < jb; ++cc)
    for (int rl = tid; rl < nrows; rl += kQcThreads) slab[rl + cc * ldp] = Ap[(row0 + rl) + int64_t(cc) * a.lda];
  __syncthreads();
  cluster.sync();                                           // every CTA of the cluster is running
  for (int i = 0; i < jb; ++i) {
    const int par = i & 1;
    const int ncols = jb - i;
    const int lo = int(imax_d(0, imin_d(int64_t(nrows), int64_t(i) - row0)));   // first local row with g >= i
    const bool own_diag = (int64_t(i) >= row0) && (int64_t(i) < row0 + nrows);
    const int rl_diag = int(int64_t(i) - row0);
    const double* pi = slab + i * ldp;
    // ---------------- A: raw dots against column i
    {
      double acc[kQcColsPerWarp];
      const double* pc[kQcColsPerWarp];
#pragma unroll
      for (int q = 0; q < kQcColsPerWarp; ++q) {
        acc[q] = 0.0;
        const int cc = i + wid + q * kQcWarps;
        pc[q] = slab + (cc < jb ? cc : i) * ldp;
      }
      for (int rl = lo + lane; rl < nrows; rl += 32) {
        const double x = pi[rl];
#pragma unroll
        for (int q = 0; q < kQcColsPerWarp; ++q) {
          double y = pc[q][rl];
          if (q == 0 && wid == 0 && own_diag && rl == rl_diag) y = 0.0;   // ss excludes the diagonal entry
          acc[q] = fma(x, y, acc[q]);
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
#pragma unroll
        for (int q = 0; q < kQcColsPerWarp; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
      }
      if (lane < csize) {
#pragma unroll
        for (int q = 0; q < kQcColsPerWarp; ++q) {
          const int c2 = wid + q * kQcWarps;
          if (c2 < ncols) *cluster.map_shared_rank(&xch[par][rank][c2], lane) = acc[q];
        }
      }
      if (own_diag) {
        for (int idx = tid; idx < ncols * csize; idx += kQcThreads) {
          const int q = idx / ncols, c2 = idx - q * ncols;
          *cluster.map_shared_rank(&prow[par][c2], q) = slab[rl_diag + (i + c2) * ldp];
        }
      }
    }
    cluster.sync();
    // ---------------- B: scalars, w, rank-1 update of the own rows
    // 16 lanes per column add the per-CTA partials in a fixed shuffle tree (deterministic, and the
    // same in every CTA); the serial 16-term chain cost ~600 cycles per column
    for (int cb = (tid >> 5) * 2; cb < ncols; cb += kQcThreads / 16) {     // warp-uniform trip count
      const int c2 = cb + ((tid >> 4) & 1), q = tid & 15;
      const bool valid = c2 < ncols;
      double t = (valid && q < csize) ? xch[par][q][c2] : 0.0;
      t += __shfl_xor_sync(0xffffffffu, t, 8);
      t += __shfl_xor_sync(0xffffffffu, t, 4);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      if (valid && q == 0) dsum[c2] = t;
    }
    __syncthreads();
    const double alpha = prow[par][0];
    double tau, beta, scl;
    householder_scalars(alpha, dsum[0], s - i, tau, beta, scl);
    const double fix = 1.0 - scl * alpha;
    if (tid >= 1 && tid < ncols) wv[tid] = tau * fma(scl, dsum[tid], prow[par][tid] * fix);
    if (tid == 0) {
      scl_s[i] = scl;
      beta_s[i] = beta;
      if (rank == 0) {
        a.tau[a.j0 + i] = tau;
        a.beta[a.j0 + i] = beta;
      }
    }
    __syncthreads();
    if (tau != 0.0) {
      double* pcw[kQcColsPerWarp];
      double wq[kQcColsPerWarp];
#pragma unroll
      for (int q = 0; q < kQcColsPerWarp; ++q) {
        const int c2 = 1 + wid + q * kQcWarps;
        pcw[q] = (c2 < ncols) ? slab + (i + c2) * ldp : nullptr;
        wq[q] = (c2 < ncols) ? wv[c2] : 0.0;
      }
      for (int rl = lo + lane; rl < nrows; rl += 32) {
        const double v = (own_diag && rl == rl_diag) ? 1.0 : scl * pi[rl];
#pragma unroll
        for (int q = 0; q < kQcColsPerWarp; ++q)
          if (pcw[q]) pcw[q][rl] = fma(-v, wq[q], pcw[q][rl]);
      }
    }
    __syncthreads();
  }
  // ---------------- write-out: R above / on the diagonal, scaled reflector tails below
  for (int cc = 0; cc < jb; ++cc) {
    const double sc = scl_s[cc], be = beta_s[cc];
    for (int rl = tid; rl < nrows; rl += kQcThreads) {
      const int64_t g = row0 + rl;
      double v = slab[rl + cc * ldp];
      if (g > cc) v *= sc;
      else if (g == cc) v = be;
      Ap[g + int64_t(cc) * a.lda] = v;
    }
  }
  cluster.sync();     // no CTA leaves while a sibling could still touch its shared memory
}

// Cluster geometry for a panel with s rows: the widest of 64 / 32 / 16 columns whose slab fits.
struct QcPlan {
  int csize = 0;       // 0: the cluster kernel cannot be used on this device
  size_t max_dyn = 0;
};

static int qc_plan(QcPlan** out) {
  static thread_local QcPlan plans[kMaxDevices];
  static thread_local bool done_dev[kMaxDevices] = {};
  QcPlan& plan = plans[device_slot()];
  bool& done = done_dev[device_slot()];
  *out = &plan;
  if (done) return TQ_OK;
  done = true;
  int dev = 0, optin = 0;
  TQ_CUDA_CHECK(cudaGetDevice(&dev));
  TQ_CUDA_CHECK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  cudaFuncAttributes fa;
  TQ_CUDA_CHECK(cudaFuncGetAttributes(&fa, qr_cluster_panel_kernel));
  const size_t max_dyn = size_t(optin) - fa.sharedSizeBytes - 1024;
  TQ_CUDA_CHECK(cudaFuncSetAttribute(qr_cluster_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     int(max_dyn)));
  TQ_CUDA_CHECK(cudaFuncSetAttribute(qr_cluster_panel_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (int cs = kQcMaxCluster; cs >= 8; cs >>= 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs);
    cfg.blockDim = dim3(kQcThreads);
    cfg.dynamicSmemBytes = max_dyn;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, qr_cluster_panel_kernel, &cfg) == cudaSuccess && nclusters >= 1) {
      plan.csize = cs;
      plan.max_dyn = max_dyn;
      break;
    }
    cudaGetLastError();
  }
  return TQ_OK;
}

// widest cluster panel (64 / 32 / 16 columns) for s rows; 0 when even 16 columns do not fit
static int qc_width(const QcPlan& plan, int64_t s, int* rpc_out) {
  if (plan.csize == 0) return 0;
  const int64_t rpc = (ceil_div(s, plan.csize) + 1) / 2 * 2;
  for (int jb = kQcMaxJb; jb >= 16; jb >>= 1) {
    if (size_t(rpc) * jb * sizeof(double) <= plan.max_dyn) {
      *rpc_out = int(rpc);
      return jb;
    }
  }
  return 0;
}

static int qc_launch(cudaStream_t st, const QcPlan& plan, const QcArgs& args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan.csize);
  cfg.blockDim = dim3(kQcThreads);
  cfg.dynamicSmemBytes = size_t(args.rpc) * args.jb * sizeof(double);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = plan.csize;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const int pslot = prof_begin_launch(st, double(args.k - args.j0) * args.jb * 16.0, TQ_PROF_QR_CLUSTER);
  TQ_CUDA_CHECK(cudaLaunchKernelEx(&cfg, qr_cluster_panel_kernel, args));
  prof_end_launch(st, pslot);
  ++g_launch_count;
  return TQ_OK;
}

struct QrcpPanelArgs {
  double* A;
  int64_t lda, k, n;
  int64_t j0;
  int jb;
  double* F;
  int64_t ldf;
  int64_t* perm;
  double* vn1;
  double* vn2;
  double* tau;
  double* beta;
  double* auxpart;  // kQrPanelWarps x kQrcpNb per-warp partials of A[c:, j0:c]^T v
  double* fpart;    // kQrPanelWarps x n       per-warp partials of A[c:, c+1:]^T v
  double* part;
  double* scal;
  QrcpCtl* ctl;
  double tol3z;
  unsigned int* bar;
  int trace;       // scal[8..15]: phase cycle counters
};

// DLAQPS panel, 3 barriers per column:
//   A  pivot = first argmax of vn1[c:] (every CTA, same result); swap columns pvt <-> c and
//      apply the panel's earlier reflectors to column c (own rows); partial norm  | barrier
//   B  scalars; F(:, i) = A[c:, c+1:]^T v and aux = A[c:, j0:c]^T v via raw dots (one CTA
//      per column); CTA 0 swaps perm / norms / F rows                             | barrier
//   C  scale v (own rows); finish F(:, i), update the pivot row, downdate the partial norms
//      (thread per trailing column), flag cancellation                            | barrier
__global__ void __launch_bounds__(kQrPanelThreads, 2) qrcp_panel_kernel(QrcpPanelArgs a) {
  extern __shared__ double2 dot_slots[];   // [kAsyncDepth][blockDim] cp.async staging of the streamed column
  __shared__ double sh[32];
  __shared__ double sval[32];
  __shared__ int64_t sidx[32];
  __shared__ int64_t spvt;
  __shared__ double frow[kQrcpNb];
  __shared__ double arow[kQrcpNb];
  double* const A = a.A;
  double* const F = a.F;
  const int64_t lda = a.lda, ldf = a.ldf, k = a.k, n = a.n, j0 = a.j0;
  const int64_t gt = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthreads = int64_t(gridDim.x) * blockDim.x;
  const unsigned int nb = gridDim.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned int bar_target = 0;
  long long tk = a.trace ? clock64() : 0;
#define TQ_PHASE(idx)                                   \
  if (a.trace && gt == 0) {                             \
    const long long now = clock64();                    \
    a.scal[8 + (idx)] += double(now - tk);              \
    tk = now;                                           \
  }
  for (int i = 0; i < a.jb; ++i) {
    const int64_t c = j0 + i;
    // ---------------- A: pivot
    {
      double best = -1.0;
      int64_t bidx = n;
      for (int64_t j = c + threadIdx.x; j < n; j += blockDim.x) {
        const double v = a.vn1[j];
        if (v > best) {
          best = v;
          bidx = j;
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int64_t oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ov > best || (ov == best && oi < bidx)) {
          best = ov;
          bidx = oi;
        }
      }
      if (lane == 0) {
        sval[wid] = best;
        sidx[wid] = bidx;
      }
      __syncthreads();
      if (wid == 0) {
        best = (lane < (blockDim.x >> 5)) ? sval[lane] : -2.0;
        bidx = (lane < (blockDim.x >> 5)) ? sidx[lane] : n;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int64_t oi = __shfl_xor_sync(0xffffffffu, bidx, o);
          if (ov > best || (ov == best && oi < bidx)) {
            best = ov;
            bidx = oi;
          }
        }
        if (lane == 0) spvt = (bidx >= n) ? c : bidx;
      }
      __syncthreads();
    }
    const int64_t pvt = spvt;
    if (threadIdx.x < i) frow[threadIdx.x] = F[(pvt - j0) + int64_t(threadIdx.x) * ldf];   // F row of the pivot column
    __syncthreads();
    // swap + column update + partial norm
    double ss = 0.0;
    for (int64_t r = gt; r < k; r += nthreads) {
      double ac = A[r + c * lda];
      if (pvt != c) {
        const double ap = A[r + pvt * lda];
        A[r + pvt * lda] = ac;
        ac = ap;
      }
      if (r >= c) {
        double s0 = 0.0, s1 = 0.0;
        int t = 0;
        for (; t + 1 < i; t += 2) {
          s0 = fma(A[r + (j0 + t) * lda], frow[t], s0);
          s1 = fma(A[r + (j0 + t + 1) * lda], frow[t + 1], s1);
        }
        if (t < i) s0 = fma(A[r + (j0 + t) * lda], frow[t], s0);
        ac -= (s0 + s1);
        if (r == c) a.scal[0] = ac;
        else ss = fma(ac, ac, ss);
      }
      A[r + c * lda] = ac;
    }
    ss = block_sum(ss, sh);
    if (threadIdx.x == 0) a.part[blockIdx.x] = ss;
    TQ_PHASE(0)
    grid_barrier(a.bar, bar_target, nb);
    TQ_PHASE(1)
    // ---------------- B
    const int64_t len = k - c;
    const double sumsq = grid_total(a.part, nb, sh);
    const double alpha = a.scal[0];
    double tau, beta, scl;
    householder_scalars(alpha, sumsq, len, tau, beta, scl);
    const double fix = 1.0 - scl * alpha;
    const int64_t ntrail = n - c - 1;
    {
      const double* u = A + c + c * lda;
      const int64_t total = ntrail + i;
      for (int64_t j = blockIdx.x; j < total; j += gridDim.x) {
        const double* col;
        double* out;
        if (j < ntrail) {
          col = A + c + (c + 1 + j) * lda;
          out = a.fpart + int64_t(wid) * n + (c + 1 + j);
        } else {
          col = A + c + (j0 + (j - ntrail)) * lda;
          out = a.auxpart + wid * kQrcpNb + (j - ntrail);
        }
        double p = scl * cta_strided_warp_dot(col, u, len, dot_slots);
        if (lane == 0) {
          if (wid == 0) p = fma(col[0], fix, p);
          *out = p;
        }
      }
    }
    if (blockIdx.x == gridDim.x - 1) {       // bookkeeping of the swap (nobody reads these in phase B)
      if (threadIdx.x == 0) {
        a.tau[c] = tau;
        a.beta[c] = beta;
        if (pvt != c) {
          const int64_t p = a.perm[pvt];
          a.perm[pvt] = a.perm[c];
          a.perm[c] = p;
          a.vn1[pvt] = a.vn1[c];
          a.vn2[pvt] = a.vn2[c];
        }
      }
      if (pvt != c && threadIdx.x < i) {
        const int t = threadIdx.x;
        const double fa = F[(pvt - j0) + int64_t(t) * ldf], fb = F[(c - j0) + int64_t(t) * ldf];
        F[(pvt - j0) + int64_t(t) * ldf] = fb;
        F[(c - j0) + int64_t(t) * ldf] = fa;
      }
    }
    TQ_PHASE(2)
    grid_barrier(a.bar, bar_target, nb);
    TQ_PHASE(3)
    // ---------------- C
    if (threadIdx.x < i) {
      double asum = 0.0;
      for (int w = 0; w < kQrPanelWarps; ++w) asum += a.auxpart[w * kQrcpNb + threadIdx.x];
      frow[threadIdx.x] = -tau * asum;                              // auxv
      arow[threadIdx.x] = A[c + (j0 + threadIdx.x) * lda];          // A[c, j0:c]
    }
    __syncthreads();
    for (int64_t r = gt; r < k; r += nthreads) {                    // v in place, beta on the diagonal
      if (r == c) A[r + c * lda] = beta;
      else if (r > c) A[r + c * lda] *= scl;
    }
    for (int64_t q = gt; q < ntrail; q += nthreads) {
      const int64_t col = c + 1 + q;
      const int64_t fr = col - j0;
      double fraw = 0.0;
      for (int w = 0; w < kQrPanelWarps; ++w) fraw += a.fpart[int64_t(w) * n + col];
      double f = tau * fraw;
      for (int t = 0; t < i; ++t) f = fma(F[fr + t * ldf], frow[t], f);
      F[fr + int64_t(i) * ldf] = f;
      double s = f;                                                 // v[c] = 1
      for (int t = 0; t < i; ++t) s = fma(F[fr + t * ldf], arow[t], s);
      const double av = A[c + col * lda] - s;
      A[c + col * lda] = av;
      if (c < k - 1) {
        const double v1 = a.vn1[col];
        if (v1 != 0.0) {
          double temp = fabs(av) / v1;
          temp = fmax(0.0, (1.0 + temp) * (1.0 - temp));
          const double qq = v1 / a.vn2[col];
          const double temp2 = temp * (qq * qq);
          if (temp2 <= a.tol3z) {
            a.vn2[col] = -1.0;
            a.ctl->stop_next = 1;
          } else {
            a.vn1[col] = v1 * sqrt(temp);
          }
        }
      }
    }
    if (gt == 0) a.ctl->kb = i + 1;
    TQ_PHASE(4)
    grid_barrier(a.bar, bar_target, nb);
    TQ_PHASE(5)
    if (a.ctl->stop_next) break;                 // uniform: read after the barrier
  }
#undef TQ_PHASE
}

// in place: on exit triu(A[0:k, 0:n]) = R (diagonal sign arbitrary)
// Two-level blocking: inner panels of 64 / 32 / 16 columns are factored by the cluster kernel
// (or, when a slab does not fit, by the grid-barrier kernel) and applied only inside the current
// outer block of kQrOb = 128 columns; the trailing matrix sees ONE compact-WY update per outer
// block, so its DGEMMs run with K = 128 (30-35 TF/s) instead of K = 64 / 32 (20 / 12 TF/s).
constexpr int kQrOb = 128;

static int qr_block_update(cublasHandle_t h, cudaStream_t st, double* A, int64_t lda, int64_t k, int64_t j0, int jb,
                           int64_t c0, int64_t nc, const double* tau, double* Vc, double* G, double* T, double* w1,
                           double* w2) {
  if (nc <= 0) return TQ_OK;
  const int64_t s = k - j0;
  dim3 grid((unsigned)imin(ceil_div(s, 256), 256), (unsigned)jb);
  copy_reflectors_kernel<<<grid, 256, 0, st>>>(A + j0 + j0 * lda, lda, s, jb, Vc, s);
  TQ_LAUNCH_CHECK();
  TQ_TRY(build_t_factor(h, st, Vc, s, s, jb, tau + j0, G, T));
  return apply_block_reflector(h, Vc, s, s, jb, T, jb, /*trans_t=*/true, A + j0 + c0 * lda, lda, nc, w1, w2);
}

// tau_out (optional, min(k, n) entries): keeps the Householder scalars for a caller that applies Q later
// (stage 1 of the two-stage tridiagonal reduction factors tall panels, k > n, with it).
int qr_r_colmajor_tau(cublasHandle_t h, cudaStream_t st, double* A, int64_t lda, int64_t k, int64_t n,
                      Workspace& ws, double* tau_out) {
  const int64_t kk = imin(k, n);                 // number of reflectors
  double* tau = tau_out ? tau_out : ws.take<double>(kk);
  double* beta = ws.take<double>(k);
  double* wdot = ws.take<double>(kQrNb * kMaxChunks);
  double* part = ws.take<double>(1024);
  double* scal = ws.take<double>(8);
  unsigned int* bar = ws.take<unsigned int>(4);
  double* Vc = ws.take<double>(size_t(k) * kQrOb);
  double* G = ws.take<double>(kQrOb * kQrOb);
  double* T = ws.take<double>(kQrOb * kQrOb);
  double* w1 = ws.take<double>(size_t(kQrOb) * n);
  double* w2 = ws.take<double>(size_t(kQrOb) * n);
  if (ws.overflow) {
    set_error("qr_r: workspace too small");
    return TQ_ERR_WORKSPACE;
  }
  QcPlan* plan = nullptr;
  TQ_TRY(qc_plan(&plan));
  static thread_local int coop_per_sm_dev[kMaxDevices] = {};
  int& coop_per_sm = coop_per_sm_dev[device_slot()];
  if (!coop_per_sm) {
    int per_sm = 0;
    TQ_CUDA_CHECK(cudaFuncSetAttribute(qr_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       int(kQrPanelSmem)));
    TQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qr_panel_kernel, kQrPanelThreads,
                                                                kQrPanelSmem));
    if (per_sm < 1) {
      set_error("qr_r: panel kernel cannot be made resident");
      return TQ_ERR_CUDA;
    }
    coop_per_sm = per_sm > 2 ? 2 : per_sm;
  }
  const int coop_blocks = num_sms() * coop_per_sm;      // num_sms() honours tq_set_sm_budget
  for (int64_t j0 = 0; j0 < kk; j0 += kQrOb) {
    const int ob = int(imin(kQrOb, kk - j0));
    const int64_t oend = j0 + ob;
    int64_t jp = j0;
    while (jp < oend) {
      const int64_t rows = k - jp;
      int rpc = 0;
      const int cw = qc_width(*plan, rows, &rpc);
      int jb;
      if (cw > 0) {
        jb = int(imin(cw, oend - jp));
        QcArgs qa{A, lda, k, jp, jb, rpc, tau, beta};
        TQ_TRY(qc_launch(st, *plan, qa));
      } else {
        // a tall-skinny panel does not need the whole machine: fewer CTAs make the barriers cheaper
        jb = int(imin(kQrNb, oend - jp));
        int blocks = int(imin(coop_blocks, imax(8, ceil_div(rows, kQrPanelThreads) * 4)));
        TQ_CUDA_CHECK(cudaMemsetAsync(bar, 0, sizeof(unsigned int), st));
        QrPanelArgs pa{A, lda, k, jp, jb, tau, beta, wdot, part, scal, bar};
        void* kargs[] = {&pa};
        TQ_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)qr_panel_kernel, dim3(blocks), dim3(kQrPanelThreads),
                                                  kargs, kQrPanelSmem, st));
        ++g_launch_count;
        set_diag_kernel<<<1, kQrNb, 0, st>>>(A, lda, jp, jb, beta);
        TQ_LAUNCH_CHECK();
      }
      // inner update: the rest of the outer block
      TQ_TRY(qr_block_update(h, st, A, lda, k, jp, jb, jp + jb, oend - (jp + jb), tau, Vc, G, T, w1, w2));
      jp += jb;
    }
    // outer update: everything to the right of the outer block, K = ob
    TQ_TRY(qr_block_update(h, st, A, lda, k, j0, ob, oend, n - oend, tau, Vc, G, T, w1, w2));
  }
  return TQ_OK;
}

int qr_r_colmajor(cublasHandle_t h, cudaStream_t st, double* A, int64_t lda, int64_t k, int64_t n,
                  Workspace& ws) {
  return qr_r_colmajor_tau(h, st, A, lda, k, n, ws, nullptr);
}

// ------------------------------------------------------------------ pivoted QR (DLAQPS)
__global__ void col_norms_kernel(const double* __restrict__ A, int64_t lda, int64_t row0, int64_t k, int64_t n,
                                 double* __restrict__ vn1, double* __restrict__ vn2, int only_flagged) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t j = warp; j < n; j += nwarps) {
    if (only_flagged && vn2[j] >= 0.0) continue;
    const double* col = A + j * lda;
    double s = 0.0;
    for (int64_t r = row0 + lane; r < k; r += 32) s = fma(col[r], col[r], s);
    s = warp_sum(s);
    if (lane == 0) {
      double nrm = sqrt(s);
      vn1[j] = nrm;
      vn2[j] = nrm;
    }
  }
}

__global__ void init_perm_kernel(int64_t* perm, int64_t n) {
  int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j < n) perm[j] = j;
}




__global__ void qrcp_panel_begin_kernel(QrcpCtl* ctl) {
  ctl->stop = 0;
  ctl->stop_next = 0;
  ctl->kb = 0;
}

// in place: triu(A[0:k, :]) = R_x (sign arbitrary), perm = column pivots
int qrcp_colmajor(cublasHandle_t h, cudaStream_t st, double* A, int64_t lda, int64_t k, int64_t n, int64_t* perm,
                  Workspace& ws) {
  double* vn1 = ws.take<double>(n);
  double* vn2 = ws.take<double>(n);
  double* F = ws.take<double>(size_t(n) * kQrcpNb);
  double* auxraw = ws.take<double>(kQrcpNb * kMaxChunks);
  double* fpart = ws.take<double>(size_t(n) * kMaxChunks);
  double* tau = ws.take<double>(k + 1);
  double* beta = ws.take<double>(k + 1);
  QrcpCtl* ctl = ws.take<QrcpCtl>(1);
  double* part = ws.take<double>(1024);
  double* scal = ws.take<double>(16);
  unsigned int* bar = ws.take<unsigned int>(4);
  if (ws.overflow) {
    set_error("qrcp: workspace too small");
    return TQ_ERR_WORKSPACE;
  }
  const double eps = 1.1102230246251565e-16;   // dlamch('Epsilon')
  const double tol3z = sqrt(eps);
  const int64_t ldf = n;
  static thread_local int coop_per_sm_dev[kMaxDevices] = {};
  int& coop_per_sm = coop_per_sm_dev[device_slot()];
  if (!coop_per_sm) {
    int per_sm = 0;
    TQ_CUDA_CHECK(cudaFuncSetAttribute(qrcp_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       int(kQrPanelSmem)));
    TQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qrcp_panel_kernel, kQrPanelThreads,
                                                                kQrPanelSmem));
    if (per_sm < 1) {
      set_error("qrcp: panel kernel cannot be made resident");
      return TQ_ERR_CUDA;
    }
    coop_per_sm = per_sm > 2 ? 2 : per_sm;
  }
  const int coop_blocks = num_sms() * coop_per_sm;      // num_sms() honours tq_set_sm_budget
  TQ_CUDA_CHECK(cudaMemsetAsync(scal, 0, sizeof(double) * 16, st));
  init_perm_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(perm, n);
  TQ_LAUNCH_CHECK();
  col_norms_kernel<<<dots_grid(n), 256, 0, st>>>(A, lda, 0, k, n, vn1, vn2, 0);
  TQ_LAUNCH_CHECK();
  int64_t j0 = 0;
  while (j0 < k) {
    const int jb = int(imin(kQrcpNb, k - j0));
    qrcp_panel_begin_kernel<<<1, 1, 0, st>>>(ctl);
    TQ_LAUNCH_CHECK();
    {
      TQ_CUDA_CHECK(cudaMemsetAsync(bar, 0, sizeof(unsigned int), st));
      QrcpPanelArgs pa{A, lda, k, n, j0, jb, F, ldf, perm, vn1, vn2, tau, beta, auxraw, fpart, part, scal, ctl, tol3z, bar,
                       trace_enabled() ? 1 : 0};
      void* kargs[] = {&pa};
      double bytes = 0.0;      // algorithmic bytes of the panel: every step streams the trailing matrix once
      for (int i = 0; i < jb; ++i) bytes += double(k - (j0 + i)) * double(n - (j0 + i) - 1 + i) * 8.0;
      const int pslot = prof_begin_launch(st, bytes, TQ_PROF_QRCP_PANEL);
      TQ_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)qrcp_panel_kernel, dim3(coop_blocks), dim3(kQrPanelThreads),
                                                kargs, kQrPanelSmem, st));
      prof_end_launch(st, pslot);
      ++g_launch_count;
    }
    QrcpCtl hc;
    TQ_CUDA_CHECK(cudaMemcpyAsync(&hc, ctl, sizeof(QrcpCtl), cudaMemcpyDeviceToHost, st));
    TQ_CUDA_CHECK(cudaStreamSynchronize(st));
    const int kb = hc.kb;
    if (kb <= 0) {
      set_error("qrcp: panel at column %lld made no progress", (long long)j0);
      return TQ_ERR_NOCONV;
    }
    const int64_t r0 = j0 + kb;          // first row / column after the factored block
    if (r0 < k && r0 < n) {
      const double one = 1.0, mone = -1.0;
      TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, int(k - r0), int(n - r0), kb, &mone,
                                  A + r0 + j0 * lda, int(lda), F + kb, int(ldf), &one, A + r0 + r0 * lda,
                                  int(lda)));
    }
    if (hc.stop_next || hc.stop) {
      col_norms_kernel<<<dots_grid(n), 256, 0, st>>>(A, lda, r0, k, n, vn1, vn2, 1);
      TQ_LAUNCH_CHECK();
    }
    j0 = r0;
  }
  if (trace_enabled()) {
    double hcnt[16];
    cudaMemcpyAsync(hcnt, scal, sizeof(hcnt), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    fprintf(stderr, "[tq-trace] qrcp phase Mcycles (CTA 0): A %.1f  bar1 %.1f  B(stream) %.1f  bar2 %.1f  C %.1f  bar3 %.1f\n",
            hcnt[8] * 1e-6, hcnt[9] * 1e-6, hcnt[10] * 1e-6, hcnt[11] * 1e-6, hcnt[12] * 1e-6, hcnt[13] * 1e-6);
  }
  return TQ_OK;
}

}  // namespace tq

using namespace tq;

static size_t qr_ws_bytes(int64_t k, int64_t n) {
  size_t b = ws_bytes_for(size_t(k) * n, 8);                       // column-major working copy
  b += ws_bytes_for(k + 1, 8) * 4 + ws_bytes_for(kQrNb, 8) + ws_bytes_for(size_t(k) * kQrOb, 8);
  b += ws_bytes_for(kQrOb * kQrOb, 8) * 2 + ws_bytes_for(size_t(kQrOb) * n, 8) * 2;
  b += ws_bytes_for(n, 8) * (2 + kMaxChunks) + ws_bytes_for(size_t(n) * kQrcpNb, 8) + ws_bytes_for(2048, 8) * 4;
  return b;
}

namespace tq {
size_t qr_stage_ws_bytes(int64_t k, int64_t n) { return qr_ws_bytes(k, n); }
}

extern "C" int tq_qr_r(const double* A, int64_t lda, int64_t k, int64_t n, double* R, int64_t ldr, void* ws,
                       size_t ws_bytes, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(A && R && k > 0 && n >= k && lda >= n && ldr >= n, "tq_qr_r: bad arguments (k=%lld n=%lld)",
             (long long)k, (long long)n);
  cudaStream_t st = (cudaStream_t)stream;
  Workspace wsp(ws, ws_bytes);
  double* Ac = wsp.take<double>(size_t(k) * n);
  if (wsp.overflow) {
    set_error("tq_qr_r: workspace too small (need %zu bytes, see tq_solver_workspace)", qr_ws_bytes(k, n));
    return TQ_ERR_WORKSPACE;
  }
  dim3 tg((unsigned)ceil_div(n, 32), (unsigned)ceil_div(k, 32));
  rowmajor_to_colmajor_kernel<<<tg, dim3(32, 8), 0, st>>>(A, lda, k, n, Ac, k);
  TQ_LAUNCH_CHECK();
  cublasHandle_t h;
  TQ_TRY(get_cublas(&h, st));
  TQ_TRY(qr_r_colmajor(h, st, Ac, k, k, n, wsp));
  emit_r_kernel<<<tg, dim3(32, 8), 0, st>>>(Ac, k, k, n, R, ldr);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

extern "C" int tq_qrcp(const double* A, int64_t lda, int64_t k, int64_t n, double* Rx, int64_t ldr, int64_t* perm,
                       void* ws, size_t ws_bytes, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(A && Rx && perm && k > 0 && n >= k && lda >= n && ldr >= n, "tq_qrcp: bad arguments (k=%lld n=%lld)",
             (long long)k, (long long)n);
  cudaStream_t st = (cudaStream_t)stream;
  Workspace wsp(ws, ws_bytes);
  double* Ac = wsp.take<double>(size_t(k) * n);
  if (wsp.overflow) {
    set_error("tq_qrcp: workspace too small (need %zu bytes, see tq_solver_workspace)", qr_ws_bytes(k, n));
    return TQ_ERR_WORKSPACE;
  }
  dim3 tg((unsigned)ceil_div(n, 32), (unsigned)ceil_div(k, 32));
  rowmajor_to_colmajor_kernel<<<tg, dim3(32, 8), 0, st>>>(A, lda, k, n, Ac, k);
  TQ_LAUNCH_CHECK();
  cublasHandle_t h;
  TQ_TRY(get_cublas(&h, st));
  TQ_TRY(qrcp_colmajor(h, st, Ac, k, k, n, perm, wsp));
  emit_r_kernel<<<tg, dim3(32, 8), 0, st>>>(Ac, k, k, n, Rx, ldr);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}
