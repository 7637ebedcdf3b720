// Kernels of the two-stage tridiagonal reduction (two_stage.cu) that carry non-trivial index arithmetic or
// inter-CTA protocol.  This header is also compiled FOR THE HOST by tests/emu/two_stage_emu.cpp (TQ_HOST_EMU: one
// OS thread per CUDA thread, pthread barriers for __syncthreads, the same source text), which is how the bulge-chase
// kernel was checked against the numpy model before its first run on a GPU.  Keep it free of CUDA-only constructs
// other than those the emulation shims define (see the top of tests/emu/two_stage_emu.cpp).
#pragma once

namespace tq {

constexpr int kBw = 64;                 // band width b; also the number of sweeps per Q2 group
constexpr int kLdb = 2 * kBw;           // rows of the band array: 0 <= i - c < 2 b
constexpr int kLds = kBw + 1;           // shared-memory leading dimension (rows and columns conflict-free)
constexpr int kChaseThreads = 256;
constexpr int kProgDone = 1 << 30;
constexpr int kChaseSmemDoubles = 2 * kBw * kLds + 2 * kBw + 4 * kBw + kChaseThreads / 32 + 2;
constexpr size_t kChaseSmem = size_t(kChaseSmemDoubles) * sizeof(double);
constexpr int kQ2H = 2 * kBw - 1;       // rows of a staircase block reflector
constexpr int kQ2Ld = 2 * kBw;          // its leading dimension
constexpr int kQ1Nb = 128;              // reflector columns per compact-WY block of the Q1 back-transformation


// ------------------------------------------------------------------------------------------------ band array
// Bd[(i - c) + c ldb] = A[i, c] for 0 <= i - c <= b (A column-major, lower triangle), zero in the bulge room.
// In the block column [j, j + b) the rows below j + b hold R of the panel QR exactly where i - c <= b.
__global__ void band_extract_kernel(const double* __restrict__ A, int64_t lda, int n, double* __restrict__ Bd) {
  const int c = blockIdx.x;
  for (int d = threadIdx.x; d < kLdb; d += blockDim.x)
    Bd[d + int64_t(c) * kLdb] = (d <= kBw && c + d < n) ? A[(c + d) + int64_t(c) * lda] : 0.0;
}

// A[c + r lda] = A[r + c lda] for r > c (column-major s x s, lower triangle valid): makes the trailing matrix of the
// band reduction fully symmetric again after a DSYR2K on its lower triangle, for the DGEMM variant of X = A22 V.
// 256 threads, 32 x 32 tiles through shared memory so that reads and writes are both coalesced.
__global__ void mirror_lower_colmajor_kernel(double* __restrict__ A, int64_t lda, int s) {
  TQ_DYN_SMEM(double, mirror_sm);                           // 32 x 33
  const int bx = blockIdx.x, by = blockIdx.y;               // tile (row block bx, column block by), bx >= by
  if (bx < by) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8) {
    const int r = bx * 32 + tx, c = by * 32 + k;
    mirror_sm[k * 33 + tx] = (r < s && c < s) ? A[r + int64_t(c) * lda] : 0.0;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int r = bx * 32 + k, c = by * 32 + tx;            // source element (r, c), written to (c, r)
    if (r < s && c < s && r > c) A[c + int64_t(r) * lda] = mirror_sm[tx * 33 + k];
  }
}

__global__ void band_diag_kernel(const double* __restrict__ Bd, int n, double* __restrict__ d, double* __restrict__ e) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) {
    d[c] = Bd[int64_t(c) * kLdb];
    if (c < n - 1) e[c] = Bd[1 + int64_t(c) * kLdb];
  }
}

// ------------------------------------------------------------------------------------------------ bulge chase
struct ChaseArgs {
  double* Bd;
  int n;
  double* Vs;       // reflector (s, k): Vs[r0 + i + s ldv], r0 = s + 1 + k b  (column s = sweep s, stacked)
  int64_t ldv;
  double* tau2;     // tau2[s + k n]
  int* prog;        // prog[s] = k + 1 once task k of sweep s has written its G block back (earlier tasks: complete)
  long long* stats; // optional (TQ_TRACE): cycles CTA 0 spent {waiting, in steps 1-2, step 3, step 4}, its task count
};

#ifndef TQ_HOST_EMU
#ifndef TQ_DYN_SMEM
#define TQ_DYN_SMEM(type, name) extern __shared__ type name[]
#endif
__device__ __forceinline__ int ld_acquire_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_s32(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// CTA-scope release / acquire on a shared-memory word (the ticket between the compute warps and the helper warp)
__device__ __forceinline__ void st_release_cta_smem(int* p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(p))), "r"(v)
               : "memory");
}
__device__ __forceinline__ int ld_acquire_cta_smem(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];"
               : "=r"(v)
               : "r"(static_cast<unsigned>(__cvta_generic_to_shared(p)))
               : "memory");
  return v;
}
// barrier over the kChaseThreads compute threads only (named barrier 1; the helper warp never joins it)
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kChaseThreads) : "memory"); }
#endif

__device__ __forceinline__ void house_scalars(double alpha, double sumsq, int len, double& tau, double& beta,
                                              double& scl) {
  if (len <= 1 || sumsq == 0.0) {
    tau = 0.0;
    beta = alpha;
    scl = 0.0;
  } else {
    beta = -copysign(sqrt(fma(alpha, alpha, sumsq)), alpha);   // entries of a Hessian: no overflow guard needed
    tau = (beta - alpha) / beta;
    scl = 1.0 / (alpha - beta);
  }
}

// Task (s, k): reflector rows [r0, r1), r0 = s + 1 + k b.  Three blocks of the band array, each a dense matrix with
// leading dimension ldb - 1 (element (li, lc) of the block with rows R0.. and columns C0.. lives at
// (R0 - C0) + C0 ldb + li + lc (ldb - 1)):
//   G  rows [r0, r1) x columns [r0 - b, r0): the bulge.  It is the E block of task k - 1 and arrives in shared
//      memory; its first column defines the reflector, the rest gets H from the left (k = 0: only column s);
//   D  rows / columns [r0, r1): H D H on the lower triangle;
//   E  rows [r1, r1 + b) x columns [r0, r1): E H, kept in shared memory as the G of task k + 1.
// Thread (ti, tq) = (tid % 64, tid / 64) owns row ti, columns tq + 4 m (m < 16) of D and E in registers; the loads
// are issued before the reflector is formed so that their L2 latency overlaps steps 1 and 2.
//
// kHelper = true (TQ_CHASE_HELPER=1, not validated on a GPU yet): a ninth warp owns the progress counter.  Measured
// on the B200 the reflector / G step is the longest of a task (6.4k of 10.7k cycles) and counts three times in the
// dependency chain between sweeps; a good part of it is thread 0's gpu-scope fence + release store, for which all
// other threads then wait at the next barrier.  With the helper the compute warps only bump a ticket in shared
// memory (CTA-scope release) and go on; the helper acquires the ticket, fences and publishes.
#define TQ_CHASE_SYNC()   \
  do {                    \
    if (kHelper)          \
      compute_sync();     \
    else                  \
      __syncthreads();    \
  } while (0)
//
// kLate = true (TQ_CHASE_LATE=1, not validated on a GPU yet): the reflector step of a task k >= 1 only works on the
// block carried in shared memory, so it does not wait for the previous sweep at all (task 0 needs prog >= 2 for
// column s); the wait for prog >= k + 3 moves in front of the D / E loads, which then follow the reflector step
// instead of overlapping it.  A task gets longer by one L2 round trip, but the dependency chain between sweeps
// shrinks from 3a + 2d + 2e to 2 (a + L + d + e) - and the chase is bound by that chain (its CTAs wait half the
// time).  Both waits are valid and minimal under half-task interleavings in scripts/prototypes/sb2st_band.py.
template <bool kHelper, bool kLate>
__global__ void __launch_bounds__(kChaseThreads + (kHelper ? 32 : 0), 1) sb2st_chase_kernel_t(ChaseArgs a) {
  TQ_DYN_SMEM(double, chase_sm);
  double* const G = chase_sm;                   // kBw x kLds
  double* const D = G + kBw * kLds;             // kBw x kLds
  double* const vs = D + kBw * kLds;            // the reflector, zero beyond its length
  double* const wsh = vs + kBw;
  double(*const red)[kBw] = reinterpret_cast<double(*)[kBw]>(wsh + kBw);   // [4][kBw] partial sums
  double* const wred = wsh + kBw + 4 * kBw;     // one slot per warp
  double& alpha_s = wred[kChaseThreads / 32];
  int* const ticket = reinterpret_cast<int*>(wred + kChaseThreads / 32 + 1);   // publish events handed to the helper
  constexpr int b = kBw, ldg = kLdb - 1;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int ti = tid & (b - 1), tq = tid >> 6;
  const int n = a.n;
  int seq = 0;                                  // publish events of this CTA so far (thread 0)
  if (kHelper) {
    if (tid == 0) *ticket = 0;
    __syncthreads();                            // the only barrier all kChaseThreads + 32 threads share
    if (tid >= kChaseThreads) {
      if (tid != kChaseThreads) return;
      // helper: one publish event per task, in task order
      int ev = 0;
      for (int s = blockIdx.x; s < n - 2; s += gridDim.x) {
        const int K = (n - 3 - s) / b + 1;
        for (int k = 0; k < K; ++k) {
          ++ev;
          while (ld_acquire_cta_smem(ticket) < ev) {
          }
          __threadfence();
          st_release_s32(a.prog + s, k == K - 1 ? kProgDone : k + 1);
        }
      }
      return;
    }
  }
  for (int s = blockIdx.x; s < n - 2; s += gridDim.x) {
    const int K = (n - 3 - s) / b + 1;
    for (int k = 0; k < K; ++k) {
      const int r0 = s + 1 + k * b;
      const int r1 = min(r0 + b, n);
      const int ln = r1 - r0;                                       // >= 2
      const int ne = (ln == b) ? min(n, r1 + b) - r1 : 0;           // rows of E
      const bool last = (k == K - 1);                               // then ne <= 1, else ne >= 2
      const bool prof = a.stats != nullptr && blockIdx.x == 0 && tid == 0;
      long long tc0 = prof ? clock64() : 0, tc1;
      if (s > 0 && tid == 0 && (!kLate || k == 0)) {
        const int need = kLate ? 2 : k + 3;
        while (ld_acquire_s32(a.prog + (s - 1)) < need) {
        }
        __threadfence();
      }
      TQ_CHASE_SYNC();
      if (prof) {
        tc1 = clock64();
        a.stats[0] += tc1 - tc0;
        a.stats[4] += 1;
        tc0 = tc1;
      }
      // ---- loads of D (lower triangle) and E, L2 only (other SMs write these lines)
      double* const Dg = a.Bd + int64_t(r0) * kLdb;
      double* const Eg = a.Bd + ln + int64_t(r0) * kLdb;
      double dreg[16], ereg[16];
      if (!kLate) {
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const int lc = tq + 4 * m;
          dreg[m] = (ti >= lc && ti < ln) ? __ldcg(Dg + ti + lc * ldg) : 0.0;
          ereg[m] = (ti < ne) ? __ldcg(Eg + ti + lc * ldg) : 0.0;
        }
      }
      // ---- step 1: reflector from the first column of G
      double x = 0.0;
      if (tid < ln) x = (k == 0) ? __ldcg(a.Bd + 1 + int64_t(s) * kLdb + tid) : G[tid];
      double sq = (tid >= 1 && tid < ln) ? x * x : 0.0;
      sq = warp_sum(sq);
      if (lane == 0) wred[wid] = sq;
      if (tid == 0) alpha_s = x;
      TQ_CHASE_SYNC();
      double tau, beta, scl;
      house_scalars(alpha_s, wred[0] + wred[1], ln, tau, beta, scl);
      if (tid < b) {
        const double v = (tid == 0) ? 1.0 : ((tid < ln) ? x * scl : 0.0);
        vs[tid] = v;
        if (tid < ln) a.Vs[r0 + tid + int64_t(s) * a.ldv] = v;
      }
      if (tid == 0) a.tau2[s + int64_t(k) * n] = tau;
      TQ_CHASE_SYNC();
      // ---- step 2: G <- H G (columns 1..b-1), column 0 <- beta e_0; back to the band array
      if (k == 0) {
        if (tid < ln) __stcg(a.Bd + 1 + int64_t(s) * kLdb + tid, tid == 0 ? beta : 0.0);
      } else {
        double acc = 0.0;                                            // column ti, rows of quarter tq
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) acc = fma(vs[tq * 16 + ii], G[tq * 16 + ii + ti * kLds], acc);
        red[tq][ti] = acc;
        TQ_CHASE_SYNC();
        const double wj = tau * ((red[0][ti] + red[1][ti]) + (red[2][ti] + red[3][ti]));
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) {
          const int i = tq * 16 + ii;
          const double g = G[i + ti * kLds];
          G[i + ti * kLds] = (ti == 0) ? (i == 0 ? beta : 0.0) : fma(-vs[i], wj, g);
        }
        TQ_CHASE_SYNC();
        double* const Gg = a.Bd + b + int64_t(r0 - b) * kLdb;
        if (ti < ln) {
#pragma unroll
          for (int m = 0; m < 16; ++m) {
            const int lc = tq + 4 * m;
            __stcg(Gg + ti + lc * ldg, G[ti + lc * kLds]);
          }
        }
      }
      // G of this task is final and every earlier task of the sweep is complete: release the next sweep now.  Sweep
      // s + 1 never touches D or E of a task it has been released for (they are in registers already, and task
      // (s+1, k) meets task (s, k+2) in ONE entry, the first of G) - checked with half-task interleavings in
      // scripts/prototypes/sb2st_band.py.
      if (!last) {
        TQ_CHASE_SYNC();
        if (tid == 0) {
          if (kHelper) {
            st_release_cta_smem(ticket, ++seq);
          } else {
            __threadfence();
            st_release_s32(a.prog + s, k + 1);
          }
        }
      }
      if (prof) {
        tc1 = clock64();
        a.stats[1] += tc1 - tc0;
        tc0 = tc1;
      }
      if (kLate) {        // second wait: D and E of this task are final once the previous sweep has published k + 3
        if (s > 0 && tid == 0) {
          while (ld_acquire_s32(a.prog + (s - 1)) < k + 3) {
          }
          __threadfence();
        }
        TQ_CHASE_SYNC();
        if (prof) {
          tc1 = clock64();
          a.stats[0] += tc1 - tc0;
          tc0 = tc1;
        }
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const int lc = tq + 4 * m;
          dreg[m] = (ti >= lc && ti < ln) ? __ldcg(Dg + ti + lc * ldg) : 0.0;
          ereg[m] = (ti < ne) ? __ldcg(Eg + ti + lc * ldg) : 0.0;
        }
      }
      // ---- step 3: D <- H D H = D - v w^T - w v^T,  p = tau D v,  w = p - (tau p^T v / 2) v
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const int lc = tq + 4 * m;
        if (ti >= lc) {
          D[ti + lc * kLds] = dreg[m];
          if (ti > lc) D[lc + ti * kLds] = dreg[m];
        }
      }
      TQ_CHASE_SYNC();
      {
        double acc = 0.0;                                            // row ti, columns of quarter tq
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) acc = fma(D[ti + (tq * 16 + jj) * kLds], vs[tq * 16 + jj], acc);
        red[tq][ti] = acc;
      }
      TQ_CHASE_SYNC();
      double p = 0.0;
      if (tid < b) p = tau * ((red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]));
      double pv = (tid < b) ? p * vs[tid] : 0.0;
      pv = warp_sum(pv);
      if (lane == 0) wred[wid] = pv;
      TQ_CHASE_SYNC();
      if (tid < b) wsh[tid] = fma(-0.5 * tau * (wred[0] + wred[1]), vs[tid], p);
      TQ_CHASE_SYNC();
      if (ti < ln) {
        const double vi = vs[ti], wi = wsh[ti];
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const int lc = tq + 4 * m;
          if (ti >= lc) __stcg(Dg + ti + lc * ldg, dreg[m] - vi * wsh[lc] - wi * vs[lc]);
        }
      }
      if (prof) {
        tc1 = clock64();
        a.stats[2] += tc1 - tc0;
        tc0 = tc1;
      }
      // ---- step 4: E <- E H = E - (tau E v) v^T
      if (ne > 0) {
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < 16; ++m) acc = fma(ereg[m], vs[tq + 4 * m], acc);
        red[tq][ti] = acc;
        TQ_CHASE_SYNC();
        const double u = tau * ((red[0][ti] + red[1][ti]) + (red[2][ti] + red[3][ti]));
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const int lc = tq + 4 * m;
          const double ev = fma(-u, vs[lc], ereg[m]);
          if (!last)
            G[ti + lc * kLds] = ev;                                  // rows >= ne stay exactly zero
          else if (ti < ne)
            __stcg(Eg + ti + lc * ldg, ev);
        }
      }
      if (last) {
        TQ_CHASE_SYNC();
        if (tid == 0) {
          if (kHelper) {
            st_release_cta_smem(ticket, ++seq);
          } else {
            __threadfence();
            st_release_s32(a.prog + s, kProgDone);
          }
        }
      }
      if (prof) a.stats[3] += clock64() - tc0;
    }
  }
}
#undef TQ_CHASE_SYNC

// ------------------------------------------------------------------------------------------------ Q2 groups
// Group (sb, k) = reflectors (s, k), s in [sb b, (sb + 1) b): a staircase of b columns, column j non-zero in rows
// [j, j + b) counted from rlo = sb b + 1 + k b.  Clean copy (zeros outside the stairs, zero column and tau = 0 for
// a task that does not exist) of `count` groups (sb0 + i, k0 + 2 i).  The copy is SHIFTED DOWN BY ONE ROW: row 0 of
// the 128 x 64 block is zero and the block acts on the rows of Z from rlo - 1 = (sb + k) b, a multiple of 64 - so
// the batched DGEMMs see 512-byte aligned operands and K = M = 128 instead of an odd start row and 127.
// With `desc` the groups are taken from a table instead: group bi = (desc[2 bi], desc[2 bi + 1]).
__global__ void copy_staircase_kernel(const double* __restrict__ Vs, int64_t ldv, const double* __restrict__ tau2,
                                      int n, int sb0, int k0, double* __restrict__ Vc, double* __restrict__ taub,
                                      const int* __restrict__ desc = nullptr) {
  const int bi = blockIdx.z, j = blockIdx.y, r = int(threadIdx.x) - 1;   // blockDim.x == kQ2Ld; r = staircase row
  const int sb = desc ? desc[2 * bi] : sb0 + bi, k = desc ? desc[2 * bi + 1] : k0 + 2 * bi;
  const int s = sb * kBw + j;
  const int r0 = s + 1 + k * kBw;
  const int64_t rlo = int64_t(sb) * kBw + 1 + int64_t(k) * kBw;
  const bool exists = (s <= n - 3) && (r0 <= n - 2);
  const int ln = exists ? min(kBw, n - r0) : 0;
  const double v = (r >= j && r < j + ln) ? Vs[rlo + r + int64_t(s) * ldv] : 0.0;
  Vc[int64_t(bi) * kQ2Ld * kBw + (r + 1) + j * kQ2Ld] = v;
  if (threadIdx.x == 0) taub[bi * kBw + j] = exists ? tau2[s + int64_t(k) * n] : 0.0;
}

}  // namespace tq
