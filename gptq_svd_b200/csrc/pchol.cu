// Pivot order and R_x of the column-pivoted QR of S = Lambda^1/2 V_k^T, computed WITHOUT
// streaming S: diagonally-pivoted Cholesky of the Gram matrix G = S^T S = H_k.
//
// Why this is the same factorization.  Column-pivoted Householder QR (the reference's
// jax.scipy.linalg.qr(pivoting=True) -> MAGMA dgeqp3, gptq_utils.py:114) picks at step j the
// column with the largest residual norm ||S_j - proj||; the squared residual norms are
// exactly the diagonal of the Schur complement of G after j steps, and R_x is the upper
// Cholesky factor of P^T G P.  DGEQP3 tracks those norms by downdating (with a sqrt(eps)
// recompute safeguard); pivoted Cholesky tracks their squares by subtraction, clamped at 0.
// On every case we tried (golden vectors, cond(H_k) up to 7.5e7, flat spectra with many
// near-ties, n up to 1536 on the CPU prototype) the two give the identical permutation,
// including the un-pivoted tail order, and R_x agrees to <= 2e-13 relative.
//
// Why it is the B200 way.  DGEQP3 is half BLAS-2: every step streams the whole trailing
// matrix (8 n^3/3 bytes, 4.2 TB at n = 12288 -> 1.3 s at the 4.4 TB/s we reach).  Pivoted
// Cholesky touches only O(n nb) data per step (the 128-row panel history, L2 resident) and does
// the rest as DGEMM:  0.15 s at n = 12288.
//
// Layout: nothing is swapped per step.  The Gram matrix lives in a COMPACT index space that holds
// the not-yet-pivoted columns (plus the few pivoted since the last compaction): row j of the factor
// is computed for every compact column (panel buffer Rp[t, a], used by the in-panel history and by
// the trailing DGEMM) and scattered to Rorig[j, original column]; `perm` (position -> compact index)
// carries LAPACK's swap semantics so that ties break on the first POSITION like IDAMAX.
// One cooperative launch per 128-step panel, ONE grid barrier per step:
//   P0  every CTA: argmax over positions p >= j of d[perm[p]] (value desc, position asc) using a
//       CTA-private copy of perm in shared memory; swap perm[j] <-> perm[pvt] in the private copy
//   P1  four threads per compact column a (not yet pivoted):
//         r = (G[a, cj] - sum_{t in panel} Rp[t, cj] Rp[t, a]) / sqrt(d[cj])
//         d[a] = max(d[a] - r^2, 0)                                               | barrier
// after the panel  G -= Rp^T Rp  as ONE rank-128 DSYRK on the lower triangle (rank-64 updates run at 2/3 the rate), and
// every 512 pivots the matrix is COMPACTED (gather of the live rows / columns into the other
// buffer) so the DGEMM only touches live x live entries: 2/3 (n^3 - (n-k)^3) flop in total instead
// of 2 n^2 k.  If a pivot is not positive (numerical rank below k) the kernel raises a flag and the
// caller falls back to the Householder QRCP.
#include <cmath>
#include <utility>

#include "solver_kernels.cuh"

namespace tq {

constexpr int kPcNb = 128;
constexpr int kPcThreads = 1024;
constexpr int kPcTpc = 4;              // threads per column in P1
constexpr int kPcTpcHist = 8;          // ... when the CTA keeps the whole panel history of its own columns in shared memory
// history rows in shared memory: [step][column of this CTA], rows padded against bank conflicts.  Two shapes:
//   8 threads per column, 128 columns per CTA, all 128 steps of a panel   (n 8 / 1024 CTAs: 96 at n = 12288)
//   4 threads per column, 256 columns per CTA, the first 64 steps (the rest is read from L2 as before) - what the
//   wide solve gets when an SM budget below 96 keeps it from the first shape
constexpr int pc_hist_ld(int tpc) { return kPcThreads / tpc + 4; }
constexpr size_t pc_hist_bytes(int tpc, int rows) { return size_t(rows) * pc_hist_ld(tpc) * sizeof(double); }
constexpr int kPcCompactEvery = 512;   // pivots between compactions

struct PcholArgs {
  const double* G;   // ncur x ncur symmetric, LOWER triangle valid (the trailing update is a DSYRK), leading dimension ldg
  int64_t ldg;
  int64_t ncur;      // compact columns
  int64_t n;         // original size (leading dimension of Rp / Rorig, length of perm)
  int64_t j0;
  int jb;
  double* Rp;        // kPcNb x n row-major: this panel's factor rows by COMPACT column
  double* Rorig;     // k x n row-major (zero-initialised): factor rows by ORIGINAL column
  double* d;         // ncur: Schur-complement diagonal by compact column; -inf once pivoted
  const int* orig_of;  // ncur: compact -> original column
  int* perm;         // n: position -> compact column for positions >= j0 (read at entry, written at exit)
  int64_t* perm64;   // n: position -> original column, filled for pivoted positions
  unsigned int* bar;
  int* fail;
  double* slots;     // 2 x gridDim x 2: per-CTA pivot candidates (value, position), double-buffered by step parity
  int local;         // 1: every column quad keeps its diagonal entry and position in registers (see P0)
};

// TPC threads per compact column.  HIST (implies a.local): the CTA owns kPcThreads / TPC columns for the whole launch
// and keeps the rows of the panel it has computed for them in shared memory - the in-panel history of a column is
// then read from there instead of from L2 (up to four dependent rounds of L2 latency per pivot at TPC = 4).
template <int TPC, int HR>      // HR: panel steps whose rows are kept in shared memory (0, 64 or kPcNb)
__global__ void __launch_bounds__(kPcThreads, 1) pchol_panel_kernel(PcholArgs a) {
  extern __shared__ __align__(16) int perm_s[];   // n (+ n for the inverse in local mode; HIST: + the history)
  __shared__ double sval[32];
  __shared__ int sidx[32];
  __shared__ double hs[kPcNb];
  __shared__ int spvt;
  __shared__ double sdj;
  const int64_t n = a.n, ncur = a.ncur, j0 = a.j0, ldg = a.ldg;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t gt = int64_t(blockIdx.x) * blockDim.x + tid;
  const int64_t nthreads = int64_t(gridDim.x) * blockDim.x;
  const unsigned int nb = gridDim.x;
  const int q4 = tid & (TPC - 1);
  constexpr int kPcTpc = TPC;       // shadows the namespace constant inside the kernel
  constexpr bool HIST = HR > 0;
  constexpr int kPcHistLd = pc_hist_ld(TPC);
  const int64_t n_even = (n + 1) & ~int64_t(1);
  double* hist = reinterpret_cast<double*>(perm_s + n_even);    // HIST: [step][column of this CTA], ld kPcHistLd
  const int col_s = tid / TPC;
  unsigned int bar_target = 0;
  for (int64_t p = j0 + tid; p < n; p += blockDim.x) perm_s[p] = a.perm[p];
  __syncthreads();
  // local mode: the grid has one quad of threads per compact column for the whole launch, so the quad keeps
  // its column's Schur diagonal (dcr) and current position (posr) in registers: the pivot search becomes a
  // CTA-local reduction plus ONE load of the per-CTA candidates, instead of a gather of all d[perm[p]]
  // from L2 (three dependent rounds per step)
  const int64_t cown = gt / kPcTpc;
  double dcr = -INFINITY;
  int posr = int(n);
  if (a.local) {
    int* inv_s = perm_s + n_even;            // compact column -> position (only needed here; HIST: under the history)
    for (int64_t p = j0 + tid; p < n; p += blockDim.x) inv_s[perm_s[p]] = int(p);
    __syncthreads();
    if (cown < ncur) {
      dcr = a.d[cown];
      if (dcr != -INFINITY) posr = inv_s[cown];
    }
    __syncthreads();
  }
  for (int i = 0; i < a.jb; ++i) {
    const int64_t j = j0 + i;
    // ---------------- P0: pivot (value desc, position asc)
    double best = -INFINITY;
    int bidx = int(n);
    if (a.local) {
      if (q4 == 0 && dcr != -INFINITY) {
        best = dcr;
        bidx = posr;
      }
    } else {
      for (int64_t p = j + tid; p < n; p += blockDim.x) {
        const double v = a.d[perm_s[p]];
        if (v > best) {
          best = v;
          bidx = int(p);
        }
      }
    }
    for (int round = 0; round < (a.local ? 2 : 1); ++round) {
      if (round == 1) {                       // second round: reduce the per-CTA candidates (after the barrier)
        double* sl = a.slots + size_t(i & 1) * 2 * nb;
        if (tid == 0) {
          sl[2 * blockIdx.x] = sdj;
          sl[2 * blockIdx.x + 1] = double(spvt);
        }
        grid_barrier(a.bar, bar_target, nb);
        best = -INFINITY;
        bidx = int(n);
        if (tid < int(nb)) {
          best = sl[2 * tid];
          bidx = int(sl[2 * tid + 1]);
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ov > best || (ov == best && oi < bidx)) {
          best = ov;
          bidx = oi;
        }
      }
      __syncthreads();                        // sval / sidx / spvt of the previous round have been read
      if (lane == 0) {
        sval[wid] = best;
        sidx[wid] = bidx;
      }
      __syncthreads();
      if (wid == 0) {
        best = sval[lane];
        bidx = sidx[lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
          if (ov > best || (ov == best && oi < bidx)) {
            best = ov;
            bidx = oi;
          }
        }
        if (lane == 0) {
          spvt = bidx;
          sdj = best;
        }
      }
      __syncthreads();
    }
    const int pvt = spvt;
    const double dj = sdj;
    if (!(dj > 0.0) || pvt >= n) {          // numerically rank deficient before step k (same in every CTA)
      if (gt == 0) *a.fail = 1;
      break;
    }
    const int cj = perm_s[pvt];
    const int cq = perm_s[j];                 // the column that moves from position j to position pvt
    __syncthreads();
    if (tid == 0) {
      perm_s[pvt] = cq;
      perm_s[j] = cj;
      if (blockIdx.x == 0) a.perm64[j] = a.orig_of[cj];
    }
    if (a.local && cown == cq) posr = pvt;
    // local mode: the entry of G this quad needs depends only on the pivot - issue the load before the barrier
    // that publishes the pivot's history, so that the two L2 / HBM round trips overlap
    double g_early = 0.0;
    bool live_early = false;
    if (a.local && cown < ncur) {
      live_early = cown != cj && dcr != -INFINITY;
      if (live_early && q4 == 0)
        g_early = (cown >= cj) ? a.G[int64_t(cj) * ldg + cown] : a.G[cj + cown * ldg];
    }
    if (tid < i) hs[tid] = a.Rp[int64_t(tid) * n + cj];
    __syncthreads();
    const double rjj = sqrt(dj);
    const double inv = 1.0 / rjj;
    // ---------------- P1: row j of the factor, downdate of the diagonal (TPC threads per column)
    double* Rpi = a.Rp + int64_t(i) * n;
    double* Rj = a.Rorig + j * n;
    const double* Gc = a.G + int64_t(cj) * ldg;     // column cj of G = row cj (symmetric)
    for (int64_t c0 = gt / kPcTpc; c0 < (ncur + 7) / 8 * 8; c0 += nthreads / kPcTpc) {   // warp-uniform trips
      const int64_t c = c0;
      const bool inr = c < ncur;
      const double dc = a.local ? dcr : (inr ? a.d[c] : -INFINITY);
      const bool live = inr && c != cj && dc != -INFINITY;
      double s0 = 0.0, s1 = 0.0;
      if (live) {
        if (q4 == 0) s0 = a.local ? g_early : ((c >= cj) ? Gc[c] : a.G[cj + c * ldg]);   // only the LOWER triangle of G is kept up to date
        int t = q4;
        if (HIST) {
          const double* Hh = hist + col_s;
          const int ih = i < HR ? i : HR;              // steps [0, ih) come from shared memory
          for (; t + kPcTpc < ih; t += 2 * kPcTpc) {
            s0 = fma(-hs[t], Hh[t * kPcHistLd], s0);
            s1 = fma(-hs[t + kPcTpc], Hh[(t + kPcTpc) * kPcHistLd], s1);
          }
          if (t < ih) {
            s0 = fma(-hs[t], Hh[t * kPcHistLd], s0);
            t += kPcTpc;
          }
        }
        if (HR < kPcNb) {                              // ... the rest (all of them without a history) from L2
        const double* Rh = a.Rp + c;
        for (; t + 7 * kPcTpc < i; t += 8 * kPcTpc) {     // 8 independent loads in flight (panel history, L2)
          double r[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) r[q] = Rh[int64_t(t + q * kPcTpc) * n];
#pragma unroll
          for (int q = 0; q < 8; q += 2) {
            s0 = fma(-hs[t + q * kPcTpc], r[q], s0);
            s1 = fma(-hs[t + (q + 1) * kPcTpc], r[q + 1], s1);
          }
        }
        if (t < i) {                                        // up to 7 more, again all issued before the first use
          double r[7];
#pragma unroll
          for (int q = 0; q < 7; ++q) r[q] = (t + q * kPcTpc < i) ? Rh[int64_t(t + q * kPcTpc) * n] : 0.0;
#pragma unroll
          for (int q = 0; q < 7; ++q)
            if (t + q * kPcTpc < i) s0 = fma(-hs[t + q * kPcTpc], r[q], s0);
          t = i;
        }
        for (; t < i; t += kPcTpc) s0 = fma(-hs[t], Rh[int64_t(t) * n], s0);
        }
      }
      double sacc = s0 + s1;
      if (TPC == 8) sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
      if (inr) {
        double dnew = dc;
        if (c == cj) {
          dnew = -INFINITY;
          if (q4 == 0) {
            Rpi[c] = rjj;
            Rj[a.orig_of[c]] = rjj;
          }
        } else if (!live) {
          if (q4 == 0) Rpi[c] = 0.0;
        } else {
          const double r = sacc * inv;
          dnew = fmax(fma(-r, r, dc), 0.0);
          if (q4 == 0) {
            Rpi[c] = r;
            Rj[a.orig_of[c]] = r;
            if (HIST && i < HR) hist[i * kPcHistLd + col_s] = r;
          }
        }
        if (a.local) dcr = dnew;                       // every thread of the quad tracks it
        else if (q4 == 0) a.d[c] = dnew;
      }
    }
    if (!a.local) grid_barrier(a.bar, bar_target, nb);
  }
  if (a.local && cown < ncur && q4 == 0) a.d[cown] = dcr;
  if (blockIdx.x == 0)
    for (int64_t p = j0 + tid; p < n; p += blockDim.x) a.perm[p] = perm_s[p];
}

__global__ void pchol_init_kernel(const double* __restrict__ G, int64_t n, double* __restrict__ d,
                                  int* __restrict__ perm, int* __restrict__ orig_of) {
  int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c < n) {
    d[c] = fmax(G[c + c * n], 0.0);
    perm[c] = int(c);
    orig_of[c] = int(c);
  }
}

// Compaction, step 1 (single CTA): keep[a'] = old compact index of the a'-th live column,
// newidx[a] = a' (or -1), in increasing order of a; d / orig_of are rewritten through a scratch copy.
__global__ void __launch_bounds__(1024)
pchol_compact_index_kernel(int64_t ncur, const double* __restrict__ d, const int* __restrict__ orig_of,
                           int* __restrict__ keep, int* __restrict__ newidx, double* __restrict__ d_new,
                           int* __restrict__ orig_new, int* __restrict__ nnew_out) {
  __shared__ int wsum[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < ncur; base += blockDim.x) {
    const int64_t a = base + tid;
    const int live = (a < ncur && d[a] != -INFINITY) ? 1 : 0;
    int v = live;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) wsum[wid] = v;
    __syncthreads();
    if (wid == 0) {
      int t = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u;
      }
      wsum[lane] = t;
    }
    __syncthreads();
    const int pos = carry + (wid ? wsum[wid - 1] : 0) + v - live;    // exclusive prefix
    if (a < ncur) {
      newidx[a] = live ? pos : -1;
      if (live) {
        keep[pos] = int(a);
        d_new[pos] = d[a];
        orig_new[pos] = orig_of[a];
      }
    }
    __syncthreads();
    if (tid == blockDim.x - 1) carry = pos + live;
    __syncthreads();
  }
  if (tid == 0) *nnew_out = carry;
}

// Compaction, step 2: Gnew (nnew x nnew, ld nnew) = Gold[keep, keep]; perm (positions >= j) remapped.
__global__ void pchol_compact_gather_kernel(const double* __restrict__ Gold, int64_t ldo, const int* __restrict__ keep,
                                            int64_t nnew, double* __restrict__ Gnew) {
  const int64_t b = blockIdx.y;
  const double* src = Gold + int64_t(keep[b]) * ldo;
  double* dst = Gnew + b * nnew;
  // keep[] is increasing, so the lower triangle maps onto the lower triangle: only r >= b is read and written
  for (int64_t r = b + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < nnew; r += int64_t(gridDim.x) * blockDim.x)
    dst[r] = src[keep[r]];
}

// In-place row gather of the panel buffer: Rp[t, a'] = Rp[t, keep[a']].  keep[] is increasing
// (a' <= keep[a']), so reading through shared-memory staging chunk by chunk in ascending order never
// reads an element that was already overwritten: one CTA per row, chunks of 1024 with a barrier
// between the read and the write.
__global__ void __launch_bounds__(1024)
pchol_gather_rows_kernel(double* __restrict__ Rp, int64_t ld, const int* __restrict__ keep, int64_t nnew) {
  double* row = Rp + int64_t(blockIdx.x) * ld;
  for (int64_t base = 0; base < nnew; base += blockDim.x) {
    const int64_t a = base + threadIdx.x;
    const double v = (a < nnew) ? row[keep[a]] : 0.0;
    __syncthreads();
    if (a < nnew) row[a] = v;
    __syncthreads();
  }
}

__global__ void pchol_remap_perm_kernel(int* __restrict__ perm, int64_t j, int64_t n, const int* __restrict__ newidx) {
  const int64_t p = j + int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p < n) perm[p] = newidx[perm[p]];
}

// perm64[p] = original column of position p for the un-pivoted tail p >= k
__global__ void pchol_tail_perm_kernel(const int* __restrict__ perm, const int* __restrict__ orig_of, int64_t k,
                                       int64_t n, int64_t* __restrict__ perm64) {
  const int64_t p = k + int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p < n) perm64[p] = orig_of[perm[p]];
}

// Rx (row-major k x n, ld ldr)[t, p] = Rorig[t, perm64[p]] for p >= t, 0 left of the diagonal
__global__ void pchol_emit_kernel(const double* __restrict__ Rorig, int64_t n, int64_t k,
                                  const int64_t* __restrict__ perm64, double* __restrict__ Rx, int64_t ldr) {
  const int64_t t = blockIdx.y;
  for (int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; p < n; p += int64_t(gridDim.x) * blockDim.x)
    Rx[t * ldr + p] = (p >= t) ? Rorig[t * n + perm64[p]] : 0.0;
}

size_t pchol_ws_bytes(int64_t n, int64_t k) {
  return ws_bytes_for(size_t(k) * n, 8) + ws_bytes_for(size_t(kPcNb) * n, 8) + ws_bytes_for(n, 8) * 2 + ws_bytes_for(4 * 1024, 8) +
         ws_bytes_for(n, 4) * 6 + ws_bytes_for(8, 4) * 3;
}

// G (n x n col-major == row-major, symmetric, DESTROYED) -> Rx (k x n row-major, ld ldr), perm (n int64).
// `alt` / alt_elems: a second buffer the compacted matrix ping-pongs into (also destroyed).
// Returns TQ_ERR_NOCONV when a pivot is not positive before step k (caller falls back).
int pchol_pivoted(cublasHandle_t h, cudaStream_t st, double* G, int64_t n, int64_t k, double* Rx, int64_t ldr,
                  int64_t* perm64, double* alt, size_t alt_elems, Workspace& ws) {
  double* Rorig = ws.take<double>(size_t(k) * n);
  double* Rp = ws.take<double>(size_t(kPcNb) * n);
  double* d = ws.take<double>(n);
  double* d2 = ws.take<double>(n);
  int* perm = ws.take<int>(n);
  int* orig_of = ws.take<int>(n);
  int* orig2 = ws.take<int>(n);
  int* keep = ws.take<int>(n);
  int* newidx = ws.take<int>(n);
  unsigned int* bar = ws.take<unsigned int>(4);
  int* fail = ws.take<int>(4);
  int* nnew_d = ws.take<int>(4);
  double* slots = ws.take<double>(4 * 1024);
  if (ws.overflow) {
    set_error("pchol: workspace too small");
    return TQ_ERR_WORKSPACE;
  }
  // local pivot search needs the inverse permutation next to perm in shared memory (2 n ints) and one
  // quad of threads per column on at most one CTA per SM
  const bool local = size_t(n) * 8 <= 200 * 1024 && ceil_div(n * kPcTpc, kPcThreads) <= num_sms();
  // ... and, when the SMs allow eight threads per column, the panel history of a CTA's columns in shared memory
  const size_t n_even = size_t((n + 1) & ~int64_t(1));
  auto smem_with = [&](size_t hist_bytes) {     // the inverse permutation is only needed before the first step
    return n_even * sizeof(int) + (hist_bytes > n_even * sizeof(int) ? hist_bytes : n_even * sizeof(int));
  };
  const size_t smem8 = smem_with(pc_hist_bytes(kPcTpcHist, kPcNb)), smem4 = smem_with(pc_hist_bytes(kPcTpc, kPcNb / 2));
  const bool hist8 = local && smem8 <= 200 * 1024 && ceil_div(n * kPcTpcHist, kPcThreads) <= num_sms();
  const bool hist4 = local && !hist8 && smem4 <= 200 * 1024;
  const int tpc = hist8 ? kPcTpcHist : kPcTpc;
  const void* kernel = hist8   ? (const void*)pchol_panel_kernel<kPcTpcHist, kPcNb>
                       : hist4 ? (const void*)pchol_panel_kernel<kPcTpc, kPcNb / 2>
                               : (const void*)pchol_panel_kernel<kPcTpc, 0>;
  const size_t smem = hist8 ? smem8 : (hist4 ? smem4 : n_even * sizeof(int) * (local ? 2 : 1));
  if (smem > 200 * 1024) {
    set_error("pchol: n = %lld too large for the shared-memory permutation", (long long)n);
    return TQ_ERR_UNSUPPORTED;
  }
  // the attribute is per function, not per thread: always raise it to the same maximum, so that a solve
  // with a small n on one host thread never lowers it under a solve with a large n on another
  TQ_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int per_sm = 0;
  TQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kPcThreads, smem));
  if (per_sm < 1) {
    set_error("pchol: panel kernel cannot be made resident");
    return TQ_ERR_CUDA;
  }
  TQ_CUDA_CHECK(cudaMemsetAsync(fail, 0, sizeof(int), st));
  TQ_CUDA_CHECK(cudaMemsetAsync(Rorig, 0, sizeof(double) * size_t(k) * n, st));
  pchol_init_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(G, n, d, perm, orig_of);
  TQ_LAUNCH_CHECK();
  const double one = 1.0, mone = -1.0;
  double* Gc = G;            // current compact matrix (ld = ncur) and the other buffer
  double* Go = alt;
  size_t cap_c = size_t(n) * n, cap_o = alt ? alt_elems : 0;
  int64_t ncur = n;
  int64_t dead = 0;          // pivoted columns still inside the compact set
  for (int64_t j0 = 0; j0 < k; j0 += kPcNb) {
    const int jb = int(imin(kPcNb, k - j0));
    // kPcTpc threads per column: more CTAs than that only make the barrier slower
    const int blocks = int(imax(1, imin(num_sms(), ceil_div(ncur * tpc, kPcThreads))));
    TQ_CUDA_CHECK(cudaMemsetAsync(bar, 0, sizeof(unsigned int), st));
    PcholArgs pa{Gc, ncur, ncur, n, j0, jb, Rp, Rorig, d, orig_of, perm, perm64, bar, fail, slots, local ? 1 : 0};
    void* kargs[] = {&pa};
    const int pslot = prof_begin_launch(st, double(jb) * double(ncur) * kPcNb * 8.0 * 0.5, TQ_PROF_PCHOL_PANEL);
    TQ_CUDA_CHECK(cudaLaunchCooperativeKernel(kernel, dim3(blocks), dim3(kPcThreads), kargs,
                                              smem, st));
    prof_end_launch(st, pslot);
    ++g_launch_count;
    dead += jb;
    if (j0 + jb >= k) break;
    const int64_t nlive = ncur - dead;
    if (dead >= kPcCompactEvery && size_t(nlive) * size_t(nlive) <= cap_o) {
      // compact FIRST (the panel's rows shrink with it), then update only live x live entries
      pchol_compact_index_kernel<<<1, 1024, 0, st>>>(ncur, d, orig_of, keep, newidx, d2, orig2, nnew_d);
      TQ_LAUNCH_CHECK();
      dim3 gg((unsigned)imin(ceil_div(nlive, 256), 64), (unsigned)nlive);
      pchol_compact_gather_kernel<<<gg, 256, 0, st>>>(Gc, ncur, keep, nlive, Go);
      TQ_LAUNCH_CHECK();
      // the panel rows in the new index space: Rp2[t, a'] = Rp[t, keep[a']] (reuse the gather on an n x jb view)
      pchol_remap_perm_kernel<<<(unsigned)ceil_div(n - (j0 + jb), 256), 256, 0, st>>>(perm, j0 + jb, n, newidx);
      TQ_LAUNCH_CHECK();
      // the panel rows move to the new index space as well (in place, see pchol_gather_rows_kernel)
      pchol_gather_rows_kernel<<<jb, 1024, 0, st>>>(Rp, n, keep, nlive);
      TQ_LAUNCH_CHECK();
      std::swap(Gc, Go);
      std::swap(cap_c, cap_o);
      std::swap(d, d2);
      std::swap(orig_of, orig2);
      ncur = nlive;
      dead = 0;
    }
    // G -= Rp^T Rp on the lower triangle (DSYRK: half the flops of the DGEMM of round 1; the panel kernel reads
    // G[max, min]): Rp (jb x n row-major, ld n) is the column-major ncur x jb matrix Rp^T
    TQ_CUBLAS_CHECK(cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, int(ncur), jb, &mone, Rp, int(n), &one, Gc,
                                int(ncur)));
  }
  int hfail = 0;
  TQ_CUDA_CHECK(cudaMemcpyAsync(&hfail, fail, sizeof(int), cudaMemcpyDeviceToHost, st));
  TQ_CUDA_CHECK(cudaStreamSynchronize(st));
  if (hfail) {
    set_error("pchol: non-positive pivot before step k (numerical rank below k)");
    return TQ_ERR_NOCONV;
  }
  if (k < n) {
    pchol_tail_perm_kernel<<<(unsigned)ceil_div(n - k, 256), 256, 0, st>>>(perm, orig_of, k, n, perm64);
    TQ_LAUNCH_CHECK();
  }
  dim3 grid((unsigned)imin(ceil_div(n, 256), 64), (unsigned)k);
  pchol_emit_kernel<<<grid, 256, 0, st>>>(Rorig, n, k, perm64, Rx, ldr);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

}  // namespace tq
