// Symmetric eigendecomposition in fp64 (reference: torch.linalg.eigh -> cuSOLVER syevd,
// gptq_utils.py:93), written from scratch for sm_100a:
//   1. sytrd  blocked Householder tridiagonal reduction (lower).  Per column the work is
//             BLAS-2 and HBM-bound (one pass over the trailing matrix per reflector);
//             per panel the rank-2k trailing update is DGEMM.
//   2. stedc  divide & conquer on the tridiagonal matrix: QL leaves (one warp per leaf),
//             then rank-one merges.  Deflation is decided on the host from d and z (O(n)
//             per merge); the secular equation (one warp per root, "middle way" rational
//             interpolation with a bisection safeguard), the Gu/Eisenstat z-hat recomputation
//             and the eigenvector formation run on the device; the eigenvector update is DGEMM.
//   3. ormtr  back-transformation Z <- Q Z with compact-WY block reflectors (DGEMM).
// All matrices are column-major with leading dimension n.
#include <algorithm>
#include <cmath>
#include <vector>

#include "solver_kernels.cuh"
#include "tcgen05.cuh"

namespace tq {

constexpr int kTrdNb = 64;   // sytrd panel width
constexpr int kLeaf = 128;   // D&C leaf size (one CTA of 128 threads; 32 in round 1: the merges of 48 - 192 rows it
                             // removes were pure launch / round-trip latency, ~100 us each, 384 of them at n = 12288)
constexpr int kOrmNb = 128;  // back-transform block (rank-128 DGEMM updates run at 30 TF/s, rank-64 at 20: profiles/r01_dgemm_probe.log)




// ----------------------------------------------------------------------- persistent panel
// One cooperative launch (four CTAs of 256 threads per SM) factors a whole panel of up to
// kTrdNb columns.  The per-column phases are separated by grid-wide barriers (grid_barrier)
// instead of kernel launches:
//   A   column update A[c:, c] -= V W[c,:]^T + W V[c,:]^T (own rows), partial sum of squares,
//       d[c]                                                                      | barrier
//   B   Householder scalars (every CTA, same order).  One CTA per trailing column streams it
//       (cp.async staged, see cta_strided_warp_dot) against the RAW column u = [alpha; x];
//       since v = [1; scl x], the per-warp partial is fixed up as
//       scl * p + col[0] (1 - scl alpha)  by the warp that owns row 0.  Per-warp partials of
//       y = A22 v, W^T v, V^T v go to ypart / tmppart; v^T y is accumulated on the fly.
//       (A variant that reads only the lower triangle - half the DRAM bytes - was measured
//       slower on B200, 1.58 s vs 1.51 s at n = 12288: it is instruction-bound on the per-column
//       warp reductions.  The tile-wise symmetric product that does pay is sytrd_panel_sym_kernel below;
//       this kernel remains for odd n and as the TQ_SYTRD_COLDOT=1 reference.)                  | barrier
//   C   v scaled in place (own rows), w = tau (y - V W^T v - W V^T v) - tau/2 (w.v) v with
//       w.v = tau (v^T y - 2 (W^T v).(V^T v)) known without another reduction.
// Row r is always handled by the same thread (r = global thread id + q * total threads), so
// values a thread wrote for its own rows need no barrier before it reads them again; the one
// foreign value the next column update needs, W[c+1, i], is recomputed by every CTA.
// 4 CTAs of 256 threads per SM; measured at n = 12288: 1.99 s (1 x 1024), 1.45 s (2 x 512), 1.33 s (4 x 256),
// 1.38 s (8 x 128)
constexpr int kPanelThreads = 256;
constexpr int kPanelWarps = kPanelThreads / 32;
constexpr size_t kPanelSmem = size_t(kAsyncDepth) * kPanelThreads * sizeof(double2);

struct TrdPanelArgs {
  double* A;
  int64_t n;
  int64_t j0;
  int jb;
  double* W;       // n x kTrdNb
  double* d;
  double* e;
  double* tau;
  double* ypart;   // kPanelWarps x n          per-warp partials of the column dots
  double* tmppart; // kPanelWarps x 2 kTrdNb   per-warp partials of W^T v | V^T v
  double* part;    // 2 x gridDim.x
  double* scal;    // scalar scratch; [8..15] phase cycle counters when tracing
  unsigned int* bar;
  int trace;
};

__device__ __forceinline__ int64_t imin_d(int64_t a, int64_t b) { return a < b ? a : b; }

__device__ __forceinline__ double grid_total(const double* part, int nb, double* sh) {
  double v = (threadIdx.x < nb) ? part[threadIdx.x] : 0.0;
  for (int q = threadIdx.x + blockDim.x; q < nb; q += blockDim.x) v += part[q];
  return block_sum(v, sh);
}

__global__ void __launch_bounds__(kPanelThreads, 4) sytrd_panel_kernel(TrdPanelArgs a) {
  extern __shared__ double2 dot_slots[];   // cp.async staging of the streamed columns
  __shared__ double sh[32];
  __shared__ double tmps[2 * kTrdNb];
  __shared__ double wrow_s;          // W[c, i-1], computed locally at the end of the previous column
  double* const A = a.A;
  double* const W = a.W;
  const int64_t n = a.n, lda = a.n, ldw = a.n, j0 = a.j0;
  const int tid = threadIdx.x;
  const int64_t gt = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthreads = int64_t(gridDim.x) * blockDim.x;
  const unsigned int nb = gridDim.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t gwarp = gt >> 5, nwarps = nthreads >> 5;
  const int sub = lane & 7, rsel = lane >> 3;      // row-wise phases: 4 rows per warp, 8 lanes per row
  double* part1 = a.part;
  double* part2 = a.part + nb;
  unsigned int bar_target = 0;
  if (threadIdx.x == 0) wrow_s = 0.0;
  __syncthreads();

  long long tk = a.trace ? clock64() : 0;
#define TQ_PHASE(idx)                                   \
  if (a.trace && gt == 0) {                             \
    const long long now = clock64();                    \
    a.scal[8 + (idx)] += double(now - tk);              \
    tk = now;                                           \
  }
  for (int i = 0; i < a.jb; ++i) {
    const int64_t c = j0 + i;
    // ---------------- A   (8 lanes per row: the panel history t < i is split over the lanes)
    double ss = 0.0;
    for (int64_t rb = 4 * gwarp; rb < n; rb += 4 * nwarps) {
      const int64_t r = rb + rsel;
      const bool act = (r < n) && (r >= c);
      double s0 = 0.0, s1 = 0.0;
      if (act) {
        for (int t = sub; t < i; t += 8) {
          const double wc = (t == i - 1) ? wrow_s : W[c + t * ldw];
          const double vc = (t == i - 1) ? 1.0 : A[c + (j0 + t) * lda];   // V[c, i-1] is the unit entry
          s0 = fma(A[r + (j0 + t) * lda], wc, s0);
          s1 = fma(W[r + t * ldw], vc, s1);
        }
      }
      double s = s0 + s1;
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (act && sub == 0) {
        const double v = A[r + c * lda] - s;
        A[r + c * lda] = v;
        if (r == c) a.d[c] = v;
        if (r == c + 1) a.scal[0] = v;
        if (r >= c + 2) ss = fma(v, v, ss);
      }
    }
    const int64_t len = n - c - 1;
    if (len <= 0) break;               // last column: only the diagonal entry (uniform across the grid)
    ss = block_sum(ss, sh);
    if (threadIdx.x == 0) part1[blockIdx.x] = ss;
    TQ_PHASE(0)
    grid_barrier(a.bar, bar_target, nb);
    TQ_PHASE(1)
    // ---------------- B
    const double sumsq = grid_total(part1, nb, sh);
    const double alpha = a.scal[0];
    double tau, beta, scl;
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
    const double fix = 1.0 - scl * alpha;
    const int64_t base = c + 1;
    const double* ucol = A + c * lda;              // raw reflector by GLOBAL row: u[g] = ucol[g], g >= base
    const double* u = ucol + base;                 // raw column [alpha; x]
    double vy = 0.0;                               // partial of v^T y (full path: lane 0 of each warp)
    {
      const int64_t total = len + 2 * i;
      for (int64_t j = blockIdx.x; j < total; j += gridDim.x) {
        const double* col;
        double* out;
        if (j < len) {
          col = A + base + (base + j) * lda;
          out = a.ypart + int64_t(wid) * n + (base + j);
        } else if (j < len + i) {
          col = W + base + (j - len) * ldw;
          out = a.tmppart + wid * (2 * kTrdNb) + (j - len);
        } else {
          col = A + base + (j0 + (j - len - i)) * lda;
          out = a.tmppart + wid * (2 * kTrdNb) + kTrdNb + (j - len - i);
        }
        double p = scl * cta_strided_warp_dot(col, u, len, dot_slots);
        if (lane == 0) {
          if (wid == 0) p = fma(col[0], fix, p);
          *out = p;
          if (j < len) vy = fma(p, (j == 0) ? 1.0 : scl * u[j], vy);
        }
      }
      vy = block_sum(lane == 0 ? vy : 0.0, sh);
      if (threadIdx.x == 0) part2[blockIdx.x] = vy;
      TQ_PHASE(2)
    }
    grid_barrier(a.bar, bar_target, nb);
    TQ_PHASE(3)
    // ---------------- C
    const double vtyv = grid_total(part2, nb, sh);
    if (threadIdx.x < 2 * kTrdNb) {     // tmp1 = W^T v | tmp2 = V^T v: fixed-order sum of the per-warp partials
      double tsum = 0.0;
      for (int w = 0; w < kPanelWarps; ++w) tsum += a.tmppart[w * (2 * kTrdNb) + threadIdx.x];
      tmps[threadIdx.x] = tsum;
    }
    __syncthreads();
    double cross = 0.0;
    for (int t = 0; t < i; ++t) cross = fma(tmps[t], tmps[kTrdNb + t], cross);
    const double wv = tau * (vtyv - 2.0 * cross);          // w'.v
    const double alpha2 = -0.5 * tau * wv;
    for (int64_t rb = 4 * gwarp; rb < n; rb += 4 * nwarps) {
      const int64_t r = rb + rsel;
      const bool act = (r < n) && (r >= c + 1);
      double acc = 0.0, s1 = 0.0;
      if (act) {
        for (int w = sub; w < kPanelWarps; w += 8) acc += a.ypart[int64_t(w) * n + r];
        for (int t = sub; t < i; t += 8) {
          acc = fma(-A[r + (j0 + t) * lda], tmps[t], acc);
          s1 = fma(-W[r + t * ldw], tmps[kTrdNb + t], s1);
        }
      }
      double ymw = acc + s1;                                  // y - V tmp1 - W tmp2
      ymw += __shfl_xor_sync(0xffffffffu, ymw, 4);
      ymw += __shfl_xor_sync(0xffffffffu, ymw, 2);
      ymw += __shfl_xor_sync(0xffffffffu, ymw, 1);
      if (act && sub == 0) {
        const double vr = (r == c + 1) ? 1.0 : scl * A[r + c * lda];
        A[r + c * lda] = vr;
        W[r + int64_t(i) * ldw] = fma(alpha2, vr, tau * ymw);
      }
    }
    __syncwarp();
    if (gt == 0) {
      a.tau[c] = tau;
      a.e[c] = beta;
    }
    if (wid == 0) {                    // W[c+1, i] for the next column update (v[c+1] = 1): every CTA repeats
      const int64_t r = c + 1;         // the owner's arithmetic (same lane split) so the value is bit-identical
      double acc = 0.0, s1 = 0.0;
      for (int w = sub; w < kPanelWarps; w += 8) acc += a.ypart[int64_t(w) * n + r];
      for (int t = sub; t < i; t += 8) {
        acc = fma(-A[r + (j0 + t) * lda], tmps[t], acc);
        s1 = fma(-W[r + t * ldw], tmps[kTrdNb + t], s1);
      }
      double ymw = acc + s1;
      ymw += __shfl_xor_sync(0xffffffffu, ymw, 4);
      ymw += __shfl_xor_sync(0xffffffffu, ymw, 2);
      ymw += __shfl_xor_sync(0xffffffffu, ymw, 1);
      if (lane == 0) wrow_s = fma(alpha2, 1.0, tau * ymw);
    }
    __syncthreads();
    TQ_PHASE(4)
  }
#undef TQ_PHASE
}

// ----------------------------------------------------------------------- symmetric TMA panel
// Same panel factorization, but the product y = A22 v reads only the LOWER triangle of the
// trailing matrix (half the DRAM bytes of the column-dot version above) and the data is staged
// by TMA instead of per-thread cp.async:
//   * the lower triangle is cut into tiles of 128 rows x 64 columns (64 KB); thread 0 of each
//     CTA feeds a 3-stage shared-memory ring with one cp.async.bulk.tensor.2d per tile (fp64
//     tensor map over A, out-of-bounds rows / columns read as zero), completion on mbarriers,
//     so 192 KB per SM are in flight without holding a single register;
//   * every tile element is used twice from shared memory: warp w owns columns 2w, 2w+1 of the
//     tile, lane l rows l, l+32, l+64, l+96:  yrow[r] += a u[c]  (kept in registers while the
//     CTA stays in the same 128-row block) and tcol[c] += a u[r] (warp-reduced per tile);
//     diagonal tiles mask r > c / r >= c;
//   * tiles are dealt to CTAs in contiguous chunks of the row-block-major order, so a CTA
//     flushes its row partial once per row block: rowpart[slot][row], slot = CTA - first CTA of
//     that row block (<= 27 slots per row), colpart[row block][column]; the consumer (phase C)
//     adds them in a fixed order - the result is deterministic, no atomics;
//   * W^T v and V^T v (panel history) are two more tiles per row block through the same ring.
// As above everything is computed against the RAW column u = [alpha; x]:  v = scl u + fix e0,
// y = scl (A22 u) + fix A22[:, 0],  v^T y = scl^2 u^T(A22 u) + 2 scl fix (A22 u)[0] + fix^2 A22[0, 0].
// 1 CTA of 1024 threads per SM; two grid barriers per column.
constexpr int kSymThreads = 1024;
constexpr int kSymWarps = kSymThreads / 32;
constexpr int kTileR = 128, kTileC = 64, kSymStages = 3;
constexpr int kTileElems = kTileR * kTileC;
constexpr int kRowSlots = 40;
constexpr size_t kSymSmem = size_t(kSymStages) * kTileElems * 8 + size_t(16) * kTileR * 8 + 128;

struct SymPanelArgs {
  double* A;
  int64_t n;
  int64_t j0;
  int jb;
  double* W;        // n x kTrdNb
  double* d;
  double* e;
  double* tau;
  double* rowpart;  // kRowSlots x n
  double* colpart;  // ceil(n / 128) x n
  double* wvpart;   // ceil(n / 128) x 2 kTrdNb   per-row-block partials of W^T u | V^T u
  double* part;     // 2 x gridDim.x
  double* scal;
  unsigned int* bar;
  int trace;
  int boxc;         // columns per TMA box (a tile is kTileC / boxc boxes)
  int align;        // row blocks start at multiples of `align` rows (TMA needs 16 bytes; 32 rows = one 256-byte L2 block)
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__global__ void __launch_bounds__(kSymThreads, 1)
sytrd_panel_sym_kernel(SymPanelArgs a, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW) {
  extern __shared__ unsigned char sym_smem_raw[];
  double* const tiles = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sym_smem_raw) + 127) & ~uintptr_t(127));
  double* const yred = tiles + kSymStages * kTileElems;        // 16 x 128
  __shared__ uint64_t full[kSymStages];
  __shared__ double sh[32];
  __shared__ double tmps[2 * kTrdNb];
  __shared__ double crow[2 * kTrdNb];   // W[c, t] | V[c, t], t < i: staged during phase C of the previous column
  __shared__ double wrow_s, yraw0_s, uty_s;
  double* const A = a.A;
  double* const W = a.W;
  const int64_t n = a.n, lda = a.n, ldw = a.n, j0 = a.j0;
  const int tid = threadIdx.x;
  const int64_t gt = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthreads = int64_t(gridDim.x) * blockDim.x;
  const unsigned int nb = gridDim.x;
  const int G = int(gridDim.x), bidx = int(blockIdx.x);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // row-wise phases: row group g (4 rows) belongs to warp (g / gridDim) of CTA (g % gridDim), so the
  // active rows r >= c stay spread over ALL SMs as c grows (a CTA-major split leaves the low CTAs idle)
  const int64_t nwarps = nthreads >> 5;
  const int64_t gwarp = int64_t(wid) * gridDim.x + blockIdx.x;
  const int sub = lane & 7, rsel = lane >> 3;
  double* part1 = a.part;
  double* part2 = a.part + nb;
  unsigned int bar_target = 0;
  unsigned int use = 0;        // tiles consumed by this CTA so far (ring position, all threads)
  unsigned int iss = 0;        // tiles issued so far (issuing thread only)
  const bool issuer = (tid == kSymThreads - 32);   // lane 0 of the last warp (rarely owns rows)
  const uint64_t l2_first = ptx::l2_policy_evict_first();
  if (tid == 0) {
    wrow_s = 0.0;
    for (int s = 0; s < kSymStages; ++s) ptx::mbar_init(&full[s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();

  long long tk = a.trace ? clock64() : 0;
  const bool tracer = (blockIdx.x == 0 && tid == 20 * 32);   // a warp that owns rows until c ~ 11840
#define TQ_PHASE(idx)                                   \
  if (a.trace && tracer) {                              \
    const long long now = clock64();                    \
    a.scal[8 + (idx)] += double(now - tk);              \
    tk = now;                                           \
  }
  for (int i = 0; i < a.jb; ++i) {
    const int64_t c = j0 + i;
    const int64_t len = n - c - 1;
    const int64_t base = c + 1;
    // ---- tile list of this column (depends only on len and i)
    // TMA needs a 16-byte aligned box origin, and a 1 KB column chunk that starts on a 256-byte boundary
    // costs exactly four 256-byte L2 blocks (a misaligned one five): row blocks start at the global row
    // base_e = base - dl, a multiple of `align` (32 when the matrix allows it, else 2); the dl rows above
    // the trailing matrix that fall into row block 0 are masked out
    const int dl = int(base % a.align);
    const int64_t base_e = base - dl;
    const int nrb = int((len + dl + kTileR - 1) / kTileR), nstrips = int((len + kTileC - 1) / kTileC);
    const int TA = nrb > 0 ? nrb * (nrb - 1) + min(2 * nrb, nstrips) : 0;
    const int T = TA + (i > 0 ? 2 * nrb : 0);
    const int ch = (T + G - 1) / G > 0 ? (T + G - 1) / G : 1;
    const int t0 = min(T, bidx * ch), t1 = min(T, t0 + ch);
    // tile t -> TMA coordinates; A-type tiles only touch the trailing matrix (read-only in this kernel)
    const bool stream_hint = len > 5000;      // lower triangle > 100 MB
    auto issue_tile = [&](int t) {
      const int stage = int(iss % kSymStages);
      double* dst = tiles + stage * kTileElems;
      ptx::mbar_expect_tx(&full[stage], kTileElems * 8);
      const CUtensorMap* map;
      int x, y;
      if (t < TA) {
        int rb = int((sqrt(4.0 * double(t) + 1.0) - 1.0) * 0.5);
        while (rb * (rb + 1) > t) --rb;
        while ((rb + 1) * (rb + 2) <= t) ++rb;
        const int cs = t - rb * (rb + 1);
        map = &tmA;
        x = int(base_e + int64_t(rb) * kTileR);
        y = int(base + int64_t(cs) * kTileC);
      } else if (t < TA + nrb) {
        map = &tmW;
        x = int(base_e + int64_t(t - TA) * kTileR);
        y = 0;
      } else {
        map = &tmA;
        x = int(base_e + int64_t(t - TA - nrb) * kTileR);
        y = int(j0);
      }
      // a trailing matrix far larger than L2 is never re-used from it: stream it with evict_first so the
      // partial sums and the panel history (phases A / C) stay resident
      if (stream_hint) {
        for (int cb = 0; cb < kTileC; cb += a.boxc)
          ptx::tma_load_2d_hint(dst + cb * kTileR, map, &full[stage], x, y + cb, l2_first);
      } else {
        for (int cb = 0; cb < kTileC; cb += a.boxc) ptx::tma_load_2d(dst + cb * kTileR, map, &full[stage], x, y + cb);
      }
      ++iss;
    };
    int issued = 0;
    // ---------------- A   (8 lanes per row: the panel history t < i is split over the lanes)
    double ss = 0.0;
    for (int64_t rb = 4 * gwarp; rb < n; rb += 4 * nwarps) {
      const int64_t r = rb + rsel;
      const bool act = (r < n) && (r >= c);
      double s0 = 0.0, s1 = 0.0, acur = 0.0;
      if (act) {
        // every load of this block is independent (i <= 64: at most 8 history entries per lane): ONE
        // DRAM / L2 round trip for the whole row; row c of the history comes from shared memory
        if (sub == 0) acur = A[r + c * lda];
        const double* ar = A + r + (j0 + sub) * lda;
        const double* wr = W + r + int64_t(sub) * ldw;
        double x[8], y[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const bool ok = sub + 8 * q < i;
          x[q] = ok ? ar[int64_t(8 * q) * lda] : 0.0;
          y[q] = ok ? wr[int64_t(8 * q) * ldw] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int t = sub + 8 * q;
          if (t < i) {
            s0 = fma(x[q], crow[t], s0);
            s1 = fma(y[q], crow[kTrdNb + t], s1);
          }
        }
      }
      double s = s0 + s1;
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (act && sub == 0) {
        const double v = acur - s;
        A[r + c * lda] = v;
        if (r == c) a.d[c] = v;
        if (r == c + 1) a.scal[0] = v;
        if (r >= c + 2) ss = fma(v, v, ss);
      }
    }
    if (len <= 0) break;               // last column: only the diagonal entry (uniform across the grid)
    // prefetch the first tiles of this column's product only now: issued at the top of the step, their
    // 28 MB (3 stages x 148 SMs) queue in front of the small phase-A loads, which then take 4 us
    if (issuer)
      while (issued < kSymStages && t0 + issued < t1 && t0 + issued < TA) issue_tile(t0 + issued++);
    TQ_PHASE(7)
    ss = block_sum(ss, sh);
    if (threadIdx.x == 0) part1[blockIdx.x] = ss;
    TQ_PHASE(0)
    grid_barrier(a.bar, bar_target, nb);
    TQ_PHASE(1)
    // ---------------- B
    if (issuer) {
      fence_proxy_async_all();           // W / V columns written with st by other CTAs -> TMA reads
      while (issued < kSymStages && t0 + issued < t1) issue_tile(t0 + issued++);
    }
    const double* u = A + base + c * lda;          // raw column [alpha; x], u[j], 0 <= j < len
    // A22[0, 0]: read BEFORE barrier 2 - its owner may already be in phase A of the next column (which
    // updates exactly this entry) while a slower CTA is still in phase C of this one
    const double a00 = A[base + base * lda];
    {
      double yrow[4] = {0.0, 0.0, 0.0, 0.0}, ur[4] = {0.0, 0.0, 0.0, 0.0};
      double uy = 0.0;
      int rb_rows = -1;        // row block whose row partials are being accumulated in yrow
      int rb_ur = -1;          // row block ur[] belongs to
      // flush the row partials of row block rb_rows: cross-warp sum in a fixed order
      auto flush_rows = [&]() {
#pragma unroll
        for (int q = 0; q < 4; ++q) uy = fma(yrow[q], ur[q], uy);
        if (wid >= 16) {
#pragma unroll
          for (int q = 0; q < 4; ++q) yred[(wid - 16) * kTileR + lane + 32 * q] = yrow[q];
        }
        __syncthreads();
        if (wid < 16) {
#pragma unroll
          for (int q = 0; q < 4; ++q) yred[wid * kTileR + lane + 32 * q] += yrow[q];
        }
        __syncthreads();
        if (tid < kTileR) {
          double sacc = 0.0;
#pragma unroll
          for (int w = 0; w < 16; ++w) sacc += yred[w * kTileR + tid];
          const int64_t rl = int64_t(rb_rows) * kTileR + tid - dl;
          const int slot = bidx - int((int64_t(rb_rows) * (rb_rows + 1)) / ch);
          if (rl >= 0 && rl < len) a.rowpart[int64_t(slot) * n + base + rl] = sacc;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) yrow[q] = 0.0;
        rb_rows = -1;
      };
      // decode the first tile; afterwards the (rb, cs) pair is advanced incrementally
      int rb = 0, cs = 0;
      if (t0 < TA) {
        rb = int((sqrt(4.0 * double(t0) + 1.0) - 1.0) * 0.5);
        while (rb * (rb + 1) > t0) --rb;
        while ((rb + 1) * (rb + 2) <= t0) ++rb;
        cs = t0 - rb * (rb + 1);
      }
      for (int t = t0; t < t1; ++t) {
        const bool a_tile = t < TA;
        const int trb = a_tile ? rb : (t < TA + nrb ? t - TA : t - TA - nrb);
        if (a_tile) {
          if (rb_rows >= 0 && rb_rows != trb) flush_rows();
        } else if (rb_rows >= 0) {
          flush_rows();
        }
        if (rb_ur != trb) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int64_t rl = int64_t(trb) * kTileR + lane + 32 * q - dl;
            ur[q] = (rl >= 0 && rl < len) ? u[rl] : 0.0;
          }
          rb_ur = trb;
        }
        const int stage = int(use % kSymStages);
        ptx::mbar_wait(&full[stage], (use / kSymStages) & 1);
        const double* ts = tiles + stage * kTileElems + (2 * wid) * kTileR + lane;
        double av[2][4];
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int q = 0; q < 4; ++q) av[e][q] = ts[e * kTileR + 32 * q];
        double tcol[2] = {0.0, 0.0};
        if (a_tile) {
          rb_rows = trb;
          const int64_t c0 = int64_t(cs) * kTileC + 2 * wid;        // local column of av[0][.]
          double uc[2];
          uc[0] = c0 < len ? u[c0] : 0.0;
          uc[1] = c0 + 1 < len ? u[c0 + 1] : 0.0;
          const int64_t r0 = int64_t(trb) * kTileR + lane - dl;     // local row of av[.][0] (-1: row c, masked)
          if (int64_t(trb) * kTileR - dl >= int64_t(cs + 1) * kTileC) {
            // whole tile strictly below the diagonal
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                yrow[q] = fma(av[e][q], uc[e], yrow[q]);
                tcol[e] = fma(av[e][q], ur[q], tcol[e]);
              }
          } else {
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int64_t rr = r0 + 32 * q, cc = c0 + e;
                if (rr >= cc) yrow[q] = fma(av[e][q], uc[e], yrow[q]);
                if (rr > cc) tcol[e] = fma(av[e][q], ur[q], tcol[e]);
              }
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            tcol[0] += __shfl_xor_sync(0xffffffffu, tcol[0], o);
            tcol[1] += __shfl_xor_sync(0xffffffffu, tcol[1], o);
          }
          if (lane == 0) {
            if (c0 < len) a.colpart[int64_t(trb) * n + base + c0] = tcol[0];
            if (c0 + 1 < len) a.colpart[int64_t(trb) * n + base + c0 + 1] = tcol[1];
            uy = fma(tcol[0], uc[0], uy);
            uy = fma(tcol[1], uc[1], uy);
          }
          if (++cs == min(2 * rb + 2, nstrips)) {
            ++rb;
            cs = 0;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int q = 0; q < 4; ++q) tcol[e] = fma(av[e][q], ur[q], tcol[e]);
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            tcol[0] += __shfl_xor_sync(0xffffffffu, tcol[0], o);
            tcol[1] += __shfl_xor_sync(0xffffffffu, tcol[1], o);
          }
          if (lane == 0) {
            double* out = a.wvpart + int64_t(trb) * (2 * kTrdNb) + (t < TA + nrb ? 0 : kTrdNb) + 2 * wid;
            out[0] = tcol[0];
            out[1] = tcol[1];
          }
        }
        __syncthreads();                 // every warp is done with this stage
        ++use;
        if (issuer && t + kSymStages < t1) issue_tile(t + kSymStages);
      }
      if (rb_rows >= 0) flush_rows();
      uy = block_sum(uy, sh);
      if (threadIdx.x == 0) part2[blockIdx.x] = uy;
    }
    // Householder scalars: not needed by the raw-u products above; part1 / scal[0] must be read before
    // barrier 2 (a faster CTA rewrites them in phase A of the next column)
    const double sumsq = grid_total(part1, nb, sh);
    const double alpha = a.scal[0];
    double tau, beta, scl;
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
    const double fix = 1.0 - scl * alpha;
    TQ_PHASE(2)
    grid_barrier(a.bar, bar_target, nb);
    TQ_PHASE(3)
    // ---------------- C
    // raw product (A22 u)[r - base] for global row r: this lane's share of the partials (8 lanes per row)
    auto yraw_lane = [&](int64_t r) -> double {
      const int64_t rl = r - base;
      const int rbr = int((rl + dl) / kTileR);           // row block holding this row
      const int64_t tstart = int64_t(rbr) * (rbr + 1);
      const int ncs = min(2 * rbr + 2, nstrips);
      const int b0 = int(tstart / ch), b1 = int((tstart + ncs - 1) / ch);
      const int rbc = int(rl / kTileC) / 2;              // first row block with a tile over this column
      const int nslots = b1 - b0 + 1, ncol = nrb - rbc;
      // fixed order: lane `sub` takes items sub, sub + 8, ...; four independent accumulators keep the
      // (L2) loads in flight
      const double* rp = a.rowpart + r;
      const double* cp = a.colpart + int64_t(rbc - nslots) * n + r;
      const int nit = nslots + ncol;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int it = sub;
      for (; it + 56 < nit; it += 64) {           // 8 independent loads per lane and round
        double x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = it + 8 * q;
          x[q] = (j < nslots) ? rp[int64_t(j) * n] : cp[int64_t(j) * n];
        }
        a0 += x[0];
        a1 += x[1];
        a2 += x[2];
        a3 += x[3];
        a0 += x[4];
        a1 += x[5];
        a2 += x[6];
        a3 += x[7];
      }
      for (; it + 24 < nit; it += 32) {
        const double x0 = (it < nslots) ? rp[int64_t(it) * n] : cp[int64_t(it) * n];
        const double x1 = (it + 8 < nslots) ? rp[int64_t(it + 8) * n] : cp[int64_t(it + 8) * n];
        const double x2 = (it + 16 < nslots) ? rp[int64_t(it + 16) * n] : cp[int64_t(it + 16) * n];
        const double x3 = (it + 24 < nslots) ? rp[int64_t(it + 24) * n] : cp[int64_t(it + 24) * n];
        a0 += x0;
        a1 += x1;
        a2 += x2;
        a3 += x3;
      }
      if (it < nit) {                              // up to three more items: independent loads again
        const double x0 = (it < nslots) ? rp[int64_t(it) * n] : cp[int64_t(it) * n];
        const double x1 = (it + 8 < nit) ? ((it + 8 < nslots) ? rp[int64_t(it + 8) * n] : cp[int64_t(it + 8) * n]) : 0.0;
        const double x2 = (it + 16 < nit) ? ((it + 16 < nslots) ? rp[int64_t(it + 16) * n] : cp[int64_t(it + 16) * n]) : 0.0;
        a0 += x0;
        a1 += x1;
        a2 += x2;
      }
      return (a0 + a1) + (a2 + a3);
    };
    if (wid == 1) {                     // u^T (A22 u): per-CTA partials, fixed order
      double p = 0.0;
      for (int q = lane; q < int(nb); q += 32) p += part2[q];
      p = warp_sum(p);
      if (lane == 0) uty_s = p;
    }
    double ybase_lane = 0.0;            // last warp: its share of (A22 u)[0] for W[c+1, i] below
    if (wid == kSymWarps - 1) ybase_lane = yraw_lane(base);
    if (wid == 0) {                     // (A22 u)[0], by the same lane split as the owner of row `base`
      double p = yraw_lane(base);
      p += __shfl_xor_sync(0xffffffffu, p, 4);
      p += __shfl_xor_sync(0xffffffffu, p, 2);
      p += __shfl_xor_sync(0xffffffffu, p, 1);
      if (lane == 0) yraw0_s = p;
    }
    {   // tmp1 = W^T v | tmp2 = V^T v: 8 threads per entry add the row-block partials (fixed tree)
      const int o = tid >> 3, part = tid & 7;
      const int t = o & (kTrdNb - 1);
      double t0s = 0.0, t1s = 0.0;
      if (t < i) {
        const double* wp = a.wvpart + o;
        for (int q0 = part; q0 < nrb; q0 += 64) {      // 8 independent loads per thread and round
          double x[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) x[q] = (q0 + 8 * q < nrb) ? wp[int64_t(q0 + 8 * q) * (2 * kTrdNb)] : 0.0;
#pragma unroll
          for (int q = 0; q < 8; q += 2) {
            t0s += x[q];
            t1s += x[q + 1];
          }
        }
      }
      double tsum = t0s + t1s;
      tsum += __shfl_xor_sync(0xffffffffu, tsum, 4);
      tsum += __shfl_xor_sync(0xffffffffu, tsum, 2);
      tsum += __shfl_xor_sync(0xffffffffu, tsum, 1);
      if (part == 0) {
        if (t < i) {
          const double first = o < kTrdNb ? W[base + t * ldw] : A[base + (j0 + t) * lda];
          tsum = fma(scl, tsum, fix * first);
        }
        tmps[o] = tsum;
      }
    }
    __syncthreads();
    TQ_PHASE(5)
    // acc -= V[r, :] tmp1, s1 -= W[r, :] tmp2 over this lane's share of the panel history
    auto history = [&](int64_t r, double& acc, double& s1) {
      const double* ar = A + r + (j0 + sub) * lda;
      const double* wr = W + r + int64_t(sub) * ldw;
      double x[8], y[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {            // i <= 64: at most 8 entries per lane, all loads independent
        const bool ok = sub + 8 * q < i;
        x[q] = ok ? ar[int64_t(8 * q) * lda] : 0.0;
        y[q] = ok ? wr[int64_t(8 * q) * ldw] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int t = sub + 8 * q;
        if (t < i) {
          acc = fma(-x[q], tmps[t], acc);
          s1 = fma(-y[q], tmps[kTrdNb + t], s1);
        }
      }
    };
    const double vtyv = scl * scl * uty_s + 2.0 * scl * fix * yraw0_s + fix * fix * a00;
    double cross = 0.0;
    for (int t = 0; t < i; ++t) cross = fma(tmps[t], tmps[kTrdNb + t], cross);
    const double wv = tau * (vtyv - 2.0 * cross);          // w'.v
    const double alpha2 = -0.5 * tau * wv;
    // (the last warp rarely owns rows: it starts on the next column's W[c+1, i] right away)
    if (wid >= kSymWarps - 3 && wid < kSymWarps - 1) {     // row c+1 of the panel history for the next phase A
      const int t = (wid - (kSymWarps - 3)) * 32 + lane;   // 64 lanes <-> t
      if (t < i) {
        crow[t] = W[(c + 1) + t * ldw];
        crow[kTrdNb + t] = A[(c + 1) + (j0 + t) * lda];
      } else if (t == i) {
        crow[kTrdNb + t] = 1.0;                             // V[c+1, i] is the unit entry; W[c+1, i] = wrow_s below
      }
    }
    if (wid == kSymWarps - 1) {        // W[c+1, i] for the next column update (v[c+1] = 1): every CTA repeats
      const int64_t r = c + 1;         // the owner's arithmetic (same lane split) so the value is bit-identical
      double acc = scl * ybase_lane, s1 = 0.0;
      if (sub == 0) acc = fma(fix, a00, acc);            // r == base: A[r + base * lda] is A22[0, 0]
      history(r, acc, s1);
      double ymw = acc + s1;
      ymw += __shfl_xor_sync(0xffffffffu, ymw, 4);
      ymw += __shfl_xor_sync(0xffffffffu, ymw, 2);
      ymw += __shfl_xor_sync(0xffffffffu, ymw, 1);
      if (lane == 0) {
        wrow_s = fma(alpha2, 1.0, tau * ymw);
        crow[i] = wrow_s;
      }
    }
    for (int64_t rb4 = 4 * gwarp; rb4 < n; rb4 += 4 * nwarps) {
      const int64_t r = rb4 + rsel;
      const bool act = (r < n) && (r >= c + 1);
      double acc = 0.0, s1 = 0.0, ucur = 0.0;
      if (act) {
        double a0v = 0.0;
        if (sub == 0) {                 // independent of the partial sums below: issue first
          ucur = A[r + c * lda];
          a0v = A[r + base * lda];
        }
        acc = scl * yraw_lane(r);
        acc = fma(fix, a0v, acc);
        history(r, acc, s1);
      }
      double ymw = acc + s1;                                  // y - V tmp1 - W tmp2
      ymw += __shfl_xor_sync(0xffffffffu, ymw, 4);
      ymw += __shfl_xor_sync(0xffffffffu, ymw, 2);
      ymw += __shfl_xor_sync(0xffffffffu, ymw, 1);
      if (act && sub == 0) {
        const double vr = (r == c + 1) ? 1.0 : scl * ucur;
        A[r + c * lda] = vr;
        W[r + int64_t(i) * ldw] = fma(alpha2, vr, tau * ymw);
        fence_proxy_async_global();      // these two entries are read by TMA (async proxy) in later columns
      }
    }
    __syncwarp();
    TQ_PHASE(6)
    if (gt == 0) {
      a.tau[c] = tau;
      a.e[c] = beta;
    }
    __syncthreads();
    TQ_PHASE(4)
  }
#undef TQ_PHASE
}

// X = [V | W], Y = [W | V]  (s x 2jb each, leading dimension ld)
__global__ void pack_vw_kernel(const double* __restrict__ V, int64_t ldv, const double* __restrict__ W, int64_t ldw,
                               int64_t s, int jb, double* __restrict__ X, double* __restrict__ Y, int64_t ld) {
  const int t = blockIdx.y;
  for (int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < s; r += int64_t(gridDim.x) * blockDim.x) {
    const double v = V[r + int64_t(t) * ldv], w = W[r + int64_t(t) * ldw];
    X[r + int64_t(t) * ld] = v;
    X[r + int64_t(jb + t) * ld] = w;
    Y[r + int64_t(t) * ld] = w;
    Y[r + int64_t(jb + t) * ld] = v;
  }
}

// Reduces A (n x n, symmetric, both triangles valid) to tridiagonal form.  On exit
// d[0:n], e[0:n-1], tau[0:n-1]; reflector c lives in A[c+1:, c] with an explicit unit at
// A[c+1, c].
int make_tmap_f64(CUtensorMap* tmap, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                  uint32_t box_rows, uint32_t box_cols);

static thread_local bool g_sytrd_notified = false;   // the stage callback already ran inside sytrd_lower

struct SymBuffers {
  double* rowpart;
  double* colpart;
  double* wvpart;
};

// largest number of row-partial slots any row block needs for trailing size len with T tiles on G CTAs
static int sym_max_slots(int64_t len, int dl, bool history, int G) {   // dl = (c + 1) % align
  const int nrb = int((len + dl + kTileR - 1) / kTileR), nstrips = int((len + kTileC - 1) / kTileC);
  if (nrb <= 0) return 0;
  const int TA = nrb * (nrb - 1) + std::min(2 * nrb, nstrips);
  const int T = TA + (history ? 2 * nrb : 0);
  const int ch = std::max(1, (T + G - 1) / G);
  int worst = 0;
  for (int rb = 0; rb < nrb; ++rb) {
    const int64_t ts = int64_t(rb) * (rb + 1);
    const int ncs = std::min(2 * rb + 2, nstrips);
    worst = std::max(worst, int((ts + ncs - 1) / ch - ts / ch + 1));
  }
  return worst;
}

static int sytrd_lower(cublasHandle_t h, cudaStream_t st, double* A, int64_t n, double* d, double* e, double* tau,
                       double* W, double* y /*warps x n*/, double* tmp /*warps*2*kTrdNb*/, double* part /*2*1024*/,
                       double* scal /*16*/, unsigned int* bar, double* XY /*2 x (n x 2 kTrdNb)*/,
                       const SymBuffers& sb) {
  const int64_t lda = n, ldw = n;
  const double one = 1.0, mone = -1.0;
  TQ_CUDA_CHECK(cudaMemsetAsync(tau, 0, sizeof(double) * n, st));
  TQ_CUDA_CHECK(cudaMemsetAsync(scal, 0, sizeof(double) * 16, st));
  TQ_CUDA_CHECK(cudaMemsetAsync(e, 0, sizeof(double) * n, st));
  static thread_local int coop_per_sm_dev[kMaxDevices] = {};
  int& coop_per_sm = coop_per_sm_dev[device_slot()];
  if (!coop_per_sm) {
    int per_sm = 0;
    TQ_CUDA_CHECK(cudaFuncSetAttribute(sytrd_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       int(kPanelSmem)));
    TQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sytrd_panel_kernel, kPanelThreads,
                                                                kPanelSmem));
    if (per_sm < 1) {
      set_error("sytrd: panel kernel cannot be made resident");
      return TQ_ERR_CUDA;
    }
    coop_per_sm = per_sm > 4 ? 4 : per_sm;
  }
  const int coop_blocks = num_sms() * coop_per_sm;      // num_sms() honours tq_set_sm_budget
  // symmetric TMA panel (half the DRAM bytes) whenever the tensor map can be built: even n (16-byte
  // row pitch); the column-dot panel stays for odd n.
  int sym_ok = 0;
  {
    int per_sm = 0;
    TQ_CUDA_CHECK(cudaFuncSetAttribute(sytrd_panel_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       int(kSymSmem)));
    TQ_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sytrd_panel_sym_kernel, kSymThreads,
                                                                kSymSmem));
    if (per_sm >= 1) sym_ok = 1;
  }
  const int sym_blocks = sym_ok ? num_sms() : 0;
  const bool use_sym = sym_blocks > 0 && (n % 2 == 0) && n >= 256 && (reinterpret_cast<uintptr_t>(A) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(W) % 16 == 0);
  CUtensorMap tmA, tmW;
  const int boxc = 64;
  const int align = (n % 32 == 0 && reinterpret_cast<uintptr_t>(A) % 256 == 0 && reinterpret_cast<uintptr_t>(W) % 256 == 0) ? 32 : 2;
  if (use_sym) {
    TQ_TRY(make_tmap_f64(&tmA, A, uint64_t(n), uint64_t(n), uint64_t(n), kTileR, boxc));
    TQ_TRY(make_tmap_f64(&tmW, W, uint64_t(n), uint64_t(kTrdNb), uint64_t(n), kTileR, boxc));
    // W is read through TMA before every column of it has been written: keep stale NaNs out of it
    TQ_CUDA_CHECK(cudaMemsetAsync(W, 0, sizeof(double) * size_t(n) * kTrdNb, st));
  }
  // (Reporting TQ_STAGE_SYTRD_DONE early, once the trailing matrix has shrunk to 4096 - 6144 rows, was measured
  // with bench.py and is a loss - 1.83 s per layer vs 1.58 s: the narrow solves' cooperative panels then
  // gang-schedule against the wide solve's remaining panel launches.  The callback fires at the end.)
  g_sytrd_notified = false;
  for (int64_t j0 = 0; j0 < n; j0 += kTrdNb) {
    const int jb = int(imin(kTrdNb, n - j0));
    if (use_sym) {
      for (int i = 0; i < jb; ++i) {
        if (sym_max_slots(n - (j0 + i) - 1, int((j0 + i + 1) % align), i > 0, sym_blocks) > kRowSlots) {
          set_error("sytrd: row-partial slots exceed %d (n=%lld)", kRowSlots, (long long)n);
          return TQ_ERR_UNSUPPORTED;
        }
      }
      TQ_CUDA_CHECK(cudaMemsetAsync(bar, 0, sizeof(unsigned int), st));
      SymPanelArgs pa{A, n, j0, jb, W, d, e, tau, sb.rowpart, sb.colpart, sb.wvpart, part, scal, bar,
                      trace_enabled() ? 1 : 0, boxc, align};
      void* kargs[] = {&pa, &tmA, &tmW};
      double bytes = 0.0;      // algorithmic bytes: every column streams the LOWER triangle of the trailing matrix once
      for (int i = 0; i < jb; ++i) {
        const double len = double(n - (j0 + i) - 1);
        bytes += len * (0.5 * len + 2.0 * i) * 8.0;
      }
      const int pslot = prof_begin_launch(st, bytes, TQ_PROF_SYTRD_SYM);
      TQ_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)sytrd_panel_sym_kernel, dim3(sym_blocks), dim3(kSymThreads),
                                                kargs, kSymSmem, st));
      prof_end_launch(st, pslot);
      ++g_launch_count;
    } else {
      TQ_CUDA_CHECK(cudaMemsetAsync(bar, 0, sizeof(unsigned int), st));
      TrdPanelArgs pa{A, n, j0, jb, W, d, e, tau, y, tmp, part, scal, bar, trace_enabled() ? 1 : 0};
      void* kargs[] = {&pa};
      double bytes = 0.0;      // algorithmic bytes of the panel: every column streams the trailing matrix once
      for (int i = 0; i < jb; ++i) {
        const double len = double(n - (j0 + i) - 1);
        bytes += len * (len + 2.0 * i) * 8.0;
      }
      const int pslot = prof_begin_launch(st, bytes, TQ_PROF_SYTRD_COLDOT);
      TQ_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)sytrd_panel_kernel, dim3(coop_blocks), dim3(kPanelThreads),
                                                kargs, kPanelSmem, st));
      prof_end_launch(st, pslot);
      ++g_launch_count;
    }
    const int64_t r0 = j0 + jb;
    const int64_t s2 = n - r0;
    if (s2 > 0) {
      if (use_sym) {
        // A22 -= V2 W2^T + W2 V2^T on the LOWER triangle only (DSYR2K, half the flops): the symmetric panel
        // kernel, phases A / C and ormtr never read above the diagonal.  844 ms vs 870 at n = 12288.
        TQ_CUBLAS_CHECK(cublasDsyr2k(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, int(s2), jb, &mone, A + r0 + j0 * lda,
                                     int(lda), W + r0, int(ldw), &one, A + r0 + r0 * lda, int(lda)));
      } else {
        // full square (the column-dot panel reads both triangles) as ONE rank-2jb DGEMM [V2 W2] [W2 V2]^T:
        // cuBLAS runs a rank-128 update at 30 TF/s, two rank-64 updates at 20.7 (profiles/r01_dgemm_probe.log)
        dim3 grid((unsigned)imin(ceil_div(s2, 256), 64), (unsigned)jb);
        pack_vw_kernel<<<grid, 256, 0, st>>>(A + r0 + j0 * lda, lda, W + r0, ldw, s2, jb, XY,
                                             XY + size_t(n) * 2 * kTrdNb, n);
        TQ_LAUNCH_CHECK();
        TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, int(s2), int(s2), 2 * jb, &mone, XY, int(n),
                                    XY + size_t(n) * 2 * kTrdNb, int(n), &one, A + r0 + r0 * lda, int(lda)));
      }
    }
  }
  return TQ_OK;
}

// ======================================================================= D&C kernels
// rank-one tears at every cut point of the D&C tree (LAPACK DLAED0): serial, ncuts < n/16
__global__ void dc_tear_serial_kernel(double* d, const double* e, const int* cuts, int ncuts) {
  for (int t = 0; t < ncuts; ++t) {
    int cp = cuts[t];
    double b = fabs(e[cp]);
    d[cp] -= b;
    d[cp + 1] -= b;
  }
}

// QL implicit (EISPACK tql2 / "tqli") on leaves of size <= kLeaf: one CTA of kLeaf threads per leaf.
// Thread 0 runs the scalar recurrence and records the plane rotations of one sweep; every
// thread then applies them to its own row of the leaf's eigenvector block (shared memory).
constexpr size_t kLeafSmem = size_t(kLeaf) * (kLeaf + 1) * sizeof(double);
__global__ void __launch_bounds__(kLeaf)
dc_leaf_kernel(double* __restrict__ d, const double* __restrict__ e, double* __restrict__ Z, int64_t ldz,
               const int* __restrict__ leaf_off, const int* __restrict__ leaf_len, int* __restrict__ fail) {
  extern __shared__ double zz_raw[];                         // zz[row][col], row stride kLeaf + 1
  __shared__ double sd[kLeaf], se[kLeaf], rc[kLeaf], rs[kLeaf];
  __shared__ int ctl[4];   // 0: first rotation index (high), 1: last rotation index (low), 2: done flag
  __shared__ int order[kLeaf];
  const int lane = threadIdx.x;
  double* zrow = zz_raw + size_t(lane) * (kLeaf + 1);
  const int off = leaf_off[blockIdx.x], len = leaf_len[blockIdx.x];
  if (lane < len) {
    sd[lane] = d[off + lane];
    se[lane] = (lane < len - 1) ? e[off + lane] : 0.0;
    for (int c = 0; c < len; ++c) zrow[c] = (lane == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  const double eps = 1.1102230246251565e-16;
  double tst1 = 0.0;         // thread 0: EISPACK tql2's running max of |d| + |e| - the scale of the leaf seen so far
  for (int l = 0; l < len; ++l) {
    int iter = 0;
    if (lane == 0) tst1 = fmax(tst1, fabs(sd[l]) + fabs(se[l]));
    while (true) {
      if (lane == 0) {
        int m = l;
        for (; m < len - 1; ++m) {
          // negligible against its neighbours (keeps the relative accuracy of a graded leaf) or against the leaf's
          // scale (tql2's test: a cluster of zero eigenvalues - a rank-deficient H - never passes the first one)
          double dd = fabs(sd[m]) + fabs(sd[m + 1]);
          if (fabs(se[m]) <= eps * dd || fabs(se[m]) <= eps * tst1) break;
        }
        if (m == l) {
          ctl[2] = 1;
        } else if (iter++ >= 60) {
          ctl[2] = 1;
          atomicExch(fail, 1);
        } else {
          ctl[2] = 0;
          double g = (sd[l + 1] - sd[l]) / (2.0 * se[l]);
          double r = hypot(g, 1.0);
          g = sd[m] - sd[l] + se[l] / (g + copysign(r, g));
          double s = 1.0, c = 1.0, p = 0.0;
          int i = m - 1;
          bool early = false;
          for (; i >= l; --i) {
            double f = s * se[i], b = c * se[i];
            r = hypot(f, g);
            se[i + 1] = r;
            if (r == 0.0) {
              sd[i + 1] -= p;
              se[m] = 0.0;
              early = true;
              break;
            }
            s = f / r;
            c = g / r;
            g = sd[i + 1] - p;
            r = (sd[i] - g) * s + 2.0 * c * b;
            p = s * r;
            sd[i + 1] = g + p;
            g = c * r - b;
            rc[i] = c;
            rs[i] = s;
          }
          ctl[0] = m - 1;
          ctl[1] = early ? i + 1 : l;
          if (!early) {
            sd[l] -= p;
            se[l] = g;
            se[m] = 0.0;
          }
        }
      }
      __syncthreads();
      const bool done = ctl[2] != 0;
      const int ihi = ctl[0], ilo = ctl[1];
      if (!done && lane < len) {
        for (int i = ihi; i >= ilo; --i) {
          double f = zrow[i + 1];
          zrow[i + 1] = rs[i] * zrow[i] + rc[i] * f;
          zrow[i] = rc[i] * zrow[i] - rs[i] * f;
        }
      }
      __syncthreads();        // thread 0 overwrites ctl / rc / rs in the next sweep
      if (done) break;
    }
  }
  // ascending order (selection sort on thread 0)
  if (lane == 0) {
    for (int i = 0; i < len; ++i) order[i] = i;
    for (int i = 0; i < len - 1; ++i) {
      int kmin = i;
      for (int j = i + 1; j < len; ++j)
        if (sd[order[j]] < sd[order[kmin]]) kmin = j;
      int t = order[i];
      order[i] = order[kmin];
      order[kmin] = t;
    }
  }
  __syncthreads();
  if (lane < len) {
    d[off + lane] = sd[order[lane]];
    for (int c = 0; c < len; ++c) Z[(off + lane) + int64_t(off + c) * ldz] = zrow[order[c]];
  }
}

// z = [last row of Q1, sign * first row of Q2] / sqrt(2)
__global__ void dc_build_z_kernel(const double* __restrict__ Zb, int64_t ldz, int n1, int len, double sign,
                                  double* __restrict__ z) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  const double r = 0.70710678118654752440;
  z[j] = (j < n1) ? Zb[(n1 - 1) + int64_t(j) * ldz] * r : sign * Zb[n1 + int64_t(j) * ldz] * r;
}

struct DcRot {
  int pj, nj;
  double c, s;
};

// columns (pj, nj) of the block: x' = c x + s y, y' = c y - s x, in list order; thread per row
__global__ void dc_apply_rot_kernel(double* __restrict__ Zb, int64_t ldz, int len, const DcRot* __restrict__ rot,
                                    int nrot) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= len) return;
  for (int q = 0; q < nrot; ++q) {
    DcRot R = rot[q];
    double x = Zb[r + int64_t(R.pj) * ldz], y = Zb[r + int64_t(R.nj) * ldz];
    Zb[r + int64_t(R.pj) * ldz] = R.c * x + R.s * y;
    Zb[r + int64_t(R.nj) * ldz] = R.c * y - R.s * x;
  }
}

// dst[:, p] = src[:, idx[p]]  (len rows)
__global__ void dc_gather_kernel(const double* __restrict__ src, int64_t lds, int len, const int* __restrict__ idx,
                                 int ncols, double* __restrict__ dst, int64_t ldd) {
  int p = blockIdx.y;
  if (p >= ncols) return;
  const double* s = src + int64_t(idx[p]) * lds;
  double* t = dst + int64_t(p) * ldd;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < len; r += gridDim.x * blockDim.x) t[r] = s[r];
}

// Secular equation 1 + sum_i z2_i / (dl_i - lam) = 0 (z2 = rho z^2, dl ascending, distinct):
// one warp per root j in (dl_j, dl_{j+1}).  Output: origin pole org[j] and offset tau[j]
// (lam_j = dl[org_j] + tau_j, with dl_i - lam_j = (dl_i - dl[org_j]) - tau_j computed
// without cancellation).
__global__ void __launch_bounds__(256)
dc_secular_kernel(const double* __restrict__ dl, const double* __restrict__ z2, int K, int* __restrict__ org,
                  double* __restrict__ tau, double* __restrict__ lam) {
  const int lane = threadIdx.x & 31;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= K) return;
  const double eps = 1.1102230246251565e-16;
  int o;
  double lo, hi;
  if (j < K - 1) {
    const double gap = dl[j + 1] - dl[j];
    double f = 0.0;
    for (int i = lane; i < K; i += 32) f += z2[i] / ((dl[i] - dl[j]) - 0.5 * gap);
    f = 1.0 + warp_sum(f);
    if (f > 0.0) {
      o = j;
      lo = 0.0;
      hi = 0.5 * gap;
    } else {
      o = j + 1;
      lo = -0.5 * gap;
      hi = 0.0;
    }
  } else {
    double s = 0.0;
    for (int i = lane; i < K; i += 32) s += z2[i];
    s = warp_sum(s);
    o = j;
    lo = 0.0;
    hi = s;
  }
  const double dorg = dl[o];
  double t = 0.5 * (lo + hi);
  for (int it = 0; it < 200; ++it) {
    double psi = 0, phi = 0, dpsi = 0, dphi = 0, sabs = 0;
    for (int i = lane; i < K; i += 32) {
      const double D = (dl[i] - dorg) - t;
      const double term = z2[i] / D;
      const double dterm = term / D;
      if (i <= j) {
        psi += term;
        dpsi += dterm;
      } else {
        phi += term;
        dphi += dterm;
      }
      sabs += fabs(term);
    }
    psi = warp_sum(psi);
    phi = warp_sum(phi);
    dpsi = warp_sum(dpsi);
    dphi = warp_sum(dphi);
    sabs = warp_sum(sabs);
    const double g = 1.0 + psi + phi;
    if (fabs(g) <= 8.0 * eps * (1.0 + sabs)) break;
    if (g < 0.0) lo = t;
    else hi = t;
    if (hi - lo <= 4.0 * eps * fmax(fabs(lo), fabs(hi))) {
      t = 0.5 * (lo + hi);
      break;
    }
    const double Dj = (dl[j] - dorg) - t;
    const double S = Dj * Dj * dpsi, s0 = psi - Dj * dpsi;
    double cand0 = NAN, cand1 = NAN;
    if (j < K - 1) {
      const double Dj1 = (dl[j + 1] - dorg) - t;
      const double Rr = Dj1 * Dj1 * dphi, r0 = phi - Dj1 * dphi;
      const double c = 1.0 + s0 + r0;
      const double a2 = c, a1 = -(c * (Dj + Dj1) + S + Rr), a0 = c * Dj * Dj1 + S * Dj1 + Rr * Dj;
      if (a2 == 0.0) {
        if (a1 != 0.0) cand0 = -a0 / a1;
      } else {
        const double disc = a1 * a1 - 4.0 * a2 * a0;
        if (disc >= 0.0) {
          const double q = -0.5 * (a1 + copysign(sqrt(disc), a1));
          if (q != 0.0) cand0 = a0 / q;
          cand1 = q / a2;
        }
      }
    } else {
      const double c = 1.0 + s0;
      if (c != 0.0) cand0 = Dj + S / c;
    }
    double tn = 0.5 * (lo + hi);
    const double x0 = t + cand0, x1 = t + cand1;
    if (x0 > lo && x0 < hi) tn = x0;
    else if (x1 > lo && x1 < hi) tn = x1;
    t = tn;
  }
  if (lane == 0) {
    org[j] = o;
    tau[j] = t;
    lam[j] = dorg + t;
  }
}

// Gu/Eisenstat: zhat_i = sign(z_i) sqrt( -prod_j (dl_i - lam_j) / prod_{j != i} (dl_i - dl_j) ); warp per i
__global__ void __launch_bounds__(256)
dc_zhat_kernel(const double* __restrict__ dl, const double* __restrict__ zl, const int* __restrict__ org,
               const double* __restrict__ tau, int K, double* __restrict__ zhat) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= K) return;
  const double di = dl[i];
  double p = 1.0;
  for (int j = lane; j < K; j += 32) {
    const double num = (di - dl[org[j]]) - tau[j];
    p *= (j == i) ? num : num / (di - dl[j]);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, o);
  if (lane == 0) zhat[i] = copysign(sqrt(fabs(p)), zl[i]);
}

// U[rowpos[i], j] = zhat_i / (dl_i - lam_j), column-normalised; CTA per column j
__global__ void __launch_bounds__(256)
dc_u_kernel(const double* __restrict__ dl, const double* __restrict__ zhat, const int* __restrict__ org,
            const double* __restrict__ tau, const int* __restrict__ rowpos, int K, double* __restrict__ U,
            int64_t ldu, int j0) {
  __shared__ double sh[32];
  const int j = j0 + blockIdx.x;
  const double dorg = dl[org[j]], t = tau[j];
  double ss = 0.0;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const double u = zhat[i] / ((dl[i] - dorg) - t);
    ss = fma(u, u, ss);
  }
  ss = block_sum(ss, sh);
  const double inv = 1.0 / sqrt(ss);
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const double u = zhat[i] / ((dl[i] - dorg) - t);
    U[rowpos[i] + int64_t(j) * ldu] = u * inv;
  }
}

// Merge the ascending lists lam[0:K) and dd[0:nd) into ascending order: src[rank] = source
// column (j for a secular root, K + t for a deflated value), dout[rank] = value.
__global__ void dc_merge_order_kernel(const double* __restrict__ lam, int K, const double* __restrict__ dd, int nd,
                                      int* __restrict__ src, double* __restrict__ dout) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= K + nd) return;
  int rank;
  double v;
  if (t < K) {
    v = lam[t];
    int lo = 0, hi = nd;          // # deflated values strictly below v
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (dd[mid] < v) lo = mid + 1;
      else hi = mid;
    }
    rank = t + lo;
  } else {
    v = dd[t - K];
    int lo = 0, hi = K;           // # roots <= v
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (lam[mid] <= v) lo = mid + 1;
      else hi = mid;
    }
    rank = (t - K) + lo;
  }
  src[rank] = t;
  dout[rank] = v;
}

// Zb[:, rank] = (src < K ? Zo[:, src] : Zg[:, src])
__global__ void dc_scatter_kernel(const double* __restrict__ Zo, const double* __restrict__ Zg, int64_t lds,
                                  int len, int K, const int* __restrict__ src, double* __restrict__ Zb,
                                  int64_t ldz, int p0) {
  int p = p0 + blockIdx.y;
  int s = src[p];
  const double* from = (s < K ? Zo : Zg) + int64_t(s) * lds;
  double* to = Zb + int64_t(p) * ldz;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < len; r += gridDim.x * blockDim.x) to[r] = from[r];
}

// ======================================================================= D&C driver
struct DcBuffers {
  double *Zg, *Zo, *U;            // n x n each
  double *z, *dl, *z2, *zl, *tau, *lam, *zhat, *dd, *dout;   // n each
  int *org, *rowpos, *gidx, *src;                             // n each
  DcRot* rot;                                                 // n
  // pinned host staging
  double *h_d, *h_z, *h_dl, *h_z2, *h_zl, *h_dd;
  int *h_rowpos, *h_gidx;
  DcRot* h_rot;
};

struct DcNode {
  int off, len;
};

static void dc_collect(int off, int len, std::vector<DcNode>& leaves, std::vector<int>& cuts,
                       std::vector<DcNode>& merges /*post-order*/, std::vector<int>& merge_n1) {
  if (len <= kLeaf) {
    leaves.push_back({off, len});
    return;
  }
  int n1 = len / 2;
  cuts.push_back(off + n1 - 1);
  dc_collect(off, n1, leaves, cuts, merges, merge_n1);
  dc_collect(off + n1, len - n1, leaves, cuts, merges, merge_n1);
  merges.push_back({off, len});
  merge_n1.push_back(n1);
}

// One rank-one merge, split in two so that a whole LEVEL of the tree shares one host round trip:
//   dc_merge_host    deflation (LAPACK DLAED2 logic) on the host copies of (d, z) of the merge's index range
//                    [off, off + len); fills the pinned staging arrays at the same offsets;
//   dc_merge_device  rotations, gather, secular solve, eigenvector update (DGEMM), reorder - launches only.
// Round 1 did both per merge with a D2H + synchronise + up to seven H2D copies each (511 merges at n = 12288 with
// leaves of 32): 90 ms alone and 200 ms next to three other solves, whose host threads compete for the driver.
struct DcMergePlan {
  int off, n1, len;
  int K, K1, K2, K3, nd, nrot;
  size_t scratch;          // element offset of this merge's len x len blocks in Zg / Zo / U
};

static int dc_merge_host(DcMergePlan& mp, double beta, DcBuffers& B) {
  const int off = mp.off, n1 = mp.n1, len = mp.len;
  const double eps = 1.1102230246251565e-16;
  double* hd = B.h_d + off;
  double* hz = B.h_z + off;
  const double rho = 2.0 * std::fabs(beta);
  double dmax = 0, zmax = 0;
  for (int j = 0; j < len; ++j) {
    dmax = std::max(dmax, std::fabs(hd[j]));
    zmax = std::max(zmax, std::fabs(hz[j]));
  }
  const double tol = 8.0 * eps * std::max(dmax, zmax);
  std::vector<int> order(len);
  for (int j = 0; j < len; ++j) order[j] = j;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return hd[a] < hd[b]; });
  std::vector<int> nondef, defl, ctype(len);
  for (int j = 0; j < len; ++j) ctype[j] = j < n1 ? 1 : 3;
  int nrot = 0;
  DcRot* hrot = B.h_rot + off;
  if (rho * zmax <= tol) {
    defl = order;
  } else {
    int pj = -1;
    for (int q = 0; q < len; ++q) {
      const int jj = order[q];
      if (rho * std::fabs(hz[jj]) <= tol) {
        defl.push_back(jj);
        continue;
      }
      if (pj < 0) {
        pj = jj;
        continue;
      }
      const int nj = jj;
      double s = hz[pj], c = hz[nj];
      const double tau_ = std::hypot(c, s);
      const double t = hd[nj] - hd[pj];
      c /= tau_;
      s = -s / tau_;
      if (std::fabs(t * c * s) <= tol) {
        hz[nj] = tau_;
        hz[pj] = 0.0;
        if (ctype[nj] != ctype[pj]) ctype[nj] = 2;
        hrot[nrot++] = DcRot{pj, nj, c, s};
        const double tt = hd[pj] * c * c + hd[nj] * s * s;
        hd[nj] = hd[pj] * s * s + hd[nj] * c * c;
        hd[pj] = tt;
        defl.push_back(pj);
        pj = nj;
      } else {
        nondef.push_back(pj);
        pj = nj;
      }
    }
    if (pj >= 0) nondef.push_back(pj);
  }
  const int K = int(nondef.size());
  const int nd = len - K;
  // deflated values ascending (rotations may perturb the order slightly)
  std::stable_sort(defl.begin(), defl.end(), [&](int a, int b) { return hd[a] < hd[b]; });
  // secular poles must be strictly increasing; DLAED2's rotation keeps them ordered, but guard anyway
  for (int q = 1; q < K; ++q) {
    if (!(hd[nondef[q]] > hd[nondef[q - 1]])) {
      set_error("stedc: non-increasing poles after deflation (merge off=%d len=%d)", off, len);
      return TQ_ERR_NOCONV;
    }
  }
  // group the non-deflated columns by type [1 | 2 | 3] for the two half GEMMs
  int K1 = 0, K2 = 0, K3 = 0;
  for (int q = 0; q < K; ++q) {
    int ty = ctype[nondef[q]];
    K1 += ty == 1;
    K2 += ty == 2;
    K3 += ty == 3;
  }
  int p1 = 0, p2 = K1, p3 = K1 + K2;
  for (int q = 0; q < K; ++q) {
    const int jj = nondef[q];
    const int ty = ctype[jj];
    const int pos = ty == 1 ? p1++ : (ty == 2 ? p2++ : p3++);
    B.h_rowpos[off + q] = pos;
    B.h_gidx[off + pos] = jj;
    B.h_dl[off + q] = hd[jj];
    B.h_zl[off + q] = hz[jj];
    B.h_z2[off + q] = rho * hz[jj] * hz[jj];
  }
  for (int t = 0; t < nd; ++t) {
    B.h_gidx[off + K + t] = defl[t];
    B.h_dd[off + t] = hd[defl[t]];
  }
  mp.K = K;
  mp.K1 = K1;
  mp.K2 = K2;
  mp.K3 = K3;
  mp.nd = nd;
  mp.nrot = nrot;
  return TQ_OK;
}

// `want` (root merge only): the caller's column chooser.  The eigenvalues of the whole matrix are known once the
// root's secular equation is solved, BEFORE its eigenvector update - the largest GEMM of the decomposition
// (n K^2 flop, three quarters of all merge flops).  The chooser is asked there, and only the columns it wants
// (the t = n - k dropped directions in the solver's usual regime, ~0.1 n) are formed, scattered and returned.
struct DcWanted {
  const EighColumnChooser* choose = nullptr;
  int64_t col0 = 0, ncols = 0;
  bool asked = false;
};

static int dc_merge_device(cublasHandle_t h, cudaStream_t st, double* d, double* Z, int64_t n, const DcMergePlan& mp,
                           DcBuffers& B, DcWanted* want = nullptr) {
  const int64_t ldz = n;
  const int off = mp.off, n1 = mp.n1, len = mp.len, K = mp.K, K1 = mp.K1, K2 = mp.K2, K3 = mp.K3, nd = mp.nd;
  double* Zb = Z + off + int64_t(off) * ldz;
  double* Zg = B.Zg + mp.scratch;
  double* Zo = B.Zo + mp.scratch;
  double* U = B.U + mp.scratch;
  if (mp.nrot > 0) {
    dc_apply_rot_kernel<<<(unsigned)ceil_div(len, 128), 128, 0, st>>>(Zb, ldz, len, B.rot + off, mp.nrot);
    TQ_LAUNCH_CHECK();
  }
  const int64_t lds = len;   // scratch matrices are packed len x len
  {
    dim3 grid((unsigned)imin(ceil_div(len, 256), 32), (unsigned)len);
    dc_gather_kernel<<<grid, 256, 0, st>>>(Zb, ldz, len, B.gidx + off, len, Zg, lds);
    TQ_LAUNCH_CHECK();
  }
  const unsigned wgrid = (unsigned)ceil_div(int64_t(K) * 32, 256);
  if (K > 0) {
    dc_secular_kernel<<<wgrid, 256, 0, st>>>(B.dl + off, B.z2 + off, K, B.org + off, B.tau + off, B.lam + off);
    TQ_LAUNCH_CHECK();
    dc_zhat_kernel<<<wgrid, 256, 0, st>>>(B.dl + off, B.zl + off, B.org + off, B.tau + off, K, B.zhat + off);
    TQ_LAUNCH_CHECK();
  }
  dc_merge_order_kernel<<<(unsigned)ceil_div(len, 256), 256, 0, st>>>(B.lam + off, K, B.dd + off, nd, B.src + off,
                                                                      B.dout + off);
  TQ_LAUNCH_CHECK();
  TQ_CUDA_CHECK(cudaMemcpyAsync(d + off, B.dout + off, sizeof(double) * len, cudaMemcpyDeviceToDevice, st));
  int p0 = 0, np = len;      // sorted positions to deliver
  int q0 = 0, q1 = K;        // secular roots (columns of U) among them
  if (want && want->choose && *want->choose && off == 0 && len == n) {
    TQ_TRY((*want->choose)(d, &want->col0, &want->ncols));
    want->asked = true;
    if (want->col0 < 0 || want->ncols < 0 || want->col0 + want->ncols > n) {
      set_error("eigh: column chooser returned [%lld, +%lld) outside [0, %lld)", (long long)want->col0,
                (long long)want->ncols, (long long)n);
      return TQ_ERR_INVALID;
    }
    p0 = int(want->col0);
    np = int(want->ncols);
    if (np < len) {
      // roots are ascending in their index and in their sorted position: the wanted ones are one contiguous range
      std::vector<int> hsrc(np > 0 ? np : 1);
      if (np > 0) {
        TQ_CUDA_CHECK(cudaMemcpyAsync(hsrc.data(), B.src + off + p0, sizeof(int) * np, cudaMemcpyDeviceToHost, st));
        TQ_CUDA_CHECK(cudaStreamSynchronize(st));
      }
      q0 = K;
      q1 = 0;
      for (int i = 0; i < np; ++i)
        if (hsrc[i] < K) {
          q0 = hsrc[i] < q0 ? hsrc[i] : q0;
          q1 = hsrc[i] + 1 > q1 ? hsrc[i] + 1 : q1;
        }
      if (q1 < q0) q0 = q1 = 0;
    }
  }
  const int nq = q1 - q0;
  if (nq > 0) {
    dc_u_kernel<<<nq, 256, 0, st>>>(B.dl + off, B.zhat + off, B.org + off, B.tau + off, B.rowpos + off, K, U, K, q0);
    TQ_LAUNCH_CHECK();
    const double one = 1.0, zero = 0.0;
    const int n2 = len - n1;
    const int K12 = K1 + K2, K23 = K2 + K3;
    const double* Uq = U + int64_t(q0) * K;
    double* Zq = Zo + int64_t(q0) * lds;
    if (K12 > 0)
      TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n1, nq, K12, &one, Zg, int(lds), Uq, K, &zero, Zq,
                                  int(lds)));
    else
      TQ_CUDA_CHECK(cudaMemset2DAsync(Zq, sizeof(double) * lds, 0, sizeof(double) * n1, nq, st));
    if (K23 > 0)
      TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n2, nq, K23, &one, Zg + n1 + int64_t(K1) * lds,
                                  int(lds), Uq + K1, K, &zero, Zq + n1, int(lds)));
    else
      TQ_CUDA_CHECK(cudaMemset2DAsync(Zq + n1, sizeof(double) * lds, 0, sizeof(double) * n2, nq, st));
  }
  if (np > 0) {
    dim3 grid((unsigned)imin(ceil_div(len, 256), 32), (unsigned)np);
    dc_scatter_kernel<<<grid, 256, 0, st>>>(Zo, Zg, lds, len, K, B.src + off, Zb, ldz, p0);
    TQ_LAUNCH_CHECK();
  }
  return TQ_OK;
}

// Eigen-decomposition of the symmetric tridiagonal (d, e): d <- eigenvalues ascending,
// Z (n x n) <- eigenvectors (columns).  e is read only.
static int stedc(cublasHandle_t h, cudaStream_t st, double* d, const double* e, double* Z, int64_t n,
                 Workspace& ws, DcWanted* want = nullptr) {
  DcBuffers B;
  B.Zg = ws.take<double>(size_t(n) * n);
  B.Zo = ws.take<double>(size_t(n) * n);
  B.U = ws.take<double>(size_t(n) * n);
  double** dv[] = {&B.z, &B.dl, &B.z2, &B.zl, &B.tau, &B.lam, &B.zhat, &B.dd, &B.dout};
  for (auto p : dv) *p = ws.take<double>(n);
  int** iv[] = {&B.org, &B.rowpos, &B.gidx, &B.src};
  for (auto p : iv) *p = ws.take<int>(n);
  B.rot = ws.take<DcRot>(n);
  int* d_leaf_off = ws.take<int>(n);
  int* d_leaf_len = ws.take<int>(n);
  int* d_cuts = ws.take<int>(n);
  int* d_fail = ws.take<int>(1);
  if (ws.overflow) {
    set_error("stedc: workspace too small");
    return TQ_ERR_WORKSPACE;
  }
  // pinned host staging: one grow-only buffer per host thread (cudaMallocHost / cudaFreeHost per
  // call cost 0.1 - 0.9 s at random in the end-to-end runs)
  const size_t hbytes = sizeof(double) * n * 7 + sizeof(int) * n * 2 + sizeof(DcRot) * n + 4096;
  static thread_local char* pinned = nullptr;
  static thread_local size_t pinned_size = 0;
  if (pinned_size < hbytes) {
    if (pinned) cudaFreeHost(pinned);
    pinned = nullptr;
    pinned_size = 0;
    TQ_CUDA_CHECK(cudaMallocHost(&pinned, hbytes));
    pinned_size = hbytes;
  }
  char* hp = pinned;
  auto htake = [&](size_t bytes) {
    char* r = hp;
    hp += (bytes + 15) / 16 * 16;
    return r;
  };
  B.h_d = (double*)htake(sizeof(double) * n);
  B.h_z = (double*)htake(sizeof(double) * n);
  B.h_dl = (double*)htake(sizeof(double) * n);
  B.h_z2 = (double*)htake(sizeof(double) * n);
  B.h_zl = (double*)htake(sizeof(double) * n);
  B.h_dd = (double*)htake(sizeof(double) * n);
  B.h_rowpos = (int*)htake(sizeof(int) * n);
  B.h_gidx = (int*)htake(sizeof(int) * n);
  B.h_rot = (DcRot*)htake(sizeof(DcRot) * n);
  double* he = (double*)htake(sizeof(double) * n);
  TQ_CUDA_CHECK(cudaMemcpyAsync(he, e, sizeof(double) * n, cudaMemcpyDeviceToHost, st));

  std::vector<DcNode> leaves, merges;
  std::vector<int> cuts, merge_n1;
  dc_collect(0, int(n), leaves, cuts, merges, merge_n1);
  std::vector<int> lo(leaves.size()), ll(leaves.size());
  for (size_t i = 0; i < leaves.size(); ++i) {
    lo[i] = leaves[i].off;
    ll[i] = leaves[i].len;
  }
  TQ_CUDA_CHECK(cudaStreamSynchronize(st));
  TQ_CUDA_CHECK(cudaMemcpyAsync(d_leaf_off, lo.data(), sizeof(int) * lo.size(), cudaMemcpyHostToDevice, st));
  TQ_CUDA_CHECK(cudaMemcpyAsync(d_leaf_len, ll.data(), sizeof(int) * ll.size(), cudaMemcpyHostToDevice, st));
  if (!cuts.empty())
    TQ_CUDA_CHECK(cudaMemcpyAsync(d_cuts, cuts.data(), sizeof(int) * cuts.size(), cudaMemcpyHostToDevice, st));
  TQ_CUDA_CHECK(cudaMemsetAsync(d_fail, 0, sizeof(int), st));
  TQ_CUDA_CHECK(cudaMemsetAsync(Z, 0, sizeof(double) * n * n, st));
  if (!cuts.empty()) {
    dc_tear_serial_kernel<<<1, 1, 0, st>>>(d, e, d_cuts, int(cuts.size()));
    TQ_LAUNCH_CHECK();
  }
  TQ_CUDA_CHECK(cudaFuncSetAttribute(dc_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kLeafSmem)));
  dc_leaf_kernel<<<(unsigned)leaves.size(), kLeaf, kLeafSmem, st>>>(d, e, Z, n, d_leaf_off, d_leaf_len, d_fail);
  TQ_LAUNCH_CHECK();
  TQ_CUDA_CHECK(cudaStreamSynchronize(st));
  // merges grouped by HEIGHT in the tree (children first): one D2H of (d, z), one synchronise and one H2D of the
  // deflation results per level instead of per merge; the merges of a level cover disjoint index ranges, so they
  // share the n-sized staging arrays at their own offsets
  std::vector<int> height(merges.size(), 0);
  {
    // the tree is fixed by (off, len): halve until a leaf, as dc_collect does
    auto height_of = [&](auto&& self, int off, int len) -> int {
      if (len <= kLeaf) return 0;
      const int n1 = len / 2;
      const int hl = self(self, off, n1), hr = self(self, off + n1, len - n1);
      return (hl > hr ? hl : hr) + 1;
    };
    for (size_t q = 0; q < merges.size(); ++q) height[q] = height_of(height_of, merges[q].off, merges[q].len);
  }
  int max_h = 0;
  for (int hq : height) max_h = hq > max_h ? hq : max_h;
  static const char* kLevelNames[] = {"dc level 1", "dc level 2", "dc level 3", "dc level 4", "dc level 5",
                                      "dc level 6", "dc level 7", "dc level 8+"};
  for (int lvl = 1; lvl <= max_h; ++lvl) {
    StageTimer level_tm(st, kLevelNames[lvl <= 8 ? lvl - 1 : 7]);
    std::vector<DcMergePlan> plans;
    size_t scratch = 0;
    for (size_t q = 0; q < merges.size(); ++q) {
      if (height[q] != lvl) continue;
      DcMergePlan mp{};
      mp.off = merges[q].off;
      mp.len = merges[q].len;
      mp.n1 = merge_n1[q];
      mp.scratch = scratch;
      scratch += size_t(mp.len) * mp.len;
      plans.push_back(mp);
    }
    if (plans.empty()) continue;
    if (scratch > size_t(n) * n) {
      set_error("stedc: level scratch exceeds n^2");
      return TQ_ERR_WORKSPACE;
    }
    for (const DcMergePlan& mp : plans) {
      const double beta = he[mp.off + mp.n1 - 1];
      double* Zb = Z + mp.off + int64_t(mp.off) * n;
      dc_build_z_kernel<<<(unsigned)ceil_div(mp.len, 256), 256, 0, st>>>(Zb, n, mp.n1, mp.len, beta < 0 ? -1.0 : 1.0,
                                                                         B.z + mp.off);
      TQ_LAUNCH_CHECK();
    }
    TQ_CUDA_CHECK(cudaMemcpyAsync(B.h_d, d, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    TQ_CUDA_CHECK(cudaMemcpyAsync(B.h_z, B.z, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    TQ_CUDA_CHECK(cudaStreamSynchronize(st));
    for (DcMergePlan& mp : plans) TQ_TRY(dc_merge_host(mp, he[mp.off + mp.n1 - 1], B));
    TQ_CUDA_CHECK(cudaMemcpyAsync(B.rot, B.h_rot, sizeof(DcRot) * n, cudaMemcpyHostToDevice, st));
    TQ_CUDA_CHECK(cudaMemcpyAsync(B.gidx, B.h_gidx, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    TQ_CUDA_CHECK(cudaMemcpyAsync(B.rowpos, B.h_rowpos, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    TQ_CUDA_CHECK(cudaMemcpyAsync(B.dd, B.h_dd, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    TQ_CUDA_CHECK(cudaMemcpyAsync(B.dl, B.h_dl, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    TQ_CUDA_CHECK(cudaMemcpyAsync(B.zl, B.h_zl, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    TQ_CUDA_CHECK(cudaMemcpyAsync(B.z2, B.h_z2, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    for (const DcMergePlan& mp : plans) TQ_TRY(dc_merge_device(h, st, d, Z, n, mp, B, want));
  }
  int fail = 0;
  TQ_CUDA_CHECK(cudaMemcpyAsync(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, st));
  TQ_CUDA_CHECK(cudaStreamSynchronize(st));
  if (fail) {
    set_error("stedc: QL iteration did not converge on a leaf");
    return TQ_ERR_NOCONV;
  }
  return TQ_OK;
}

// ======================================================================= ormtr
// Z <- Q Z, Q = H_0 H_1 ... H_{n-2} from sytrd_lower.
static int ormtr_lower(cublasHandle_t h, cudaStream_t st, const double* A, const double* tau, int64_t n, double* Z,
                       int64_t ncolsZ, double* Vc, double* G, double* T, double* w1, double* w2) {
  const int64_t lda = n, ldz = n;
  const int64_t nref = n - 1;
  if (nref <= 0) return TQ_OK;
  const int64_t nblk = ceil_div(nref, kOrmNb);
  for (int64_t b = nblk - 1; b >= 0; --b) {
    const int64_t j0 = b * kOrmNb;
    const int jb = int(imin(kOrmNb, nref - j0));
    const int64_t s = n - j0 - 1;
    dim3 grid((unsigned)imin(ceil_div(s, 256), 256), (unsigned)jb);
    copy_reflectors_kernel<<<grid, 256, 0, st>>>(A + (j0 + 1) + j0 * lda, lda, s, jb, Vc, s);
    TQ_LAUNCH_CHECK();
    TQ_TRY(build_t_factor(h, st, Vc, s, s, jb, tau + j0, G, T));
    TQ_TRY(apply_block_reflector(h, Vc, s, s, jb, T, jb, /*trans_t=*/false, Z + (j0 + 1), ldz, ncolsZ, w1, w2));
  }
  return TQ_OK;
}

// A (n x n, both triangles) = the symmetric matrix defined by the LOWER triangle of the
// row-major H, like torch.linalg.eigh's default UPLO='L' (gptq_utils.py:93).
__global__ void copy_sym_kernel(const double* __restrict__ H, int64_t ldh, int64_t n, double* __restrict__ A) {
  int64_t c = blockIdx.y;
  for (int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < n; r += int64_t(gridDim.x) * blockDim.x)
    A[r + c * n] = (r >= c) ? H[r * ldh + c] : H[c * ldh + r];
}

int copy_symmetric_lower(cudaStream_t st, const double* H, int64_t ldh, int64_t n, double* A) {
  dim3 grid((unsigned)imin(ceil_div(n, 256), 64), (unsigned)n);
  copy_sym_kernel<<<grid, 256, 0, st>>>(H, ldh, n, A);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

// two_stage.cu (experimental, off by default)
bool two_stage_usable(int64_t n);
size_t two_stage_ws_bytes(int64_t n);
int two_stage_reduce(cublasHandle_t h, cudaStream_t st, double* A, int64_t n, double* d, double* e, Workspace& ws);
int two_stage_back(cublasHandle_t h, cudaStream_t st, const double* A, int64_t n, double* Z, int64_t ncols,
                   Workspace scratch);

size_t eigh_ws_bytes(int64_t n) {
  size_t b = two_stage_usable(n) ? two_stage_ws_bytes(n) : 0;
  b += ws_bytes_for(size_t(n) * n, 8) * 4;                 // A, Zg, Zo, U
  b += ws_bytes_for(n, 8) * (14 + 32) + ws_bytes_for(2 * kTrdNb * 32, 8) + ws_bytes_for(n, 4) * 8 + ws_bytes_for(n, sizeof(DcRot));
  b += ws_bytes_for(size_t(n) * kTrdNb, 8) * 5 + ws_bytes_for(size_t(n) * kOrmNb, 8);   // W, XY, Vc
  b += ws_bytes_for(size_t(kRowSlots) * n, 8) + ws_bytes_for(size_t((n + kTileR - 1) / kTileR + 2) * n, 8) + ws_bytes_for(size_t((n + kTileR - 1) / kTileR + 2) * 2 * kTrdNb, 8) + 1024;   // symmetric panel partials
  b += ws_bytes_for(size_t(kOrmNb) * n, 8) * 2;            // w1, w2
  b += ws_bytes_for(kOrmNb * kOrmNb, 8) * 2 + ws_bytes_for(4 * kTrdNb, 8);
  return b;
}

// w (n, ascending) and Zout (n x n column-major, ld n: column i = eigenvector i).
// `choose` (optional) runs once all eigenvalues are on the device (after the divide & conquer, before any
// eigenvector is back-transformed) and names the columns [col0, col0 + ncols) the caller needs: only those are
// back-transformed - the other columns of Zout keep eigenvectors of the TRIDIAGONAL matrix and must not be used.
int eigh_colmajor(cublasHandle_t h, cudaStream_t st, const double* H, int64_t ldh, int64_t n, double* w,
                  double* Zout, Workspace& ws, const EighColumnChooser& choose) {
  double* A = ws.take<double>(size_t(n) * n);
  double* e = ws.take<double>(n);
  double* tau = ws.take<double>(n);
  double* y = ws.take<double>(size_t(n) * kPanelWarps);
  double* W = ws.take<double>(size_t(n) * kTrdNb);
  double* XY = ws.take<double>(size_t(n) * 4 * kTrdNb);
  const size_t nrb_max = size_t((n + kTileR - 1) / kTileR) + 2;
  SymBuffers sb;
  sb.rowpart = ws.take<double>(size_t(kRowSlots) * n);
  sb.colpart = ws.take<double>(nrb_max * n);
  sb.wvpart = ws.take<double>(nrb_max * 2 * kTrdNb);
  double* tmp = ws.take<double>(2 * kTrdNb * kPanelWarps);
  double* part = ws.take<double>(4096);
  double* scal = ws.take<double>(16);
  unsigned int* bar = ws.take<unsigned int>(4);
  double* G = ws.take<double>(kOrmNb * kOrmNb);
  double* T = ws.take<double>(kOrmNb * kOrmNb);
  if (ws.overflow) {
    set_error("eigh: workspace too small");
    return TQ_ERR_WORKSPACE;
  }
  {
    StageTimer tm(st, "copy_sym");
    dim3 grid((unsigned)imin(ceil_div(n, 256), 64), (unsigned)n);
    copy_sym_kernel<<<grid, 256, 0, st>>>(H, ldh, n, A);
    TQ_LAUNCH_CHECK();
  }
  const bool two_stage = two_stage_usable(n);
  if (two_stage) {
    // band reduction + bulge chasing instead of the one-stage reduction (two_stage.cu)
    StageTimer tm(st, "sytrd (two-stage)");
    g_sytrd_notified = false;
    TQ_TRY(two_stage_reduce(h, st, A, n, w, e, ws));
  } else {
    StageTimer tm(st, "sytrd");
    TQ_TRY(sytrd_lower(h, st, A, n, w, e, tau, W, y, tmp, part, scal, bar, XY, sb));
    if (trace_enabled()) {
      double hc[16];
      cudaMemcpyAsync(hc, scal, sizeof(hc), cudaMemcpyDeviceToHost, st);
      cudaStreamSynchronize(st);
      fprintf(stderr, "[tq-trace] sytrd phase Mcycles (CTA 0): A %.1f  barrier1 %.1f  B(stream) %.1f  barrier2 %.1f  C %.1f (C1 %.1f C2 %.1f)  [A row loop %.1f]\n",
              (hc[8] + hc[15]) * 1e-6, hc[9] * 1e-6, hc[10] * 1e-6, hc[11] * 1e-6, (hc[12] + hc[13] + hc[14]) * 1e-6, hc[13] * 1e-6, hc[14] * 1e-6, hc[15] * 1e-6);
    }
  }
  if (stage_callback_set() && !g_sytrd_notified) {   // the bandwidth-bound part is over once the stream drains
    TQ_CUDA_CHECK(cudaStreamSynchronize(st));
    notify_stage(TQ_STAGE_SYTRD_DONE);
  }
  {
    Workspace sub = ws;   // D&C scratch is released afterwards
    DcWanted wanted;
    {
      StageTimer tm(st, "stedc");
      wanted.choose = &choose;
      TQ_TRY(stedc(h, st, w, e, Zout, n, sub, &wanted));
    }
    if (sub.overflow) return TQ_ERR_WORKSPACE;
    // back-transform scratch overlays the D&C scratch
    Workspace sub2 = ws;
    double* Vc = sub2.take<double>(size_t(n) * kOrmNb);
    double* w1 = sub2.take<double>(size_t(kOrmNb) * n);
    double* w2 = sub2.take<double>(size_t(kOrmNb) * n);
    if (sub2.overflow) {
      set_error("eigh: workspace too small (ormtr)");
      return TQ_ERR_WORKSPACE;
    }
    int64_t col0 = 0, ncols = n;
    if (wanted.asked) {            // the root merge asked already and formed only those columns
      col0 = wanted.col0;
      ncols = wanted.ncols;
    } else if (choose) {
      TQ_TRY(choose(w, &col0, &ncols));
      if (col0 < 0 || ncols < 0 || col0 + ncols > n) {
        set_error("eigh: column chooser returned [%lld, +%lld) outside [0, %lld)", (long long)col0, (long long)ncols,
                  (long long)n);
        return TQ_ERR_INVALID;
      }
    }
    if (ncols == 0) return TQ_OK;
    if (two_stage) {
      TQ_TRY(two_stage_back(h, st, A, n, Zout + col0 * n, ncols, ws));
      return TQ_OK;
    }
    StageTimer tm(st, "ormtr");
    TQ_TRY(ormtr_lower(h, st, A, tau, n, Zout + col0 * n, ncols, Vc, G, T, w1, w2));
  }
  return TQ_OK;
}

}  // namespace tq

using namespace tq;

extern "C" int tq_eigh(const double* H, int64_t ldh, int64_t n, double* w, double* V, int64_t ldv, void* ws,
                       size_t ws_bytes, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(H && w && V && n > 0 && ldh >= n, "tq_eigh: bad arguments");
  TQ_REQUIRE(ldv == n, "tq_eigh: V must be contiguous (ldv == n)");
  TQ_REQUIRE(n < (1 << 30), "tq_eigh: n too large");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace wsp(ws, ws_bytes);
  cublasHandle_t h;
  TQ_TRY(get_cublas(&h, st));
  // column-major Z (column i = eigenvector i) is the same memory as row-major V (row i = eigenvector i)
  return eigh_colmajor(h, st, H, ldh, n, w, V, wsp, nullptr);
}
