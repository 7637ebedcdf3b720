// Two-stage tridiagonal reduction for tq_eigh (EXPERIMENTAL, off by default: tq_set_eigh_two_stage(1) or
// TQ_EIGH_TWO_STAGE=1; n % 64 == 0, n >= 256).  Written at the end of round 1 with no GPU time left: it compiles for
// sm_100a and has been executed on the CPU - the kernels of two_stage_kernels.cuh and this file's host driver are
// compiled unchanged by tests/emu/ (one OS thread per CUDA thread, reference BLAS in place of cuBLAS) and checked
// against the numpy model scripts/prototypes/sb2st_band.py and numpy.linalg.eigh (tests/test_two_stage_emu.py) -
// and then ran correctly on a B200 the first time (profiles/r01_two_stage_probe.log): at n = 12288 the reduction
// takes 196 (sy2sb) + 205 (sb2st) ms against 822 ms one-stage, the extra back-transformation 429 ms, tq_eigh 1051 ms
// against 1048 ms - level, hence still off.  tests/test_gpu_two_stage.py (TQ_TEST_TWO_STAGE=1) is its parity test.
//
// Why.  The one-stage reduction (eigh.cu) streams the lower triangle of the trailing matrix once per COLUMN:
// 4 n^3 / 3 bytes, and sytrd_panel_sym_kernel already runs that stream at ~0.93 of the HBM peak (0.82 s at
// n = 12288, 52 % of a decoder layer).  Only fewer bytes help:
//   stage 1  sy2sb   A = Q1 B Q1^T, B a band of width b = 64: QR of the block column below the band (cluster /
//            DSMEM panel kernel of qr.cu) + ONE symmetric rank-2b update per block column (DSYMM + DSYR2K): the
//            trailing matrix is read once per 64 columns, 4 n^3 / 3 flop of fp64 tensor-core work;
//   stage 2  sb2st   B = Q2 T Q2^T: bulge chasing on the (2 b) x n band array, which stays in L2 (12.6 MB at
//            n = 12288).  One persistent kernel; a CTA owns a sweep and walks its tasks, the bulge block travels
//            from task to task in shared memory, consecutive sweeps run ~2.4 tasks apart under acquire / release
//            progress counters (the protocol is checked with half-task interleavings in the numpy model);
//   back     Z = Q1 (Q2 Z_T): Q2 in wavefronts of row-disjoint (127 x 64) staircase block reflectors - one
//            strided-batched DGEMM triple per wavefront, groups of a wavefront sit 3 b rows apart in Z; Q1 like
//            ormtr with the reflector staircase shifted down by b rows.
// The bulge chase is bound by its dependency chain (its CTAs wait half the time; per task 6.4k cycles for the
// reflector / G block / progress fence, 3.0k for D, 1.3k for E): DESIGN.md 3.10 lists what to change.
#include <stdlib.h>

#include <vector>
#include "solver_kernels.cuh"
#include "two_stage_kernels.cuh"

namespace tq {

int qr_r_colmajor_tau(cublasHandle_t h, cudaStream_t st, double* A, int64_t lda, int64_t k, int64_t n, Workspace& ws,
                      double* tau_out);
int two_stage_setting();    // -1 automatic, 0 off, 1 on (tq_set_eigh_two_stage)

static inline int64_t chase_tasks(int64_t s, int64_t n) { return s > n - 3 ? 0 : (n - 3 - s) / kBw + 1; }

// ------------------------------------------------------------------------------------------------ buffers
struct TwoStageBuffers {
  double* Vs = nullptr;     // n x n
  double* Bd = nullptr;     // kLdb x n
  double* tau1 = nullptr;   // n
  double* tau2 = nullptr;   // n x (n / b + 2)
  int* prog = nullptr;      // n
  long long* stats = nullptr;   // 8 cycle counters of the bulge-chase kernel (TQ_TRACE)
};

size_t two_stage_ws_bytes(int64_t n) {
  size_t b = ws_bytes_for(size_t(n) * n, 8) + ws_bytes_for(size_t(kLdb) * n, 8) + ws_bytes_for(n, 8);
  b += ws_bytes_for(size_t(n) * (n / kBw + 2), 8) + ws_bytes_for(n, 4) + ws_bytes_for(8, 8);
  return b;
}

static int take_two_stage(Workspace& ws, int64_t n, TwoStageBuffers& tb) {
  tb.Vs = ws.take<double>(size_t(n) * n);
  tb.Bd = ws.take<double>(size_t(kLdb) * n);
  tb.tau1 = ws.take<double>(n);
  tb.tau2 = ws.take<double>(size_t(n) * (n / kBw + 2));
  tb.prog = ws.take<int>(n);
  tb.stats = ws.take<long long>(8);
  if (ws.overflow) {
    set_error("eigh (two-stage): workspace too small - query tq_solver_workspace after tq_set_eigh_two_stage(1)");
    return TQ_ERR_WORKSPACE;
  }
  return TQ_OK;
}

// ------------------------------------------------------------------------------------------------ stage 1
// A (n x n column-major, LOWER triangle referenced and updated) -> band of width b in place; the reflectors of
// block column j stay below R in A[j + b :, j : j + b), their scalars in tau1[j : j + b).
// scratch: overlays the D&C workspace (>= 3 n^2 doubles).
// Side stream of the look-ahead (one per host thread, created on first use): its own cuBLAS handle, two events.
struct SideStream {
  cudaStream_t s = nullptr;
  cublasHandle_t h = nullptr;
  cudaEvent_t ev_cols = nullptr, ev_qr = nullptr;
  int dev = -1;
};
static int get_side_stream(SideStream** out) {
  static thread_local SideStream ss;
  int dev = 0;
  TQ_CUDA_CHECK(cudaGetDevice(&dev));
  if (ss.s == nullptr || ss.dev != dev) {
    TQ_CUDA_CHECK(cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking));
    TQ_CUDA_CHECK(cudaEventCreateWithFlags(&ss.ev_cols, cudaEventDisableTiming));
    TQ_CUDA_CHECK(cudaEventCreateWithFlags(&ss.ev_qr, cudaEventDisableTiming));
    TQ_CUBLAS_CHECK(cublasCreate(&ss.h));
    TQ_CUBLAS_CHECK(cublasSetMathMode(ss.h, CUBLAS_PEDANTIC_MATH));
    TQ_CUBLAS_CHECK(cublasSetStream(ss.h, ss.s));
    ss.dev = dev;
  }
  *out = &ss;
  return TQ_OK;
}

static int sy2sb(cublasHandle_t h, cudaStream_t st, double* A, int64_t n, double* tau1, Workspace scratch) {
  constexpr int b = kBw;
  const int64_t lda = n;
  const double one = 1.0, zero = 0.0, mone = -1.0, mhalf = -0.5;
  double* Vc = scratch.take<double>(size_t(n) * b);
  double* X = scratch.take<double>(size_t(n) * b);
  double* X2 = scratch.take<double>(size_t(n) * b);
  double* G = scratch.take<double>(b * b);
  double* T = scratch.take<double>(b * b);
  double* M1 = scratch.take<double>(b * b);
  double* M2 = scratch.take<double>(b * b);
  if (scratch.overflow) {
    set_error("sy2sb: workspace too small");
    return TQ_ERR_WORKSPACE;
  }
  TQ_CUDA_CHECK(cudaMemsetAsync(tau1, 0, sizeof(double) * n, st));
  // (Measured and dropped, profiles/r02_two_stage_variants.log: X = A22 V as a DGEMM on a mirrored trailing
  // matrix instead of DSYMM - 205 vs 198 ms at n = 12288; a look-ahead that factors the next panel on a side
  // stream while the main stream runs the DSYR2K - 200 vs 198 ms.)
  const int use_gemm = 0, lookahead = 0;
  SideStream* side = nullptr;
  bool panel_ready = false;          // the current panel has been factored already (by the previous iteration)
  for (int64_t j = 0; j + b < n; j += b) {
    const int64_t r0 = j + b, s = n - r0;
    double* P = A + r0 + j * lda;
    double* A22 = A + r0 + r0 * lda;
    if (panel_ready) {
      TQ_CUDA_CHECK(cudaStreamWaitEvent(st, side->ev_qr, 0));      // its QR ran on the side stream
    } else {
      Workspace qws = scratch;                       // the panel factorisation's own scratch, released afterwards
      TQ_TRY(qr_r_colmajor_tau(h, st, P, lda, s, b, qws, tau1 + j));
    }
    dim3 grid((unsigned)imin(ceil_div(s, 256), 256), (unsigned)b);
    TQ_LAUNCH(copy_reflectors_kernel, grid, 256, 0, st, P, lda, s, b, Vc, s);
    TQ_LAUNCH_CHECK();
    TQ_TRY(build_t_factor(h, st, Vc, s, s, b, tau1 + j, G, T));
    // X = A22 V (lower triangle of A22 only), X2 = X T
    if (use_gemm) {
      if (j > 0) {     // the DSYR2K of the previous block column left only the lower triangle up to date
        const unsigned nt = unsigned(ceil_div(s, 32));
        TQ_LAUNCH(mirror_lower_colmajor_kernel, dim3(nt, nt), 256, 32 * 33 * sizeof(double), st, A22, lda, int(s));
        TQ_LAUNCH_CHECK();
      }
      TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, int(s), b, int(s), &one, A22, int(lda), Vc, int(s),
                                  &zero, X, int(s)));
    } else {
      TQ_CUBLAS_CHECK(cublasDsymm(h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, int(s), b, &one, A22, int(lda), Vc,
                                  int(s), &zero, X, int(s)));
    }
    TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, int(s), b, b, &one, X, int(s), T, b, &zero, X2, int(s)));
    // M2 = T^T (V^T X2);  W = X2 - V M2 / 2  (in X2)
    TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, b, b, int(s), &one, Vc, int(s), X2, int(s), &zero, M1, b));
    TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, b, b, b, &one, T, b, M1, b, &zero, M2, b));
    TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, int(s), b, b, &mhalf, Vc, int(s), M2, b, &one, X2, int(s)));
    panel_ready = false;
    if (lookahead && s > b) {
      // block column [0, b) of A22 (the next diagonal block and the next panel below it) first:
      //   A22[:, 0:b] -= V W[0:b, :]^T + W V[0:b, :]^T
      TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, int(s), b, b, &mone, Vc, int(s), X2, int(s), &one, A22,
                                  int(lda)));
      TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, int(s), b, b, &mone, X2, int(s), Vc, int(s), &one, A22,
                                  int(lda)));
      // the next panel (rows >= b of that block column) is final: factor it on the side stream ...
      TQ_CUDA_CHECK(cudaEventRecord(side->ev_cols, st));
      TQ_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->ev_cols, 0));
      Workspace qws = scratch;
      TQ_TRY(qr_r_colmajor_tau(side->h, side->s, A22 + b, lda, s - b, b, qws, tau1 + j + b));
      TQ_CUDA_CHECK(cudaEventRecord(side->ev_qr, side->s));
      panel_ready = true;
      // ... while the main stream updates the rest: rows and columns >= b
      TQ_CUBLAS_CHECK(cublasDsyr2k(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, int(s - b), b, &mone, Vc + b, int(s), X2 + b,
                                   int(s), &one, A22 + b + b * lda, int(lda)));
    } else {
      // A22 -= V W^T + W V^T
      TQ_CUBLAS_CHECK(cublasDsyr2k(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, int(s), b, &mone, Vc, int(s), X2, int(s), &one,
                                   A22, int(lda)));
    }
  }
  if (panel_ready) TQ_CUDA_CHECK(cudaStreamWaitEvent(st, side->ev_qr, 0));
  return TQ_OK;
}

// ------------------------------------------------------------------------------------------------ stage 2
static int sb2st(cudaStream_t st, const double* A, int64_t n, const TwoStageBuffers& tb, double* d, double* e,
                 double* band_out = nullptr) {
  TQ_LAUNCH(band_extract_kernel, unsigned(n), 128, 0, st, A, n, int(n), tb.Bd);
  TQ_LAUNCH_CHECK();
  if (band_out)     // debugging: the band matrix as stage 1 left it
    TQ_CUDA_CHECK(cudaMemcpyAsync(band_out, tb.Bd, sizeof(double) * size_t(kLdb) * n, cudaMemcpyDeviceToDevice, st));
  TQ_CUDA_CHECK(cudaMemsetAsync(tb.prog, 0, sizeof(int) * n, st));
  TQ_CUDA_CHECK(cudaMemsetAsync(tb.tau2, 0, sizeof(double) * size_t(n) * (n / kBw + 2), st));
  if (n > 2) {
    // the variant with the second wait in front of the D / E loads ("late loads": 182 ms at n = 12288 against 202
    // for the first version and 211 / 183 with a helper warp owning the progress counter,
    // profiles/r02_two_stage_variants.log); the other instantiations stay in two_stage_kernels.cuh for the
    // CPU emulation tests
    const int helper = 0, late = 1;
    const void* kfn = helper ? (late ? (const void*)sb2st_chase_kernel_t<true, true> : (const void*)sb2st_chase_kernel_t<true, false>)
                             : (late ? (const void*)sb2st_chase_kernel_t<false, true> : (const void*)sb2st_chase_kernel_t<false, false>);
    const int kthreads = kChaseThreads + (helper ? 32 : 0);
    TQ_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kChaseSmem)));
    // sweeps run 3 tasks apart, so at most K_0 / 3 + 1 of them are in flight; the launch is cooperative only for
    // its guarantee that every CTA is resident (a waiting sweep's predecessor must be running)
    int64_t want = chase_tasks(0, n) / 3 + 2;
    const int grid = int(imax(1, imin(num_sms(), want)));
    const bool trace = trace_enabled();
    if (trace) TQ_CUDA_CHECK(cudaMemsetAsync(tb.stats, 0, 8 * sizeof(long long), st));
    ChaseArgs ca{tb.Bd, int(n), tb.Vs, n, tb.tau2, tb.prog, trace ? tb.stats : nullptr};
    void* kargs[] = {&ca};
    double tasks = 0.0;
    for (int64_t s = 0; s < n; ++s) tasks += double(chase_tasks(s, n));
    const int pslot = prof_begin_launch(st, tasks * 3.0 * kBw * kBw * 16.0, TQ_PROF_CHASE);
    TQ_CUDA_CHECK(cudaLaunchCooperativeKernel(kfn, dim3(grid), dim3(kthreads), kargs, kChaseSmem, st));
    prof_end_launch(st, pslot);
    ++g_launch_count;
    if (trace) {
      long long hs[8];
      TQ_CUDA_CHECK(cudaMemcpyAsync(hs, tb.stats, sizeof(hs), cudaMemcpyDeviceToHost, st));
      TQ_CUDA_CHECK(cudaStreamSynchronize(st));
      const double t = double(hs[4] > 0 ? hs[4] : 1);
      fprintf(stderr, "[tq-trace] sb2st chase%s: %d CTAs; CTA 0 ran %lld tasks, cycles per task: wait %.0f  reflector+G %.0f  "
              "D %.0f  E %.0f\n", helper ? (late ? " (helper warp, late loads)" : " (helper warp)") : (late ? " (late loads)" : ""), grid, hs[4], hs[0] / t, hs[1] / t, hs[2] / t, hs[3] / t);
    }
  }
  TQ_CUDA_CHECK(cudaMemsetAsync(e, 0, sizeof(double) * n, st));      // e[n-1] = 0 like sytrd_lower leaves it
  TQ_LAUNCH(band_diag_kernel, unsigned(ceil_div(n, 256)), 256, 0, st, tb.Bd, int(n), d, e);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

// ------------------------------------------------------------------------------------------------ Z <- Q2 Z
// Of two overlapping reflectors the later generated one acts first; for groups that means sweep blocks DESCENDING
// and, inside a block, chase index ASCENDING.  Group (sb, k) overlaps (sb, k - 1) and (sb + 1, k - 2 .. k), so
// w = k + 2 (M - sb) is a valid wavefront number, and the groups of one wavefront start 3 b rows apart.
// The T factors and V T of the groups do not depend on Z: they are formed for a whole CHUNK of wavefronts at once
// (four launches per chunk instead of four per wavefront - with t = n - k ~ 0.1 n columns to transform, the chain of
// six small dependent launches per wavefront was most of this stage), which leaves two batched GEMMs over Z per
// wavefront.  Group g of the chunk: Vc / VT + g * kQ2Ld * b, the groups of one wavefront are consecutive.
static int q2_prepare_chunk(cublasHandle_t h, cudaStream_t st, const TwoStageBuffers& tb, int64_t n, const int* desc_dev,
                            int groups, double* Vc, double* taub, double* Gb, double* Tb, double* VT) {
  constexpr int b = kBw;
  const double one = 1.0, zero = 0.0;
  const long long sv = (long long)kQ2Ld * b, st_t = (long long)b * b;
  for (int g0 = 0; g0 < groups; g0 += 65535) {      // gridDim.z limit
    const int cnt = groups - g0 < 65535 ? groups - g0 : 65535;
    TQ_LAUNCH(copy_staircase_kernel, dim3(1, b, cnt), kQ2Ld, 0, st, tb.Vs, n, tb.tau2, int(n), 0, 0, Vc + g0 * sv,
              taub + int64_t(g0) * b, desc_dev + 2 * g0);
    TQ_LAUNCH_CHECK();
  }
  // rows past the matrix end are zero in the clean copy, so every group takes the full window here
  TQ_CUBLAS_CHECK(cublasDgemmStridedBatched(h, CUBLAS_OP_T, CUBLAS_OP_N, b, b, kQ2H + 1, &one, Vc, kQ2Ld, sv, Vc, kQ2Ld,
                                            sv, &zero, Gb, b, st_t, groups));
  TQ_LAUNCH(larft_kernel, groups, kLarftThreads, size_t(b) * b * 10, st, Gb, b, taub, b, Tb, b, int64_t(st_t), int64_t(b), int64_t(st_t));
  TQ_LAUNCH_CHECK();
  // (I - V T V^T) Z = Z - (V T)(V^T Z): T is folded into V once per group (128 x 64 x 64) instead of being applied to
  // the 64 x ncols product - one batched GEMM over Z fewer (20 % of the flops of this stage)
  TQ_CUBLAS_CHECK(cublasDgemmStridedBatched(h, CUBLAS_OP_N, CUBLAS_OP_N, kQ2H + 1, b, b, &one, Vc, kQ2Ld, sv, Tb, b, st_t,
                                            &zero, VT, kQ2Ld, sv, groups));
  return TQ_OK;
}

static int apply_q2_batch(cublasHandle_t h, cudaStream_t st, int sb0, int k0, int count, int hg, double* Z, int64_t ldz,
                          int64_t ncols, const double* Vc, const double* VT, double* w1) {
  constexpr int b = kBw;
  const double one = 1.0, zero = 0.0, mone = -1.0;
  const int64_t row0 = (int64_t(sb0) + k0) * b;       // one row above the staircase: see copy_staircase_kernel
  const int hw = hg + 1;                              // rows of the window (128 unless the group is clipped)
  const long long sv = (long long)kQ2Ld * b, sw = (long long)b * ncols, sz = 3 * b;
  double* Zb = Z + row0;
  TQ_CUBLAS_CHECK(cublasDgemmStridedBatched(h, CUBLAS_OP_T, CUBLAS_OP_N, b, int(ncols), hw, &one, Vc, kQ2Ld, sv, Zb,
                                            int(ldz), sz, &zero, w1, b, sw, count));
  TQ_CUBLAS_CHECK(cublasDgemmStridedBatched(h, CUBLAS_OP_N, CUBLAS_OP_N, hw, int(ncols), b, &mone, VT, kQ2Ld, sv, w1, b,
                                            sw, &one, Zb, int(ldz), sz, count));
  return TQ_OK;
}

static inline int q2_max_batch(int64_t n) { return int(n / (3 * kBw)) + 2; }

// The schedule: calls f(sb0, k0, count, hg) for every batch of `count` row-disjoint groups (sb0 + i, k0 + 2 i) of
// height hg, in an order that respects the dependencies above.  Only the lowest group of a wavefront can run past
// row n; it becomes a batch of its own (hg < kQ2H).  tests/test_two_stage_emu.py compares this enumeration with the
// numpy model at the Qwen3-8B sizes.
template <class F>
static int q2_for_each_batch(int64_t n, F&& f) {
  constexpr int b = kBw;
  const int64_t nsweeps = n - 2;
  if (nsweeps <= 0) return TQ_OK;
  const int M = int((nsweeps - 1) / b);
  const int maxb = q2_max_batch(n);
  const int kmax0 = int(chase_tasks(0, n)) - 1;
  for (int w = 0; w <= kmax0 + 2 * M; ++w) {
    int sb_lo = -1, sb_hi = -1;
    for (int sb = 0; sb <= M; ++sb) {
      const int k = w - 2 * (M - sb);
      if (k < 0 || k > int(chase_tasks(int64_t(sb) * b, n)) - 1) continue;
      if (sb_lo < 0) sb_lo = sb;
      if (sb_hi >= 0 && sb != sb_hi + 1) {
        set_error("apply_q2: wavefront %d is not contiguous", w);
        return TQ_ERR_UNSUPPORTED;
      }
      sb_hi = sb;
    }
    if (sb_lo < 0) continue;
    int count = sb_hi - sb_lo + 1;
    if (count > maxb) {
      set_error("apply_q2: wavefront %d has %d groups (max %d)", w, count, maxb);
      return TQ_ERR_UNSUPPORTED;
    }
    const int k_lo = w - 2 * (M - sb_lo), k_hi = w - 2 * (M - sb_hi);
    const int64_t rlo_last = int64_t(sb_hi) * b + 1 + int64_t(k_hi) * b;
    const int hg_last = int(imin(kQ2H, n - rlo_last));
    if (hg_last < kQ2H) {
      TQ_TRY(f(sb_hi, k_hi, 1, hg_last));
      --count;
    }
    if (count > 0) TQ_TRY(f(sb_lo, k_lo, count, kQ2H));
  }
  return TQ_OK;
}

// grow-only pinned staging buffer of this host thread (group tables of apply_q2)
static int pinned_scratch(size_t bytes, void** out) {
  static thread_local void* buf = nullptr;
  static thread_local size_t cap = 0;
  if (cap < bytes) {
    if (buf) cudaFreeHost(buf);
    buf = nullptr;
    cap = 0;
    TQ_CUDA_CHECK(cudaHostAlloc(&buf, bytes, cudaHostAllocPortable));
    cap = bytes;
  }
  *out = buf;
  return TQ_OK;
}

struct Q2Batch {
  int sb0, k0, count, hg;
};

static int apply_q2(cublasHandle_t h, cudaStream_t st, const TwoStageBuffers& tb, int64_t n, double* Z, int64_t ldz,
                    int64_t ncols, Workspace scratch) {
  constexpr int b = kBw;
  if (n <= 2) return TQ_OK;
  const int maxb = q2_max_batch(n);
  std::vector<Q2Batch> batches;
  TQ_TRY(q2_for_each_batch(n, [&](int sb0, int k0, int count, int hg) -> int {
    batches.push_back({sb0, k0, count, hg});
    return TQ_OK;
  }));
  size_t total = 0;
  for (const Q2Batch& q : batches) total += size_t(q.count);
  if (total == 0) return TQ_OK;
  double* w1 = scratch.take<double>(size_t(maxb) * b * ncols);
  int* desc_dev = scratch.take<int>(2 * total);
  // chunk size: what the scratch (an overlay of the D&C workspace) holds, at least one wavefront
  const size_t per_group = (size_t(2) * kQ2Ld * b + size_t(2) * b * b + b) * sizeof(double);
  size_t room = scratch.size > align_up(scratch.off, 256) + 8192 ? scratch.size - align_up(scratch.off, 256) - 8192 : 0;
  size_t gmax = room / per_group;
  if (gmax > 4096) gmax = 4096;
  if (gmax < size_t(maxb)) gmax = size_t(maxb);
  double* Vc = scratch.take<double>(gmax * kQ2Ld * b);
  double* VT = scratch.take<double>(gmax * kQ2Ld * b);
  double* Gb = scratch.take<double>(gmax * b * b);
  double* Tb = scratch.take<double>(gmax * b * b);
  double* taub = scratch.take<double>(gmax * b);
  if (scratch.overflow) {
    set_error("apply_q2: workspace too small");
    return TQ_ERR_WORKSPACE;
  }
  // group table of the whole schedule: one upload (pinned staging of this host thread)
  int* desc_host = nullptr;
  TQ_TRY(pinned_scratch(sizeof(int) * 2 * total, reinterpret_cast<void**>(&desc_host)));
  {
    size_t g = 0;
    for (const Q2Batch& q : batches)
      for (int i = 0; i < q.count; ++i, ++g) {
        desc_host[2 * g] = q.sb0 + i;
        desc_host[2 * g + 1] = q.k0 + 2 * i;
      }
  }
  TQ_CUDA_CHECK(cudaMemcpyAsync(desc_dev, desc_host, sizeof(int) * 2 * total, cudaMemcpyHostToDevice, st));
  const long long sv = (long long)kQ2Ld * b;
  size_t bi = 0, g_done = 0;
  while (bi < batches.size()) {
    size_t be = bi, groups = 0;
    while (be < batches.size() && groups + size_t(batches[be].count) <= gmax) groups += size_t(batches[be++].count);
    TQ_TRY(q2_prepare_chunk(h, st, tb, n, desc_dev + 2 * g_done, int(groups), Vc, taub, Gb, Tb, VT));
    size_t g = 0;
    for (; bi < be; ++bi) {
      const Q2Batch& q = batches[bi];
      TQ_TRY(apply_q2_batch(h, st, q.sb0, q.k0, q.count, q.hg, Z, ldz, ncols, Vc + g * sv, VT + g * sv, w1));
      g += size_t(q.count);
    }
    g_done += groups;
  }
  // desc_host is reused by the next call of this thread: the upload above must have been consumed
  TQ_CUDA_CHECK(cudaStreamSynchronize(st));
  return TQ_OK;
}

// ------------------------------------------------------------------------------------------------ Z <- Q1 Z
// Reflector c of stage 1 (c < n - b) has its unit entry in row c + b: the blocks are clean staircases that start b
// rows below the diagonal; otherwise exactly ormtr_lower (eigh.cu).
static int apply_q1(cublasHandle_t h, cudaStream_t st, const double* A, const double* tau1, int64_t n, double* Z,
                    int64_t ldz, int64_t ncols, Workspace scratch) {
  const int64_t nref = n - kBw;
  if (nref <= 0) return TQ_OK;
  double* Vc = scratch.take<double>(size_t(n) * kQ1Nb);
  double* G = scratch.take<double>(kQ1Nb * kQ1Nb);
  double* T = scratch.take<double>(kQ1Nb * kQ1Nb);
  double* w1 = scratch.take<double>(size_t(kQ1Nb) * ncols);
  double* w2 = scratch.take<double>(size_t(kQ1Nb) * ncols);
  if (scratch.overflow) {
    set_error("apply_q1: workspace too small");
    return TQ_ERR_WORKSPACE;
  }
  const int64_t nblk = ceil_div(nref, kQ1Nb);
  for (int64_t blk = nblk - 1; blk >= 0; --blk) {
    const int64_t j0 = blk * kQ1Nb;
    const int jb = int(imin(kQ1Nb, nref - j0));
    const int64_t s = n - j0 - kBw;
    dim3 grid((unsigned)imin(ceil_div(s, 256), 256), (unsigned)jb);
    TQ_LAUNCH(copy_reflectors_kernel, grid, 256, 0, st, A + (j0 + kBw) + j0 * n, n, s, jb, Vc, s);
    TQ_LAUNCH_CHECK();
    TQ_TRY(build_t_factor(h, st, Vc, s, s, jb, tau1 + j0, G, T));
    TQ_TRY(apply_block_reflector(h, Vc, s, s, jb, T, jb, /*trans_t=*/false, Z + (j0 + kBw), ldz, ncols, w1, w2));
  }
  return TQ_OK;
}

// ------------------------------------------------------------------------------------------------ entry points
// Automatic choice (tq_set_eigh_two_stage(-1), the default): two stages from n = 8192 on.  Measured on a B200
// (profiles/r02_solver_stages.log): tridiagonal reduction 380 ms against 820 ms one-stage at n = 12288; at
// n = 4096 the two are level (85 vs 101 ms) and the narrow solves of a decoder block run side by side under SM
// budgets, where the one-stage panel kernel has been tuned - so the small orders keep the one-stage path.
constexpr int64_t kTwoStageMinN = 8192;
bool two_stage_usable(int64_t n) {
  const int setting = two_stage_setting();
  if (setting == 0) return false;
  if (!(n % kBw == 0 && n >= 4 * kBw && n < (1 << 30))) return false;
  return setting == 1 || n >= kTwoStageMinN;
}

struct TwoStageState {
  TwoStageBuffers tb;
};
static thread_local TwoStageState g_ts;

// A -> (d, e); the reflectors of both stages stay in A / in buffers taken from `ws` (by reference: they must
// outlive the D&C stage); everything after them in `ws` is scratch until this returns.
int two_stage_reduce(cublasHandle_t h, cudaStream_t st, double* A, int64_t n, double* d, double* e, Workspace& ws) {
  TQ_TRY(take_two_stage(ws, n, g_ts.tb));
  Workspace scratch = ws;
  {
    StageTimer tm(st, "sy2sb");
    TQ_TRY(sy2sb(h, st, A, n, g_ts.tb.tau1, scratch));
  }
  if (stage_callback_set()) {      // the DGEMM-bound band reduction is over; the bulge chase occupies <= n / 192 + 2 SMs
    TQ_CUDA_CHECK(cudaStreamSynchronize(st));
    notify_stage(TQ_STAGE_BAND_DONE);
  }
  {
    StageTimer tm(st, "sb2st");
    TQ_TRY(sb2st(st, A, n, g_ts.tb, d, e));
  }
  return TQ_OK;
}

// Z (n x ncols, column-major, ld n) <- Q1 Q2 Z
int two_stage_back(cublasHandle_t h, cudaStream_t st, const double* A, int64_t n, double* Z, int64_t ncols,
                   Workspace scratch) {
  {
    StageTimer tm(st, "apply_q2");
    TQ_TRY(apply_q2(h, st, g_ts.tb, n, Z, n, ncols, scratch));
  }
  {
    StageTimer tm(st, "apply_q1");
    TQ_TRY(apply_q1(h, st, A, g_ts.tb.tau1, n, Z, n, ncols, scratch));
  }
  return TQ_OK;
}

int copy_symmetric_lower(cudaStream_t st, const double* H, int64_t ldh, int64_t n, double* A);

}  // namespace tq

using namespace tq;

extern "C" int tq_two_stage_debug(const double* H, int64_t ldh, int64_t n, double* band_out, double* d, double* e,
                                  void* ws, size_t ws_bytes, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(H && d && e && ldh >= n, "tq_two_stage_debug: bad arguments");
  TQ_REQUIRE(n % kBw == 0 && n >= 4 * kBw && n < (1 << 30), "tq_two_stage_debug: n must be a multiple of %d, >= %d", kBw,
             4 * kBw);
  cudaStream_t st = (cudaStream_t)stream;
  Workspace wsp(ws, ws_bytes);
  double* A = wsp.take<double>(size_t(n) * n);
  TwoStageBuffers tb;
  TQ_TRY(take_two_stage(wsp, n, tb));
  if (wsp.overflow) {
    set_error("tq_two_stage_debug: workspace too small");
    return TQ_ERR_WORKSPACE;
  }
  cublasHandle_t h;
  TQ_TRY(get_cublas(&h, st));
  TQ_TRY(copy_symmetric_lower(st, H, ldh, n, A));
  TQ_TRY(sy2sb(h, st, A, n, tb.tau1, wsp));
  return sb2st(st, A, n, tb, d, e, band_out);
}
