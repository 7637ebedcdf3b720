/*
 * truncgptq.h - C ABI of libtruncgptq.so: the B200-native (sm_100a) TruncGPTQ
 * solve-and-quantize hot path.
 *
 * The reference (davidtweedle/gptq-svd) has no FFI boundary: its hot path is the
 * Python module src/TruncGPTQ/gptq_utils.py imported at src/TruncGPTQ/quantize.py:14.
 * Each entry point below replaces the body of one of those Python functions; the
 * host-side mirror (gptq_svd_b200/gptq_utils.py) keeps the Python signatures and
 * calls these through ctypes.  INTEGRATION.md shows the binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless
 *     the parameter name ends in _host;
 *   - matrices are row-major with an explicit leading dimension (elements);
 *   - `stream` is a cudaStream_t passed as void*; calls are stream-ordered.  Host synchronisations (every one of
 *     them a cudaStreamSynchronize on `stream`, nothing device-wide):
 *       tq_rank_select, tq_spectral_solve   one 16-byte read-back of (k, clamp flag): k sizes the later launches;
 *       tq_eigh / tq_spectral_solve         divide & conquer: one O(n) read-back per LEVEL of the merge tree
 *                                           (deflation runs on the host, 7 levels at n = 12288), one at the root
 *                                           merge for the columns the caller wants, one per solve for the leaf
 *                                           status; the two-stage back-transformation one at its end (it releases
 *                                           a pinned staging table); the stage callback synchronises before it runs;
 *       tq_spectral_solve                   one status read-back per pivoted Cholesky (non-positive pivot), one per
 *                                           Cholesky of the R-from-R_x stage; the Householder QRCP path
 *                                           (TQ_SOLVE_HOUSEHOLDER_QRCP) one per DLAQPS panel (its length is data-
 *                                           dependent);
 *       tq_cholesky_solve                   one status read-back per factorisation of the damping ladder;
 *       tq_gptq_loop                        one 4-byte read-back that validates `perm` (range, bijection) before
 *                                           anything is gathered;
 *     tq_syrk_accum*, tq_cast_*, tq_hessian_*, tq_find_params, tq_pack_*, tq_quant_error and
 *     tq_sketch_accum never synchronise.  TQ_TRACE=1 adds one per stage (timers), TQ_TRACE=2 times the
 *     stages with events and adds one at the end of a solve;
 *   - workspace is caller-provided (query with the *_workspace function);
 *   - return value: TQ_OK (0) or a negative status; tq_last_error() returns a
 *     thread-local description of the last failure;
 *   - no global mutable state, one host thread per GPU; there is NO CPU fallback:
 *     every compute entry point fails with TQ_ERR_CUDA when no sm_100 device is
 *     current.
 */
#ifndef TRUNCGPTQ_H_
#define TRUNCGPTQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TQ_VERSION 100 /* 0.1.0 */

/* status codes */
#define TQ_OK 0
#define TQ_ERR_INVALID (-1)     /* bad argument (shape, alignment, null pointer)    */
#define TQ_ERR_CUDA (-2)        /* CUDA runtime / driver failure                    */
#define TQ_ERR_WORKSPACE (-3)   /* workspace too small                              */
#define TQ_ERR_NOCONV (-4)      /* an iterative stage did not converge              */
#define TQ_ERR_UNSUPPORTED (-5) /* valid request outside the implemented envelope   */

/* element types */
#define TQ_F16 0
#define TQ_BF16 1
#define TQ_F32 2
#define TQ_F64 3

/* rank rules of process_hessian_alt (gptq_utils.py:97-108) */
#define TQ_RANK_ENERGY 0
#define TQ_RANK_MEAN_TRIMMED 1
#define TQ_RANK_FULL 2
/* OR-ed into `method` of tq_spectral_solve: obtain perm / R_x from the Householder
 * column-pivoted QR of S (LAPACK DLAQPS semantics, BLAS-2 bound) instead of the default
 * diagonally pivoted Cholesky of S^T S (same pivots and factor, BLAS-3 bound). */
#define TQ_SOLVE_HOUSEHOLDER_QRCP 0x100

/* loop arithmetic (gptq_utils.py:507-534) */
#define TQ_LOOP_TRITON 0 /* Triton kernel semantics: half-up, un-scaled error (:345-386) */
#define TQ_LOOP_TORCH 1  /* torch fallback semantics: half-even, error / diag (:516-534)  */
/* OR-ed into `semantics`: run the lazy trailing update as a strict-fp32 SIMT GEMM instead of
 * the default tcgen05 3xTF32 GEMM (both meet the 99.9 % code-parity bar; the reference runs
 * this product in fp32 with TF32 disabled, :474-475). */
#define TQ_LOOP_STRICT_FP32 0x100

int tq_version(void);
const char* tq_last_error(void);

/* Cap the number of SMs the calling thread's persistent (co-resident) kernels occupy; 0 = all.
 * The panel kernels of the solver are cooperative launches sized to this budget, so several
 * solves - one host thread and one stream each - can be in flight on one GPU
 * (gptq_svd_b200/concurrent.py).  Partial sums are added in a fixed, grid-dependent order: results are
 * reproducible for a given budget and agree across budgets to rounding (R within 2e-12 relative). */
int tq_set_sm_budget(int sms);

/* Per thread.  tq_eigh / tq_spectral_solve reduce H to tridiagonal form either in one stage (blocked DSYTRD with a
 * TMA-staged symmetric panel kernel) or in TWO stages (band reduction of width 64 by QR panels and DSYMM / DSYR2K,
 * then bulge chasing on the L2-resident band; two back-transformations).  -1 (default) = automatic: two stages for
 * n >= 8192 with n % 64 == 0 (380 ms against 820 ms at n = 12288 on a B200, level at n = 4096), one stage otherwise;
 * 1 = two stages whenever n % 64 == 0 and n >= 256; 0 = always one stage.  The workspace query depends on this
 * setting: query after changing it.  gptq_svd_b200/csrc/two_stage.cu. */
int tq_set_eigh_two_stage(int on);

/* Debugging aid for that path: runs the two reductions of an n x n symmetric H (n % 64 == 0, n >= 256) and returns
 * band_out (optional): the band matrix after stage 1 as a column-major 128 x n array, entry (i - c) + 128 c holds
 * B[i, c] for 0 <= i - c <= 64; d, e (n each): the tridiagonal matrix after stage 2.  H, band and T must share
 * their eigenvalues - which tells a failing stage from a sound one.  Workspace: tq_solver_workspace(n) bytes
 * queried with the two-stage path switched on. */
int tq_two_stage_debug(const double* H, int64_t ldh, int64_t n, double* band_out, double* d, double* e, void* ws,
                       size_t ws_bytes, void* stream);

/* Per-thread stage callback: `cb(stage, user)` runs on the calling host thread inside
 * tq_spectral_solve / tq_eigh when a stage boundary has been reached ON THE DEVICE (the stream is
 * synchronised first).  TQ_STAGE_SYTRD_DONE (once per solve): the tridiagonal reduction - the bandwidth-bound
 * part of a solve - is complete; what follows is latency- and DGEMM-bound, so a scheduler may lower this thread's
 * SM budget and start other solves next to it.  NULL removes the callback. */
#define TQ_STAGE_SYTRD_DONE 1
/* two-stage reduction only, before TQ_STAGE_SYTRD_DONE: the band reduction (DGEMM-bound, whole GPU) is complete and
 * the bulge chase that follows is a persistent kernel on at most n / 192 + 2 SMs (66 at n = 12288) */
#define TQ_STAGE_BAND_DONE 2
int tq_set_stage_callback(void (*cb)(int stage, void* user), void* user);

/* --------------------------------------------------------------------------
 * (1) Hessian accumulation - replaces HessianAccumulator.add_batch / get_hessian
 *     (gptq_utils.py:218-228).
 * -------------------------------------------------------------------------- */

/* H (n x n fp64, fully symmetric on return) += X^T X, X = rows x n (x_dtype TQ_F16 or
 * TQ_BF16), computed as a tcgen05/TMEM SYRK: fp16 products are exact in fp32, TMEM
 * accumulates `kc_tokens` tokens at a time (0 = default 256), chunk sums are added in
 * fp32 registers and the batch total is added to H in fp64.  Requires ldx % 8 == 0 and a
 * 16-byte aligned X (TMA).  gptq_utils.py:221-222. */
int tq_syrk_accum(double* H, int64_t ldh, const void* X, int x_dtype, int64_t rows, int64_t n,
                  int64_t ldx, int kc_tokens, void* stream);

/* out = H / n_samples (out = H when n_samples == 0).  gptq_utils.py:225-228. */
int tq_hessian_scale(const double* H, int64_t ldh, int64_t n, int64_t n_samples, double* out,
                     int64_t ldo, void* stream);

/* Same with H += alpha * X^T X, alpha read from DEVICE memory when the kernel runs (NULL = 1): the
 * epilogue of the SYRK multiplies the fp32 batch total by alpha in fp64.  Used with
 * tq_cast_to_f16_scaled, where alpha = 1 / scale^2 is a power of two (exact). */
int tq_syrk_accum_scaled(double* H, int64_t ldh, const void* X, int x_dtype, int64_t rows, int64_t n,
                         int64_t ldx, int kc_tokens, const double* alpha_dev, void* stream);

/* dst(fp16) = src (TQ_F32 / TQ_F64 / TQ_BF16 rows x n), saturating at +-65504, for activations that
 * arrive in another float type (the reference casts to fp64, gptq_utils.py:221). */
int tq_cast_to_f16(const void* src, int src_dtype, int64_t rows, int64_t n, int64_t lds, void* dst,
                   int64_t ldd, void* stream);

/* Range-safe version: dst(fp16) = src * scale with scale = 2^e chosen ON THE DEVICE from the batch's
 * largest magnitude so that it lands in [2^13, 2^14) - no overflow for any finite fp32 / fp64 input, the
 * full fp16 significand for the largest entries.  scratch32_dev: 32 bytes, 16-byte aligned, written as
 * {u32 amax bits, f32 scale, -, -, f64 alpha = 1 / scale^2 at byte 16}; pass scratch32_dev + 16 as
 * alpha_dev to tq_syrk_accum_scaled.  A NaN or an infinity in src sets *status_dev |= 1 (the caller reads
 * it back when it wants to fail: HessianAccumulator.get_hessian does).  No host synchronisation. */
int tq_cast_to_f16_scaled(const void* src, int src_dtype, int64_t rows, int64_t n, int64_t lds, void* dst,
                          int64_t ldd, void* scratch32_dev, int* status_dev, void* stream);

/* Guard against a corrupted accumulation (cheap: one pass over X per batch, one over H per check).
 * With the fixed sign vector v (v_j = +-1 from a hash of j): v^T (X^T X) v = ||X v||^2.
 *   tq_hessian_probe_accum: *probe_dev += alpha * sum_rows (x_row . v)^2 for the fp16 / bf16 batch the
 *       SYRK consumed (alpha_dev as above, NULL = 1);
 *   tq_hessian_probe_check: *vhv_dev = v^T H v; *status_dev |= 2 unless
 *       |v^T H v - probe| <= tol * max(|v^T H v|, |probe|)   (also set when either is NaN). */
int tq_hessian_probe_accum(const void* X, int x_dtype, int64_t rows, int64_t n, int64_t ldx,
                           const double* alpha_dev, double* probe_dev, void* stream);
int tq_hessian_probe_check(const double* H, int64_t ldh, int64_t n, const double* probe_dev, double tol,
                           double* vhv_dev, int* status_dev, void* stream);

/* --------------------------------------------------------------------------
 * (2) Spectral solver - replaces process_hessian_alt (gptq_utils.py:87-126).
 * -------------------------------------------------------------------------- */

int tq_solver_workspace(int64_t n, size_t* bytes);

/* Whole solver.  H: n x n symmetric (fp64).  Outputs (all device, caller-allocated):
 *   R, Rx   n x n row-major buffers, leading dimension n; rows [0,k) are written
 *           (upper-trapezoidal, diag > 0): R^T R = P^T H_k^+ P, Rx^T Rx = P^T H_k P;
 *   perm    n int64 (column pivots of the QRCP, gptq_utils.py:114-115);
 *   eigvals n fp64, clamped at 1e-12, DESCENDING (gptq_utils.py:94);
 *   k_host  retained rank, written on the host after an internal 8-byte D2H copy.
 * threshold / method as gptq_utils.py:97-108. */
int tq_spectral_solve(const double* H, int64_t ldh, int64_t n, double threshold, int method,
                      double* R, double* Rx, int64_t* perm, double* eigvals, int64_t* k_host,
                      void* ws, size_t ws_bytes, void* stream);

/* Stages of the solver, exposed for stage-wise parity tests and profiling. */

/* Symmetric eigendecomposition (blocked tridiagonal reduction + divide & conquer +
 * Householder back-transform).  w: n eigenvalues ascending; V: n x n, eigenvector i is
 * ROW i of V (row-major), i.e. V = torch.linalg.eigh(H)[1].T.  gptq_utils.py:93. */
int tq_eigh(const double* H, int64_t ldh, int64_t n, double* w, double* V, int64_t ldv, void* ws,
            size_t ws_bytes, void* stream);

/* eig_desc = clamp(w, 1e-12) reversed; k by the rank rule.  gptq_utils.py:94-108. */
int tq_rank_select(const double* w_asc, int64_t n, double threshold, int method, double* eig_desc,
                   int64_t* k_host, void* ws, size_t ws_bytes, void* stream);

/* Column-pivoted Householder QR (LAPACK dgeqp3/dlaqps semantics) of the k x n matrix
 * A (row-major, lda); returns Rx (k x n upper-trapezoidal, sign-normalised to diag > 0)
 * and perm.  A is not modified.  gptq_utils.py:114-116,122-123. */
int tq_qrcp(const double* A, int64_t lda, int64_t k, int64_t n, double* Rx, int64_t ldr,
            int64_t* perm, void* ws, size_t ws_bytes, void* stream);

/* R factor (sign-normalised) of the unpivoted Householder QR of the k x n matrix A
 * (row-major).  Q is never formed.  gptq_utils.py:120-121,124. */
int tq_qr_r(const double* A, int64_t lda, int64_t k, int64_t n, double* R, int64_t ldr, void* ws,
            size_t ws_bytes, void* stream);

/* --------------------------------------------------------------------------
 * (3) Quantisation grid and blocked GPTQ loop - replaces Quantizer.find_params
 *     (gptq_utils.py:249-266) and gptq_fwrd (:459-565) incl. the Triton kernel (:298-386).
 * -------------------------------------------------------------------------- */

/* scale, zero: m x (n / g) fp32, g = group > 0 ? group : n.  n % g != 0 -> TQ_ERR_INVALID
 * (the reference asserts, gptq_utils.py:253). */
int tq_find_params(const float* W, int64_t ldw, int64_t m, int64_t n, int bits, int group, int sym,
                   float* scale, float* zero, void* stream);

int tq_gptq_loop_workspace(int64_t m, int64_t n, int64_t k, size_t* bytes);

/* Blocked GPTQ column loop.
 *   W      m x n fp32, original column order (not modified);
 *   R      k x n upper-trapezoidal factor (r_dtype TQ_F64 or TQ_F32; cast to fp32 first,
 *          gptq_utils.py:483), row-major, ldr;
 *   perm   n int64; scale/zero from tq_find_params (static grid of the ORIGINAL W);
 *   ref_block  the reference's block_size: pairs (c, j) inside one ref_block use
 *          R[c,j] * (1/R[c,c]), pairs across blocks use R[c,j] / R[c,c]
 *          (gptq_utils.py:374-377 vs :541); the kernel's own tiling is independent of it;
 *   semantics  TQ_LOOP_TRITON or TQ_LOOP_TORCH;
 *   Wq_out m x n fp32 dequantised weights, original column order (gptq_utils.py:556-557);
 *   codes_out  optional m x n uint8, code - min_q, original column order (new: the
 *          reference has no integer output, README.md:133). */
int tq_gptq_loop(const float* W, int64_t ldw, const void* R, int r_dtype, int64_t ldr, int64_t k,
                 const int64_t* perm, const float* scale, const float* zero, int64_t m, int64_t n,
                 int bits, int group, int sym, int ref_block, int semantics, float* Wq_out,
                 int64_t ldq, uint8_t* codes_out, int64_t ldc, void* ws, size_t ws_bytes,
                 void* stream);

/* Pack biased codes (m x n uint8, value < 2^bits) LSB-first along the input dimension into
 * little-endian uint32 words, ceil(n*bits/32) words per row. */
int tq_pack_codes(const uint8_t* codes, int64_t ldc, int64_t m, int64_t n, int bits,
                  uint32_t* packed, int64_t ldp, void* stream);

/* GPTQ / AutoGPTQ / vLLM checkpoint layout (the packed-weight output the reference lists as a roadmap item,
 * README.md:133; sequential groups, no act-order g_idx, README.md:43).  `codes`: rows x cols uint8, value < 2^bits.
 * out: [ceil(cols * bits / 32), rows] uint32 with leading dimension ldo: word w of row j holds bits [32 w, 32 w + 32)
 * of that row's LSB-first bitstream - for 2 / 4 / 8 bits 32 / bits consecutive columns per word, for 3 bits
 * AutoGPTQ's 32-values-in-3-words scheme (the same bitstream).  `add` is added to every code modulo 2^bits.
 *   qweight = tq_pack_gptq(codes [out_features x in_features], add = 0)        -> [in * bits / 32, out]
 *   qzeros  = tq_pack_gptq(zeros^T [n_groups x out_features], add = -1) ^T     (v1 stores zero - 1; see
 *             gptq_svd_b200/gptq_utils.py: export_gptq) */
int tq_pack_gptq(const uint8_t* codes, int64_t ldc, int64_t rows, int64_t cols, int bits, int add, uint32_t* out,
                 int64_t ldo, void* stream);

int tq_quant_error_workspace(int64_t m, int64_t n, int64_t k, size_t* bytes);

/* out2[0] = ||(W - Wq)[:,perm] Rx^T||_F^2, out2[1] = ||W[:,perm] Rx^T||_F^2 in fp32
 * arithmetic with fp64 norm accumulation.  Replaces log_quantization_error
 * (gptq_utils.py:275-291). */
int tq_quant_error(const float* W, int64_t ldw, const float* Wq, int64_t ldq, const void* Rx,
                   int rx_dtype, int64_t ldr, int64_t k, const int64_t* perm, int64_t m, int64_t n,
                   double* out2, void* ws, size_t ws_bytes, void* stream);

/* --------------------------------------------------------------------------
 * (4) Callers either side of the hot path: the two other
 *     front ends of the reference that end in the same gptq_fwrd loop.
 * -------------------------------------------------------------------------- */

int tq_cholesky_workspace(int64_t n, size_t* bytes);

/* Reference-GPTQ factor - replaces process_hessian (gptq_utils.py:129-165).
 *   H          n x n symmetric fp64 (row-major, ldh);
 *   perm       optional n int64: act-order permutation (argsort of diag(H), descending,
 *              :137-139, computed by the caller); the factor is taken of H[perm][:, perm];
 *   damp_percent  damping ladder damp = 10^e * damp_percent, e = 0..4 (:148-160): the first e for
 *              which both Cholesky factorizations succeed wins; *damp_exp_host = e, or -1 when all
 *              five fail and the identity is returned (:161-163);
 *   Hinv_chol  n x n row-major UPPER factor U with U^T U = (H + damp mean(diag H) I)^-1.
 * Host synchronisation: one flag read-back per factorization. */
int tq_cholesky_solve(const double* H, int64_t ldh, int64_t n, const int64_t* perm, double damp_percent,
                      double* Hinv_chol, int64_t ldo, int* damp_exp_host, void* ws, size_t ws_bytes,
                      void* stream);

int tq_sketch_accum_workspace(int64_t rank, int64_t rows, int64_t n, size_t* bytes);

/* Sketch accumulation - replaces the GEMM of Sketcher.hook_fn (gptq_utils.py:185-203):
 * Y (rank x n fp32) += Rb (rank x rows fp32, the caller's Gaussian block) @ float32(X) with X
 * rows x n (TQ_F16 / TQ_BF16 / TQ_F32 / TQ_F64).  fp32-grade arithmetic on the tensor cores (3xTF32 with
 * 128-deep TMEM accumulations added in fp32 registers; the reference runs an fp32 matmul with TF32 off).
 * ws: tq_sketch_accum_workspace(rank, rows, n) bytes. */
int tq_sketch_accum(float* Y, int64_t ldy, const float* Rb, int64_t ldr, const void* X, int x_dtype,
                    int64_t ldx, int64_t rank, int64_t rows, int64_t n, void* ws, size_t ws_bytes,
                    void* stream);

int tq_sketch_workspace(int64_t rank, int64_t n, size_t* bytes);

/* Sketch solver - replaces process_sketch (gptq_utils.py:33-84) on the scaled sketch Y
 * (rank x n fp32).  The singular values / right singular vectors the reference takes from
 * geqrf + svd are obtained as the eigen-decomposition of Y^T Y (fp64); the rank rule runs over
 * min(rank, n) values with a floor of 1 (:49-64); the pivot order and R then follow the same
 * stages as tq_spectral_solve.  R: n x n row-major buffer, rows [0, k) written. */
int tq_sketch_solve(const float* Y, int64_t ldy, int64_t rank, int64_t n, double threshold, int method,
                    double* R, int64_t* perm, int64_t* k_host, void* ws, size_t ws_bytes, void* stream);

/* --------------------------------------------------------------------------
 * Diagnostics (measurement only; no effect on results).
 * -------------------------------------------------------------------------- */

/* Number of libtruncgptq kernels launched by the calling thread since load (cuBLAS calls
 * are not counted). */
int64_t tq_launch_count(void);

/* Sampled timing of the library's own hot kernels.  After tq_profile_begin(every), every `every`-th launch of each
 * instrumented kernel BY THE CALLING THREAD is bracketed by CUDA events on its own stream.
 *   tq_profile_kernel(kind, ...) - may be called any number of times before tq_profile_end - synchronises the
 *     sampled launches of one kernel and returns their algorithmic work (bytes or flops, see the table), their
 *     milliseconds, how many were sampled / launched, the work of ALL launches of the kind, and
 *     sm_ms = sum over the sampled launches of (milliseconds x SM budget of the launching thread), so that
 *     sm_ms / ms is the average number of SMs the launches were confined to (tq_set_sm_budget);
 *   tq_profile_end returns the sums over the three BLAS-2 panel kernels of the solver (kinds 0..2, bytes) and
 *     releases the events.
 * kind                    kernel                                       work unit
 * TQ_PROF_SYTRD_SYM       sytrd_panel_sym_kernel (one-stage reduction)  bytes: per column len (len / 2 + 2 i) 8
 * TQ_PROF_SYTRD_COLDOT    sytrd_panel_kernel (odd n)                     bytes: rows x columns x 8
 * TQ_PROF_QRCP_PANEL      qrcp_panel_kernel (Householder QRCP)           bytes: rows x columns x 8
 * TQ_PROF_CHASE           sb2st_chase_kernel (two-stage, stage 2)        bytes the tasks move through L2: tasks x 3 x 64^2 x 8 x 2
 * TQ_PROF_PCHOL_PANEL     pchol_panel_kernel                             bytes: steps x live columns x 128 x 8 (panel history)
 * TQ_PROF_QR_CLUSTER      qr_cluster_panel_kernel                        bytes: rows x jb x 8 x 2
 * TQ_PROF_LOOP_BLOCK      gptq_block_kernel                              bytes: m x 128 x 4 x 2 (W block read + written)
 * TQ_PROF_TRAILING_TC     trailing_tc_kernel (tcgen05 3xTF32)            flops: 2 m N K (algorithmic; 3x are issued)
 * TQ_PROF_TRAILING_SEQ    trailing_update_kernel (in-block pairs, SIMT)  flops: 2 m N K
 * TQ_PROF_SYRK            syrk_tcgen05_kernel                            flops: rows n (n + 1) (one triangle)
 * TQ_PROF_METRIC          metric_tc_kernel (error metric, tcgen05 TF32)  flops: 2 (2 m) (k n - k^2 / 2), Rx upper trapezoidal */
#define TQ_PROF_SYTRD_SYM 0
#define TQ_PROF_SYTRD_COLDOT 1
#define TQ_PROF_QRCP_PANEL 2
#define TQ_PROF_CHASE 3
#define TQ_PROF_PCHOL_PANEL 4
#define TQ_PROF_QR_CLUSTER 5
#define TQ_PROF_LOOP_BLOCK 6
#define TQ_PROF_TRAILING_TC 7
#define TQ_PROF_TRAILING_SEQ 8
#define TQ_PROF_SYRK 9
#define TQ_PROF_METRIC 10
#define TQ_PROF_KINDS 11
int tq_profile_begin(int sample_every);
int tq_profile_kernel(int kind, double* work, double* ms, int64_t* sampled, int64_t* total, double* work_all,
                      double* sm_ms);
int tq_profile_end(double* alg_bytes, double* ms, int64_t* sampled, int64_t* total);

#ifdef __cplusplus
}
#endif
#endif /* TRUNCGPTQ_H_ */
