// R from R_x alone - the last stage of process_hessian_alt (reference gptq_utils.py:118-124) without the k retained
// eigenvectors.
//
// The reference obtains R (k x n upper trapezoidal, diag > 0, R^T R = P^T H_k^+ P) as the R factor of an unpivoted QR
// of B = Lambda_k^-1/2 V_k^T [:, perm]: k eigenvectors back-transformed (2 n^2 k flop, twice that on the two-stage
// path) and a k x n Householder QR (2 n k^2 - 2/3 k^3).  R is unique given (H_k, perm), and it follows from
// R_x = [R11 R12] (R_x^T R_x = P^T H_k P, from the pivoted Cholesky) by BLAS-3 work on k x k and k x t matrices
// (t = n - k):
//     Z   = R11^-1 R12                          (k x t)
//     M   = (I + Z Z^T) A11 (I + Z Z^T),        A11 = R11^T R11 = (P^T H_k P)[:k, :k]
//     M   = U U^T   (U upper: Cholesky from the bottom-right corner),     T11 = U^-1
//     R   = [T11, T11 Z]
// Proof: with G = R_x R_x^T = R11 (I + Z Z^T) R11^T, (P^T H_k P)^+ = R_x^T G^-2 R_x, so R = C R_x with C upper
// triangular and C^T C = G^-2; then T11 = C R11 satisfies (T11^T T11)^-1 = (I + Z Z^T) A11 (I + Z Z^T).
// A11 is taken from the DATA (the solver's copy of H_k, gathered through perm), not from R11^T R11, so the
// only ill-conditioned steps are one Cholesky and one triangular inverse: the error against the reference route is
// ~3e-16 cond(H_k) (bar: 2e-14 cond(H_k); numpy prototype at n = 768 / 1024, eps 1e-2 .. 0).
// Work: 2/3 k^3 + ~7 k^2 t, all DGEMM / DSYRK / DTRSM-shaped, against 2 n^2 k + 2 n k^2 - 2/3 k^3 before.
//
// Index flip.  "Cholesky from the bottom-right corner" is the ordinary lower Cholesky of J M J (J = index
// reversal): everything k-indexed is built flipped (r' = k - 1 - r), M' = L' L'^T, and T11 = J L'^-1 J.
#include <vector>

#include "solver_kernels.cuh"

namespace tq {

int chol_lower(cublasHandle_t h, cudaStream_t st, double* A, int64_t n, int* fail);

constexpr int kRfLeaf = 64;   // 2 x 64 x 65 doubles of shared memory per leaf (128 would need 264 KB)

// Zt (t x k, ld t) = rows [k, n) of the column-major L = R_x^T (n x k, ld n)
__global__ void rf_copy_l21_kernel(const double* __restrict__ L, int64_t n, int64_t k, int64_t t,
                                   double* __restrict__ Zt) {
  const int64_t c = blockIdx.y;
  for (int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < t; r += int64_t(gridDim.x) * blockDim.x)
    Zt[r + c * t] = L[(k + r) + c * n];
}

// Zf (k x t, ld k): Zf[r, i] = Zt[i, k - 1 - r]   (Z with its rows flipped), through a 32 x 32 tile
__global__ void rf_flip_transpose_kernel(const double* __restrict__ Zt, int64_t t, int64_t k, double* __restrict__ Zf) {
  __shared__ double tile[32][33];
  const int64_t i0 = int64_t(blockIdx.x) * 32, c0 = int64_t(blockIdx.y) * 32;   // i: row of Zt, c: column of Zt
  for (int a = threadIdx.y; a < 32; a += blockDim.y) {
    const int64_t i = i0 + threadIdx.x, c = c0 + a;
    tile[a][threadIdx.x] = (i < t && c < k) ? Zt[i + c * t] : 0.0;
  }
  __syncthreads();
  for (int a = threadIdx.y; a < 32; a += blockDim.y) {
    const int64_t c = c0 + threadIdx.x, i = i0 + a;
    if (i < t && c < k) Zf[(k - 1 - c) + i * k] = tile[threadIdx.x][a];
  }
}

// M (k x k, ld k) = Hk[p', p'] with p'[r] = perm[k - 1 - r]   (Hk symmetric, leading dimension ldh: read along one line)
__global__ void rf_gather_h_kernel(const double* __restrict__ H, int64_t ldh, const int64_t* __restrict__ perm,
                                   int64_t k, double* __restrict__ M) {
  const int64_t c = blockIdx.y;
  const int64_t pc = perm[k - 1 - c];
  for (int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < k; r += int64_t(gridDim.x) * blockDim.x)
    M[r + c * k] = H[pc * ldh + perm[k - 1 - r]];      // Hk is fully symmetric (both triangles written by the solver)
}

// In-place inverse of one lower-triangular diagonal block (jb <= 128) per CTA: thread c owns column c of the
// inverse (forward substitution against the identity), the block sits in shared memory.
struct RfLeaf {
  int off, len;
};
__global__ void __launch_bounds__(kRfLeaf)
rf_trtri_leaf_kernel(double* __restrict__ A, int64_t lda, const RfLeaf* __restrict__ leaves) {
  extern __shared__ double rf_sm[];        // L (len x len, ld len + 1) then X (same)
  const RfLeaf lf = leaves[blockIdx.x];
  const int len = lf.len, ld = len + 1;
  double* Ls = rf_sm;
  double* Xs = rf_sm + size_t(kRfLeaf) * (kRfLeaf + 1);
  double* Ab = A + lf.off + int64_t(lf.off) * lda;
  for (int idx = threadIdx.x; idx < len * len; idx += blockDim.x) {
    const int r = idx % len, c = idx / len;
    Ls[r + c * ld] = (r >= c) ? Ab[r + int64_t(c) * lda] : 0.0;
  }
  __syncthreads();
  const int c = threadIdx.x;
  if (c < len) {
    for (int r = c; r < len; ++r) {
      double s0 = (r == c) ? 1.0 : 0.0, s1 = 0.0;
      int q = c;
      for (; q + 1 < r; q += 2) {
        s0 = fma(-Ls[r + q * ld], Xs[q + c * ld], s0);
        s1 = fma(-Ls[r + (q + 1) * ld], Xs[(q + 1) + c * ld], s1);
      }
      if (q < r) s0 = fma(-Ls[r + q * ld], Xs[q + c * ld], s0);
      Xs[r + c * ld] = (s0 + s1) / Ls[r + r * ld];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < len * len; idx += blockDim.x) {
    const int r = idx % len, cc = idx / len;
    if (r >= cc) Ab[r + int64_t(cc) * lda] = Xs[r + cc * ld];
  }
}

struct RfMerge {
  int off, n1, len;
};
static void rf_collect(int off, int len, std::vector<RfLeaf>& leaves, std::vector<RfMerge>& merges) {
  if (len <= kRfLeaf) {
    leaves.push_back({off, len});
    return;
  }
  // split on a multiple of the leaf size so that every leaf but the last is full
  int n1 = ((len / 2 + kRfLeaf - 1) / kRfLeaf) * kRfLeaf;
  if (n1 >= len) n1 = len - kRfLeaf;
  rf_collect(off, n1, leaves, merges);
  rf_collect(off + n1, len - n1, leaves, merges);
  merges.push_back({off, n1, len});          // post-order: both halves are inverted when a merge runs
}

// In-place inverse of the lower-triangular A (k x k, ld lda): all diagonal leaves in one launch, then
// inv([A 0; B C]) = [A^-1 0; -C^-1 B A^-1  C^-1] bottom-up with two triangular products per merge (k^3 / 3 flop).
static int trtri_lower(cublasHandle_t h, cudaStream_t st, double* A, int64_t lda, int64_t k, RfLeaf* d_leaves) {
  std::vector<RfLeaf> leaves;
  std::vector<RfMerge> merges;
  rf_collect(0, int(k), leaves, merges);
  TQ_CUDA_CHECK(cudaMemcpyAsync(d_leaves, leaves.data(), sizeof(RfLeaf) * leaves.size(), cudaMemcpyHostToDevice, st));
  const size_t smem = size_t(2) * kRfLeaf * (kRfLeaf + 1) * sizeof(double);
  TQ_CUDA_CHECK(cudaFuncSetAttribute(rf_trtri_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  rf_trtri_leaf_kernel<<<unsigned(leaves.size()), kRfLeaf, smem, st>>>(A, lda, d_leaves);
  TQ_LAUNCH_CHECK();
  // (`leaves` is pageable: cudaMemcpyAsync returns once it has been staged, so it may go out of scope)
  const double one = 1.0, mone = -1.0;
  for (const RfMerge& mg : merges) {
    const int n2 = mg.len - mg.n1;
    double* Ai = A + mg.off + int64_t(mg.off) * lda;                       // A^-1 (n1 x n1)
    double* Ci = A + (mg.off + mg.n1) + int64_t(mg.off + mg.n1) * lda;     // C^-1 (n2 x n2)
    double* B = A + (mg.off + mg.n1) + int64_t(mg.off) * lda;              // n2 x n1
    TQ_CUBLAS_CHECK(cublasDtrmm(h, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n2,
                                mg.n1, &one, Ai, int(lda), B, int(lda), B, int(lda)));
    TQ_CUBLAS_CHECK(cublasDtrmm(h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n2,
                                mg.n1, &mone, Ci, int(lda), B, int(lda), B, int(lda)));
  }
  return TQ_OK;
}

// R (row-major k x n, ld ldr), columns [0, k): R[i, j] = Linv[k - 1 - i, k - 1 - j] for j >= i, 0 below the diagonal.
// 32 x 32 tiles: read Linv (column-major, ld k) along its columns, write R along its rows.
__global__ void rf_emit_t11_kernel(const double* __restrict__ Linv, int64_t k, double* __restrict__ R, int64_t ldr) {
  __shared__ double tile[32][33];
  const int64_t i0 = int64_t(blockIdx.y) * 32, j0 = int64_t(blockIdx.x) * 32;   // tile of R: rows i, columns j
  // source element for (i, j): Linv[(k-1-i) + (k-1-j) k]; read with threadIdx.x along i (contiguous in Linv)
  for (int a = threadIdx.y; a < 32; a += blockDim.y) {
    const int64_t i = i0 + threadIdx.x, j = j0 + a;
    tile[a][threadIdx.x] = (i < k && j < k && j >= i) ? Linv[(k - 1 - i) + (k - 1 - j) * k] : 0.0;
  }
  __syncthreads();
  for (int a = threadIdx.y; a < 32; a += blockDim.y) {
    const int64_t i = i0 + a, j = j0 + threadIdx.x;
    if (i < k && j < k) R[i * ldr + j] = tile[threadIdx.x][a];
  }
}

size_t rfactor_ws_bytes(int64_t n, int64_t k) {
  const int64_t t = n - k;
  return ws_bytes_for(size_t(k) * k, 8) + ws_bytes_for(size_t(k) * imax(t, 1), 8) * 3 + ws_bytes_for(size_t(t) * t + 1, 8) +
         ws_bytes_for(size_t(k / kRfLeaf + 2), sizeof(RfLeaf)) + ws_bytes_for(4, 4) + 4096;
}

// Hk: the truncated Hessian H_k = H - V_t diag(w_t) V_t^T (symmetric, leading dimension ldh) - the solver keeps the
// copy it made for the pivoted Cholesky; it may alias R (it is gathered before R is written).  Rx: row-major k x n
// (ld ldrx) from the pivoted Cholesky of H_k, perm its pivots.  R: row-major k x n (ld ldr), written completely.
int r_from_rx(cublasHandle_t h, cudaStream_t st, const double* Hk, int64_t ldh, int64_t n, int64_t k, const double* Rx,
              int64_t ldrx, const int64_t* perm, double* R, int64_t ldr, Workspace ws) {
  const int64_t t = n - k;
  double* M = ws.take<double>(size_t(k) * k);
  double* Zt = ws.take<double>(size_t(imax(t, 1)) * k);
  double* Zf = ws.take<double>(size_t(imax(t, 1)) * k);
  double* Q = ws.take<double>(size_t(imax(t, 1)) * k);
  double* S = ws.take<double>(size_t(t) * t + 1);
  RfLeaf* d_leaves = ws.take<RfLeaf>(size_t(k / kRfLeaf + 2));
  int* fail = ws.take<int>(4);
  if (ws.overflow) {
    set_error("r_from_rx: workspace too small");
    return TQ_ERR_WORKSPACE;
  }
  const double one = 1.0, zero = 0.0, half = 0.5;
  TQ_REQUIRE(ldrx == n && ldr == n, "r_from_rx: R and R_x must have leading dimension n");
  const double* L = Rx;                   // column-major n x k view of the row-major R_x: L = R_x^T
  double* Rt = R;                         // column-major n x k view of R
  dim3 gk((unsigned)imin(ceil_div(k, 256), 64), (unsigned)k);
  {
    StageTimer tm(st, "rfac: A11");
    rf_gather_h_kernel<<<gk, 256, 0, st>>>(Hk, ldh, perm, k, M);       // A11' = H_k[p', p']
    TQ_LAUNCH_CHECK();
  }
  if (t > 0) {
    StageTimer tm(st, "rfac: M");
    // Zt = L21 L11^-1  (t x k)
    dim3 gz((unsigned)imin(ceil_div(t, 256), 64), (unsigned)k);
    rf_copy_l21_kernel<<<gz, 256, 0, st>>>(L, n, k, t, Zt);
    TQ_LAUNCH_CHECK();
    TQ_CUBLAS_CHECK(cublasDtrsm(h, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, int(t),
                                int(k), &one, L, int(n), Zt, int(t)));
    dim3 gf((unsigned)ceil_div(t, 32), (unsigned)ceil_div(k, 32));
    rf_flip_transpose_kernel<<<gf, dim3(32, 8), 0, st>>>(Zt, t, k, Zf);
    TQ_LAUNCH_CHECK();
    // Q = A11' Zf;  S = Zf^T Q;  Q += Zf S / 2;  M' = A11' + Q Zf^T + Zf Q^T
    TQ_CUBLAS_CHECK(cublasDsymm(h, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, int(k), int(t), &one, M, int(k), Zf, int(k),
                                &zero, Q, int(k)));
    TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, int(t), int(t), int(k), &one, Zf, int(k), Q, int(k), &zero,
                                S, int(t)));
    TQ_CUBLAS_CHECK(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, int(k), int(t), int(t), &half, Zf, int(k), S, int(t), &one,
                                Q, int(k)));
    TQ_CUBLAS_CHECK(cublasDsyr2k(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, int(k), int(t), &one, Q, int(k), Zf, int(k),
                                 &one, M, int(k)));
  }
  {
    StageTimer tm(st, "rfac: chol");
    const int rc = chol_lower(h, st, M, k, fail);
    if (rc == TQ_ERR_NOCONV) set_error("r_from_rx: (I + Z Z^T) A11 (I + Z Z^T) is not positive definite");
    if (rc != TQ_OK) return rc;
  }
  {
    StageTimer tm(st, "rfac: trtri");
    TQ_TRY(trtri_lower(h, st, M, k, k, d_leaves));
  }
  {
    StageTimer tm(st, "rfac: emit");
    dim3 ge((unsigned)ceil_div(k, 32), (unsigned)ceil_div(k, 32));
    rf_emit_t11_kernel<<<ge, dim3(32, 8), 0, st>>>(M, k, R, ldr);
    TQ_LAUNCH_CHECK();
    if (t > 0)      // T12^T = Zt T11^T: rows [k, n) of the column-major view of R
      TQ_CUBLAS_CHECK(cublasDtrmm(h, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT,
                                  int(t), int(k), &one, Rt, int(n), Zt, int(t), Rt + k, int(n)));
  }
  return TQ_OK;
}

}  // namespace tq
