// Per-layer output reconstruction error  ||(W-Q)[:,perm] Rx^T||_F / ||W[:,perm] Rx^T||_F
// (reference log_quantization_error, gptq_utils.py:275-291): a logged diagnostic whose time is part of the
// reference's published per-Linear numbers.  The reference forms both m x k products with fp32 GEMMs and takes
// their Frobenius norms.  Here ONE tcgen05 GEMM of the stacked 2m x n operand [D; W_o] against Rx^T computes both,
// and the products never reach memory: the epilogue squares its fp32 accumulators and adds the tile's sum to the
// fp64 numerator (rows < m) or denominator (rows >= m).
//   * kind::tf32 on the fp32 operands as they are (the tensor core reads 10 mantissa bits): the ratio moves by
//     < 1e-4 relative against strict fp32 (bar 1 %) - the sum of 4.5e7 squared entries averages the rounding out;
//   * both operands K-major (D / W_o are m x n row-major, Rx is k x n row-major: no transposes), TMA boxes of
//     32 k x 128 / 256 rows with the 128-byte swizzle, 4-stage ring of 48 KB, M=128 N=256 K=8;
//   * 256 k per TMEM accumulation, chunks added in fp32 registers with round-to-nearest (two TMEM buffers), as in
//     trailing_tc.cu: the tensor core truncates when it accumulates.
//   * Rx is the R of a QR: rows [n0, n0 + 256) are zero left of column n0, so a tile's reduction starts there
//     (the cast kernel checks the structure and raises a flag that disables the skip);
// Roofline: per tile and k, (128 + 256) x 4 bytes for 2 x 128 x 256 flop = 43 flop per byte of L2 traffic.
// n % 4 != 0 (TMA needs 16-byte row pitches) falls back to the two cuBLAS GEMMs of round 1.
#include <cstdlib>
#include "blas.cuh"
#include "common.cuh"
#include "tcgen05.cuh"

namespace tq {

int get_cublas(cublasHandle_t* out, cudaStream_t stream) {
  static thread_local cublasHandle_t h = nullptr;
  static thread_local int h_dev = -1;
  int dev = 0;
  TQ_CUDA_CHECK(cudaGetDevice(&dev));
  if (h == nullptr || h_dev != dev) {
    TQ_CUBLAS_CHECK(cublasCreate(&h));
    TQ_CUBLAS_CHECK(cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH));
    h_dev = dev;
  }
  TQ_CUBLAS_CHECK(cublasSetStream(h, stream));
  *out = h;
  return TQ_OK;
}

// also raises *below when an entry left of the diagonal is not zero (Rx is then not upper trapezoidal and the
// GEMM must not skip the k range left of a tile's first row)
template <typename TR>
__global__ void cast_rx_kernel(const TR* __restrict__ R, int64_t ldr, int64_t k, int64_t n,
                               float* __restrict__ out, int* __restrict__ below) {
  int64_t r = blockIdx.y;
  for (int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; c < n; c += int64_t(gridDim.x) * blockDim.x) {
    const float v = float(R[r * ldr + c]);
    out[r * n + c] = v;
    if (c < r && v != 0.f) *below = 1;
  }
}

int make_tmap_2d(CUtensorMap* tmap, const void* base, int dtype, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);

constexpr int kMtM = 128, kMtN = 256, kMtKStage = 32;
constexpr int kMtStages = 4;
constexpr int kMtBand = 12;                               // tile columns per band of the rasterisation
constexpr int kMtChunkStages = 8;                        // 256 k per TMEM accumulation
constexpr int kMtABytes = kMtM * kMtKStage * 4;          // 16 KB
constexpr int kMtBBytes = kMtN * kMtKStage * 4;          // 32 KB
constexpr int kMtStageBytes = kMtABytes + kMtBBytes;     // 48 KB
constexpr int kMtCtrlWarps = 4, kMtEpiWarps = 8;
constexpr int kMtThreads = (kMtCtrlWarps + kMtEpiWarps) * 32;
constexpr size_t kMtSmem = size_t(kMtStages) * kMtStageBytes + 1024 + 256;

struct MtBarriers {
  uint64_t full[kMtStages];
  uint64_t empty[kMtStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// out2[0] += sum over rows < m_split of Y^2, out2[1] += the same over rows >= m_split, Y = A (rows x n) . B^T (k x n)
__global__ void __launch_bounds__(kMtThreads, 1)
metric_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 int64_t rows, int64_t m_split, int num_kstages, uint32_t idesc, double* __restrict__ out2,
                 const int* __restrict__ below, int tiles_x, int tiles_y) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  MtBarriers* bars = reinterpret_cast<MtBarriers*>(smem + size_t(kMtStages) * kMtStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Tile order: bands of kMtBand tile columns (rows of Rx), inside a band x fastest, then y.  The ~148 CTAs in
  // flight then cover a 12 x 12 patch: 12 tiles of each operand are shared through L2.  With the plain x-fastest
  // order a wave covered 44 x 3 tiles and every wave streamed the whole Rx again: ncu measured 20 GB of DRAM
  // reads for 0.7 GB of operands (79 % DRAM busy - the kernel was HBM-bound on re-reads).
  const int lin = int(blockIdx.x);
  const int band = lin / (kMtBand * tiles_y);
  const int in_band = lin - band * (kMtBand * tiles_y);
  const int wb = min(kMtBand, tiles_x - band * kMtBand);
  const int ty = in_band / wb, tx = band * kMtBand + (in_band - ty * wb);
  const int64_t n0 = int64_t(tx) * kMtN, m0 = int64_t(ty) * kMtM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMtStages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&bars->tmem_full[b], 1);
      ptx::mbar_init(&bars->tmem_empty[b], kMtEpiWarps);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&map_a);
    ptx::prefetch_tmap(&map_b);
  }
  if (warp == 1) {
    ptx::tmem_alloc(&bars->tmem_base, 2 * kMtN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  // rows [n0, n0 + 256) of an upper-trapezoidal Rx are zero left of column n0: start the reduction there
  // (k ~ 0.9 n: 45 % of the k stages of the grid)
  const int ks_first = (*below == 0) ? int(n0 / kMtKStage) : 0;
  const int num_chunks = (num_kstages - ks_first + kMtChunkStages - 1) / kMtChunkStages;

  if (warp >= kMtCtrlWarps) {
    ptx::setmaxnreg_inc<224>();
    const int quarter = warp & 3;
    const int half = (warp - kMtCtrlWarps) >> 2;
    float acc[128];
#pragma unroll
    for (int c = 0; c < 128; ++c) acc[c] = 0.f;
    for (int ch = 0; ch < num_chunks; ++ch) {
      const int buf = ch & 1;
      ptx::mbar_wait(&bars->tmem_full[buf], (ch >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(buf * kMtN + half * 128);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr + g * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) acc[g * 32 + q] = __fadd_rn(acc[g * 32 + q], __uint_as_float(v[q]));
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[buf]);
    }
    // rows beyond `rows` and columns beyond k were zero-filled by TMA: they add 0
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 128; c += 8) {        // 8 squares in fp32, the running sum in fp64
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) t = fmaf(acc[c + q], acc[c + q], t);
      s += double(t);
    }
    const int64_t row = m0 + quarter * 32 + lane;
    double s_num = row < m_split ? s : 0.0, s_den = row < m_split ? 0.0 : s;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s_num += __shfl_xor_sync(0xffffffffu, s_num, o);
      s_den += __shfl_xor_sync(0xffffffffu, s_den, o);
    }
    if (lane == 0) {
      if (s_num != 0.0) atomicAdd(out2, s_num);
      if (s_den != 0.0) atomicAdd(out2 + 1, s_den);
    }
  } else {
    ptx::setmaxnreg_dec<56>();
    if (warp == 0) {
      if (lane == 0) {
        for (int ks = ks_first; ks < num_kstages; ++ks) {
          const int s = (ks - ks_first) % kMtStages;
          const uint32_t ph = ((ks - ks_first) / kMtStages) & 1;
          ptx::mbar_wait(&bars->empty[s], ph ^ 1);
          ptx::mbar_expect_tx(&bars->full[s], kMtStageBytes);
          uint8_t* st = smem + size_t(s) * kMtStageBytes;
          ptx::tma_load_2d(st, &map_a, &bars->full[s], ks * kMtKStage, int(m0));
          ptx::tma_load_2d(st + kMtABytes, &map_b, &bars->full[s], ks * kMtKStage, int(n0));
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        int ks = ks_first;
        for (int ch = 0; ch < num_chunks; ++ch) {
          const int buf = ch & 1;
          ptx::mbar_wait(&bars->tmem_empty[buf], ((ch >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + uint32_t(buf) * kMtN;
          const int ks_end = min(num_kstages, ks + kMtChunkStages);
          bool first = true;
          for (; ks < ks_end; ++ks) {
            const int s = (ks - ks_first) % kMtStages;
            const uint32_t ph = ((ks - ks_first) / kMtStages) & 1;
            ptx::mbar_wait(&bars->full[s], ph);
            ptx::tc_fence_after();
            const uint32_t a_addr = ptx::smem_u32(smem + size_t(s) * kMtStageBytes);
            const uint32_t b_addr = a_addr + kMtABytes;
#pragma unroll
            for (int k = 0; k < kMtKStage / 8; ++k) {
              const uint64_t da = ptx::make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
              const uint64_t db = ptx::make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
              ptx::mma_tf32_ss(tmem_d, da, db, idesc, first ? 0u : 1u);
              first = false;
            }
            ptx::tc_commit(&bars->empty[s]);
          }
          ptx::tc_commit(&bars->tmem_full[buf]);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * kMtN);
  }
}

__global__ void gather_diff_kernel(const float* __restrict__ W, int64_t ldw, const float* __restrict__ Wq,
                                   int64_t ldq, const int64_t* __restrict__ perm, int64_t m, int64_t n,
                                   float* __restrict__ Wo, float* __restrict__ D) {
  for (int64_t r = blockIdx.y; r < m; r += gridDim.y)
    for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += int64_t(gridDim.x) * blockDim.x) {
      int64_t p = perm[j];
      float w = W[r * ldw + p];
      Wo[r * n + j] = w;
      D[r * n + j] = __fsub_rn(w, Wq[r * ldq + p]);
    }
}

__global__ void sumsq_kernel(const float* __restrict__ x, int64_t count, double* __restrict__ out) {
  double s = 0.0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
    double v = x[i];
    s += v * v;
  }
  __shared__ double sh[32];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

}  // namespace tq

using namespace tq;

extern "C" int tq_quant_error_workspace(int64_t m, int64_t n, int64_t k, size_t* bytes) {
  TQ_REQUIRE(bytes && m > 0 && n > 0 && k > 0, "tq_quant_error_workspace: bad arguments");
  *bytes = ws_bytes_for(size_t(k) * n, 4) + ws_bytes_for(size_t(2) * m * n, 4) + ws_bytes_for(size_t(m) * k, 4) + 4096;
  return TQ_OK;
}

extern "C" int tq_quant_error(const float* W, int64_t ldw, const float* Wq, int64_t ldq, const void* Rx,
                              int rx_dtype, int64_t ldr, int64_t k, const int64_t* perm, int64_t m, int64_t n,
                              double* out2, void* ws, size_t ws_bytes, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(W && Wq && Rx && perm && out2, "tq_quant_error: null pointer");
  TQ_REQUIRE(m > 0 && n > 0 && k > 0 && k <= n && ldw >= n && ldq >= n && ldr >= n, "tq_quant_error: bad shape");
  TQ_REQUIRE(rx_dtype == TQ_F64 || rx_dtype == TQ_F32, "tq_quant_error: Rx must be fp64 or fp32");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace wsp(ws, ws_bytes);
  float* R32 = wsp.take<float>(size_t(k) * n);
  float* D = wsp.take<float>(size_t(2) * m * n);        // stacked [D; W_o]: rows [0, m) and [m, 2 m)
  float* Wo = D + size_t(m) * n;
  float* Y = wsp.take<float>(size_t(m) * k);            // cuBLAS fallback only
  int* below = wsp.take<int>(4);
  if (wsp.overflow) {
    set_error("tq_quant_error: workspace too small (%zu < %zu)", ws_bytes, wsp.off);
    return TQ_ERR_WORKSPACE;
  }
  {
    dim3 grid((unsigned)imin(ceil_div(n, 256), 64), (unsigned)k);
    TQ_CUDA_CHECK(cudaMemsetAsync(below, 0, sizeof(int), st));
    if (rx_dtype == TQ_F64) cast_rx_kernel<double><<<grid, 256, 0, st>>>((const double*)Rx, ldr, k, n, R32, below);
    else cast_rx_kernel<float><<<grid, 256, 0, st>>>((const float*)Rx, ldr, k, n, R32, below);
    TQ_LAUNCH_CHECK();
  }
  {
    dim3 grid((unsigned)imin(ceil_div(n, 256), 64), (unsigned)imin(m, 65535));
    gather_diff_kernel<<<grid, 256, 0, st>>>(W, ldw, Wq, ldq, perm, m, n, Wo, D);
    TQ_LAUNCH_CHECK();
  }
  TQ_CUDA_CHECK(cudaMemsetAsync(out2, 0, 2 * sizeof(double), st));
  if (n % 4 == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0 && (reinterpret_cast<uintptr_t>(R32) & 15) == 0) {
    CUtensorMap map_a, map_b;
    TQ_TRY(make_tmap_2d(&map_a, D, TQ_F32, uint64_t(n), uint64_t(2 * m), uint64_t(n) * 4, kMtKStage, kMtM));
    TQ_TRY(make_tmap_2d(&map_b, R32, TQ_F32, uint64_t(n), uint64_t(k), uint64_t(n) * 4, kMtKStage, kMtN));
    static thread_local bool attr_done[kMaxDevices] = {};
    if (!attr_done[device_slot()]) {
      TQ_CUDA_CHECK(cudaFuncSetAttribute(metric_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMtSmem));
      attr_done[device_slot()] = true;
    }
    const uint32_t idesc = ptx::make_idesc(/*TF32*/ 2u, /*A K-major*/ 0u, /*B K-major*/ 0u, kMtM, kMtN);
    const int tiles_x = int(ceil_div(k, kMtN)), tiles_y = int(ceil_div(2 * m, kMtM));
    dim3 grid((unsigned)(int64_t(tiles_x) * tiles_y));
    // algorithmic flops of the upper-trapezoidal case (the R of a QR): row i of Rx has n - i entries; a dense Rx
    // (flag raised by the cast) executes 2 (2m) k n and is under-reported by this figure
    const double work = 2.0 * double(2 * m) * (double(k) * double(n) - 0.5 * double(k) * double(k));
    const int pslot = prof_begin_launch(st, work, TQ_PROF_METRIC);
    metric_tc_kernel<<<grid, kMtThreads, kMtSmem, st>>>(map_a, map_b, 2 * m, m, int(ceil_div(n, kMtKStage)), idesc, out2, below, tiles_x,
                                                         tiles_y);
    prof_end_launch(st, pslot);
    TQ_LAUNCH_CHECK();
    return TQ_OK;
  }
  cublasHandle_t h;
  TQ_TRY(get_cublas(&h, st));
  const float one = 1.f, zero = 0.f;
  // row-major Y (m x k) = A (m x n) . R32^T  <=>  col-major Y^T (k x m) = R32 . A^T
  const float* srcs[2] = {D, Wo};
  const int strict = 0;
  for (int i = 0; i < 2; ++i) {
    if (strict) {
      TQ_CUBLAS_CHECK(cublasSgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, int(k), int(m), int(n), &one, R32, int(n), srcs[i],
                                  int(n), &zero, Y, int(k)));
    } else {
      cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);
      cublasStatus_t cs = cublasGemmEx(h, CUBLAS_OP_T, CUBLAS_OP_N, int(k), int(m), int(n), &one, R32, CUDA_R_32F,
                                       int(n), srcs[i], CUDA_R_32F, int(n), &zero, Y, CUDA_R_32F, int(k),
                                       CUBLAS_COMPUTE_32F_FAST_TF32, CUBLAS_GEMM_DEFAULT);
      cublasSetMathMode(h, CUBLAS_PEDANTIC_MATH);
      if (cs != CUBLAS_STATUS_SUCCESS) {
        set_error("tq_quant_error: cublasGemmEx failed with status %d", int(cs));
        return TQ_ERR_CUDA;
      }
    }
    sumsq_kernel<<<296, 256, 0, st>>>(Y, m * k, out2 + i);
    TQ_LAUNCH_CHECK();
  }
  return TQ_OK;
}
