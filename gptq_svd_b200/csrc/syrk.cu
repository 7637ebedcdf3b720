// Hessian accumulation H += X^T X as a tcgen05 / TMEM SYRK with TMA-staged activation
// tiles (sm_100a).  Replaces HessianAccumulator.add_batch / get_hessian
// (reference gptq_utils.py:218-228), which runs a full fp64 GEMM on fp16 data.
//
// Mapping.  X is rows x n row-major (tokens x features), so both UMMA operands are
// "MN-major": the contraction index (token) is the slow one.  One CTA owns one
// 128 (i) x 256 (j) tile of X^T X and walks all tokens of the batch:
//   warp 0      TMA producer: per 64-token K-block, 2 + 4 boxes of [64 tokens x 64 features]
//               (128-byte swizzle) into a 4-stage ring (48 KB / stage)
//   warp 1      tcgen05.mma issuer (one elected lane), M=128 N=256 K=16, fp32 accumulators
//               in TMEM, two 256-column buffers alternating every `kc` K-blocks
//   warps 4..11 epilogue (two warpgroups, registers raised with setmaxnreg): drain a finished TMEM buffer with tcgen05.ld, add it into fp32
//               registers (round-to-nearest across chunks), and after the last chunk add
//               the batch total into the fp64 H
// Only tiles that touch j >= i are computed (T n^2 flop instead of 2 T n^2); the tile is
// stored TRANSPOSED (H[j][i], coalesced along i) and a mirror pass fills the other
// triangle, so H is fully symmetric when the call returns, like the reference's .H.
// fp16 x fp16 products are exact in fp32; the only rounding is the fp32 accumulation
// inside a kc-token chunk, the fp32 chunk sums and one fp64 add per batch.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace tq {

constexpr int kTileI = 128;   // UMMA M
constexpr int kTileJ = 256;   // UMMA N
constexpr int kTokBlk = 64;   // tokens per pipeline stage
constexpr int kStages = 4;
constexpr int kChunkBytes = kTokBlk * 128;                                   // one TMA box: 8 KB
constexpr int kStageBytes = (kTileI / 64 + kTileJ / 64) * kChunkBytes;       // 48 KB
constexpr int kEpiWarps = 8;
constexpr int kCtrlWarps = 4;  // warpgroup 0: TMA producer, MMA issuer, two idle warps (setmaxnreg is per warpgroup)
constexpr int kSyrkThreads = (kCtrlWarps + kEpiWarps) * 32;
constexpr size_t kSyrkSmem = size_t(kStages) * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;

struct SyrkBarriers {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void epilogue_role(SyrkBarriers* bars, uint32_t tmem_base, int num_chunks, int warp,
                                              int lane, int i0, int j0, int n, double* __restrict__ H,
                                              int64_t ldh, double alpha) {
  const int ew = warp - kCtrlWarps;
  const int quarter = warp & 3;   // TMEM lane quarter this warp may access
  const int half = ew >> 2;       // column half of the 256-wide tile
  float acc[128];
#pragma unroll
  for (int c = 0; c < 128; ++c) acc[c] = 0.f;
  for (int ch = 0; ch < num_chunks; ++ch) {
    const int buf = ch & 1;
    ptx::mbar_wait(&bars->tmem_full[buf], (ch >> 1) & 1);
    ptx::tc_fence_after();
    const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(buf * kTileJ + half * 128);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(taddr + g * 32, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) acc[g * 32 + q] += __uint_as_float(v[q]);
    }
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&bars->tmem_empty[buf]);
  }
  const int i = i0 + quarter * 32 + lane;
  if (i < n) {
#pragma unroll
    for (int c = 0; c < 128; ++c) {
      const int j = j0 + half * 128 + c;
      if (j < n) {
        double* p = H + int64_t(j) * ldh + i;
        *p += alpha * double(acc[c]);
      }
    }
  }
}

__global__ void __launch_bounds__(kSyrkThreads, 1)
syrk_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap, double* __restrict__ H, int64_t ldh, int n,
                    int num_kblocks, int kc_blocks, int n_iblk, uint32_t idesc, const double* __restrict__ alpha_dev) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  SyrkBarriers* bars = reinterpret_cast<SyrkBarriers*>(smem + size_t(kStages) * kStageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile decode: tiles are enumerated by j-block, each with min(2*bj+2, n_iblk) i-blocks
  int bj = 0, bi = 0;
  {
    int t = blockIdx.x;
    for (;; ++bj) {
      int c = min(2 * bj + 2, n_iblk);
      if (t < c) {
        bi = t;
        break;
      }
      t -= c;
    }
  }
  const int i0 = bi * kTileI, j0 = bj * kTileJ;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&bars->tmem_full[b], 1);
      ptx::mbar_init(&bars->tmem_empty[b], kEpiWarps);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmap);
  }
  if (warp == 1) {
    ptx::tmem_alloc(&bars->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int num_chunks = (num_kblocks + kc_blocks - 1) / kc_blocks;

  if (warp >= kCtrlWarps) {
    // ------------------------------------------------ epilogue (8 warps)
    ptx::setmaxnreg_inc<224>();
    epilogue_role(bars, tmem_base, num_chunks, warp, lane, i0, j0, n, H, ldh, alpha_dev ? *alpha_dev : 1.0);
  } else {
  ptx::setmaxnreg_dec<56>();
  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int kb = 0; kb < num_kblocks; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        ptx::mbar_wait(&bars->empty[s], ph ^ 1);
        ptx::mbar_expect_tx(&bars->full[s], kStageBytes);
        uint8_t* st = smem + size_t(s) * kStageBytes;
#pragma unroll
        for (int c = 0; c < kTileI / 64; ++c)
          ptx::tma_load_2d(st + c * kChunkBytes, &tmap, &bars->full[s], i0 + 64 * c, kb * kTokBlk);
#pragma unroll
        for (int c = 0; c < kTileJ / 64; ++c)
          ptx::tma_load_2d(st + (kTileI / 64 + c) * kChunkBytes, &tmap, &bars->full[s], j0 + 64 * c,
                           kb * kTokBlk);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      int kb = 0;
      for (int ch = 0; ch < num_chunks; ++ch) {
        const int buf = ch & 1;
        ptx::mbar_wait(&bars->tmem_empty[buf], ((ch >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(buf) * kTileJ;
        const int kb_end = min(num_kblocks, kb + kc_blocks);
        bool first = true;
        for (; kb < kb_end; ++kb) {
          const int s = kb % kStages;
          const uint32_t ph = (kb / kStages) & 1;
          ptx::mbar_wait(&bars->full[s], ph);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + size_t(s) * kStageBytes);
          const uint32_t b_addr = a_addr + (kTileI / 64) * kChunkBytes;
#pragma unroll
          for (int k = 0; k < kTokBlk / 16; ++k) {
            // 16 tokens = 16 rows of 128 B = two 8-row swizzle atoms further down
            const uint64_t da = ptx::make_smem_desc_sw128(a_addr + k * 2048, kChunkBytes, 1024);
            const uint64_t db = ptx::make_smem_desc_sw128(b_addr + k * 2048, kChunkBytes, 1024);
            ptx::mma_f16_ss(tmem_d, da, db, idesc, first ? 0u : 1u);
            first = false;
          }
          ptx::tc_commit(&bars->empty[s]);   // frees the smem stage when these MMAs retire
        }
        ptx::tc_commit(&bars->tmem_full[buf]);  // accumulator chunk complete
      }
    }
  }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// H[r][c] = H[c][r] for r < c (the SYRK writes the r >= c triangle).
__global__ void mirror_lower_to_upper_kernel(double* __restrict__ H, int64_t ldh, int64_t n) {
  __shared__ double t[32][33];
  int bx = blockIdx.x, by = blockIdx.y;  // tile (by: row block, bx: col block) of the UPPER part
  if (bx < by) return;
  int64_t r = int64_t(bx) * 32 + threadIdx.y, c = int64_t(by) * 32 + threadIdx.x;  // source (lower)
  for (int k = 0; k < 32; k += 8) {
    int64_t rr = r + k;
    t[threadIdx.y + k][threadIdx.x] = (rr < n && c < n) ? H[rr * ldh + c] : 0.0;
  }
  __syncthreads();
  int64_t orow = int64_t(by) * 32 + threadIdx.y, ocol = int64_t(bx) * 32 + threadIdx.x;
  for (int k = 0; k < 32; k += 8) {
    int64_t rr = orow + k;
    if (rr < n && ocol < n && rr < ocol) H[rr * ldh + ocol] = t[threadIdx.x][threadIdx.y + k];
  }
}

__global__ void scale_matrix_kernel(const double* __restrict__ H, int64_t ldh, int64_t n, double inv,
                                    int divide, double denom, double* __restrict__ out, int64_t ldo) {
  int64_t r = blockIdx.y;
  for (int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; c < n; c += int64_t(gridDim.x) * blockDim.x) {
    double v = H[r * ldh + c];
    out[r * ldo + c] = divide ? v / denom : v;
  }
  (void)inv;
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(double v) { return float(v); }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }

// dst = fp16(src * scale), saturating at +-65504 (NaN stays NaN); scale_dev == nullptr: scale = 1
template <typename T>
__global__ void cast_to_f16_kernel(const T* __restrict__ src, int64_t lds, int64_t rows, int64_t n,
                                   __half* __restrict__ dst, int64_t ldd, const float* __restrict__ scale_dev) {
  const int64_t total = rows * n;
  const float sc = scale_dev ? *scale_dev : 1.0f;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    int64_t r = idx / n, c = idx - r * n;
    float v = to_f32(src[r * lds + c]) * sc;
    v = fminf(fmaxf(v, -65504.f), 65504.f);      // fminf / fmaxf return the non-NaN operand: NaN -> flagged by amax
    dst[r * ldd + c] = __float2half_rn(v);
  }
}

// amax_bits = max over the batch of |x| as an ordered unsigned (non-negative floats compare like their bit
// patterns); a NaN or an infinity anywhere sets status |= 1.
template <typename T>
__global__ void amax_kernel(const T* __restrict__ src, int64_t lds, int64_t rows, int64_t n,
                            unsigned int* __restrict__ amax_bits, int* __restrict__ status) {
  const int64_t total = rows * n;
  float m = 0.f;
  bool bad = false;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    int64_t r = idx / n, c = idx - r * n;
    const float v = fabsf(to_f32(src[r * lds + c]));
    if (!(v <= 3.0e38f)) bad = true;
    else m = fmaxf(m, v);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  bad = __any_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(amax_bits, __float_as_uint(m));
    if (bad) atomicOr(status, 1);
  }
}

// scale = 2^e with amax * scale in [2^13, 2^14) (fp16 keeps its full 11-bit significand for the largest
// entries and |x| * scale never reaches the fp16 overflow threshold); alpha = 1 / scale^2 (exact).
__global__ void pick_scale_kernel(const unsigned int* __restrict__ amax_bits, float* __restrict__ scale,
                                  double* __restrict__ alpha) {
  const float m = __uint_as_float(*amax_bits);
  int e = 0;
  if (m > 0.f) {
    int ex;
    frexpf(m, &ex);          // m = f * 2^ex, f in [0.5, 1)
    e = 14 - ex;
  }
  e = max(-100, min(100, e));
  *scale = ldexpf(1.0f, e);
  *alpha = ldexp(1.0, -2 * e);
}

// Probe of the accumulated Hessian (guard against a corrupted accumulation): with the fixed sign vector
// v_j = +-1 (hash of j), v^T (X^T X) v = ||X v||^2.  probe += alpha * sum over rows of (x_row . v)^2.
__device__ __forceinline__ float probe_sign(int64_t j) {
  uint32_t h = uint32_t(j) * 2654435761u;
  h ^= h >> 15;
  h *= 2246822519u;
  h ^= h >> 13;
  return (h & 0x10000u) ? 1.f : -1.f;
}

template <typename T>
__global__ void __launch_bounds__(256)
probe_rows_kernel(const T* __restrict__ X, int64_t ldx, int64_t rows, int64_t n, const double* __restrict__ alpha_dev,
                  double* __restrict__ probe) {
  // one warp per token row, 16-byte loads (8 values per lane): ldx % 8 == 0 and a 16-byte aligned X are
  // preconditions of the SYRK this probe accompanies
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warps = int64_t(gridDim.x) * 8;
  const int64_t n8 = n / 8;
  double local = 0.0;
  for (int64_t r = int64_t(blockIdx.x) * 8 + wid; r < rows; r += warps) {
    const T* xr = X + r * ldx;
    const uint4* x4 = reinterpret_cast<const uint4*>(xr);
    float s = 0.f;
    for (int64_t q = lane; q < n8; q += 32) {
      const uint4 v = x4[q];
      const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += to_f32(e[i]) * probe_sign(q * 8 + i);
    }
    for (int64_t j = n8 * 8 + lane; j < n; j += 32) s += to_f32(xr[j]) * probe_sign(j);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    local += double(s) * double(s);
  }
  __shared__ double sh[8];
  if (lane == 0) sh[wid] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(probe, t * (alpha_dev ? *alpha_dev : 1.0));
  }
}

// vhv += sum_r v_r * (H[r,:] . v) for the rows of this block
__global__ void __launch_bounds__(256)
probe_h_kernel(const double* __restrict__ H, int64_t ldh, int64_t n, double* __restrict__ vhv) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warps = int64_t(gridDim.x) * 8;
  double local = 0.0;
  for (int64_t r = int64_t(blockIdx.x) * 8 + wid; r < n; r += warps) {
    const double* hr = H + r * ldh;
    double s = 0.0;
    for (int64_t j = lane; j < n; j += 32) s += hr[j] * double(probe_sign(j));
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    local += s * double(probe_sign(r));
  }
  __shared__ double sh[8];
  if (lane == 0) sh[wid] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(vhv, t);
  }
}

__global__ void probe_compare_kernel(const double* __restrict__ vhv, const double* __restrict__ probe, double tol,
                                     int* __restrict__ status) {
  const double a = *vhv, b = *probe;
  const double scale = fmax(fabs(a), fabs(b));
  if (!(fabs(a - b) <= tol * scale)) atomicOr(status, 2);     // also true for NaN
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_2d(CUtensorMap* tmap, const void* base, int dtype, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return TQ_ERR_CUDA;
  }
  CUtensorMapDataType dt = dtype == TQ_F16    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                           : dtype == TQ_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                              : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tmap, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu stride=%llu)", int(r),
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes);
    return TQ_ERR_CUDA;
  }
  return TQ_OK;
}

// fp64 column-major matrix (rows contiguous), dense box of box_rows x box_cols, no swizzle,
// out-of-bounds elements read as zero (used by the symmetric SYMV of the tridiagonal reduction)
int make_tmap_f64(CUtensorMap* tmap, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                  uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return TQ_ERR_CUDA;
  }
  cuuint64_t dims[2] = {rows, cols};
  cuuint64_t strides[1] = {ld_elems * 8};
  cuuint32_t box[2] = {box_rows, box_cols};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (f64) failed with CUresult %d (rows=%llu cols=%llu ld=%llu)", int(r),
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems);
    return TQ_ERR_CUDA;
  }
  return TQ_OK;
}

}  // namespace tq

using namespace tq;

extern "C" int tq_syrk_accum_scaled(double* H, int64_t ldh, const void* X, int x_dtype, int64_t rows, int64_t n,
                                    int64_t ldx, int kc_tokens, const double* alpha_dev, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(H && (X || rows == 0), "tq_syrk_accum: null pointer");
  TQ_REQUIRE(n > 0 && rows >= 0 && ldh >= n && ldx >= n, "tq_syrk_accum: bad shape rows=%lld n=%lld",
             (long long)rows, (long long)n);
  TQ_REQUIRE(x_dtype == TQ_F16 || x_dtype == TQ_BF16,
             "tq_syrk_accum: X must be fp16 or bf16 (cast other types with tq_cast_to_f16_scaled)");
  TQ_REQUIRE(ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0,
             "tq_syrk_accum: TMA needs a 16-byte aligned X and ldx %% 8 == 0 (ldx=%lld)", (long long)ldx);
  TQ_REQUIRE(n < (1 << 30) && rows < (int64_t(1) << 31), "tq_syrk_accum: shape too large");
  if (rows == 0) return TQ_OK;
  if (kc_tokens <= 0) kc_tokens = 256;
  int kc_blocks = max(1, kc_tokens / kTokBlk);
  cudaStream_t st = (cudaStream_t)stream;

  CUtensorMap tmap;
  TQ_TRY(make_tmap_2d(&tmap, X, x_dtype, uint64_t(n), uint64_t(rows), uint64_t(ldx) * 2, 64, kTokBlk));

  // cudaFuncSetAttribute is per function and per device: cheap enough to call every time
  TQ_CUDA_CHECK(cudaFuncSetAttribute(syrk_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kSyrkSmem));
  const int n_iblk = int(ceil_div(n, kTileI)), n_jblk = int(ceil_div(n, kTileJ));
  int64_t tiles = 0;
  for (int bj = 0; bj < n_jblk; ++bj) tiles += min(2 * bj + 2, n_iblk);
  const int num_kblocks = int(ceil_div(rows, kTokBlk));
  const uint32_t idesc = ptx::make_idesc(x_dtype == TQ_BF16 ? 1u : 0u, 1u, 1u, kTileI, kTileJ);
  const int pslot = prof_begin_launch(st, double(rows) * double(n) * double(n + 1), TQ_PROF_SYRK);
  syrk_tcgen05_kernel<<<(unsigned)tiles, kSyrkThreads, kSyrkSmem, st>>>(tmap, H, ldh, int(n), num_kblocks,
                                                                        kc_blocks, n_iblk, idesc, alpha_dev);
  prof_end_launch(st, pslot);
  TQ_LAUNCH_CHECK();
  dim3 g((unsigned)ceil_div(n, 32), (unsigned)ceil_div(n, 32));
  mirror_lower_to_upper_kernel<<<g, dim3(32, 8), 0, st>>>(H, ldh, n);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

extern "C" int tq_syrk_accum(double* H, int64_t ldh, const void* X, int x_dtype, int64_t rows, int64_t n,
                             int64_t ldx, int kc_tokens, void* stream) {
  return tq_syrk_accum_scaled(H, ldh, X, x_dtype, rows, n, ldx, kc_tokens, nullptr, stream);
}

template <typename T>
static int probe_rows_launch(const void* X, int64_t ldx, int64_t rows, int64_t n, const double* alpha_dev,
                             double* probe, cudaStream_t st) {
  const unsigned grid = (unsigned)imin(ceil_div(rows, 8), 148 * 8);
  probe_rows_kernel<T><<<grid, 256, 0, st>>>((const T*)X, ldx, rows, n, alpha_dev, probe);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

extern "C" int tq_hessian_probe_accum(const void* X, int x_dtype, int64_t rows, int64_t n, int64_t ldx,
                                      const double* alpha_dev, double* probe_dev, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(probe_dev && (X || rows == 0) && n > 0 && rows >= 0 && ldx >= n, "tq_hessian_probe_accum: bad arguments");
  TQ_REQUIRE(ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0,
             "tq_hessian_probe_accum: needs a 16-byte aligned X and ldx %% 8 == 0 like tq_syrk_accum");
  if (rows == 0) return TQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == TQ_F16) return probe_rows_launch<__half>(X, ldx, rows, n, alpha_dev, probe_dev, st);
  if (x_dtype == TQ_BF16) return probe_rows_launch<__nv_bfloat16>(X, ldx, rows, n, alpha_dev, probe_dev, st);
  set_error("tq_hessian_probe_accum: X must be fp16 or bf16 (the tensor the SYRK consumed)");
  return TQ_ERR_INVALID;
}

extern "C" int tq_hessian_probe_check(const double* H, int64_t ldh, int64_t n, const double* probe_dev, double tol,
                                      double* vhv_dev, int* status_dev, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(H && probe_dev && vhv_dev && status_dev && n > 0 && ldh >= n && tol > 0,
             "tq_hessian_probe_check: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  TQ_CUDA_CHECK(cudaMemsetAsync(vhv_dev, 0, sizeof(double), st));
  probe_h_kernel<<<(unsigned)imin(ceil_div(n, 8), 148 * 8), 256, 0, st>>>(H, ldh, n, vhv_dev);
  TQ_LAUNCH_CHECK();
  probe_compare_kernel<<<1, 1, 0, st>>>(vhv_dev, probe_dev, tol, status_dev);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

extern "C" int tq_hessian_scale(const double* H, int64_t ldh, int64_t n, int64_t n_samples, double* out,
                                int64_t ldo, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(H && out && n > 0 && ldh >= n && ldo >= n && n_samples >= 0, "tq_hessian_scale: bad arguments");
  dim3 grid((unsigned)imin(ceil_div(n, 256), 64), (unsigned)n);
  scale_matrix_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(H, ldh, n, 0.0, n_samples > 0 ? 1 : 0,
                                                              double(n_samples), out, ldo);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

template <typename T>
static int cast_launch(const void* src, int64_t lds, int64_t rows, int64_t n, void* dst, int64_t ldd,
                       const float* scale_dev, unsigned int* amax_bits, int* status_dev, cudaStream_t st) {
  const unsigned grid = (unsigned)imin(ceil_div(rows * n, 256), 148 * 16);
  if (amax_bits) {
    amax_kernel<T><<<grid, 256, 0, st>>>((const T*)src, lds, rows, n, amax_bits, status_dev);
    TQ_LAUNCH_CHECK();
    pick_scale_kernel<<<1, 1, 0, st>>>(amax_bits, const_cast<float*>(scale_dev),
                                       reinterpret_cast<double*>(amax_bits + 4));
    TQ_LAUNCH_CHECK();
  }
  cast_to_f16_kernel<T><<<grid, 256, 0, st>>>((const T*)src, lds, rows, n, (__half*)dst, ldd, scale_dev);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

static int cast_dispatch(const void* src, int src_dtype, int64_t lds, int64_t rows, int64_t n, void* dst, int64_t ldd,
                         const float* scale_dev, unsigned int* amax_bits, int* status_dev, cudaStream_t st) {
  if (src_dtype == TQ_F32) return cast_launch<float>(src, lds, rows, n, dst, ldd, scale_dev, amax_bits, status_dev, st);
  if (src_dtype == TQ_F64) return cast_launch<double>(src, lds, rows, n, dst, ldd, scale_dev, amax_bits, status_dev, st);
  if (src_dtype == TQ_BF16)
    return cast_launch<__nv_bfloat16>(src, lds, rows, n, dst, ldd, scale_dev, amax_bits, status_dev, st);
  if (src_dtype == TQ_F16) return cast_launch<__half>(src, lds, rows, n, dst, ldd, scale_dev, amax_bits, status_dev, st);
  set_error("tq_cast_to_f16: unsupported source dtype %d", src_dtype);
  return TQ_ERR_INVALID;
}

extern "C" int tq_cast_to_f16(const void* src, int src_dtype, int64_t rows, int64_t n, int64_t lds, void* dst,
                              int64_t ldd, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(src && dst && rows > 0 && n > 0 && lds >= n && ldd >= n, "tq_cast_to_f16: bad arguments");
  return cast_dispatch(src, src_dtype, lds, rows, n, dst, ldd, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int tq_cast_to_f16_scaled(const void* src, int src_dtype, int64_t rows, int64_t n, int64_t lds, void* dst,
                                     int64_t ldd, void* scratch32_dev, int* status_dev, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(src && dst && scratch32_dev && status_dev && rows > 0 && n > 0 && lds >= n && ldd >= n,
             "tq_cast_to_f16_scaled: bad arguments");
  TQ_REQUIRE((reinterpret_cast<uintptr_t>(scratch32_dev) & 15) == 0, "tq_cast_to_f16_scaled: scratch must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  // scratch layout (32 bytes): [0] amax bits (u32)  [4] scale (f32)  [16] alpha = 1 / scale^2 (f64)
  unsigned int* amax_bits = reinterpret_cast<unsigned int*>(scratch32_dev);
  TQ_CUDA_CHECK(cudaMemsetAsync(scratch32_dev, 0, 32, st));
  return cast_dispatch(src, src_dtype, lds, rows, n, dst, ldd, reinterpret_cast<float*>(amax_bits + 1), amax_bits,
                       status_dev, st);
}
