// Lazy-batch trailing update of the GPTQ loop as a tcgen05 GEMM (sm_100a):
//     C[m x N] -= E[m x K] . U[K x N],   K = 128 inside a macro block, K <= 1024 beyond it,
// replaces the cuBLAS SGEMM + two ATen passes of the reference (gptq_utils.py:539-545).
//
// The reference runs this product in strict fp32 (TF32 off, :474-475); a single TF32 pass
// fails the 99.9 % code-parity bar by a wide margin, so both operands are
// split x = hi + lo with hi, lo representable in TF32 (round-to-nearest) and three MMAs
// accumulate hi.hi + hi.lo + lo.hi in the same fp32 TMEM accumulator ("3xTF32").  The
// splits are produced by the kernels that write E (gptq_block_kernel) and U (prep_u_kernel).
//
// One CTA per 128 x 256 tile of C:
//   warp 0   TMA producer: per 32-deep K stage, E_hi / E_lo boxes [128 rows x 32 k] and
//            U^T_hi / U^T_lo boxes [256 n x 32 k] (both K-major; U is stored transposed by
//            split_transpose_u_kernel because N-major TF32 operands need the 32-byte-base
//            swizzle atom), 128-byte swizzle, 2-stage ring
//   warp 1   tcgen05.mma.kind::tf32 issuer, M=128 N=256 K=8, 3 MMAs per k-step
//   warps 2-5 epilogue: tcgen05.ld 32 x 32 sub-tiles -> shared-memory transpose -> coalesced
//            C -= acc (fp32 subtraction, one rounding, as `W[:, i2:] -= Global_delta`)
#include "common.cuh"
#include "tcgen05.cuh"

namespace tq {

int make_tmap_2d(CUtensorMap* tmap, const void* base, int dtype, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);

constexpr int kTcM = 128, kTcN = 256, kTcKStage = 32;
constexpr int kTcStages = 2;
constexpr int kTcABytes = kTcM * kTcKStage * 4;          // 16 KB (one of hi / lo)
constexpr int kTcBBytes = kTcN * kTcKStage * 4;          // 32 KB (one of hi / lo)
constexpr int kTcStageBytes = 2 * kTcABytes + 2 * kTcBBytes;   // 96 KB
constexpr int kTcThreads = 6 * 32;
constexpr size_t kTcSmem = size_t(kTcStages) * kTcStageBytes + 1024 + 256 + 4 * 32 * 33 * 4;

struct TcBarriers {
  uint64_t full[kTcStages];
  uint64_t empty[kTcStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kTcThreads, 1)
trailing_tc_kernel(const __grid_constant__ CUtensorMap map_ehi, const __grid_constant__ CUtensorMap map_elo,
                   const __grid_constant__ CUtensorMap map_uhi, const __grid_constant__ CUtensorMap map_ulo,
                   float* __restrict__ C, int64_t ldc, int64_t m, int64_t N, int e_col0 /*first E column*/,
                   int k_row0 /*first U row*/, int u_col0 /*first U column of this update*/, int num_kstages,
                   uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(smem + size_t(kTcStages) * kTcStageBytes);
  float* tr = reinterpret_cast<float*>(smem + size_t(kTcStages) * kTcStageBytes + 256);   // 4 warps x 32 x 33

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n0 = int64_t(blockIdx.x) * kTcN, m0 = int64_t(blockIdx.y) * kTcM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    ptx::mbar_init(&bars->tmem_full, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&map_ehi);
    ptx::prefetch_tmap(&map_elo);
    ptx::prefetch_tmap(&map_uhi);
    ptx::prefetch_tmap(&map_ulo);
  }
  if (warp == 1) {
    ptx::tmem_alloc(&bars->tmem_base, kTcN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      for (int ks = 0; ks < num_kstages; ++ks) {
        const int s = ks % kTcStages;
        const uint32_t ph = (ks / kTcStages) & 1;
        ptx::mbar_wait(&bars->empty[s], ph ^ 1);
        ptx::mbar_expect_tx(&bars->full[s], kTcStageBytes);
        uint8_t* st = smem + size_t(s) * kTcStageBytes;
        ptx::tma_load_2d(st, &map_ehi, &bars->full[s], e_col0 + ks * kTcKStage, int(m0));
        ptx::tma_load_2d(st + kTcABytes, &map_elo, &bars->full[s], e_col0 + ks * kTcKStage, int(m0));
        uint8_t* bh = st + 2 * kTcABytes;
        uint8_t* bl = bh + kTcBBytes;
        ptx::tma_load_2d(bh, &map_uhi, &bars->full[s], k_row0 + ks * kTcKStage, u_col0 + int(n0));
        ptx::tma_load_2d(bl, &map_ulo, &bars->full[s], k_row0 + ks * kTcKStage, u_col0 + int(n0));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      bool first = true;
      for (int ks = 0; ks < num_kstages; ++ks) {
        const int s = ks % kTcStages;
        const uint32_t ph = (ks / kTcStages) & 1;
        ptx::mbar_wait(&bars->full[s], ph);
        ptx::tc_fence_after();
        const uint32_t ahi = ptx::smem_u32(smem + size_t(s) * kTcStageBytes);
        const uint32_t alo = ahi + kTcABytes;
        const uint32_t bhi = ahi + 2 * kTcABytes;
        const uint32_t blo = bhi + kTcBBytes;
#pragma unroll
        for (int k = 0; k < kTcKStage / 8; ++k) {
          // both operands K-major: rows of 128 B (32 tf32), 8-row swizzle atoms; 8 tf32 = 32 B per k-step
          const uint64_t dah = ptx::make_smem_desc_sw128(ahi + k * 32, 16, 1024);
          const uint64_t dal = ptx::make_smem_desc_sw128(alo + k * 32, 16, 1024);
          const uint64_t dbh = ptx::make_smem_desc_sw128(bhi + k * 32, 16, 1024);
          const uint64_t dbl = ptx::make_smem_desc_sw128(blo + k * 32, 16, 1024);
          ptx::mma_tf32_ss(tmem_base, dah, dbh, idesc, first ? 0u : 1u);
          first = false;
          ptx::mma_tf32_ss(tmem_base, dah, dbl, idesc, 1u);
          ptx::mma_tf32_ss(tmem_base, dal, dbh, idesc, 1u);
        }
        ptx::tc_commit(&bars->empty[s]);
      }
      ptx::tc_commit(&bars->tmem_full);
    }
  } else {
    const int quarter = warp & 3;
    float* mytr = tr + (warp - 2) * 32 * 33;
    ptx::mbar_wait(&bars->tmem_full, 0);
    ptx::tc_fence_after();
    const int64_t rbase = m0 + quarter * 32;
#pragma unroll 1
    for (int g = 0; g < kTcN / 32; ++g) {
      const int64_t cbase = n0 + g * 32;
      if (cbase >= N) break;
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(g * 32), v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) mytr[lane * 33 + q] = __uint_as_float(v[q]);
      __syncwarp();
      const int64_t col = cbase + lane;
      if (col < N) {
        // all 32 row loads of the 32 x 32 sub-tile are issued before the first use: the loop was
        // latency-bound (4 loads in flight per warp, 130 us per tile; profiles/r01_launches_loop_down.txt)
        float cv[32];
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          const int64_t row = rbase + rr;
          cv[rr] = (row < m) ? C[row * ldc + col] : 0.f;
        }
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          const int64_t row = rbase + rr;
          if (row < m) C[row * ldc + col] = __fsub_rn(cv[rr], mytr[rr * 33 + lane]);
        }
      }
      __syncwarp();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTcN);
  }
}

// Host-side state of one tq_gptq_loop call: the four tensor maps are encoded once.
struct TrailingTc {
  CUtensorMap ehi, elo, uhi, ulo;
  uint32_t idesc;
};

// UT_hi / UT_lo: n x kpad (U transposed, k contiguous, kpad = k rounded up to 4, zero padded)
int trailing_tc_prepare(TrailingTc* t, const float* E_hi, const float* E_lo, int64_t m, const float* UT_hi,
                        const float* UT_lo, int64_t kpad, int64_t n) {
  TQ_TRY(make_tmap_2d(&t->ehi, E_hi, TQ_F32, 1024, uint64_t(m), 1024 * 4, kTcKStage, kTcM));
  TQ_TRY(make_tmap_2d(&t->elo, E_lo, TQ_F32, 1024, uint64_t(m), 1024 * 4, kTcKStage, kTcM));
  TQ_TRY(make_tmap_2d(&t->uhi, UT_hi, TQ_F32, uint64_t(kpad), uint64_t(n), uint64_t(kpad) * 4, kTcKStage, kTcN));
  TQ_TRY(make_tmap_2d(&t->ulo, UT_lo, TQ_F32, uint64_t(kpad), uint64_t(n), uint64_t(kpad) * 4, kTcKStage, kTcN));
  t->idesc = ptx::make_idesc(/*TF32*/ 2u, /*A K-major*/ 0u, /*B K-major*/ 0u, kTcM, kTcN);
  return TQ_OK;
}

// C (m x N, ldc) -= E[:, e_col0 : e_col0 + kcount] . U[u_row0 : u_row0 + kcount, u_col0 : u_col0 + N]
int trailing_tc_launch(const TrailingTc* t, float* C, int64_t ldc, int64_t m, int64_t N, int64_t e_col0,
                       int64_t u_row0, int kcount, int64_t u_col0, cudaStream_t st) {
  static thread_local bool attr_done[kMaxDevices] = {};
  if (!attr_done[device_slot()]) {
    TQ_CUDA_CHECK(cudaFuncSetAttribute(trailing_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kTcSmem));
    attr_done[device_slot()] = true;
  }
  dim3 grid((unsigned)ceil_div(N, kTcN), (unsigned)ceil_div(m, kTcM));
  const int num_kstages = int(ceil_div(kcount, kTcKStage));
  const int pslot = prof_begin_launch(st, 2.0 * double(m) * double(N) * double(kcount), TQ_PROF_TRAILING_TC);
  trailing_tc_kernel<<<grid, kTcThreads, kTcSmem, st>>>(t->ehi, t->elo, t->uhi, t->ulo, C, ldc, m, N, int(e_col0),
                                                        int(u_row0), int(u_col0), num_kstages, t->idesc);
  prof_end_launch(st, pslot);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

}  // namespace tq
