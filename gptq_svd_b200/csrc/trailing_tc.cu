// Lazy-batch trailing update of the GPTQ loop as a tcgen05 GEMM (sm_100a):
//     C[m x N] -= E[m x K] . U[K x N],   K <= 1024 (one macro block of errors against every later column),
// replaces the cuBLAS SGEMM + two ATen passes of the reference (gptq_utils.py:539-545).
//
// Precision.  The reference runs this product in strict fp32 (TF32 off, :474-475); a single TF32 pass
// fails the 99.9 % code-parity bar by a wide margin, so both operands are
// split x = hi + lo with hi, lo representable in TF32 (round-to-nearest) and three MMAs
// accumulate hi.hi + hi.lo + lo.hi ("3xTF32").  The tensor core TRUNCATES when it adds into its fp32
// accumulator, a bias that grows linearly with the number of accumulation steps (measured on the SYRK:
// 2.0e-6 relative at 512 tokens, 4.4e-6 at 1024).  With all 1024 k of a macro block in ONE TMEM accumulation the
// codes of an ill-conditioned n = 4096 Hessian agreed with the fp32 reference to 99.71 % only; the strict-fp32
// SIMT product reaches 99.9998 % (tests/test_gpu_parity_large.py).  So TMEM accumulates kTcChunk = 128 k at a
// time, the epilogue warps drain each chunk and add it into fp32 REGISTERS with round-to-nearest (two 256-column
// TMEM buffers alternate, so the drain of one chunk runs under the MMAs of the next), and C -= acc happens
// once per update, like the reference's W[:, i2:] -= E @ Scale.
//
// One CTA per 128 x 256 tile of C (same roles as the SYRK, syrk.cu):
//   warp 0     TMA producer: per 32-deep K stage, E_hi / E_lo boxes [128 rows x 32 k] and
//              U^T_hi / U^T_lo boxes [256 n x 32 k] (both K-major; U is stored transposed by
//              split_transpose_u_kernel because N-major TF32 operands need the 32-byte-base
//              swizzle atom), 128-byte swizzle, 2-stage ring of 96 KB
//   warp 1     tcgen05.mma.kind::tf32 issuer, M=128 N=256 K=8, 3 MMAs per k-step
//   warps 4-11 epilogue (registers raised with setmaxnreg): tcgen05.ld -> fp32 register accumulators ->
//              C -= acc with 16-byte accesses along the thread's row
// Roofline: per tile and k, (128 + 256) x 8 bytes of operands for 2 x 128 x 256 flop = 21 flop per
// byte through L2; the E / U^T planes of one update (<= 16 MB + 92 MB) are L2-resident.
#include "common.cuh"
#include "tcgen05.cuh"

namespace tq {

int make_tmap_2d(CUtensorMap* tmap, const void* base, int dtype, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);

constexpr int kTcM = 128, kTcN = 256, kTcKStage = 32;
constexpr int kTcStages = 2;
constexpr int kTcChunkStages = 4;                        // 128 k per TMEM accumulation
constexpr int kTcABytes = kTcM * kTcKStage * 4;          // 16 KB (one of hi / lo)
constexpr int kTcBBytes = kTcN * kTcKStage * 4;          // 32 KB (one of hi / lo)
constexpr int kTcStageBytes = 2 * kTcABytes + 2 * kTcBBytes;   // 96 KB
constexpr int kTcCtrlWarps = 4, kTcEpiWarps = 8;
constexpr int kTcThreads = (kTcCtrlWarps + kTcEpiWarps) * 32;
constexpr size_t kTcSmem = size_t(kTcStages) * kTcStageBytes + 1024 + 256;

struct TcBarriers {
  uint64_t full[kTcStages];
  uint64_t empty[kTcStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kTcThreads, 1)
trailing_tc_kernel(const __grid_constant__ CUtensorMap map_ehi, const __grid_constant__ CUtensorMap map_elo,
                   const __grid_constant__ CUtensorMap map_uhi, const __grid_constant__ CUtensorMap map_ulo,
                   float* __restrict__ C, int64_t ldc, int64_t m, int64_t N, int e_col0 /*first E column*/,
                   int k_row0 /*first U row*/, int u_col0 /*first U column of this update*/, int num_kstages,
                   uint32_t idesc, float sign /* C += sign * (E . U): -1 for the GPTQ update, +1 for the sketch */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(smem + size_t(kTcStages) * kTcStageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n0 = int64_t(blockIdx.x) * kTcN, m0 = int64_t(blockIdx.y) * kTcM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      ptx::mbar_init(&bars->full[s], 1);
      ptx::mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&bars->tmem_full[b], 1);
      ptx::mbar_init(&bars->tmem_empty[b], kTcEpiWarps);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&map_ehi);
    ptx::prefetch_tmap(&map_elo);
    ptx::prefetch_tmap(&map_uhi);
    ptx::prefetch_tmap(&map_ulo);
  }
  if (warp == 1) {
    ptx::tmem_alloc(&bars->tmem_base, 2 * kTcN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int num_chunks = (num_kstages + kTcChunkStages - 1) / kTcChunkStages;

  if (warp >= kTcCtrlWarps) {
    // ------------------------------------------------ epilogue (8 warps)
    ptx::setmaxnreg_inc<224>();
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int half = (warp - kTcCtrlWarps) >> 2;   // column half of the 256-wide tile
    float acc[128];
#pragma unroll
    for (int c = 0; c < 128; ++c) acc[c] = 0.f;
    for (int ch = 0; ch < num_chunks; ++ch) {
      const int buf = ch & 1;
      ptx::mbar_wait(&bars->tmem_full[buf], (ch >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(buf * kTcN + half * 128);
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
    // C -= acc: thread = one row of the tile, 128 consecutive columns
    const int64_t row = m0 + quarter * 32 + lane;
    const int64_t col0 = n0 + half * 128;
    if (row < m && col0 < N) {
      float* crow = C + row * ldc + col0;
      if ((reinterpret_cast<uintptr_t>(crow) & 15) == 0 && col0 + 128 <= N) {
#pragma unroll
        for (int c = 0; c < 128; c += 4) {
          float4 v = *reinterpret_cast<float4*>(crow + c);
          v.x = __fadd_rn(v.x, sign * acc[c]);          // sign = -1: exactly v - acc
          v.y = __fadd_rn(v.y, sign * acc[c + 1]);
          v.z = __fadd_rn(v.z, sign * acc[c + 2]);
          v.w = __fadd_rn(v.w, sign * acc[c + 3]);
          *reinterpret_cast<float4*>(crow + c) = v;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (col0 + c < N) crow[c] = __fadd_rn(crow[c], sign * acc[c]);
      }
    }
  } else {
    ptx::setmaxnreg_dec<56>();
    if (warp == 0) {
      // ---------------------------------------------- TMA producer
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
      // ---------------------------------------------- MMA issuer
      if (lane == 0) {
        int ks = 0;
        for (int ch = 0; ch < num_chunks; ++ch) {
          const int buf = ch & 1;
          ptx::mbar_wait(&bars->tmem_empty[buf], ((ch >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + uint32_t(buf) * kTcN;
          const int ks_end = min(num_kstages, ks + kTcChunkStages);
          bool first = true;
          for (; ks < ks_end; ++ks) {
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
              // smallest terms first: the cross terms are 2^-11 of the main term
              ptx::mma_tf32_ss(tmem_d, dah, dbl, idesc, first ? 0u : 1u);
              first = false;
              ptx::mma_tf32_ss(tmem_d, dal, dbh, idesc, 1u);
              ptx::mma_tf32_ss(tmem_d, dah, dbh, idesc, 1u);
            }
            ptx::tc_commit(&bars->empty[s]);      // frees the smem stage when these MMAs retire
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
    ptx::tmem_dealloc(tmem_base, 2 * kTcN);
  }
}

// Host-side state of one tq_gptq_loop call: the four tensor maps are encoded once.
struct TrailingTc {
  CUtensorMap ehi, elo, uhi, ulo;
  uint32_t idesc;
};

// A_hi / A_lo: m x lda planes of the left operand (k contiguous, `kdim` valid columns);
// BT_hi / BT_lo: n x kpad planes of the right operand TRANSPOSED (k contiguous, kpad = k rounded up to 4, zero padded)
int trailing_tc_prepare_ex(TrailingTc* t, const float* A_hi, const float* A_lo, int64_t m, int64_t lda, int64_t kdim,
                           const float* BT_hi, const float* BT_lo, int64_t kpad, int64_t n) {
  TQ_TRY(make_tmap_2d(&t->ehi, A_hi, TQ_F32, uint64_t(kdim), uint64_t(m), uint64_t(lda) * 4, kTcKStage, kTcM));
  TQ_TRY(make_tmap_2d(&t->elo, A_lo, TQ_F32, uint64_t(kdim), uint64_t(m), uint64_t(lda) * 4, kTcKStage, kTcM));
  TQ_TRY(make_tmap_2d(&t->uhi, BT_hi, TQ_F32, uint64_t(kpad), uint64_t(n), uint64_t(kpad) * 4, kTcKStage, kTcN));
  TQ_TRY(make_tmap_2d(&t->ulo, BT_lo, TQ_F32, uint64_t(kpad), uint64_t(n), uint64_t(kpad) * 4, kTcKStage, kTcN));
  t->idesc = ptx::make_idesc(/*TF32*/ 2u, /*A K-major*/ 0u, /*B K-major*/ 0u, kTcM, kTcN);
  return TQ_OK;
}

// UT_hi / UT_lo: n x kpad (U transposed, k contiguous, kpad = k rounded up to 4, zero padded); E planes m x 1024
int trailing_tc_prepare(TrailingTc* t, const float* E_hi, const float* E_lo, int64_t m, const float* UT_hi,
                        const float* UT_lo, int64_t kpad, int64_t n) {
  return trailing_tc_prepare_ex(t, E_hi, E_lo, m, 1024, 1024, UT_hi, UT_lo, kpad, n);
}

// C (m x N, ldc) += sign * E[:, e_col0 : e_col0 + kcount] . U[u_row0 : u_row0 + kcount, u_col0 : u_col0 + N]
int trailing_tc_launch_ex(const TrailingTc* t, float* C, int64_t ldc, int64_t m, int64_t N, int64_t e_col0,
                          int64_t u_row0, int64_t kcount, int64_t u_col0, float sign, cudaStream_t st) {
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
                                                        int(u_row0), int(u_col0), num_kstages, t->idesc, sign);
  prof_end_launch(st, pslot);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

int trailing_tc_launch(const TrailingTc* t, float* C, int64_t ldc, int64_t m, int64_t N, int64_t e_col0,
                       int64_t u_row0, int kcount, int64_t u_col0, cudaStream_t st) {
  return trailing_tc_launch_ex(t, C, ldc, m, N, e_col0, u_row0, kcount, u_col0, -1.0f, st);
}

}  // namespace tq
