// Quantisation grid + blocked GPTQ column loop (strict fp32) for sm_100a.
//
// Replaces Quantizer.find_params (reference gptq_utils.py:249-266), gptq_fwrd
// (:459-565) and the Triton in-block kernel (:298-386).  Rows of W are independent
// in the GPTQ recurrence, so rows are the parallel dimension and columns are the
// serial one.  Layout in HBM (all fp32, row-major):
//   Wp  m x n   working copy of W in PERMUTED column order; column c is overwritten
//               by its dequantised value once quantised, so Wp ends as Q_final[:, perm]
//   U   k x n   propagation factors U[c,j] = R[c,j]/R[c,c] (rounded as the reference
//               rounds them, see prep_u_kernel)
//   E   m x 1024 quantisation errors of the current macro block (the GEMM operand of the lazy
//       updates; TF32 hi / lo halves for the tcgen05 path)
//   Cp  m x n   uint8 codes (permuted order), optional
// Per block of 128 permuted columns: gptq_block_kernel (register-resident column loop,
// one warp per 8 rows, lanes across columns, error broadcast by warp shuffle) then
// the lazy trailing update  W[:, c0+128:] -= E . U[c0:c0+128, c0+128:]: a tcgen05 3xTF32
// GEMM (trailing_tc.cu) by default, or the strict-fp32 SIMT GEMM below with
// TQ_LOOP_STRICT_FP32 / when n is not a multiple of 4.
#include <cuda.h>

#include "common.cuh"

namespace tq {

struct TrailingTc {
  CUtensorMap ehi, elo, uhi, ulo;
  uint32_t idesc;
};
int trailing_tc_prepare(TrailingTc* t, const float* E_hi, const float* E_lo, int64_t m, const float* UT_hi,
                        const float* UT_lo, int64_t kpad, int64_t n);
int trailing_tc_launch(const TrailingTc* t, float* C, int64_t ldc, int64_t m, int64_t N, int64_t e_col0,
                       int64_t u_row0, int kcount, int64_t u_col0, cudaStream_t st);

// x = hi + lo with hi, lo representable in TF32 (round to nearest): operands of the 3xTF32 update
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  const float rest = __fsub_rn(x, hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(rest));
  lo = __uint_as_float(l);
}

constexpr int kBlk = 128;        // columns per block step
constexpr int kMacro = 1024;     // lazy-batch width of the far trailing update (the reference's block_size scale)
constexpr int kRowsPerWarp = 2;  // rows interleaved per warp (8 rows x 4 warps left one warp per SM sub-partition:
constexpr int kWarpsPerCta = 16; // 1.5 us per column, latency-bound); 16 warps x 2 rows keep the 32-row CTA

// ------------------------------------------------------------------ find_params
// One warp per (row, group).  Bit-exact with torch: amin/amax, (mx-mn).clamp(1e-5)/max_q,
// round-half-even of -mn/scale, clamp to [0, max_q]  (gptq_utils.py:257-266).
__global__ void find_params_kernel(const float* __restrict__ W, int64_t ldw, int64_t m, int64_t n,
                                   int g, int ng, float max_q, int sym, float* __restrict__ scale,
                                   float* __restrict__ zero) {
  int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= m * ng) return;
  int64_t r = warp / ng;
  int gi = int(warp % ng);
  const float* p = W + r * ldw + int64_t(gi) * g;
  float mn = INFINITY, mx = -INFINITY;
  for (int j = lane; j < g; j += 32) {
    float v = p[j];
    if (sym) v = fabsf(v);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) {
    float s, z;
    if (sym) {
      s = __fdiv_rn(fmaxf(mx, 1e-5f), max_q);
      z = 0.f;
    } else {
      s = __fdiv_rn(fmaxf(__fsub_rn(mx, mn), 1e-5f), max_q);
      z = fminf(fmaxf(rintf(__fdiv_rn(-mn, s)), 0.f), max_q);
    }
    scale[warp] = s;
    zero[warp] = z;
  }
}

// ------------------------------------------------------------------ U preparation
// Triton semantics: pairs inside one reference block use R[c,j] * (1/R[c,c])
// (gptq_utils.py:374-377), pairs across reference blocks use R[c,j] / R[c,c] (:541).
// Torch semantics: U = R (the error itself is divided by the diagonal, :528-531).
template <typename TR>
__global__ void prep_u_kernel(const TR* __restrict__ R, int64_t ldr, int64_t k, int64_t n,
                              int ref_block, int semantics, float* __restrict__ U,
                              float* __restrict__ dvec) {
  int64_t c = blockIdx.y;
  float d = float(R[c * ldr + c]);
  float inv = __fdiv_rn(1.0f, d);
  int64_t blk_end = (c / ref_block + 1) * int64_t(ref_block);
  if (blk_end > k) blk_end = k;
  for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < n;
       j += int64_t(gridDim.x) * blockDim.x) {
    float r = float(R[c * ldr + j]);
    float u;
    if (semantics == TQ_LOOP_TORCH) u = r;
    else if (j < c) u = 0.f;
    else if (j < blk_end) u = __fmul_rn(r, inv);
    else u = __fdiv_rn(r, d);
    U[c * n + j] = u;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) dvec[c] = d;
}

// UT_hi / UT_lo (n x kpad, zero padded) = TF32 hi / lo split of U^T: the K-major B operand of
// the tcgen05 trailing update.  32 x 32 tiles through shared memory, coalesced both ways.
__global__ void split_transpose_u_kernel(const float* __restrict__ U, int64_t k, int64_t n, int64_t kpad,
                                         float* __restrict__ UT_hi, float* __restrict__ UT_lo) {
  __shared__ float t[32][33];
  const int64_t c0 = int64_t(blockIdx.y) * 32, j0 = int64_t(blockIdx.x) * 32;
  for (int a = threadIdx.y; a < 32; a += blockDim.y) {
    const int64_t c = c0 + a, j = j0 + threadIdx.x;
    t[a][threadIdx.x] = (c < k && j < n) ? U[c * n + j] : 0.f;
  }
  __syncthreads();
  for (int a = threadIdx.y; a < 32; a += blockDim.y) {
    const int64_t j = j0 + a, c = c0 + threadIdx.x;
    if (j < n && c < kpad) {
      float hi, lo;
      split_tf32(t[threadIdx.x][a], hi, lo);
      UT_hi[j * kpad + c] = hi;
      UT_lo[j * kpad + c] = lo;
    }
  }
}

// invperm must be filled with -1 first: an entry out of range or seen twice sets *bad (the reference raises an
// IndexError / silently duplicates columns; here the call fails with TQ_ERR_INVALID before anything is gathered)
__global__ void perm_meta_kernel(const int64_t* __restrict__ perm, int64_t n, int g,
                                 int* __restrict__ invperm, int* __restrict__ gidx,
                                 int* __restrict__ bad) {
  int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int64_t p = perm[j];
  if (p < 0 || p >= n) {
    atomicExch(bad, 1);
    gidx[j] = 0;
    return;
  }
  if (atomicCAS(&invperm[p], -1, int(j)) != -1) atomicExch(bad, 1);
  gidx[j] = int(p / g);
}

__global__ void gather_cols_kernel(const float* __restrict__ W, int64_t ldw, int64_t m, int64_t n,
                                   const int64_t* __restrict__ perm, float* __restrict__ Wp) {
  for (int64_t r = blockIdx.y; r < m; r += gridDim.y)          // gridDim.y <= 65535: rows loop (lm_head-sized m)
    for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < n;
         j += int64_t(gridDim.x) * blockDim.x)
      Wp[r * n + j] = W[r * ldw + perm[j]];
}

// ------------------------------------------------------------------ quantise one value
template <bool kHalfUp>
__device__ __forceinline__ void quantize(float w, float s, float z, float min_q, float max_q,
                                         float& q, float& qv) {
  float x = __fadd_rn(__fdiv_rn(w, s), z);
  q = kHalfUp ? floorf(__fadd_rn(x, 0.5f)) : rintf(x);
  q = fminf(fmaxf(q, min_q), max_q);
  qv = __fmul_rn(__fsub_rn(q, z), s);
}

// ------------------------------------------------------------------ in-block column loop
// CTA = 16 warps x 2 rows.  Lane L owns block columns 4L..4L+3 of its rows in
// registers; the 128 x 128 diagonal block of U sits in shared memory.  For column c
// the owner lane quantises its rows (independent chains -> ILP), the errors are
// broadcast with __shfl_sync and every lane applies the rank-1 update to its columns > c.
template <int SEM>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
gptq_block_kernel(float* __restrict__ Wp, int64_t n, const float* __restrict__ U,
                  const float* __restrict__ dvec, const float* __restrict__ scale,
                  const float* __restrict__ zero, int ng, const int* __restrict__ gidx, int64_t m,
                  int64_t c0, int cnt, float min_q, float max_q, float* __restrict__ E /* + column offset */,
                  float* __restrict__ E_hi /* non-null: TF32 hi / lo halves for the tcgen05 update */,
                  float* __restrict__ E_lo, uint8_t* __restrict__ Cp) {
  extern __shared__ float Us[];  // [kBlk][kBlk]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int idx = tid; idx < kBlk * kBlk; idx += blockDim.x) {
    int rr = idx / kBlk, cc = idx % kBlk;
    Us[idx] = (rr < cnt && cc < cnt) ? U[(c0 + rr) * n + c0 + cc] : 0.f;
  }
  const int64_t r0 = (int64_t(blockIdx.x) * kWarpsPerCta + warp) * kRowsPerWarp;
  float w[kRowsPerWarp][4], s[kRowsPerWarp][4], z[kRowsPerWarp][4], eo[kRowsPerWarp][4];
  uint32_t code[kRowsPerWarp];
#pragma unroll
  for (int rr = 0; rr < kRowsPerWarp; ++rr) {
    int64_t r = r0 + rr;
    code[rr] = 0;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      int col = 4 * lane + cc;
      bool ok = (r < m) && (col < cnt);
      w[rr][cc] = ok ? Wp[r * n + c0 + col] : 0.f;
      int g = ok ? gidx[c0 + col] : 0;
      s[rr][cc] = ok ? scale[r * ng + g] : 1.f;
      z[rr][cc] = ok ? zero[r * ng + g] : 0.f;
      eo[rr][cc] = 0.f;
    }
  }
  __syncthreads();

  for (int L = 0; L < 32; ++L) {
    if (4 * L >= cnt) break;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c = 4 * L + cc;
      if (c < cnt) {
        const float4 u = *reinterpret_cast<const float4*>(&Us[c * kBlk + 4 * lane]);
        float dd = 1.f;
        if (SEM == TQ_LOOP_TORCH) dd = dvec[c0 + c];
        float e[kRowsPerWarp];
#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp; ++rr) {
          float q, qv;
          quantize<SEM == TQ_LOOP_TRITON>(w[rr][cc], s[rr][cc], z[rr][cc], min_q, max_q, q, qv);
          float ev = __fsub_rn(w[rr][cc], qv);
          if (SEM == TQ_LOOP_TORCH) ev = __fdiv_rn(ev, dd);
          if (lane == L) {
            w[rr][cc] = qv;
            eo[rr][cc] = ev;
            code[rr] |= uint32_t(int(q - min_q) & 0xff) << (8 * cc);
          }
          e[rr] = __shfl_sync(0xffffffffu, ev, L);
        }
        const bool p0 = 4 * lane + 0 > c, p1 = 4 * lane + 1 > c, p2 = 4 * lane + 2 > c,
                   p3 = 4 * lane + 3 > c;
#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp; ++rr) {
          if (SEM == TQ_LOOP_TRITON) {
            if (p0) w[rr][0] = fmaf(-e[rr], u.x, w[rr][0]);
            if (p1) w[rr][1] = fmaf(-e[rr], u.y, w[rr][1]);
            if (p2) w[rr][2] = fmaf(-e[rr], u.z, w[rr][2]);
            if (p3) w[rr][3] = fmaf(-e[rr], u.w, w[rr][3]);
          } else {
            if (p0) w[rr][0] = __fsub_rn(w[rr][0], __fmul_rn(e[rr], u.x));
            if (p1) w[rr][1] = __fsub_rn(w[rr][1], __fmul_rn(e[rr], u.y));
            if (p2) w[rr][2] = __fsub_rn(w[rr][2], __fmul_rn(e[rr], u.z));
            if (p3) w[rr][3] = __fsub_rn(w[rr][3], __fmul_rn(e[rr], u.w));
          }
        }
      }
    }
  }

#pragma unroll
  for (int rr = 0; rr < kRowsPerWarp; ++rr) {
    int64_t r = r0 + rr;
    if (r >= m) continue;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      int col = 4 * lane + cc;
      const float ev = (col < cnt) ? eo[rr][cc] : 0.f;
      E[r * kMacro + col] = ev;
      if (E_hi) {
        float hi, lo;
        split_tf32(ev, hi, lo);
        E_hi[r * kMacro + col] = hi;
        E_lo[r * kMacro + col] = lo;
      }
      if (col < cnt) {
        Wp[r * n + c0 + col] = w[rr][cc];
        if (Cp) Cp[r * n + c0 + col] = uint8_t((code[rr] >> (8 * cc)) & 0xff);
      }
    }
  }
}

// ------------------------------------------------------------------ fused macro-block kernel
// One launch per 1024-column macro block = what the reference's Triton kernel does per block (gptq_utils.py:298-386:
// a tile of rows walks all columns of the block, every column quantised and its error propagated to the later
// columns of the block with one rank-1 FMA update per column), restructured for the SM.  A CTA owns 32 rows and
// walks the macro block in sub-blocks of 64 columns with two kinds of warps:
//   * the QUANTISER warp (warp 0, LANE = ROW; the other warps of its scheduler, 4 / 8 / 12, stay idle while it
//     works: next to four busy worker warps it received a fifth of the issue slots and took 480 - 650 cycles per
//     column, measured with the kernel's cycle counters) keeps its row's 64 values in registers, quantises column c and
//     applies w[j] = fma(-e, U[c, j], w[j]) to the later columns of the sub-block; U[c, :] is a shared-memory
//     broadcast.  No lane computes a redundant quantisation (the 2-rows-per-warp kernel above executes ~150
//     instructions per column and warp, 16 warps per 32 rows: issue-bound at 1070 cycles per column,
//     profiles/r02_ncu_gptq_block.txt);
//   * 12 WORKER warps (the warps of the other three schedulers) apply the previous sub-block's 64 rank-1 updates to the later columns of the macro block, in
//     column order, each one FMA - bit-identical to the Triton kernel's sequence.  Part A: the 64 columns the
//     quantiser needs next (all workers, a few hundred cycles); part B: everything beyond them, WHILE the
//     quantiser works on the next sub-block.  A worker warp owns 8 rows x a third of the columns (<= 80
//     accumulators per lane, one shared-memory value of U feeds 8 FMAs); U rows stream through shared memory in
//     double-buffered chunks of 8; the errors are stored k-major so that 8 rows are two 16-byte broadcasts.
// W stays in global memory (the 16 - 50 MB slab of a macro block is L2-resident); E / codes / dequantised values
// leave through shared memory with coalesced stores.  Requires ref_block % 1024 == 0 (every pair inside the macro
// block is then inside one reference block) and n % 4 == 0; other shapes take the kernels above.
constexpr int kFBlk = 64;
constexpr int kFRows = 32;
constexpr int kFWorkers = 12;                            // worker warps: the warps with id % 4 != 0
constexpr int kFThreads = 16 * 32;                       // warp 0 = quantiser, warps 4 / 8 / 12 only load and store
constexpr int kFChunk = 8;
constexpr int kFStages = 3;                              // U-chunk ring of part B
constexpr int kFMaxLen = kMacro - 2 * kFBlk;             // part B: at most 896 columns
constexpr int kFLd = kFBlk + 1;                          // lane = row: stride 65 is conflict-free
constexpr int kFNi = (kFMaxLen / 3 + 31) / 32;           // 10 column groups of 32 per lane and third
struct FusedSmem {
  // double-buffered per sub-block (the next sub-block's copies arrive with cp.async during this one's work)
  float Us[2][kFBlk][kFBlk];              // U[c0 + c, c0 + j]: the sub-block's own triangle
  float Ua[2][kFBlk][kFBlk];              // U[c0 - 64 + k, c0 + j]: previous sub-block's rows against this one's columns
  float Sb[2][kFRows][kFLd];
  float Zb[2][kFRows][kFLd];
  float Wb[kFRows][kFLd];
  float Et[2][kFBlk][kFRows];             // errors, k-major, double-buffered
  float Dv[2][kFBlk];
  uint8_t Cb[kFBlk][kFRows];              // codes, column-major: a warp store is 32 consecutive bytes
  float Uc[kFStages][kFChunk][kFMaxLen + 64];      // + 64: a third rounded up to 32 columns may reach past 896 (values unused)
};

__device__ __forceinline__ void cp_async_f4(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_f1(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, 384;" ::: "memory"); }

// part B for one worker warp: NI column groups of 32 per lane.  3-stage ring of U chunks, ONE barrier per chunk:
// chunk ck + 2 is issued after the barrier of chunk ck, i.e. when every worker has finished chunk ck - 1, whose
// buffer it overwrites.
template <int SEM, int NI>
__device__ __forceinline__ void fused_part_b(FusedSmem& sm, float* __restrict__ Wp, int64_t n, const float* __restrict__ U,
                                             int64_t m, int64_t r0, int64_t krow0 /*first U row*/, int kcnt,
                                             int64_t j0 /*first column*/, int len, int quarter, int ebuf, int wtid) {
  const int lane = wtid & 31, widx = wtid >> 5;                   // worker index 0 .. 11
  const int rg = widx / 3, cq = widx % 3;
  const int jbase = cq * quarter + lane;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(U) & 15) == 0) && (n % 4 == 0) && (j0 % 4 == 0);
  const int nchunks = (kcnt + kFChunk - 1) / kFChunk;
  auto load_chunk = [&](int ck) {
    if (ck < nchunks) {
      const int buf = ck % kFStages;
      const int ngran = (len + 3) / 4;
      for (int idx = wtid; idx < kFChunk * ngran; idx += kFWorkers * 32) {
        const int kk = idx / ngran, g = idx % ngran;
        const int krow = ck * kFChunk + kk;
        float* dst = &sm.Uc[buf][kk][4 * g];
        if (krow < kcnt) {
          const float* src = U + (krow0 + krow) * n + j0 + 4 * g;
          if (vec_ok && 4 * g + 4 <= len) {
            cp_async_f4(dst, src);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (4 * g + e < len) cp_async_f1(dst + e, src + e);
          }
        }
      }
    }
    cp_async_commit_group();            // one group per call, empty or not: the wait counts below stay exact
  };
  load_chunk(0);
  load_chunk(1);
  float acc[8][NI];
  float* const wbase = Wp + (r0 + rg * 8) * n + j0 + jbase;      // row a of the group: wbase + a n
  const int64_t rleft = m - (r0 + rg * 8);
  const int nrows = rleft < 8 ? int(rleft) : 8;                     // rows of this group inside the matrix (may be <= 0)
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const bool okc = jbase + 32 * i < len;
#pragma unroll
    for (int a = 0; a < 8; ++a) acc[a][i] = (okc && a < nrows) ? __ldcg(wbase + a * n + 32 * i) : 0.f;
  }
  for (int ck = 0; ck < nchunks; ++ck) {
    const int buf = ck % kFStages;
    cp_async_wait_group<1>();             // chunk ck has landed (chunk ck + 1 may still be in flight)
    worker_barrier();
    load_chunk(ck + 2);
    const int kmax = min(kFChunk, kcnt - ck * kFChunk);
#pragma unroll 2
    for (int kk = 0; kk < kmax; ++kk) {
      const float4 ea = *reinterpret_cast<const float4*>(&sm.Et[ebuf][ck * kFChunk + kk][rg * 8]);
      const float4 eb = *reinterpret_cast<const float4*>(&sm.Et[ebuf][ck * kFChunk + kk][rg * 8 + 4]);
      const float e[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
      const float* ucol = &sm.Uc[buf][kk][jbase];                // jbase + 32 i < 3 * 320 = 960: inside the padded row
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const float u = ucol[32 * i];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          if (SEM == TQ_LOOP_TRITON) acc[a][i] = fmaf(-e[a], u, acc[a][i]);
          else acc[a][i] = __fsub_rn(acc[a][i], __fmul_rn(e[a], u));
        }
      }
    }
  }
  cp_async_wait_group<0>();
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const bool okc = jbase + 32 * i < len;
#pragma unroll
    for (int a = 0; a < 8; ++a)
      if (okc && a < nrows) wbase[a * n + 32 * i] = acc[a][i];
  }
}

// The quantiser warp's work on one sub-block (LANE = ROW), a function of its own so that the register allocator
// sees only this loop: inlined into the kernel it spilled the row's values around every column.  s / z arrive one
// column ahead of their use; the value the NEXT column's quantisation depends on, w[c + 1], is updated first, the
// other updates fill the shadow of its latency.
template <int SEM>
__device__ __noinline__ void fused_quantise(FusedSmem& sm, int eb, int lane, int cnt, float min_q, float max_q) {
      float w[kFBlk];
#pragma unroll
      for (int j = 0; j < kFBlk; ++j) w[j] = sm.Wb[lane][j];
      float s_cur = sm.Sb[eb][lane][0], z_cur = sm.Zb[eb][lane][0];
#pragma unroll
      for (int c = 0; c < kFBlk; ++c) {
        // next column's grid parameters and the one U value on the critical path: issued before this column's chain
        const float s_nxt = (c + 1 < kFBlk) ? sm.Sb[eb][lane][c + 1 < kFBlk ? c + 1 : c] : 1.f;
        const float z_nxt = (c + 1 < kFBlk) ? sm.Zb[eb][lane][c + 1 < kFBlk ? c + 1 : c] : 0.f;
        if (c < cnt) {
          const float unext = (c + 1 < kFBlk) ? sm.Us[eb][c][c + 1 < kFBlk ? c + 1 : c] : 0.f;
          float qq, qv;
          quantize<SEM == TQ_LOOP_TRITON>(w[c], s_cur, z_cur, min_q, max_q, qq, qv);
          float ev = __fsub_rn(w[c], qv);
          if (SEM == TQ_LOOP_TORCH) ev = __fdiv_rn(ev, sm.Dv[eb][c]);
          if (c + 1 < kFBlk) {
            if (SEM == TQ_LOOP_TRITON) w[c + 1] = fmaf(-ev, unext, w[c + 1]);
            else w[c + 1] = __fsub_rn(w[c + 1], __fmul_rn(ev, unext));
          }
          sm.Wb[lane][c] = qv;
          sm.Et[eb][c][lane] = ev;
          sm.Cb[c][lane] = uint8_t(int(qq - min_q) & 0xff);
#pragma unroll
          for (int j = c + 2; j < kFBlk; ++j) {
            const float u = sm.Us[eb][c][j];                   // broadcast
            if (SEM == TQ_LOOP_TRITON) w[j] = fmaf(-ev, u, w[j]);
            else w[j] = __fsub_rn(w[j], __fmul_rn(ev, u));
          }
        } else {
          sm.Et[eb][c][lane] = 0.f;
        }
        s_cur = s_nxt;
        z_cur = z_nxt;
      }
}

template <int SEM>
__global__ void __launch_bounds__(kFThreads, 1)
gptq_macro_kernel(float* __restrict__ Wp, int64_t n, const float* __restrict__ U, const float* __restrict__ dvec,
                  const float* __restrict__ scale, const float* __restrict__ zero, int ng,
                  const int* __restrict__ gidx, int64_t m, int64_t M0, int cntM, float min_q, float max_q,
                  float* __restrict__ E, float* __restrict__ E_hi, float* __restrict__ E_lo,
                  uint8_t* __restrict__ Cp, long long* __restrict__ stats /* TQ_TRACE: cycle counters of CTA 0 */) {
  extern __shared__ __align__(16) uint8_t fused_raw[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>(fused_raw);      // no integer round trip: keeps the accesses LDS / STS
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool quantiser = warp == 0;
  const bool worker = (warp & 3) != 0;
  const int wtid = ((warp >> 2) * 3 + (warp & 3) - 1) * 32 + lane;     // dense worker thread id 0 .. 383
  const int64_t r0 = int64_t(blockIdx.x) * kFRows;
  const bool timed = stats != nullptr && blockIdx.x == 0 && lane == 0 && (warp == 1 || quantiser);
  long long tmark = timed ? clock64() : 0;
  auto lap = [&](int slot) {
    if (timed) {
      const long long t = clock64();
      atomicAdd(reinterpret_cast<unsigned long long*>(stats + slot), (unsigned long long)(t - tmark));
      tmark = t;
    }
  };
  const bool u_vec = ((reinterpret_cast<uintptr_t>(U) & 15) == 0) && (n % 4 == 0) && (M0 % 4 == 0);

  // U triangle, U rows of the previous sub-block, scales / zeros (through the permutation) and diagonal of sub-block
  // `bi` into buffer bi & 1: asynchronous copies, none of them depends on W
  auto prefetch = [&](int bi) {
    const int b0 = bi * kFBlk;
    if (b0 < cntM) {
      const int cnt = min(kFBlk, cntM - b0);
      const int64_t c0 = M0 + b0;
      const int pb = bi & 1;
      for (int idx = tid; idx < kFBlk * (kFBlk / 4); idx += kFThreads) {
        const int c = idx / (kFBlk / 4), j4 = (idx % (kFBlk / 4)) * 4;
        float* du = &sm.Us[pb][c][j4];
        float* da = &sm.Ua[pb][c][j4];
        if (u_vec && c < cnt && j4 + 4 <= cnt) {
          cp_async_f4(du, U + (c0 + c) * n + c0 + j4);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (c < cnt && j4 + e < cnt) cp_async_f1(du + e, U + (c0 + c) * n + c0 + j4 + e);
            else du[e] = 0.f;
          }
        }
        if (b0 > 0) {
          if (u_vec && j4 + 4 <= cnt) {
            cp_async_f4(da, U + (c0 - kFBlk + c) * n + c0 + j4);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (j4 + e < cnt) cp_async_f1(da + e, U + (c0 - kFBlk + c) * n + c0 + j4 + e);
              else da[e] = 0.f;
            }
          }
        }
      }
      for (int idx = tid; idx < kFRows * kFBlk; idx += kFThreads) {
        const int r = idx / kFBlk, j = idx % kFBlk;
        if ((r0 + r < m) && (j < cnt)) {
          const int g = gidx[c0 + j];
          cp_async_f1(&sm.Sb[pb][r][j], scale + (r0 + r) * ng + g);
          cp_async_f1(&sm.Zb[pb][r][j], zero + (r0 + r) * ng + g);
        } else {
          sm.Sb[pb][r][j] = 1.f;
          sm.Zb[pb][r][j] = 0.f;
        }
      }
      if (tid < kFBlk) sm.Dv[pb][tid] = (SEM == TQ_LOOP_TORCH && tid < cnt) ? dvec[c0 + tid] : 1.f;
    }
    cp_async_commit_group();
  };
  prefetch(0);

  for (int b0 = 0, bi = 0; b0 < cntM; b0 += kFBlk, ++bi) {
    const int cnt = min(kFBlk, cntM - b0);
    const int64_t c0 = M0 + b0;
    const int eb = bi & 1;                 // this sub-block's buffers; the previous one's are eb ^ 1
    // ---- W sub-block (after the previous iteration's part B has written it)
    for (int idx = tid; idx < kFRows * kFBlk; idx += kFThreads) {
      const int r = idx / kFBlk, j = idx % kFBlk;
      sm.Wb[r][j] = ((r0 + r < m) && (j < cnt)) ? __ldcg(&Wp[(r0 + r) * n + c0 + j]) : 0.f;
    }
    cp_async_wait_group<0>();              // this sub-block's prefetch (issued one iteration ago)
    __syncthreads();
    prefetch(bi + 1);                      // lands while the quantiser and part B work
    lap(quantiser ? 7 : 0);                                        // 0: load phase

    // ---- part A (workers): the previous sub-block's 64 updates on THIS sub-block's columns, in shared memory
    if (b0 > 0) {
      const int j = tid & 63, r4 = tid >> 6;                    // 8 groups of 4 rows x 64 columns
      float a0 = sm.Wb[r4 * 4 + 0][j], a1 = sm.Wb[r4 * 4 + 1][j], a2 = sm.Wb[r4 * 4 + 2][j], a3 = sm.Wb[r4 * 4 + 3][j];
#pragma unroll
      for (int k8 = 0; k8 < kFBlk; k8 += 8) {
        float4 e[8];
        float u[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {                            // all loads of 8 steps first, then the dependent FMAs
          e[q] = *reinterpret_cast<const float4*>(&sm.Et[eb ^ 1][k8 + q][r4 * 4]);
          u[q] = sm.Ua[eb][k8 + q][j];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (SEM == TQ_LOOP_TRITON) {
            a0 = fmaf(-e[q].x, u[q], a0);
            a1 = fmaf(-e[q].y, u[q], a1);
            a2 = fmaf(-e[q].z, u[q], a2);
            a3 = fmaf(-e[q].w, u[q], a3);
          } else {
            a0 = __fsub_rn(a0, __fmul_rn(e[q].x, u[q]));
            a1 = __fsub_rn(a1, __fmul_rn(e[q].y, u[q]));
            a2 = __fsub_rn(a2, __fmul_rn(e[q].z, u[q]));
            a3 = __fsub_rn(a3, __fmul_rn(e[q].w, u[q]));
          }
        }
      }
      sm.Wb[r4 * 4 + 0][j] = a0;
      sm.Wb[r4 * 4 + 1][j] = a1;
      sm.Wb[r4 * 4 + 2][j] = a2;
      sm.Wb[r4 * 4 + 3][j] = a3;
    }
    __syncthreads();
    lap(quantiser ? 7 : 1);                                        // 1: part A

    if (quantiser) {
      fused_quantise<SEM>(sm, eb, lane, cnt, min_q, max_q);
    } else if (worker && b0 > 0) {
      // ---- part B (workers): the previous sub-block's updates on the columns beyond this sub-block
      const int len = cntM - (b0 + kFBlk);
      if (len > 0) {
        const int quarter = ((len + 2) / 3 + 31) / 32 * 32;       // columns per worker third
        const int64_t j0 = M0 + b0 + kFBlk;
        const int64_t krow0 = c0 - kFBlk;
#define TQ_PART_B(NI) fused_part_b<SEM, NI>(sm, Wp, n, U, m, r0, krow0, kFBlk, j0, len, quarter, eb ^ 1, wtid)
        switch (quarter / 32) {
          case 1: TQ_PART_B(1); break;
          case 2: TQ_PART_B(2); break;
          case 3: TQ_PART_B(3); break;
          case 4: TQ_PART_B(4); break;
          case 5: TQ_PART_B(5); break;
          case 6: TQ_PART_B(6); break;
          case 7: TQ_PART_B(7); break;
          case 8: TQ_PART_B(8); break;
          case 9: TQ_PART_B(9); break;
          default: TQ_PART_B(kFNi); break;
        }
#undef TQ_PART_B
      }
    }
    lap(quantiser ? 2 : 3);                                        // 2: quantiser, 3: part B (own work, before the barrier)
    __syncthreads();
    lap(quantiser ? 7 : 4);                                        // 4: workers waiting for the quantiser

    // ---- store: dequantised values, codes, errors (coalesced along the columns)
    for (int idx = tid; idx < kFRows * kFBlk; idx += kFThreads) {
      const int r = idx / kFBlk, j = idx % kFBlk;
      if (r0 + r >= m) continue;
      const float ev = sm.Et[eb][j][r];
      const int64_t eo = (r0 + r) * kMacro + b0 + j;
      E[eo] = ev;
      if (E_hi) {
        float hi, lo;
        split_tf32(ev, hi, lo);
        E_hi[eo] = hi;
        E_lo[eo] = lo;
      }
      if (j < cnt) {
        Wp[(r0 + r) * n + c0 + j] = sm.Wb[r][j];
        if (Cp) Cp[(r0 + r) * n + c0 + j] = sm.Cb[j][r];
      }
    }
    __syncthreads();       // Wb is reloaded next; part B's global writes precede the next loads
    lap(quantiser ? 7 : 5);                                        // 5: store phase
  }
  cp_async_wait_group<0>();
}

// ------------------------------------------------------------------ lazy trailing update (SIMT)
// C[m x N] -= A[m x K] . B[K x N], K <= 1024, strict fp32 (the reference disables TF32,
// gptq_utils.py:474-475).  64 x 64 tile per CTA, 4 x 4 per thread, k ascending.  Three arithmetics:
//   kSeqDelta   acc = sum_k a b (FMA chain from 0), C = C - acc: the structure of the reference's cross-block
//               update W[:, i2:] -= E @ Scale (:539-545; an SGEMM, then one subtraction);
//   kSeqFma     acc = C, acc = fma(-a, b, acc) for k ascending: BIT-IDENTICAL to applying the K rank-1 updates of the
//               Triton kernel one after another (W -= e (x) corr contracted to an FMA, :380-386).  Pairs (c, j) inside
//               one reference block must be updated this way: every step rounds at the magnitude of W, and a
//               trajectory that rounds differently flips ~0.2-0.5 % of the codes at n = 4096 on an ill-conditioned
//               spectrum (measured: the reference arithmetic on a CPU, with against without FMA contraction);
//   kSeqMulSub  acc = C, acc = acc - rn(a b): the torch loop's unfused W1[:, c:] -= e * R[c, c:] (:531).
constexpr int kTM = 64, kTN = 64, kTK = 32;
constexpr int kSeqDelta = 0, kSeqFma = 1, kSeqMulSub = 2;
template <int MODE>
__global__ void __launch_bounds__(256)
trailing_update_kernel(float* __restrict__ C, int64_t ldc, const float* __restrict__ A, int64_t lda,
                       const float* __restrict__ B, int64_t ldb, int64_t m, int64_t N, int K) {
  __shared__ float As[kTK][kTM + 4];  // As[k][row]
  __shared__ float Bs[kTK][kTN];      // Bs[k][col]
  const int tid = threadIdx.x;
  const int64_t row0 = int64_t(blockIdx.y) * kTM, col0 = int64_t(blockIdx.x) * kTN;
  const int tr = (tid / 16) * 4, tc = (tid % 16) * 4;
  float acc[4][4] = {};
  if (MODE != kSeqDelta) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t gr = row0 + tr + i, gc = col0 + tc + j;
        acc[i][j] = (gr < m && gc < N) ? C[gr * ldc + gc] : 0.f;
      }
  }
  for (int k0 = 0; k0 < K; k0 += kTK) {
    for (int idx = tid; idx < kTM * kTK; idx += 256) {
      int r = idx / kTK, kk = idx % kTK;
      int64_t gr = row0 + r;
      As[kk][r] = (gr < m && k0 + kk < K) ? A[gr * lda + k0 + kk] : 0.f;
    }
    for (int idx = tid; idx < kTK * kTN; idx += 256) {
      int kk = idx / kTN, c = idx % kTN;
      int64_t gc = col0 + c;
      Bs[kk][c] = (gc < N && k0 + kk < K) ? B[int64_t(k0 + kk) * ldb + gc] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][tr]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tc]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (MODE == kSeqDelta) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
          else if (MODE == kSeqFma) acc[i][j] = fmaf(-av[i], bv[j], acc[i][j]);
          else acc[i][j] = __fsub_rn(acc[i][j], __fmul_rn(av[i], bv[j]));
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t gr = row0 + tr + i;
    if (gr >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t gc = col0 + tc + j;
      if (gc < N) C[gr * ldc + gc] = (MODE == kSeqDelta) ? __fsub_rn(C[gr * ldc + gc], acc[i][j]) : acc[i][j];
    }
  }
}

// ------------------------------------------------------------------ tail RTN (columns >= k)
// Half-even rounding of the error-compensated tail, no propagation (gptq_utils.py:547-553).
__global__ void tail_rtn_kernel(float* __restrict__ Wp, int64_t n, int64_t m, int64_t k,
                                const float* __restrict__ scale, const float* __restrict__ zero,
                                int ng, const int* __restrict__ gidx, float min_q, float max_q,
                                uint8_t* __restrict__ Cp) {
  for (int64_t r = blockIdx.y; r < m; r += gridDim.y)
    for (int64_t j = k + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < n;
         j += int64_t(gridDim.x) * blockDim.x) {
      int g = gidx[j];
      float q, qv;
      quantize<false>(Wp[r * n + j], scale[r * ng + g], zero[r * ng + g], min_q, max_q, q, qv);
      Wp[r * n + j] = qv;
      if (Cp) Cp[r * n + j] = uint8_t(int(q - min_q));
    }
}

// Restore the original column order (gptq_utils.py:556-557) with coalesced writes.
__global__ void unpermute_kernel(const float* __restrict__ Wp, const uint8_t* __restrict__ Cp,
                                 int64_t m, int64_t n, const int* __restrict__ invperm,
                                 float* __restrict__ Wq, int64_t ldq, uint8_t* __restrict__ codes,
                                 int64_t ldc) {
  for (int64_t r = blockIdx.y; r < m; r += gridDim.y)
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += int64_t(gridDim.x) * blockDim.x) {
      int j = invperm[i];
      Wq[r * ldq + i] = Wp[r * n + j];
      if (codes) codes[r * ldc + i] = Cp[r * n + j];
    }
}

// ------------------------------------------------------------------ packing
// Row bit-stream, value j at bits [j*bits, (j+1)*bits), little-endian uint32 words.
__global__ void pack_codes_kernel(const uint8_t* __restrict__ codes, int64_t ldc, int64_t m,
                                  int64_t n, int bits, uint32_t* __restrict__ packed, int64_t ldp,
                                  int64_t nwords) {
  for (int64_t r = blockIdx.y; r < m; r += gridDim.y)
  for (int64_t w = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; w < nwords;
       w += int64_t(gridDim.x) * blockDim.x) {
    int64_t bit0 = w * 32;
    int64_t j = bit0 / bits;
    int off = int(bit0 - j * bits);  // bits of value j already consumed by the previous word
    uint64_t acc = 0;
    int have = 0;
    if (off) {
      acc = uint64_t(codes[r * ldc + j]) >> off;
      have = bits - off;
      ++j;
    }
    while (have < 32 && j < n) {
      acc |= uint64_t(codes[r * ldc + j]) << have;
      have += bits;
      ++j;
    }
    packed[r * ldp + w] = uint32_t(acc & 0xffffffffu);
  }
}

// GPTQ / AutoGPTQ checkpoint layout: words [ceil(n bits / 32), m] int32, word w of OUTPUT row j holds bits
// [32 w, 32 w + 32) of that row's LSB-first bitstream over the INPUT dimension (for 2 / 4 / 8 bits: 32 / bits
// consecutive input columns per word; for 3 bits AutoGPTQ's 32-values-in-3-words scheme, which is the same
// bitstream).  `add` is added to every code first (modulo 2^bits) - qzeros store zero - 1 in the v1 format.
__global__ void pack_gptq_kernel(const uint8_t* __restrict__ codes, int64_t ldc, int64_t m, int64_t n, int bits,
                                 int add, uint32_t* __restrict__ out, int64_t ldo, int64_t nwords) {
  const int64_t w = blockIdx.y;
  const uint32_t mask = (1u << bits) - 1u;
  for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < m; j += int64_t(gridDim.x) * blockDim.x) {
    const int64_t bit0 = w * 32;
    int64_t i = bit0 / bits;
    int off = int(bit0 - i * bits);
    uint64_t acc = 0;
    int have = 0;
    if (off) {
      acc = uint64_t((uint32_t(codes[j * ldc + i]) + uint32_t(add)) & mask) >> off;
      have = bits - off;
      ++i;
    }
    while (have < 32 && i < n) {
      acc |= uint64_t((uint32_t(codes[j * ldc + i]) + uint32_t(add)) & mask) << have;
      have += bits;
      ++i;
    }
    out[w * ldo + j] = uint32_t(acc & 0xffffffffu);
  }
}

}  // namespace tq

using namespace tq;

extern "C" int tq_pack_gptq(const uint8_t* codes, int64_t ldc, int64_t m, int64_t n, int bits, int add,
                            uint32_t* out, int64_t ldo, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(codes && out && m > 0 && n > 0 && ldc >= n && ldo >= m, "tq_pack_gptq: bad arguments");
  TQ_REQUIRE(bits >= 2 && bits <= 8, "tq_pack_gptq: bits=%d outside [2,8]", bits);
  const int64_t nwords = (n * bits + 31) / 32;
  TQ_REQUIRE(nwords <= 65535, "tq_pack_gptq: too many words per row");
  dim3 grid((unsigned)imin(ceil_div(m, 128), 1024), (unsigned)nwords);
  pack_gptq_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(codes, ldc, m, n, bits, add, out, ldo, nwords);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

extern "C" int tq_find_params(const float* W, int64_t ldw, int64_t m, int64_t n, int bits, int group,
                              int sym, float* scale, float* zero, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(W && scale && zero, "tq_find_params: null pointer");
  TQ_REQUIRE(m > 0 && n > 0 && ldw >= n, "tq_find_params: bad shape m=%lld n=%lld ldw=%lld",
             (long long)m, (long long)n, (long long)ldw);
  TQ_REQUIRE(bits >= 2 && bits <= 8, "tq_find_params: bits=%d outside [2,8]", bits);
  int64_t g = group > 0 ? group : n;
  TQ_REQUIRE(n % g == 0, "tq_find_params: in_features %lld not divisible by group size %lld",
             (long long)n, (long long)g);
  int ng = int(n / g);
  float max_q = sym ? float((1 << (bits - 1)) - 1) : float((1 << bits) - 1);
  int64_t warps = m * ng;
  int threads = 256;
  int64_t blocks = ceil_div(warps * 32, threads);
  find_params_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      W, ldw, m, n, int(g), ng, max_q, sym, scale, zero);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}

extern "C" int tq_gptq_loop_workspace(int64_t m, int64_t n, int64_t k, size_t* bytes) {
  TQ_REQUIRE(bytes && m > 0 && n > 0 && k >= 0 && k <= n, "tq_gptq_loop_workspace: bad arguments");
  size_t b = 0;
  b += ws_bytes_for(size_t(m) * n, 4);     // Wp
  b += ws_bytes_for(size_t(k) * n, 4);     // U
  b += ws_bytes_for(size_t(m) * kMacro, 4) * 3;  // E, E_hi, E_lo: m x 1024
  b += ws_bytes_for(size_t(k + 4) * n, 4) * 2; // UT_hi, UT_lo
  b += ws_bytes_for(size_t(m) * n, 1);     // Cp
  b += ws_bytes_for(size_t(n), 4) * 2;     // invperm, gidx
  b += ws_bytes_for(size_t(k) + 1, 4);     // dvec
  b += ws_bytes_for(1, 4);                 // bad flag
  b += ws_bytes_for(8, 8);                 // cycle counters of the fused kernel (TQ_TRACE)
  *bytes = b;
  return TQ_OK;
}

extern "C" int tq_gptq_loop(const float* W, int64_t ldw, const void* R, int r_dtype, int64_t ldr,
                            int64_t k, const int64_t* perm, const float* scale, const float* zero,
                            int64_t m, int64_t n, int bits, int group, int sym, int ref_block,
                            int semantics, float* Wq_out, int64_t ldq, uint8_t* codes_out,
                            int64_t ldc, void* ws, size_t ws_bytes, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(W && perm && scale && zero && Wq_out, "tq_gptq_loop: null pointer");
  TQ_REQUIRE(m > 0 && n > 0 && k >= 0 && k <= n, "tq_gptq_loop: bad shape m=%lld n=%lld k=%lld",
             (long long)m, (long long)n, (long long)k);
  TQ_REQUIRE(k == 0 || R, "tq_gptq_loop: null R");
  TQ_REQUIRE(ldw >= n && ldq >= n && (k == 0 || ldr >= n) && (!codes_out || ldc >= n),
             "tq_gptq_loop: leading dimension too small");
  TQ_REQUIRE(bits >= 2 && bits <= 8, "tq_gptq_loop: bits=%d outside [2,8]", bits);
  TQ_REQUIRE(r_dtype == TQ_F64 || r_dtype == TQ_F32, "tq_gptq_loop: R must be fp64 or fp32");
  const bool strict_fp32 = (semantics & TQ_LOOP_STRICT_FP32) != 0;
  semantics &= ~TQ_LOOP_STRICT_FP32;
  TQ_REQUIRE(semantics == TQ_LOOP_TRITON || semantics == TQ_LOOP_TORCH, "tq_gptq_loop: bad semantics");
  // the tcgen05 path needs 16-byte row strides for TMA; other shapes take the SIMT fp32 GEMM
  const bool use_tc = !strict_fp32 && k > 0;
  TQ_REQUIRE(ref_block > 0, "tq_gptq_loop: ref_block must be positive");
  int64_t g = group > 0 ? group : n;
  TQ_REQUIRE(n % g == 0, "tq_gptq_loop: in_features %lld not divisible by group size %lld",
             (long long)n, (long long)g);
  const int ng = int(n / g);
  cudaStream_t st = (cudaStream_t)stream;

  Workspace wsp(ws, ws_bytes);
  float* Wp = wsp.take<float>(size_t(m) * n);
  float* U = wsp.take<float>(size_t(k) * n);
  float* E = wsp.take<float>(size_t(m) * kMacro);
  float* E_hi = wsp.take<float>(size_t(m) * kMacro);
  float* E_lo = wsp.take<float>(size_t(m) * kMacro);
  const int64_t kpad = (k + 3) / 4 * 4;
  float* U_hi = wsp.take<float>(size_t(kpad) * n);   // U^T hi / lo, n x kpad
  float* U_lo = wsp.take<float>(size_t(kpad) * n);
  uint8_t* Cp = wsp.take<uint8_t>(size_t(m) * n);
  int* invperm = wsp.take<int>(n);
  int* gidx = wsp.take<int>(n);
  float* dvec = wsp.take<float>(k + 1);
  int* bad = wsp.take<int>(1);
  long long* fstats = wsp.take<long long>(8);
  if (wsp.overflow) {
    set_error("tq_gptq_loop: workspace too small (%zu < %zu)", ws_bytes, wsp.off);
    return TQ_ERR_WORKSPACE;
  }
  if (!codes_out) Cp = nullptr;

  const float max_q = sym ? float((1 << (bits - 1)) - 1) : float((1 << bits) - 1);
  const float min_q = sym ? -max_q : 0.f;

  TQ_CUDA_CHECK(cudaMemsetAsync(bad, 0, sizeof(int), st));
  TQ_CUDA_CHECK(cudaMemsetAsync(invperm, 0xff, sizeof(int) * n, st));        // -1
  perm_meta_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(perm, n, int(g), invperm, gidx, bad);
  TQ_LAUNCH_CHECK();
  {
    int hbad = 0;      // the one host synchronisation of this call: a 4-byte read-back before anything is gathered
    TQ_CUDA_CHECK(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    TQ_CUDA_CHECK(cudaStreamSynchronize(st));
    TQ_REQUIRE(hbad == 0, "tq_gptq_loop: perm is not a permutation of 0 .. %lld", (long long)(n - 1));
  }
  {
    dim3 grid((unsigned)imin(ceil_div(n, 256), 64), (unsigned)imin(m, 65535));
    gather_cols_kernel<<<grid, 256, 0, st>>>(W, ldw, m, n, perm, Wp);
    TQ_LAUNCH_CHECK();
  }
  if (k > 0) {
    dim3 grid((unsigned)imin(ceil_div(n, 256), 64), (unsigned)k);
    if (r_dtype == TQ_F64)
      prep_u_kernel<double><<<grid, 256, 0, st>>>((const double*)R, ldr, k, n, ref_block, semantics, U, dvec);
    else
      prep_u_kernel<float><<<grid, 256, 0, st>>>((const float*)R, ldr, k, n, ref_block, semantics, U, dvec);
    TQ_LAUNCH_CHECK();
    if (use_tc) {
      dim3 tg((unsigned)ceil_div(n, 32), (unsigned)ceil_div(kpad, 32));
      split_transpose_u_kernel<<<tg, dim3(32, 8), 0, st>>>(U, k, n, kpad, U_hi, U_lo);
      TQ_LAUNCH_CHECK();
    }
  }

  const size_t smem = size_t(kBlk) * kBlk * sizeof(float);
  TQ_CUDA_CHECK(cudaFuncSetAttribute(gptq_block_kernel<TQ_LOOP_TRITON>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TQ_CUDA_CHECK(cudaFuncSetAttribute(gptq_block_kernel<TQ_LOOP_TORCH>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TrailingTc tc;
  if (use_tc) TQ_TRY(trailing_tc_prepare(&tc, E_hi, E_lo, m, U_hi, U_lo, kpad, n));
  TQ_CUDA_CHECK(cudaMemsetAsync(E, 0, sizeof(float) * size_t(m) * kMacro, st));
  TQ_CUDA_CHECK(cudaMemsetAsync(E_hi, 0, sizeof(float) * size_t(m) * kMacro, st));
  TQ_CUDA_CHECK(cudaMemsetAsync(E_lo, 0, sizeof(float) * size_t(m) * kMacro, st));
  const unsigned row_ctas = (unsigned)ceil_div(m, kRowsPerWarp * kWarpsPerCta);
  // Two-level lazy batching: 128-column block steps update only the rest of their 1024-column
  // macro block (K = 128); the far columns are updated once per macro block with K = 1024.
  // Pairs (c, j) inside one REFERENCE block are updated with the reference's own arithmetic - the K rank-1 updates
  // one after another, each rounded at the magnitude of W (trailing_update_kernel<kSeqFma / kSeqMulSub>) - and
  // pairs across reference blocks as a product that is subtracted once (tcgen05 3xTF32 GEMM, or the SIMT
  // kSeqDelta kernel with TQ_LOOP_STRICT_FP32), like the reference's E @ Scale (gptq_utils.py:539-545).
  auto trailing = [&](int64_t j0, int64_t N, int64_t e_col0, int64_t u_row0, int kcount, int mode) -> int {
    if (N <= 0 || kcount <= 0) return TQ_OK;
    if (mode == kSeqDelta && use_tc) return trailing_tc_launch(&tc, Wp + j0, n, m, N, e_col0, u_row0, kcount, j0, st);
    dim3 grid((unsigned)ceil_div(N, kTN), (unsigned)ceil_div(m, kTM));
    const int pslot = prof_begin_launch(st, 2.0 * double(m) * double(N) * double(kcount), TQ_PROF_TRAILING_SEQ);
    float* Cj = Wp + j0;
    const float* Ej = E + e_col0;
    const float* Uj = U + u_row0 * n + j0;
    if (mode == kSeqFma)
      trailing_update_kernel<kSeqFma><<<grid, 256, 0, st>>>(Cj, n, Ej, kMacro, Uj, n, m, N, kcount);
    else if (mode == kSeqMulSub)
      trailing_update_kernel<kSeqMulSub><<<grid, 256, 0, st>>>(Cj, n, Ej, kMacro, Uj, n, m, N, kcount);
    else
      trailing_update_kernel<kSeqDelta><<<grid, 256, 0, st>>>(Cj, n, Ej, kMacro, Uj, n, m, N, kcount);
    prof_end_launch(st, pslot);
    TQ_LAUNCH_CHECK();
    return TQ_OK;
  };
  const int seq_mode = semantics == TQ_LOOP_TRITON ? kSeqFma : kSeqMulSub;
  // fused macro-block kernel whenever every pair inside a macro block lies inside one reference block
  const bool fused = (ref_block % kMacro == 0) && (n % 4 == 0);
  const size_t fsmem = sizeof(FusedSmem) + 16;
  const bool ftrace = fused && trace_enabled();
  if (ftrace) TQ_CUDA_CHECK(cudaMemsetAsync(fstats, 0, 8 * sizeof(long long), st));
  if (fused) {
    TQ_CUDA_CHECK(cudaFuncSetAttribute(gptq_macro_kernel<TQ_LOOP_TRITON>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)fsmem));
    TQ_CUDA_CHECK(cudaFuncSetAttribute(gptq_macro_kernel<TQ_LOOP_TORCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)fsmem));
  }
  for (int64_t M0 = 0; M0 < k; M0 += kMacro) {
    const int64_t M1 = imin(M0 + kMacro, k);
    if (fused) {
      const unsigned fgrid = (unsigned)ceil_div(m, kFRows);
      const int cntM = int(M1 - M0);
      const int bslot = prof_begin_launch(st, double(m) * cntM * 8.0, TQ_PROF_LOOP_BLOCK);
      if (semantics == TQ_LOOP_TRITON)
        gptq_macro_kernel<TQ_LOOP_TRITON><<<fgrid, kFThreads, fsmem, st>>>(
            Wp, n, U, dvec, scale, zero, ng, gidx, m, M0, cntM, min_q, max_q, E, use_tc ? E_hi : nullptr,
            use_tc ? E_lo : nullptr, Cp, ftrace ? fstats : nullptr);
      else
        gptq_macro_kernel<TQ_LOOP_TORCH><<<fgrid, kFThreads, fsmem, st>>>(
            Wp, n, U, dvec, scale, zero, ng, gidx, m, M0, cntM, min_q, max_q, E, use_tc ? E_hi : nullptr,
            use_tc ? E_lo : nullptr, Cp, ftrace ? fstats : nullptr);
      prof_end_launch(st, bslot);
      TQ_LAUNCH_CHECK();
    } else {
    for (int64_t c0 = M0; c0 < M1; c0 += kBlk) {
      const int cnt = int(imin(kBlk, M1 - c0));
      float* Eb = E + (c0 - M0);
      float* Ehb = use_tc ? E_hi + (c0 - M0) : nullptr;
      float* Elb = use_tc ? E_lo + (c0 - M0) : nullptr;
      const int bslot = prof_begin_launch(st, double(m) * cnt * 8.0, TQ_PROF_LOOP_BLOCK);
      if (semantics == TQ_LOOP_TRITON)
        gptq_block_kernel<TQ_LOOP_TRITON><<<row_ctas, kWarpsPerCta * 32, smem, st>>>(
            Wp, n, U, dvec, scale, zero, ng, gidx, m, c0, cnt, min_q, max_q, Eb, Ehb, Elb, Cp);
      else
        gptq_block_kernel<TQ_LOOP_TORCH><<<row_ctas, kWarpsPerCta * 32, smem, st>>>(
            Wp, n, U, dvec, scale, zero, ng, gidx, m, c0, cnt, min_q, max_q, Eb, Ehb, Elb, Cp);
      prof_end_launch(st, bslot);
      TQ_LAUNCH_CHECK();
      const int64_t j0 = c0 + cnt;
      // rest of the macro block: up to the end of the reference block sequentially, beyond it as a product
      const int64_t ref_end = imin((c0 / ref_block + 1) * int64_t(ref_block), M1);
      if (ref_end > j0) TQ_TRY(trailing(j0, ref_end - j0, c0 - M0, c0, cnt, seq_mode));
      const int64_t d0 = imax(j0, ref_end);
      TQ_TRY(trailing(d0, M1 - d0, c0 - M0, c0, cnt, kSeqDelta));
    }
    }
    // everything beyond the macro block: columns that still belong to the macro block's reference block
    // (ref_block > 1024) sequentially, the rest as one K = 1024 product
    const int64_t ref_end_far = imin((M0 / ref_block + 1) * int64_t(ref_block), k);
    if (ref_end_far > M1) TQ_TRY(trailing(M1, ref_end_far - M1, 0, M0, int(M1 - M0), seq_mode));
    const int64_t f0 = imax(M1, ref_end_far);
    TQ_TRY(trailing(f0, n - f0, 0, M0, int(M1 - M0), kSeqDelta));
  }
  if (ftrace) {
    long long hs[8];
    TQ_CUDA_CHECK(cudaMemcpyAsync(hs, fstats, sizeof(hs), cudaMemcpyDeviceToHost, st));
    TQ_CUDA_CHECK(cudaStreamSynchronize(st));
    fprintf(stderr, "[tq-trace] gptq_macro_kernel CTA 0, kcycles over %lld macro blocks: load %.0f  part A %.0f  quantiser %.0f  "
            "part B %.0f  workers waiting %.0f  store %.0f\n", (long long)ceil_div(k, kMacro), hs[0] * 1e-3, hs[1] * 1e-3,
            hs[2] * 1e-3, hs[3] * 1e-3, hs[4] * 1e-3, hs[5] * 1e-3);
  }
  if (k < n) {
    dim3 grid((unsigned)imin(ceil_div(n - k, 256), 64), (unsigned)imin(m, 65535));
    tail_rtn_kernel<<<grid, 256, 0, st>>>(Wp, n, m, k, scale, zero, ng, gidx, min_q, max_q, Cp);
    TQ_LAUNCH_CHECK();
  }
  {
    dim3 grid((unsigned)imin(ceil_div(n, 256), 64), (unsigned)imin(m, 65535));
    unpermute_kernel<<<grid, 256, 0, st>>>(Wp, Cp, m, n, invperm, Wq_out, ldq, codes_out, ldc);
    TQ_LAUNCH_CHECK();
  }
  return TQ_OK;
}

extern "C" int tq_pack_codes(const uint8_t* codes, int64_t ldc, int64_t m, int64_t n, int bits,
                             uint32_t* packed, int64_t ldp, void* stream) {
  TQ_TRY(check_device());
  TQ_REQUIRE(codes && packed && m > 0 && n > 0 && ldc >= n, "tq_pack_codes: bad arguments");
  TQ_REQUIRE(bits >= 1 && bits <= 8, "tq_pack_codes: bits=%d outside [1,8]", bits);
  int64_t nwords = (n * bits + 31) / 32;
  TQ_REQUIRE(ldp >= nwords, "tq_pack_codes: ldp %lld < %lld words", (long long)ldp, (long long)nwords);
  dim3 grid((unsigned)imin(ceil_div(nwords, 128), 64), (unsigned)imin(m, 65535));
  pack_codes_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(codes, ldc, m, n, bits, packed, ldp, nwords);
  TQ_LAUNCH_CHECK();
  return TQ_OK;
}
