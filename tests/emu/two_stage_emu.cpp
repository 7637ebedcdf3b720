// Host emulation of gptq_svd_b200/csrc/two_stage_kernels.cuh (test infrastructure, CPU only).
//
// The kernel SOURCE TEXT is compiled for the host: one OS thread per CUDA thread, a pthread barrier per CTA for
// __syncthreads(), a per-warp exchange buffer for the shuffle reduction, GCC atomics for the acquire / release
// progress counters.  CTAs really run concurrently, so the inter-sweep protocol of the bulge chase is exercised,
// not only its index arithmetic.  tests/test_two_stage_emu.py builds this file with g++ and compares the kernels'
// outputs with the numpy model scripts/prototypes/sb2st_band.py.
#define TQ_HOST_EMU 1
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>

#include <algorithm>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct EmuDim3 {
  int x = 0, y = 0, z = 0;
};
static thread_local EmuDim3 threadIdx, blockIdx, gridDim, blockDim;

struct EmuCta {
  pthread_barrier_t bar;
  pthread_barrier_t wbar[32];
  std::vector<double> shfl;
  std::vector<double> smem;
};
static thread_local EmuCta* g_cta = nullptr;

#define TQ_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(g_cta->smem.data())

static inline void __syncthreads() { pthread_barrier_wait(&g_cta->bar); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline double __ldcg(const double* p) { return *reinterpret_cast<const volatile double*>(p); }
static inline void __stcg(double* p, double v) { *reinterpret_cast<volatile double*>(p) = v; }
static inline int ld_acquire_s32(const int* p) {
  const int v = __atomic_load_n(p, __ATOMIC_ACQUIRE);
  sched_yield();
  return v;
}
static inline void st_release_s32(int* p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline long long clock64() { return 0; }
using std::min;

namespace tq {
// xor-butterfly over the 32 lanes of a warp, same association order as the __shfl_xor_sync loop on the GPU
static inline double warp_sum(double v) {
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  double* buf = g_cta->shfl.data() + w * 32;
  buf[lane] = v;
  pthread_barrier_wait(&g_cta->wbar[w]);
  double t[32], u[32];
  for (int l = 0; l < 32; ++l) t[l] = buf[l];
  for (int o = 16; o; o >>= 1) {
    for (int l = 0; l < 32; ++l) u[l] = t[l] + t[l ^ o];
    for (int l = 0; l < 32; ++l) t[l] = u[l];
  }
  pthread_barrier_wait(&g_cta->wbar[w]);
  return t[lane];
}
}  // namespace tq

#include "../../gptq_svd_b200/csrc/two_stage_kernels.cuh"

// kernels without barriers: every (block, thread) in turn on the calling thread
template <class F>
static void run_serial(EmuDim3 grid, EmuDim3 block, F f) {
  gridDim = grid;
  blockDim = block;
  for (int bz = 0; bz < std::max(1, grid.z); ++bz)
    for (int by = 0; by < std::max(1, grid.y); ++by)
      for (int bx = 0; bx < grid.x; ++bx)
        for (int tx = 0; tx < block.x; ++tx) {
          blockIdx.x = bx, blockIdx.y = by, blockIdx.z = bz;
          threadIdx.x = tx;
          f();
        }
}

extern "C" {

void emu_constants(int* out) {
  out[0] = tq::kBw;
  out[1] = tq::kLdb;
  out[2] = tq::kQ2H;
  out[3] = tq::kQ2Ld;
  out[4] = tq::kChaseThreads;
}

void emu_band_extract(const double* A, int64_t lda, int n, double* Bd) {
  EmuDim3 g, b;
  g.x = n, g.y = g.z = 1, b.x = 128;
  run_serial(g, b, [&] { tq::band_extract_kernel(A, lda, n, Bd); });
}

void emu_band_diag(const double* Bd, int n, double* d, double* e) {
  EmuDim3 g, b;
  g.x = (n + 255) / 256, g.y = g.z = 1, b.x = 256;
  run_serial(g, b, [&] { tq::band_diag_kernel(Bd, n, d, e); });
}

void emu_copy_staircase(const double* Vs, int64_t ldv, const double* tau2, int n, int sb0, int k0, int count,
                        double* Vc, double* taub) {
  EmuDim3 g, b;
  g.x = 1, g.y = tq::kBw, g.z = count, b.x = tq::kQ2Ld;
  run_serial(g, b, [&] { tq::copy_staircase_kernel(Vs, ldv, tau2, n, sb0, k0, Vc, taub); });
}

// the persistent bulge-chase kernel on `grid` concurrently running CTAs of kChaseThreads OS threads each
void emu_chase(double* Bd, int n, double* Vs, int64_t ldv, double* tau2, int* prog, int grid) {
  const int T = tq::kChaseThreads;
  std::vector<EmuCta> ctas(grid);
  for (auto& c : ctas) {
    pthread_barrier_init(&c.bar, nullptr, T);
    for (int w = 0; w < T / 32; ++w) pthread_barrier_init(&c.wbar[w], nullptr, 32);
    c.shfl.assign(T, 0.0);
    c.smem.assign(tq::kChaseSmemDoubles, NAN);            // uninitialised shared memory must never be consumed
  }
  long long stats[8] = {0};
  tq::ChaseArgs args{Bd, n, Vs, ldv, tau2, prog, stats};     // exercises the instrumented path too
  std::vector<std::thread> th;
  th.reserve(size_t(grid) * T);
  for (int bx = 0; bx < grid; ++bx)
    for (int tx = 0; tx < T; ++tx)
      th.emplace_back([&, bx, tx] {
        g_cta = &ctas[bx];
        gridDim.x = grid, blockDim.x = T;
        blockIdx.x = bx, threadIdx.x = tx;
        tq::sb2st_chase_kernel(args);
      });
  for (auto& t : th) t.join();
  for (auto& c : ctas) {
    pthread_barrier_destroy(&c.bar);
    for (int w = 0; w < T / 32; ++w) pthread_barrier_destroy(&c.wbar[w]);
  }
}
}
