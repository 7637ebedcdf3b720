// Host emulation runtime for the CUDA sources that tests/emu compiles with g++ (test infrastructure, CPU only).
// One OS thread per CUDA thread; __syncthreads() is a pthread barrier per CTA, __shfl_xor_sync() an exchange buffer
// per warp, acquire / release accesses are GCC atomics.  Kernels are launched through TQ_LAUNCH (CTAs one after
// another) or emu_launch_concurrent (all CTAs at once, for the cooperative bulge-chase kernel).
#pragma once
#define TQ_HOST_EMU 1
#include <cuda_runtime.h>   // types only (dim3, double2, cudaStream_t ...): nothing of libcudart is linked
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>

#include <algorithm>
#include <functional>
#include <thread>
#include <vector>

#undef __launch_bounds__
#define __launch_bounds__(...)

static thread_local dim3 threadIdx, blockIdx, gridDim, blockDim;

struct EmuCta {
  pthread_barrier_t bar;
  pthread_barrier_t cbar;                      // compute_sync(): the first 256 threads only
  std::vector<pthread_barrier_t> wbar;
  std::vector<double> shfl;
  std::vector<double> smem;
  int nthreads = 0;
  void init(int threads, size_t smem_bytes) {
    nthreads = threads;
    pthread_barrier_init(&bar, nullptr, threads);
    pthread_barrier_init(&cbar, nullptr, std::min(threads, 256));
    const int nw = (threads + 31) / 32;
    wbar.resize(nw);
    for (int w = 0; w < nw; ++w) pthread_barrier_init(&wbar[w], nullptr, std::min(32, threads - 32 * w));
    shfl.assign(size_t(nw) * 32, 0.0);
    smem.assign((smem_bytes + 7) / 8 + 1, NAN);          // uninitialised shared memory must never be consumed
  }
  void destroy() {
    pthread_barrier_destroy(&bar);
    pthread_barrier_destroy(&cbar);
    for (auto& b : wbar) pthread_barrier_destroy(&b);
  }
};
static thread_local EmuCta* g_cta = nullptr;

#define TQ_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(g_cta->smem.data())

static inline void __syncthreads() { pthread_barrier_wait(&g_cta->bar); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline double __ldcg(const double* p) { return *reinterpret_cast<const volatile double*>(p); }
static inline void __stcg(double* p, double v) { *reinterpret_cast<volatile double*>(p) = v; }
static inline long long clock64() { return 0; }
static inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }
static inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int ld_acquire_s32(const int* p) {
  const int v = __atomic_load_n(p, __ATOMIC_ACQUIRE);
  sched_yield();
  return v;
}
static inline void st_release_s32(int* p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline void st_release_cta_smem(int* p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline int ld_acquire_cta_smem(const int* p) {
  const int v = __atomic_load_n(p, __ATOMIC_ACQUIRE);
  sched_yield();
  return v;
}
// named barrier over the first `compute_threads` threads of the CTA (the chase kernel's helper warp stays out)
static inline void compute_sync() { pthread_barrier_wait(&g_cta->cbar); }
using std::max;
using std::min;

// all lanes of the (full) warp call this together, like the converged shuffle it stands for
static inline double __shfl_xor_sync(unsigned, double v, int o) {
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  double* buf = g_cta->shfl.data() + w * 32;
  buf[lane] = v;
  pthread_barrier_wait(&g_cta->wbar[w]);
  const double r = buf[lane ^ o];
  pthread_barrier_wait(&g_cta->wbar[w]);
  return r;
}

// run `body` as grid x block CUDA threads; concurrent = every CTA at the same time (co-resident kernels)
static inline void emu_run(dim3 grid, dim3 block, size_t smem_bytes, bool concurrent, const std::function<void()>& body) {
  const int T = int(block.x);
  const int nctas = int(grid.x * grid.y * grid.z);
  auto run_ctas = [&](int c0, int c1) {
    std::vector<EmuCta> ctas(c1 - c0);
    for (auto& c : ctas) c.init(T, smem_bytes);
    std::vector<std::thread> th;
    th.reserve(size_t(c1 - c0) * T);
    for (int c = c0; c < c1; ++c)
      for (int t = 0; t < T; ++t)
        th.emplace_back([&, c, t] {
          g_cta = &ctas[c - c0];
          gridDim = grid, blockDim = block;
          blockIdx = dim3(c % grid.x, (c / grid.x) % grid.y, c / (grid.x * grid.y));
          threadIdx = dim3(t, 0, 0);
          body();
        });
    for (auto& t : th) t.join();
    for (auto& c : ctas) c.destroy();
  };
  if (concurrent) {
    run_ctas(0, nctas);
  } else {
    for (int c = 0; c < nctas; ++c) run_ctas(c, c + 1);
  }
}

// barrier-free kernels: every CUDA thread in turn on the calling OS thread (fast)
static inline void emu_run_serial(dim3 grid, dim3 block, const std::function<void()>& body) {
  gridDim = grid, blockDim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned tx = 0; tx < block.x; ++tx) {
          blockIdx = dim3(bx, by, bz);
          threadIdx = dim3(tx, 0, 0);
          body();
        }
}

// kernels that contain barriers or shuffles are listed by name; everything else runs serially
#define EMU_NEEDS_THREADS(kernel) (std::string(#kernel) == "larft_kernel" || std::string(#kernel) == "mirror_lower_colmajor_kernel")
#include <string>
#define TQ_LAUNCH(kernel, grid, block, smem, stream, ...)                                      \
  do {                                                                                         \
    if (EMU_NEEDS_THREADS(kernel))                                                             \
      emu_run(dim3(grid), dim3(block), (smem), false, [&] { kernel(__VA_ARGS__); });           \
    else                                                                                       \
      emu_run_serial(dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); });                   \
  } while (0)
