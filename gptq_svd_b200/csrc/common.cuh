// Shared host/device helpers for libtruncgptq (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/truncgptq.h"

namespace tq {

void set_error(const char* fmt, ...);
int check_device();  // TQ_OK when the current device is sm_100 (B200)

#define TQ_CUDA_CHECK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      tq::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return TQ_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define TQ_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      tq::set_error(__VA_ARGS__);  \
      return TQ_ERR_INVALID;       \
    }                              \
  } while (0)

#define TQ_TRY(expr)            \
  do {                          \
    int _s = (expr);            \
    if (_s != TQ_OK) return _s; \
  } while (0)

// Kernel launch and dynamic shared memory go through these two macros in the files that tests/emu compiles for
// the HOST (TQ_HOST_EMU: the emulation defines its own versions before including this header).
#ifndef TQ_HOST_EMU
#define TQ_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define TQ_DYN_SMEM(type, name) extern __shared__ type name[]
#endif

extern thread_local int64_t g_launch_count;
#define TQ_LAUNCH_CHECK()                \
  do {                                   \
    ++tq::g_launch_count;                \
    TQ_CUDA_CHECK(cudaGetLastError());   \
  } while (0)

// sampled event timing of the library's own hot kernels (see tq_profile_begin / tq_profile_kernel); `kind` is one
// of the TQ_PROF_* ids of truncgptq.h, `work` the algorithmic bytes or flops of the launch
int prof_begin_launch(cudaStream_t st, double work, int kind = 0);   // returns a slot or -1
void prof_end_launch(cudaStream_t st, int slot);

static inline int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }
static inline int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Bump allocator over the caller-provided workspace.
struct Workspace {
  char* base;
  size_t size;
  size_t off;
  bool overflow;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), off(0), overflow(false) {}
  template <typename T>
  T* take(size_t count) {
    size_t a = align_up(off, 256);
    size_t need = count * sizeof(T);
    if (base == nullptr || a + need > size) {
      overflow = true;
      off = a + need;
      return nullptr;
    }
    off = a + need;
    return reinterpret_cast<T*>(base + a);
  }
};

static inline size_t ws_bytes_for(size_t count, size_t elem) { return align_up(count * elem, 256) + 256; }

int num_sms();
// index of the current device for the per-device one-shot caches (function attributes, occupancy): a host
// thread may work on several GPUs, and cudaFuncSetAttribute / occupancy results are per device
constexpr int kMaxDevices = 64;
static inline int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev % kMaxDevices;
}

// optional per-thread stage callback (tq_set_stage_callback)
bool stage_callback_set();
void notify_stage(int stage);

// TQ_TRACE=1: print per-stage milliseconds (synchronises the stream at stage boundaries)
bool trace_enabled();
void trace_flush(cudaStream_t st);     // TQ_TRACE=2: print the event-timed stages of the call that just ended
struct StageTimer {
  cudaStream_t st;
  const char* name;
  double t0;
  StageTimer(cudaStream_t s, const char* n);
  ~StageTimer();
};

}  // namespace tq
