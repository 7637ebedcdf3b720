// Host emulation of the KERNELS in gptq_svd_b200/csrc/two_stage_kernels.cuh (test infrastructure, CPU only).
//
// The kernel SOURCE TEXT is compiled for the host on the emulation runtime (emu_runtime.h): one OS thread per CUDA
// thread, pthread barriers for __syncthreads() / the named compute barrier, an exchange buffer per warp for the
// shuffles, GCC atomics for the acquire / release progress counters and the helper warp's ticket.  CTAs really run
// concurrently, so the inter-sweep protocol of the bulge chase is exercised, not only its index arithmetic.
// tests/test_two_stage_emu.py builds this file with g++ and compares the kernels' outputs with the numpy model
// scripts/prototypes/sb2st_band.py.  (two_stage_host_emu.cpp does the same for the host driver around them.)
#include "emu_runtime.h"

namespace tq {
// as in solver_kernels.cuh
static inline double warp_sum(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
}  // namespace tq

#include "../../gptq_svd_b200/csrc/two_stage_kernels.cuh"

extern "C" {

void emu_constants(int* out) {
  out[0] = tq::kBw;
  out[1] = tq::kLdb;
  out[2] = tq::kQ2H;
  out[3] = tq::kQ2Ld;
  out[4] = tq::kChaseThreads;
}

void emu_band_extract(const double* A, int64_t lda, int n, double* Bd) {
  emu_run_serial(dim3(n), dim3(128), [&] { tq::band_extract_kernel(A, lda, n, Bd); });
}

void emu_band_diag(const double* Bd, int n, double* d, double* e) {
  emu_run_serial(dim3((n + 255) / 256), dim3(256), [&] { tq::band_diag_kernel(Bd, n, d, e); });
}

void emu_copy_staircase(const double* Vs, int64_t ldv, const double* tau2, int n, int sb0, int k0, int count,
                        double* Vc, double* taub) {
  emu_run_serial(dim3(1, tq::kBw, count), dim3(tq::kQ2Ld),
                 [&] { tq::copy_staircase_kernel(Vs, ldv, tau2, n, sb0, k0, Vc, taub); });
}

// the persistent bulge-chase kernel on `grid` concurrently running CTAs; variant bit 0: the ninth warp publishes the
// progress counters, bit 1: second wait in front of the D / E loads.  The cycle counters are switched on to
// exercise the instrumented path too.
void emu_chase(double* Bd, int n, double* Vs, int64_t ldv, double* tau2, int* prog, int grid, int variant) {
  long long stats[8] = {0};
  tq::ChaseArgs args{Bd, n, Vs, ldv, tau2, prog, stats};
  const dim3 g(grid), t256(tq::kChaseThreads), t288(tq::kChaseThreads + 32);
  switch (variant & 3) {
    case 0: emu_run(g, t256, tq::kChaseSmem, true, [&] { tq::sb2st_chase_kernel_t<false, false>(args); }); break;
    case 1: emu_run(g, t288, tq::kChaseSmem, true, [&] { tq::sb2st_chase_kernel_t<true, false>(args); }); break;
    case 2: emu_run(g, t256, tq::kChaseSmem, true, [&] { tq::sb2st_chase_kernel_t<false, true>(args); }); break;
    default: emu_run(g, t288, tq::kChaseSmem, true, [&] { tq::sb2st_chase_kernel_t<true, true>(args); }); break;
  }
}
}
