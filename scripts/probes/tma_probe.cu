// Standalone probe: which fp64 TMA tile loads work on sm_100a (dtype, coordinate alignment, OOB).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu   Run: ./tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, int x, int y, int box_bytes, double* out, int ndoubles) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  double* tile = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sm) + 127) & ~uintptr_t(127));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(box_bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  for (int i = threadIdx.x; i < ndoubles; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int n = 300;
  std::vector<double> h(size_t(n) * n);
  for (int c = 0; c < n; ++c) for (int r = 0; r < n; ++r) h[r + size_t(c) * n] = r + 1000.0 * c;
  double *d, *out;
  cudaMalloc(&d, sizeof(double) * n * n);
  cudaMemcpy(d, h.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice);
  const int BR = 128, BC = 16;
  cudaMalloc(&out, sizeof(double) * BR * BC);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)p;
  CUtensorMap tm;
  bool f32view = variant >= 3 && variant <= 5;
  bool u64 = variant >= 6;
  cuuint64_t dims[2] = {cuuint64_t(f32view ? 2 * n : n), cuuint64_t(n)};
  cuuint64_t strides[1] = {cuuint64_t(n) * 8};
  cuuint32_t box[2] = {cuuint32_t(f32view ? 2 * BR : BR), BC};
  cuuint32_t es[2] = {1, 1};
  CUtensorMapDataType dt = f32view ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (u64 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64);
  CUresult r = fn(&tm, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d encode -> %d\n", variant, int(r));
  int xr = 0, yc = 0;                       // row / column coordinates in doubles
  switch (variant % 3) { case 0: xr = 0; yc = 0; break; case 1: xr = 1; yc = 1; break; case 2: xr = 257; yc = 290; break; }
  int x = f32view ? 2 * xr : xr;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, BR * BC * 8 + 256);
  probe<<<1, 128, BR * BC * 8 + 256>>>(tm, x, yc, BR * BC * 8, out, BR * BC);
  cudaError_t e = cudaDeviceSynchronize();
  printf("  launch -> %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<double> ho(BR * BC);
  cudaMemcpy(ho.data(), out, sizeof(double) * BR * BC, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int c = 0; c < BC; ++c) for (int rr = 0; rr < BR; ++rr) {
    int gr = xr + rr, gc = yc + c;
    double want = (gr < n && gc < n) ? h[gr + size_t(gc) * n] : 0.0;
    if (ho[rr + c * BR] != want) ++bad;
  }
  printf("  mismatches: %d of %d\n", bad, BR * BC);
  return 0;
}
