// Microbenchmark: per-SM throughput of tcgen05.ld / tcgen05.st (32x32b shapes) with 4, 8 or 16 resident warps.
// Every warp w touches TMEM lanes 32*(w%4)..+31 (the only lanes it may access), columns rotating over the allocation.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../glue-factory_b200/csrc/common.cuh"
using namespace lgb;

template <int X>
__device__ __forceinline__ void ld_x(uint32_t addr, float* r);
template <>
__device__ __forceinline__ void ld_x<16>(uint32_t addr, float* r) { tmem_ld16(addr, r); }
template <>
__device__ __forceinline__ void ld_x<32>(uint32_t addr, float* r) { tmem_ld32(addr, r); }
template <>
__device__ __forceinline__ void ld_x<64>(uint32_t addr, float* r) { tmem_ld32(addr, r); tmem_ld32(addr + 32, r + 32); }

template <int OP, int X>  // OP 0: ld.x{X}, 1: st.x8 / st.x16 (X = 8 or 16)
__global__ void __launch_bounds__(512, 1) k(float* out, long long* clk, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  float acc = 0.f;
  float r[64];
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = ((it + (warp >> 2)) * 64) & 511;
    if (OP == 0) {
      ld_x<X>(base + (col & (512 - X)), r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < X; i += 8) acc += r[i];
    } else {
      if (X == 8) tmem_st8(base + col, w); else tmem_st16(base + col, w);
      tmem_st_wait();
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}
template <int OP, int X>
void run(const char* name, int nwarps) {
  float* o; long long* c; cudaMalloc(&o, 148 * 512 * 4); cudaMalloc(&c, 8);
  const int iters = 4096;
  k<OP, X><<<148, nwarps * 32>>>(o, c, iters); cudaDeviceSynchronize();
  k<OP, X><<<148, nwarps * 32>>>(o, c, iters); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  const double bytes = (double)iters * nwarps * 32 * X * 4;  // per SM
  printf("%-12s %2d warps: %7.1f B/clk/SM   %6.1f clk/instr/warp  (%s)\n", name, nwarps, bytes / h, (double)h / iters,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(o); cudaFree(c);
}
int main() {
  for (int nw : {4, 8, 16}) {
    run<0, 16>("ld.x16", nw); run<0, 32>("ld.x32", nw); run<0, 64>("ld.2*x32", nw);
    run<1, 8>("st.x8", nw); run<1, 16>("st.x16", nw);
  }
  return 0;
}
