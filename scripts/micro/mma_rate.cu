// Microbenchmark: sustained issue rate of tcgen05.mma (kind::f16, M=128, K=16) for N in {64,128,256},
// operands from shared memory (SS) or A from tensor memory (TS).  One CTA per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../glue-factory_b200/csrc/common.cuh"
namespace lgb { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }
using namespace lgb;

// AMN / BMN: the operand is MN-major in shared memory (as the dQ = dS K MMA of the fused attention backward reads both)
template <int N, bool TS, bool AMN = false, bool BMN = false>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (32768 + N * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, AMN ? 1 : 0, BMN ? 1 : 0);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t db = BMN ? make_smem_desc(b + kk * 2048, 8192, 1024) : make_smem_desc(b + kk * 32, 16, 1024);
        const uint64_t da = AMN ? make_smem_desc(a + kk * 2048, 16384, 1024) : make_smem_desc(a + kk * 32, 16, 1024);
        if (TS) umma_bf16_ts(tb + 256, tb + kk * 8, db, idesc, 1u);
        else umma_bf16(tb + 256, da, db, idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N, bool TS, bool AMN = false, bool BMN = false> void run(const char* name) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 2048, smem = 32768 + N * 128 + 1024;
  cudaFuncSetAttribute(k<N, TS, AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, TS, AMN, BMN><<<148, 128, smem>>>(d, iters); cudaDeviceSynchronize();
  k<N, TS, AMN, BMN><<<148, 128, smem>>>(d, iters); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * 4);
  printf("%-10s N=%3d : %.1f clk/MMA  -> %.0f FLOP/clk/SM (%s)\n", name, N, per, 2.0 * 128 * N * 16 / per, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main() {
  run<64, false>("SS"); run<128, false>("SS"); run<256, false>("SS");
  run<64, true>("TS"); run<128, true>("TS"); run<256, true>("TS");
  run<64, false, false, true>("SS A-K B-MN"); run<64, false, true, false>("SS A-MN B-K"); run<64, false, true, true>("SS A-MN B-MN");
  run<64, true, false, true>("TS B-MN");
  return 0;
}
