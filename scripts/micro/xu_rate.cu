// Microbenchmark: per-SM throughput of MUFU.EX2, F2FP (cvt.rn.bf16x2.f32), FFMA and SHFL with 16 resident warps.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
template <int OP>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* clk, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = 0.001f * (threadIdx.x + i);
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 1) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[(i + 1) & 15])); acc ^= r; }
      if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
      if (OP == 3) x[i] = __shfl_sync(0xffffffffu, x[i], (i + it) & 31);
      if (OP == 5 && (i & 1) == 0) { float2 a = make_float2(x[i], x[i + 1]); a = __ffma2_rn(a, a, a); x[i] = a.x; x[i + 1] = a.y; }
      if (OP == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[(i + 1) & 15])); acc ^= r; }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int OP> void run(const char* name, int per) {
  float* o; long long* c; cudaMalloc(&o, 148 * 512 * 4); cudaMalloc(&c, 8);
  const int iters = 4096;
  k<OP><<<148, 512>>>(o, c, iters); cudaDeviceSynchronize();
  k<OP><<<148, 512>>>(o, c, iters); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  double thread_ops = (double)iters * 16 * 512 * per;  // per SM
  printf("%-22s %.2f thread-ops/clk/SM  (%s)\n", name, thread_ops / h, cudaGetErrorString(cudaGetLastError()));
  cudaFree(o); cudaFree(c);
}
int main() {
  run<0>("MUFU.EX2", 1); run<1>("F2FP.BF16x2 (per instr)", 1); run<2>("FFMA", 1); run<3>("SHFL", 1); run<4>("EX2 + F2FP pair", 1);
  run<5>("FFMA2 (instr; 2 FMA each)", 1);
  return 0;
}
