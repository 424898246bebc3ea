// Batched bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulator in TMEM, operands
// staged by TMA into 128-byte-swizzled shared memory).  Used for the assignment similarity
// sim = mdesc0 . mdesc1^T (lightglue.py:283) and its two backward contractions, which need every
// combination of K-major / MN-major operands.
//
// One CTA computes a 128x128 tile of C.  Warp roles: warps 0-3 epilogue (TMEM -> registers ->
// global), warp 4 TMA producer (one elected lane), warp 5 MMA issuer (one elected lane) + TMEM owner.
// 4-stage smem ring, 64-wide K blocks (one 128-byte swizzle atom per operand row).
#include "common.cuh"
#include "host_util.h"
#include "lgb200.h"

namespace lgb {

constexpr int GB_M = 128, GB_N = 128, GB_K = 64, G_STAGES = 4;
constexpr int G_TILE = GB_M * GB_K * 2;  // 16 KiB per operand per stage
constexpr int G_SMEM = G_STAGES * 2 * G_TILE + 256;

template <bool A_MN, bool B_MN, typename OutT>
__global__ void __launch_bounds__(192, 1)
    gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     OutT* __restrict__ C, int M, int N, int K, int64_t ldc, int64_t strideC, float alpha) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + G_STAGES * G_TILE;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * G_STAGES * G_TILE);
  uint64_t* empty = full + G_STAGES;
  uint64_t* done = empty + G_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N, b = blockIdx.z;
  const int nk = (K + GB_K - 1) / GB_K;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 5) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % G_STAGES;
        const uint32_t ph = (kb / G_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], 2 * G_TILE);
        uint8_t* a = sA + s * G_TILE;
        uint8_t* bb = sB + s * G_TILE;
        if (!A_MN) {
          tma_load_3d(a, &tmA, &full[s], kb * GB_K, m0, b);
        } else {
          tma_load_3d(a, &tmA, &full[s], m0, kb * GB_K, b);
          tma_load_3d(a + 8192, &tmA, &full[s], m0 + 64, kb * GB_K, b);
        }
        if (!B_MN) {
          tma_load_3d(bb, &tmB, &full[s], kb * GB_K, n0, b);
        } else {
          tma_load_3d(bb, &tmB, &full[s], n0, kb * GB_K, b);
          tma_load_3d(bb + 8192, &tmB, &full[s], n0 + 64, kb * GB_K, b);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GB_M, GB_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % G_STAGES;
        const uint32_t ph = (kb / G_STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < GB_K / 16; ++kk) {
          const uint64_t ad = A_MN ? make_smem_desc(a0 + s * G_TILE + kk * 2048, 8192, 1024)
                                   : make_smem_desc(a0 + s * G_TILE + kk * 32, 16, 1024);
          const uint64_t bd = B_MN ? make_smem_desc(b0 + s * G_TILE + kk * 2048, 8192, 1024)
                                   : make_smem_desc(b0 + s * G_TILE + kk * 32, 16, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (kb | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(done);
    }
  } else {
    mbar_wait(done, 0);
    tc_fence_after();
    const int row = m0 + warp * 32 + lane;
    OutT* crow = C + (int64_t)b * strideC + (int64_t)row * ldc + n0;
    const bool vec_ok = ((ldc * sizeof(OutT)) % 16 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                        ((strideC * sizeof(OutT)) % 16 == 0);
#pragma unroll 1
    for (int c = 0; c < GB_N / 32; ++c) {
      float v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] *= alpha;
      if (row < M) {
        const int col0 = n0 + c * 32;
        if (vec_ok && col0 + 32 <= N) {
          if constexpr (sizeof(OutT) == 4) {
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              *reinterpret_cast<float4*>(crow + c * 32 + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < 32; e += 8) {
              uint4 u;
              u.x = pack_bf16(v[e], v[e + 1]); u.y = pack_bf16(v[e + 2], v[e + 3]);
              u.z = pack_bf16(v[e + 4], v[e + 5]); u.w = pack_bf16(v[e + 6], v[e + 7]);
              *reinterpret_cast<uint4*>(crow + c * 32 + e) = u;
            }
          }
        } else {
          for (int e = 0; e < 32; ++e)
            if (col0 + e < N) {
              if constexpr (sizeof(OutT) == 4) crow[c * 32 + e] = v[e];
              else crow[c * 32 + e] = __float2bfloat16(v[e]);
            }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 128);
}

template <bool A_MN, bool B_MN, typename OutT>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, void* C, int batch, int M, int N, int K,
                       int64_t ldc, int64_t strideC, float alpha, cudaStream_t stream) {
  auto kern = gemm_bf16_kernel<A_MN, B_MN, OutT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
    LGB_REQUIRE(e == cudaSuccess, kErrCuda, "gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid((N + GB_N - 1) / GB_N, (M + GB_M - 1) / GB_M, batch);
  kern<<<grid, 192, G_SMEM, stream>>>(ta, tb, static_cast<OutT*>(C), M, N, K, ldc, strideC, alpha);
  return check_launch("gemm_bf16");
}

}  // namespace lgb

using namespace lgb;

extern "C" int lgb200_gemm_bf16(const void* A, const void* B, void* C, int batch, int M, int N, int K, int a_mn_major,
                                int b_mn_major, int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA,
                                int64_t strideB, int64_t strideC, int c_dtype, float alpha, cudaStream_t stream) {
  LGB_REQUIRE(A && B && C, kErrInvalid, "gemm_bf16: null pointer");
  LGB_REQUIRE(batch > 0 && M > 0 && N > 0 && K > 0, kErrInvalid, "gemm_bf16: empty problem %dx%dx%dx%d", batch, M, N, K);
  LGB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, kErrInvalid, "gemm_bf16: lda/ldb must be multiples of 8 elements");
  LGB_REQUIRE(batch == 1 || (strideA % 8 == 0 && strideB % 8 == 0), kErrInvalid,
              "gemm_bf16: batch strides must be multiples of 8 elements");
  CUtensorMap ta, tb;
  {
    // A: K-major -> dims {K, M, batch}; MN-major (stored [K,M]) -> dims {M, K, batch}
    const uint64_t inner = a_mn_major ? (uint64_t)M : (uint64_t)K, outer = a_mn_major ? (uint64_t)K : (uint64_t)M;
    const uint64_t dims[3] = {inner, outer, (uint64_t)batch};
    const uint64_t str[2] = {(uint64_t)lda * 2, (uint64_t)(batch > 1 ? strideA : (int64_t)outer * lda) * 2};
    const uint32_t box[3] = {64, a_mn_major ? 64u : 128u, 1};
    int rc = make_tmap_bf16(&ta, A, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t inner = b_mn_major ? (uint64_t)N : (uint64_t)K, outer = b_mn_major ? (uint64_t)K : (uint64_t)N;
    const uint64_t dims[3] = {inner, outer, (uint64_t)batch};
    const uint64_t str[2] = {(uint64_t)ldb * 2, (uint64_t)(batch > 1 ? strideB : (int64_t)outer * ldb) * 2};
    const uint32_t box[3] = {64, b_mn_major ? 64u : 128u, 1};
    int rc = make_tmap_bf16(&tb, B, 3, dims, str, box);
    if (rc) return rc;
  }
#define LGB_GEMM_CASE(AM, BMJ)                                                                              \
  if (a_mn_major == AM && b_mn_major == BMJ) {                                                              \
    if (c_dtype == LGB200_F32) return launch_gemm<AM, BMJ, float>(ta, tb, C, batch, M, N, K, ldc, strideC, alpha, stream); \
    if (c_dtype == LGB200_BF16)                                                                             \
      return launch_gemm<AM, BMJ, __nv_bfloat16>(ta, tb, C, batch, M, N, K, ldc, strideC, alpha, stream);  \
  }
  LGB_GEMM_CASE(0, 0)
  LGB_GEMM_CASE(0, 1)
  LGB_GEMM_CASE(1, 0)
  LGB_GEMM_CASE(1, 1)
#undef LGB_GEMM_CASE
  LGB_REQUIRE(false, kErrInvalid, "gemm_bf16: bad major/dtype flags");
}
