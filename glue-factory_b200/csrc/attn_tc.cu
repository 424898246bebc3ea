// Flash attention on the 5th-generation tensor cores (sm_100a): out = softmax(scale * q k^T) v
// for head_dim 64, bf16 operands, fp32 accumulation, never materialising the N x N similarity.
// (reference: lightglue.py:118-121 self-attention; :207-216 cross-attention.)
//
// FORWARD.  One CTA = one (batch, head, 128-query tile).  Q/K/V tiles are brought in by TMA
// (4-D tensor map over the token-major [B,N,H,64] layout, 128-byte swizzle) and consumed straight
// from shared memory by tcgen05.mma:
//     S (128 x 128 fp32, TMEM cols [0,128))   = Q K_j^T        4 MMAs  (M128 N128 K16)
//     O_j (128 x 64 fp32, TMEM cols [128,192)) = P_j V_j        8 MMAs  (M128 N64  K16)
// Warps 0-3 are the softmax warpgroup (thread == query row == TMEM lane): two sweeps over S with
// tcgen05.ld (row max, then exp2 / row sum / bf16 pack), P_j written to shared memory in the
// K-major 128B-swizzled layout the MMA expects, O_j folded into a register accumulator with the
// usual online-softmax rescale.  Warp 4 = TMA producer, warp 5 = MMA issuer + TMEM owner.
// Two CTAs fit per SM (112 KiB smem, 256 TMEM columns each), so one CTA's MMAs overlap the
// other's softmax.
#include <math.h>

#include "common.cuh"
#include "host_util.h"
#include "lgb200.h"

namespace lgb {

constexpr int FA_BM = 128;     // queries per CTA
constexpr int FA_BN = 128;     // keys per iteration
constexpr int FA_D = 64;
constexpr int FA_STAGES = 2;
constexpr int FA_QBYTES = FA_BM * FA_D * 2;   // 16 KiB
constexpr int FA_KBYTES = FA_BN * FA_D * 2;   // 16 KiB
constexpr int FA_PBYTES = FA_BM * FA_BN * 2;  // 32 KiB
constexpr int FA_SMEM = FA_QBYTES + FA_STAGES * 2 * FA_KBYTES + FA_PBYTES + 256;
constexpr int FA_TMEM_COLS = 256;
constexpr int FA_S_COL = 0, FA_O_COL = 128;

__global__ void __launch_bounds__(192, 2)
    attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out,
                       float* __restrict__ lse, int B, int Nq, int Nk, int H, int kv_shift, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + FA_QBYTES;
  uint8_t* sV = sK + FA_STAGES * FA_KBYTES;
  uint8_t* sP = sV + FA_STAGES * FA_KBYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + FA_PBYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;               // [FA_STAGES]
  uint64_t* kv_empty = kv_full + FA_STAGES;   // [FA_STAGES]
  uint64_t* s_full = kv_empty + FA_STAGES;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * FA_BM, h = blockIdx.y, b = blockIdx.z;
  const int kb = (b + kv_shift) % B;
  const int ntiles = (Nk + FA_BN - 1) / FA_BN;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 5) tmem_alloc(tmem_slot, FA_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, FA_QBYTES);
      tma_load_4d(sQ, &tmQ, q_full, 0, h, q0, b);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % FA_STAGES;
        const uint32_t ph = (j / FA_STAGES) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_expect_tx(&kv_full[s], 2 * FA_KBYTES);
        tma_load_4d(sK + s * FA_KBYTES, &tmK, &kv_full[s], 0, h, j * FA_BN, kb);
        tma_load_4d(sV + s * FA_KBYTES, &tmV, &kv_full[s], 0, h, j * FA_BN, kb);
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(FA_BM, FA_BN, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(FA_BM, FA_D, 0, 1);   // P (K-major) x V (MN-major)
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
      auto issue_s = [&](int j) {
        const int s = j % FA_STAGES;
        mbar_wait(&kv_full[s], (j / FA_STAGES) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < FA_D / 16; ++kk)
          umma_bf16(tmem_base + FA_S_COL, make_smem_desc(aQ + kk * 32, 16, 1024),
                    make_smem_desc(aK + s * FA_KBYTES + kk * 32, 16, 1024), idesc_s, kk != 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j % FA_STAGES;
        mbar_wait(p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < FA_BN / 16; ++kk)
          umma_bf16(tmem_base + FA_O_COL, make_smem_desc(aP + (kk >> 2) * (FA_BM * 128) + (kk & 3) * 32, 16, 1024),
                    make_smem_desc(aV + s * FA_KBYTES + kk * 2048, 8192, 1024), idesc_o, kk != 0 ? 1u : 0u);
        umma_commit(&kv_empty[s]);
        umma_commit(o_full);
        if (j + 1 < ntiles) issue_s(j + 1);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroup
    const int r = warp * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    float o_acc[FA_D];
#pragma unroll
    for (int d = 0; d < FA_D; ++d) o_acc[d] = 0.f;
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
    uint8_t* prow = sP + r * 128;
    const int sw = r & 7;
    for (int j = 0; j < ntiles; ++j) {
      const int kbase = j * FA_BN;
      const bool tail = kbase + FA_BN > Nk;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // sweep 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < FA_BN / 32; ++c) {
        float v[32];
        tmem_ld32(t_lane + FA_S_COL + c * 32, v);
        tmem_ld_wait();
        if (tail) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (kbase + c * 32 + e >= Nk) v[e] = -INFINITY;
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) mx = fmaxf(mx, v[e]);
      }
      const float m_new = fmaxf(m, mx * scale_log2);
      const float alpha = fast_exp2(m - m_new);  // first tile: exp2(-inf) = 0
      // fold the previous tile's P V into the register accumulator
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < FA_D / 32; ++c) {
          float v[32];
          tmem_ld32(t_lane + FA_O_COL + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) o_acc[c * 32 + e] = fmaf(o_acc[c * 32 + e], alpha_prev, v[e]);
        }
      }
      // sweep 2: p = exp2(s*scale - m_new), row sum, bf16 pack, swizzled store
      float lsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < FA_BN / 32; ++c) {
        float v[32];
        tmem_ld32(t_lane + FA_S_COL + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float p = fast_exp2(fmaf(v[e], scale_log2, -m_new));
          if (tail && kbase + c * 32 + e >= Nk) p = 0.f;
          v[e] = p;
          lsum += p;
        }
        // columns [c*32, c*32+32) -> k-block (c>>1), 16-byte chunks (c&1)*4 .. +3
        uint8_t* pblk = prow + (c >> 1) * (FA_BM * 128);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]); u.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
          u.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]); u.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
          const int chunk = (c & 1) * 4 + g;
          *reinterpret_cast<uint4*>(pblk + ((chunk ^ sw) << 4)) = u;
        }
      }
      l = l * alpha + lsum;
      m = m_new;
      alpha_prev = alpha;
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // last tile's P V
    mbar_wait(o_full, (ntiles - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < FA_D / 32; ++c) {
      float v[32];
      tmem_ld32(t_lane + FA_O_COL + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) o_acc[c * 32 + e] = fmaf(o_acc[c * 32 + e], alpha_prev, v[e]);
    }
    const int row = q0 + r;
    if (row < Nq) {
      const float inv = 1.f / l;
      __nv_bfloat16* orow = out + (((int64_t)b * Nq + row) * H + h) * FA_D;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 u;
        u.x = pack_bf16(o_acc[g * 8 + 0] * inv, o_acc[g * 8 + 1] * inv);
        u.y = pack_bf16(o_acc[g * 8 + 2] * inv, o_acc[g * 8 + 3] * inv);
        u.z = pack_bf16(o_acc[g * 8 + 4] * inv, o_acc[g * 8 + 5] * inv);
        u.w = pack_bf16(o_acc[g * 8 + 6] * inv, o_acc[g * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(orow + g * 8) = u;
      }
      lse[((int64_t)b * H + h) * Nq + row] = (m + log2f(l)) * 0.6931471805599453f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, FA_TMEM_COLS);
}

// token-major [B, N, H, 64] bf16 -> 4-D tensor map {64, H, N, B}, box {64, 1, rows, 1}
static int make_qkv_tmap(CUtensorMap* tm, const void* base, int B, int N, int H, int box_rows) {
  const uint64_t dims[4] = {64, (uint64_t)H, (uint64_t)N, (uint64_t)B};
  const uint64_t str[3] = {64 * 2, (uint64_t)H * 64 * 2, (uint64_t)N * H * 64 * 2};
  const uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
  return make_tmap_bf16(tm, base, 4, dims, str, box);
}

int attn_fwd_tc(const void* q, const void* k, const void* v, void* out, float* lse, int B, int Nq, int Nk, int H,
                int kv_shift, float scale, cudaStream_t stream) {
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_qkv_tmap(&tq, q, B, Nq, H, FA_BM))) return rc;
  if ((rc = make_qkv_tmap(&tk, k, B, Nk, H, FA_BN))) return rc;
  if ((rc = make_qkv_tmap(&tv, v, B, Nk, H, FA_BN))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM);
    LGB_REQUIRE(e == cudaSuccess, kErrCuda, "attn_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid((Nq + FA_BM - 1) / FA_BM, H, B);
  attn_fwd_tc_kernel<<<grid, 192, FA_SMEM, stream>>>(tq, tk, tv, static_cast<__nv_bfloat16*>(out), lse, B, Nq, Nk, H,
                                                     kv_shift, scale * 1.4426950408889634f);
  return check_launch("attn_fwd_tc");
}

// Backward on tensor cores is the next milestone; until it lands the bf16 path uses the CUDA-core
// kernels (same math, bf16 I/O, fp32 accumulation).
int attn_bwd_tc(const void* q, const void* k, const void* v, const void* out, const float* lse, const void* dout,
                void* dq, void* dk, void* dv, float* delta, int B, int Nq, int Nk, int H, int kv_shift, float scale,
                cudaStream_t stream) {
  return attn_bwd_simt<__nv_bfloat16>(q, k, v, out, lse, dout, dq, dk, dv, delta, B, Nq, Nk, H, kv_shift, scale, stream);
}

}  // namespace lgb
