// Assignment head fused with its similarity GEMM (SURVEY K8+K9; lightglue.py:256-290, losses.py:6-25):
// sim = mdesc0 . mdesc1^T is produced tile by tile in tensor memory and consumed there -- it is never written to HBM
// (round 1 wrote 16.8 MB fp32 per pair per supervised layer, re-read it three times and kept it for backward).
//
// Three passes over the same skeleton (persistent CTAs, one per SM):
//   work item = (direction, pair b, 128-row panel).  direction 0: panel rows = keypoints of image 0, columns = all
//   keypoints of image 1 (thread == row i == TMEM lane: everything per ROW of sim); direction 1: the transposed
//   problem (thread == column j of sim: everything per COLUMN).  Running both directions doubles the (cheap, tensor-
//   core) contraction but makes every reduction a per-thread running value: no column partials, no merge kernels,
//   no atomics, deterministic.
//   The panel (A, 128 x D bf16) stays resident in shared memory for the item; 128-column tiles of the other image
//   (B, 128 x D) stream through a 2-stage TMA ring; S = A B^T (M128 N128, D/16 MMAs) lands in one of two TMEM buffers
//   so the MMA warp runs a tile ahead of the 16 epilogue warps (four per TMEM lane quarter, 32 columns each).
//
//   LSE  : running (max, sum exp2) per thread -> lse_row [B,M] (dir 0), lse_col [B,N] (dir 1).
//   STATS: scores in the reference's association order ((s - lse_r) + (s - lse_c)) + (lsig0 + lsig1); running
//          max / first argmax per thread -> rowmax/rowarg (dir 0), colmax/colarg (dir 1); dir 0 also gathers
//          sum_j gt_ij (2 s_ij - lse_r_i - lse_c_j) from the boolean assignment.  No MUFU work at all.
//   BWD  : dsim = gc (2 gt - softmax_row rowcnt - softmax_col colcnt) recomputed from S and the two LSE vectors,
//          rounded to bf16 and written IN PLACE over the S columns it came from (tcgen05.st), then
//          d(mdesc) panel (128 x D fp32, TMEM) += dsim (TMEM A operand) . B tile (the same smem bytes, MN-major):
//          dir 0 yields d(mdesc0) = dsim mdesc1, dir 1 yields d(mdesc1) = dsim^T mdesc0.  dsim never exists in HBM.
// All three are exact restatements of assign_lse / assign_scores / assign_bwd (assign.cu) on the fp32 accumulator.
#include "common.cuh"
#include "host_util.h"
#include "lgb200.h"

namespace lgb {

constexpr int AT_R = 128;                  // panel rows
constexpr int AT_C = 128;                  // columns per tile
constexpr int AT_KBLK = 64;                // K elements per smem block (one 128-byte swizzle row)
constexpr int AT_BLK = AT_R * AT_KBLK * 2; // 16 KiB
constexpr int AT_MAXKB = 4;                // D <= 256
constexpr int AT_STAGES = 2;
constexpr int AT_EWARPS = 16;
constexpr int AT_THREADS = (AT_EWARPS + 2) * 32;
constexpr int AT_SCR = AT_EWARPS * 2 * 32 * 4 + 3 * 384 * 4;  // per-warp column vectors + quarter-merge scratch
constexpr int AT_SMEM = (1 + AT_STAGES) * AT_MAXKB * AT_BLK + AT_SCR + 256;
constexpr int AT_ACC_COL = 256;            // TMEM: S0 [0,128) S1 [128,256) | d(mdesc) accumulator [256, 256 + D)

enum { AT_LSE = 0, AT_STATS = 1, AT_BWD = 2 };

struct AssignTcArgs {
  float alpha;                         // sim = alpha * (md0 . md1^T)
  int B, M, N, D;
  float* lse_row; float* lse_col;      // [B,M], [B,N]   (LSE: out; STATS / BWD: in)
  const float* ls0; const float* ls1;  // STATS: log sigmoid(z0) [B,M], log sigmoid(z1) [B,N]
  const uint8_t* gt;                   // [B,M,N] bool (STATS: optional; BWD: required)
  const uint8_t* gt_t;                 // [B,N,M] bool, the transposed copy (BWD, dir 1)
  float* rowmax; int* rowarg; float* colmax; int* colarg; float* pos_row_sum;  // STATS out
  const float* gcoef;                  // BWD: [B]
  const float* rowcnt; const float* colcnt;  // BWD: [B,M], [B,N]
  __nv_bfloat16* dmd0; __nv_bfloat16* dmd1;  // BWD out: [B*M, D], [B*N, D]
};

__device__ __forceinline__ void bar_epi() { asm volatile("bar.sync 1, %0;" ::"n"(AT_EWARPS * 32) : "memory"); }

template <int MODE>
__global__ void __launch_bounds__(AT_THREADS, 1)
    assign_tc_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1, AssignTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + AT_MAXKB * AT_BLK;
  float* sVec = reinterpret_cast<float*>(smem + (1 + AT_STAGES) * AT_MAXKB * AT_BLK);  // [16 warps][2][32]
  float* sMerge = sVec + AT_EWARPS * 2 * 32;                                              // [3][3][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (1 + AT_STAGES) * AT_MAXKB * AT_BLK + AT_SCR);
  uint64_t* a_full = bars;          // 1
  uint64_t* a_empty = bars + 1;     // 1
  uint64_t* b_full = bars + 2;      // [2]
  uint64_t* b_empty = bars + 4;     // [2]
  uint64_t* s_full = bars + 6;      // [2]
  uint64_t* s_done = bars + 8;      // [2]  epilogue finished with S buffer (LSE/STATS: free; BWD: dsim written)
  uint64_t* acc_full = bars + 10;   // BWD
  uint64_t* acc_empty = bars + 11;  // BWD
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = a.B, M = a.M, N = a.N, D = a.D;
  const int KB = D / AT_KBLK;
  const int PM = (M + AT_R - 1) / AT_R, PN = (N + AT_R - 1) / AT_R;
  const int nitems = B * (PM + PN);

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_done[s], AT_EWARPS);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, AT_EWARPS);
    mbar_fence_init();
  }
  if (warp == AT_EWARPS && lane == 0) {
    tma_prefetch_desc(&tm0);
    tma_prefetch_desc(&tm1);
  }
  if (warp == AT_EWARPS + 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (dir, b, panel); items of one pair are consecutive so its tiles are shared through L2
  auto decode = [&](int item, int& dir, int& b, int& p) {
    b = item / (PM + PN);
    const int r = item - b * (PM + PN);
    dir = r >= PM;
    p = dir ? r - PM : r;
  };

  if (warp == AT_EWARPS) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0, g = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
        int dir, b, p;
        decode(item, dir, b, p);
        const CUtensorMap* tA = dir ? &tm1 : &tm0;
        const CUtensorMap* tB = dir ? &tm0 : &tm1;
        const int ncols = dir ? M : N;
        const int ntiles = (ncols + AT_C - 1) / AT_C;
        mbar_wait(a_empty, (it & 1) ^ 1);
        mbar_expect_tx(a_full, KB * AT_BLK);
        for (int kb = 0; kb < KB; ++kb) tma_load_3d(sA + kb * AT_BLK, tA, a_full, kb * AT_KBLK, p * AT_R, b);
        for (int t = 0; t < ntiles; ++t, ++g) {
          const int s = g & 1;
          mbar_wait(&b_empty[s], ((g >> 1) & 1) ^ 1);
          mbar_expect_tx(&b_full[s], KB * AT_BLK);
          for (int kb = 0; kb < KB; ++kb)
            tma_load_3d(sB + (s * AT_MAXKB + kb) * AT_BLK, tB, &b_full[s], kb * AT_KBLK, t * AT_C, b);
        }
      }
    }
  } else if (warp == AT_EWARPS + 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-convergent, elected lane)
    constexpr uint32_t idesc_s = make_idesc_bf16(AT_R, AT_C, 0, 0);
    const uint32_t idesc_acc = make_idesc_bf16(AT_R, D, 0, 1);  // dsim (TMEM) x B tile MN-major, N = D
    const uint64_t dA = make_smem_desc(smem_u32(sA), 16, 1024);
    const uint64_t dBk = make_smem_desc(smem_u32(sB), 16, 1024);
    const uint64_t dBm = make_smem_desc(smem_u32(sB), AT_BLK, 1024);  // 64-wide channel chunks AT_BLK apart
    const bool leader = elect_one();
    int it = 0, g = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
      int dir, b, p;
      decode(item, dir, b, p);
      const int ncols = dir ? M : N;
      const int ntiles = (ncols + AT_C - 1) / AT_C;
      mbar_wait(a_full, it & 1);
      tc_fence_after();
      auto issue_acc = [&](int gg, bool first) {  // d(mdesc) += dsim(gg) . B(gg)
        const int s = gg & 1;
        mbar_wait(&s_done[s], (gg >> 1) & 1);
        tc_fence_after();
        if (first) {
          mbar_wait(acc_empty, (it & 1) ^ 1);
          tc_fence_after();
        }
        if (leader) {
          const uint64_t so = (uint64_t)((s * AT_MAXKB * AT_BLK) >> 4);
          const uint32_t abase = tmem_base + s * AT_C;
#pragma unroll
          for (int kk = 0; kk < AT_C / 16; ++kk)  // quarter cq of the tile wrote its packed dsim at column cq*32
            umma_bf16_ts(tmem_base + AT_ACC_COL, abase + (kk >> 1) * 32 + (kk & 1) * 8, dBm + so + (uint64_t)(kk * 128),
                         idesc_acc, (first && kk == 0) ? 0u : 1u);
          umma_commit(&b_empty[s]);
        }
        __syncwarp();
      };
      for (int t = 0; t < ntiles; ++t, ++g) {
        const int s = g & 1;
        mbar_wait(&b_full[s], (g >> 1) & 1);
        if (MODE != AT_BWD) mbar_wait(&s_done[s], ((g >> 1) & 1) ^ 1);  // epilogue drained this S buffer (2 tiles ago)
        tc_fence_after();
        if (leader) {
          const uint64_t so = (uint64_t)((s * AT_MAXKB * AT_BLK) >> 4);
          for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
            for (int kk = 0; kk < AT_KBLK / 16; ++kk)
              umma_bf16(tmem_base + s * AT_C, dA + (uint64_t)((kb * AT_BLK) >> 4) + (uint64_t)(kk * 2),
                        dBk + so + (uint64_t)((kb * AT_BLK) >> 4) + (uint64_t)(kk * 2), idesc_s, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&s_full[s]);
          if (MODE != AT_BWD) umma_commit(&b_empty[s]);
          if (t == ntiles - 1) umma_commit(a_empty);
        }
        __syncwarp();
        if (MODE == AT_BWD && t >= 1) issue_acc(g - 1, t == 1);
      }
      if (MODE == AT_BWD) {
        issue_acc(g - 1, ntiles == 1);
        if (leader) umma_commit(acc_full);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // 16 warps = 4 per scheduler: TMEM lane quarter q (rows), column quarter cq (32 of the tile's 128 columns)
    const int q = warp & 3, cq = warp >> 2;
    const int r = q * 32 + lane;                   // row within the panel == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* myvec = sVec + warp * 64;               // [2][32]: this warp's slice of the two per-column vectors
    constexpr float kLog2e = 1.4426950408889634f;
    const float c2 = a.alpha * kLog2e;
    int it = 0, g = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
      int dir, b, p;
      decode(item, dir, b, p);
      const int nrows = dir ? N : M, ncols = dir ? M : N;
      const int ntiles = (ncols + AT_C - 1) / AT_C;
      const int row = p * AT_R + r;
      const bool rok = row < nrows;
      const int64_t ro = (int64_t)b * nrows + (rok ? row : 0);
      // per-row constants
      float lse_mine = 0.f, ls_mine = 0.f, coef_mine = 0.f, gc = 0.f;
      if (MODE != AT_LSE) lse_mine = (dir ? a.lse_col : a.lse_row)[ro];
      if (MODE == AT_STATS) ls_mine = (dir ? a.ls1 : a.ls0)[ro];
      if (MODE == AT_BWD) {
        gc = a.gcoef[b];
        coef_mine = rok ? gc * (dir ? a.colcnt : a.rowcnt)[ro] : 0.f;
        lse_mine *= kLog2e;
      }
      const float* cvec_lse = (dir ? a.lse_row : a.lse_col) + (int64_t)b * ncols;  // per-column vectors of the tile
      const float* cvec_b = MODE == AT_STATS ? (dir ? a.ls0 : a.ls1) + (int64_t)b * ncols
                            : MODE == AT_BWD ? (dir ? a.rowcnt : a.colcnt) + (int64_t)b * ncols : nullptr;
      const uint8_t* grow = nullptr;  // this thread's row of the boolean assignment (or of its transpose)
      if (MODE == AT_BWD) grow = (dir ? a.gt_t : a.gt) + ro * ncols;
      if (MODE == AT_STATS && !dir && a.gt) grow = a.gt + ro * ncols;
      const bool gvec = (ncols & 15) == 0;
      // running values
      float m_run = -INFINITY, l_run = 0.f;     // LSE (log2 domain)
      float best = -INFINITY, psum = 0.f;       // STATS
      int besti = 0x7fffffff;

      for (int t = 0; t < ntiles; ++t, ++g) {
        const int s = g & 1;
        const int j0 = t * AT_C + cq * 32;      // first column of this warp's quarter
        const bool tail = j0 + 32 > ncols;
        if (MODE != AT_LSE) {                   // stage this warp's 32 per-column values, coalesced
          __syncwarp();
          const int j = j0 + lane;
          float v0 = 0.f, v1 = 0.f;
          if (j < ncols) {
            v0 = cvec_lse[j];
            v1 = cvec_b[j];
          }
          if (MODE == AT_BWD) {
            v0 = j < ncols ? v0 * kLog2e : INFINITY;  // exp2(x - inf) = 0 for columns past the end
            v1 *= gc;
          } else {
            v0 = j < ncols ? v1 - v0 : -INFINITY;     // STATS: other side's (log sigmoid - lse); -inf masks the tail
          }
          myvec[lane] = v0;
          myvec[32 + lane] = v1;
          __syncwarp();
        }
        uint32_t gw[8];  // 32 mask bytes of this thread's row (registers: constant indices only)
#pragma unroll
        for (int e = 0; e < 8; ++e) gw[e] = 0u;
        if (grow && !tail && gvec) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint4 u = *reinterpret_cast<const uint4*>(grow + j0 + 16 * e);
            gw[4 * e] = u.x; gw[4 * e + 1] = u.y; gw[4 * e + 2] = u.z; gw[4 * e + 3] = u.w;
          }
        } else if (grow) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (j0 + 4 * e + u < ncols) gw[e] |= (uint32_t)grow[j0 + 4 * e + u] << (8 * u);
          }
        }
        mbar_wait(&s_full[s], (g >> 1) & 1);
        tc_fence_after();
        float sv[32];
        tmem_ld32(t_lane + s * AT_C + cq * 32, sv);
        tmem_ld_wait();
        if (MODE == AT_LSE) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_done[s]);  // S is in registers: the MMA warp may overwrite the buffer
          if (tail) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (j0 + e >= ncols) sv[e] = -INFINITY;
          }
          float mx = sv[0];
#pragma unroll
          for (int e = 1; e < 32; ++e) mx = fmaxf(mx, sv[e]);
          // alpha > 0: max commutes with the scale.  All-masked quarter (j0 >= ncols): mx = -inf, contributes nothing.
          const float m_new = fmaxf(m_run, mx * c2);
          if (m_new > -INFINITY) {
            float sum = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) sum += fast_exp2(fmaf(sv[e], c2, -m_new));
            l_run = l_run * fast_exp2(m_run - m_new) + sum;
            m_run = m_new;
          }
        } else if (MODE == AT_STATS) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_done[s]);
          // score_ij = (2 alpha S_ij + [lsig - lse]_other(j)) + [lsig - lse]_mine: the second bracket is constant along
          // this thread's sweep, so the running max / first argmax is taken over the first term only (an FMA and an
          // FMNMX per element) and the constant is added when the result is written.  (assign_scores in assign.cu keeps
          // the reference's association order for the fp32 parity path; the two agree to fp32 rounding.)
          const float a2 = 2.f * a.alpha;
          float cmax = -INFINITY;
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4) {
            const float4 cv = *reinterpret_cast<const float4*>(myvec + e4 * 4);
            sv[e4 * 4] = fmaf(sv[e4 * 4], a2, cv.x); sv[e4 * 4 + 1] = fmaf(sv[e4 * 4 + 1], a2, cv.y);
            sv[e4 * 4 + 2] = fmaf(sv[e4 * 4 + 2], a2, cv.z); sv[e4 * 4 + 3] = fmaf(sv[e4 * 4 + 3], a2, cv.w);
            cmax = fmaxf(fmaxf(cmax, fmaxf(sv[e4 * 4], sv[e4 * 4 + 1])), fmaxf(sv[e4 * 4 + 2], sv[e4 * 4 + 3]));
            if (gw[e4] != 0u) {  // a ground-truth positive in these four columns (rare): sum_j gt (2 s - lse_r - lse_c)
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if ((gw[e4] >> (8 * u)) & 0xffu) {
                  const int e = e4 * 4 + u;
                  // 2 alpha S = sv[e] - cv; the other side's lse comes from memory, mine from the register
                  const float two_x = sv[e] - myvec[e];
                  psum += two_x - lse_mine - cvec_lse[j0 + e];
                }
            }
          }
          if (cmax > best) {  // a new running maximum is rare (~ln(columns) times per row): find its first position
            best = cmax;
#pragma unroll
            for (int e = 31; e >= 0; --e)
              if (sv[e] == cmax) besti = j0 + e;
          }
        } else {  // AT_BWD
          const float two_gc = 2.f * gc;
          uint32_t pw[16];
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4) {
            const float4 lc = *reinterpret_cast<const float4*>(myvec + e4 * 4);
            const float4 cc = *reinterpret_cast<const float4*>(myvec + 32 + e4 * 4);
            const float lcv[4] = {lc.x, lc.y, lc.z, lc.w}, ccv[4] = {cc.x, cc.y, cc.z, cc.w};
            float d[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int e = e4 * 4 + u;
              float v = ((gw[e4] >> (8 * u)) & 0xffu) ? two_gc : 0.f;
              v = fmaf(-fast_exp2(fmaf(sv[e], c2, -lse_mine)), coef_mine, v);
              if (ccv[u] != 0.f) v = fmaf(-fast_exp2(fmaf(sv[e], c2, -lcv[u])), ccv[u], v);  // warp-uniform branch
              d[u] = (tail && j0 + e >= ncols) ? 0.f : v;
            }
            pw[e4 * 2] = pack_bf16(d[0], d[1]);
            pw[e4 * 2 + 1] = pack_bf16(d[2], d[3]);
          }
          // dsim (bf16, two columns per 32-bit word) over the first 16 columns of this warp's own 32 S columns
          tmem_st16(t_lane + s * AT_C + cq * 32, pw);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_done[s]);
        }
      }

      // ---- end of item: merge the four column quarters (LSE / STATS) or drain the accumulator (BWD)
      if (MODE == AT_LSE) {
        if (cq) { sMerge[(cq - 1) * 256 + r] = m_run; sMerge[(cq - 1) * 256 + 128 + r] = l_run; }
        bar_epi();
        if (cq == 0) {
          float mn = m_run;
#pragma unroll
          for (int k = 0; k < 3; ++k) mn = fmaxf(mn, sMerge[k * 256 + r]);
          float l = l_run * fast_exp2(m_run - mn);
#pragma unroll
          for (int k = 0; k < 3; ++k) l += sMerge[k * 256 + 128 + r] * fast_exp2(sMerge[k * 256 + r] - mn);
          if (rok) (dir ? a.lse_col : a.lse_row)[ro] = (mn + log2f(l)) * 0.6931471805599453f;
        }
        bar_epi();
      } else if (MODE == AT_STATS) {
        if (cq) {
          sMerge[(cq - 1) * 384 + r] = best;
          reinterpret_cast<int*>(sMerge)[(cq - 1) * 384 + 128 + r] = besti;
          sMerge[(cq - 1) * 384 + 256 + r] = psum;
        }
        bar_epi();
        if (cq == 0 && rok) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {  // ascending column quarters: on equal values the earlier (lower) index stays
            const float b2 = sMerge[k * 384 + r];
            if (b2 > best) { best = b2; besti = reinterpret_cast<int*>(sMerge)[k * 384 + 128 + r]; }
            psum += sMerge[k * 384 + 256 + r];
          }
          (dir ? a.colmax : a.rowmax)[ro] = best + (ls_mine - lse_mine);
          (dir ? a.colarg : a.rowarg)[ro] = besti;
          if (!dir && a.pos_row_sum) a.pos_row_sum[ro] = psum;
        }
        bar_epi();
      } else {
        mbar_wait(acc_full, it & 1);
        tc_fence_after();
        __nv_bfloat16* out = (dir ? a.dmd1 : a.dmd0) + ro * D;
        const int cpw = D / 4;  // channels per column quarter
#pragma unroll 1
        for (int c = 0; c < cpw; c += 16) {
          float v[16];
          tmem_ld16(t_lane + AT_ACC_COL + cq * cpw + c, v);
          tmem_ld_wait();
          if (rok) {
#pragma unroll
            for (int e8 = 0; e8 < 2; ++e8) {
              uint4 u;
              u.x = pack_bf16(v[e8 * 8], v[e8 * 8 + 1]); u.y = pack_bf16(v[e8 * 8 + 2], v[e8 * 8 + 3]);
              u.z = pack_bf16(v[e8 * 8 + 4], v[e8 * 8 + 5]); u.w = pack_bf16(v[e8 * 8 + 6], v[e8 * 8 + 7]);
              *reinterpret_cast<uint4*>(out + cq * cpw + c + e8 * 8) = u;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AT_EWARPS + 1) tmem_dealloc(tmem_base, 512);
}

static int make_md_map(CUtensorMap* tm, const void* base, int B, int rows, int D) {
  const uint64_t dims[3] = {(uint64_t)D, (uint64_t)rows, (uint64_t)B};
  const uint64_t str[2] = {(uint64_t)D * 2, (uint64_t)rows * D * 2};
  const uint32_t box[3] = {AT_KBLK, AT_R, 1};
  return make_tmap_bf16(tm, base, 3, dims, str, box);
}

template <int MODE>
static int launch_assign_tc(const void* md0, const void* md1, const AssignTcArgs& a, cudaStream_t stream) {
  LGB_REQUIRE(md0 && md1, kErrInvalid, "assign_fused: null mdesc pointer");
  LGB_REQUIRE(a.B > 0 && a.M > 0 && a.N > 0, kErrInvalid, "assign_fused: empty input B=%d M=%d N=%d", a.B, a.M, a.N);
  LGB_REQUIRE(a.D % 64 == 0 && a.D >= 64 && a.D <= 256, kErrUnsupported,
              "assign_fused: descriptor width %d not a multiple of 64 in [64, 256]", a.D);
  LGB_REQUIRE(a.alpha > 0.f, kErrInvalid, "assign_fused: alpha must be positive");
  CUtensorMap t0, t1;
  int rc;
  if ((rc = make_md_map(&t0, md0, a.B, a.M, a.D))) return rc;
  if ((rc = make_md_map(&t1, md1, a.B, a.N, a.D))) return rc;
  auto kern = assign_tc_kernel<MODE>;
  if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), AT_SMEM))) return rc;
  const int items = a.B * ((a.M + AT_R - 1) / AT_R + (a.N + AT_R - 1) / AT_R);
  const int sms = device_sm_count();
  kern<<<items < sms ? items : sms, AT_THREADS, AT_SMEM, stream>>>(t0, t1, a);
  return check_launch("assign_fused");
}

}  // namespace lgb

using namespace lgb;

extern "C" {

int lgb200_assign_fused_lse(const void* md0, const void* md1, float alpha, float* lse_row, float* lse_col, int B, int M,
                            int N, int D, cudaStream_t stream) {
  LGB_REQUIRE(lse_row && lse_col, kErrInvalid, "assign_fused_lse: null pointer");
  AssignTcArgs a = {};
  a.alpha = alpha; a.B = B; a.M = M; a.N = N; a.D = D;
  a.lse_row = lse_row; a.lse_col = lse_col;
  return launch_assign_tc<AT_LSE>(md0, md1, a, stream);
}

int lgb200_assign_fused_stats(const void* md0, const void* md1, float alpha, const float* lse_row, const float* lse_col,
                              const float* ls0, const float* ls1, const uint8_t* gt, float* rowmax, int* rowarg,
                              float* colmax, int* colarg, float* pos_row_sum, int B, int M, int N, int D,
                              cudaStream_t stream) {
  LGB_REQUIRE(lse_row && lse_col && ls0 && ls1 && rowmax && rowarg && colmax && colarg, kErrInvalid,
              "assign_fused_stats: null pointer");
  LGB_REQUIRE((gt == nullptr) == (pos_row_sum == nullptr), kErrInvalid,
              "assign_fused_stats: gt and pos_row_sum must both be set or both be null");
  AssignTcArgs a = {};
  a.alpha = alpha; a.B = B; a.M = M; a.N = N; a.D = D;
  a.lse_row = const_cast<float*>(lse_row); a.lse_col = const_cast<float*>(lse_col);
  a.ls0 = ls0; a.ls1 = ls1; a.gt = gt;
  a.rowmax = rowmax; a.rowarg = rowarg; a.colmax = colmax; a.colarg = colarg; a.pos_row_sum = pos_row_sum;
  return launch_assign_tc<AT_STATS>(md0, md1, a, stream);
}

int lgb200_assign_fused_bwd(const void* md0, const void* md1, float alpha, const float* lse_row, const float* lse_col,
                            const uint8_t* gt, const uint8_t* gt_t, const float* gcoef, const float* rowcnt,
                            const float* colcnt, void* dmd0, void* dmd1, int B, int M, int N, int D,
                            cudaStream_t stream) {
  LGB_REQUIRE(lse_row && lse_col && gt && gt_t && gcoef && rowcnt && colcnt && dmd0 && dmd1, kErrInvalid,
              "assign_fused_bwd: null pointer");
  AssignTcArgs a = {};
  a.alpha = alpha; a.B = B; a.M = M; a.N = N; a.D = D;
  a.lse_row = const_cast<float*>(lse_row); a.lse_col = const_cast<float*>(lse_col);
  a.gt = gt; a.gt_t = gt_t; a.gcoef = gcoef; a.rowcnt = rowcnt; a.colcnt = colcnt;
  a.dmd0 = static_cast<__nv_bfloat16*>(dmd0); a.dmd1 = static_cast<__nv_bfloat16*>(dmd1);
  return launch_assign_tc<AT_BWD>(md0, md1, a, stream);
}

}  // extern "C"
