// Assignment head kernels: sigmoid_log_double_softmax, deep-supervision NLL
// terms, row/col argmax and match filtering -- the HBM-bound part of the
// LightGlue matcher (reference: gluefactory/models/matchers/lightglue.py:256-309,
// gluefactory/models/utils/losses.py:6-73).
//
// Layout: sim [B,M,N] fp32 row-major.  A CTA of 256 threads (8 warps) owns a
// strip of 32 consecutive rows and walks it in 128-column chunks; lane l of a
// warp owns 4 consecutive columns (one 128-bit load), warp w owns rows
// w, w+8, w+16, w+24 of the strip.  Row statistics stay in registers and are
// reduced with warp shuffles; column statistics are reduced across the 8 warps
// through shared memory and written as per-strip partials that a tiny second
// kernel merges in strip order (deterministic, no atomics).
#include <math.h>

#include "common.cuh"
#include "lgb200.h"

namespace lgb {

constexpr int kStripRows = 32;
constexpr int kChunk = 128;
constexpr int kWarps = 8;
constexpr int kRowsPerWarp = kStripRows / kWarps;  // 4
constexpr float kNegInf = -INFINITY;

__device__ __forceinline__ float4 ld4(const float* p, bool vec, int valid) {
  // loads up to 4 floats starting at p; lanes beyond `valid` are -inf
  float4 r;
  if (vec && valid >= 4) {
    r = *reinterpret_cast<const float4*>(p);
  } else {
    r.x = valid > 0 ? p[0] : kNegInf;
    r.y = valid > 1 ? p[1] : kNegInf;
    r.z = valid > 2 ? p[2] : kNegInf;
    r.w = valid > 3 ? p[3] : kNegInf;
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// pass 1: row LSE (final) and per-strip column (max, sumexp) partials
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) assign_lse_kernel(const float* __restrict__ sim, float* __restrict__ lse_row,
                                                        float* __restrict__ colpart_m,
                                                        float* __restrict__ colpart_s, int M, int N, int nstrips) {
  __shared__ float s_cmax[2][kWarps][kChunk];
  __shared__ float s_csum[2][kWarps][kChunk];
  const int b = blockIdx.y, strip = blockIdx.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec = (N & 3) == 0;
  const float* base = sim + (int64_t)b * M * N;

  float rm[kRowsPerWarp], rs[kRowsPerWarp];
  int rows[kRowsPerWarp];
#pragma unroll
  for (int t = 0; t < kRowsPerWarp; ++t) {
    rm[t] = kNegInf;
    rs[t] = 0.f;
    rows[t] = strip * kStripRows + w + kWarps * t;
  }
  auto load_chunk = [&](int c0, float4* dst) {
    const int col = c0 + lane * 4;
    const int valid = min(4, N - col);  // may be <= 0
#pragma unroll
    for (int t = 0; t < kRowsPerWarp; ++t)
      dst[t] = (rows[t] < M && valid > 0) ? ld4(base + (int64_t)rows[t] * N + col, vec, valid)
                                           : make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
  };
  int buf = 0;
  float4 xn[kRowsPerWarp];
  load_chunk(0, xn);
  for (int c0 = 0; c0 < N; c0 += kChunk, buf ^= 1) {
    float4 x[kRowsPerWarp];
#pragma unroll
    for (int t = 0; t < kRowsPerWarp; ++t) x[t] = xn[t];
    if (c0 + kChunk < N) load_chunk(c0 + kChunk, xn);  // next chunk in flight while this one is reduced
    // ---- row statistics (online, rescale only when the running max moves)
#pragma unroll
    for (int t = 0; t < kRowsPerWarp; ++t) {
      float cm = fmaxf(fmaxf(x[t].x, x[t].y), fmaxf(x[t].z, x[t].w));
      if (cm > rm[t]) {
        rs[t] *= __expf(rm[t] - cm);  // rm=-inf -> exp(-inf)=0, rs is 0 anyway
        rm[t] = cm;
      }
      const float ref = (rm[t] == kNegInf) ? 0.f : rm[t];
      rs[t] += (__expf(x[t].x - ref) + __expf(x[t].y - ref)) + (__expf(x[t].z - ref) + __expf(x[t].w - ref));
    }
    // ---- column statistics of this strip: max first (no exp), then one exp per element
    float4 cm4;
    cm4.x = fmaxf(fmaxf(x[0].x, x[1].x), fmaxf(x[2].x, x[3].x));
    cm4.y = fmaxf(fmaxf(x[0].y, x[1].y), fmaxf(x[2].y, x[3].y));
    cm4.z = fmaxf(fmaxf(x[0].z, x[1].z), fmaxf(x[2].z, x[3].z));
    cm4.w = fmaxf(fmaxf(x[0].w, x[1].w), fmaxf(x[2].w, x[3].w));
    *reinterpret_cast<float4*>(&s_cmax[buf][w][lane * 4]) = cm4;
    __syncthreads();
    float4 mx = *reinterpret_cast<float4*>(&s_cmax[buf][0][lane * 4]);
#pragma unroll
    for (int ww = 1; ww < kWarps; ++ww) {
      float4 o = *reinterpret_cast<float4*>(&s_cmax[buf][ww][lane * 4]);
      mx.x = fmaxf(mx.x, o.x); mx.y = fmaxf(mx.y, o.y); mx.z = fmaxf(mx.z, o.z); mx.w = fmaxf(mx.w, o.w);
    }
    float4 rf;  // reference (0 when the whole strip column is -inf)
    rf.x = mx.x == kNegInf ? 0.f : mx.x; rf.y = mx.y == kNegInf ? 0.f : mx.y;
    rf.z = mx.z == kNegInf ? 0.f : mx.z; rf.w = mx.w == kNegInf ? 0.f : mx.w;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < kRowsPerWarp; ++t) {
      cs.x += __expf(x[t].x - rf.x); cs.y += __expf(x[t].y - rf.y);
      cs.z += __expf(x[t].z - rf.z); cs.w += __expf(x[t].w - rf.w);
    }
    *reinterpret_cast<float4*>(&s_csum[buf][w][lane * 4]) = cs;
    __syncthreads();
    if (threadIdx.x < kChunk && c0 + threadIdx.x < N) {
      float m = s_cmax[buf][0][threadIdx.x], s = s_csum[buf][0][threadIdx.x];
#pragma unroll
      for (int ww = 1; ww < kWarps; ++ww) {
        m = fmaxf(m, s_cmax[buf][ww][threadIdx.x]);
        s += s_csum[buf][ww][threadIdx.x];
      }
      const int64_t o = ((int64_t)b * nstrips + strip) * N + c0 + threadIdx.x;
      colpart_m[o] = m;
      colpart_s[o] = s;
    }
  }
#pragma unroll
  for (int t = 0; t < kRowsPerWarp; ++t) {
    float m = rm[t], s = rs[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      lse_merge(m, s, m2, s2);
    }
    if (lane == 0 && rows[t] < M) lse_row[(int64_t)b * M + rows[t]] = m + logf(s);
  }
}

// 32 columns x 8 strip-groups per CTA: max pass (no exp), then one exp per partial; groups merged through smem
__global__ void __launch_bounds__(256) assign_col_lse_merge_kernel(const float* __restrict__ colpart_m,
                                                                  const float* __restrict__ colpart_s,
                                                                  float* __restrict__ lse_col, int N, int nstrips) {
  __shared__ float s_a[8][33];
  const int b = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  const bool ok = j < N;
  const float* pm = colpart_m + (int64_t)b * nstrips * N + (ok ? j : 0);
  const float* ps = colpart_s + (int64_t)b * nstrips * N + (ok ? j : 0);
  float m = kNegInf;
  for (int st = ty; st < nstrips; st += 8) m = fmaxf(m, pm[(int64_t)st * N]);
  s_a[ty][tx] = m;
  __syncthreads();
  float gm = s_a[0][tx];
#pragma unroll
  for (int g = 1; g < 8; ++g) gm = fmaxf(gm, s_a[g][tx]);
  const float ref = (gm == kNegInf) ? 0.f : gm;
  float sum = 0.f;
  for (int st = ty; st < nstrips; st += 8) sum += ps[(int64_t)st * N] * __expf(pm[(int64_t)st * N] - ref);
  __syncthreads();
  s_a[ty][tx] = sum;
  __syncthreads();
  if (ty == 0 && ok) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += s_a[g][tx];  // fixed order: deterministic
    lse_col[(int64_t)b * N + j] = gm + logf(t);
  }
}

// ---------------------------------------------------------------------------------------------
// pass 2: scores (optional dense write), row/col max+argmax, positive-weighted sums
// ---------------------------------------------------------------------------------------------
struct ScoreArgs {
  const float* sim;
  const float* lse_row;
  const float* lse_col;
  const float* ls0;   // log sigmoid(z0)   [B,M]
  const float* ls1;   // log sigmoid(z1)   [B,N]
  const float* dust0; // log sigmoid(-z0)  [B,M]  (dustbin column)
  const float* dust1; // log sigmoid(-z1)  [B,N]  (dustbin row)
  const uint8_t* gt;  // [B,M,N] or null
  float* scores;      // [B,M+1,N+1] or null
  float* rowmax; int* rowarg;       // [B,M]
  float* colpart_v; int* colpart_i; // [B,nstrips,N]
  float* pos_row_sum;  // [B,M] or null : sum_j gt_ij * (2 sim_ij - lse_row_i - lse_col_j)
  float* row_expsum;   // [B,M] or null : sum_{j<=N} exp(scores_ij)
  int M, N, nstrips;
};

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__global__ void __launch_bounds__(256) assign_scores_kernel(ScoreArgs a) {
  __shared__ float s_v[2][kWarps][kChunk];
  __shared__ int s_i[2][kWarps][kChunk];
  const int b = blockIdx.y, strip = blockIdx.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = a.M, N = a.N;
  const bool vec = (N & 3) == 0;
  const float* base = a.sim + (int64_t)b * M * N;
  const uint8_t* gbase = a.gt ? a.gt + (int64_t)b * M * N : nullptr;

  int rows[kRowsPerWarp];
  float lr[kRowsPerWarp], l0[kRowsPerWarp], bestv[kRowsPerWarp], psum[kRowsPerWarp], esum[kRowsPerWarp];
  int besti[kRowsPerWarp];
#pragma unroll
  for (int t = 0; t < kRowsPerWarp; ++t) {
    rows[t] = strip * kStripRows + w + kWarps * t;
    const bool ok = rows[t] < M;
    lr[t] = ok ? a.lse_row[(int64_t)b * M + rows[t]] : 0.f;
    l0[t] = ok ? a.ls0[(int64_t)b * M + rows[t]] : 0.f;
    bestv[t] = kNegInf;
    besti[t] = 0x7fffffff;
    psum[t] = 0.f;
    esum[t] = 0.f;
  }
  auto load_chunk = [&](int c0, float4* dx, uint32_t* dg) {
    const int col = c0 + lane * 4;
    const int valid = min(4, N - col);
#pragma unroll
    for (int t = 0; t < kRowsPerWarp; ++t) {
      dx[t] = make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
      dg[t] = 0;
      if (rows[t] >= M || valid <= 0) continue;
      const int64_t roff = (int64_t)rows[t] * N + col;
      dx[t] = ld4(base + roff, vec, valid);
      if (gbase) {
        if (vec && valid >= 4) dg[t] = *reinterpret_cast<const uint32_t*>(gbase + roff);
        else
          for (int e = 0; e < valid; ++e) dg[t] |= (uint32_t)gbase[roff + e] << (8 * e);
      }
    }
  };
  int buf = 0;
  float4 xn[kRowsPerWarp];
  uint32_t gn[kRowsPerWarp];
  load_chunk(0, xn, gn);
  for (int c0 = 0; c0 < N; c0 += kChunk, buf ^= 1) {
    const int col = c0 + lane * 4;
    const int valid = min(4, N - col);
    float4 xc[kRowsPerWarp];
    uint32_t gc[kRowsPerWarp];
#pragma unroll
    for (int t = 0; t < kRowsPerWarp; ++t) { xc[t] = xn[t]; gc[t] = gn[t]; }
    if (c0 + kChunk < N) load_chunk(c0 + kChunk, xn, gn);  // next chunk in flight while this one is processed
    float lc[4], l1[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      lc[e] = e < valid ? a.lse_col[(int64_t)b * N + col + e] : 0.f;
      l1[e] = e < valid ? a.ls1[(int64_t)b * N + col + e] : 0.f;
    }
    float cbv[4];
    int cbi[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { cbv[e] = kNegInf; cbi[e] = 0x7fffffff; }
#pragma unroll
    for (int t = 0; t < kRowsPerWarp; ++t) {
      if (rows[t] >= M || valid <= 0) continue;
      const float x[4] = {xc[t].x, xc[t].y, xc[t].z, xc[t].w};
      const uint32_t g = gc[t];
      float sc[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // same association as the reference: (log_softmax_row + log_softmax_col) + (lsig0 + lsig1)
        sc[e] = ((x[e] - lr[t]) + (x[e] - lc[e])) + (l0[t] + l1[e]);
        if (e < valid) {
          if (sc[e] > bestv[t]) { bestv[t] = sc[e]; besti[t] = col + e; }
          if (better(sc[e], rows[t], cbv[e], cbi[e])) { cbv[e] = sc[e]; cbi[e] = rows[t]; }
          if ((g >> (8 * e)) & 0xffu) psum[t] += (x[e] - lr[t]) + (x[e] - lc[e]);
          if (a.row_expsum) esum[t] += __expf(sc[e]);
        }
      }
      if (a.scores) {
        float* o = a.scores + ((int64_t)b * (M + 1) + rows[t]) * (N + 1) + col;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (e < valid) o[e] = sc[e];
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { s_v[buf][w][lane * 4 + e] = cbv[e]; s_i[buf][w][lane * 4 + e] = cbi[e]; }
    __syncthreads();
    if (threadIdx.x < kChunk && c0 + threadIdx.x < N) {
      float v = s_v[buf][0][threadIdx.x];
      int i = s_i[buf][0][threadIdx.x];
#pragma unroll
      for (int ww = 1; ww < kWarps; ++ww) {
        float v2 = s_v[buf][ww][threadIdx.x];
        int i2 = s_i[buf][ww][threadIdx.x];
        if (better(v2, i2, v, i)) { v = v2; i = i2; }
      }
      const int64_t o = ((int64_t)b * a.nstrips + strip) * N + c0 + threadIdx.x;
      a.colpart_v[o] = v;
      a.colpart_i[o] = i;
    }
    // s_v/s_i[buf] are rewritten two chunks later, after the next __syncthreads
  }
#pragma unroll
  for (int t = 0; t < kRowsPerWarp; ++t) {
    float v = bestv[t];
    int i = besti[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float v2 = __shfl_xor_sync(0xffffffffu, v, o);
      int i2 = __shfl_xor_sync(0xffffffffu, i, o);
      if (better(v2, i2, v, i)) { v = v2; i = i2; }
    }
    float ps = warp_sum(psum[t]);
    float es = a.row_expsum ? warp_sum(esum[t]) : 0.f;
    if (lane == 0 && rows[t] < M) {
      const int64_t r = (int64_t)b * M + rows[t];
      a.rowmax[r] = v;
      a.rowarg[r] = i;
      if (a.pos_row_sum) a.pos_row_sum[r] = ps;
      const float d0 = a.dust0[r];
      if (a.row_expsum) a.row_expsum[r] = es + __expf(d0);
      if (a.scores) a.scores[((int64_t)b * (M + 1) + rows[t]) * (N + 1) + N] = d0;
    }
  }
}

__global__ void __launch_bounds__(256) assign_col_arg_merge_kernel(const float* __restrict__ colpart_v,
                                                                  const int* __restrict__ colpart_i,
                                                                  const float* __restrict__ dust1,
                                                                  float* __restrict__ colmax, int* __restrict__ colarg,
                                                                  float* __restrict__ scores, int M, int N,
                                                                  int nstrips) {
  __shared__ float s_v[8][33];
  __shared__ int s_i[8][33];
  const int b = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float v = kNegInf;
  int i = 0x7fffffff;
  if (j < N) {
    for (int st = ty; st < nstrips; st += 8) {
      const int64_t o = ((int64_t)b * nstrips + st) * N + j;
      const float v2 = colpart_v[o];
      const int i2 = colpart_i[o];
      if (better(v2, i2, v, i)) { v = v2; i = i2; }
    }
  }
  s_v[ty][tx] = v;
  s_i[ty][tx] = i;
  __syncthreads();
  if (ty != 0 || j > N) return;
  if (j == N) {
    if (scores) scores[((int64_t)b * (M + 1) + M) * (N + 1) + N] = 0.f;
    return;
  }
#pragma unroll
  for (int g = 1; g < 8; ++g)
    if (better(s_v[g][tx], s_i[g][tx], v, i)) { v = s_v[g][tx]; i = s_i[g][tx]; }
  colmax[(int64_t)b * N + j] = v;
  colarg[(int64_t)b * N + j] = i;
  if (scores) scores[((int64_t)b * (M + 1) + M) * (N + 1) + j] = dust1[(int64_t)b * N + j];
}

// ---------------------------------------------------------------------------------------------
// backward of S_pos(sim) = sum_ij gt_ij (2 sim_ij - lse_row_i - lse_col_j):
//   dsim_ij = 2 c_b gt_ij - exp(sim_ij - lse_row_i) a_row_i - exp(sim_ij - lse_col_j) a_col_j
// (here a_row_i = c_b * rowcount_i(gt), a_col_j = c_b * colcount_j(gt): the
//  kernel takes the counts and applies c_b itself)
// ---------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(256) assign_bwd_kernel(const float* __restrict__ sim,
                                                        const float* __restrict__ lse_row,
                                                        const float* __restrict__ lse_col,
                                                        const uint8_t* __restrict__ gt, const float* __restrict__ gcoef,
                                                        const float* __restrict__ a_row,
                                                        const float* __restrict__ a_col, OutT* __restrict__ dsim, int M,
                                                        int N) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const bool vec = (N & 3) == 0;
  const float gc = gcoef[b];
  const float c2 = 2.f * gc;
  const float lr = lse_row[(int64_t)b * M + row];
  const float ar = gc * a_row[(int64_t)b * M + row];
  const int64_t roff = ((int64_t)b * M + row) * N;
  for (int col = lane * 4; col < N; col += 128) {
    const int valid = min(4, N - col);
    float4 x4 = ld4(sim + roff + col, vec, valid);
    float x[4] = {x4.x, x4.y, x4.z, x4.w};
    uint32_t g = 0;
    if (vec && valid >= 4) g = *reinterpret_cast<const uint32_t*>(gt + roff + col);
    else
      for (int e = 0; e < valid; ++e) g |= (uint32_t)gt[roff + col + e] << (8 * e);
    float d[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      d[e] = 0.f;
      if (e < valid) {
        const float ac = gc * a_col[(int64_t)b * N + col + e];
        float v = ((g >> (8 * e)) & 0xffu) ? c2 : 0.f;
        if (ar != 0.f) v -= __expf(x[e] - lr) * ar;
        if (ac != 0.f) v -= __expf(x[e] - lse_col[(int64_t)b * N + col + e]) * ac;
        d[e] = v;
      }
    }
    OutT* o = dsim + roff + col;
    if constexpr (sizeof(OutT) == 4) {
      if (vec && valid >= 4) *reinterpret_cast<float4*>(o) = make_float4(d[0], d[1], d[2], d[3]);
      else
        for (int e = 0; e < valid; ++e) o[e] = d[e];
    } else {
      if (vec && valid >= 4) {
        uint2 p;
        p.x = pack_bf16(d[0], d[1]);
        p.y = pack_bf16(d[2], d[3]);
        *reinterpret_cast<uint2*>(o) = p;
      } else {
        for (int e = 0; e < valid; ++e) o[e] = __float2bfloat16(d[e]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// filter_matches (lightglue.py:293-309) from the row/col argmax of pass 2
// ---------------------------------------------------------------------------------------------
__global__ void filter_matches_kernel(const float* __restrict__ rowmax, const int* __restrict__ rowarg,
                                      const int* __restrict__ colarg, float th, int64_t* __restrict__ m0,
                                      int64_t* __restrict__ m1, float* __restrict__ ms0, float* __restrict__ ms1,
                                      int M, int N) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int* ra = rowarg + (int64_t)b * M;
  const int* ca = colarg + (int64_t)b * N;
  const float* rm = rowmax + (int64_t)b * M;
  if (t < M) {
    const int j = ra[t];
    const bool mutual = (j >= 0 && j < N) && ca[j] == t;
    const float s = mutual ? expf(rm[t]) : 0.f;
    const bool valid = mutual && s > th;
    m0[(int64_t)b * M + t] = valid ? j : -1;
    ms0[(int64_t)b * M + t] = s;
  }
  if (t < N) {
    const int i = ca[t];
    const bool mutual1 = (i >= 0 && i < M) && ra[i] == t;
    float s = 0.f;
    bool valid = false;
    if (mutual1) {
      // mutual1 implies mutual0 of row i (ca[ra[i]] == i)
      s = expf(rm[i]);
      valid = s > th;
    }
    m1[(int64_t)b * N + t] = valid ? i : -1;
    ms1[(int64_t)b * N + t] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// O(M+N) terms of one supervised layer, fused: matchability log-sigmoids, the NLL of losses.py:6-25
// (given the similarity part sum_j gt_ij(2 sim - lse_r - lse_c) per row from pass 2), and the
// TokenConfidence BCE of lightglue.py:81-94.  zt holds [matchability logit, token logit] per token,
// tokens ordered [image0 (B*M); image1 (B*N)].  One CTA per pair; deterministic block reduction.
// ---------------------------------------------------------------------------------------------
struct HeadArgs {
  const float* zt;           // [T, 2]
  const float* pos_row_sum;  // [B,M]
  const float* rowcnt; const float* colcnt; const float* neg0; const float* neg1;
  const float* rowmax; const int* rowarg; const float* colmax; const int* colarg;
  const int* fin0; const int* fin1;  // final-layer argmax incl. dustbin, or null (no confidence term)
  const float* num_pos; const float* num_neg;  // [B]
  float bal;
  int B, M, N;
  float* nll; float* nll_pos; float* nll_neg; float* conf;  // [B]
};

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

// grid (chunks of 256 tokens over [side-0 rows | side-1 columns], B); per-chunk partial sums, then the last
// CTA of each pair to finish adds the partials in chunk order (deterministic) and writes the four outputs.
__global__ void __launch_bounds__(256) head_terms_fwd_kernel(HeadArgs a, float* __restrict__ part,
                                                            unsigned* __restrict__ counter) {
  __shared__ float red[8];
  __shared__ bool s_last;
  const int b = blockIdx.y, M = a.M, N = a.N;
  const int64_t t0 = (int64_t)a.B * M;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  float pos = 0.f, neg = 0.f, c0 = 0.f, c1 = 0.f;
  if (idx < M) {
    const int64_t r = (int64_t)b * M + idx;
    const float2 z = *reinterpret_cast<const float2*>(a.zt + 2 * r);
    const float ls = log_sigmoid(z.x), du = ls - z.x;
    pos = a.pos_row_sum[r] + a.rowcnt[r] * ls;
    neg = a.neg0[r] * du;
    if (a.fin0) {
      const int arg = du > a.rowmax[r] ? N : a.rowarg[r];
      const float y = (a.fin0[r] == arg) ? 1.f : 0.f;
      c0 = fmaxf(z.y, 0.f) - z.y * y + log1pf(expf(-fabsf(z.y)));
    }
  } else if (idx < M + N) {
    const int64_t r = (int64_t)b * N + (idx - M);
    const float2 z = *reinterpret_cast<const float2*>(a.zt + 2 * (t0 + r));
    const float ls = log_sigmoid(z.x), du = ls - z.x;
    pos = a.colcnt[r] * ls;
    neg = a.neg1[r] * du;
    if (a.fin1) {
      const int arg = du > a.colmax[r] ? M : a.colarg[r];
      const float y = (a.fin1[r] == arg) ? 1.f : 0.f;
      c1 = fmaxf(z.y, 0.f) - z.y * y + log1pf(expf(-fabsf(z.y)));
    }
  }
  pos = block_sum_256(pos, red);
  neg = block_sum_256(neg, red);
  c0 = block_sum_256(c0, red);
  c1 = block_sum_256(c1, red);
  const int nchunk = gridDim.x;
  if (threadIdx.x == 0) {
    float4* dst = reinterpret_cast<float4*>(part) + (int64_t)b * nchunk + blockIdx.x;
    *dst = make_float4(pos, neg, c0, c1);
    __threadfence();
    const unsigned prev = atomicAdd(&counter[b], 1u);
    s_last = (prev == (unsigned)nchunk - 1);
    if (s_last) counter[b] = 0u;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    float sp = 0.f, sn = 0.f, s0 = 0.f, s1 = 0.f;
    for (int c = 0; c < nchunk; ++c) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(part) + (int64_t)b * nchunk + c);
      sp += v.x; sn += v.y; s0 += v.z; s1 += v.w;
    }
    const float np = -sp / a.num_pos[b], nn = -sn / a.num_neg[b];
    a.nll_pos[b] = np;
    a.nll_neg[b] = nn;
    a.nll[b] = a.bal * np + (1.f - a.bal) * nn;
    a.conf[b] = a.fin0 ? 0.5f * (s0 / M + s1 / N) : 0.f;
  }
}

// d(zt) for upstream gradients g_nll[b], g_conf[b]
__global__ void __launch_bounds__(256) head_terms_bwd_kernel(HeadArgs a, const float* __restrict__ g_nll,
                                                            const float* __restrict__ g_conf,
                                                            float* __restrict__ dzt) {
  const int M = a.M, N = a.N;
  const int64_t t0 = (int64_t)a.B * M, T = t0 + (int64_t)a.B * N;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const bool side1 = t >= t0;
  const int64_t r = side1 ? t - t0 : t;
  const int L = side1 ? N : M;
  const int b = (int)(r / L);
  const float2 z = *reinterpret_cast<const float2*>(a.zt + 2 * t);
  const float g = g_nll[b];
  const float gpos = -a.bal * g / a.num_pos[b];
  const float gneg = -(1.f - a.bal) * g / a.num_neg[b];
  const float sg = 1.f / (1.f + expf(-z.x));
  const float cnt = side1 ? a.colcnt[r] : a.rowcnt[r];
  const float ng = side1 ? a.neg1[r] : a.neg0[r];
  float2 d;
  d.x = gpos * cnt * (1.f - sg) - gneg * ng * sg;
  d.y = 0.f;
  const int* fin = side1 ? a.fin1 : a.fin0;
  if (fin) {
    const float ls = log_sigmoid(z.x), du = ls - z.x;
    const float mx = side1 ? a.colmax[r] : a.rowmax[r];
    const int ag = side1 ? a.colarg[r] : a.rowarg[r];
    const int arg = du > mx ? (side1 ? M : N) : ag;
    const float y = (fin[r] == arg) ? 1.f : 0.f;
    d.y = g_conf[b] * (0.5f / L) * (1.f / (1.f + expf(-z.y)) - y);
  }
  *reinterpret_cast<float2*>(dzt + 2 * t) = d;
}

// ls = log sigmoid(z), du = log sigmoid(-z) from the strided logits (feeds pass 2)
__global__ void __launch_bounds__(256) logsig_kernel(const float* __restrict__ zt, float* __restrict__ ls,
                                                    float* __restrict__ du, int64_t T) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float z = zt[2 * t];
  const float l = log_sigmoid(z);
  ls[t] = l;
  du[t] = l - z;
}

}  // namespace lgb

using namespace lgb;

extern "C" {

int lgb200_head_logsig(const float* zt, float* ls, float* du, int64_t T, cudaStream_t stream) {
  LGB_REQUIRE(zt && ls && du && T > 0, kErrInvalid, "head_logsig: bad arguments");
  logsig_kernel<<<(unsigned)((T + 255) / 256), 256, 0, stream>>>(zt, ls, du, T);
  return check_launch("head_logsig");
}

static int fill_head_args(HeadArgs& a, const float* zt, const float* pos_row_sum, const float* rowcnt,
                          const float* colcnt, const float* neg0, const float* neg1, const float* rowmax,
                          const int* rowarg, const float* colmax, const int* colarg, const int* fin0, const int* fin1,
                          const float* num_pos, const float* num_neg, float bal, int B, int M, int N) {
  LGB_REQUIRE(zt && pos_row_sum && rowcnt && colcnt && neg0 && neg1 && rowmax && rowarg && colmax && colarg &&
                  num_pos && num_neg,
              kErrInvalid, "head_terms: null pointer");
  LGB_REQUIRE((fin0 == nullptr) == (fin1 == nullptr), kErrInvalid, "head_terms: fin0/fin1 must both be set or null");
  LGB_REQUIRE(B > 0 && M > 0 && N > 0, kErrInvalid, "head_terms: empty input");
  a.zt = zt; a.pos_row_sum = pos_row_sum; a.rowcnt = rowcnt; a.colcnt = colcnt; a.neg0 = neg0; a.neg1 = neg1;
  a.rowmax = rowmax; a.rowarg = rowarg; a.colmax = colmax; a.colarg = colarg; a.fin0 = fin0; a.fin1 = fin1;
  a.num_pos = num_pos; a.num_neg = num_neg; a.bal = bal; a.B = B; a.M = M; a.N = N;
  a.nll = a.nll_pos = a.nll_neg = a.conf = nullptr;
  return 0;
}

int lgb200_head_terms_fwd(const float* zt, const float* pos_row_sum, const float* rowcnt, const float* colcnt,
                          const float* neg0, const float* neg1, const float* rowmax, const int* rowarg,
                          const float* colmax, const int* colarg, const int* fin0, const int* fin1,
                          const float* num_pos, const float* num_neg, float bal, float* nll, float* nll_pos,
                          float* nll_neg, float* conf, float* ws, unsigned* counters, int B, int M, int N,
                          cudaStream_t stream) {
  HeadArgs a;
  int rc = fill_head_args(a, zt, pos_row_sum, rowcnt, colcnt, neg0, neg1, rowmax, rowarg, colmax, colarg, fin0, fin1,
                          num_pos, num_neg, bal, B, M, N);
  if (rc) return rc;
  LGB_REQUIRE(nll && nll_pos && nll_neg && conf && ws && counters, kErrInvalid, "head_terms_fwd: null output");
  a.nll = nll; a.nll_pos = nll_pos; a.nll_neg = nll_neg; a.conf = conf;
  head_terms_fwd_kernel<<<dim3((M + N + 255) / 256, B), 256, 0, stream>>>(a, ws, counters);
  return check_launch("head_terms_fwd");
}

int lgb200_head_terms_bwd(const float* zt, const float* rowcnt, const float* colcnt, const float* neg0,
                          const float* neg1, const float* rowmax, const int* rowarg, const float* colmax,
                          const int* colarg, const int* fin0, const int* fin1, const float* num_pos,
                          const float* num_neg, float bal, const float* g_nll, const float* g_conf, float* dzt, int B,
                          int M, int N, cudaStream_t stream) {
  HeadArgs a;
  int rc = fill_head_args(a, zt, rowcnt /*unused*/, rowcnt, colcnt, neg0, neg1, rowmax, rowarg, colmax, colarg, fin0,
                          fin1, num_pos, num_neg, bal, B, M, N);
  if (rc) return rc;
  LGB_REQUIRE(g_nll && g_conf && dzt, kErrInvalid, "head_terms_bwd: null pointer");
  const int64_t T = (int64_t)B * (M + N);
  head_terms_bwd_kernel<<<(unsigned)((T + 255) / 256), 256, 0, stream>>>(a, g_nll, g_conf, dzt);
  return check_launch("head_terms_bwd");
}

size_t lgb200_assign_ws_bytes(int B, int M, int N) {
  const int nstrips = (M + kStripRows - 1) / kStripRows;
  return (size_t)B * nstrips * N * 8;
}

int lgb200_assign_lse(const float* sim, float* lse_row, float* lse_col, void* ws, int B, int M, int N,
                      cudaStream_t stream) {
  LGB_REQUIRE(sim && lse_row && lse_col && ws, kErrInvalid, "assign_lse: null pointer");
  LGB_REQUIRE(B > 0 && M > 0 && N > 0, kErrInvalid, "assign_lse: empty input B=%d M=%d N=%d", B, M, N);
  const int nstrips = (M + kStripRows - 1) / kStripRows;
  float* pm = static_cast<float*>(ws);
  float* ps = pm + (size_t)B * nstrips * N;
  assign_lse_kernel<<<dim3(nstrips, B), 256, 0, stream>>>(sim, lse_row, pm, ps, M, N, nstrips);
  assign_col_lse_merge_kernel<<<dim3((N + 31) / 32, B), 256, 0, stream>>>(pm, ps, lse_col, N, nstrips);
  return check_launch("assign_lse");
}

int lgb200_assign_scores(const float* sim, const float* lse_row, const float* lse_col, const float* ls0,
                         const float* ls1, const float* dust0, const float* dust1, const uint8_t* gt, float* scores,
                         float* rowmax, int* rowarg, float* colmax, int* colarg, float* pos_row_sum,
                         float* row_expsum, void* ws, int B, int M, int N, cudaStream_t stream) {
  LGB_REQUIRE(sim && lse_row && lse_col && ls0 && ls1 && dust0 && dust1 && rowmax && rowarg && colmax && colarg && ws,
              kErrInvalid, "assign_scores: null pointer");
  LGB_REQUIRE(B > 0 && M > 0 && N > 0, kErrInvalid, "assign_scores: empty input");
  LGB_REQUIRE(!pos_row_sum || gt, kErrInvalid, "assign_scores: pos_row_sum requested without gt");
  const int nstrips = (M + kStripRows - 1) / kStripRows;
  ScoreArgs a;
  a.sim = sim; a.lse_row = lse_row; a.lse_col = lse_col; a.ls0 = ls0; a.ls1 = ls1; a.dust0 = dust0; a.dust1 = dust1;
  a.gt = gt; a.scores = scores; a.rowmax = rowmax; a.rowarg = rowarg;
  a.colpart_v = static_cast<float*>(ws);
  a.colpart_i = reinterpret_cast<int*>(a.colpart_v + (size_t)B * nstrips * N);
  a.pos_row_sum = pos_row_sum; a.row_expsum = row_expsum; a.M = M; a.N = N; a.nstrips = nstrips;
  assign_scores_kernel<<<dim3(nstrips, B), 256, 0, stream>>>(a);
  assign_col_arg_merge_kernel<<<dim3((N + 1 + 31) / 32, B), 256, 0, stream>>>(a.colpart_v, a.colpart_i, dust1, colmax,
                                                                              colarg, scores, M, N, nstrips);
  return check_launch("assign_scores");
}

int lgb200_assign_bwd(const float* sim, const float* lse_row, const float* lse_col, const uint8_t* gt,
                      const float* gcoef, const float* a_row, const float* a_col, void* dsim, int out_dtype, int B,
                      int M, int N, cudaStream_t stream) {
  LGB_REQUIRE(sim && lse_row && lse_col && gt && gcoef && a_row && a_col && dsim, kErrInvalid,
              "assign_bwd: null pointer");
  LGB_REQUIRE(B > 0 && M > 0 && N > 0, kErrInvalid, "assign_bwd: empty input");
  dim3 grid((M + kWarps - 1) / kWarps, B);
  if (out_dtype == LGB200_F32)
    assign_bwd_kernel<float><<<grid, 256, 0, stream>>>(sim, lse_row, lse_col, gt, gcoef, a_row, a_col,
                                                       static_cast<float*>(dsim), M, N);
  else if (out_dtype == LGB200_BF16)
    assign_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(sim, lse_row, lse_col, gt, gcoef, a_row, a_col,
                                                               static_cast<__nv_bfloat16*>(dsim), M, N);
  else
    LGB_REQUIRE(false, kErrInvalid, "assign_bwd: bad out_dtype %d", out_dtype);
  return check_launch("assign_bwd");
}

int lgb200_filter_matches(const float* rowmax, const int* rowarg, const int* colarg, float th, int64_t* m0,
                          int64_t* m1, float* ms0, float* ms1, int B, int M, int N, cudaStream_t stream) {
  LGB_REQUIRE(rowmax && rowarg && colarg && m0 && m1 && ms0 && ms1, kErrInvalid, "filter_matches: null pointer");
  LGB_REQUIRE(B > 0 && M > 0 && N > 0, kErrInvalid, "filter_matches: empty input");
  const int L = M > N ? M : N;
  filter_matches_kernel<<<dim3((L + 255) / 256, B), 256, 0, stream>>>(rowmax, rowarg, colarg, th, m0, m1, ms0, ms1, M,
                                                                      N);
  return check_launch("filter_matches");
}

}  // extern "C"
