// The two other assignment heads on the matcher path:
//   * log_double_softmax with a learned bin score (GlueStick, gluestick.py:772-783)
//   * log-domain Sinkhorn optimal transport with a dustbin (SuperGlue head,
//     gluefactory_nonfree/superglue.py:186-214), forward only.
// Both are HBM/L2-bound sweeps over the [B,M,N] similarity; the bordered (M+1)x(N+1) coupling
// matrix is never materialised: the dustbin row/column is handled analytically.
#include <math.h>

#include "common.cuh"
#include "lgb200.h"

namespace lgb {

constexpr float kNInf = -INFINITY;

__device__ __forceinline__ float logaddexp_f(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == kNInf) return kNInf;
  return m + log1pf(expf(-fabsf(a - b)));
}

// scores from precomputed lse_row / lse_col of the UNBORDERED sim
__global__ void __launch_bounds__(256) lds_scores_kernel(const float* __restrict__ sim, const float* __restrict__ lse_row,
                                                        const float* __restrict__ lse_col, float beta,
                                                        float* __restrict__ scores, int M, int N) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i > M) return;
  float* orow = scores + ((int64_t)b * (M + 1) + i) * (N + 1);
  if (i == M) {
    for (int j = lane; j <= N; j += 32)
      orow[j] = j < N ? beta - logaddexp_f(lse_col[(int64_t)b * N + j], beta) : 0.f;
    return;
  }
  const float lr = logaddexp_f(lse_row[(int64_t)b * M + i], beta);
  const float* srow = sim + ((int64_t)b * M + i) * N;
  for (int j = lane; j < N; j += 32) {
    const float x = srow[j];
    const float lc = logaddexp_f(lse_col[(int64_t)b * N + j], beta);
    orow[j] = ((x - lr) + (x - lc)) * 0.5f;
  }
  if (lane == 0) orow[N] = beta - lr;
}

// ---- Sinkhorn -------------------------------------------------------------------------------------
// u_i = log_mu_i - LSE_j (Z_ij + v_j), i in [0, M]   (Z_iN = Z_Mj = alpha)
__global__ void __launch_bounds__(256) sk_row_kernel(const float* __restrict__ sim, const float* __restrict__ v,
                                                    float* __restrict__ u, float alpha, float norm, int M, int N) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i > M) return;
  const float* vb = v + (int64_t)b * (N + 1);
  float m = kNInf, s = 0.f;
  if (i < M) {
    const float* srow = sim + ((int64_t)b * M + i) * N;
    for (int j = lane; j < N; j += 32) {
      const float x = srow[j] + vb[j];
      if (x > m) { s *= __expf(m - x); m = x; }
      s += __expf(x - (m == kNInf ? 0.f : m));
    }
  } else {
    for (int j = lane; j < N; j += 32) {
      const float x = alpha + vb[j];
      if (x > m) { s *= __expf(m - x); m = x; }
      s += __expf(x - (m == kNInf ? 0.f : m));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_merge(m, s, m2, s2);
  }
  if (lane == 0) {
    lse_merge(m, s, alpha + vb[N], 1.f);
    const float log_mu = i < M ? norm : logf((float)N) + norm;
    u[(int64_t)b * (M + 1) + i] = log_mu - (m + logf(s));
  }
}

constexpr int kSkSeg = 64;
// per-segment column partials of (sim_ij + u_i), thread per column
__global__ void __launch_bounds__(128) sk_col_part_kernel(const float* __restrict__ sim, const float* __restrict__ u,
                                                         float* __restrict__ pm, float* __restrict__ ps, int M, int N,
                                                         int nseg) {
  const int b = blockIdx.z, seg = blockIdx.y;
  const int j = blockIdx.x * 128 + threadIdx.x;
  if (j >= N) return;
  const float* ub = u + (int64_t)b * (M + 1);
  const int i1 = min(M, (seg + 1) * kSkSeg);
  float m = kNInf, s = 0.f;
  for (int i = seg * kSkSeg; i < i1; ++i) {
    const float x = sim[((int64_t)b * M + i) * N + j] + ub[i];
    if (x > m) { s *= __expf(m - x); m = x; }
    s += __expf(x - (m == kNInf ? 0.f : m));
  }
  const int64_t o = ((int64_t)b * nseg + seg) * N + j;
  pm[o] = m;
  ps[o] = s;
}
// v_j = log_nu_j - LSE_i (Z_ij + u_i), j in [0, N]
__global__ void __launch_bounds__(128) sk_col_merge_kernel(const float* __restrict__ pm, const float* __restrict__ ps,
                                                          const float* __restrict__ u, float* __restrict__ v,
                                                          float alpha, float norm, int M, int N, int nseg) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * 128 + threadIdx.x;
  if (j > N) return;
  const float* ub = u + (int64_t)b * (M + 1);
  float m = kNInf, s = 0.f;
  if (j < N) {
    for (int sg = 0; sg < nseg; ++sg) {
      const int64_t o = ((int64_t)b * nseg + sg) * N + j;
      lse_merge(m, s, pm[o], ps[o]);
    }
  } else {
    for (int i = 0; i < M; ++i) lse_merge(m, s, alpha + ub[i], 1.f);
  }
  lse_merge(m, s, alpha + ub[M], 1.f);
  const float log_nu = j < N ? norm : logf((float)M) + norm;
  v[(int64_t)b * (N + 1) + j] = log_nu - (m + logf(s));
}
__global__ void __launch_bounds__(256) sk_out_kernel(const float* __restrict__ sim, const float* __restrict__ u,
                                                    const float* __restrict__ v, float* __restrict__ out, float alpha,
                                                    float norm, int M, int N) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i > M) return;
  const float ui = u[(int64_t)b * (M + 1) + i] - norm;
  const float* vb = v + (int64_t)b * (N + 1);
  float* orow = out + ((int64_t)b * (M + 1) + i) * (N + 1);
  for (int j = lane; j <= N; j += 32) {
    const float z = (i < M && j < N) ? sim[((int64_t)b * M + i) * N + j] : alpha;
    orow[j] = z + ui + vb[j];
  }
}
__global__ void fill_kernel(float* p, int64_t n, float val) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = val;
}

}  // namespace lgb

using namespace lgb;

extern "C" {

size_t lgb200_heads_ws_bytes(int B, int M, int N) {
  return (size_t)4 * B * (M + N + 2) + (size_t)8 * B * ((M + 31) / 32) * N;
}

int lgb200_log_double_softmax(const float* sim, float bin_score, float* scores, void* ws, int B, int M, int N,
                              cudaStream_t stream) {
  LGB_REQUIRE(sim && scores && ws, kErrInvalid, "log_double_softmax: null pointer");
  LGB_REQUIRE(B > 0 && M > 0 && N > 0, kErrInvalid, "log_double_softmax: empty input");
  // ws layout: [lse_row B*M][lse_col B*N][assign_lse scratch]
  float* lse_row = static_cast<float*>(ws);
  float* lse_col = lse_row + (size_t)B * M;
  void* sub = lse_col + (size_t)B * N;
  int rc = lgb200_assign_lse(sim, lse_row, lse_col, sub, B, M, N, stream);
  if (rc) return rc;
  lds_scores_kernel<<<dim3((M + 1 + 7) / 8, B), 256, 0, stream>>>(sim, lse_row, lse_col, bin_score, scores, M, N);
  return check_launch("log_double_softmax");
}

int lgb200_sinkhorn(const float* sim, float alpha, int iters, float* out, void* ws, int B, int M, int N,
                    cudaStream_t stream) {
  LGB_REQUIRE(sim && out && ws, kErrInvalid, "sinkhorn: null pointer");
  LGB_REQUIRE(B > 0 && M > 0 && N > 0 && iters >= 0, kErrInvalid, "sinkhorn: bad arguments");
  const int nseg = (M + kSkSeg - 1) / kSkSeg;
  // ws layout: [u B*(M+1)][v B*(N+1)][pm B*nseg*N][ps B*nseg*N]
  float* u = static_cast<float*>(ws);
  float* v = u + (size_t)B * (M + 1);
  float* pm = v + (size_t)B * (N + 1);
  float* ps = pm + (size_t)B * nseg * N;
  const float norm = -logf((float)(M + N));
  const int64_t nv = (int64_t)B * (N + 1), nu = (int64_t)B * (M + 1);
  fill_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, stream>>>(v, nv, 0.f);
  fill_kernel<<<(unsigned)((nu + 255) / 256), 256, 0, stream>>>(u, nu, 0.f);
  for (int it = 0; it < iters; ++it) {
    sk_row_kernel<<<dim3((M + 1 + 7) / 8, B), 256, 0, stream>>>(sim, v, u, alpha, norm, M, N);
    sk_col_part_kernel<<<dim3((N + 127) / 128, nseg, B), 128, 0, stream>>>(sim, u, pm, ps, M, N, nseg);
    sk_col_merge_kernel<<<dim3((N + 1 + 127) / 128, B), 128, 0, stream>>>(pm, ps, u, v, alpha, norm, M, N, nseg);
  }
  sk_out_kernel<<<dim3((M + 1 + 7) / 8, B), 256, 0, stream>>>(sim, u, v, out, alpha, norm, M, N);
  return check_launch("sinkhorn");
}

}  // extern "C"
