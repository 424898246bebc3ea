// Ground-truth correspondences from a homography, on the device.
// Restates gluefactory/geometry/gt_generation.py:109-161 (gt_matches_from_homography) without its ~10 dense
// [B,M,N] temporaries: the O(M+N) point warps stay in torch (warp_points_torch, homography.py:161-180); this file
// does the O(M N) part -- squared reprojection distances both ways, their max, row / column argmins, mutual check
// and thresholds -- and emits matches0/1 (int64: index, -1 unmatched, -2 ignored) plus the boolean assignment.
// Distances are formed with explicitly rounded multiplies and adds (no FMA contraction) so that every comparison
// sees exactly the value torch's elementwise ops produce; results are bit-identical up to argmin ties.
#include "common.cuh"
#include "lgb200.h"

namespace lgb {

constexpr int kGtRows = 64;  // rows of the distance matrix per CTA (8 per warp)

__device__ __forceinline__ float sqdist(float2 a, float2 b) {
  const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y);
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// torch.min order on (value, index): a NaN is the minimum (it propagates, first NaN wins), otherwise the smaller
// value, and the lower index among equals -- so an all-inf / NaN row still yields a valid index (degenerate H).
__device__ __forceinline__ bool gt_better(float d, int i, float bd, int bi) {
  const bool dn = d != d, bn = bd != bd;
  if (dn != bn) return dn;
  if (!dn && d != bd) return d < bd;
  return i < bi;
}
// running minimum that propagates NaN like torch.min
__device__ __forceinline__ float gt_min(float a, float b) { return (a != a || b != b) ? NAN : fminf(a, b); }

// rows: min_j dist, argmin_j dist, min_j d0.  columns: the same over this strip's rows, one partial per strip.
__global__ void __launch_bounds__(256) gt_h_scan_kernel(const float2* __restrict__ kp0, const float2* __restrict__ kp1,
                                                       const float2* __restrict__ kp0_1,
                                                       const float2* __restrict__ kp1_0, float* __restrict__ row_dist,
                                                       int* __restrict__ row_arg, float* __restrict__ row_d0,
                                                       float* __restrict__ part_dist, int* __restrict__ part_arg,
                                                       float* __restrict__ part_d1, int M, int N, int nstrips,
                                                       const uint8_t* __restrict__ vis0 /* [B,M] or null */,
                                                       const uint8_t* __restrict__ vis1 /* [B,N] or null */) {
  // vis0 / vis1 (pose + depth ground truth, gt_generation.py:47-54): the joint distance of a pair counts only when both
  // points are visible in the other view (else +inf); the one-way minima d0 / d1 are taken over ALL pairs.
  const int b = blockIdx.y, strip = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint8_t* v0 = vis0 ? vis0 + (int64_t)b * M : nullptr;
  const uint8_t* v1 = vis1 ? vis1 + (int64_t)b * N : nullptr;
  const float2* p0 = kp0 + (int64_t)b * M;
  const float2* p01 = kp0_1 + (int64_t)b * M;
  const float2* p1 = kp1 + (int64_t)b * N;
  const float2* p10 = kp1_0 + (int64_t)b * N;
  const int r_base = strip * kGtRows;
  // ---- row pass: this warp's 8 rows against every column (lanes stride the columns)
  {
    float2 a0[8], a1[8];
    float bd[8], bd0[8];
    int bi[8];
    bool rv[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int i = r_base + warp * 8 + t;
      a0[t] = i < M ? p01[i] : make_float2(0.f, 0.f);
      a1[t] = i < M ? p0[i] : make_float2(0.f, 0.f);
      rv[t] = (v0 && i < M) ? v0[i] != 0 : true;
      bd[t] = INFINITY; bd0[t] = INFINITY; bi[t] = 0x7fffffff;
    }
    for (int j = lane; j < N; j += 32) {
      const float2 c1 = p1[j], c10 = p10[j];
      const bool cv = v1 ? v1[j] != 0 : true;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float d0 = sqdist(a0[t], c1), d1 = sqdist(a1[t], c10);
        float d = (d0 != d0 || d1 != d1) ? NAN : fmaxf(d0, d1);  // torch.max propagates NaN
        if (!(rv[t] && cv)) d = INFINITY;                          // torch.where(mask_visible, dist, inf)
        if (d < bd[t] || (d != d && bd[t] == bd[t])) { bd[t] = d; bi[t] = j; }  // j ascends per lane: first minimum kept
        bd0[t] = gt_min(bd0[t], d0);
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t)
      if (bi[t] == 0x7fffffff && lane < N) bi[t] = lane;  // every distance of this lane was +inf: its first column
#pragma unroll
    for (int t = 0; t < 8; ++t) {
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd[t], o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi[t], o);
        if (gt_better(od, oi, bd[t], bi[t])) { bd[t] = od; bi[t] = oi; }
        bd0[t] = gt_min(bd0[t], __shfl_xor_sync(0xffffffffu, bd0[t], o));
      }
      const int i = r_base + warp * 8 + t;
      if (lane == 0 && i < M) {
        row_dist[(int64_t)b * M + i] = bd[t];
        row_arg[(int64_t)b * M + i] = bi[t];
        row_d0[(int64_t)b * M + i] = bd0[t];
      }
    }
  }
  // ---- column pass: thread owns columns j = tid, tid + 256, ...; all rows of this strip
  __shared__ float2 s_a0[kGtRows], s_a1[kGtRows];
  __shared__ uint8_t s_rv[kGtRows];
  if (threadIdx.x < kGtRows) {
    const int i = r_base + threadIdx.x;
    s_a0[threadIdx.x] = i < M ? p01[i] : make_float2(0.f, 0.f);
    s_a1[threadIdx.x] = i < M ? p0[i] : make_float2(0.f, 0.f);
    s_rv[threadIdx.x] = (v0 && i < M) ? v0[i] : 1;
  }
  __syncthreads();
  const int nrows = min(kGtRows, M - r_base);
  for (int j = threadIdx.x; j < N; j += 256) {
    const float2 c1 = p1[j], c10 = p10[j];
    float bd = INFINITY, bd1 = INFINITY;
    int bi = 0x7fffffff;
    const bool cv = v1 ? v1[j] != 0 : true;
    for (int t = 0; t < nrows; ++t) {
      const float d0 = sqdist(s_a0[t], c1), d1 = sqdist(s_a1[t], c10);
      float d = (d0 != d0 || d1 != d1) ? NAN : fmaxf(d0, d1);
      if (!(s_rv[t] && cv)) d = INFINITY;
      if (d < bd || (d != d && bd == bd)) { bd = d; bi = r_base + t; }
      bd1 = gt_min(bd1, d1);
    }
    if (bi == 0x7fffffff) bi = r_base;  // all +inf: first row of the strip (nrows >= 1)
    const int64_t o = ((int64_t)b * nstrips + strip) * N + j;
    part_dist[o] = bd;
    part_arg[o] = bi;
    part_d1[o] = bd1;
  }
}

// merge the strip partials of every column (strip order == ascending row index: first minimum kept)
__global__ void __launch_bounds__(256) gt_h_colmerge_kernel(const float* __restrict__ part_dist,
                                                           const int* __restrict__ part_arg,
                                                           const float* __restrict__ part_d1,
                                                           float* __restrict__ col_dist, int* __restrict__ col_arg,
                                                           float* __restrict__ col_d1, int N, int nstrips) {
  const int b = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
  if (j >= N) return;
  float bd = INFINITY, bd1 = INFINITY;
  int bi = 0x7fffffff;
  for (int s = 0; s < nstrips; ++s) {
    const int64_t o = ((int64_t)b * nstrips + s) * N + j;
    const float d = part_dist[o];
    if (gt_better(d, part_arg[o], bd, bi)) { bd = d; bi = part_arg[o]; }
    bd1 = gt_min(bd1, part_d1[o]);
  }
  col_dist[(int64_t)b * N + j] = bd;
  col_arg[(int64_t)b * N + j] = bi;
  col_d1[(int64_t)b * N + j] = bd1;
}

// labels: mutual nearest neighbours below pos_th^2 are positives (index), points whose best one-way reprojection
// error exceeds neg_th^2 are unmatched (-1), everything else is ignored (-2)
__global__ void __launch_bounds__(256) gt_h_label_kernel(const float* __restrict__ row_dist,
                                                        const int* __restrict__ row_arg,
                                                        const float* __restrict__ row_d0,
                                                        const float* __restrict__ col_dist,
                                                        const int* __restrict__ col_arg,
                                                        const float* __restrict__ col_d1, float pos2, float neg2,
                                                        int64_t* __restrict__ m0, int64_t* __restrict__ m1,
                                                        uint8_t* __restrict__ assignment, int M, int N,
                                                        const uint8_t* __restrict__ valid0,
                                                        const uint8_t* __restrict__ valid1,
                                                        uint8_t* __restrict__ assignment_t /* [B,N,M] or null */) {
  // valid0 / valid1 (or null): "unmatched" additionally requires a valid depth (gt_generation.py:66-67)
  const int b = blockIdx.y, t = blockIdx.x * 256 + threadIdx.x;
  if (t < M) {
    const int64_t o = (int64_t)b * M + t;
    const int j = row_arg[o];
    const bool pos = (unsigned)j < (unsigned)N && col_arg[(int64_t)b * N + j] == t && row_dist[o] < pos2;
    int64_t m = pos ? (int64_t)j : -2;
    if (row_d0[o] > neg2 && (!valid0 || valid0[o])) m = -1;
    m0[o] = m;
    if (pos && assignment) assignment[((int64_t)b * M + t) * N + j] = 1;
    if (pos && assignment_t) assignment_t[((int64_t)b * N + j) * M + t] = 1;  // the transposed copy, for column-wise readers
  }
  if (t < N) {
    const int64_t o = (int64_t)b * N + t;
    const int i = col_arg[o];
    const bool pos = (unsigned)i < (unsigned)M && row_arg[(int64_t)b * M + i] == t && col_dist[o] < pos2;
    int64_t m = pos ? (int64_t)i : -2;
    if (col_d1[o] > neg2 && (!valid1 || valid1[o])) m = -1;
    m1[o] = m;
  }
}

// Extra "unmatched" labels from the epipolar geometry (gt_generation.py:82-90): a point WITHOUT valid depth whose label is
// still "ignore" becomes unmatched when every still-ignored point of the other view lies further than th from its
// epipolar line (symmetric distance, geometry/epipolar.py:59-72).  Pass 1 evaluates the per-point flags from a snapshot
// of the labels, pass 2 applies them (both sides use the labels BEFORE the update, like the reference).
__device__ __forceinline__ float sym_epi(float2 a, float2 bq, const float* F, float3 Fa /* F (a,1) */) {
  // |b^T F a| / |(F a)_xy|  and  / |(F^T b)_xy|, averaged
  const float num = fabsf(bq.x * Fa.x + bq.y * Fa.y + Fa.z);
  const float ftx = F[0] * bq.x + F[3] * bq.y + F[6], fty = F[1] * bq.x + F[4] * bq.y + F[7];
  const float d0 = num / sqrtf(Fa.x * Fa.x + Fa.y * Fa.y + 1e-15f);
  const float d1 = num / sqrtf(ftx * ftx + fty * fty + 1e-15f);
  return (d0 + d1) * 0.5f;
}

__global__ void __launch_bounds__(256) gt_epi_flags_kernel(const float2* __restrict__ kp0, const float2* __restrict__ kp1,
                                                          const float* __restrict__ Fm, const int64_t* __restrict__ m0,
                                                          const int64_t* __restrict__ m1,
                                                          const uint8_t* __restrict__ valid0,
                                                          const uint8_t* __restrict__ valid1, float th,
                                                          uint8_t* __restrict__ ex0, uint8_t* __restrict__ ex1, int M,
                                                          int N) {
  // one warp per point (rows of view 0 first, then the columns = points of view 1)
  const int b = blockIdx.y, w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= M + N) return;
  const float* F = Fm + (int64_t)b * 9;
  const bool side1 = w >= M;
  const int idx = side1 ? w - M : w;
  const int64_t self = side1 ? (int64_t)b * N + idx : (int64_t)b * M + idx;
  const bool cand = (side1 ? m1[self] : m0[self]) == -2 && !(side1 ? valid1[self] : valid0[self]);
  uint8_t* out = side1 ? ex1 + self : ex0 + self;
  if (!cand) {
    if (lane == 0) *out = 0;
    return;
  }
  const float2 p = (side1 ? kp1 : kp0)[self];
  const int other_n = side1 ? M : N;
  const float2* op = side1 ? kp0 + (int64_t)b * M : kp1 + (int64_t)b * N;
  const int64_t* om = side1 ? m0 + (int64_t)b * M : m1 + (int64_t)b * N;
  float best = INFINITY;
  for (int j = lane; j < other_n; j += 32) {
    if (om[j] != -2) continue;
    const float2 a = side1 ? op[j] : p, bq = side1 ? p : op[j];  // a in view 0, bq in view 1
    const float3 Fa = make_float3(F[0] * a.x + F[1] * a.y + F[2], F[3] * a.x + F[4] * a.y + F[5],
                                  F[6] * a.x + F[7] * a.y + F[8]);
    best = fminf(best, sym_epi(a, bq, F, Fa));
  }
  best = -warp_max(-best);
  if (lane == 0) *out = best > th ? 1 : 0;  // no ignored partner at all: min over an all-inf row = inf > th
}

__global__ void __launch_bounds__(256) gt_epi_apply_kernel(int64_t* __restrict__ m0, int64_t* __restrict__ m1,
                                                          const uint8_t* __restrict__ ex0,
                                                          const uint8_t* __restrict__ ex1, int64_t n0, int64_t n1) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t < n0 && ex0[t]) m0[t] = -1;
  if (t < n1 && ex1[t]) m1[t] = -1;
}

}  // namespace lgb

using namespace lgb;

extern "C" {

size_t lgb200_gt_homography_ws_bytes(int B, int M, int N) {
  const size_t nstrips = (size_t)(M + kGtRows - 1) / kGtRows;
  return (size_t)B * (3 * (size_t)M + 3 * (size_t)N + 3 * nstrips * (size_t)N) * 4;
}

int lgb200_gt_from_reprojection(const float* kp0, const float* kp1, const float* kp0_1, const float* kp1_0,
                                const uint8_t* vis0, const uint8_t* vis1, const uint8_t* valid0, const uint8_t* valid1,
                                float pos_th, float neg_th, int64_t* m0, int64_t* m1, uint8_t* assignment,
                                uint8_t* assignment_t, void* ws, int B, int M, int N, cudaStream_t stream) {
  LGB_REQUIRE(kp0 && kp1 && kp0_1 && kp1_0 && m0 && m1 && ws, kErrInvalid, "gt_from_reprojection: null pointer");
  LGB_REQUIRE(B > 0 && M > 0 && N > 0, kErrInvalid, "gt_from_reprojection: empty input B=%d M=%d N=%d", B, M, N);
  LGB_REQUIRE(((reinterpret_cast<uintptr_t>(kp0) | reinterpret_cast<uintptr_t>(kp1) |
                reinterpret_cast<uintptr_t>(kp0_1) | reinterpret_cast<uintptr_t>(kp1_0)) & 7) == 0,
              kErrInvalid, "gt_from_reprojection: keypoint arrays must be 8-byte aligned");
  LGB_REQUIRE((vis0 == nullptr) == (vis1 == nullptr) && (valid0 == nullptr) == (valid1 == nullptr), kErrInvalid,
              "gt_from_reprojection: visibility / validity masks come in pairs");
  const int nstrips = (M + kGtRows - 1) / kGtRows;
  float* row_dist = static_cast<float*>(ws);
  int* row_arg = reinterpret_cast<int*>(row_dist + (size_t)B * M);
  float* row_d0 = reinterpret_cast<float*>(row_arg + (size_t)B * M);
  float* col_dist = row_d0 + (size_t)B * M;
  int* col_arg = reinterpret_cast<int*>(col_dist + (size_t)B * N);
  float* col_d1 = reinterpret_cast<float*>(col_arg + (size_t)B * N);
  float* part_dist = col_d1 + (size_t)B * N;
  int* part_arg = reinterpret_cast<int*>(part_dist + (size_t)B * nstrips * N);
  float* part_d1 = reinterpret_cast<float*>(part_arg + (size_t)B * nstrips * N);
  if (assignment) {
    cudaError_t e = cudaMemsetAsync(assignment, 0, (size_t)B * M * N, stream);
    LGB_REQUIRE(e == cudaSuccess, kErrCuda, "gt_from_reprojection: memset: %s", cudaGetErrorString(e));
  }
  if (assignment_t) {
    cudaError_t e = cudaMemsetAsync(assignment_t, 0, (size_t)B * M * N, stream);
    LGB_REQUIRE(e == cudaSuccess, kErrCuda, "gt_from_reprojection: memset: %s", cudaGetErrorString(e));
  }
  gt_h_scan_kernel<<<dim3(nstrips, B), 256, 0, stream>>>(
      reinterpret_cast<const float2*>(kp0), reinterpret_cast<const float2*>(kp1),
      reinterpret_cast<const float2*>(kp0_1), reinterpret_cast<const float2*>(kp1_0), row_dist, row_arg, row_d0,
      part_dist, part_arg, part_d1, M, N, nstrips, vis0, vis1);
  gt_h_colmerge_kernel<<<dim3((N + 255) / 256, B), 256, 0, stream>>>(part_dist, part_arg, part_d1, col_dist, col_arg,
                                                                     col_d1, N, nstrips);
  const int L = M > N ? M : N;
  gt_h_label_kernel<<<dim3((L + 255) / 256, B), 256, 0, stream>>>(row_dist, row_arg, row_d0, col_dist, col_arg, col_d1,
                                                                  pos_th * pos_th, neg_th * neg_th, m0, m1, assignment,
                                                                  M, N, valid0, valid1, assignment_t);
  return check_launch("gt_from_reprojection");
}

int lgb200_gt_from_homography(const float* kp0, const float* kp1, const float* kp0_1, const float* kp1_0, float pos_th,
                              float neg_th, int64_t* m0, int64_t* m1, uint8_t* assignment, void* ws, int B, int M,
                              int N, cudaStream_t stream) {
  return lgb200_gt_from_reprojection(kp0, kp1, kp0_1, kp1_0, nullptr, nullptr, nullptr, nullptr, pos_th, neg_th, m0, m1,
                                     assignment, nullptr, ws, B, M, N, stream);
}

int lgb200_gt_epipolar_unmatched(const float* kp0, const float* kp1, const float* F, const uint8_t* valid0,
                                 const uint8_t* valid1, float th, int64_t* m0, int64_t* m1, uint8_t* ws, int B, int M,
                                 int N, cudaStream_t stream) {
  LGB_REQUIRE(kp0 && kp1 && F && valid0 && valid1 && m0 && m1 && ws, kErrInvalid, "gt_epipolar_unmatched: null pointer");
  LGB_REQUIRE(B > 0 && M > 0 && N > 0, kErrInvalid, "gt_epipolar_unmatched: empty input");
  uint8_t* ex0 = ws;
  uint8_t* ex1 = ws + (size_t)B * M;
  gt_epi_flags_kernel<<<dim3((M + N + 7) / 8, B), 256, 0, stream>>>(reinterpret_cast<const float2*>(kp0),
                                                                    reinterpret_cast<const float2*>(kp1), F, m0, m1, valid0,
                                                                    valid1, th, ex0, ex1, M, N);
  const int64_t n0 = (int64_t)B * M, n1 = (int64_t)B * N, L = n0 > n1 ? n0 : n1;
  gt_epi_apply_kernel<<<(unsigned)((L + 255) / 256), 256, 0, stream>>>(m0, m1, ex0, ex1, n0, n1);
  return check_launch("gt_epipolar_unmatched");
}

}  // extern "C"
