// fp32 CUDA-core attention (forward + backward) -- the full-precision parity path
// (`precision="fp32"`) and the on-device cross-check for the tcgen05 kernels.
//
//   out = softmax(scale * q k^T) v       per (batch, head), head_dim = 64
// (reference: gluefactory/models/matchers/lightglue.py:118-121 self-attention,
//  :207-216 cross-attention evaluated as two one-directional passes).
//
// Tensors are token-major [B, N, H, 64].  The key/value batch of query batch b is
// (b + kv_shift) % B, which expresses the two directions of cross-attention on the
// concatenated [image0; image1] batch as ONE launch.
//
// Work split: two threads per row ("row" = query in fwd/dq, key in dkv), each owning 32 of the 64
// channels; dot products are completed with one shuffle.  The opposite operand is staged through
// shared memory in 64-row tiles and read as warp-wide broadcasts.
#include <math.h>

#include "common.cuh"
#include "lgb200.h"

namespace lgb {

constexpr int kD = 64;
constexpr int kHalf = 32;
constexpr int kTile = 64;  // rows per smem tile == rows per CTA

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// cooperative load of `rows` x 64 (row stride ld elements) into smem [kTile][kD], zero-filled past nvalid
template <typename T>
__device__ __forceinline__ void load_tile(float (*dst)[kD], const T* src, int64_t ld, int nvalid) {
  for (int idx = threadIdx.x; idx < kTile * kD; idx += blockDim.x) {
    const int r = idx >> 6, c = idx & 63;
    dst[r][c] = r < nvalid ? to_f<T>(src[(int64_t)r * ld + c]) : 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(128) attn_fwd_simt_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                           const T* __restrict__ v, T* __restrict__ out,
                                                           float* __restrict__ lse, int B, int Nq, int Nk, int H,
                                                           int kv_shift, float scale) {
  __shared__ float sk[kTile][kD];
  __shared__ float sv[kTile][kD];
  const int b = blockIdx.z, h = blockIdx.y;
  const int kb = (b + kv_shift) % B;
  const int row = blockIdx.x * kTile + (threadIdx.x >> 1);
  const int half = threadIdx.x & 1;
  const int64_t ld = (int64_t)H * kD;
  const bool rvalid = row < Nq;
  float qr[kHalf], o[kHalf];
#pragma unroll
  for (int d = 0; d < kHalf; ++d) {
    qr[d] = rvalid ? to_f<T>(q[((int64_t)b * Nq + row) * ld + h * kD + half * kHalf + d]) * scale : 0.f;
    o[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < Nk; j0 += kTile) {
    const int nv = min(kTile, Nk - j0);
    __syncthreads();
    load_tile<T>(sk, k + ((int64_t)kb * Nk + j0) * ld + h * kD, ld, nv);
    load_tile<T>(sv, v + ((int64_t)kb * Nk + j0) * ld + h * kD, ld, nv);
    __syncthreads();
    for (int j = 0; j < nv; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < kHalf; ++d) s = fmaf(qr[d], sk[j][half * kHalf + d], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (s > m) {
        const float a = expf(m - s);
        l *= a;
#pragma unroll
        for (int d = 0; d < kHalf; ++d) o[d] *= a;
        m = s;
      }
      const float p = expf(s - m);
      l += p;
#pragma unroll
      for (int d = 0; d < kHalf; ++d) o[d] = fmaf(p, sv[j][half * kHalf + d], o[d]);
    }
  }
  if (rvalid) {
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < kHalf; ++d)
      out[((int64_t)b * Nq + row) * ld + h * kD + half * kHalf + d] = from_f<T>(o[d] * inv);
    if (half == 0) lse[((int64_t)b * H + h) * Nq + row] = m + logf(l);
  }
}

// delta[b,h,i] = sum_d dout[b,i,h,d] * out[b,i,h,d]   (8 lanes per row)
template <typename T>
__global__ void __launch_bounds__(256) attn_delta_kernel(const T* __restrict__ out, const T* __restrict__ dout,
                                                        float* __restrict__ delta, int64_t nrows /*B*N*H*/, int N,
                                                        int H) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = gid >> 3;  // (b, n, h) flattened token-major
  const int sub = (int)(gid & 7);
  float acc = 0.f;
  if (r < nrows) {
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += to_f<T>(out[r * kD + sub * 8 + e]) * to_f<T>(dout[r * kD + sub * 8 + e]);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (r < nrows && sub == 0) {
    const int h = (int)(r % H);
    const int64_t bn = r / H;
    const int n = (int)(bn % N);
    const int64_t b = bn / N;
    delta[(b * H + h) * N + n] = acc;
  }
}

// dq_i = scale * sum_j p_ij (dp_ij - delta_i) k_j
template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_dq_simt_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                              const T* __restrict__ v, const T* __restrict__ dout,
                                                              const float* __restrict__ lse,
                                                              const float* __restrict__ delta, T* __restrict__ dq,
                                                              int B, int Nq, int Nk, int H, int kv_shift,
                                                              float scale) {
  __shared__ float sk[kTile][kD];
  __shared__ float sv[kTile][kD];
  const int b = blockIdx.z, h = blockIdx.y;
  const int kb = (b + kv_shift) % B;
  const int row = blockIdx.x * kTile + (threadIdx.x >> 1);
  const int half = threadIdx.x & 1;
  const int64_t ld = (int64_t)H * kD;
  const bool rvalid = row < Nq;
  float qr[kHalf], go[kHalf], acc[kHalf];
#pragma unroll
  for (int d = 0; d < kHalf; ++d) {
    const int64_t off = ((int64_t)b * Nq + row) * ld + h * kD + half * kHalf + d;
    qr[d] = rvalid ? to_f<T>(q[off]) * scale : 0.f;
    go[d] = rvalid ? to_f<T>(dout[off]) : 0.f;
    acc[d] = 0.f;
  }
  const float L = rvalid ? lse[((int64_t)b * H + h) * Nq + row] : 0.f;
  const float Dl = rvalid ? delta[((int64_t)b * H + h) * Nq + row] : 0.f;
  for (int j0 = 0; j0 < Nk; j0 += kTile) {
    const int nv = min(kTile, Nk - j0);
    __syncthreads();
    load_tile<T>(sk, k + ((int64_t)kb * Nk + j0) * ld + h * kD, ld, nv);
    load_tile<T>(sv, v + ((int64_t)kb * Nk + j0) * ld + h * kD, ld, nv);
    __syncthreads();
    for (int j = 0; j < nv; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < kHalf; ++d) {
        s = fmaf(qr[d], sk[j][half * kHalf + d], s);
        dp = fmaf(go[d], sv[j][half * kHalf + d], dp);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      dp += __shfl_xor_sync(0xffffffffu, dp, 1);
      const float ds = expf(s - L) * (dp - Dl) * scale;
#pragma unroll
      for (int d = 0; d < kHalf; ++d) acc[d] = fmaf(ds, sk[j][half * kHalf + d], acc[d]);
    }
  }
  if (rvalid) {
#pragma unroll
    for (int d = 0; d < kHalf; ++d)
      dq[((int64_t)b * Nq + row) * ld + h * kD + half * kHalf + d] = from_f<T>(acc[d]);
  }
}

// dv_j = sum_i p_ij do_i ;  dk_j = scale * sum_i p_ij (dp_ij - delta_i) q_i
// launched over KEY batches kb; the query batch is qb = (kb - kv_shift) mod B.
template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_dkv_simt_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                               const T* __restrict__ v, const T* __restrict__ dout,
                                                               const float* __restrict__ lse,
                                                               const float* __restrict__ delta, T* __restrict__ dk,
                                                               T* __restrict__ dv, int B, int Nq, int Nk, int H,
                                                               int kv_shift, float scale) {
  __shared__ float sq[kTile][kD];
  __shared__ float sdo[kTile][kD];
  __shared__ float sl[kTile], sd[kTile];
  const int kb = blockIdx.z, h = blockIdx.y;
  const int qb = ((kb - kv_shift) % B + B) % B;
  const int row = blockIdx.x * kTile + (threadIdx.x >> 1);
  const int half = threadIdx.x & 1;
  const int64_t ld = (int64_t)H * kD;
  const bool rvalid = row < Nk;
  float kr[kHalf], vr[kHalf], gk[kHalf], gv[kHalf];
#pragma unroll
  for (int d = 0; d < kHalf; ++d) {
    const int64_t off = ((int64_t)kb * Nk + row) * ld + h * kD + half * kHalf + d;
    kr[d] = rvalid ? to_f<T>(k[off]) * scale : 0.f;
    vr[d] = rvalid ? to_f<T>(v[off]) : 0.f;
    gk[d] = gv[d] = 0.f;
  }
  for (int i0 = 0; i0 < Nq; i0 += kTile) {
    const int nv = min(kTile, Nq - i0);
    __syncthreads();
    load_tile<T>(sq, q + ((int64_t)qb * Nq + i0) * ld + h * kD, ld, nv);
    load_tile<T>(sdo, dout + ((int64_t)qb * Nq + i0) * ld + h * kD, ld, nv);
    if (threadIdx.x < kTile) {
      const bool ok = (int)threadIdx.x < nv;
      sl[threadIdx.x] = ok ? lse[((int64_t)qb * H + h) * Nq + i0 + threadIdx.x] : 0.f;
      sd[threadIdx.x] = ok ? delta[((int64_t)qb * H + h) * Nq + i0 + threadIdx.x] : 0.f;
    }
    __syncthreads();
    for (int i = 0; i < nv; ++i) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < kHalf; ++d) {
        s = fmaf(kr[d], sq[i][half * kHalf + d], s);
        dp = fmaf(vr[d], sdo[i][half * kHalf + d], dp);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      dp += __shfl_xor_sync(0xffffffffu, dp, 1);
      const float p = expf(s - sl[i]);
      const float ds = p * (dp - sd[i]) * scale;
#pragma unroll
      for (int d = 0; d < kHalf; ++d) {
        gv[d] = fmaf(p, sdo[i][half * kHalf + d], gv[d]);
        gk[d] = fmaf(ds, sq[i][half * kHalf + d], gk[d]);
      }
    }
  }
  if (rvalid) {
#pragma unroll
    for (int d = 0; d < kHalf; ++d) {
      const int64_t off = ((int64_t)kb * Nk + row) * ld + h * kD + half * kHalf + d;
      dk[off] = from_f<T>(gk[d]);
      dv[off] = from_f<T>(gv[d]);
    }
  }
}

template <typename T>
int attn_fwd_simt(const void* q, const void* k, const void* v, void* out, float* lse, int B, int Nq, int Nk, int H,
                  int kv_shift, float scale, cudaStream_t stream) {
  dim3 grid((Nq + kTile - 1) / kTile, H, B);
  attn_fwd_simt_kernel<T><<<grid, 128, 0, stream>>>((const T*)q, (const T*)k, (const T*)v, (T*)out, lse, B, Nq, Nk, H,
                                                    kv_shift, scale);
  return check_launch("attn_fwd_simt");
}

template <typename T>
int attn_delta(const void* out, const void* dout, float* delta, int B, int N, int H, cudaStream_t stream) {
  const int64_t nrows = (int64_t)B * N * H;
  const unsigned grid = (unsigned)((nrows * 8 + 255) / 256);
  attn_delta_kernel<T><<<grid, 256, 0, stream>>>((const T*)out, (const T*)dout, delta, nrows, N, H);
  return check_launch("attn_delta");
}

template <typename T>
int attn_bwd_simt(const void* q, const void* k, const void* v, const void* out, const float* lse, const void* dout,
                  void* dq, void* dk, void* dv, float* delta, int B, int Nq, int Nk, int H, int kv_shift, float scale,
                  cudaStream_t stream) {
  int rc = attn_delta<T>(out, dout, delta, B, Nq, H, stream);
  if (rc) return rc;
  dim3 gq((Nq + kTile - 1) / kTile, H, B);
  attn_bwd_dq_simt_kernel<T><<<gq, 128, 0, stream>>>((const T*)q, (const T*)k, (const T*)v, (const T*)dout, lse, delta,
                                                     (T*)dq, B, Nq, Nk, H, kv_shift, scale);
  dim3 gk((Nk + kTile - 1) / kTile, H, B);
  attn_bwd_dkv_simt_kernel<T><<<gk, 128, 0, stream>>>((const T*)q, (const T*)k, (const T*)v, (const T*)dout, lse,
                                                      delta, (T*)dk, (T*)dv, B, Nq, Nk, H, kv_shift, scale);
  return check_launch("attn_bwd_simt");
}

template int attn_fwd_simt<float>(const void*, const void*, const void*, void*, float*, int, int, int, int, int, float,
                                  cudaStream_t);
template int attn_fwd_simt<__nv_bfloat16>(const void*, const void*, const void*, void*, float*, int, int, int, int, int,
                                          float, cudaStream_t);
template int attn_bwd_simt<float>(const void*, const void*, const void*, const void*, const float*, const void*, void*,
                                  void*, void*, float*, int, int, int, int, int, float, cudaStream_t);
template int attn_bwd_simt<__nv_bfloat16>(const void*, const void*, const void*, const void*, const float*,
                                          const void*, void*, void*, void*, float*, int, int, int, int, int, float,
                                          cudaStream_t);
template int attn_delta<float>(const void*, const void*, float*, int, int, int, cudaStream_t);
template int attn_delta<__nv_bfloat16>(const void*, const void*, float*, int, int, int, cudaStream_t);

}  // namespace lgb
