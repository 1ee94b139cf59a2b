// Memory-bound kernels of the trunk on channels-last activations:
//   * stem max-pool  k[1,3,3] s[1,2,2] p[0,1,1]  (altfreezing/slowfast/models/stem_helper.py:166-168,177)
//   * pathway0_pool  k=s=[2,1,1]                 (video_model_builder.py:474-480,566-568)
//   * head: AvgPool3d over the whole [T/2,7,7] map + Linear(C->1) (+sigmoid)
//           (head_helper.py:74-95; sigmoid is the callers', demo.py:328)
//   * NDHWC -> NCTHW fp32 export for stage-parity tests
// Every thread moves 16 bytes of contiguous channels; grids are flat over output vectors.
#include "common.cuh"
#include "../../include/afb200.h"

namespace afb {
namespace {

template <typename T> struct Vec;   // 16-byte vector of T
template <> struct Vec<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ void load(const bf16* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
  }
  __device__ __forceinline__ void store(bf16* p) const {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

template <typename T>
__global__ void maxpool_spatial_kernel(const T* __restrict__ x, T* __restrict__ y, int BT, int H, int W,
                                       int C, int Ho, int Wo) {
  constexpr int V = Vec<T>::N;
  const int cv = C / V;
  const long long total = (long long)BT * Ho * Wo * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    unsigned r = (unsigned)i;              // total < 2^31 checked by the launcher
    const int c = (int)(r % (unsigned)cv) * V; r /= (unsigned)cv;
    const int wo = (int)(r % (unsigned)Wo); r /= (unsigned)Wo;
    const int ho = (int)(r % (unsigned)Ho); r /= (unsigned)Ho;
    const long long bt = r;
    Vec<T> m;
#pragma unroll
    for (int k = 0; k < V; ++k) m.v[k] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int h = ho * 2 - 1 + dy;
      if ((unsigned)h >= (unsigned)H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int w = wo * 2 - 1 + dx;
        if ((unsigned)w >= (unsigned)W) continue;
        Vec<T> t;
        t.load(x + ((bt * H + h) * W + w) * C + c);
#pragma unroll
        for (int k = 0; k < V; ++k) m.v[k] = fmaxf(m.v[k], t.v[k]);
      }
    }
    m.store(y + ((bt * Ho + ho) * Wo + wo) * C + c);
  }
}

template <typename T>
__global__ void maxpool_temporal_kernel(const T* __restrict__ x, T* __restrict__ y, long long BTo,
                                        long long frame_vecs) {
  constexpr int V = Vec<T>::N;
  const long long total = BTo * frame_vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long bt = i / frame_vecs, e = (i - bt * frame_vecs) * V;
    Vec<T> a, b;
    a.load(x + (2 * bt) * frame_vecs * V + e);
    b.load(x + (2 * bt + 1) * frame_vecs * V + e);
#pragma unroll
    for (int k = 0; k < V; ++k) a.v[k] = fmaxf(a.v[k], b.v[k]);
    a.store(y + bt * frame_vecs * V + e);
  }
}

// Sum over a slice of the P positions for a slab of channels: grid (C/(V*32), B, PZ), block (32, 8).
// threadIdx.x -> channel vector, threadIdx.y -> position phase, blockIdx.z -> position slice.  With PZ == 1 the block
// writes the mean; with PZ > 1 it writes its partial sum to feat[(b*PZ + z)*C + c] and head_fc_kernel adds the PZ
// partials in a fixed order and divides (a single clip is then pooled by 8*PZ blocks instead of 8: the head was the
// slowest kernel of the batch-1 pass).  PZ depends only on P, never on the batch, so results are batch-invariant.
constexpr int HEAD_PZ = 16;
template <typename T>
__global__ void head_pool_kernel(const T* __restrict__ x, float* __restrict__ feat, int P, int C, int PZ) {
  constexpr int V = Vec<T>::N;
  __shared__ float part[8][32 * V];
  const int b = blockIdx.y, z = blockIdx.z;
  const int c = (blockIdx.x * 32 + threadIdx.x) * V;
  const int per = (P + PZ - 1) / PZ, p_lo = z * per, p_hi = min(P, p_lo + per);
  float s[V];
#pragma unroll
  for (int k = 0; k < V; ++k) s[k] = 0.f;
  if (c < C) {
    const T* xb = x + (long long)b * P * C + c;
    for (int p = p_lo + threadIdx.y; p < p_hi; p += 8) {
      Vec<T> t;
      t.load(xb + (long long)p * C);
#pragma unroll
      for (int k = 0; k < V; ++k) s[k] += t.v[k];
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) part[threadIdx.y][threadIdx.x * V + k] = s[k];
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) t += part[j][threadIdx.x * V + k];
      feat[((long long)b * PZ + z) * C + c + k] = PZ == 1 ? t / (float)P : t;
    }
  }
}

// logits[b] = dot(feat[b], w) + bias ; optional sigmoid; optional copy of the features.  feat holds PZ partial sums
// per clip (PZ > 1: summed here in slice order and divided by P) or the finished mean (PZ == 1).
__global__ void head_fc_kernel(const float* __restrict__ feat, int PZ, float P, const float* __restrict__ w, float bias,
                               int C, float* __restrict__ feat_out, float* __restrict__ logits,
                               float* __restrict__ scores) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float f = feat[(long long)b * PZ * C + c];
    if (PZ > 1) {
      for (int z = 1; z < PZ; ++z) f += feat[((long long)b * PZ + z) * C + c];
      f = f / P;
    }
    if (feat_out) feat_out[(long long)b * C + c] = f;
    s = fmaf(f, w[c], s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) {
      const float l = s + bias;
      if (logits) logits[b] = l;
      if (scores) scores[b] = 1.f / (1.f + expf(-l));
    }
  }
}

template <typename T>
__global__ void ndhwc_to_ncthw_kernel(const T* __restrict__ x, float* __restrict__ y, int B, long long P, int C) {
  // tile transpose of [P, C] -> [C, P] per batch element
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const long long p = p0 + j;
    const int c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (p < P && c < C) ? (float)x[((long long)b * P + p) * C + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j;
    const long long p = p0 + threadIdx.x;
    if (p < P && c < C) y[((long long)b * C + c) * P + p] = tile[threadIdx.x][j];
  }
}

inline int flat_grid(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

int maxpool_spatial_launch(const void* x, void* y, int B, int T, int H, int W, int C, bool is_bf16,
                           cudaStream_t s) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int V = is_bf16 ? 8 : 4;
  if (C % V) { set_error("maxpool_spatial: C=%d not a multiple of %d", C, V); return AF_ERR_INVALID; }
  const long long total = (long long)B * T * Ho * Wo * (C / V);
  if (total >= (1LL << 31)) { set_error("maxpool_spatial: chunk too large"); return AF_ERR_INVALID; }
  if (is_bf16)
    maxpool_spatial_kernel<bf16><<<flat_grid(total, 256), 256, 0, s>>>((const bf16*)x, (bf16*)y, B * T, H, W, C, Ho, Wo);
  else
    maxpool_spatial_kernel<float><<<flat_grid(total, 256), 256, 0, s>>>((const float*)x, (float*)y, B * T, H, W, C, Ho, Wo);
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

int maxpool_temporal_launch(const void* x, void* y, int B, int T, int H, int W, int C, bool is_bf16,
                            cudaStream_t s) {
  const int V = is_bf16 ? 8 : 4;
  if (T % 2 || C % V) { set_error("maxpool_temporal: T=%d C=%d unsupported", T, C); return AF_ERR_INVALID; }
  const long long fv = (long long)H * W * C / V, BTo = (long long)B * (T / 2);
  if (is_bf16)
    maxpool_temporal_kernel<bf16><<<flat_grid(BTo * fv, 256), 256, 0, s>>>((const bf16*)x, (bf16*)y, BTo, fv);
  else
    maxpool_temporal_kernel<float><<<flat_grid(BTo * fv, 256), 256, 0, s>>>((const float*)x, (float*)y, BTo, fv);
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

int head_pool_slices(int P) { return P >= 8 * HEAD_PZ ? HEAD_PZ : 1; }

// features_ws: [B * head_pool_slices(P), C] floats
int head_launch(const void* x, int B, int P, int C, bool is_bf16, const float* fc_w, float fc_b,
                float* features_ws, float* features_out, float* logits, float* scores,
                cudaStream_t s) {
  const int V = is_bf16 ? 8 : 4;
  if (C % (V * 32)) { set_error("head: C=%d not a multiple of %d", C, V * 32); return AF_ERR_INVALID; }
  const int PZ = head_pool_slices(P);
  dim3 grid(C / (V * 32), B, PZ), block(32, 8);
  if (is_bf16) head_pool_kernel<bf16><<<grid, block, 0, s>>>((const bf16*)x, features_ws, P, C, PZ);
  else head_pool_kernel<float><<<grid, block, 0, s>>>((const float*)x, features_ws, P, C, PZ);
  head_fc_kernel<<<B, 256, 0, s>>>(features_ws, PZ, (float)P, fc_w, fc_b, C, features_out, logits, scores);
  g_launches += 2;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

int spatial_mean_launch(const void* x, int N, int P, int C, bool is_bf16, float* out, cudaStream_t s) {
  const int V = is_bf16 ? 8 : 4;
  if (C % (V * 32)) { set_error("spatial_mean: C=%d not a multiple of %d", C, V * 32); return AF_ERR_INVALID; }
  dim3 grid(C / (V * 32), N), block(32, 8);
  if (is_bf16) head_pool_kernel<bf16><<<grid, block, 0, s>>>((const bf16*)x, out, P, C, 1);
  else head_pool_kernel<float><<<grid, block, 0, s>>>((const float*)x, out, P, C, 1);
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

int ndhwc_to_ncthw_f32_launch(const void* x, float* y, int B, int T, int H, int W, int C,
                              bool is_bf16, cudaStream_t s) {
  const long long P = (long long)T * H * W;
  dim3 grid((unsigned)((P + 31) / 32), (C + 31) / 32, B), block(32, 8);
  if (is_bf16) ndhwc_to_ncthw_kernel<bf16><<<grid, block, 0, s>>>((const bf16*)x, y, B, P, C);
  else ndhwc_to_ncthw_kernel<float><<<grid, block, 0, s>>>((const float*)x, y, B, P, C);
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

}  // namespace afb
