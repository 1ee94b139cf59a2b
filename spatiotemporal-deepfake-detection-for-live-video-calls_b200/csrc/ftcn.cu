// Kernels that only the FTCN-TT plugin needs (altfreezing/model/classifier/i3d_temporal_var_fix_dropout_tt_cfg.py):
// the same slowfast ResNet-50 trunk with every spatial kernel collapsed to 1x1 and every spatial stride replaced by
// a MaxPool3d((1,2,2)) behind the conv's BatchNorm (temporal_only_conv, :207-289), stopped after s4 (:315-321), and a
// one-layer pre-norm transformer over the 16 per-frame mean features (TransformerHead :126-196,
// time_transformer.py:29-88,219-279) instead of the average-pool + Linear head.
//
//   ftcn_stem_kernel     Conv3d(3->64, k[5,1,1]) + folded BN + MaxPool3d(1,2,2) + ReLU + MaxPool3d k[1,3,3] s[1,2,2]
//                        p[0,1,1] in one pass over the padded NDHWC4 clip (the 224x224x64 conv output never exists)
//   maxpool_hw2_kernel   MaxPool3d((1,2,2)) on NDHWC
//   tt_* kernels         the transformer head in fp32 (17 tokens x 1024 channels per clip)
// The trunk's kt x 1 x 1 convs run on the engine's ordinary conv kernels (conv_umma / conv_tsweep / conv_simt).
#include <math.h>

#include "../../include/afb200.h"
#include "common.cuh"

namespace afb {
namespace {

// ------------------------------------------------------------------------------------------------ stem
constexpr int FS_OX = 8, FS_OY = 4;                    // final (56x56-level) outputs per block
constexpr int FS_MX = 2 * FS_OX + 1, FS_MY = 2 * FS_OY + 1;   // 112-level map region incl. the 3x3/2 pool halo: 17 x 9
constexpr int FS_IX = 2 * FS_MX, FS_IY = 2 * FS_MY;    // input pixels under it: 34 x 18
constexpr int FS_C = 64;
template <typename T> constexpr int fs_smem_bytes() { return 5 * FS_IY * FS_IX * 16 + FS_MY * FS_MX * FS_C * 4 + 5 * FS_IY * FS_IX * 4 * (int)sizeof(T); }

template <typename T> __device__ __forceinline__ float4 load_px4(const T* p);
template <> __device__ __forceinline__ float4 load_px4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 load_px4<bf16>(const bf16* p) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&t.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
  return make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void store4<bf16>(bf16* p, float4 v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 t;
  t.x = *reinterpret_cast<const uint32_t*>(&lo);
  t.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = t;
}

__device__ __forceinline__ void cp_async_px(void* smem_dst, const void* gsrc, int bytes, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int src_bytes = valid ? bytes : 0;                  // 0 source bytes = zero fill
  if (bytes == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

// clip: padded NDHWC4 (element strides sB,sT,sH,sW; `clip` points at logical (b=0,t=0,y=0,x=0); >= 2 zero frames
// and >= 3 zero rows / columns around every clip).  w: [5][4][64] fp32 (tap, channel, cout).  y: [B*T, S/4, S/4, 64].
// Persistent blocks (2 per SM) walk tiles of 8 x 4 final outputs of one frame.  Per tile: the 5-frame input patch
// arrives by cp.async into a raw staging buffer WHILE the previous tile is computed, is widened to fp32 once, then
// thread = 4 consecutive output channels x one of 16 position lanes (every 16-byte smem read feeds 12 FMAs) builds
// the 2x2-max-pooled, ReLU'd 17 x 9 map tile in smem and the 3x3/2 max-pool writes the 32 outputs.
// (Round-1 history: one channel per thread 13 TFLOP/s, bound by the smem reads; 4 channels per thread with a
// synchronous fill 22 TFLOP/s, bound by the fill phase at 2 blocks/SM; reading pixels straight from global 15.)
template <typename T>
__global__ void __launch_bounds__(256, 2) ftcn_stem_kernel(const T* __restrict__ clip, long long sB, long long sT, long long sH,
                                                           long long sW, int T_, int S, const float* __restrict__ w,
                                                           const float* __restrict__ bias, T* __restrict__ y, int tiles_x,
                                                           int tiles_y, int n_tiles) {
  extern __shared__ float4 fs_smem[];
  float4* in_s = fs_smem;                                                   // [5][FS_IY][FS_IX] fp32 pixels
  float* map_s = reinterpret_cast<float*>(fs_smem + 5 * FS_IY * FS_IX);     // [FS_MY*FS_MX][64]
  T* raw_s = reinterpret_cast<T*>(map_s + FS_MY * FS_MX * FS_C);            // [5][FS_IY][FS_IX][4] as stored in the clip
  const int cg = (threadIdx.x & 15) * 4, pl = threadIdx.x >> 4;
  float4 wr[5][3];
#pragma unroll
  for (int f = 0; f < 5; ++f)
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) wr[f][ch] = __ldg(reinterpret_cast<const float4*>(w + (f * 4 + ch) * FS_C + cg));
  const float4 bc = __ldg(reinterpret_cast<const float4*>(bias + cg));
  const int M2 = S / 2, O = S / 4;

  auto issue = [&](int tile) {                             // start the patch of `tile` on its way into raw_s
    const int tx = tile % tiles_x, r0 = tile / tiles_x, ty = r0 % tiles_y, bt = r0 / tiles_y;
    const int b = bt / T_, t = bt - b * T_;
    const int ix0 = 2 * (2 * tx * FS_OX - 1), iy0 = 2 * (2 * ty * FS_OY - 1);   // may be -2: inside the pads
    const T* src = clip + b * sB + (long long)(t - 2) * sT;
    for (int i = threadIdx.x; i < 5 * FS_IY * FS_IX; i += 256) {
      const int x = i % FS_IX, r = i / FS_IX, yy = r % FS_IY, f = r / FS_IY;
      const int gx = ix0 + x, gy = iy0 + yy;
      const bool valid = gx < S && gy < S;
      cp_async_px(raw_s + (size_t)i * 4, src + f * sT + (long long)(valid ? gy : 0) * sH + (long long)(valid ? gx : 0) * sW,
                  (int)(4 * sizeof(T)), valid);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int tile = blockIdx.x;
  if (tile < n_tiles) issue(tile);
  for (; tile < n_tiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, r0 = tile / tiles_x, ty = r0 % tiles_y, bt = r0 / tiles_y;
    const int ox0 = tx * FS_OX, oy0 = ty * FS_OY;
    const int mx0 = 2 * ox0 - 1, my0 = 2 * oy0 - 1;                         // map origin (may be -1)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    for (int i = threadIdx.x; i < 5 * FS_IY * FS_IX; i += 256) in_s[i] = load_px4<T>(raw_s + (size_t)i * 4);   // own copies
    __syncthreads();                                                        // in_s complete, raw_s free
    if (tile + (int)gridDim.x < n_tiles) issue(tile + gridDim.x);
    for (int pos = pl; pos < FS_MY * FS_MX; pos += 16) {
      const int my = pos / FS_MX, mx = pos - my * FS_MX;
      float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
      for (int q = 0; q < 4; ++q) {                      // MaxPool3d((1,2,2)) over the conv + BN outputs
        const int px = (2 * my + (q >> 1)) * FS_IX + 2 * mx + (q & 1);
        float4 a = bc;
#pragma unroll
        for (int f = 0; f < 5; ++f) {
          const float4 v = in_s[f * FS_IY * FS_IX + px];
          a.x = fmaf(v.x, wr[f][0].x, a.x); a.y = fmaf(v.x, wr[f][0].y, a.y); a.z = fmaf(v.x, wr[f][0].z, a.z); a.w = fmaf(v.x, wr[f][0].w, a.w);
          a.x = fmaf(v.y, wr[f][1].x, a.x); a.y = fmaf(v.y, wr[f][1].y, a.y); a.z = fmaf(v.y, wr[f][1].z, a.z); a.w = fmaf(v.y, wr[f][1].w, a.w);
          a.x = fmaf(v.z, wr[f][2].x, a.x); a.y = fmaf(v.z, wr[f][2].y, a.y); a.z = fmaf(v.z, wr[f][2].z, a.z); a.w = fmaf(v.z, wr[f][2].w, a.w);
        }
        m.x = fmaxf(m.x, a.x); m.y = fmaxf(m.y, a.y); m.z = fmaxf(m.z, a.z); m.w = fmaxf(m.w, a.w);
      }
      // ReLU, then positions outside the 112x112 map must not win the 3x3 max: after ReLU 0 is neutral
      const bool inside = (unsigned)(my0 + my) < (unsigned)M2 && (unsigned)(mx0 + mx) < (unsigned)M2;
      if (!inside) m = make_float4(0.f, 0.f, 0.f, 0.f);
      m.x = fmaxf(m.x, 0.f); m.y = fmaxf(m.y, 0.f); m.z = fmaxf(m.z, 0.f); m.w = fmaxf(m.w, 0.f);
      *reinterpret_cast<float4*>(map_s + pos * FS_C + cg) = m;
    }
    __syncthreads();                                                        // map complete; in_s may be overwritten
#pragma unroll
    for (int j = 0; j < (FS_OX * FS_OY) / 16; ++j) {
      const int o = pl * ((FS_OX * FS_OY) / 16) + j, oy = o / FS_OX, ox = o - oy * FS_OX;
      if (oy0 + oy >= O || ox0 + ox >= O) continue;
      float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float4 v = *reinterpret_cast<const float4*>(map_s + ((2 * oy + dy) * FS_MX + 2 * ox + dx) * FS_C + cg);
          m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
      store4<T>(y + (((long long)bt * O + oy0 + oy) * O + ox0 + ox) * FS_C + cg, m);
    }
    // the next iteration's two barriers (after the widening pass) separate this pool phase from the next map writes
  }
}

// ------------------------------------------------------------------------------------------------ 2x2 max-pool
template <typename T> struct V16;
template <> struct V16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ uint4 vmax(uint4 a, uint4 b) {
    float4 x = *reinterpret_cast<float4*>(&a), z = *reinterpret_cast<float4*>(&b);
    x.x = fmaxf(x.x, z.x); x.y = fmaxf(x.y, z.y); x.z = fmaxf(x.z, z.z); x.w = fmaxf(x.w, z.w);
    return *reinterpret_cast<uint4*>(&x);
  }
};
template <> struct V16<bf16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ uint4 vmax(uint4 a, uint4 b) {
    __nv_bfloat162* x = reinterpret_cast<__nv_bfloat162*>(&a);
    const __nv_bfloat162* z = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = __hmax2(x[i], z[i]);
    return a;
  }
};

template <typename T>
__global__ void __launch_bounds__(256) maxpool_hw2_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C,
                                                          long long total) {
  constexpr int V = V16<T>::N;
  const int cv = C / V, Ho = H / 2, Wo = W / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int c = (int)(r % cv) * V; r /= cv;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho); r /= Ho;
    const T* p = x + ((r * H + 2 * ho) * W + 2 * wo) * C + c;
    uint4 m = *reinterpret_cast<const uint4*>(p);
    m = V16<T>::vmax(m, *reinterpret_cast<const uint4*>(p + C));
    m = V16<T>::vmax(m, *reinterpret_cast<const uint4*>(p + (long long)W * C));
    m = V16<T>::vmax(m, *reinterpret_cast<const uint4*>(p + (long long)W * C + C));
    *reinterpret_cast<uint4*>(y + ((r * Ho + ho) * Wo + wo) * C + c) = m;
  }
}

// ------------------------------------------------------------------------------------------------ transformer head
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                                    // red may still be read from a previous call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// x[b, 0, :] = cls + pos[0];  x[b, 1+i, :] = tok[b, i, :] + pos[1+i]          (TimeTransformer.forward :262-266)
__global__ void tt_embed_kernel(const float* __restrict__ tok, const float* __restrict__ cls, const float* __restrict__ pos,
                                float* __restrict__ x, int n, int D) {
  const int row = blockIdx.x, b = row / (n + 1), i = row - b * (n + 1);
  const float* src = i == 0 ? cls : tok + ((long long)b * n + (i - 1)) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) x[(long long)row * D + d] = src[d] + pos[(long long)i * D + d];
}

// nn.LayerNorm(D), eps 1e-5, biased variance; one block per row (rows `stride_rows` apart in the input)
__global__ void tt_layernorm_kernel(const float* __restrict__ x, long long in_row_stride, const float* __restrict__ g,
                                    const float* __restrict__ be, float* __restrict__ y, int D) {
  __shared__ float red[32];
  const float* xr = x + (long long)blockIdx.x * in_row_stride;
  float s = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) s += xr[d];
  const float mean = block_sum(s, red) / (float)D;
  float q = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) { const float t = xr[d] - mean; q = fmaf(t, t, q); }
  const float rstd = rsqrtf(block_sum(q, red) / (float)D + 1e-5f);
  for (int d = threadIdx.x; d < D; d += blockDim.x) y[(long long)blockIdx.x * D + d] = (xr[d] - mean) * rstd * g[d] + be[d];
}

// Y[M,N] = X[M,K] . W[N,K]^T (+ bias) (-> exact GELU) (+ R[M,N]).  64 x 64 output tile per block (4 x 4 per thread),
// K in steps of 16.
constexpr int TL_BM = 64, TL_BN = 64, TL_BK = 16;
__global__ void __launch_bounds__(256) tt_linear_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                        const float* __restrict__ bias, const float* __restrict__ R,
                                                        float* __restrict__ Y, int M, int N, int K, int gelu) {
  __shared__ __align__(16) float xs[TL_BK][TL_BM + 4], ws[TL_BK][TL_BN + 4];
  const int m0 = blockIdx.y * TL_BM, n0 = blockIdx.x * TL_BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // 16 x 16 threads; each 4 rows x 4 columns
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TL_BK) {
    for (int i = threadIdx.x; i < TL_BM * TL_BK / 4; i += 256) {      // 16-byte loads along K, transposed into smem
      const int k4 = (i & 3) * 4, r = i >> 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M) v = *reinterpret_cast<const float4*>(X + (long long)(m0 + r) * K + k0 + k4);
      xs[k4][r] = v.x; xs[k4 + 1][r] = v.y; xs[k4 + 2][r] = v.z; xs[k4 + 3][r] = v.w;
    }
    for (int i = threadIdx.x; i < TL_BN * TL_BK / 4; i += 256) {
      const int k4 = (i & 3) * 4, r = i >> 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + r < N) v = *reinterpret_cast<const float4*>(W + (long long)(n0 + r) * K + k0 + k4);
      ws[k4][r] = v.x; ws[k4 + 1][r] = v.y; ws[k4 + 2][r] = v.z; ws[k4 + 3][r] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TL_BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&xs[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&ws[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bw[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (gelu) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));     // nn.GELU(approximate='none')
      if (R) v += R[(long long)m * N + n];
      Y[(long long)m * N + n] = v;
    }
  }
}

// softmax(q k^T * scale) v for one (clip, head): qkv [B, n1, 3*H*dh] in to_qkv's (q | k | v), (h d) order
// (time_transformer.py:49-64) -> out [B, n1, H*dh].  n1 <= 32, dh = 64.
constexpr int TA_MAXN = 32, TA_DH = 64;
__global__ void __launch_bounds__(128) tt_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int n1,
                                                           int heads, float scale) {
  __shared__ float q[TA_MAXN][TA_DH + 1], k[TA_MAXN][TA_DH + 1], v[TA_MAXN][TA_DH + 1], p[TA_MAXN][TA_MAXN + 1];
  const int b = blockIdx.x / heads, h = blockIdx.x - b * heads;
  const int inner = heads * TA_DH;
  for (int i = threadIdx.x; i < n1 * TA_DH; i += 128) {
    const int r = i / TA_DH, d = i - r * TA_DH;
    const float* row = qkv + ((long long)b * n1 + r) * 3 * inner + h * TA_DH + d;
    q[r][d] = row[0]; k[r][d] = row[inner]; v[r][d] = row[2 * inner];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n1 * n1; i += 128) {
    const int r = i / n1, c = i - r * n1;
    float s = 0.f;
#pragma unroll 16
    for (int d = 0; d < TA_DH; ++d) s = fmaf(q[r][d], k[c][d], s);
    p[r][c] = s * scale;
  }
  __syncthreads();
  if (threadIdx.x < n1) {
    const int r = threadIdx.x;
    float mx = -INFINITY;
    for (int c = 0; c < n1; ++c) mx = fmaxf(mx, p[r][c]);
    float sum = 0.f;
    for (int c = 0; c < n1; ++c) { const float e = expf(p[r][c] - mx); p[r][c] = e; sum += e; }
    const float inv = 1.f / sum;
    for (int c = 0; c < n1; ++c) p[r][c] *= inv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n1 * TA_DH; i += 128) {
    const int r = i / TA_DH, d = i - r * TA_DH;
    float s = 0.f;
    for (int c = 0; c < n1; ++c) s = fmaf(p[r][c], v[c][d], s);
    out[((long long)b * n1 + r) * inner + h * TA_DH + d] = s;
  }
}

// mlp_head on the cls token: LayerNorm -> Linear(D -> 1) (time_transformer.py:247,270-273); optional sigmoid and a
// copy of the normalised cls vector (the input of the last nn.Linear, which altfreezing/feature.py:106-114 hooks).
__global__ void tt_final_kernel(const float* __restrict__ x, long long row_stride, const float* __restrict__ g,
                                const float* __restrict__ be, const float* __restrict__ w, float bias, int D,
                                float* __restrict__ feat_out, float* __restrict__ logits, float* __restrict__ scores) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float* xr = x + (long long)b * row_stride;
  float s = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) s += xr[d];
  const float mean = block_sum(s, red) / (float)D;
  float q = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) { const float t = xr[d] - mean; q = fmaf(t, t, q); }
  const float rstd = rsqrtf(block_sum(q, red) / (float)D + 1e-5f);
  float dot = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float nv = (xr[d] - mean) * rstd * g[d] + be[d];
    if (feat_out) feat_out[(long long)b * D + d] = nv;
    dot = fmaf(nv, w[d], dot);
  }
  const float l = block_sum(dot, red) + bias;
  if (threadIdx.x == 0) {
    if (logits) logits[b] = l;
    if (scores) scores[b] = 1.f / (1.f + expf(-l));
  }
}

inline int flat_grid(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

int ftcn_stem_launch(const ClipLayout& clip, int clip0, int B, const float* w_tap_c_cout, const float* bias, void* y,
                     cudaStream_t s) {
  if (clip.S % 4 || B <= 0) { set_error("ftcn_stem: clip size %d not a multiple of 4", clip.S); return AF_ERR_INVALID; }
  static bool configured[64] = {};
  static int sms[64] = {};
  int dev = 0;
  AFB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("ftcn_stem: device index %d out of range", dev); return AF_ERR_INVALID; }
  if (!configured[dev]) {
    AFB_CUDA(cudaFuncSetAttribute(ftcn_stem_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs_smem_bytes<bf16>()));
    AFB_CUDA(cudaFuncSetAttribute(ftcn_stem_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs_smem_bytes<float>()));
    AFB_CUDA(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev));
    configured[dev] = true;
  }
  const int O = clip.S / 4;
  const int tiles_x = (O + FS_OX - 1) / FS_OX, tiles_y = (O + FS_OY - 1) / FS_OY;
  const long long n_tiles_ll = (long long)tiles_x * tiles_y * B * clip.T;
  if (n_tiles_ll >= (1LL << 31)) { set_error("ftcn_stem: chunk too large"); return AF_ERR_INVALID; }
  const int n_tiles = (int)n_tiles_ll;
  if (clip.is_bf16) {
    const int grid = n_tiles < 2 * sms[dev] ? n_tiles : 2 * sms[dev];            // 2 resident blocks per SM
    ftcn_stem_kernel<bf16><<<grid, 256, fs_smem_bytes<bf16>(), s>>>((const bf16*)clip.base + (long long)clip0 * clip.sB, clip.sB,
                                                                    clip.sT, clip.sH, clip.sW, clip.T, clip.S, w_tap_c_cout, bias,
                                                                    (bf16*)y, tiles_x, tiles_y, n_tiles);
  } else {
    const int grid = n_tiles < sms[dev] ? n_tiles : sms[dev];                    // fp32 staging: 1 block per SM
    ftcn_stem_kernel<float><<<grid, 256, fs_smem_bytes<float>(), s>>>((const float*)clip.base + (long long)clip0 * clip.sB, clip.sB,
                                                                      clip.sT, clip.sH, clip.sW, clip.T, clip.S, w_tap_c_cout,
                                                                      bias, (float*)y, tiles_x, tiles_y, n_tiles);
  }
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

int maxpool_hw2_launch(const void* x, void* y, long long BT, int H, int W, int C, bool is_bf16, cudaStream_t s) {
  const int V = is_bf16 ? 8 : 4;
  if ((H & 1) || (W & 1) || C % V) { set_error("maxpool_hw2: H=%d W=%d C=%d unsupported", H, W, C); return AF_ERR_INVALID; }
  const long long total = BT * (H / 2) * (W / 2) * (C / V);
  if (is_bf16) maxpool_hw2_kernel<bf16><<<flat_grid(total, 256), 256, 0, s>>>((const bf16*)x, (bf16*)y, H, W, C, total);
  else maxpool_hw2_kernel<float><<<flat_grid(total, 256), 256, 0, s>>>((const float*)x, (float*)y, H, W, C, total);
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

int tt_head_launch(const TTHeadDev& h, const float* tokens, int B, float* ws, float* features_out, float* logits,
                   float* scores, cudaStream_t s) {
  const int n1 = h.tokens + 1, D = h.dim, inner = h.heads * h.dim_head, M = B * n1;
  if (n1 > TA_MAXN || h.dim_head != TA_DH) { set_error("tt_head: %d tokens / dim_head %d unsupported", h.tokens, h.dim_head); return AF_ERR_INVALID; }
  // workspace carve-up (floats): x [M,D] | ln [M,D] | qkv [M,3*inner] | att [M,inner] | hid [M,mlp]
  float* x = ws;
  float* ln = x + (long long)M * D;
  float* qkv = ln + (long long)M * D;
  float* att = qkv + (long long)M * 3 * inner;
  float* hid = att + (long long)M * inner;
  auto linear = [&](const float* X, const float* W, const float* bias, const float* R, float* Y, int N, int K, int gelu) {
    dim3 grid((N + TL_BN - 1) / TL_BN, (M + TL_BM - 1) / TL_BM);
    tt_linear_kernel<<<grid, 256, 0, s>>>(X, W, bias, R, Y, M, N, K, gelu);
    ++g_launches;
  };
  tt_embed_kernel<<<M, 256, 0, s>>>(tokens, h.cls_token, h.pos_embedding, x, h.tokens, D);
  ++g_launches;
  for (const TTLayerDev& L : h.layers) {
    tt_layernorm_kernel<<<M, 256, 0, s>>>(x, D, L.ln1_w, L.ln1_b, ln, D);
    linear(ln, L.qkv_w, nullptr, nullptr, qkv, 3 * inner, D, 0);
    tt_attention_kernel<<<B * h.heads, 128, 0, s>>>(qkv, att, n1, h.heads, 1.0f / sqrtf((float)h.dim_head));
    linear(att, L.out_w, L.out_b, x, x, D, inner, 0);                 // x = to_out(att) + x   (Residual, :8-12)
    tt_layernorm_kernel<<<M, 256, 0, s>>>(x, D, L.ln2_w, L.ln2_b, ln, D);
    linear(ln, L.fc1_w, L.fc1_b, nullptr, hid, h.mlp_dim, D, 1);
    linear(hid, L.fc2_w, L.fc2_b, x, x, D, h.mlp_dim, 0);             // x = ff(ln) + x
    g_launches += 3;
  }
  tt_final_kernel<<<B, 256, 0, s>>>(x, (long long)n1 * D, h.norm_w, h.norm_b, h.fc_w, h.fc_b, D, features_out, logits, scores);
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

long long tt_head_workspace_floats(const TTHeadDev& h, int B) {
  const long long M = (long long)B * (h.tokens + 1), inner = (long long)h.heads * h.dim_head;
  return M * (2LL * h.dim + 4 * inner + h.mlp_dim);
}

}  // namespace afb
