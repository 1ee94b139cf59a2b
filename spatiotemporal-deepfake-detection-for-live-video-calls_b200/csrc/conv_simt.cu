// Generic channels-last Conv3d (+bias +residual +ReLU) on CUDA cores, fp32 accumulate.
//
// This is the arithmetic of nn.Conv3d + eval BatchNorm3d (+ add + ReLU) as the reference
// chains them (altfreezing/slowfast/models/resnet_helper.py:311-326,438-444;
// stem_helper.py:173-178), written as an implicit GEMM: M = B*To*Ho*Wo output positions,
// N = Cout, K = taps*Cin.  It is the fp32 ("TF32-free") parity path of the engine and the
// on-device cross-check for the tcgen05 kernel; the bf16 trunk only uses it for shapes the
// tensor-core kernel does not take.
//
// Tiling: 64x64 outputs per 256-thread CTA, 4x4 per thread, K in slabs of 16 staged in
// shared memory.  Requires Cin % 4 == 0 and Cout % 64 == 0 (the stem is packed to Cin=4).
#include "common.cuh"
#include "../../include/afb200.h"

namespace afb {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename T> struct Ld4;
template <> struct Ld4<float> {
  static __device__ __forceinline__ void load(const float* p, float v[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float v[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Ld4<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float v[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void store(bf16* p, const float v[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

template <typename T>
__global__ void __launch_bounds__(NT) conv_simt_kernel(ConvProblem p) {
  __shared__ float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];

  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int K = p.kt * p.kh * p.kw * p.Cin;

  // A-load role: row ar (0..63), k-quad aq (0..3)
  const int ar = tid >> 2, aq = (tid & 3) * 4;
  const long long am = m0 + ar;
  const bool arow_ok = am < p.M;
  int ab = 0, at = 0, ah = 0, aw = 0;
  if (arow_ok) {
    long long r = am;
    aw = (int)(r % p.Wo); r /= p.Wo;
    ah = (int)(r % p.Ho); r /= p.Ho;
    at = (int)(r % p.To); r /= p.To;
    ab = (int)r;
  }
  const int it0 = at * p.st - p.pt, ih0 = ah * p.sh - p.ph, iw0 = aw * p.sw - p.pw;
  const T* xb = reinterpret_cast<const T*>(p.x) + (long long)ab * p.xsB;

  // B-load role: k row bk (0..15), 4 columns at bn
  const int bk = tid >> 4, bn = (tid & 15) * 4;
  const float* wp = reinterpret_cast<const float*>(p.w);

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    {  // A slab
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int k = k0 + aq;
      if (arow_ok && k < K) {
        const int tap = k / p.Cin, c = k - tap * p.Cin;
        const int dx = tap % p.kw, r = tap / p.kw, dy = r % p.kh, dt = r / p.kh;
        const int it = it0 + dt, ih = ih0 + dy, iw = iw0 + dx;
        if ((unsigned)it < (unsigned)p.Ti && (unsigned)ih < (unsigned)p.Hi && (unsigned)iw < (unsigned)p.Wi)
          Ld4<T>::load(xb + it * p.xsT + ih * p.xsH + iw * p.xsW + c, v);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) As[aq + j][ar] = v[j];
    }
    {  // B slab
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int k = k0 + bk;
      if (k < K) v = *reinterpret_cast<const float4*>(wp + (long long)k * p.Cout + n0 + bn);
      *reinterpret_cast<float4*>(&Bs[bk][bn]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int n = n0 + tx * 4;
  const float4 bias = *reinterpret_cast<const float4*>(p.bias + n);
  const float bv[4] = {bias.x, bias.y, bias.z, bias.w};
  T* y = reinterpret_cast<T*>(p.y);
  const T* res = reinterpret_cast<const T*>(p.res);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = acc[i][j] + bv[j];
    if (res) {
      float r[4];
      Ld4<T>::load(res + m * p.Cout + n, r);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] += r[j];
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    Ld4<T>::store(y + m * p.Cout + n, o);
  }
}

}  // namespace

int conv_simt_launch(const ConvProblem& p, bool is_bf16, cudaStream_t s) {
  if (p.Cin % 4 != 0 || p.Cout % BN != 0) {
    set_error("conv_simt: needs Cin %% 4 == 0 and Cout %% 64 == 0 (got %d, %d)", p.Cin, p.Cout);
    return AF_ERR_INVALID;
  }
  dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)(p.Cout / BN));
  if (is_bf16) conv_simt_kernel<bf16><<<grid, NT, 0, s>>>(p);
  else conv_simt_kernel<float><<<grid, NT, 0, s>>>(p);
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

}  // namespace afb
