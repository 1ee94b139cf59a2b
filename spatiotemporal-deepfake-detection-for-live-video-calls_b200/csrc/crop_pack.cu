// K1: crop / similarity-warp / normalise, and the clip packers.
//
// crop_kernel reproduces, bit for bit, what the reference does per frame on the CPU
//   canvas = zeros((h,w,3)); canvas[y:y+ih, x:x+iw] = crop; cv2.warpAffine(canvas, tfm, (S,S))
// (altfreezing/test_tools/faster_crop_align_xray.py:77-88) without materialising the canvas:
// canvas pixel (x,y) is frame pixel (x+left_top.x, y+left_top.y) masked to the frame's own
// big box.  The arithmetic is OpenCV's legacy fixed-point remap for u8 (un-vendored
// dependency; imgwarp.cpp WarpAffineInvoker + remapBilinear): the inverse map in f64 with
// round-half-even (cvRound), coordinates quantised to 1/32 px (AB_BITS 10, INTER_BITS 5),
// int16-saturated bilinear weights with 15 fractional bits, (sum + 2^14) >> 15.
// All f64/f32 products use explicit _rn intrinsics so nvcc cannot contract them into FMAs.
//
// With a ClipLayout destination the kernel also applies the callers' pack step
//   x = (float(u8) - 255*mean_c) / (255*std_c)      (altfreezing/demo.py:84-87,317-319)
// in IEEE fp32 and writes the engine's padded NDHWC4 clip directly (bf16 or fp32), so the
// aligned u8 clip never round-trips through HBM.
#include "common.cuh"
#include "../../include/afb200.h"

namespace afb {
namespace {

__device__ __forceinline__ int cv_round(double v) { return __double2int_rn(v); }

__device__ __forceinline__ int bilinear_w(int fa, int fb) {
  // float32((1 - a/32)) * float32((1 - b/32)) style entries of OpenCV's BilinearTab_i,
  // scaled by 2^15 and saturated to int16.
  const float v = __fmul_rn(fa * (1.0f / 32.0f), fb * (1.0f / 32.0f));
  const int w = __float2int_rn(__fmul_rn(v, 32768.0f));
  return w > 32767 ? 32767 : w;
}

template <typename T> __device__ __forceinline__ void store_px4(T* p, float a, float b, float c);
template <> __device__ __forceinline__ void store_px4<float>(float* p, float a, float b, float c) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, 0.f);
}
template <> __device__ __forceinline__ void store_px4<bf16>(bf16* p, float a, float b, float c) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, 0.f);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&lo);
  t.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = t;
}

struct Norm { float mean[3], stdv[3]; };

__device__ __forceinline__ float norm1(float v, float m, float s) { return __fdiv_rn(__fsub_rn(v, m), s); }

template <typename T, bool kToClip>
__global__ void __launch_bounds__(256) crop_kernel(const FrameDesc* __restrict__ frames,
                                                   const ClipGeom* __restrict__ geom, int T_, int S,
                                                   int bgr, uint8_t* __restrict__ out_u8, T* dst,
                                                   long long sB, long long sT, long long sH,
                                                   long long sW, Norm nrm) {
  // Per block (32 columns x 8 rows of one frame): the f64 part of OpenCV's coordinate maths is done
  // once per column (adelta, bdelta) and once per row (X0, Y0) by 40 threads and shared through smem.
  __shared__ int s_ad[32], s_bd[32], s_x0[8], s_y0[8];
  const int bt = blockIdx.z;
  const int b = bt / T_, t = bt - b * T_;
  const ClipGeom g = geom[b];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < 40) {
    // cv::invertAffineTransform (f64)
    double D = __dsub_rn(__dmul_rn(g.tfm[0], g.tfm[4]), __dmul_rn(g.tfm[1], g.tfm[3]));
    D = D != 0.0 ? __ddiv_rn(1.0, D) : 0.0;
    const double A11 = __dmul_rn(g.tfm[4], D), A22 = __dmul_rn(g.tfm[0], D);
    const double A12 = __dmul_rn(-g.tfm[1], D), A21 = __dmul_rn(-g.tfm[3], D);
    if (tid < 32) {
      const double xx = (double)(blockIdx.x * 32 + tid);
      s_ad[tid] = cv_round(__dmul_rn(__dmul_rn(A11, xx), 1024.0));
      s_bd[tid] = cv_round(__dmul_rn(__dmul_rn(A21, xx), 1024.0));
    } else {
      const double b1 = __dsub_rn(__dmul_rn(-A11, g.tfm[2]), __dmul_rn(A12, g.tfm[5]));
      const double b2 = __dsub_rn(__dmul_rn(-A21, g.tfm[2]), __dmul_rn(A22, g.tfm[5]));
      const double yy = (double)(blockIdx.y * 8 + tid - 32);
      s_x0[tid - 32] = cv_round(__dmul_rn(__dadd_rn(__dmul_rn(A12, yy), b1), 1024.0)) + 16;
      s_y0[tid - 32] = cv_round(__dmul_rn(__dadd_rn(__dmul_rn(A22, yy), b2), 1024.0)) + 16;
    }
  }
  __syncthreads();
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= S || y >= S) return;
  const FrameDesc f = frames[bt];
  const int adelta = s_ad[threadIdx.x], bdelta = s_bd[threadIdx.x];
  const int X0 = s_x0[threadIdx.y], Y0 = s_y0[threadIdx.y];
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  int sx = X >> 5, sy = Y >> 5;
  sx = max(-32768, min(32767, sx));
  sy = max(-32768, min(32767, sy));
  const int fx = X & 31, fy = Y & 31;
  const int w00 = bilinear_w(32 - fy, 32 - fx), w01 = bilinear_w(32 - fy, fx);
  const int w10 = bilinear_w(fy, 32 - fx), w11 = bilinear_w(fy, fx);

  const int bx1 = max(f.box[0], 0), by1 = max(f.box[1], 0);
  const int bx2 = min(f.box[2], f.width), by2 = min(f.box[3], f.height);
  int acc[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cx = sx + (k & 1), cy = sy + (k >> 1);
    const int wgt = k == 0 ? w00 : k == 1 ? w01 : k == 2 ? w10 : w11;
    const int px = cx + g.left_top[0], py = cy + g.left_top[1];
    const bool ok = cx >= 0 && cx < g.canvas_wh[0] && cy >= 0 && cy < g.canvas_wh[1] &&
                    px >= bx1 && px < bx2 && py >= by1 && py < by2;
    if (ok && wgt != 0) {
      const uint8_t* p = f.data + (long long)py * f.pitch + (long long)px * 3;
      acc[0] += wgt * (int)p[bgr ? 2 : 0];
      acc[1] += wgt * (int)p[1];
      acc[2] += wgt * (int)p[bgr ? 0 : 2];
    }
  }
  int o[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = max(0, min(255, (acc[c] + 16384) >> 15));

  if (kToClip) {
    T* q = dst + b * sB + t * sT + y * sH + x * sW;
    store_px4<T>(q, norm1((float)o[0], nrm.mean[0], nrm.stdv[0]), norm1((float)o[1], nrm.mean[1], nrm.stdv[1]),
                 norm1((float)o[2], nrm.mean[2], nrm.stdv[2]));
  } else {
    uint8_t* q = out_u8 + (((long long)bt * S + y) * S + x) * 3;
    q[0] = (uint8_t)o[0]; q[1] = (uint8_t)o[1]; q[2] = (uint8_t)o[2];
  }
}

// u8 [B,T,S,S,3] aligned clips -> normalised padded NDHWC4 clip (A4 of SURVEY.md §8a).
template <typename T>
__global__ void __launch_bounds__(256) pack_u8_kernel(const uint8_t* __restrict__ src, int T_, int S, T* dst,
                                                      long long sB, long long sT, long long sH, long long sW,
                                                      long long npix, Norm nrm) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int x = (int)(r % S); r /= S;
    const int y = (int)(r % S); r /= S;
    const int t = (int)(r % T_); r /= T_;
    const uint8_t* p = src + i * 3;
    store_px4<T>(dst + r * sB + t * sT + y * sH + x * sW, norm1((float)p[0], nrm.mean[0], nrm.stdv[0]),
                 norm1((float)p[1], nrm.mean[1], nrm.stdv[1]), norm1((float)p[2], nrm.mean[2], nrm.stdv[2]));
  }
}

template <typename TS> __device__ __forceinline__ float to_f(TS v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

// already-normalised clip tensor [B,3,T,S,S] with arbitrary element strides -> padded NDHWC4.
template <typename TS, typename T>
__global__ void __launch_bounds__(256) pack_clip_kernel(const TS* __restrict__ src, long long qB, long long qC,
                                                        long long qT, long long qH, long long qW, int T_, int S,
                                                        T* dst, long long sB, long long sT, long long sH,
                                                        long long sW, long long npix) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int x = (int)(r % S); r /= S;
    const int y = (int)(r % S); r /= S;
    const int t = (int)(r % T_); r /= T_;
    const TS* p = src + r * qB + t * qT + y * qH + x * qW;
    store_px4<T>(dst + r * sB + t * sT + y * sH + x * sW, to_f<TS>(p[0]), to_f<TS>(p[qC]), to_f<TS>(p[2 * qC]));
  }
}

// Stem unfold (bf16 engine only): U[b][t][r][xo][64] gathers, for output column xo and input
// row pair r (rows 2r-3, 2r-2), the 2 x 7 x 4 window values X[t][2r-3+dyy][2xo-3+dx][c] at
// k = dyy*28 + dx*4 + c (k >= 56 zero).  On U the reference's stem conv
// (k[5,7,7] s[1,2,2] p[2,3,3], altfreezing/slowfast/models/stem_helper.py:156-163) becomes a
// dense 64-channel conv with kernel [5,4,1], stride 1, pad [2,0,0] that the tcgen05 kernel takes.
// One thread writes one 16-byte group (two 4-channel pixels).
__global__ void __launch_bounds__(256) stem_unfold_kernel(const bf16* __restrict__ clip, long long sB, long long sT,
                                                          long long sH, long long sW, int T_, int Hu, int Wo,
                                                          bf16* __restrict__ U, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // 32-bit decode (total < 2^31 is checked by the launcher): 64-bit div/mod would dominate
    unsigned r_ = (unsigned)i;
    const int j = (int)(r_ & 7u); r_ >>= 3;
    const int xo = (int)(r_ % (unsigned)Wo); r_ /= (unsigned)Wo;
    const int r = (int)(r_ % (unsigned)Hu); r_ /= (unsigned)Hu;
    const int t = (int)(r_ % (unsigned)T_); r_ /= (unsigned)T_;
    const bf16* src = clip + r_ * sB + t * sT;
    uint2 px[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = 2 * j + h;                 // pixel slot 0..15
      px[h] = make_uint2(0u, 0u);
      if (q < 14) {
        const int dyy = q >= 7 ? 1 : 0, dx = q - 7 * dyy;
        px[h] = *reinterpret_cast<const uint2*>(src + (long long)(2 * r - 3 + dyy) * sH + (long long)(2 * xo - 3 + dx) * sW);
      }
    }
    *reinterpret_cast<uint4*>(U + i * 8) = make_uint4(px[0].x, px[0].y, px[1].x, px[1].y);
  }
}

inline int flat_grid(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

int crop_launch(const FrameDesc* frames, const ClipGeom* geom, int B, int T, int S, int bgr,
                uint8_t* out_u8, const ClipLayout* dst, const float mean[3], const float stdv[3],
                cudaStream_t s) {
  if (B <= 0) return AF_OK;
  dim3 grid((S + 31) / 32, (S + 7) / 8, B * T), block(32, 8);
  Norm n = {{0, 0, 0}, {1, 1, 1}};
  if (mean && stdv)
    for (int c = 0; c < 3; ++c) { n.mean[c] = mean[c]; n.stdv[c] = stdv[c]; }
  if (dst == nullptr) {
    crop_kernel<float, false><<<grid, block, 0, s>>>(frames, geom, T, S, bgr, out_u8, nullptr, 0, 0, 0, 0, n);
  } else if (dst->is_bf16) {
    crop_kernel<bf16, true><<<grid, block, 0, s>>>(frames, geom, T, S, bgr, nullptr, (bf16*)dst->base, dst->sB,
                                                   dst->sT, dst->sH, dst->sW, n);
  } else {
    crop_kernel<float, true><<<grid, block, 0, s>>>(frames, geom, T, S, bgr, nullptr, (float*)dst->base, dst->sB,
                                                    dst->sT, dst->sH, dst->sW, n);
  }
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

int stem_unfold_launch(const ClipLayout& clip, int clip0, int B, void* U, cudaStream_t s) {
  const int Wo = clip.S / 2, Hu = clip.S / 2 + 3;
  const long long total = (long long)B * clip.T * Hu * Wo * 8;
  if (total >= (1LL << 31)) { set_error("stem_unfold: chunk too large"); return AF_ERR_INVALID; }
  const bf16* base = (const bf16*)clip.base + (long long)clip0 * clip.sB;
  stem_unfold_kernel<<<flat_grid(total, 256), 256, 0, s>>>(base, clip.sB, clip.sT, clip.sH, clip.sW, clip.T, Hu, Wo,
                                                           (bf16*)U, total);
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

int pack_u8_launch(const uint8_t* src, int B, const float mean[3], const float stdv[3],
                   const ClipLayout& d, cudaStream_t s) {
  const long long npix = (long long)B * d.T * d.S * d.S;
  Norm n;
  for (int c = 0; c < 3; ++c) { n.mean[c] = mean[c]; n.stdv[c] = stdv[c]; }
  if (d.is_bf16)
    pack_u8_kernel<bf16><<<flat_grid(npix, 256), 256, 0, s>>>(src, d.T, d.S, (bf16*)d.base, d.sB, d.sT, d.sH, d.sW, npix, n);
  else
    pack_u8_kernel<float><<<flat_grid(npix, 256), 256, 0, s>>>(src, d.T, d.S, (float*)d.base, d.sB, d.sT, d.sH, d.sW, npix, n);
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

template <typename TS>
static int pack_clip_typed(const void* src, const long long q[5], int B, const ClipLayout& d, cudaStream_t s) {
  const long long npix = (long long)B * d.T * d.S * d.S;
  if (d.is_bf16)
    pack_clip_kernel<TS, bf16><<<flat_grid(npix, 256), 256, 0, s>>>((const TS*)src, q[0], q[1], q[2], q[3], q[4], d.T, d.S,
                                                                    (bf16*)d.base, d.sB, d.sT, d.sH, d.sW, npix);
  else
    pack_clip_kernel<TS, float><<<flat_grid(npix, 256), 256, 0, s>>>((const TS*)src, q[0], q[1], q[2], q[3], q[4], d.T, d.S,
                                                                     (float*)d.base, d.sB, d.sT, d.sH, d.sW, npix);
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

int pack_clip_launch(const void* src, int dtype, const long long strides[5], int B,
                     const ClipLayout& dst, cudaStream_t s) {
  switch (dtype) {
    case AF_F32: return pack_clip_typed<float>(src, strides, B, dst, s);
    case AF_BF16: return pack_clip_typed<bf16>(src, strides, B, dst, s);
    case AF_F16: return pack_clip_typed<__half>(src, strides, B, dst, s);
    default: set_error("af_forward: unsupported clip dtype %d", dtype); return AF_ERR_INVALID;
  }
}

}  // namespace afb
