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
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "../../include/afb200.h"

namespace afb {
namespace {

__device__ __forceinline__ int cv_round(double v) { return __double2int_rn(v); }

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

template <typename T> __device__ __forceinline__ void store_1(T* p, float v);
template <> __device__ __forceinline__ void store_1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store_1<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

struct Norm { float mean[3], stdv[3]; };

__device__ __forceinline__ float norm1(float v, float m, float s) { return __fdiv_rn(__fsub_rn(v, m), s); }

// 6 consecutive bytes (two RGB pixels) starting at an arbitrary address, fetched as two aligned 8-byte words.
// Only used where the 16-byte window is known to lie inside the frame buffer.
__device__ __forceinline__ uint2 load6(const uint8_t* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint2* q = reinterpret_cast<const uint2*>(a & ~(uintptr_t)7);
  const uint2 lo = __ldg(q), hi = __ldg(q + 1);
  const uint32_t sh = (uint32_t)(a & 7) * 8;
  // 128-bit window {lo.x, lo.y, hi.x, hi.y} shifted right by sh bits (sh in {0,8,..,56}); keep the low 48 bits
  const uint32_t w0 = sh < 32 ? lo.x : lo.y, w1 = sh < 32 ? lo.y : hi.x, w2 = sh < 32 ? hi.x : hi.y;
  const uint32_t s2 = sh & 31;
  return make_uint2(__funnelshift_r(w0, w1, s2), __funnelshift_r(w1, w2, s2));
}

// Per block (32 columns x 32 rows of one frame, 4 rows per thread): the f64 part of OpenCV's coordinate maths is done once per column
// (adelta, bdelta) and once per row (X0, Y0) by 40 threads and shared through smem, together with the frame's
// descriptor reduced to "valid canvas window" form.  lut: fp32 [3][256] = (v - mean_c) / std_c for every u8 value
// (built by norm_lut_kernel with the same IEEE division the callers' pack step performs).
//
// Pixel path.  The four taps of an output pixel are the 2x2 source pixels at (sx, sy); a tap outside the frame's
// valid window contributes zero (cv2's BORDER_CONSTANT and the canvas' black surround).  The FAST path (block-uniform:
// the window is at least 2x2 and does not touch the frame's first or last row, so 16-byte fetch windows stay inside the
// frame buffer) is branch-free: the 2x2 fetch position is clamped INTO the window, the weights of taps outside the
// window are zeroed, and when clamping moved the position by one pixel / row the fetched data is shifted back under
// the taps (a 24-bit funnel shift / a row select).  Pixels whose taps are all outside get four zero weights.  One code
// path for interior, edge and outside pixels means no warp divergence along the rotated crop's borders.
constexpr int CROP_ROWS = 112;     // output rows per block (8 thread rows x 14): amortises the per-block prologue (fp64 geometry, descriptor, LUT fill: a quarter of the time with 32-row blocks; measured 0.351 -> 0.334 (56 rows) -> 0.300 ms (112 rows) per 32 clips)
template <typename T, bool kToClip>
__global__ void __launch_bounds__(256) crop_kernel(const FrameDesc* __restrict__ frames,
                                                   const ClipGeom* __restrict__ geom, int T_, int S,
                                                   int bgr, uint8_t* __restrict__ out_u8, T* dst,
                                                   long long sB, long long sT, long long sH,
                                                   long long sW, long long sC, const float* __restrict__ lut) {
  // block-uniform values are packed so that a thread fetches them with a few vector LDS (the kernel is issue-bound)
  __shared__ int2 s_col[32], s_row[CROP_ROWS];   // (adelta, bdelta) per column, (X0, Y0) per row
  __shared__ int4 s_win;                       // valid canvas window x_lo, x_hi, y_lo, y_hi
  __shared__ int s_fast;                       // fast path usable for this frame
  struct __align__(16) Org { const uint8_t* p; long long pitch; };
  __shared__ Org s_org;                        // address of canvas pixel (0,0) in the frame, row pitch
  __shared__ float s_lut[768];
  const int bt = blockIdx.z;
  const int b = bt / T_, t = bt - b * T_;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (kToClip)
    for (int i = tid; i < 768; i += 256) s_lut[i] = __ldg(lut + i);
  if (tid < 32 + CROP_ROWS) {
    const ClipGeom g = geom[b];
    // cv::invertAffineTransform (f64)
    double D = __dsub_rn(__dmul_rn(g.tfm[0], g.tfm[4]), __dmul_rn(g.tfm[1], g.tfm[3]));
    D = D != 0.0 ? __ddiv_rn(1.0, D) : 0.0;
    const double A11 = __dmul_rn(g.tfm[4], D), A22 = __dmul_rn(g.tfm[0], D);
    const double A12 = __dmul_rn(-g.tfm[1], D), A21 = __dmul_rn(-g.tfm[3], D);
    if (tid < 32) {
      const double xx = (double)(blockIdx.x * 32 + tid);
      s_col[tid] = make_int2(cv_round(__dmul_rn(__dmul_rn(A11, xx), 1024.0)), cv_round(__dmul_rn(__dmul_rn(A21, xx), 1024.0)));
    } else {
      const double b1 = __dsub_rn(__dmul_rn(-A11, g.tfm[2]), __dmul_rn(A12, g.tfm[5]));
      const double b2 = __dsub_rn(__dmul_rn(-A21, g.tfm[2]), __dmul_rn(A22, g.tfm[5]));
      const double yy = (double)(blockIdx.y * CROP_ROWS + tid - 32);
      s_row[tid - 32] = make_int2(cv_round(__dmul_rn(__dadd_rn(__dmul_rn(A12, yy), b1), 1024.0)) + 16,
                                  cv_round(__dmul_rn(__dadd_rn(__dmul_rn(A22, yy), b2), 1024.0)) + 16);
    }
  } else if (tid == 32 + CROP_ROWS) {
    // canvas pixel (cx,cy) is frame pixel (cx+ltx, cy+lty); it exists iff it is inside the canvas AND inside this
    // frame's own big box clipped to the frame (faster_crop_align_xray.py:77-83)
    const ClipGeom g = geom[b];
    const FrameDesc f = frames[bt];
    const int ltx = g.left_top[0], lty = g.left_top[1];
    const int bx1 = max(f.box[0], 0), by1 = max(f.box[1], 0);
    const int bx2 = min(f.box[2], f.width), by2 = min(f.box[3], f.height);
    const int4 win = make_int4(max(0, bx1 - ltx), min(g.canvas_wh[0], bx2 - ltx), max(0, by1 - lty), min(g.canvas_wh[1], by2 - lty));
    s_win = win;
    // fast path: window >= 2x2, inside frame rows 1 .. height-2 (the aligned 16-byte fetches of load6 then stay
    // inside the frame buffer) and 32-bit byte offsets suffice
    s_fast = (win.y - win.x >= 2 && win.w - win.z >= 2 && win.z + lty >= 1 && win.w + lty <= f.height - 1 && f.pitch >= 16 &&
              f.pitch * (long long)(f.height + 1) < (1LL << 31))
                 ? 1 : 0;
    s_org.p = f.data + (long long)lty * f.pitch + (long long)ltx * 3;
    s_org.pitch = f.pitch;
  }
  __syncthreads();
  const int x = blockIdx.x * 32 + threadIdx.x;
  if (x >= S) return;
  const int2 col = s_col[threadIdx.x];
  const int4 win = s_win;
  const int xlo = win.x, xhi = win.y, ylo = win.z, yhi = win.w;
  const Org o_ = s_org;
  const uint8_t* org = o_.p;
  const bool fast = s_fast != 0;
  // INTERIOR blocks (block-uniform): the source coordinates are an affine function of (x, y) up to rounding, so if the
  // four corner pixels of the block sample at least one pixel inside the valid window, every pixel's four taps are
  // inside it and the validity / clamp / shift-back logic of the general fast path is not needed at all.
  bool interior = fast;
  if (fast) {
    const int cx1 = min(31, S - 1 - (int)blockIdx.x * 32), ry1 = min(CROP_ROWS - 1, S - 1 - (int)blockIdx.y * CROP_ROWS);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int2 c = s_col[(k & 1) ? cx1 : 0], r = s_row[(k >> 1) ? ry1 : 0];
      const int qx = (r.x + c.x) >> 10, qy = (r.y + c.y) >> 10;
      interior = interior && qx - 1 >= xlo && qx + 2 < xhi && qy - 1 >= ylo && qy + 2 < yhi;
    }
  }
  // destination of this thread's first row; every further row is 8 image rows down (64-bit address maths out of the loop)
  T* q_row = kToClip ? dst + b * sB + t * sT + (long long)(blockIdx.y * CROP_ROWS + threadIdx.y) * sH + x * sW : nullptr;
  uint8_t* q8_row = kToClip ? nullptr : out_u8 + (((long long)bt * S + blockIdx.y * CROP_ROWS + threadIdx.y) * S + x) * 3;
  const long long q_step = 8 * sH;
#pragma unroll 1
  for (int ry = threadIdx.y; ry < CROP_ROWS; ry += 8, q_row += q_step, q8_row += 8 * S * 3) {
  const int y = blockIdx.y * CROP_ROWS + ry;
  if (y >= S) break;
  const int2 row = s_row[ry];
  const int X = (row.x + col.x) >> 5, Y = (row.y + col.y) >> 5;
  int sx = X >> 5, sy = Y >> 5;
  sx = max(-32768, min(32767, sx));
  sy = max(-32768, min(32767, sy));
  const int fx = X & 31, fy = Y & 31;
  // OpenCV's BilinearTab_i entry: rint(float(a/32) * float(b/32) * 32768) saturated to int16.  Both factors and the
  // product are exact in fp32, so this is the integer 32*a*b (<= 32768) with the one saturating case a = b = 32.
  int w00 = min((32 - fy) * (32 - fx) * 32, 32767), w01 = (32 - fy) * fx * 32;
  int w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
  int acc[3] = {0, 0, 0};                 // in memory channel order; BGR frames are swapped at the end
  if (interior) {
    const int pitch = (int)o_.pitch;
    const uint8_t* p = org + (sy * pitch + sx * 3);
    const uint2 r0 = load6(p), r1 = load6(p + pitch);
    const int a0 = r0.x & 255, a1 = (r0.x >> 8) & 255, a2 = (r0.x >> 16) & 255;
    const int b0 = r0.x >> 24, b1 = r0.y & 255, b2 = (r0.y >> 8) & 255;
    const int d0 = r1.x & 255, d1 = (r1.x >> 8) & 255, d2 = (r1.x >> 16) & 255;
    const int e0 = r1.x >> 24, e1 = r1.y & 255, e2 = (r1.y >> 8) & 255;
    acc[0] = w00 * a0 + w01 * b0 + w10 * d0 + w11 * e0;
    acc[1] = w00 * a1 + w01 * b1 + w10 * d1 + w11 * e1;
    acc[2] = w00 * a2 + w01 * b2 + w10 * d2 + w11 * e2;
  } else if (fast) {
    const int pitch = (int)o_.pitch;
    // taps outside the window weigh nothing
    const bool vx0 = sx >= xlo && sx < xhi, vx1 = sx + 1 >= xlo && sx + 1 < xhi;
    const bool vy0 = sy >= ylo && sy < yhi, vy1 = sy + 1 >= ylo && sy + 1 < yhi;
    w00 = (vx0 && vy0) ? w00 : 0; w01 = (vx1 && vy0) ? w01 : 0;
    w10 = (vx0 && vy1) ? w10 : 0; w11 = (vx1 && vy1) ? w11 : 0;
    // fetch position clamped into the window; ddx / ddy = how far the wanted position is from the fetched one
    const int csx = min(max(sx, xlo), xhi - 2), csy = min(max(sy, ylo), yhi - 2);
    const int ddx = sx - csx, ddy = sy - csy;
    const uint8_t* p = org + (csy * pitch + csx * 3);
    uint2 r0 = load6(p), r1 = load6(p + pitch);
    // rows: wanted (sy, sy+1) = fetched (csy, csy+1) shifted by ddy (only |ddy| <= 1 leaves a tap with weight)
    const uint2 t0 = ddy == 1 ? r1 : r0, t1 = ddy == -1 ? r0 : r1;
    r0 = t0; r1 = t1;
    // columns: 6 bytes = pixels (csx, csx+1); ddx = +1: wanted pixel sx is the second one -> shift right by 3 bytes;
    // ddx = -1: wanted pixel sx+1 is the first one -> shift left by 3 bytes
    if (ddx == 1) {
      r0 = make_uint2(__funnelshift_r(r0.x, r0.y, 24), r0.y >> 24);
      r1 = make_uint2(__funnelshift_r(r1.x, r1.y, 24), r1.y >> 24);
    } else if (ddx == -1) {
      r0 = make_uint2(r0.x << 24, __funnelshift_l(r0.x, r0.y, 24));
      r1 = make_uint2(r1.x << 24, __funnelshift_l(r1.x, r1.y, 24));
    }
    const int a0 = r0.x & 255, a1 = (r0.x >> 8) & 255, a2 = (r0.x >> 16) & 255;
    const int b0 = r0.x >> 24, b1 = r0.y & 255, b2 = (r0.y >> 8) & 255;
    const int d0 = r1.x & 255, d1 = (r1.x >> 8) & 255, d2 = (r1.x >> 16) & 255;
    const int e0 = r1.x >> 24, e1 = r1.y & 255, e2 = (r1.y >> 8) & 255;
    acc[0] = w00 * a0 + w01 * b0 + w10 * d0 + w11 * e0;
    acc[1] = w00 * a1 + w01 * b1 + w10 * d1 + w11 * e1;
    acc[2] = w00 * a2 + w01 * b2 + w10 * d2 + w11 * e2;
  } else {
    const long long pitch = o_.pitch;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cx = sx + (k & 1), cy = sy + (k >> 1);
      const int wgt = k == 0 ? w00 : k == 1 ? w01 : k == 2 ? w10 : w11;
      if (cx >= xlo && cx < xhi && cy >= ylo && cy < yhi && wgt != 0) {
        const uint8_t* p = org + (long long)cy * pitch + (long long)cx * 3;
        acc[0] += wgt * (int)p[0];
        acc[1] += wgt * (int)p[1];
        acc[2] += wgt * (int)p[2];
      }
    }
  }
  int o[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = (acc[c] + 16384) >> 15;      // in [0,255] by construction: weights >= 0, sum <= 2^15
  if (bgr) { const int t0 = o[0]; o[0] = o[2]; o[2] = t0; }

  if (kToClip) {
    T* q = q_row;
    if (sC == 0) {
      store_px4<T>(q, s_lut[o[0]], s_lut[256 + o[1]], s_lut[512 + o[2]]);
    } else {      // caller tensor viewed as [B,3,T,S,S] with its own strides (af_crop_pack)
      store_1<T>(q, s_lut[o[0]]);
      store_1<T>(q + sC, s_lut[256 + o[1]]);
      store_1<T>(q + 2 * sC, s_lut[512 + o[2]]);
    }
  } else {
    uint8_t* q = q8_row;
    q[0] = (uint8_t)o[0]; q[1] = (uint8_t)o[1]; q[2] = (uint8_t)o[2];
  }
  }
}

// lut[c][v] = (float(v) - mean_c) / std_c in IEEE fp32 (demo.py:84-87,317-319), one entry per u8 value
__global__ void norm_lut_kernel(Norm nrm, float* __restrict__ lut) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 768) lut[i] = norm1((float)(i & 255), nrm.mean[i >> 8], nrm.stdv[i >> 8]);
}

// u8 [B,T,S,S,3] aligned clips -> normalised padded NDHWC4 clip (A4 of SURVEY.md §8a).
// One block per 8 image rows of (b, t): no per-pixel index division; S % 4 == 0: a thread converts 4 pixels (three aligned
// 32-bit loads, four 8-byte stores), otherwise one pixel at a time.  lut as in crop_kernel: fp32 [3][256], built with
// the callers' IEEE subtraction / division, so the result equals the fp32 pack lines bit for bit.
template <typename T>
__global__ void __launch_bounds__(64) pack_u8_kernel(const uint8_t* __restrict__ src, int T_, int S, T* dst, long long sB,
                                                     long long sT, long long sH, long long sW,
                                                     const float* __restrict__ lut) {
  __shared__ float s_lut[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = __ldg(lut + i);
  __syncthreads();
  const int t = blockIdx.y, b = blockIdx.z;
  for (int y = blockIdx.x * 8; y < min(S, blockIdx.x * 8 + 8); ++y) {
  const uint8_t* row = src + (((long long)b * T_ + t) * S + y) * (long long)S * 3;
  T* out = dst + b * sB + t * sT + y * sH;
  if ((S & 3) == 0 && ((reinterpret_cast<uintptr_t>(row) & 3) == 0)) {
    const uint32_t* row4 = reinterpret_cast<const uint32_t*>(row);
    for (int g = threadIdx.x; g < S / 4; g += blockDim.x) {
      const uint32_t w0 = __ldg(row4 + 3 * g), w1 = __ldg(row4 + 3 * g + 1), w2 = __ldg(row4 + 3 * g + 2);
      T* q = out + (long long)(4 * g) * sW;
      store_px4<T>(q, s_lut[w0 & 255], s_lut[256 + ((w0 >> 8) & 255)], s_lut[512 + ((w0 >> 16) & 255)]);
      store_px4<T>(q + sW, s_lut[w0 >> 24], s_lut[256 + (w1 & 255)], s_lut[512 + ((w1 >> 8) & 255)]);
      store_px4<T>(q + 2 * sW, s_lut[(w1 >> 16) & 255], s_lut[256 + (w1 >> 24)], s_lut[512 + (w2 & 255)]);
      store_px4<T>(q + 3 * sW, s_lut[(w2 >> 8) & 255], s_lut[256 + ((w2 >> 16) & 255)], s_lut[512 + (w2 >> 24)]);
    }
  } else {
    for (int x = threadIdx.x; x < S; x += blockDim.x) {
      const uint8_t* p = row + 3 * x;
      store_px4<T>(out + (long long)x * sW, s_lut[p[0]], s_lut[256 + p[1]], s_lut[512 + p[2]]);
    }
  }
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

// Normalisation tables live in device memory per (device, mean, std); a table is built the first time its
// constants are seen (callers use one or two sets: demo.py's and the services' rounding of 255*mean).
struct NormLutCache { float* dev[4] = {}; Norm key[4] = {}; int used = 0, next = 0; };
static NormLutCache g_norm_lut[64];
static std::mutex g_norm_lut_mutex;      // engines on different host threads share the per-device tables

static int norm_lut_for(const Norm& n, cudaStream_t s, const float** out) {
  int dev = 0;
  AFB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("crop: device index %d out of range", dev); return AF_ERR_INVALID; }
  std::lock_guard<std::mutex> lock(g_norm_lut_mutex);
  NormLutCache& c = g_norm_lut[dev];
  for (int i = 0; i < c.used; ++i)
    if (memcmp(&c.key[i], &n, sizeof(Norm)) == 0) { *out = c.dev[i]; return AF_OK; }
  int slot = c.used < 4 ? c.used : c.next;
  if (c.used == 4) {
    AFB_CUDA(cudaDeviceSynchronize());       // a launch in flight may still read the table being replaced
    c.next = (c.next + 1) & 3;
  }
  if (!c.dev[slot]) AFB_CUDA(cudaMalloc(&c.dev[slot], 768 * sizeof(float)));
  norm_lut_kernel<<<3, 256, 0, s>>>(n, c.dev[slot]);
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  AFB_CUDA(cudaStreamSynchronize(s));        // other streams may use the table as soon as this call returns
  c.key[slot] = n;
  if (c.used < 4) ++c.used;
  *out = c.dev[slot];
  return AF_OK;
}

int crop_launch(const FrameDesc* frames, const ClipGeom* geom, int B, int T, int S, int bgr,
                uint8_t* out_u8, const ClipLayout* dst, const float mean[3], const float stdv[3],
                cudaStream_t s) {
  if (B <= 0) return AF_OK;
  dim3 grid((S + 31) / 32, (S + CROP_ROWS - 1) / CROP_ROWS, B * T), block(32, 8);
  Norm n = {{0, 0, 0}, {1, 1, 1}};
  if (mean && stdv)
    for (int c = 0; c < 3; ++c) { n.mean[c] = mean[c]; n.stdv[c] = stdv[c]; }
  const float* lut = nullptr;
  if (dst != nullptr) {
    int rc = norm_lut_for(n, s, &lut);
    if (rc) return rc;
  }
  if (dst == nullptr) {
    crop_kernel<float, false><<<grid, block, 0, s>>>(frames, geom, T, S, bgr, out_u8, nullptr, 0, 0, 0, 0, 0, nullptr);
  } else if (dst->is_bf16) {
    crop_kernel<bf16, true><<<grid, block, 0, s>>>(frames, geom, T, S, bgr, nullptr, (bf16*)dst->base, dst->sB,
                                                   dst->sT, dst->sH, dst->sW, dst->sC, lut);
  } else {
    crop_kernel<float, true><<<grid, block, 0, s>>>(frames, geom, T, S, bgr, nullptr, (float*)dst->base, dst->sB,
                                                    dst->sT, dst->sH, dst->sW, dst->sC, lut);
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
  if (B <= 0) return AF_OK;
  Norm n;
  for (int c = 0; c < 3; ++c) { n.mean[c] = mean[c]; n.stdv[c] = stdv[c]; }
  const float* lut = nullptr;
  int rc = norm_lut_for(n, s, &lut);
  if (rc) return rc;
  dim3 grid((d.S + 7) / 8, d.T, B);
  if (d.is_bf16)
    pack_u8_kernel<bf16><<<grid, 64, 0, s>>>(src, d.T, d.S, (bf16*)d.base, d.sB, d.sT, d.sH, d.sW, lut);
  else
    pack_u8_kernel<float><<<grid, 64, 0, s>>>(src, d.T, d.S, (float*)d.base, d.sB, d.sT, d.sH, d.sW, lut);
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
