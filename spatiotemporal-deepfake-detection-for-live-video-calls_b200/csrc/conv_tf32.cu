// K2t: the trunk convolutions of the fp32 engine on the tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM).
//
// The reference runs its fp32 model through cuDNN with TF32 allowed (torch's default for convolutions:
// altfreezing/demo.py:317-324 under torch.backends.cudnn.allow_tf32); this is that path, hand-written: fp32 NDHWC
// activations and fp32 weights go through TMA into 128B-swizzled K-major shared memory as they are, the tensor core
// reads 10 mantissa bits of each operand and accumulates in fp32; bias, residual and ReLU are applied in fp32 and the
// output is stored as fp32, rounded to the nearest TF32 value (weights are rounded the same way at upload), so the
// truncating operand read of the next layer is exact and the rounding error stays unbiased.  This is precision "tf32"
// of the engine (af_create, AF_PREC_TF32); precision "fp32" stays on the exact FFMA kernel (conv_simt.cu), which
// remains the <= 1e-3 parity engine.
//
// Same implicit GEMM as conv_umma.cu with 4-byte elements: a 128-byte swizzle row is 32 channels, so a k-block is
// 32 channels (K = 8 per MMA, four MMAs per k-block); A tiles come from one im2col-mode TMA (or a 2-D map for 1x1x1
// stride-1 convs), B tiles from W[tap][Cout][Cin].  One persistent CTA per SM: warp 0 TMA producer, warp 1 MMA issuer
// (two TMEM accumulators), warps 2-5 epilogue on 32-column chunks (tcgen05.ld -> +bias +residual -> ReLU -> swizzled
// smem -> TMA store, two staging slots).  The stem (3 input channels) stays on the FFMA kernel.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/afb200.h"
#include "common.cuh"
#include "umma_ptx.cuh"

namespace afb {
namespace {

constexpr int T_M = 128;
constexpr int T_KB = 32;                      // fp32 channels per k-block = one 128-byte swizzle row
constexpr int T_THREADS = 192;                // TMA warp, MMA warp, one epilogue warpgroup
constexpr int T_A_BYTES = T_M * 128;          // 16 KB
constexpr int T_OUT_BYTES = T_M * 128;        // one 32-column fp32 output chunk
constexpr int T_MAX_STAGES = 8;

struct Tf32Params {
  const float* bias;
  const float* res;
  long long M;
  int Cout, Cin;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
  int To, Ho, Wo;
  int num_m_tiles, num_n_tiles;
  int relu, im2col, stages;
  int res_slots;            // residual layers: in-place residual / output slots (the rest: 2 plain staging slots)
  int x_tiles, y_tiles, T;  // stem form: 8 x 16-pixel tiles per frame, frames per clip
};

__device__ __forceinline__ void t_tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3,
                                              int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];" ::"r"(smem_u32(dst)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void t_tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// kind::tf32 instruction descriptor: D = f32, A = B = tf32 (format 2), both K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(T_M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// fp32 -> nearest TF32 value (10 mantissa bits), returned as fp32
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// kStem: the stem conv (k[5,7,7] s[1,2,2] p[2,3,3], 3 -> 64; stem_helper.py:156-163) straight from the engine's padded
// NDHWC4 fp32 clip: for output column xo the 7 dx taps x 4 channels of one input row are 28 contiguous floats (+4 that
// meet zero weights) = exactly one 128-byte k-block, so a 5-D tiled map whose xo stride (32 bytes) is smaller than its
// 128-byte inner extent, with traversal stride 2 over the rows, hands the tensor core the windows of an 8 x 16-pixel
// output tile directly; 35 (dt,dy) taps, one k-block each.
template <int BLOCK_N, bool kStem>
__global__ void __launch_bounds__(T_THREADS, 1)
conv_tf32_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_r, const Tf32Params p) {
  constexpr int B_BYTES = BLOCK_N * 128;
  constexpr int STAGE_BYTES = T_A_BYTES + B_BYTES;
  constexpr int CHUNKS = BLOCK_N / 32;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  const bool has_res = !kStem && p.res != nullptr;
  const int n_slots = has_res ? p.res_slots : 2;
  uint8_t* smem_out = smem + p.stages * STAGE_BYTES;            // staging slots (residual layers: residual in, output out)
  float* bias_s = reinterpret_cast<float*>(smem_out + n_slots * T_OUT_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + BLOCK_N);
  uint64_t* empty_bar = full_bar + T_MAX_STAGES;
  uint64_t* tmem_full = empty_bar + T_MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_full = tmem_empty + 2;                          // [4]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_full + 4);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int cblocks = p.Cin / T_KB;
  const int num_kb = kStem ? 35 : p.kt * p.kh * p.kw * cblocks;
  const int stages = p.stages;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_y);
    tma_prefetch_desc(&tm_r);
    for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
    for (int i = 0; i < 4; ++i) mbar_init(&res_full[i], 1);
    fence_barrier_init();
  } else if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  pdl_wait_prior_grid();

  if (warp == 0) {
    // ===================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      if (kStem) {
        int r = tile;
        const int xt = r % p.x_tiles; r /= p.x_tiles;
        const int yt = r % p.y_tiles; r /= p.y_tiles;         // r = b*T + to
        const int b = r / p.T, to = r - b * p.T;
        for (int tap = 0; tap < 35; ++tap) {
          const int dt = tap / 7, dy = tap - dt * 7;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          if (elect_one()) {
            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
            // padded clip: row 2*yo + dy, frame to + dt, window of padded pixels 2*xo .. 2*xo + 7
            t_tma_load_5d(sa, &tm_a, &full_bar[stage], 0, xt * 8, 32 * yt + dy, to + dt, b);
            tma_load_2d(sa + T_A_BYTES, &tm_b, &full_bar[stage], 0, tap * BLOCK_N);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      const int m_tile = tile / p.num_n_tiles, n_tile = tile - m_tile * p.num_n_tiles;
      const int m0 = m_tile * T_M, n0 = n_tile * BLOCK_N;
      int r = m0;
      const int wo = r % p.Wo; r /= p.Wo;
      const int ho = r % p.Ho; r /= p.Ho;
      const int to = r % p.To; r /= p.To;
      const int b = r;
      const int wb = wo * p.sw - p.pw, hb = ho * p.sh - p.ph, tb = to * p.st - p.pt;
      int tap = 0, cb = 0, dx = 0, dy = 0, dt = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (p.im2col)
            tma_load_im2col_5d(sa, &tm_a, &full_bar[stage], cb * T_KB, wb, hb, tb, b, (uint16_t)dx, (uint16_t)dy, (uint16_t)dt);
          else
            tma_load_2d(sa, &tm_a, &full_bar[stage], cb * T_KB, m0);
          tma_load_2d(sa + T_A_BYTES, &tm_b, &full_bar[stage], cb * T_KB, tap * p.Cout + n0);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
        if (++cb == cblocks) {
          cb = 0; ++tap;
          if (++dx == p.kw) { dx = 0; if (++dy == p.kh) { dy = 0; ++dt; } }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc = make_idesc_tf32(BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BLOCK_N;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
        const uint64_t adesc = make_smem_desc(a_addr), bdesc = make_smem_desc(a_addr + T_A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)      // K = 8 fp32 = 32 bytes per MMA: +2 in the descriptor's >>4 address field
            umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full[as]);
      __syncwarp();
    }
  } else {
    // ===================================================== epilogue (warps 2-5)
    const int et = threadIdx.x - 64;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int it = 0;
    // Residual layers work in place (as conv_umma.cu): the residual tile of a chunk is TMA-prefetched into one of R slots,
    // the accumulator is added into it, the TMA store leaves from the same slot and the slot is refilled once that store
    // has read it.  (pre_tile, pre_chunk) walks the chunk sequence R chunks ahead.
    const int R = p.res_slots;
    int pre_tile = blockIdx.x, pre_chunk = 0;
    auto issue_res = [&](int slot) {
      const int m_tile = pre_tile / p.num_n_tiles, n_tile = pre_tile - m_tile * p.num_n_tiles;
      mbar_expect_tx(&res_full[slot], T_OUT_BYTES);
      tma_load_2d(smem_out + slot * T_OUT_BYTES, &tm_r, &res_full[slot], n_tile * BLOCK_N + pre_chunk * 32, m_tile * T_M);
      if (++pre_chunk == CHUNKS) { pre_chunk = 0; pre_tile += gridDim.x; }
    };
    if (has_res && et == 0)
      for (int j = 0; j < R; ++j)
        if (pre_tile < num_tiles) issue_res(j);
    int rslot = 0, prev_rslot = 0;
    uint32_t rphase = 0, k = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int m_tile = tile / p.num_n_tiles, n_tile = tile - m_tile * p.num_n_tiles;
      int xt = 0, yt = 0, bt = 0;
      if (kStem) {
        int r = tile;
        xt = r % p.x_tiles; r /= p.x_tiles;
        yt = r % p.y_tiles; bt = r / p.y_tiles;
        n_tile = 0;
      }
      const long long m0 = (long long)m_tile * T_M;
      const int n0 = n_tile * BLOCK_N;
      const int as = it & 1;
      epi_bar_sync(0);                              // every thread is past the previous tile's bias reads
      for (int i = et; i < BLOCK_N; i += 128) bias_s[i] = __ldg(p.bias + n0 + i);
      mbar_wait(&tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < CHUNKS; ++chunk, ++k) {
        uint8_t* sout = smem_out + (has_res ? rslot : (int)(k & 1)) * T_OUT_BYTES;
        uint32_t v[32];
        TMEM_LD_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + as * BLOCK_N + chunk * 32, v);
        if (!has_res && et == 0) tma_store_wait_read<1>();   // the store that last read this slot (two chunks ago) has drained it
        epi_bar_sync(0);                            // (also publishes bias_s)
        if (has_res) mbar_wait(&res_full[rslot], rphase);
        tmem_ld_wait();
        if (chunk == CHUNKS - 1) {
          tc_fence_before();
          mbar_arrive(&tmem_empty[as]);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float* pa = reinterpret_cast<float*>(sout + row * 128 + ((q ^ (row & 7)) << 4));
          float4 o;
          o.x = __uint_as_float(v[q * 4 + 0]) + bias_s[chunk * 32 + q * 4 + 0];
          o.y = __uint_as_float(v[q * 4 + 1]) + bias_s[chunk * 32 + q * 4 + 1];
          o.z = __uint_as_float(v[q * 4 + 2]) + bias_s[chunk * 32 + q * 4 + 2];
          o.w = __uint_as_float(v[q * 4 + 3]) + bias_s[chunk * 32 + q * 4 + 3];
          if (has_res) {
            const float4 t = *reinterpret_cast<const float4*>(pa);
            o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
          }
          if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          // The next conv's tensor-core read keeps 10 mantissa bits by TRUNCATION, which biases every product towards
          // zero and adds up coherently over 52 layers (measured: 8.5e-3 on the logit); rounding to nearest here, where
          // the value is produced, makes that read exact and the error unbiased.
          o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w);
          *reinterpret_cast<float4*>(pa) = o;
        }
        fence_proxy_async_smem();
        epi_bar_sync(0);
        if (et == 0) {
          if (kStem) t_tma_store_4d(&tm_y, sout, chunk * 32, xt * 8, yt * 16, bt);
          else tma_store_2d(&tm_y, sout, n0 + chunk * 32, (int)m0);
          tma_store_commit();
          if (has_res && k >= 1) {                  // the PREVIOUS chunk's store has read its slot: refill that one
            tma_store_wait_read<1>();
            if (pre_tile < num_tiles) issue_res(prev_rslot);
          }
        }
        if (has_res) {
          prev_rslot = rslot;
          if (++rslot == R) { rslot = 0; rphase ^= 1; }
        }
      }
    }
    if (et == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_t_tiled = nullptr;
EncodeIm2colFn g_t_im2col = nullptr;
int g_t_sms = 0, g_t_max_smem = 0;

int t_encode_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, const char* what) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 4};
  cuuint32_t box[2] = {T_KB, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_t_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(tf32 %s) failed: %d", what, (int)r); return AF_ERR_CUDA; }
  return AF_OK;
}

int t_encode_im2col(CUtensorMap* map, const ConvProblem& p) {
  cuuint64_t dims[5] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Wi, (cuuint64_t)p.Hi, (cuuint64_t)p.Ti, (cuuint64_t)p.B};
  cuuint64_t strides[4] = {(cuuint64_t)p.Cin * 4, (cuuint64_t)p.Wi * p.Cin * 4, (cuuint64_t)p.Hi * p.Wi * p.Cin * 4,
                           (cuuint64_t)p.Ti * p.Hi * p.Wi * p.Cin * 4};
  int lower[3] = {-p.pw, -p.ph, -p.pt};
  int upper[3] = {p.pw - (p.kw - 1), p.ph - (p.kh - 1), p.pt - (p.kt - 1)};
  cuuint32_t es[5] = {1, (cuuint32_t)p.sw, (cuuint32_t)p.sh, (cuuint32_t)p.st, 1};
  CUresult r = g_t_im2col(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(p.x), dims, strides, lower, upper, T_KB,
                          T_M, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeIm2col(tf32) failed: %d", (int)r); return AF_ERR_CUDA; }
  return AF_OK;
}

template <int BLOCK_N, bool kStem>
int t_launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ty, const CUtensorMap& tr, Tf32Params tp,
             cudaStream_t s) {
  static bool configured[64] = {};
  auto kern = conv_tf32_kernel<BLOCK_N, kStem>;
  int dev = 0;
  AFB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    AFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_t_max_smem));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const int stage_bytes = T_A_BYTES + BLOCK_N * 128;
  tp.res_slots = (!kStem && tp.res) ? 4 : 0;
  const int fixed = (tp.res_slots ? tp.res_slots : 2) * T_OUT_BYTES + BLOCK_N * 4 + (2 * T_MAX_STAGES + 8) * 8 + 16 + 1024;
  int stages = (g_t_max_smem - fixed) / stage_bytes;
  if (stages > T_MAX_STAGES) stages = T_MAX_STAGES;
  if (stages < 2) { set_error("conv_tf32: shared memory budget too small"); return AF_ERR_INVALID; }
  tp.stages = stages;
  const int tiles = tp.num_m_tiles * tp.num_n_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(limit_grid(tiles, g_t_sms));
  cfg.blockDim = dim3(T_THREADS); cfg.dynamicSmemBytes = fixed + stages * stage_bytes; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  AFB_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, ty, tr, tp));
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

}  // namespace

int conv_tf32_init() {
  if (g_t_tiled && g_t_im2col) return AF_OK;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  AFB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available"); return AF_ERR_UNSUPPORTED; }
  g_t_tiled = (EncodeTiledFn)fn;
  fn = nullptr;
  AFB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeIm2col not available"); return AF_ERR_UNSUPPORTED; }
  g_t_im2col = (EncodeIm2colFn)fn;
  int dev = 0;
  AFB_CUDA(cudaGetDevice(&dev));
  AFB_CUDA(cudaDeviceGetAttribute(&g_t_sms, cudaDevAttrMultiProcessorCount, dev));
  AFB_CUDA(cudaDeviceGetAttribute(&g_t_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return AF_OK;
}

// fp32 dense NDHWC input [B,Ti,Hi,Wi,Cin], Cin % 32 == 0, Cout % 64 == 0; p.w = fp32 [taps][Cout][Cin]
bool conv_tf32_supported(const ConvProblem& p) {
  if (!g_t_tiled || !g_t_im2col) return false;
  if (p.Cin % T_KB != 0 || p.Cout % 64 != 0 || p.pool_t || p.pool_hw || p.x2) return false;
  if (p.xsW != p.Cin || p.xsH != (long long)p.Wi * p.Cin || p.xsT != (long long)p.Hi * p.Wi * p.Cin ||
      p.xsB != (long long)p.Ti * p.Hi * p.Wi * p.Cin)
    return false;
  if (p.kt > 16 || p.kh > 16 || p.kw > 16 || p.M >= (1LL << 31)) return false;
  return true;
}

int conv_tf32_launch(const ConvProblem& p, cudaStream_t s) {
  Tf32Params tp;
  tp.bias = p.bias; tp.res = (const float*)p.res; tp.M = p.M; tp.Cout = p.Cout; tp.Cin = p.Cin;
  tp.kt = p.kt; tp.kh = p.kh; tp.kw = p.kw; tp.st = p.st; tp.sh = p.sh; tp.sw = p.sw; tp.pt = p.pt; tp.ph = p.ph; tp.pw = p.pw;
  tp.To = p.To; tp.Ho = p.Ho; tp.Wo = p.Wo; tp.relu = p.relu; tp.stages = 0;
  tp.res_slots = 0; tp.x_tiles = tp.y_tiles = tp.T = 0;
  const bool pointwise = p.kt == 1 && p.kh == 1 && p.kw == 1 && p.st == 1 && p.sh == 1 && p.sw == 1;
  tp.im2col = pointwise ? 0 : 1;
  tp.num_m_tiles = (int)((p.M + T_M - 1) / T_M);
  // tile width: the widest of 256 / 128 / 64 that still gives every SM ~2 tiles (an N = 64 TF32 MMA reads 6 KB of operands
  // per 32 clocks and is shared-memory bound like its bf16 counterpart; N = 256 needs 12 KB per 128 clocks)
  int bn = 64;
  for (int cand : {256, 128})
    if (p.Cout % cand == 0 && (long long)tp.num_m_tiles * (p.Cout / cand) >= 2LL * g_t_sms) { bn = cand; break; }
  static const char* fbn = getenv("AFB200_TF32_BLOCK_N");
  if (fbn && (atoi(fbn) == 64 || atoi(fbn) == 128 || atoi(fbn) == 256) && p.Cout % atoi(fbn) == 0) bn = atoi(fbn);
  tp.num_n_tiles = p.Cout / bn;
  alignas(64) CUtensorMap ta, tb, ty, tr;
  int rc = tp.im2col ? t_encode_im2col(&ta, p) : t_encode_2d(&ta, p.x, (uint64_t)p.M, (uint64_t)p.Cin, T_M, "A");
  if (rc) return rc;
  const int taps = p.kt * p.kh * p.kw;
  rc = t_encode_2d(&tb, p.w, (uint64_t)taps * p.Cout, (uint64_t)p.Cin, (uint32_t)bn, "W");
  if (rc) return rc;
  rc = t_encode_2d(&ty, p.y, (uint64_t)p.M, (uint64_t)p.Cout, T_M, "Y");
  if (rc) return rc;
  rc = t_encode_2d(&tr, p.res ? p.res : p.y, (uint64_t)p.M, (uint64_t)p.Cout, T_M, "R");
  if (rc) return rc;
  return bn == 256 ? t_launch<256, false>(ta, tb, ty, tr, tp, s)
                   : bn == 128 ? t_launch<128, false>(ta, tb, ty, tr, tp, s) : t_launch<64, false>(ta, tb, ty, tr, tp, s);
}

// The stem on the tensor cores with TF32 operands.  clip_phys: the engine's padded fp32 NDHWC4 clip
// [B, T+4, S+6, S+8, 4] (logical pixel (0,0,0) at padded (2,3,3), pads zero); w35: fp32 [35 = dt*7+dy][64][32 = dx*4+c]
// (zero for c = 3 and dx = 7); y: fp32 [B*T, S/2, S/2, 64] = relu(conv + bias), the map the fp32 max-pool kernel takes.
int conv_tf32_stem_launch(const void* clip_phys, int B, int T, int S, const void* w35, const float* bias, void* y,
                          cudaStream_t s) {
  if (!g_t_tiled) { set_error("conv_tf32_stem: not initialised"); return AF_ERR_INVALID; }
  const int Ho = S / 2, Wo = S / 2, Tp = T + 4, Hp = S + 6, Wp = S + 8;
  if ((S & 1) || Wo % 8 || Ho % 16) { set_error("conv_tf32_stem: unsupported clip size %d", S); return AF_ERR_INVALID; }
  Tf32Params tp = {};
  tp.bias = bias; tp.res = nullptr; tp.M = (long long)B * T * Ho * Wo; tp.Cout = 64; tp.Cin = 32;
  tp.relu = 1; tp.x_tiles = Wo / 8; tp.y_tiles = Ho / 16; tp.T = T;
  tp.num_m_tiles = B * T * tp.x_tiles * tp.y_tiles; tp.num_n_tiles = 1;
  alignas(64) CUtensorMap ta, tb, ty;
  {
    const cuuint64_t rowpitch = (cuuint64_t)Wp * 16;
    cuuint64_t dims[5] = {32, (cuuint64_t)Wo, (cuuint64_t)Hp, (cuuint64_t)Tp, (cuuint64_t)B};
    cuuint64_t strides[4] = {32, rowpitch, rowpitch * Hp, rowpitch * Hp * Tp};
    cuuint32_t box[5] = {32, 8, 32, 1, 1};               // 16 rows at traversal stride 2
    cuuint32_t es[5] = {1, 1, 2, 1, 1};
    CUresult r = g_t_tiled(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(clip_phys), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(tf32 stem A) failed: %d", (int)r); return AF_ERR_CUDA; }
  }
  int rc = t_encode_2d(&tb, w35, (uint64_t)35 * 64, 32, 64, "stem W");
  if (rc) return rc;
  {
    cuuint64_t dims[4] = {64, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)B * T};
    cuuint64_t strides[3] = {256, (cuuint64_t)Wo * 256, (cuuint64_t)Ho * Wo * 256};
    cuuint32_t box[4] = {32, 8, 16, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = g_t_tiled(&ty, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(tf32 stem Y) failed: %d", (int)r); return AF_ERR_CUDA; }
  }
  return t_launch<64, true>(ta, tb, ty, ty, tp, s);
}

}  // namespace afb
