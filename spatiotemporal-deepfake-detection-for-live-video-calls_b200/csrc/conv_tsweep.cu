// K2c: temporal-sweep tcgen05 kernel for the 3x1x1 `a` convs of s2 (Cout = 64, Cin <= 256).
//
// In the generic implicit-GEMM kernel every temporal tap re-loads its own 128-pixel x 64-channel activation
// tile and weight tile, which keeps these N=64 layers on the L2->SM bandwidth limit although their HBM
// traffic is only input-once + output-once.  Here a work unit is one 128-pixel tile over FOUR consecutive
// output frames (4 accumulators in TMEM, 2 units double buffered = 512 columns): the 3 x Cin/64 weight tiles
// stay resident in shared memory, and each (input frame, channel block) tile is loaded ONCE and used by every
// output frame it feeds (input frame f -> output g with dt = f - g), 6 frame loads per 4 outputs instead of 12.
// Temporal zero padding and the ragged last pixel tile of a frame come from TMA out-of-bounds fill / clipping
// on a 4-D (C, H*W, T, B) view.
//
// Replaces nn.Conv3d([3,1,1], pad [1,0,0]) + BatchNorm3d(eval) + ReLU of BottleneckTransform.a
// (altfreezing/slowfast/models/resnet_helper.py:268-281,313-316).  Warp roles / issue discipline as in
// conv_umma.cu and conv_rows.cu.
#include <cuda.h>
#include <string.h>

#include "../../include/afb200.h"
#include "common.cuh"
#include "umma_ptx.cuh"

namespace afb {
namespace {

constexpr int TS_N = 64;            // output channels (== Cout)
constexpr int TS_G = 4;             // output frames per work unit
constexpr int TS_THREADS = 320;
constexpr int TS_A_BYTES = 128 * 64 * 2;      // one activation tile, 16 KB
constexpr int TS_W_TILE = TS_N * 64 * 2;      // one weight tile, 8 KB
constexpr int TS_OUT_BYTES = 128 * 64 * 2;
constexpr int TS_MAX_STAGES = 8;

struct TsParams {
  float bias_v[64];    // launch-parameter copy of the bias (constant bank operands in the epilogue)
  int cblocks;          // Cin / 64 (1..4)
  int B, T, HW;
  int ptiles;           // ceil(HW / 128)
  int num_units;        // B * (T/4) * ptiles
  int relu;
  int stages;
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d_ts(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   (uint64_t)m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__global__ void __launch_bounds__(TS_THREADS, 1)
conv_tsweep_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                   const __grid_constant__ CUtensorMap tm_y, const TsParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  const int w_bytes = 3 * p.cblocks * TS_W_TILE;
  uint8_t* smem_w = smem;                                   // [3 taps][cblocks][64 x 64] bf16, SWIZZLE_128B
  uint8_t* smem_a = smem_w + w_bytes;                       // [stages] x 16 KB
  uint8_t* smem_out = smem_a + p.stages * TS_A_BYTES;       // [2] x 16 KB
  float* bias_s = reinterpret_cast<float*>(smem_out + 2 * TS_OUT_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + TS_N);
  uint64_t* empty_bar = full_bar + TS_MAX_STAGES;
  uint64_t* w_full = empty_bar + TS_MAX_STAGES;
  uint64_t* tmem_full = w_full + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 2 * TS_G * TS_N;   // 512
  const int tgroups = p.T / TS_G;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_y);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 256); }
    fence_barrier_init();
  } else if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {                               // weights are constants: load before waiting for the prior grid
      mbar_expect_tx(w_full, w_bytes);
      for (int dt = 0; dt < 3; ++dt)
        for (int c = 0; c < p.cblocks; ++c)
          tma_load_2d(smem_w + (dt * p.cblocks + c) * TS_W_TILE, &tm_w, w_full, c * 64, dt * TS_N);
    }
    __syncwarp();
    pdl_wait_prior_grid();
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int r = unit;
      const int pt = r % p.ptiles; r /= p.ptiles;
      const int tg = r % tgroups;
      const int b = r / tgroups;
      for (int f = 0; f < TS_G + 2; ++f) {           // input frame tg*4 - 1 + f (out of range -> zero fill)
        for (int c = 0; c < p.cblocks; ++c) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&full_bar[stage], TS_A_BYTES);
            tma_load_4d(smem_a + stage * TS_A_BYTES, &tm_a, &full_bar[stage], c * 64, pt * 128, tg * TS_G - 1 + f, b);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    pdl_wait_prior_grid();
    constexpr uint32_t idesc = make_idesc(TS_N);
    mbar_wait(w_full, 0);
    tc_fence_after();
    const uint32_t w_base = smem_u32(smem_w);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++it) {
      const int as = it & 1;
      mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int f = 0; f < TS_G + 2; ++f) {
        const int g_lo = f > 2 ? f - 2 : 0, g_hi = f < TS_G - 1 ? f : TS_G - 1;
        for (int c = 0; c < p.cblocks; ++c) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * TS_A_BYTES);
          if (elect_one()) {
            const uint64_t adesc = make_smem_desc(a_addr);
            for (int g = g_lo; g <= g_hi; ++g) {     // output frame g sees input frame f as tap dt = f - g
              const int dt = f - g;
              const uint32_t d_tmem = tmem_base + (as * TS_G + g) * TS_N;
              const uint64_t bdesc = make_smem_desc(w_base + (dt * p.cblocks + c) * TS_W_TILE);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (dt | c | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (elect_one()) umma_commit(&tmem_full[as]);
      __syncwarp();
    }
  } else {
    // ===================================================== epilogue: two warpgroups, alternate output frames
    pdl_wait_prior_grid();
    const int eg = (warp - 2) >> 2;
    const int et = (threadIdx.x - 64) & 127;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint8_t* sout = smem_out + eg * TS_OUT_BYTES;
    int it = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++it) {
      const int as = it & 1;
      mbar_wait(&tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
      int r = unit;
      const int pt = r % p.ptiles; r /= p.ptiles;
      const int tg = r % tgroups;
      const int b = r / tgroups;
#pragma unroll 1
      for (int g = eg; g < TS_G; g += 2) {
        if (et == 0) tma_store_wait_read<0>();
        epi_bar_sync(eg);
        uint32_t v[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (as * TS_G + g) * TS_N;
        TMEM_LD_32x32b_x32(taddr, v);
        TMEM_LD_32x32b_x32(taddr + 32, (v + 32));
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]) + p.bias_v[q * 8 + e];
          uint4 o;
          uint32_t* o2 = reinterpret_cast<uint32_t*>(&o);
          if (p.relu) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o2[e] = pack_bf16x2_relu(f[2 * e], f[2 * e + 1]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) o2[e] = pack_bf16x2(f[2 * e], f[2 * e + 1]);
          }
          *reinterpret_cast<uint4*>(sout + row * 128 + ((q ^ (row & 7)) << 4)) = o;
        }
        fence_proxy_async_smem();
        epi_bar_sync(eg);
        if (et == 0) {
          tma_store_4d_ts(&tm_y, sout, 0, pt * 128, tg * TS_G + g, b);     // rows past H*W are clipped
          tma_store_commit();
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_ts_encode = nullptr;
int g_ts_sms = 0, g_ts_max_smem = 0;

int ts_encode(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
              const cuuint32_t* box, const char* what) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = g_ts_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                           es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(%s) failed: %d", what, (int)r); return AF_ERR_CUDA; }
  return AF_OK;
}

}  // namespace

int conv_tsweep_init() {
  static bool configured[64] = {};
  int dev = 0;
  AFB_CUDA(cudaGetDevice(&dev));
  if (!g_ts_encode) {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    AFB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available"); return AF_ERR_UNSUPPORTED; }
    g_ts_encode = (EncodeTiledFn)fn;
    AFB_CUDA(cudaDeviceGetAttribute(&g_ts_sms, cudaDevAttrMultiProcessorCount, dev));
    AFB_CUDA(cudaDeviceGetAttribute(&g_ts_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  }
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    AFB_CUDA(cudaFuncSetAttribute(conv_tsweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_ts_max_smem));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  return AF_OK;
}

bool conv_tsweep_supported(const ConvProblem& p) {
  if (!g_ts_encode) return false;
  static const bool off = getenv("AFB200_NO_TSWEEP") != nullptr;
  if (off) return false;
  if (p.Cout != TS_N || p.Cin % 64 != 0 || p.Cin > 256 || p.res != nullptr || p.pool_t || p.pool_hw || !p.bias_host) return false;
  if (p.kt != 3 || p.kh != 1 || p.kw != 1 || p.st != 1 || p.sh != 1 || p.sw != 1 || p.pt != 1 || p.ph != 0 || p.pw != 0) return false;
  if (p.Ti % TS_G != 0 || p.To != p.Ti) return false;
  if (p.xsW != p.Cin || p.xsH != (long long)p.Wi * p.Cin || p.xsT != (long long)p.Hi * p.Wi * p.Cin ||
      p.xsB != (long long)p.Ti * p.Hi * p.Wi * p.Cin)
    return false;
  return true;
}

int conv_tsweep_launch(const ConvProblem& p, cudaStream_t s) {
  TsParams tp;
  memcpy(tp.bias_v, p.bias_host, sizeof(tp.bias_v)); tp.cblocks = p.Cin / 64; tp.B = p.B; tp.T = p.Ti; tp.HW = p.Hi * p.Wi;
  tp.ptiles = (tp.HW + 127) / 128;
  tp.num_units = p.B * (p.Ti / TS_G) * tp.ptiles;
  tp.relu = p.relu;
  const int fixed = 3 * tp.cblocks * TS_W_TILE + 2 * TS_OUT_BYTES + TS_N * 4 + (2 * TS_MAX_STAGES + 8) * 8 + 16 + 1024;
  tp.stages = (g_ts_max_smem - fixed) / TS_A_BYTES;
  if (tp.stages > TS_MAX_STAGES) tp.stages = TS_MAX_STAGES;
  if (tp.stages < 2) { set_error("conv_tsweep: not enough shared memory"); return AF_ERR_INVALID; }
  const int dyn = fixed + tp.stages * TS_A_BYTES;

  alignas(64) CUtensorMap ta, tw, ty;
  {
    cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)tp.HW, (cuuint64_t)p.Ti, (cuuint64_t)p.B};
    cuuint64_t strides[3] = {(cuuint64_t)p.Cin * 2, (cuuint64_t)tp.HW * p.Cin * 2, (cuuint64_t)p.Ti * tp.HW * p.Cin * 2};
    cuuint32_t box[4] = {64, 128, 1, 1};
    int rc = ts_encode(&ta, p.x, 4, dims, strides, box, "tsweep A");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)3 * p.Cout};
    cuuint64_t strides[1] = {(cuuint64_t)p.Cin * 2};
    cuuint32_t box[2] = {64, TS_N};
    int rc = ts_encode(&tw, p.w, 2, dims, strides, box, "tsweep W");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)p.Cout, (cuuint64_t)tp.HW, (cuuint64_t)p.To, (cuuint64_t)p.B};
    cuuint64_t strides[3] = {(cuuint64_t)p.Cout * 2, (cuuint64_t)tp.HW * p.Cout * 2, (cuuint64_t)p.To * tp.HW * p.Cout * 2};
    cuuint32_t box[4] = {64, 128, 1, 1};
    int rc = ts_encode(&ty, p.y, 4, dims, strides, box, "tsweep Y");
    if (rc) return rc;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(limit_grid(tp.num_units, g_ts_sms));
  cfg.blockDim = dim3(TS_THREADS); cfg.dynamicSmemBytes = dyn; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  AFB_CUDA(cudaLaunchKernelEx(&cfg, conv_tsweep_kernel, ta, tw, ty, tp));
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

}  // namespace afb
