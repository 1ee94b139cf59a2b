// K2b / K3: "row-halo" tcgen05 conv for the wide, shallow layers — the stem and the 1x3x3 convs of s2 —
// where Cout = 64 makes the generic implicit-GEMM kernel L2-bandwidth bound (every filter tap would re-load
// its own 128x64 activation tile and its weight tile).
//
// An output tile is an 8-wide x 16-tall patch of one frame.  For each (channel block, dt, dx) "column-tap
// group" ONE TMA box is loaded; the kh vertical taps are different views of that box: the UMMA descriptor's
// start address moves by one image row of 8 pixels (a whole swizzle atom, so the swizzle phase is unchanged).
// Zero padding comes from TMA out-of-bounds fill (generic mode) or the clip's physical pads (direct stem).
// The group's kh weight tiles are loaded once per work unit of G=4 tiles (4 accumulators live in TMEM, 2 units
// double buffered = 512 columns).
//
//   conv_rows_kernel<false>  NDHWC-64 activations, 128-byte K rows, SWIZZLE_128B: s2 `b` convs (K2b) and the
//                            unfolded-stem fallback
//   conv_rows_kernel<true>   the stem straight from the padded NDHWC4 clip (K3): overlapping-window tensor map,
//                            64-byte K rows, SWIZZLE_64B, output rows two input rows (1024 B) apart
//
// Warp roles as in conv_umma.cu: warp 0 TMA producer, warp 1 MMA issuer / TMEM owner (both walk their loops
// with all lanes and issue through elect.sync so operands stay in uniform registers), warps 2-9 two epilogue
// warpgroups (bias + ReLU -> bf16 -> swizzled smem -> 4-D TMA store, or the fused 3x3/2 max-pool: interior
// windows stored, tile-border windows merged with 16-byte red.global.max.bf16x2).  Programmatic dependent launch.
// Stride 1 only; no residual (neither the stem nor `b` convs have one:
// altfreezing/slowfast/models/stem_helper.py:173-178, resnet_helper.py:311-326).
#include <cuda.h>
#include <string.h>

#include "../../include/afb200.h"
#include "common.cuh"
#include "umma_ptx.cuh"

namespace afb {
namespace {

constexpr int RB_N = 64;          // output channels per tile (== Cout)
constexpr int RB_X = 8, RB_R = 16;  // tile: 8 columns x 16 rows = 128 output pixels
constexpr int RB_G = 4;           // tiles per work unit (accumulators sharing one weight load)
constexpr int RB_THREADS = 320;   // TMA warp, MMA warp, 2 epilogue warpgroups
constexpr int RB_OUT_BYTES = 128 * 64 * 2;

struct RowsParams {
  // the bias travels in the launch parameters (constant bank): the epilogues add it as constant-cache operands instead of
  // shared-memory loads, which compete with the tensor cores' operand reads for the shared-memory data pipe
  float bias_v[RB_N];
  int Cin, kt, kh, kw, pt, ph, pw;
  int B, To, Ho, Wo;
  int x_tiles, y_tiles, num_tiles, num_units;
  int relu;
  int stages;          // A ring depth
  int a_stage_bytes;   // ring slot size (1024-aligned)
  int a_tx_bytes;      // bytes one A box actually delivers (mbarrier expect_tx)
  int w_buf_bytes;     // kh * 64 * 128
  int w_resident;      // all phases' weights fit in smem: loaded once per CTA instead of once per unit
  int prefetch;        // stem sweep: L2-prefetch the next unit's boxes
  int pool;            // fused 3x3/2 max-pool epilogue
  bf16* pool_out;      // [B*To, Ho/2, Wo/2, 64], zero-initialised by the caller
};

__device__ __forceinline__ void tma_load_tile_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                                 int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];" ::"r"(smem_u32(dst)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   (uint64_t)m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// One 8x16-pixel output tile of the epilogue: accumulator (TMEM, fp32) -> +bias, ReLU -> bf16 tile in swizzled
// smem -> either a 4-D TMA store or the fused 3x3/2 max-pool.  Called by all 128 threads of one epilogue group.
__device__ __forceinline__ void rows_store_or_pool(const RowsParams& p, const CUtensorMap* tm_y_ptr, uint8_t* sout, int eg, int et,
                                                   int xt, int yt, int r);

__device__ __forceinline__ void rows_epilogue_tile(const RowsParams& p, const CUtensorMap* tm_y_ptr, uint8_t* sout,
                                                   const float* bias_s, uint32_t taddr, int eg, int et, int row, int xt,
                                                   int yt, int r) {
  if (et == 0) tma_store_wait_read<0>();
  epi_bar_sync(eg);
  uint32_t v[64];
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
  rows_store_or_pool(p, tm_y_ptr, sout, eg, et, xt, yt, r);
}

// Second half of the epilogue: the finished 8x16-pixel bf16 tile in swizzled smem -> a 4-D TMA store, or the fused
// 3x3/2 max-pool.  Called by all 128 threads of one epilogue group.
__device__ __forceinline__ void rows_store_or_pool(const RowsParams& p, const CUtensorMap* tm_y_ptr, uint8_t* sout, int eg, int et,
                                                   int xt, int yt, int r) {
  const CUtensorMap& tm_y = *tm_y_ptr;
  if (!p.pool) {
    fence_proxy_async_smem();
    epi_bar_sync(eg);
    if (et == 0) {
      tma_store_4d(&tm_y, sout, 0, xt * RB_X, yt * RB_R, r);
      tma_store_commit();
    }
  } else {
    // Fused MaxPool3d k[1,3,3] s[1,2,2] p[0,1,1] (stem_helper.py:166-168): the 8x16 conv tile in
    // smem feeds 5x9 pooled positions.  The 3x7 interior ones are complete and stored plainly;
    // the border ones are shared with neighbouring tiles and merged with a 16-byte vector
    // red.max into the zero-initialised output (post-ReLU values are >= 0).
    epi_bar_sync(eg);
    const int Hp = p.Ho >> 1, Wp = p.Wo >> 1;
    const int py0 = yt * (RB_R / 2), px0 = xt * (RB_X / 2);
    const int ly_max = min(RB_R - 1, p.Ho - 1 - yt * RB_R);      // rows past the image hold garbage
    for (int item = et; item < 45 * 8; item += 128) {
      const int ch = item & 7, pos = item >> 3;
      const int k = pos / 5, j = pos - k * 5;
      const int py = py0 + k, px = px0 + j;
      if (py >= Hp || px >= Wp) continue;
      // window = conv rows 2k-1..2k+1 x columns 2j-1..2j+1 clipped to the tile: fully unrolled, loads predicated
      uint4 m = make_uint4(0u, 0u, 0u, 0u);
      __nv_bfloat162* m2 = reinterpret_cast<__nv_bfloat162*>(&m);
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int ly = 2 * k - 1 + dy;
        if (ly < 0 || ly > ly_max) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int lx = 2 * j - 1 + dx;
          if (lx < 0 || lx >= RB_X) continue;
          const uint4 t = *reinterpret_cast<const uint4*>(sout + ly * (RB_X * 128) + lx * 128 + ((ch ^ lx) << 4));
          const __nv_bfloat162* t2 = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
          for (int e = 0; e < 4; ++e) m2[e] = __hmax2(m2[e], t2[e]);
        }
      }
      bf16* dst = p.pool_out + (((long long)r * Hp + py) * Wp + px) * RB_N + ch * 8;
      const bool interior = k >= 1 && k <= 7 && j >= 1 && j <= 3;
      if (interior) {
        *reinterpret_cast<uint4*>(dst) = m;
      } else {
        asm volatile("red.global.v4.bf16x2.max.noftz [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(m.x), "r"(m.y), "r"(m.z),
                     "r"(m.w)
                     : "memory");
      }
    }
  }
}

// kDirect selects the operand geometry at compile time so the MMA issue loop stays branch-free:
//   false: NDHWC-64 activations, 128-byte K rows (SWIZZLE_128B), 4 MMAs per vertical tap
//   true : padded NDHWC4 clip windows, 64-byte K rows (SWIZZLE_64B), 2 MMAs per vertical tap
template <bool kDirect>
__global__ void __launch_bounds__(RB_THREADS, 1)
conv_rows_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                 const __grid_constant__ CUtensorMap tm_y, const RowsParams p) {
  pdl_launch_dependents();
  constexpr int ROW_BYTES = kDirect ? 64 : 128;
  constexpr int KSTEPS = ROW_BYTES / 32;
  constexpr uint32_t LAYOUT = kDirect ? 4u : 2u;
  constexpr uint32_t A_TAP_BYTES = RB_X * ROW_BYTES, W_TILE_BYTES = RB_N * ROW_BYTES, SBO_B = 8 * ROW_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  // carve-up: [A ring][W double buffer][out ring x2][bias][barriers][tmem ptr]
  uint8_t* smem_a = smem;
  uint8_t* smem_w = smem_a + p.stages * p.a_stage_bytes;
  const int cblocks = p.Cin / 64;
  const int num_phases = cblocks * p.kt * p.kw;
  uint8_t* smem_out = smem_w + (p.w_resident ? num_phases : 2) * p.w_buf_bytes;
  float* bias_s = reinterpret_cast<float*>(smem_out + 2 * RB_OUT_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + RB_N);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* w_full = empty_bar + 8;
  uint64_t* w_empty = w_full + 2;
  uint64_t* tmem_full = w_empty + 2;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 2 * RB_G * RB_N;   // 512

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_y);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1);
      mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 256);
    }
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
  pdl_wait_prior_grid();      // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ===================================================== TMA producer (all lanes walk, one issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t wcount = 0;                 // weight-buffer uses so far
      if (p.w_resident) {                  // every phase's kh weight tiles, once per CTA
        if (elect_one()) {
          mbar_expect_tx(&w_full[0], num_phases * p.w_buf_bytes);
          for (int ph = 0; ph < num_phases; ++ph) {
            const int dx = ph % p.kw, q = ph / p.kw, dt = q % p.kt, cb = q / p.kt;
            for (int dy = 0; dy < p.kh; ++dy)
              tma_load_2d(smem_w + ph * p.w_buf_bytes + dy * W_TILE_BYTES, &tm_w, &w_full[0], cb * 64,
                          ((dt * p.kh + dy) * p.kw + dx) * RB_N);
          }
        }
        __syncwarp();
      }
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        for (int ph = 0; ph < num_phases; ++ph) {
          // phase -> (channel block, dt, dx)
          const int dx = ph % p.kw, q = ph / p.kw, dt = q % p.kt, cb = q / p.kt;
          if (!p.w_resident) {
            const int wb = wcount & 1;
            mbar_wait(&w_empty[wb], ((wcount >> 1) & 1) ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&w_full[wb], p.w_buf_bytes);
              for (int dy = 0; dy < p.kh; ++dy) {
                const int tap = (dt * p.kh + dy) * p.kw + dx;
                tma_load_2d(smem_w + wb * p.w_buf_bytes + dy * W_TILE_BYTES, &tm_w, &w_full[wb], cb * 64, tap * RB_N);
              }
            }
            __syncwarp();
            ++wcount;
          }
          for (int g = 0; g < RB_G; ++g) {
            const int tile = unit * RB_G + g;
            if (tile >= p.num_tiles) break;
            int r = tile;
            const int xt = r % p.x_tiles; r /= p.x_tiles;
            const int yt = r % p.y_tiles; r /= p.y_tiles;
            const int to = r % p.To;
            const int b = r / p.To;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&full_bar[stage], p.a_tx_bytes);
              if (kDirect)   // physical pads in the clip: output (yo,xo) reads rows 2yo.., pixels 2xo.. of frame to+dt
                tma_load_tile_5d(smem_a + stage * p.a_stage_bytes, &tm_a, &full_bar[stage], 0, xt * RB_X, 2 * yt * RB_R,
                                 to + dt, b);
              else
                tma_load_tile_5d(smem_a + stage * p.a_stage_bytes, &tm_a, &full_bar[stage], cb * 64,
                                 xt * RB_X + dx - p.pw, yt * RB_R - p.ph, to + dt - p.pt, b);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (all lanes walk the loop so the
    // operands stay warp-uniform; one elected lane issues)
    {
      constexpr uint32_t idesc = make_idesc(RB_N);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t wcount = 0;
      int it = 0;
      if (p.w_resident) { mbar_wait(&w_full[0], 0); tc_fence_after(); }
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        for (int ph = 0; ph < num_phases; ++ph) {
          const int wb = p.w_resident ? ph : (int)(wcount & 1);
          if (!p.w_resident) {
            mbar_wait(&w_full[wb], (wcount >> 1) & 1);
            tc_fence_after();
          }
          const uint32_t w_addr = smem_u32(smem_w + wb * p.w_buf_bytes);
          for (int g = 0; g < RB_G; ++g) {
            const int tile = unit * RB_G + g;
            if (tile >= p.num_tiles) break;
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem_a + stage * p.a_stage_bytes);
            const uint32_t d_tmem = tmem_base + (as * RB_G + g) * RB_N;
            if (elect_one()) {
              for (int dy = 0; dy < p.kh; ++dy) {
                // vertical tap dy = the same box viewed one image row (8 pixels) further down
                const uint64_t adesc = make_smem_desc_ex(a_addr + dy * A_TAP_BYTES, 1024, LAYOUT);
                const uint64_t bdesc = make_smem_desc_ex(w_addr + dy * W_TILE_BYTES, SBO_B, LAYOUT);
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k)
                  umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (ph | dy | k) != 0 ? 1u : 0u);
              }
              umma_commit(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          if (!p.w_resident) {
            if (elect_one()) umma_commit(&w_empty[wb]);
            __syncwarp();
            ++wcount;
          }
        }
        if (elect_one()) umma_commit(&tmem_full[as]);
        __syncwarp();
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..9): two warpgroups,
    // alternate tiles of the unit, one output slot each
    const int eg = (warp - 2) >> 2;
    const int et = (threadIdx.x - 64) & 127;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint8_t* sout = smem_out + eg * RB_OUT_BYTES;
    int it = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int g = eg; g < RB_G; g += 2) {
        const int tile = unit * RB_G + g;
        if (tile >= p.num_tiles) break;
        int r = tile;
        const int xt = r % p.x_tiles; r /= p.x_tiles;
        const int yt = r % p.y_tiles; r /= p.y_tiles;   // r = b*To + to
        rows_epilogue_tile(p, &tm_y, sout, bias_s, tmem_base + ((uint32_t)(quad * 32) << 16) + (as * RB_G + g) * RB_N, eg, et, row,
                           xt, yt, r);
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


// K3 (default stem path): temporal-sweep variant of the direct stem.  A work unit is one 8x16 spatial tile over
// FOUR consecutive output frames (4 accumulators).  The 35 (dt,dy) weight tiles (140 KB) stay resident in shared
// memory for the CTA's lifetime, and each of the unit's 8 input-frame boxes is loaded ONCE and used by every
// (output frame, dt) pair it participates in (input frame f feeds output g with dt = f - g): 2 boxes per output
// tile instead of 5 boxes + 5 weight groups, which takes the stem off the L2->SM bandwidth limit.  All n <= 4 output
// frames a box feeds are issued as ONE 128 x 64n x 16 MMA (adjacent weight tiles, adjacent accumulators).
constexpr int SW_W_BYTES = 35 * RB_N * 64;        // 143360
constexpr int SW_A_STAGES = 2;

__global__ void __launch_bounds__(RB_THREADS, 1)
stem_sweep_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                  const __grid_constant__ CUtensorMap tm_y, const RowsParams p) {
  pdl_launch_dependents();
  constexpr uint32_t LAYOUT = 4u, A_TAP_BYTES = RB_X * 64, W_TILE_BYTES = RB_N * 64, SBO_B = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* smem_w = smem;                                        // [35][64][32] bf16, SWIZZLE_64B
  uint8_t* smem_a = smem_w + SW_W_BYTES;                         // [2] boxes
  uint8_t* smem_out = smem_a + SW_A_STAGES * p.a_stage_bytes;    // [2] x 16 KB
  float* bias_s = reinterpret_cast<float*>(smem_out + 2 * RB_OUT_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + RB_N);
  uint64_t* empty_bar = full_bar + SW_A_STAGES;
  uint64_t* w_full = empty_bar + SW_A_STAGES;
  uint64_t* tmem_full = w_full + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 2 * RB_G * RB_N;   // 512
  const int tgroups = p.To / RB_G;                  // output-frame groups per clip (To % 4 == 0)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < SW_A_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
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
    if (elect_one()) {                              // weights are constants: no need to wait for the prior grid
      mbar_expect_tx(w_full, SW_W_BYTES);
      // tile (dt,dy) of W35 lands at slot dy*5 + (4-dt): the tiles of taps dt and dt-1 (same dy) are then adjacent, so
      // one N=128 MMA can feed two consecutive output frames from the same input box (see the MMA issuer)
      for (int i = 0; i < 35; ++i) {
        const int dt = i / 7, dy = i - dt * 7;
        tma_load_2d(smem_w + (dy * 5 + 4 - dt) * W_TILE_BYTES, &tm_w, w_full, 0, i * RB_N);
      }
    }
    __syncwarp();
    pdl_wait_prior_grid();
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int r = unit;
      const int xt = r % p.x_tiles; r /= p.x_tiles;
      const int yt = r % p.y_tiles; r /= p.y_tiles;
      const int tg = r % tgroups;
      const int b = r / tgroups;
      if (p.prefetch && unit + (int)gridDim.x < p.num_units) {
        // the NEXT unit's 8 boxes go to L2 now: with two ring slots the HBM latency of a box is longer than its MMAs
        int rn = unit + gridDim.x;
        const int xn = rn % p.x_tiles; rn /= p.x_tiles;
        const int yn = rn % p.y_tiles; rn /= p.y_tiles;
        const int tn = rn % tgroups, bn = rn / tgroups;
        if (elect_one())
          for (int f = 0; f < RB_G + 4; ++f) tma_prefetch_l2_5d(&tm_a, 0, xn * RB_X, 2 * yn * RB_R, tn * RB_G + f, bn);
        __syncwarp();
      }
      for (int f = 0; f < RB_G + 4; ++f) {          // padded frame index of logical frame tg*4 - 2 + f
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], p.a_tx_bytes);
          tma_load_tile_5d(smem_a + stage * p.a_stage_bytes, &tm_a, &full_bar[stage], 0, xt * RB_X, 2 * yt * RB_R,
                           tg * RB_G + f, b);
        }
        __syncwarp();
        if (++stage == SW_A_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    pdl_wait_prior_grid();
    constexpr uint32_t idesc = make_idesc(RB_N);
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
      for (int f = 0; f < RB_G + 4; ++f) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + stage * p.a_stage_bytes);
        if (elect_one()) {
          const int g_lo = f > 4 ? f - 4 : 0, g_hi = f < RB_G - 1 ? f : RB_G - 1;
          // Output frame g of the unit sees this input frame as tap dt = f - g.  The n = g_hi - g_lo + 1 output frames
          // fed by this box use taps (dt, dt-1, ...) of the SAME activation rows: their weight tiles are adjacent in smem
          // and their accumulators adjacent in TMEM, so ONE 128 x 64n x 16 MMA does them all and reads the A operand
          // once -- an N=64 MMA is bound by its shared-memory operand reads (6 KB per 32 tensor clocks), the wide
          // ones are not (n = 4: 12 KB per 128 clocks).  Only the very first MMA of a frame whose accumulator opens here
          // (dt = 0: g_hi = f) is issued on its own, because the accumulate flag is per instruction.
          const int n = g_hi - g_lo + 1;
          const int dt_lo = f - g_lo;                 // tap of the first (oldest) output frame
          const bool opens = f < RB_G;                // g_hi = f receives its first contribution (tap 0)
          const uint32_t d_tmem = tmem_base + (as * RB_G + g_lo) * RB_N;
          const uint32_t idesc_n = make_idesc(RB_N * n), idesc_old = make_idesc(RB_N * (n > 1 ? n - 1 : 1));
#pragma unroll
          for (int dy = 0; dy < 7; ++dy) {
            const uint64_t adesc = make_smem_desc_ex(a_addr + dy * A_TAP_BYTES, 1024, LAYOUT);
            const uint64_t bdesc = make_smem_desc_ex(w_base + (dy * 5 + 4 - dt_lo) * W_TILE_BYTES, SBO_B, LAYOUT);
            if (dy == 0 && opens) {
              if (n > 1) umma_bf16(d_tmem, adesc, bdesc, idesc_old, 1u);
              umma_bf16(d_tmem + (n - 1) * RB_N, adesc, bdesc + (uint64_t)((n - 1) * (W_TILE_BYTES >> 4)), idesc, 0u);
            } else {
              umma_bf16(d_tmem, adesc, bdesc, idesc_n, 1u);
            }
            umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc_n, 1u);
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == SW_A_STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full[as]);
      __syncwarp();
    }
  } else {
    // ===================================================== epilogue: two warpgroups, alternate frames of the unit
    pdl_wait_prior_grid();
    const int eg = (warp - 2) >> 2;
    const int et = (threadIdx.x - 64) & 127;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint8_t* sout = smem_out + eg * RB_OUT_BYTES;
    int it = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++it) {
      const int as = it & 1;
      mbar_wait(&tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
      int r = unit;
      const int xt = r % p.x_tiles; r /= p.x_tiles;
      const int yt = r % p.y_tiles; r /= p.y_tiles;
      const int tg = r % tgroups;
      const int b = r / tgroups;
#pragma unroll 1
      for (int g = eg; g < RB_G; g += 2)
        rows_epilogue_tile(p, &tm_y, sout, bias_s, tmem_base + ((uint32_t)(quad * 32) << 16) + (as * RB_G + g) * RB_N, eg, et,
                           row, xt, yt, b * p.To + tg * RB_G + g);
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


// K6 (FTCN-TT plugin): the temporal-only stem on the tensor cores.
//   Conv3d(3->64, k[5,1,1], p[2,0,0]) + folded BN + MaxPool3d(1,2,2) + ReLU + MaxPool3d k[1,3,3] s[1,2,2] p[0,1,1]
//   (i3d_temporal_var_fix_dropout_tt_cfg.py:207-289 applied to stem_helper.py:156-178)
// K = 5 taps x 3 channels is far too shallow for an MMA as it stands, so the GEMM is laid out around the padded NDHWC4
// clip (origin at padded row 4 / column 4 for this variant: pixel PAIRS are 16-byte aligned and row pairs start even):
//   A row    = one horizontal pixel pair of one input row: 2 x 4 bf16 = 16 bytes = exactly one K chunk of the
//              un-swizzled K-major operand layout; the 5 (+1 zero-weight) temporal taps are 6 such chunks, one per
//              frame: K = 48, three K=16 MMAs.  ONE 5-D TMA box brings all 6 frames x 32 rows x 8 pairs of a tile
//   N = 128  = 64 channels for the pair's even pixel (weights on chunk elements 0..2) | 64 for its odd pixel (4..6)
//   M = 128  = 8 pairs x 16 rows of ONE ROW PARITY (the box interleaves the parities in 8-row groups, so a parity is
//              the 8-row groups 256 bytes apart: the descriptor's SBO); the other parity is a second accumulator
// so the four pixels of every 2x2 pooling window are the same accumulator ROW (= epilogue thread) in four column /
// accumulator ranges: the first max-pool is four register-local fmaxf per channel.  The pooled 8x16 tile (112x112 level)
// then goes through the same smem tile + 3x3/2 pooling epilogue as the I3D stem (rows_store_or_pool).
constexpr int FT_FRAME_BYTES = 256 * 16;               // one frame of a tile: 16 row pairs x 2 parities x 8 pixel pairs x 16 B
constexpr int FT_STAGE_BYTES = 6 * FT_FRAME_BYTES;     // 6 frames
constexpr int FT_W_BYTES = 6 * 128 * 16;               // [6 chunks][128 rows][8 bf16]
constexpr int FT_STAGES = 4;

// un-swizzled K-major operand: core matrix = 8 rows x 16 bytes contiguous; SBO = distance between 8-row groups,
// LBO = distance between the two 16-byte K chunks of one K=16 MMA (cute/atom/mma_traits_sm100.hpp, LayoutType::INTERLEAVE)
__device__ __forceinline__ uint64_t make_smem_desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__global__ void __launch_bounds__(RB_THREADS, 1)
ftcn_stem_umma_kernel(const __grid_constant__ CUtensorMap tm_a, const bf16* __restrict__ w2, const RowsParams p,
                      int frames_padded) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem_w + FT_W_BYTES;
  uint8_t* smem_out = smem_a + FT_STAGES * FT_STAGE_BYTES;
  float* bias_s = reinterpret_cast<float*>(smem_out + 2 * RB_OUT_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + RB_N);
  uint64_t* empty_bar = full_bar + FT_STAGES;
  uint64_t* tmem_full = empty_bar + FT_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 512;                  // 2 tiles in flight x (2 row parities x 128 columns)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    for (int i = 0; i < FT_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
    fence_barrier_init();
  } else if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // weights (constants, 12 KB, already in operand layout) and bias: plain copies
  for (int i = threadIdx.x; i < FT_W_BYTES / 16; i += RB_THREADS)
    reinterpret_cast<uint4*>(smem_w)[i] = __ldg(reinterpret_cast<const uint4*>(w2) + i);
  fence_proxy_async_smem();                            // the tensor core reads smem_w through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  pdl_wait_prior_grid();

  if (warp == 0) {
    // ===================================================== TMA producer: 12 boxes per tile
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int r = tile;
      const int xt = r % p.x_tiles; r /= p.x_tiles;
      const int yt = r % p.y_tiles; r /= p.y_tiles;
      const int to = r % p.To;
      const int b = r / p.To;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[stage], FT_STAGE_BYTES);
        // clip view: (8 el of a pixel pair, pair, row parity, row pair, padded frame of any clip).  Logical pixel
        // (y, x) sits at padded (y + 4, x + 4): pair x/2 + 2, row pair y/2 + 2, parity y & 1.  Output frame `to` reads
        // padded frames to .. to+4 (chunk 5 meets zero weights but must hold finite data: frame to+5, zero-filled by
        // TMA when it lies past the last clip)
        tma_load_tile_5d(smem_a + stage * FT_STAGE_BYTES, &tm_a, &full_bar[stage], 0, xt * RB_X + 2, 0, yt * RB_R + 2,
                         b * frames_padded + to);
      }
      __syncwarp();
      if (++stage == FT_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer: 2 accumulators x 3 MMAs (128 x 128 x 16) per tile
    constexpr uint32_t idesc = make_idesc(128);
    const uint32_t w_addr = smem_u32(smem_w);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem_a + stage * FT_STAGE_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int par = 0; par < 2; ++par)
#pragma unroll
          for (int j = 0; j < 3; ++j)
            umma_bf16(tmem_base + as * 256 + par * 128,
                      make_smem_desc_noswz(a_addr + 2 * j * FT_FRAME_BYTES + par * 128, FT_FRAME_BYTES, 256),
                      make_smem_desc_noswz(w_addr + 2 * j * 2048, 2048, 128), idesc, j != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full[as]);
      }
      __syncwarp();
      if (++stage == FT_STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===================================================== epilogue: two warpgroups on alternate tiles
    const int eg = (warp - 2) >> 2;
    const int et = (threadIdx.x - 64) & 127;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint8_t* sout = smem_out + eg * RB_OUT_BYTES;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != eg) continue;
      const int as = it & 1;
      mbar_wait(&tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
      int r = tile;
      const int xt = r % p.x_tiles; r /= p.x_tiles;
      const int yt = r % p.y_tiles; r /= p.y_tiles;   // r = b*To + to
      epi_bar_sync(eg);                               // every thread is done reading the previous tile in `sout`
      const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + as * 256;
      // 16 channels at a time: 4 window pixels x 16 accumulators; the loads of the next 16 are in flight while these
      // are reduced (two register sets)
      uint32_t v[2][4][16];
      TMEM_LD_32x32b_x16(tbase, v[0][0]);             // even row, even pixel
      TMEM_LD_32x32b_x16(tbase + 64, v[0][1]);        // even row, odd pixel
      TMEM_LD_32x32b_x16(tbase + 128, v[0][2]);       // odd row, even pixel
      TMEM_LD_32x32b_x16(tbase + 192, v[0][3]);       // odd row, odd pixel
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        tmem_ld_wait();
        if (q + 1 < 4) {
          TMEM_LD_32x32b_x16(tbase + (q + 1) * 16, v[(q + 1) & 1][0]);
          TMEM_LD_32x32b_x16(tbase + 64 + (q + 1) * 16, v[(q + 1) & 1][1]);
          TMEM_LD_32x32b_x16(tbase + 128 + (q + 1) * 16, v[(q + 1) & 1][2]);
          TMEM_LD_32x32b_x16(tbase + 192 + (q + 1) * 16, v[(q + 1) & 1][3]);
        }
        const uint32_t(*w)[16] = v[q & 1];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint4 o;
          __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float f[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int i = h * 8 + 2 * e + u;
              const float m = fmaxf(fmaxf(__uint_as_float(w[0][i]), __uint_as_float(w[1][i])),
                                    fmaxf(__uint_as_float(w[2][i]), __uint_as_float(w[3][i])));
              f[u] = fmaxf(m + p.bias_v[q * 16 + i], 0.f);
            }
            o2[e] = __floats2bfloat162_rn(f[0], f[1]);
          }
          const int chunk = q * 2 + h;
          *reinterpret_cast<uint4*>(sout + row * 128 + ((chunk ^ (row & 7)) << 4)) = o;
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
      rows_store_or_pool(p, &tm_a, sout, eg, et, xt, yt, r);      // pool mode: the tensor map is not used
    }
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
EncodeTiledFn g_rows_encode = nullptr;
int g_rows_sms = 0;
int g_rows_max_smem = 0;

int encode_nd(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
              const cuuint32_t* box, const char* what, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = g_rows_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides,
                             box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(%s) failed: %d", what, (int)r); return AF_ERR_CUDA; }
  return AF_OK;
}

}  // namespace

int conv_rows_init() {
  static bool configured[64] = {};            // function attributes are per device
  int dev = 0;
  AFB_CUDA(cudaGetDevice(&dev));
  if (!g_rows_encode) {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    AFB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available"); return AF_ERR_UNSUPPORTED; }
    g_rows_encode = (EncodeTiledFn)fn;
    AFB_CUDA(cudaDeviceGetAttribute(&g_rows_sms, cudaDevAttrMultiProcessorCount, dev));
    AFB_CUDA(cudaDeviceGetAttribute(&g_rows_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  }
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    AFB_CUDA(cudaFuncSetAttribute(conv_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_rows_max_smem));
    AFB_CUDA(cudaFuncSetAttribute(conv_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_rows_max_smem));
    AFB_CUDA(cudaFuncSetAttribute(stem_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_rows_max_smem));
    AFB_CUDA(cudaFuncSetAttribute(ftcn_stem_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_rows_max_smem));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  return AF_OK;
}

bool conv_rows_supported(const ConvProblem& p) {
  if (!g_rows_encode) return false;
  if (p.Cout != RB_N || p.Cin % 64 != 0 || p.res != nullptr || !p.bias_host) return false;
  if (p.st != 1 || p.sh != 1 || p.sw != 1) return false;
  if (p.Wo % RB_X != 0 || p.kh < 2 || p.kh > 8) return false;    // needs vertical taps to pay off
  if (p.pool_hw && (!p.relu || (p.Ho & 1) || (p.Wo & 1))) return false;
  if (p.xsW != p.Cin || p.xsH != (long long)p.Wi * p.Cin || p.xsT != (long long)p.Hi * p.Wi * p.Cin ||
      p.xsB != (long long)p.Ti * p.Hi * p.Wi * p.Cin)
    return false;
  return true;
}

int conv_rows_launch(const ConvProblem& p, cudaStream_t s) {
  RowsParams rp = {};
  memcpy(rp.bias_v, p.bias_host, sizeof(rp.bias_v)); rp.Cin = p.Cin; rp.kt = p.kt; rp.kh = p.kh; rp.kw = p.kw; rp.pt = p.pt; rp.ph = p.ph; rp.pw = p.pw;
  rp.B = p.B; rp.To = p.To; rp.Ho = p.Ho; rp.Wo = p.Wo; rp.relu = p.relu;
  rp.pool = p.pool_hw; rp.pool_out = (bf16*)p.y;
  rp.x_tiles = p.Wo / RB_X;
  rp.y_tiles = (p.Ho + RB_R - 1) / RB_R;
  rp.num_tiles = p.B * p.To * rp.y_tiles * rp.x_tiles;
  rp.num_units = (rp.num_tiles + RB_G - 1) / RB_G;
  rp.a_stage_bytes = (RB_R + p.kh - 1) * RB_X * 128;
  rp.a_tx_bytes = rp.a_stage_bytes;
  rp.w_buf_bytes = p.kh * RB_N * 128;

  const int phases = (p.Cin / 64) * p.kt * p.kw;
  static const bool no_res = getenv("AFB200_NO_RESIDENT_W") != nullptr;
  rp.w_resident = (!no_res && phases * rp.w_buf_bytes <= 96 * 1024) ? 1 : 0;
  const int fixed = (rp.w_resident ? phases : 2) * rp.w_buf_bytes + 2 * RB_OUT_BYTES + RB_N * 4 + 32 * 8 + 16 + 1024;
  rp.stages = (g_rows_max_smem - fixed) / rp.a_stage_bytes;
  if (rp.stages > 8) rp.stages = 8;
  if (rp.stages < 2) { set_error("conv_rows: not enough shared memory"); return AF_ERR_INVALID; }
  const int dyn = fixed + rp.stages * rp.a_stage_bytes;

  alignas(64) CUtensorMap ta, tw, ty;
  {
    cuuint64_t dims[5] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Wi, (cuuint64_t)p.Hi, (cuuint64_t)p.Ti, (cuuint64_t)p.B};
    cuuint64_t strides[4] = {(cuuint64_t)p.Cin * 2, (cuuint64_t)p.Wi * p.Cin * 2, (cuuint64_t)p.Hi * p.Wi * p.Cin * 2,
                             (cuuint64_t)p.Ti * p.Hi * p.Wi * p.Cin * 2};
    cuuint32_t box[5] = {64, RB_X, (cuuint32_t)(RB_R + p.kh - 1), 1, 1};
    int rc = encode_nd(&ta, p.x, 5, dims, strides, box, "rows A");
    if (rc) return rc;
  }
  {
    const int taps = p.kt * p.kh * p.kw;
    cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)taps * p.Cout};
    cuuint64_t strides[1] = {(cuuint64_t)p.Cin * 2};
    cuuint32_t box[2] = {64, RB_N};
    int rc = encode_nd(&tw, p.w, 2, dims, strides, box, "rows W");
    if (rc) return rc;
  }
  {
    const int yo_w = p.pool_hw ? p.Wo / 2 : p.Wo, yo_h = p.pool_hw ? p.Ho / 2 : p.Ho;   // (map unused in pool mode)
    cuuint64_t dims[4] = {(cuuint64_t)p.Cout, (cuuint64_t)yo_w, (cuuint64_t)yo_h, (cuuint64_t)p.B * p.To};
    cuuint64_t strides[3] = {(cuuint64_t)p.Cout * 2, (cuuint64_t)yo_w * p.Cout * 2, (cuuint64_t)yo_h * yo_w * p.Cout * 2};
    cuuint32_t box[4] = {64, RB_X, RB_R, 1};
    int rc = encode_nd(&ty, p.y, 4, dims, strides, box, "rows Y");
    if (rc) return rc;
  }
  const int grid = limit_grid(rp.num_units, g_rows_sms);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(RB_THREADS); cfg.dynamicSmemBytes = dyn; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  AFB_CUDA(cudaLaunchKernelEx(&cfg, conv_rows_kernel<false>, ta, tw, ty, rp));
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}


// The reference stem, Conv3d(3->64, k[5,7,7], s[1,2,2], p[2,3,3]) (stem_helper.py:156-163), straight from the
// engine's padded NDHWC4 clip [B, T+4, S+6, S+8, 4] -- no unfolded copy.  For output column xo the 7 dx taps x 4
// channels of one input row are 28 contiguous bf16 (+4 that meet zero weights) starting at padded pixel 2xo:
// a tensor map whose xo-stride (16 B) is smaller than its 64-byte inner extent hands TMA exactly those
// overlapping windows.  GEMM K = 35 (dt,dy) taps x 32; the 7 dy taps are views of one 37-row box (+512 B each),
// output rows are 2 input rows = 1024 B apart (the descriptor's SBO).  Weights: [35][64][32] bf16.
// y is the zero-initialised POOLED output (fused MaxPool3d [1,3,3]/[1,2,2]) or the dense conv output.
int conv_stem_direct_launch(const void* clip_phys, int B, int T, int S, const void* w35, const float* bias_host, void* y,
                            int pool, cudaStream_t s, int force_per_frame) {
  if (!g_rows_encode) { set_error("conv_stem_direct: not initialised"); return AF_ERR_INVALID; }
  const int Ho = S / 2, Wo = S / 2, Tp = T + 4, Hp = S + 6, Wp = S + 8;
  if ((S & 1) || Wo % RB_X) { set_error("conv_stem_direct: unsupported clip size %d", S); return AF_ERR_INVALID; }
  RowsParams rp = {};
  memcpy(rp.bias_v, bias_host, sizeof(rp.bias_v)); rp.Cin = 64; rp.kt = 5; rp.kh = 7; rp.kw = 1; rp.pt = 0; rp.ph = 0; rp.pw = 0;
  rp.B = B; rp.To = T; rp.Ho = Ho; rp.Wo = Wo; rp.relu = 1;
  rp.x_tiles = Wo / RB_X; rp.y_tiles = (Ho + RB_R - 1) / RB_R;
  rp.num_tiles = B * T * rp.y_tiles * rp.x_tiles;
  rp.num_units = (rp.num_tiles + RB_G - 1) / RB_G;
  const int box_rows = 2 * RB_R + 5;                       // input rows 2*yo0 .. 2*yo0+36
  rp.a_tx_bytes = box_rows * RB_X * 64;
  rp.a_stage_bytes = (rp.a_tx_bytes + 1023) / 1024 * 1024;
  rp.w_buf_bytes = 7 * RB_N * 64;
  rp.pool = pool; rp.pool_out = (bf16*)y;
  rp.w_resident = 0;
  static const char* pf = getenv("AFB200_L2_PREFETCH");      // measured: 1.28 -> 1.45 ms with it (the clip is L2-hot from K1)
  rp.prefetch = pf ? atoi(pf) : 0;
  const int fixed = 2 * rp.w_buf_bytes + 2 * RB_OUT_BYTES + RB_N * 4 + 32 * 8 + 16 + 1024;
  rp.stages = (g_rows_max_smem - fixed) / rp.a_stage_bytes;
  if (rp.stages > 8) rp.stages = 8;
  const int dyn = fixed + rp.stages * rp.a_stage_bytes;

  alignas(64) CUtensorMap ta, tw, ty;
  {
    const cuuint64_t rowpitch = (cuuint64_t)Wp * 8;
    cuuint64_t dims[5] = {32, (cuuint64_t)Wo, (cuuint64_t)Hp, (cuuint64_t)Tp, (cuuint64_t)B};
    cuuint64_t strides[4] = {16, rowpitch, rowpitch * Hp, rowpitch * Hp * Tp};
    cuuint32_t box[5] = {32, RB_X, (cuuint32_t)box_rows, 1, 1};
    int rc = encode_nd(&ta, clip_phys, 5, dims, strides, box, "stem-direct A (overlapping windows)", CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[2] = {32, (cuuint64_t)35 * RB_N};
    cuuint64_t strides[1] = {64};
    cuuint32_t box[2] = {32, RB_N};
    int rc = encode_nd(&tw, w35, 2, dims, strides, box, "stem-direct W", CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  {
    const int yo_w = pool ? Wo / 2 : Wo, yo_h = pool ? Ho / 2 : Ho;
    cuuint64_t dims[4] = {64, (cuuint64_t)yo_w, (cuuint64_t)yo_h, (cuuint64_t)B * T};
    cuuint64_t strides[3] = {128, (cuuint64_t)yo_w * 128, (cuuint64_t)yo_h * yo_w * 128};
    cuuint32_t box[4] = {64, RB_X, RB_R, 1};
    int rc = encode_nd(&ty, y, 4, dims, strides, box, "stem-direct Y");
    if (rc) return rc;
  }
  static const bool no_sweep = getenv("AFB200_NO_STEM_SWEEP") != nullptr;
  const int sweep_dyn = SW_W_BYTES + SW_A_STAGES * rp.a_stage_bytes + 2 * RB_OUT_BYTES + RB_N * 4 + 16 * 8 + 16 + 1024;
  const bool sweep = !no_sweep && !force_per_frame && (T % RB_G == 0) && sweep_dyn <= g_rows_max_smem;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(RB_THREADS); cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (sweep) {
    rp.num_units = B * (T / RB_G) * rp.y_tiles * rp.x_tiles;      // unit = spatial tile x 4 consecutive output frames
    cfg.gridDim = dim3(limit_grid(rp.num_units, g_rows_sms));
    cfg.dynamicSmemBytes = sweep_dyn;
    AFB_CUDA(cudaLaunchKernelEx(&cfg, stem_sweep_kernel, ta, tw, ty, rp));
  } else {
    cfg.gridDim = dim3(limit_grid(rp.num_units, g_rows_sms));
    cfg.dynamicSmemBytes = dyn;
    AFB_CUDA(cudaLaunchKernelEx(&cfg, conv_rows_kernel<true>, ta, tw, ty, rp));
  }
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

// FTCN-TT stem on the tensor cores (ftcn_stem_umma_kernel).  clip_phys: padded NDHWC4 bf16 clip [B, T+4, S+6, S+8, 4]
// whose logical pixel (0,0) sits at padded (4, 4); w2: [6][128][8] bf16 in operand layout (api.cu: upload_ftcn_stem_w2);
// y: ZERO-INITIALISED pooled output [B*T, S/4, S/4, 64].
int ftcn_stem_umma_launch(const void* clip_phys, int B, int T, int S, const void* w2, const float* bias_host, void* y, cudaStream_t s) {
  if (!g_rows_encode) { set_error("ftcn_stem_umma: not initialised"); return AF_ERR_INVALID; }
  const int M2 = S / 2;                                  // 112-level map
  if ((S % 32) != 0) { set_error("ftcn_stem_umma: clip size %d is not a multiple of 32", S); return AF_ERR_INVALID; }
  const int Tp = T + 4, Hp = S + 6, Wp = S + 8;
  RowsParams rp = {};
  memcpy(rp.bias_v, bias_host, sizeof(rp.bias_v)); rp.B = B; rp.To = T; rp.Ho = M2; rp.Wo = M2; rp.relu = 1;
  rp.x_tiles = M2 / RB_X; rp.y_tiles = (M2 + RB_R - 1) / RB_R;
  rp.num_tiles = B * T * rp.y_tiles * rp.x_tiles;
  rp.pool = 1; rp.pool_out = (bf16*)y;
  alignas(64) CUtensorMap ta;
  {
    const cuuint64_t rowpitch = (cuuint64_t)Wp * 8;
    cuuint64_t dims[5] = {8, (cuuint64_t)Wp / 2, 2, (cuuint64_t)Hp / 2, (cuuint64_t)Tp * B};
    cuuint64_t strides[4] = {16, rowpitch, 2 * rowpitch, rowpitch * Hp};
    cuuint32_t box[5] = {8, RB_X, 2, RB_R, 6};
    int rc = encode_nd(&ta, clip_phys, 5, dims, strides, box, "ftcn stem A (pixel pairs)", CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  const int dyn = FT_W_BYTES + FT_STAGES * FT_STAGE_BYTES + 2 * RB_OUT_BYTES + RB_N * 4 + 16 * 8 + 16 + 1024;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(limit_grid(rp.num_tiles, g_rows_sms));
  cfg.blockDim = dim3(RB_THREADS); cfg.dynamicSmemBytes = dyn; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  AFB_CUDA(cudaLaunchKernelEx(&cfg, ftcn_stem_umma_kernel, ta, (const bf16*)w2, rp, Tp));
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

}  // namespace afb
