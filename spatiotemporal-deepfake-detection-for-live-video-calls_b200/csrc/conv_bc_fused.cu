// K2d: the `b` -> `c` tail of an s2 bottleneck block in ONE kernel:
//     y = relu( W_c . relu( conv1x3x3(x; W_b) + bias_b ) + bias_c + residual )
// (BottleneckTransform.forward b, b_bn, b_relu, c, c_bn and ResBlock.forward's add + ReLU:
//  altfreezing/slowfast/models/resnet_helper.py:311-326,438-444; BN folded by the host.)
//
// Why: in s2 (56x56 maps, 64 -> 64 -> 256 channels) the 1x3x3 conv is tensor-pipe bound (N = 64 caps it at the
// shared-memory operand read rate) and the 1x1x1 conv is an HBM stream (51 MB residual in, 51 MB out per clip).  As two
// kernels they run back to back; fused, the `b` MMAs run underneath the `c` epilogue's HBM traffic and the 64-channel
// intermediate (12.8 MB per clip, written and read back) never leaves the SM.
//
// Work unit: one 8-wide x 16-tall output tile (128 pixels) of one frame.
//   warp 0      TMA producer: the three dx-shifted 18-row halo boxes of the tile (the kh = 3 vertical taps are views of
//               one box, as in conv_rows.cu); W_b (9 x 64 x 64) and W_c (256 x 64) are loaded once and stay resident
//   warp 1      MMA issuer: b(i) -> acc_b (128 x 64 fp32, TMEM), and c(i-1): Yb(i-1) [128 x 64 bf16, TMEM] x W_c^T
//               -> acc_c (128 x 256): c of a tile rides inside b of the NEXT one so the first epilogue hides behind it
//   warps 2-5   epilogue 1: acc_b -> +bias_b, ReLU, bf16 -> Yb written back to TENSOR MEMORY (tcgen05.st, two channels per
//               column): the c MMA takes its A operand from TMEM, so the intermediate costs neither 16 KB of shared
//               memory (spent on a third halo-ring slot instead) nor any shared-memory bandwidth
//   warps 6-13  epilogue 2 (two warpgroups, one per 128-channel HALF of acc_c): acc_c -> +bias_c +residual -> ReLU -> bf16,
//               IN PLACE in the slot the residual tile was TMA-loaded into, then a TMA store from that slot.  The c GEMM
//               is issued as two N = 128 halves at different points of the next tile's b MMAs (after its first and its
//               last halo box), each with its own full / empty barriers: the two groups work out of phase instead of
//               bursting on the same issue slots, and a half's MMAs only wait for that half to be drained
// TMEM: acc_b (64 columns) + Yb x2 (2 x 32) + acc_c (256) [+ 128 held output, kPoolT].  Shared memory: 72 + 32 KB weights,
// 3 x 18 KB halo ring, 4 x 16 KB residual/output slots = 223 KB.
//
// kPoolT (last block of s2): the next stage's MaxPool3d k = s = [2,1,1] (video_model_builder.py:474-480,566-568) is
// fused as well.  A CTA's consecutive tiles are the SAME spatial tile of frames 2j and 2j+1; the finished bf16 output of
// the even frame waits in the 128 spare TMEM columns (tcgen05.st, 2 channels per column), the odd frame's epilogue takes
// the element-wise max with it and stores the pooled tile: the un-pooled 256-channel tensor is never written.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/afb200.h"
#include "common.cuh"
#include "umma_ptx.cuh"

namespace afb {
namespace {

constexpr int FX = 8, FR = 16;                  // tile: 8 columns x 16 rows
constexpr int F_MID = 64, F_OUT = 256;          // channels: b 64 -> 64, c 64 -> 256
constexpr int F_THREADS = 32 * 14;
constexpr int F_A_STAGES = 3;
constexpr int F_A_BYTES = (FR + 2) * FX * 128;  // 18 rows x 8 pixels x 64 bf16
constexpr int F_WB_BYTES = 9 * F_MID * 128;
constexpr int F_WC_BYTES = F_OUT * 128;
constexpr int F_TILE_BYTES = 128 * 128;         // 128 pixels x 64 channels bf16
// kShortcut (first block of the stage): W_c is followed by the projection shortcut's weights (second K block of the c
// GEMM, operand = the block INPUT tile, which travels through the halo ring as a 4th box); no residual tile is loaded,
// so two of the four 16 KB slots suffice.
constexpr int f_smem(bool shortcut) {
  return F_WB_BYTES + (shortcut ? 2 : 1) * F_WC_BYTES + F_A_STAGES * F_A_BYTES + (shortcut ? 2 : 4) * F_TILE_BYTES +
         32 * 8 + 16 + 1024;
}

struct FusedParams {
  // biases travel in the kernel parameters (constant bank): the epilogues add them as constant-cache operands instead of
  // shared-memory loads -- the shared-memory data pipe is this kernel's limiter (tensor-core operand reads + staging)
  float bias_b[F_MID];
  float bias_c[F_OUT];
  int x_tiles, y_tiles, frames;     // tiles per row / column of a frame, B*T frames
  int num_tiles;                    // work items per grid: tiles, or (kPoolT) frame-pair units of two tiles each
  int prefetch;                     // L2-prefetch the inputs of the tile this many tiles ahead (0 = off; AFB200_L2_PREFETCH)
};

// The j-th tile of this CTA.  Plain: tile = blockIdx.x + j * gridDim.x over (x tile, y tile, frame).  kPoolT: unit
// = blockIdx.x + (j >> 1) * gridDim.x over (x tile, y tile, frame pair); the unit's two tiles are frames 2*pair + (j & 1).
template <bool kPoolT>
struct TileIter {
  int x_tiles, y_tiles, num, first, stride;
  __device__ __forceinline__ TileIter(const FusedParams& p)
      : x_tiles(p.x_tiles), y_tiles(p.y_tiles), num(p.num_tiles), first(blockIdx.x), stride(gridDim.x) {}
  __device__ __forceinline__ bool valid(int j) const { return first + (kPoolT ? (j >> 1) : j) * stride < num; }
  __device__ __forceinline__ void coords(int j, int& xt, int& yt, int& r) const {
    int u = first + (kPoolT ? (j >> 1) : j) * stride;
    xt = u % x_tiles; u /= x_tiles;
    yt = u % y_tiles; u /= y_tiles;
    r = kPoolT ? 2 * u + (j & 1) : u;
  }
};

__device__ __forceinline__ void f_tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3,
                                              int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];" ::"r"(smem_u32(dst)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void f_tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void f_tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// named barrier among the 128 threads of one warpgroup (ids 1..3; 0 is __syncthreads)
__device__ __forceinline__ void wg_bar_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// tm_r: residual [M,256] (kShortcut = false) / tm_x: the block input [B*T,H,W,64] and tm_ws: shortcut weights (true)
template <bool kShortcut, bool kPoolT>
__global__ void __launch_bounds__(F_THREADS, 1)
conv_bc_fused_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_wb,
                     const __grid_constant__ CUtensorMap tm_wc, const __grid_constant__ CUtensorMap tm_r,
                     const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_x,
                     const __grid_constant__ CUtensorMap tm_ws, const FusedParams p) {
  constexpr int WC_BYTES = (kShortcut ? 2 : 1) * F_WC_BYTES;
  constexpr int N_SLOTS = kShortcut ? 2 : 4;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* smem_wb = smem;
  uint8_t* smem_wc = smem_wb + F_WB_BYTES;
  uint8_t* smem_a = smem_wc + WC_BYTES;
  uint8_t* smem_slot = smem_a + F_A_STAGES * F_A_BYTES;        // [2 groups][2 (1 with kShortcut) slots] x 16 KB
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_slot + N_SLOTS * F_TILE_BYTES);
  uint64_t* a_empty = a_full + F_A_STAGES;
  uint64_t* w_full = a_empty + F_A_STAGES;
  uint64_t* accb_full = w_full + 1;
  uint64_t* accb_empty = accb_full + 1;
  uint64_t* yb_full = accb_empty + 1;     // [2]
  uint64_t* yb_empty = yb_full + 2;       // [2]
  uint64_t* accc_full = yb_empty + 2;     // [2 halves]
  uint64_t* accc_empty = accc_full + 2;   // [2 halves]
  uint64_t* res_full = accc_empty + 2;    // [2 groups][2 slots]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_full + 4);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t YB_COL = F_MID;                   // two bf16 [128 x 64] tiles, 32 columns each
  constexpr uint32_t ACCC_COL = 2 * F_MID;
  constexpr uint32_t HOLD_COL = ACCC_COL + F_OUT;      // kPoolT: the even frame's bf16 output, 2 channels per column
  static_assert(!(kShortcut && kPoolT), "the pooled variant is the identity-residual block");
  const TileIter<kPoolT> tiles(p);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_wb);
    tma_prefetch_desc(&tm_wc);
    if (kShortcut) { tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_ws); }
    tma_prefetch_desc(&tm_r);
    tma_prefetch_desc(&tm_y);
    for (int i = 0; i < F_A_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    mbar_init(w_full, 1);
    mbar_init(accb_full, 1);
    mbar_init(accb_empty, 128);
    for (int i = 0; i < 2; ++i) { mbar_init(&yb_full[i], 128); mbar_init(&yb_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&accc_full[i], 1); mbar_init(&accc_empty[i], 128); }
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

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {                            // weights are constants: no need to wait for the prior grid
      mbar_expect_tx(w_full, F_WB_BYTES + WC_BYTES);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(smem_wb + tap * (F_MID * 128), &tm_wb, w_full, 0, tap * F_MID);
      tma_load_2d(smem_wc, &tm_wc, w_full, 0, 0);
      if (kShortcut) tma_load_2d(smem_wc + F_WC_BYTES, &tm_ws, w_full, 0, 0);
    }
    __syncwarp();
    pdl_wait_prior_grid();
    int stage = 0;
    uint32_t phase = 0;
    // kShortcut: the block input under a tile (no halo; operand of the shortcut's K block) travels through the ring
    // BEHIND the halo boxes of the CTA's next tile, because that is where the MMA warp consumes it (both halves of c of a
    // tile are issued after b of the next one; holding the slot across b's boxes would stall the ring)
    auto load_x = [&](int j) {
      int xt, yt, r;
      tiles.coords(j, xt, yt, r);
      mbar_wait(&a_empty[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&a_full[stage], F_TILE_BYTES);
        f_tma_load_5d(smem_a + stage * F_A_BYTES, &tm_x, &a_full[stage], 0, xt * FX, yt * FR, r, 0);
      }
      __syncwarp();
      if (++stage == F_A_STAGES) { stage = 0; phase ^= 1; }
    };
    // Everything tile j+1 will read from global memory is pulled into L2 while tile j is loaded: the centre halo box
    // (its 1-pixel side columns are neighbouring tiles' centres), the shortcut input tile or the four residual chunks.
    auto prefetch_l2 = [&](int jn) {
      if (!p.prefetch || !tiles.valid(jn)) return;
      int xt, yt, r;
      tiles.coords(jn, xt, yt, r);
      if (elect_one()) {
        tma_prefetch_l2_5d(&tm_a, 0, xt * FX, yt * FR - 1, r, 0);
        if (kShortcut) {
          tma_prefetch_l2_5d(&tm_x, 0, xt * FX, yt * FR, r, 0);
        } else {
          for (int c = 0; c < 4; ++c) tma_prefetch_l2_4d(&tm_r, c * 64, xt * FX, yt * FR, r);
        }
      }
      __syncwarp();
    };
    int j = 0;
    const int pf_dist = p.prefetch;                       // tiles ahead (0 = off)
    for (int q = 0; q < pf_dist; ++q) prefetch_l2(q);
    for (; tiles.valid(j); ++j) {
      int xt, yt, r;                                      // r = frame index b*T + t
      tiles.coords(j, xt, yt, r);
      prefetch_l2(j + pf_dist);
      for (int dx = 0; dx < 3; ++dx) {
        mbar_wait(&a_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&a_full[stage], F_A_BYTES);
          // padding 1 on every side comes from TMA out-of-bounds zero fill
          f_tma_load_5d(smem_a + stage * F_A_BYTES, &tm_a, &a_full[stage], 0, xt * FX + dx - 1, yt * FR - 1, r, 0);
        }
        __syncwarp();
        if (++stage == F_A_STAGES) { stage = 0; phase ^= 1; }
      }
      if (kShortcut && j >= 1) load_x(j - 1);
    }
    if (kShortcut && j >= 1) load_x(j - 1);
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc_b = make_idesc(F_MID), idesc_c = make_idesc(F_OUT / 2);
    mbar_wait(w_full, 0);
    tc_fence_after();
    const uint32_t wb_addr = smem_u32(smem_wb), wc_addr = smem_u32(smem_wc);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    int x_stage = 0;
    // kShortcut: the block-input tile of the tile whose c is about to be issued is the next box of the ring
    auto take_x = [&]() {
      mbar_wait(&a_full[stage], phase);
      tc_fence_after();
      x_stage = stage;
      if (++stage == F_A_STAGES) { stage = 0; phase ^= 1; }
    };
    // half h (output channels [128h, 128h + 128)) of c of this CTA's j-th tile: Yb x W_c^T -> acc_c
    auto issue_c_half = [&](int j, int h) {
      const int ys = j & 1;
      if (h == 0) mbar_wait(&yb_full[ys], (j >> 1) & 1);
      mbar_wait(&accc_empty[h], (j & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_tmem = tmem_base + YB_COL + ys * (F_MID / 2);
        const uint32_t d_c = tmem_base + ACCC_COL + h * (F_OUT / 2);
        const uint64_t bdesc = make_smem_desc(wc_addr + h * (F_WC_BYTES / 2));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ts(d_c, a_tmem + 8 * k, bdesc + 2 * k, idesc_c, k != 0 ? 1u : 0u);
        if (h == 1) umma_commit(&yb_empty[ys]);   // this Yb buffer may be overwritten once both halves have read it
        umma_commit(&accc_full[h]);
      }
      __syncwarp();
    };
    // shortcut form: c and the shortcut's K block at full width (N = 256: the block-input tile is read from shared memory
    // once per k-step; two N = 128 halves measured slower here), both halves' barriers signalled together
    auto issue_c_shortcut = [&](int j) {
      const int ys = j & 1;
      mbar_wait(&yb_full[ys], (j >> 1) & 1);
      mbar_wait(&accc_empty[0], (j & 1) ^ 1);
      mbar_wait(&accc_empty[1], (j & 1) ^ 1);
      take_x();
      if (elect_one()) {
        constexpr uint32_t idesc_full = make_idesc(F_OUT);
        const uint32_t a_tmem = tmem_base + YB_COL + ys * (F_MID / 2);
        const uint64_t bdesc = make_smem_desc(wc_addr);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem_base + ACCC_COL, a_tmem + 8 * k, bdesc + 2 * k, idesc_full, k != 0 ? 1u : 0u);
        umma_commit(&yb_empty[ys]);
        const uint64_t adesc = make_smem_desc(smem_u32(smem_a + x_stage * F_A_BYTES));
        const uint64_t bdesc2 = make_smem_desc(wc_addr + F_WC_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + ACCC_COL, adesc + 2 * k, bdesc2 + 2 * k, idesc_full, 1u);
        umma_commit(&a_empty[x_stage]);
        umma_commit(&accc_full[0]);
        umma_commit(&accc_full[1]);
      }
      __syncwarp();
    };
    for (; tiles.valid(it); ++it) {
      // acc_b is single-buffered: c of the previous tile is queued inside this b, which is all the time the first
      // epilogue needs to read it out
      mbar_wait(accb_empty, (it & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base;
      for (int dx = 0; dx < 3; ++dx) {
        mbar_wait(&a_full[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + stage * F_A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            // vertical tap dy = the same box viewed one image row (8 pixels = 1024 bytes) further down
            const uint64_t adesc = make_smem_desc(a_addr + dy * (FX * 128));
            const uint64_t bdesc = make_smem_desc(wb_addr + (dy * 3 + dx) * (F_MID * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_b, (dx | dy | k) != 0 ? 1u : 0u);
          }
          umma_commit(&a_empty[stage]);
        }
        __syncwarp();
        if (++stage == F_A_STAGES) { stage = 0; phase ^= 1; }
        // c of the previous tile rides inside this b (epilogue 1 has had a whole tile to produce its Yb): the first half
        // behind the first halo box, the second half behind the last one (shortcut form: both behind the last one, where
        // the block-input tile arrives)
        if (!kShortcut && dx == 0 && it >= 1) issue_c_half(it - 1, 0);
      }
      if (elect_one()) umma_commit(accb_full);
      __syncwarp();
      if (it >= 1) {
        if (kShortcut) issue_c_shortcut(it - 1); else issue_c_half(it - 1, 1);
      }
    }
    if (it >= 1) {
      if (kShortcut) {
        issue_c_shortcut(it - 1);
      } else {
        issue_c_half(it - 1, 0);
        issue_c_half(it - 1, 1);
      }
    }
  } else if (warp < 6) {
    // ===================================================== epilogue 1 (warps 2-5): acc_b -> Yb
    pdl_wait_prior_grid();
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;
    int it = 0;
    for (; tiles.valid(it); ++it) {
      mbar_wait(accb_full, it & 1);
      tc_fence_after();
      uint32_t v[64];
      const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
      TMEM_LD_32x32b_x32(tlane, v);
      TMEM_LD_32x32b_x32(tlane + 32, (v + 32));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(accb_empty);
      uint32_t y[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        // even channel (even K) in the low half
        y[j] = pack_bf16x2_relu(__uint_as_float(v[2 * j]) + p.bias_b[2 * j], __uint_as_float(v[2 * j + 1]) + p.bias_b[2 * j + 1]);
      }
      const int ys = it & 1;
      mbar_wait(&yb_empty[ys], ((it >> 1) & 1) ^ 1);      // c of the tile two back has read this Yb buffer
      tc_fence_after();
      TMEM_ST_32x32b_x32(tlane + YB_COL + ys * (F_MID / 2), y);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&yb_full[ys]);
    }
  } else {
    // ===================================================== epilogue 2 (warps 6-13): acc_c + residual -> y
    pdl_wait_prior_grid();
    const int eg = (warp - 6) >> 2;               // group 0: chunks 0, 1 (half 0 of acc_c); group 1: chunks 2, 3
    const int et = (threadIdx.x - 192) & 127;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint8_t* slot_g = smem_slot + eg * (N_SLOTS / 2) * F_TILE_BYTES;
    uint64_t* res_bar = res_full + eg * 2;
    // chunk sequence of this group: (tile it, chunk 2eg), (it, 2eg + 1), (it + 1, 2eg), ...
    int pre_j = 0, pre_chunk = 2 * eg;
    auto issue_res = [&](int slot) {
      int xt, yt, r;
      tiles.coords(pre_j, xt, yt, r);
      mbar_expect_tx(&res_bar[slot], F_TILE_BYTES);
      f_tma_load_4d(slot_g + slot * F_TILE_BYTES, &tm_r, &res_bar[slot], pre_chunk * 64, xt * FX, yt * FR, r);
      if (pre_chunk == 2 * eg) ++pre_chunk; else { pre_chunk = 2 * eg; ++pre_j; }
    };
    if (!kShortcut && et == 0)
      for (int j = 0; j < 2; ++j)
        if (tiles.valid(pre_j)) issue_res(j);
    uint32_t k = 0;
    int it = 0;
    if constexpr (kShortcut) {
      // No residual tile to wait for, and only 16 KB of staging per group: the output goes out in 32-channel half chunks
      // through two 8 KB sub-slots (SWIZZLE_64B boxes), so the TMA store of one half is read out of shared memory while
      // the next half is computed -- with whole 64-channel chunks every chunk waited for the previous store's read.
      for (; tiles.valid(it); ++it) {
        int xt, yt, r;
        tiles.coords(it, xt, yt, r);
        mbar_wait(&accc_full[eg], it & 1);
        tc_fence_after();
#pragma unroll 1
        for (int half = 4 * eg; half < 4 * eg + 4; ++half, ++k) {      // 32-channel pieces 4eg .. 4eg+3 of 8
          uint8_t* s_io = slot_g + (k & 1) * (F_TILE_BYTES / 2);
          uint32_t v[32];
          TMEM_LD_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + ACCC_COL + half * 32, v);
          if (et == 0) tma_store_wait_read<1>();  // the store that last read this sub-slot (two halves ago) has drained it
          wg_bar_sync(2 + eg);
          tmem_ld_wait();
          if (half == 4 * eg + 3) {               // this group's last read of its half of acc_c for the tile
            tc_fence_before();
            mbar_arrive(&accc_empty[eg]);
          }
          const float* bias = p.bias_c + half * 32;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            uint32_t* o2 = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o2[e] = pack_bf16x2_relu(__uint_as_float(v[q * 8 + 2 * e]) + bias[q * 8 + 2 * e],
                                       __uint_as_float(v[q * 8 + 2 * e + 1]) + bias[q * 8 + 2 * e + 1]);
            *reinterpret_cast<uint4*>(s_io + row * 64 + ((q ^ ((row >> 1) & 3)) << 4)) = o;      // SWIZZLE_64B
          }
          fence_proxy_async_smem();
          wg_bar_sync(2 + eg);
          if (et == 0) {
            f_tma_store_4d(&tm_r, s_io, half * 32, xt * FX, yt * FR, r);      // tm_r: the 32-channel-box view of y
            tma_store_commit();
          }
        }
      }
    }
    for (; !kShortcut && tiles.valid(it); ++it) {
      int xt, yt, r;
      tiles.coords(it, xt, yt, r);
      const bool hold_phase = kPoolT && !(it & 1);      // even frame of a pair: the output waits in TMEM
      const bool max_phase = kPoolT && (it & 1);        // odd frame: max with the held output, store the pooled tile
      mbar_wait(&accc_full[eg], it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 2 * eg; chunk < 2 * eg + 2; ++chunk, ++k) {
        const int slot = kShortcut ? 0 : (int)(k & 1);
        uint8_t* s_io = slot_g + slot * F_TILE_BYTES;
        uint32_t v[64];
        uint32_t h[kPoolT ? 32 : 1];
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t taddr = tlane + ACCC_COL + chunk * 64;
        TMEM_LD_32x32b_x32(taddr, v);
        TMEM_LD_32x32b_x32(taddr + 32, (v + 32));
        if (kPoolT) {
          if (max_phase) TMEM_LD_32x32b_x32(tlane + HOLD_COL + chunk * 32, h);
        }
        if (kShortcut) {                          // plain output staging: the previous store must have read the slot
          if (et == 0) tma_store_wait_read<0>();
          wg_bar_sync(2 + eg);
        } else {
          mbar_wait(&res_bar[slot], (k >> 1) & 1u);
        }
        tmem_ld_wait();
        if (chunk == 2 * eg + 1) {                // this group's last read of its half of acc_c for the tile
          tc_fence_before();
          mbar_arrive(&accc_empty[eg]);
        }
        const float* bias = p.bias_c + chunk * 64;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint8_t* pa = s_io + row * 128 + ((q ^ (row & 7)) << 4);
          uint4 t = make_uint4(0u, 0u, 0u, 0u);
          if (!kShortcut) t = *reinterpret_cast<const uint4*>(pa);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&t);
          uint4 o;
          __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float f0 = __uint_as_float(v[q * 8 + 2 * e]) + bias[q * 8 + 2 * e] + __low2float(h2[e]);
            const float f1 = __uint_as_float(v[q * 8 + 2 * e + 1]) + bias[q * 8 + 2 * e + 1] + __high2float(h2[e]);
            const uint32_t pk = pack_bf16x2_relu(f0, f1);
            o2[e] = *reinterpret_cast<const __nv_bfloat162*>(&pk);
          }
          if (kPoolT) {
            if (hold_phase) {
              h[q * 4 + 0] = o.x; h[q * 4 + 1] = o.y; h[q * 4 + 2] = o.z; h[q * 4 + 3] = o.w;
            } else {
              const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&h[q * 4]);
#pragma unroll
              for (int e = 0; e < 4; ++e) o2[e] = __hmax2(o2[e], g2[e]);
              *reinterpret_cast<uint4*>(pa) = o;
            }
          } else {
            *reinterpret_cast<uint4*>(pa) = o;
          }
        }
        if (kPoolT && hold_phase) {
          TMEM_ST_32x32b_x32(tlane + HOLD_COL + chunk * 32, h);
          tmem_st_wait();
          wg_bar_sync(2 + eg);                    // every thread of the group has consumed the residual slot
          if (et == 0 && tiles.valid(pre_j)) issue_res(slot);
          continue;
        }
        fence_proxy_async_smem();
        wg_bar_sync(2 + eg);                      // tile chunk complete in the slot
        if (et == 0) {
          f_tma_store_4d(&tm_y, s_io, chunk * 64, xt * FX, yt * FR, kPoolT ? (r >> 1) : r);
          tma_store_commit();
          if (!kShortcut) {
            tma_store_wait_read<0>();             // the store has read the slot: refill it with the residual two chunks on
            if (tiles.valid(pre_j)) issue_res(slot);
          }
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
EncodeTiledFn g_f_encode = nullptr;
int g_f_sms = 0, g_f_max_smem = 0;

int f_encode(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
             const char* what, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = g_f_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(%s) failed: %d", what, (int)r); return AF_ERR_CUDA; }
  return AF_OK;
}

}  // namespace

int conv_bc_fused_init() {
  static bool configured[64] = {};
  int dev = 0;
  AFB_CUDA(cudaGetDevice(&dev));
  if (!g_f_encode) {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    AFB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available"); return AF_ERR_UNSUPPORTED; }
    g_f_encode = (EncodeTiledFn)fn;
    AFB_CUDA(cudaDeviceGetAttribute(&g_f_sms, cudaDevAttrMultiProcessorCount, dev));
    AFB_CUDA(cudaDeviceGetAttribute(&g_f_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  }
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    if (f_smem(false) <= g_f_max_smem) {
      AFB_CUDA(cudaFuncSetAttribute(conv_bc_fused_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, f_smem(false)));
      AFB_CUDA(cudaFuncSetAttribute(conv_bc_fused_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, f_smem(false)));
    }
    if (f_smem(true) <= g_f_max_smem)
      AFB_CUDA(cudaFuncSetAttribute(conv_bc_fused_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, f_smem(true)));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  return AF_OK;
}

// b: dense NDHWC [B,T,H,W,64] -> 1x3x3, stride 1, pad [0,1,1], 64 -> 64; c: 1x1x1 64 -> 256; residual / y dense [M,256].
// With c.x2 set (projection shortcut fused into c: pointwise, stride 1, 64 -> 256 over the block input c.x2, biases
// pre-summed in c.bias) there is no residual.  With c.pool_t set, y is the temporally max-pooled output [B,T/2,H,W,256].
bool conv_bc_fused_supported(const ConvProblem& b, const ConvProblem& c) {
  if (!g_f_encode || f_smem(c.x2 != nullptr) > g_f_max_smem || !b.bias_host || !c.bias_host) return false;
  if (b.Cin != F_MID || b.Cout != F_MID || b.kt != 1 || b.kh != 3 || b.kw != 3 || b.st != 1 || b.sh != 1 || b.sw != 1 ||
      b.pt != 0 || b.ph != 1 || b.pw != 1 || !b.relu || b.res || b.pool_hw || b.pool_t || b.x2)
    return false;
  if (c.Cin != F_MID || c.Cout != F_OUT || c.kt != 1 || c.kh != 1 || c.kw != 1 || c.st != 1 || c.sh != 1 || c.sw != 1 ||
      !c.relu || c.pool_hw)
    return false;
  if (c.pool_t && (c.x2 || (b.To & 1))) return false;      // pooled variant: identity residual, whole frame pairs
  if (c.x2 ? (c.res || !c.w2 || c.Cin2 != F_MID || c.sh2 != 1 || c.sw2 != 1 || c.T2 != c.To || c.H2 != c.Ho || c.W2 != c.Wo)
           : !c.res)
    return false;
  if (b.Wo % FX != 0 || b.M != c.M) return false;
  if (b.xsW != b.Cin || b.xsH != (long long)b.Wi * b.Cin || b.xsT != (long long)b.Hi * b.Wi * b.Cin ||
      b.xsB != (long long)b.Ti * b.Hi * b.Wi * b.Cin)
    return false;
  return true;
}

int conv_bc_fused_launch(const ConvProblem& b, const ConvProblem& c, cudaStream_t s) {
  FusedParams fp;
  memcpy(fp.bias_b, b.bias_host, sizeof(fp.bias_b));
  memcpy(fp.bias_c, c.bias_host, sizeof(fp.bias_c));
  fp.x_tiles = b.Wo / FX; fp.y_tiles = (b.Ho + FR - 1) / FR; fp.frames = b.B * b.To;
  const bool shortcut = c.x2 != nullptr, pool_t = c.pool_t != 0;
  // measured on B200 (32 clips): prefetching the next tile helps the shortcut form (0.586 -> 0.563 ms) and costs the
  // residual forms 2-20 % (the residual chunks are already TMA-prefetched two chunks ahead; extra L2 requests only compete)
  static const char* pf = getenv("AFB200_L2_PREFETCH");
  fp.prefetch = pf ? atoi(pf) : (shortcut ? 1 : 0);
  fp.num_tiles = (pool_t ? fp.frames / 2 : fp.frames) * fp.y_tiles * fp.x_tiles;
  alignas(64) CUtensorMap ta, twb, twc, tr, ty, tx, tws;
  {
    cuuint64_t dims[5] = {(cuuint64_t)F_MID, (cuuint64_t)b.Wi, (cuuint64_t)b.Hi, (cuuint64_t)b.B * b.Ti, 1};
    cuuint64_t strides[4] = {(cuuint64_t)F_MID * 2, (cuuint64_t)b.Wi * F_MID * 2, (cuuint64_t)b.Hi * b.Wi * F_MID * 2,
                             (cuuint64_t)b.B * b.Ti * b.Hi * b.Wi * F_MID * 2};
    cuuint32_t box[5] = {64, FX, FR + 2, 1, 1};
    int rc = f_encode(&ta, b.x, 5, dims, strides, box, "fused A");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)F_MID, (cuuint64_t)9 * F_MID};
    cuuint64_t strides[1] = {(cuuint64_t)F_MID * 2};
    cuuint32_t box[2] = {64, F_MID};
    int rc = f_encode(&twb, b.w, 2, dims, strides, box, "fused Wb");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)F_MID, (cuuint64_t)F_OUT};
    cuuint64_t strides[1] = {(cuuint64_t)F_MID * 2};
    cuuint32_t box[2] = {64, F_OUT};
    int rc = f_encode(&twc, c.w, 2, dims, strides, box, "fused Wc");
    if (rc) return rc;
  }
  for (int which = 0; which < 2; ++which) {
    cuuint64_t dims[4] = {(cuuint64_t)F_OUT, (cuuint64_t)b.Wo, (cuuint64_t)b.Ho,
                          (cuuint64_t)((which && pool_t) ? fp.frames / 2 : fp.frames)};
    cuuint64_t strides[3] = {(cuuint64_t)F_OUT * 2, (cuuint64_t)b.Wo * F_OUT * 2, (cuuint64_t)b.Ho * b.Wo * F_OUT * 2};
    // shortcut form: there is no residual; `tr` is the 32-channel-box (SWIZZLE_64B) view of y its epilogue stores through
    const bool y32 = !which && shortcut;
    cuuint32_t box[4] = {(cuuint32_t)(y32 ? 32 : 64), FX, FR, 1};
    int rc = f_encode(which ? &ty : &tr, (which || shortcut) ? c.y : c.res, 4, dims, strides, box, which ? "fused Y" : "fused R",
                      y32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  tx = ta; tws = twc;
  if (shortcut) {
    cuuint64_t dims[5] = {(cuuint64_t)F_MID, (cuuint64_t)c.W2, (cuuint64_t)c.H2, (cuuint64_t)b.B * c.T2, 1};
    cuuint64_t strides[4] = {(cuuint64_t)F_MID * 2, (cuuint64_t)c.W2 * F_MID * 2, (cuuint64_t)c.H2 * c.W2 * F_MID * 2,
                             (cuuint64_t)b.B * c.T2 * c.H2 * c.W2 * F_MID * 2};
    cuuint32_t box[5] = {64, FX, FR, 1, 1};
    int rc = f_encode(&tx, c.x2, 5, dims, strides, box, "fused X");
    if (rc) return rc;
    cuuint64_t wdims[2] = {(cuuint64_t)F_MID, (cuuint64_t)F_OUT};
    cuuint64_t wstrides[1] = {(cuuint64_t)F_MID * 2};
    cuuint32_t wbox[2] = {64, F_OUT};
    rc = f_encode(&tws, c.w2, 2, wdims, wstrides, wbox, "fused Ws");
    if (rc) return rc;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(limit_grid(fp.num_tiles, g_f_sms));
  cfg.blockDim = dim3(F_THREADS); cfg.dynamicSmemBytes = f_smem(shortcut); cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (shortcut) AFB_CUDA(cudaLaunchKernelEx(&cfg, conv_bc_fused_kernel<true, false>, ta, twb, twc, tr, ty, tx, tws, fp));
  else if (pool_t) AFB_CUDA(cudaLaunchKernelEx(&cfg, conv_bc_fused_kernel<false, true>, ta, twb, twc, tr, ty, tx, tws, fp));
  else AFB_CUDA(cudaLaunchKernelEx(&cfg, conv_bc_fused_kernel<false, false>, ta, twb, twc, tr, ty, tx, tws, fp));
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

}  // namespace afb
