// K2: channels-last Conv3d (+folded-BN bias, +residual, +ReLU) as an implicit GEMM on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands staged by TMA.
//
// Replaces, for every trunk convolution of the reference (1x3x3, 3x1x1 and 1x1x1, stride 1
// or [1,2,2]; altfreezing/slowfast/models/resnet_helper.py:255-326,411-444), the chain
// nn.Conv3d -> BatchNorm3d(eval) [-> + shortcut] [-> ReLU] that PyTorch runs as 2-4 kernels.
//
// GEMM view: D[M, N] = sum over taps, channel blocks of A_tap[M, 64] * W_tap[N, 64]^T with
//   M = B*To*Ho*Wo flattened output pixels (NDHWC order), N = Cout, K = taps*Cin.
// A tiles: 128 consecutive output pixels x 64 input channels, loaded by ONE im2col-mode TMA
//   (cp.async.bulk.tensor.5d...im2col): the hardware walks the pixels across rows / frames /
//   clips, applies the conv stride, adds the filter-tap offset and zero-fills the padding halo.
//   1x1x1 stride-1 convs use a plain 2-D tiled map over [M, Cin].
// B tiles: BLOCK_N output channels x 64 input channels of one tap from W[tap][Cout][Cin].
// Both land in 128B-swizzled K-major shared memory, which is what the UMMA descriptors read.
//
// One persistent CTA per SM, warp-specialised (320 threads):
//   warp 0   : TMA producer                   -- operand ring of `stages` slots, full/empty mbarriers
//   warp 1   : tcgen05.mma issuer             -- owns the TMEM allocation (2 accumulators)
//              (both walk their loops with all 32 lanes and issue through elect.sync, so descriptors and
//               coordinates live in uniform registers and UTCHMMA/UTMALDG go out back to back)
//   warps 2-9: two epilogue warpgroups on alternate 64-column chunks: tcgen05.ld -> +bias (+residual tile,
//              TMA-prefetched into one of 3-4 slots and updated IN PLACE) -> ReLU inside the bf16x2 conversion ->
//              swizzled smem -> TMA store; optional fused temporal max-pool (tile = 64 pixels x 2 frames).
//              Double-buffered against the next tile's main loop through the tmem_full/tmem_empty barriers.
// Shared memory is carved at launch: residual layers trade operand stages for residual/output slots; small launches
// (batch 1-2) shrink the staging and deepen the operand ring.  kPair: 2-CTA clusters with cta_group::2 MMAs.
// Programmatic dependent launch overlaps each kernel's prologue with its predecessor's tail.
#include <cuda.h>

#include <cstdio>
#include <cstring>
#include <unordered_map>

#include "../../include/afb200.h"
#include "common.cuh"
#include "umma_ptx.cuh"

namespace afb {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;             // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;          // TMA warp, MMA warp, 2 epilogue warpgroups
constexpr int EPI_THREADS = 128;          // per epilogue warpgroup
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KiB
constexpr int OUT_STAGE_BYTES = BLOCK_M * 64 * 2;      // one 64-column output chunk, 16 KiB

struct UmmaParams {
  const float* bias;
  const bf16* res;
  long long M;
  int Cout, Cin;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
  int To, Ho, Wo;
  int num_m_tiles, num_n_tiles;
  int relu;
  int im2col;
  int stages;          // depth of the A/B operand ring
  int out_per_group;   // output staging slots per epilogue group (1 or 2; 0 for residual layers: in place)
  int res_slots;       // residual layers: slots per epilogue group (3 or 4); the output is computed in place in them
  int pool_t;          // fused temporal max-pool: tile = 64 pixels x 2 consecutive frames (rows r, r+64)
  int hw;              // pixels per frame (pool_t only)
  // second operand source (fused projection shortcut): a pointwise conv of stride [1,sh2,sw2] over another
  // tensor, accumulated into the same TMEM tile after the primary taps (0 channel blocks = none)
  int cblocks2, im2col2, sh2, sw2;
  unsigned long long* dbg;   // AFB200_TIMELINE=1: CTA 0 stamps %globaltimer at its phase boundaries (bring-up aid)
};

constexpr int MAX_STAGES = 8;

// Walks, in order, the (tile, 64-column chunk) pairs one epilogue group handles: with
// BLOCK_N >= 128 the two groups take alternate chunks of every tile; with BLOCK_N = 64 they take
// alternate tiles.
template <int CHUNKS> struct EpiIter {
  int tile, it, chunk, num_tiles, stride, eg;
  __device__ __forceinline__ void init(int first_tile, int grid_stride, int n_tiles, int group) {
    num_tiles = n_tiles; stride = grid_stride; eg = group;
    if (CHUNKS == 1) { it = group; tile = first_tile + group * grid_stride; chunk = 0; }
    else { it = 0; tile = first_tile; chunk = group; }
  }
  __device__ __forceinline__ bool valid() const { return tile < num_tiles; }
  __device__ __forceinline__ bool first_in_tile() const { return CHUNKS == 1 || chunk == eg; }
  __device__ __forceinline__ bool last_in_tile() const { return CHUNKS == 1 || chunk + 2 >= CHUNKS; }
  __device__ __forceinline__ void next() {
    if (CHUNKS == 1) { it += 2; tile += 2 * stride; }
    else if (chunk + 2 < CHUNKS) { chunk += 2; }
    else { chunk = eg; ++it; tile += stride; }
  }
};

__device__ __forceinline__ void stamp(const UmmaParams& p, int slot) {
  if (p.dbg && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.dbg[slot] = t;
  }
}

// kPair: the CTA is one half of a 2-CTA cluster running cta_group::2 MMAs on 256-row tiles (see umma_ptx.cuh): it
// loads its own 128 rows of A and HALF of every weight tile, the leader CTA issues the MMAs and commits to the
// barriers of both CTAs, each CTA runs the epilogue of its own 128 accumulator rows.
template <int BLOCK_N, bool kPair>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_r,
                 const __grid_constant__ CUtensorMap tm_a2, const __grid_constant__ CUtensorMap tm_b2,
                 const UmmaParams p) {
  constexpr int B_ROWS = kPair ? BLOCK_N / 2 : BLOCK_N;      // weight rows this CTA stages
  constexpr int B_STAGE_BYTES = B_ROWS * BLOCK_K * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  constexpr int CHUNKS = BLOCK_N / 64;
  pdl_launch_dependents();
  if (threadIdx.x == 0) stamp(p, 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  const bool has_res = p.res != nullptr;
  // carve-up: [operand ring][out slots: 2 groups x out_per_group | residual/output slots: 2 groups x res_slots][bias x2][barriers]
  uint8_t* smem_out = smem + p.stages * STAGE_BYTES;
  uint8_t* smem_res = smem_out + 2 * p.out_per_group * OUT_STAGE_BYTES;
  float* bias_s = reinterpret_cast<float*>(smem_res + (has_res ? 2 * p.res_slots : 0) * OUT_STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + 2 * BLOCK_N);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_full = tmem_empty + 2;                 // [8]
  uint64_t* tmem_empty_peer = res_full + 8;            // [2] leader only: the peer CTA's epilogue has drained the accumulator
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_peer + 2);

  // warp index through a shuffle so the compiler knows it is warp-uniform (keeps the role loops'
  // address arithmetic in uniform registers: tcgen05/TMA operands must be uniform)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // pair mode: a tile is 256 rows (m tiles 2*mp, 2*mp+1) x BLOCK_N; an odd m-tile count leaves the last pair's
  // second CTA with a tile that is entirely out of range (TMA zero-fills its loads and clips its stores)
  const int num_tiles = (kPair ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles) * p.num_n_tiles;
  const int first_tile = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_stride = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int cblocks = p.Cin / BLOCK_K;
  const int num_kb = p.kt * p.kh * p.kw * cblocks;
  const int num_kb_all = num_kb + p.cblocks2;        // primary taps, then the fused shortcut's channel blocks
  const int stages = p.stages;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_y);
    tma_prefetch_desc(&tm_r);
    if (p.cblocks2) { tma_prefetch_desc(&tm_a2); tma_prefetch_desc(&tm_b2); }
    for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], (CHUNKS >= 2 ? 2 : 1) * EPI_THREADS); }
    for (int i = 0; i < 8; ++i) mbar_init(&res_full[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&tmem_empty_peer[i], 1);
    fence_barrier_init();
  } else if (warp == 1) {
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                   "n"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                   "n"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();      // barrier inits visible to the peer before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  if (threadIdx.x == 0) stamp(p, 1);
  // everything above overlapped the previous kernel's tail; the producer warp additionally starts fetching WEIGHTS
  // (never written by a kernel) before it waits for the prior grid
  if (warp != 0) pdl_wait_prior_grid();
  if (threadIdx.x == 0) stamp(p, 2);

  if (warp == 0) {
    // ===================================================== TMA producer (all lanes walk the loop,
    // one elected lane issues; see the MMA warp for why)
    {
      // weight tiles of the first tile's first `pre` k-blocks: issued before griddepcontrol.wait, so their HBM / L2
      // latency (the longest of the first operands: nothing has touched these bytes since the previous pass) overlaps
      // the predecessor's drain.  The stage's barrier is armed here for the whole stage (A arrives later).
      int pre = 0;
      if (first_tile < num_tiles) {
        pre = stages < num_kb ? stages : num_kb;
        const int n_tile0 = first_tile % p.num_n_tiles;
        const int n00 = n_tile0 * BLOCK_N + (kPair ? (int)cta_rank * B_ROWS : 0);
        if (elect_one()) {
          for (int kb = 0; kb < pre; ++kb) {
            uint8_t* sb = smem + kb * STAGE_BYTES + A_STAGE_BYTES;
            const int tap0 = kb / cblocks, cb0 = kb - tap0 * cblocks;
            if (kPair) {
              if (leader) mbar_expect_tx(&full_bar[kb], 2 * STAGE_BYTES);
              tma_load_2d_pair(sb, &tm_b, mapa_u32(&full_bar[kb], 0), cb0 * BLOCK_K, tap0 * p.Cout + n00);
            } else {
              mbar_expect_tx(&full_bar[kb], STAGE_BYTES);
              tma_load_2d(sb, &tm_b, &full_bar[kb], cb0 * BLOCK_K, tap0 * p.Cout + n00);
            }
          }
        }
        __syncwarp();
      }
      pdl_wait_prior_grid();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
        const bool w_ahead = tile == first_tile;      // this tile's first `pre` weight tiles are already in flight
        const int mp = tile / p.num_n_tiles, n_tile = tile - mp * p.num_n_tiles;
        const int m_tile = kPair ? 2 * mp + (int)cta_rank : mp;
        const int m0 = m_tile * BLOCK_M;
        const int n0 = n_tile * BLOCK_N + (kPair ? (int)cta_rank * B_ROWS : 0);      // this CTA's weight rows
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
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (elect_one()) {
            if (kPair) {
              // both CTAs' bytes are counted on the LEADER's full barrier, which alone is armed (for both)
              const uint32_t fb = mapa_u32(&full_bar[stage], 0);
              const bool have_w = w_ahead && kb < pre;
              if (leader && !have_w) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
              if (p.im2col) {
                tma_load_im2col_5d_pair(sa, &tm_a, fb, cb * BLOCK_K, wb, hb, tb, b, (uint16_t)dx, (uint16_t)dy, (uint16_t)dt);
              } else if (p.pool_t) {
                const int ptiles = p.hw >> 6, pair = m_tile / ptiles, pt_ = m_tile - pair * ptiles;
                const int r0 = pair * 2 * p.hw + pt_ * 64;
                tma_load_2d_pair(sa, &tm_a, fb, cb * BLOCK_K, r0);
                tma_load_2d_pair(sa + 64 * 128, &tm_a, fb, cb * BLOCK_K, r0 + p.hw);
              } else {
                tma_load_2d_pair(sa, &tm_a, fb, cb * BLOCK_K, m0);
              }
              if (!have_w) tma_load_2d_pair(sb, &tm_b, fb, cb * BLOCK_K, tap * p.Cout + n0);
            } else {
              const bool have_w = w_ahead && kb < pre;
              if (!have_w) mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
              if (p.im2col) {
                tma_load_im2col_5d(sa, &tm_a, &full_bar[stage], cb * BLOCK_K, wb, hb, tb, b, (uint16_t)dx, (uint16_t)dy,
                                   (uint16_t)dt);
              } else if (p.pool_t) {
                // tile = (frame pair, 64-pixel block): rows 0..63 from frame 2j, rows 64..127 from frame 2j+1
                const int ptiles = p.hw >> 6, pair = m_tile / ptiles, pt_ = m_tile - pair * ptiles;
                const int r0 = pair * 2 * p.hw + pt_ * 64;
                tma_load_2d(sa, &tm_a, &full_bar[stage], cb * BLOCK_K, r0);
                tma_load_2d(sa + 64 * 128, &tm_a, &full_bar[stage], cb * BLOCK_K, r0 + p.hw);
              } else {
                tma_load_2d(sa, &tm_a, &full_bar[stage], cb * BLOCK_K, m0);
              }
              if (!have_w) tma_load_2d(sb, &tm_b, &full_bar[stage], cb * BLOCK_K, tap * p.Cout + n0);
            }
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
          if (++cb == cblocks) {                 // next filter tap (dx fastest, then dy, then dt)
            cb = 0; ++tap;
            if (++dx == p.kw) { dx = 0; if (++dy == p.kh) { dy = 0; ++dt; } }
          }
        }
        // fused projection shortcut (resnet_helper.py:411-423,438-441): x2[b, to, ho*sh2, wo*sw2, :] . W2^T
        for (int cb2 = 0; cb2 < p.cblocks2; ++cb2) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (elect_one()) {
            if (kPair) {
              const uint32_t fb = mapa_u32(&full_bar[stage], 0);
              if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
              if (p.im2col2)
                tma_load_im2col_5d_pair(sa, &tm_a2, fb, cb2 * BLOCK_K, wo * p.sw2, ho * p.sh2, to, b, 0, 0, 0);
              else
                tma_load_2d_pair(sa, &tm_a2, fb, cb2 * BLOCK_K, m0);
              tma_load_2d_pair(sb, &tm_b2, fb, cb2 * BLOCK_K, n0);
            } else {
              mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
              if (p.im2col2)
                tma_load_im2col_5d(sa, &tm_a2, &full_bar[stage], cb2 * BLOCK_K, wo * p.sw2, ho * p.sh2, to, b, 0, 0, 0);
              else
                tma_load_2d(sa, &tm_a2, &full_bar[stage], cb2 * BLOCK_K, m0);
              tma_load_2d(sb, &tm_b2, &full_bar[stage], cb2 * BLOCK_K, n0);
            }
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    // All 32 lanes walk the loop (uniform control flow, operands in uniform registers); one
    // elected lane issues the tcgen05 instructions.
    if (kPair && !leader) {
      // peer CTA: no MMAs to issue.  This warp relays "my epilogue has drained accumulator `as`" to the leader.
      int it = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride, ++it) {
        mbar_wait(&tmem_empty[it & 1], (it >> 1) & 1);
        if (lane == 0) mbar_arrive_remote(mapa_u32(&tmem_empty_peer[it & 1], 0));
        __syncwarp();
      }
    } else {
      constexpr uint32_t idesc = kPair ? make_idesc_mn(256, BLOCK_N) : make_idesc(BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        if (kPair) mbar_wait(&tmem_empty_peer[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_kb_all; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (it == 0 && kb == 0 && lane == 0) stamp(p, 3);
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = make_smem_desc(a_addr);
          const uint64_t bdesc = make_smem_desc(a_addr + A_STAGE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // advance 32 bytes (16 bf16) along K inside the swizzle atom: +2 in the >>4 address field
              if (kPair) umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            // frees the smem slot (in both CTAs of a pair) when these MMAs retire
            if (kPair) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) { if (kPair) umma_commit_pair(&tmem_full[as]); else umma_commit(&tmem_full[as]); }   // accumulator complete -> epilogue(s)
        __syncwarp();
        if (lane == 0) stamp(p, 4);                     // (last tile's value survives)
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..9)
    // Two warpgroups, each with its own named barrier, bias copy, output slots and two residual
    // slots.  Residual tiles are TMA-prefetched two chunks ahead (a whole tile ahead for the
    // usual BLOCK_N = 256), so their HBM latency hides behind the current chunk's work.
    const int eg = (warp - 2) >> 2;                // epilogue group 0 / 1
    const int et = (threadIdx.x - 64) & 127;       // thread within the group
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;              // accumulator row == output pixel within the tile
    float* bias_g = bias_s + eg * BLOCK_N;
    uint8_t* out_g = smem_out + eg * p.out_per_group * OUT_STAGE_BYTES;
    const int R = p.res_slots;
    uint8_t* res_g = smem_res + eg * R * OUT_STAGE_BYTES;
    uint64_t* res_bar = res_full + eg * 4;

    EpiIter<CHUNKS> cur, pre;
    cur.init(first_tile, tile_stride, num_tiles, eg);
    pre = cur;
    auto issue_res = [&](const EpiIter<CHUNKS>& w, int slot) {
      const int mp = w.tile / p.num_n_tiles, n_tile = w.tile - mp * p.num_n_tiles;
      const int m_tile = kPair ? 2 * mp + (int)cta_rank : mp;
      mbar_expect_tx(&res_bar[slot], OUT_STAGE_BYTES);
      uint8_t* dst = res_g + slot * OUT_STAGE_BYTES;
      if (p.pool_t) {
        const int ptiles = p.hw >> 6, pair = m_tile / ptiles, pt_ = m_tile - pair * ptiles;
        const int r0 = pair * 2 * p.hw + pt_ * 64;
        tma_load_2d(dst, &tm_r, &res_bar[slot], n_tile * BLOCK_N + w.chunk * 64, r0);
        tma_load_2d(dst + 64 * 128, &tm_r, &res_bar[slot], n_tile * BLOCK_N + w.chunk * 64, r0 + p.hw);
      } else {
        tma_load_2d(dst, &tm_r, &res_bar[slot], n_tile * BLOCK_N + w.chunk * 64, m_tile * BLOCK_M);
      }
    };
    // Residual layers work IN PLACE: the residual tile of a chunk is TMA-prefetched into one of R slots (R - 1 chunks
    // ahead), the epilogue adds the accumulator into it, the TMA store leaves from the same slot and the slot is refilled
    // once that store has read it.  No separate output staging: the smem buys a third / fourth tile in flight, and a
    // chunk never waits for the previous chunk's store.
    if (has_res && et == 0) {
      for (int j = 0; j < R; ++j)
        if (pre.valid()) { issue_res(pre, j); pre.next(); }
    }
    int n0 = 0;
    long long m0 = 0;
    int as = 0;
    int rslot = 0, prev_rslot = 0;
    uint32_t rphase = 0;
#pragma unroll 1
    for (uint32_t k = 0; cur.valid(); cur.next(), ++k) {
      const int slot = k & 1;
      uint8_t* sout = has_res ? res_g + rslot * OUT_STAGE_BYTES : out_g + (p.out_per_group == 2 ? slot : 0) * OUT_STAGE_BYTES;
      const uint8_t* sres = sout;
      if (cur.first_in_tile()) {
        const int mp = cur.tile / p.num_n_tiles, n_tile = cur.tile - mp * p.num_n_tiles;
        const int m_tile = kPair ? 2 * mp + (int)cta_rank : mp;
        m0 = p.pool_t ? (long long)m_tile * 64 : (long long)m_tile * BLOCK_M;   // pooled tiles are 64 rows
        n0 = n_tile * BLOCK_N;
        as = cur.it & 1;
        // safe to overwrite bias_g: every thread passed the closing barrier of the previous chunk
        for (int i = et; i < BLOCK_N; i += EPI_THREADS) bias_g[i] = __ldg(p.bias + n0 + i);
        mbar_wait(&tmem_full[as], (cur.it >> 1) & 1);
        tc_fence_after();
      }
      if (!has_res && et == 0) {                   // the store that last read this out slot has drained it
        if (p.out_per_group == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
      }
      epi_bar_sync(eg);                            // (also publishes bias_g)
      uint32_t v[64];
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BLOCK_N + cur.chunk * 64;
      TMEM_LD_32x32b_x32(taddr, v);
      TMEM_LD_32x32b_x32(taddr + 32, (v + 32));
      if (has_res) mbar_wait(&res_bar[rslot], rphase);
      tmem_ld_wait();
      if (cur.last_in_tile()) {                    // this group's last TMEM read of the accumulator
        tc_fence_before();
        mbar_arrive(&tmem_empty[as]);
      }
      const int cbase = cur.chunk * 64;
#pragma unroll
      for (int q = 0; q < 8; ++q) {                // 8 x 16 bytes of output per row
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]) + bias_g[cbase + q * 8 + e];
        const int off = row * 128 + ((q ^ (row & 7)) << 4);
        if (has_res) {
          const uint4 t = *reinterpret_cast<const uint4*>(sres + off);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            f[2 * e] += __low2float(h2[e]);
            f[2 * e + 1] += __high2float(h2[e]);
          }
        }
        uint4 o;
        uint32_t* o2 = reinterpret_cast<uint32_t*>(&o);
        if (p.relu) {                              // ReLU inside the conversion (F2FP...RELU): no FMNMX per element
#pragma unroll
          for (int e = 0; e < 4; ++e) o2[e] = pack_bf16x2_relu(f[2 * e], f[2 * e + 1]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) o2[e] = pack_bf16x2(f[2 * e], f[2 * e + 1]);
        }
        *reinterpret_cast<uint4*>(sout + off) = o;
      }
      if (p.pool_t) {
        // MaxPool3d k=s=[2,1,1] (pathway0_pool): rows r and r+64 are the same pixel in frames 2j, 2j+1
        epi_bar_sync(eg);
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2) {
          const int item = et + i2 * EPI_THREADS, r2 = item >> 3, q2 = item & 7;
          uint8_t* pa = sout + r2 * 128 + ((q2 ^ (r2 & 7)) << 4);
          uint4 a4 = *reinterpret_cast<uint4*>(pa);
          const uint4 b4 = *reinterpret_cast<const uint4*>(pa + 64 * 128);     // (r2+64)&7 == r2&7
          __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&a4);
          const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b4);
#pragma unroll
          for (int e = 0; e < 4; ++e) a2[e] = __hmax2(a2[e], b2[e]);
          *reinterpret_cast<uint4*>(pa) = a4;
        }
      }
      fence_proxy_async_smem();
      epi_bar_sync(eg);                            // out tile complete; residual slot fully consumed
      if (et == 0) {
        tma_store_2d(&tm_y, sout, n0 + cbase, (int)m0);
        tma_store_commit();
        if (has_res && k >= 1) {                   // the PREVIOUS chunk's store has read its slot: refill that one
          tma_store_wait_read<1>();
          if (pre.valid()) { issue_res(pre, prev_rslot); pre.next(); }
        }
      }
      if (has_res) {
        prev_rslot = rslot;
        if (++rslot == R) { rslot = 0; rphase ^= 1; }
      }
    }
    if (et == 0 && eg == 0) stamp(p, 5);
    if (et == 0) tma_store_wait<0>();
    if (et == 0 && eg == 0) stamp(p, 6);
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();     // pair: neither CTA leaves while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
  if (threadIdx.x == 0) stamp(p, 7);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;
int g_num_sms = 0;
int g_driver_version = 0;
bool g_corner_dhw = false;    // AFB200_IM2COL_CORNERS=dhw flips the corner array order (bring-up knob)

int g_max_smem = 0;
int g_pair_mode = -1;         // af_set_global_option("pair"): 0 never / 1 wherever possible / -1 planner's choice
int g_force_block_n = 0;      // af_set_global_option("block_n"): 64 / 128 / 256 forces the tile width where Cout allows (tests)
unsigned long long* g_timeline = nullptr;      // AFB200_TIMELINE=1
struct TimelineInfo { int grid, tiles, kblocks, bn; };
TimelineInfo g_timeline_info[256];
long long g_timeline_n = 0;

template <int BLOCK_N, bool kPair>
int launch_t(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ty, const CUtensorMap& tr,
             const CUtensorMap& ta2, const CUtensorMap& tb2, UmmaParams up, cudaStream_t s) {
  static bool configured[64] = {};            // the opt-in shared-memory size is a per-device function attribute
  auto kern = conv_umma_kernel<BLOCK_N, kPair>;
  int dev = 0;
  AFB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    AFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  // pair mode: tiles are 256 rows tall, one 2-CTA cluster (the two SMs of a TPC) per tile
  const int tiles = (kPair ? (up.num_m_tiles + 1) / 2 : up.num_m_tiles) * up.num_n_tiles;
  const int grid = kPair ? 2 * limit_grid(tiles, g_num_sms / 2) : limit_grid(tiles, g_num_sms);
  // shared-memory budget: residual layers trade operand stages for 2 x (3 or 4) residual slots that double as output
  // staging.  Small launches (batch 1-2: one or two tiles per CTA) never cycle through that many slots, and their K
  // loops run at "ring depth per L2 round trip" (measured 160 ns per 64-channel block with 6 slots of operands in
  // flight), so there the staging shrinks to what a CTA's few chunks can use and the ring takes the rest.
  const int stage_bytes = A_STAGE_BYTES + (kPair ? BLOCK_N / 2 : BLOCK_N) * BLOCK_K * 2;
  const bool has_res = up.res != nullptr;
  const int ctas = kPair ? grid / 2 : grid;
  const int tiles_per_cta = (tiles + ctas - 1) / ctas;
  const int chunks_per_group = (tiles_per_cta * (BLOCK_N / 64) + 1) / 2;
  up.out_per_group = has_res ? 0 : ((BLOCK_N <= 128 && chunks_per_group >= 2) ? 2 : 1);
  static const char* rs256 = getenv("AFB200_RES_SLOTS_256");
  const int res_default = BLOCK_N <= 128 ? 4 : (rs256 ? atoi(rs256) : 3);
  up.res_slots = has_res ? (chunks_per_group < 2 ? 2 : chunks_per_group < res_default ? chunks_per_group : res_default) : 0;
  const int fixed = (2 * up.out_per_group + 2 * up.res_slots) * OUT_STAGE_BYTES + 2 * BLOCK_N * 4 +
                    (2 * MAX_STAGES + 14) * 8 + 16 + 1024;
  int stages = (g_max_smem - fixed) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) { set_error("conv_umma: shared memory budget too small"); return AF_ERR_INVALID; }
  up.stages = stages;
  const int dyn = fixed + stages * stage_bytes;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = dyn; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = kPair ? 2 : 1;
  up.dbg = nullptr;
  if (g_timeline) {                      // ring of 256 launches x 8 stamps, dumped by conv_umma_timeline_dump()
    up.dbg = g_timeline + (size_t)(g_timeline_n % 256) * 8;
    g_timeline_info[g_timeline_n % 256] = {grid, tiles, up.kt * up.kh * up.kw * (up.Cin / BLOCK_K) + up.cblocks2, BLOCK_N};
    ++g_timeline_n;
  }
  AFB_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, ty, tr, ta2, tb2, up));
  ++g_launches;
  AFB_CUDA(cudaGetLastError());
  return AF_OK;
}

int encode_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, const char* what) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {BLOCK_K, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s: rows=%llu cols=%llu box_rows=%u) failed: %d", what,
              (unsigned long long)rows, (unsigned long long)cols, box_rows, (int)r);
    return AF_ERR_CUDA;
  }
  return AF_OK;
}


// im2col-mode map over a dense NDHWC tensor: 128 output pixels x 64 channels per box; the hardware applies
// the conv stride, the tap offset and the zero halo.
int encode_im2col(CUtensorMap* map, const void* x, int B, int Ti, int Hi, int Wi, int Cin, int kt, int kh, int kw,
                  int st, int sh, int sw, int pt, int ph, int pw) {
  cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)Ti, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)Wi * Cin * 2, (cuuint64_t)Hi * Wi * Cin * 2,
                           (cuuint64_t)Ti * Hi * Wi * Cin * 2};
  int lower[3] = {-pw, -ph, -pt};
  int upper[3] = {pw - (kw - 1), ph - (kh - 1), pt - (kt - 1)};
  if (g_corner_dhw) { int t0 = lower[0]; lower[0] = lower[2]; lower[2] = t0; t0 = upper[0]; upper[0] = upper[2]; upper[2] = t0; }
  cuuint32_t es[5] = {1, (cuuint32_t)sw, (cuuint32_t)sh, (cuuint32_t)st, 1};
  CUresult r = g_encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, lower,
                               upper, BLOCK_K, BLOCK_M, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed: %d (C=%d W=%d H=%d T=%d B=%d k=%dx%dx%d)", (int)r, Cin, Wi, Hi, Ti, B, kt,
              kh, kw);
    return AF_ERR_CUDA;
  }
  // CUTLASS (copy_traits_sm90_im2col.hpp) clears bit 21 of descriptor word 1 for tensors
  // smaller than 128 KiB on drivers <= 13.1 to work around a driver encoding issue.
  const unsigned long long bytes = (unsigned long long)B * Ti * Hi * Wi * Cin * 2ULL;
  if (g_driver_version <= 13010 && bytes < 131072ULL) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ULL << 21);
  return AF_OK;
}

}  // namespace

void conv_umma_force_block_n(int bn) { g_force_block_n = bn; }
void conv_umma_set_pair_mode(int mode) { g_pair_mode = mode; }

int conv_umma_init() {
  if (g_encode_tiled && g_encode_im2col) return AF_OK;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  AFB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available in this driver"); return AF_ERR_UNSUPPORTED; }
  g_encode_tiled = (EncodeTiledFn)fn;
  fn = nullptr;
  AFB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeIm2col not available in this driver"); return AF_ERR_UNSUPPORTED; }
  g_encode_im2col = (EncodeIm2colFn)fn;
  int dev = 0;
  AFB_CUDA(cudaGetDevice(&dev));
  AFB_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  AFB_CUDA(cudaDriverGetVersion(&g_driver_version));
  AFB_CUDA(cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const char* tl = getenv("AFB200_TIMELINE");
  if (tl && tl[0] == '1' && !g_timeline) {
    AFB_CUDA(cudaMalloc(&g_timeline, 256 * 8 * sizeof(unsigned long long)));
    AFB_CUDA(cudaMemset(g_timeline, 0, 256 * 8 * sizeof(unsigned long long)));
  }
  const char* c = getenv("AFB200_IM2COL_CORNERS");
  g_corner_dhw = c && c[0] == 'd';
  return AF_OK;
}

// AFB200_TIMELINE=1: print, for the last `n` conv_umma launches, CTA 0's phase boundaries in microseconds relative to
// the previous launch's end: entry, prologue done, prior grid done, first operands landed, last MMA issued, last store
// issued, stores drained, exit.
void conv_umma_timeline_dump(int n) {
  if (!g_timeline) return;
  cudaDeviceSynchronize();
  static unsigned long long host[256 * 8];
  cudaMemcpy(host, g_timeline, sizeof(host), cudaMemcpyDeviceToHost);
  long long first = g_timeline_n - n < 0 ? 0 : g_timeline_n - n;
  unsigned long long prev_end = 0;
  for (long long i = first; i < g_timeline_n; ++i) {
    const unsigned long long* t = host + (size_t)(i % 256) * 8;
    const TimelineInfo& f = g_timeline_info[i % 256];
    const double base = prev_end ? (double)prev_end : (double)t[0];
    fprintf(stderr, "[timeline] bn=%3d grid=%3d tiles=%4d kb=%3d | entry %+7.2f prol %+7.2f dep %+7.2f data %+7.2f mma %+7.2f st %+7.2f drain %+7.2f exit %+7.2f | dur %6.2f us\n",
            f.bn, f.grid, f.tiles, f.kblocks, (t[0] - base) / 1e3, (t[1] - base) / 1e3, (t[2] - base) / 1e3, (t[3] - base) / 1e3,
            (t[4] - base) / 1e3, (t[5] - base) / 1e3, (t[6] - base) / 1e3, (t[7] - base) / 1e3, (t[7] - t[0]) / 1e3);
    prev_end = t[7];
  }
}

bool conv_umma_supported(const ConvProblem& p) {
  if (!g_encode_tiled) return false;
  if (p.Cin % BLOCK_K != 0 || p.Cout % 64 != 0) return false;
  // dense NDHWC input only
  if (p.xsW != p.Cin || p.xsH != (long long)p.Wi * p.Cin || p.xsT != (long long)p.Hi * p.Wi * p.Cin ||
      p.xsB != (long long)p.Ti * p.Hi * p.Wi * p.Cin)
    return false;
  if (p.kt > 16 || p.kh > 16 || p.kw > 16) return false;
  if (p.M >= (1LL << 31)) return false;
  if (p.pool_t) {
    const bool pointwise = p.kt == 1 && p.kh == 1 && p.kw == 1 && p.st == 1 && p.sh == 1 && p.sw == 1;
    if (!pointwise || !p.relu || (p.To & 1) || ((p.Ho * p.Wo) & 63)) return false;
  }
  if (p.x2) {      // fused projection shortcut: same output grid through a pointwise stride-[1,sh2,sw2] conv of x2
    if (p.pool_t || p.res || !p.w2 || p.Cin2 <= 0 || p.Cin2 % BLOCK_K != 0) return false;
    if (p.T2 != p.To || (p.H2 - 1) / p.sh2 + 1 != p.Ho || (p.W2 - 1) / p.sw2 + 1 != p.Wo) return false;
  }
  return true;
}

namespace {
// Everything a launch needs that depends only on the problem (pointers, shapes): tile width, kernel
// parameters and the four tensor maps.  The engine's buffers are fixed, so plans are memoised — encoding four
// tensor maps per conv is most of the host cost of a batch-1 pass.
struct UmmaPlan {
  alignas(64) CUtensorMap ta, tb, ty, tr, ta2, tb2;
  UmmaParams up;
  int bn;
  int pair;      // 2-CTA clusters, cta_group::2 MMAs (weight tiles split between the CTAs of a pair)
};
struct PlanKey {
  const void *x, *w, *y, *res, *x2, *w2; const float* bias;
  int v[28];
  bool operator==(const PlanKey& o) const { return memcmp(this, &o, sizeof(PlanKey)) == 0; }
};
struct PlanKeyHash {
  size_t operator()(const PlanKey& k) const {
    const uint64_t* q = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < sizeof(PlanKey) / 8; ++i) { h ^= q[i]; h *= 1099511628211ULL; }
    return (size_t)h;
  }
};
thread_local std::unordered_map<PlanKey, UmmaPlan, PlanKeyHash> g_plans;

int build_plan(const ConvProblem& p, UmmaPlan& plan);
}  // namespace

int conv_umma_launch(const ConvProblem& p, cudaStream_t s) {
  PlanKey key;
  memset(&key, 0, sizeof(key));
  key.x = p.x; key.w = p.w; key.y = p.y; key.res = p.res; key.bias = p.bias; key.x2 = p.x2; key.w2 = p.w2;
  const int vals[28] = {g_pair_mode, g_force_block_n, p.B, p.Ti, p.Hi, p.Wi, p.Cin, p.To, p.Ho, p.Wo, p.Cout, p.kt, p.kh, p.kw, p.st, p.sh, p.sw,
                        p.pt, p.ph, p.pw, p.relu, p.pool_t, p.Cin2, p.T2, p.H2, p.W2, p.sh2, p.sw2};
  memcpy(key.v, vals, sizeof(vals));
  auto it = g_plans.find(key);
  if (it == g_plans.end()) {
    if (g_plans.size() > 4096) g_plans.clear();
    UmmaPlan plan;
    int rc = build_plan(p, plan);
    if (rc) return rc;
    it = g_plans.emplace(key, plan).first;
  }
  const UmmaPlan& pl = it->second;
  if (pl.pair) {
    if (pl.bn == 256) return launch_t<256, true>(pl.ta, pl.tb, pl.ty, pl.tr, pl.ta2, pl.tb2, pl.up, s);
    return launch_t<128, true>(pl.ta, pl.tb, pl.ty, pl.tr, pl.ta2, pl.tb2, pl.up, s);
  }
  switch (pl.bn) {
    case 256: return launch_t<256, false>(pl.ta, pl.tb, pl.ty, pl.tr, pl.ta2, pl.tb2, pl.up, s);
    case 128: return launch_t<128, false>(pl.ta, pl.tb, pl.ty, pl.tr, pl.ta2, pl.tb2, pl.up, s);
    default: return launch_t<64, false>(pl.ta, pl.tb, pl.ty, pl.tr, pl.ta2, pl.tb2, pl.up, s);
  }
}

namespace {
int build_plan(const ConvProblem& p, UmmaPlan& plan) {
  UmmaParams& up = plan.up;
  CUtensorMap &ta = plan.ta, &tb = plan.tb, &ty = plan.ty, &tr = plan.tr;
  up.bias = p.bias; up.res = (const bf16*)p.res; up.M = p.M; up.Cout = p.Cout; up.Cin = p.Cin;
  up.kt = p.kt; up.kh = p.kh; up.kw = p.kw; up.st = p.st; up.sh = p.sh; up.sw = p.sw;
  up.pt = p.pt; up.ph = p.ph; up.pw = p.pw; up.To = p.To; up.Ho = p.Ho; up.Wo = p.Wo;
  up.relu = p.relu;
  up.pool_t = p.pool_t; up.hw = p.Ho * p.Wo;
  const bool pointwise = p.kt == 1 && p.kh == 1 && p.kw == 1 && p.st == 1 && p.sh == 1 && p.sw == 1;
  static const char* force = getenv("AFB200_FORCE_IM2COL");
  up.im2col = (pointwise && !(force && force[0] == '1')) ? 0 : 1;
  up.num_m_tiles = (int)((p.M + BLOCK_M - 1) / BLOCK_M);      // pool_t: M/128 tiles of (2 frames x 64 pixels) too

  // tile width: the widest of 256/128/64 that still gives every SM at least ~2 tiles
  int bn = 64;
  for (int cand : {256, 128}) {
    if (p.Cout % cand == 0 && (long long)up.num_m_tiles * (p.Cout / cand) >= 2LL * g_num_sms) { bn = cand; break; }
  }
  // Residual layers with a very short K (s2 / s3 `c` convs: K = 64 / 128) are pure HBM streams: narrower tiles give
  // them two output slots per epilogue group and a deeper operand ring (measured on B200, 32 clips: K=64 5.4 -> 6.5 TB/s
  // at BLOCK_N 64, K=128 5.45 -> 5.7 TB/s at BLOCK_N 128; wider K is better off at 256)
  if (p.res && !p.x2 && pointwise) {
    static const char* rk = getenv("AFB200_RES_BN128_MAXK");
    const int maxk128 = rk ? atoi(rk) : 128;
    if (p.Cin <= 64) bn = 64;
    else if (p.Cin <= maxk128 && bn > 128) bn = 128;
  }
  static const char* fbn = getenv("AFB200_BLOCK_N");
  int force_bn = g_force_block_n ? g_force_block_n : (fbn ? atoi(fbn) : 0);
  if ((force_bn == 64 || force_bn == 128 || force_bn == 256) && p.Cout % force_bn == 0) bn = force_bn;
  up.num_n_tiles = p.Cout / bn;
  // CTA pairs wherever the weight tile can be split in two halves of >= 64 rows (g_pair_mode: 0 never, 1 always,
  // -1 planner's choice)
  static const char* fpair = getenv("AFB200_PAIR");
  const int pair_mode = g_pair_mode >= 0 ? g_pair_mode : (fpair ? atoi(fpair) : -1);
  // (measured on B200, 32 clips: pairs gain 5-20 % on every K >= 256 layer; the K <= 128 HBM streams of s2 lose 10 %)
  const int k_total = p.kt * p.kh * p.kw * p.Cin + (p.x2 ? p.Cin2 : 0);
  plan.pair = (bn >= 128 && g_num_sms >= 2 && pair_mode != 0 && (pair_mode == 1 || k_total > 128)) ? 1 : 0;
  const uint32_t b_box_rows = (uint32_t)(plan.pair ? bn / 2 : bn);

  if (up.im2col) {
    int rc = encode_im2col(&ta, p.x, p.B, p.Ti, p.Hi, p.Wi, p.Cin, p.kt, p.kh, p.kw, p.st, p.sh, p.sw, p.pt, p.ph, p.pw);
    if (rc) return rc;
  } else {
    int rc = encode_2d(&ta, p.x, (uint64_t)p.M, (uint64_t)p.Cin, p.pool_t ? 64 : BLOCK_M, "A");
    if (rc) return rc;
  }
  const int taps = p.kt * p.kh * p.kw;
  int rc = encode_2d(&tb, p.w, (uint64_t)taps * p.Cout, (uint64_t)p.Cin, b_box_rows, "W");
  if (rc) return rc;
  if (p.pool_t) {       // pooled output has M/2 rows; every box is 64 rows tall
    rc = encode_2d(&ty, p.y, (uint64_t)p.M / 2, (uint64_t)p.Cout, 64, "Y(pooled)");
    if (rc) return rc;
    rc = encode_2d(&tr, p.res ? p.res : p.x, (uint64_t)p.M, (uint64_t)(p.res ? p.Cout : p.Cin), 64, "R");
    if (rc) return rc;
  } else {
    rc = encode_2d(&ty, p.y, (uint64_t)p.M, (uint64_t)p.Cout, BLOCK_M, "Y");
    if (rc) return rc;
    rc = encode_2d(&tr, p.res ? p.res : p.y, (uint64_t)p.M, (uint64_t)p.Cout, BLOCK_M, "R");
    if (rc) return rc;
  }

  // fused projection shortcut: pointwise conv of stride [1,sh2,sw2] over x2 [B,T2,H2,W2,Cin2], weights w2 [Cout][Cin2]
  up.cblocks2 = 0; up.im2col2 = 0; up.sh2 = up.sw2 = 1;
  plan.ta2 = plan.ta; plan.tb2 = plan.tb;
  if (p.x2) {
    up.cblocks2 = p.Cin2 / BLOCK_K;
    up.sh2 = p.sh2; up.sw2 = p.sw2;
    up.im2col2 = (p.sh2 != 1 || p.sw2 != 1) ? 1 : 0;
    if (up.im2col2) rc = encode_im2col(&plan.ta2, p.x2, p.B, p.T2, p.H2, p.W2, p.Cin2, 1, 1, 1, 1, p.sh2, p.sw2, 0, 0, 0);
    else rc = encode_2d(&plan.ta2, p.x2, (uint64_t)p.M, (uint64_t)p.Cin2, BLOCK_M, "A2");
    if (rc) return rc;
    rc = encode_2d(&plan.tb2, p.w2, (uint64_t)p.Cout, (uint64_t)p.Cin2, b_box_rows, "W2");
    if (rc) return rc;
  }

  plan.bn = bn;
  return AF_OK;
}
}  // namespace

}  // namespace afb
