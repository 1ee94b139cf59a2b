// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core conv kernels (sm_100a).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace afb {
namespace {

constexpr int UMMA_BLOCK_M = 128;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Dynamic shared memory rounded up to 1024 bytes (SWIZZLE_128B atoms) WITHOUT a round trip through an integer: the
// result stays a pointer the compiler can prove to be in the shared window, so the epilogues' loads / stores compile
// to LDS / STS (32-bit addresses, short scoreboard) instead of generic LD / ST (64-bit address arithmetic, long scoreboard).
__device__ __forceinline__ uint8_t* align_smem_1024(uint8_t* raw) {
  return raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(raw) & 1023u)) & 1023u);
}

// One lane of a fully converged warp (PTX elect.sync): lets ptxas emit the warp-level tcgen05/TMA
// instructions once instead of a per-active-thread waterfall loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// Programmatic dependent launch: let the next kernel of the stream start its prologue while this one
// drains, and do not touch global memory before the previous kernel has completed.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint expires)
// instead of re-issuing the poll every few dozen clocks -- ten polling warps per SM cost issue slots and, on a
// power-capped part, clock speed.
#ifndef AFB_MBAR_SUSPEND_NS
#define AFB_MBAR_SUSPEND_NS 20000
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"((uint32_t)AFB_MBAR_SUSPEND_NS)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// named barrier among the 128 threads of one epilogue warpgroup (id 1 or 2)
__device__ __forceinline__ void epi_bar_sync(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c, int w, int h,
                                                   int d, int n, uint16_t ow, uint16_t oh, uint16_t od) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2], {%8, %9, %10};" ::"r"(smem_u32(dst)),
      "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "h"(ow), "h"(oh), "h"(od)
      : "memory");
}
// L2 prefetch of a tile (no shared memory, no barrier): issued a work unit ahead so that the later TMA load of the same box
// finds its lines in L2 -- with only two or three ring slots per CTA the HBM latency of a box (1.5-2 us under load) is
// otherwise longer than the MMAs of the boxes in flight.
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"((uint64_t)m), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_5d(const CUtensorMap* m, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];" ::"l"((uint64_t)m), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)m),
               "r"(smem_u32(src)), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// K-major, 128B-swizzled shared-memory matrix descriptor (rows of 128 bytes, 8-row groups
// 1024 bytes apart). Encoding per the PTX ISA "shared memory descriptor" / CUTLASS
// cute/arch/mma_sm100_desc.hpp: addr>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
// General form: `sbo` = byte distance between consecutive 8-row groups, `layout` = 2 (SWIZZLE_128B,
// 128-byte rows) or 4 (SWIZZLE_64B, 64-byte rows).
__device__ __forceinline__ uint64_t make_smem_desc_ex(uint32_t addr, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// Measured on B200 (round 1): an operand may start at any 128-byte multiple inside a TMA-written SWIZZLE_128B
// box, with SBO not a multiple of 1024 (e.g. a 10-pixel-wide halo box, tile starting dy rows / dx pixels in):
// results are exact with the base-offset field (bits [49,52)) left 0 — the hardware de-swizzles on absolute
// shared-memory address bits [7,10), exactly as TMA swizzled when writing.  Setting base offset = (addr>>7)&7
// on top of that gives wrong results.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BLOCK_N.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(UMMA_BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// A operand from tensor memory (lane = row, one 32-bit column = two consecutive K elements, even K in the low half;
// a K = 16 step is 8 columns), B from shared memory: cute SM100_MMA_F16BF16_TS.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of a 2-CTA cluster (the two SMs of a TPC) run ONE 256-row MMA: each CTA holds its own 128 rows of A and
// HALF of the B tile in its shared memory and its own 128 accumulator rows in its TMEM; the leader (cluster rank 0)
// issues the instruction and both CTAs' operands are read at the same shared-memory offsets.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p`'s counterpart in the CTA of cluster rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: the data lands in THIS CTA's shared memory, the transaction bytes are counted on the
// barrier at cluster address `bar_cluster` (the leader's full barrier).
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"((uint64_t)m), "r"(bar_cluster), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_5d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c, int w,
                                                        int h, int d, int n, uint16_t ow, uint16_t oh, uint16_t od) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6, %7}], [%2], {%8, %9, %10};" ::"r"(smem_u32(dst)),
      "l"((uint64_t)m), "r"(bar_cluster), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "h"(ow), "h"(oh), "h"(od)
      : "memory");
}
// kind::f16 instruction descriptor with an explicit M (256 for cta_group::2)
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrives (once the pair's MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

#define TMEM_LD_32x32b_x32(taddr, r)                                                                               \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                    \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                    \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),      \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),     \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                   \
      : "r"(taddr)                                                                                                 \
      : "memory")

#define TMEM_LD_32x32b_x16(taddr, r)                                                                               \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                                    \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                             \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                    \
      : "r"(taddr)                                                                                                 \
      : "memory")

// Two fp32 -> packed bf16x2 (round to nearest even), `lo` in the low half.  The _relu form clamps negatives to zero inside
// the same conversion instruction (F2FP...RELU): relu(round(x)) == round(relu(x)), so the epilogues need no FMNMX.
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// registers -> TMEM (own lane quadrant, 32 consecutive columns): used to park finished bf16x2 outputs in spare columns
#define TMEM_ST_32x32b_x32(taddr, r)                                                                               \
  asm volatile(                                                                                                    \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                              \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                                   \
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),             \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), \
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),  \
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),  \
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                               \
      : "memory")
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }


}  // namespace
}  // namespace afb
