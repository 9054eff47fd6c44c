// tc05.cuh -- thin PTX wrappers for the sm_100a tensor path: tcgen05.mma
// (kind::tf32, accumulators in TMEM), TMEM allocation / loads, MMA-completion
// commits onto mbarriers, and the shared-memory / instruction descriptors.
//
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" /
// "instruction descriptor" tables (the same fields CUTLASS names
// UMMA::SmemDescriptor and UMMA::InstrDescriptor).
#pragma once

#include "common.cuh"

namespace pmb {
namespace tc {

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (64 bit):
//   [0,14)  start address >> 4        [16,30) leading-dimension byte offset >> 4
//   [32,46) stride-dimension byte offset >> 4     [46,48) version = 1 (sm_100)
//   [49,52) base offset (0: atoms are 1024 B aligned)   [61,64) layout type
enum : uint64_t { kLayoutNone = 0, kLayoutSw128Base32 = 1, kLayoutSw128 = 2, kLayoutSw64 = 4, kLayoutSw32 = 6 };

__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint64_t layout) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (layout << 61);
}

// Instruction descriptor (32 bit) for kind::tf32, fp32 accumulate:
//   [4,6) D format (1 = F32)  [7,10) A format (2 = TF32)  [10,13) B format (2 = TF32)
//   [15] A major (0 = K, 1 = MN)  [16] B major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- TMEM
// One warp allocates `cols` (power of two >= 32) columns; the base address lands in *slot (smem).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.  accumulate == 0 overwrites D.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-convergent variants: EVERY lane of the warp executes the statement, one elected lane issues.
// Keeping the issue out of an `if (lane == 0)` branch matters: under divergence ptxas wraps each
// uniform-datapath instruction (UTCHMMA / UTCBAR) in an ELECT ... BRA.U.ANY loop, and the MMA warp
// becomes issue-bound.
__device__ __forceinline__ void mma_tf32_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(smem_u32(bar))
      : "memory");
}
// All MMAs issued so far by this thread arrive (once) on the mbarrier when they complete.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base + i).
// taddr = tmem_base + ((lane_base) << 16) + column, lane_base = 32 * (warp_id % 4).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Same load without the wait: issue several, then tmem_ld_wait() once (tcgen05.wait::ld covers every
// outstanding load of the thread).  The destination registers must not be read before the wait.
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mbarrier arrive (count 1) by a thread of this CTA
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of a cluster on the two SMs of a TPC execute ONE tcgen05.mma: M = 256 rows (each CTA holds 128
// rows of A in its own shared memory and 128 lanes of D in its own TMEM), B is split along N (each CTA
// holds N / 2 columns).  Descriptors and TMEM addresses are the same numbers in both CTAs, so both must
// use identical shared-memory layouts.  Only the CTA of cluster rank 0 issues the instruction.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// executed by the warp of the same index in BOTH CTAs of the pair, same slot address
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void mma_tf32_pair_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                    uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// when the pair's MMAs issued so far complete, arrive (once) on the mbarrier at this shared-memory offset in
// every CTA of `cta_mask`
__device__ __forceinline__ void mma_commit_pair_elect(uint64_t* bar, uint32_t cta_mask) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      ".reg .b16 lo, hi;\n\t"
      "mov.b32 {lo, hi}, %1;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], lo;\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(cta_mask)
      : "memory");
}
// arrive (count 1, release at cluster scope) on the mbarrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// wait with acquire at cluster scope (the arrivals come from both CTAs of the pair)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}

// ---------------------------------------------------------------- operand layouts (byte offsets)
// K-major, no swizzle: 8-row x 16-byte core matrices (128 contiguous bytes);
// core matrices adjacent in K are `lbo` apart, adjacent in M/N `sbo` apart.
__device__ __forceinline__ uint32_t off_kmajor(int row, int k /*fp32 index*/, uint32_t lbo, uint32_t sbo) {
  return (uint32_t)(row >> 3) * sbo + (uint32_t)(k >> 2) * lbo + (uint32_t)(row & 7) * 16u + (uint32_t)(k & 3) * 4u;
}
// MN-major fp32/tf32 operands: the only layout the tensor core accepts is the 128-byte swizzle with
// a 32-byte base (layout type 1): an atom is 4 k-rows of 128 B (32 fp32 along M/N); the 32-byte chunk
// index (address bits 5-6) is XOR-ed with the k-row index (address bits 7-8).  Atoms adjacent in M/N
// are `lbo` apart, adjacent in K (4 rows) `sbo` apart; one K = 8 instruction spans two atoms in K.
__device__ __forceinline__ uint32_t off_mnmajor_sw128b32(int mn, int k, uint32_t lbo, uint32_t sbo) {
  const uint32_t k4 = (uint32_t)k & 3u, chunk = ((uint32_t)mn & 31u) >> 3;
  return ((uint32_t)mn >> 5) * lbo + ((uint32_t)k >> 2) * sbo + k4 * 128u + ((chunk ^ k4) << 5) + ((uint32_t)mn & 7u) * 4u;
}

}  // namespace tc
}  // namespace pmb
