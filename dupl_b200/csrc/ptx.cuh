// Thin inline-PTX wrappers for the sm_100a features the hot path uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and proxy fences.
// Bit layouts of the UMMA shared-memory and instruction descriptors follow the PTX ISA
// (cross-checked against cute/arch/mma_sm100_desc.hpp shipped with CUTLASS 4.5).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dupl {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
// 2-D tiled load global -> shared, completion signalled on `bar` (complete_tx::bytes).
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 3-D tiled load (innermost coordinate first).
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// ---------------------------------------------------------------- CTA pairs (cta_group::2)
// shared::cluster address of `p` (a shared-memory location of the calling CTA) in CTA `rank` of the cluster.
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of one CTA of a pair: data lands in this CTA's shared memory, the bytes are accounted on the
// mbarrier at shared::cluster address `bar_cluster_addr` (the leader CTA's barrier).
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 256 x N x 16 MMA executed by the two SMs of a CTA pair (each holds 128 rows of A and N/2 rows of B in
// its own shared memory, and 128 rows of the accumulator in its own TMEM).  Issued by the leader CTA only.
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Commit of the pair's MMAs, arriving on the barrier at this offset in both CTAs.
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t.reg .b16 m;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b16 m, 3;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}
// ---------------------------------------------------------------- clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Make generic-proxy writes to shared memory visible to the async proxy (UMMA / TMA reads).
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// Arrive on an mbarrier once every tcgen05 op previously issued by the elected thread has completed.
// Warp-collective like tc_mma_f16 (same elected leader).
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16/fp16 inputs, fp32 accumulate.
// Called by ALL lanes of the issuing warp with warp-uniform operands: the leader is elected inside the
// asm block, so the surrounding C++ stays convergent and descriptors live in uniform registers
// (a divergent `if (lane == 0)` around the issue loop costs ~14 SASS instructions per MMA).
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand read from tensor memory (lane = row, two bf16 per 32-bit column, 8 columns per
// UMMA_K = 16): halves the shared-memory bandwidth an MMA needs.
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns store: thread i of the warp writes TMEM lane (base_lane + i), columns c..c+31.
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread i of the warp receives TMEM lane (base_lane + i), columns c..c+31.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 16 columns load / 32 lanes x 8 columns store (a thread of a 16-warp CTA owns 16 columns of its row).
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
      : "memory");
}

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor for a tile stored as rows of 128 bytes with the 128-byte
// swizzle (exactly what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B and a 128-byte inner box):
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (unused here: 1)
//   bits [32,46) stride byte offset >> 4 = 1024 B between 8-row groups
//   bits [46,48) version = 1 (sm_100)       bits [61,64) layout = 2 (SWIZZLE_128B)
// The same descriptor serves K-major operands (contraction along the 128-byte row; advance
// the start address by 32 B per UMMA_K=16 bf16) and MN-major operands (contraction along rows;
// advance by 2048 B = 16 rows per UMMA_K); which one is meant is selected in the instruction
// descriptor's a_major / b_major bits.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return static_cast<uint64_t>((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major operand wider than one 64-element block: block j (elements 64j .. 64j+63 of the M / N dimension) is its own
// [k rows x 128 B] region `block_bytes` after block j-1 — the descriptor's leading byte offset; 8-row groups along k stay
// 1024 B apart (canonical layout ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in 16-byte units).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t saddr, uint32_t block_bytes) {
  return static_cast<uint64_t>((saddr >> 4) & 0x3FFF) | (static_cast<uint64_t>((block_bytes >> 4) & 0x3FFF) << 16) | (64ull << 32) |
         (1ull << 46) | (2ull << 61);
}
// Same for rows of 64 bytes with the 64-byte swizzle (K-major only): 512 B between 8-row groups, layout = 4.
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
  return static_cast<uint64_t>((saddr >> 4) & 0x3FFF) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// Descriptor of the same tile `bytes` further on (no carry out of the 14-bit address field within 227 KB).
__device__ __forceinline__ uint64_t umma_desc_advance(uint64_t desc, uint32_t bytes) { return desc + (bytes >> 4); }
// Instruction descriptor, kind::f16: bf16 x bf16 -> f32, M = 128.
//   [4,6) c_format F32=1   [7,10) a_format BF16=1   [10,13) b_format BF16=1
//   [15] a_major  [16] b_major (0 = K-major, 1 = MN-major)   [17,23) N>>3   [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n, int a_mn_major, int b_mn_major, int m = 128) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ---------------------------------------------------------------- batched issue of a 3-pass split product
// The 12 MMAs of one 3-pass split product (64-deep contraction = 4 k-steps per pass) as ONE asm block: one elect.sync for the
// batch and the descriptor increments as immediates.  Issued one C++ call per MMA (elect + setp + two 64-bit adds each, every
// MMA's predicate depending on its own elect), a single warp got out one 128x64x16 MMA per ~45 cycles — more than the 32-48
// cycles the tensor pipe needs for it, so the ISSUING warp bounded both backward kernels (ncu: 34-40 % tensor-pipe activity,
// arithmetic warps waiting on the S / dP barrier).
#define DUPL_MMA_SS(A, B, KA, KB, ACC)                                                                     \
  "add.u64 ta, " A ", " #KA ";\n\tadd.u64 tb, " B ", " KB ";\n\t"                                            \
  "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ta, tb, %5, " ACC ";\n\t"
#define DUPL_MMA_SS_PASS(A, B, S1, S2, S3, ACC0)                                                           \
  "@e tcgen05.mma.cta_group::1.kind::f16 [%0], " A ", " B ", %5, " ACC0 ";\n\t"                               \
  DUPL_MMA_SS(A, B, 2, S1, "one") DUPL_MMA_SS(A, B, 4, S2, "one") DUPL_MMA_SS(A, B, 6, S3, "one")
// D[tmem] (+)= A[smem] B[smem]^T, 3-pass split; BSTEP = descriptor step of B per k-step in 16-byte units: 2 (K-major B,
// 32 B) or 128 (MN-major B, 2048 B = 16 rows)
template <int BSTEP>
__device__ __forceinline__ void mma_ss_split(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                             uint32_t idesc, bool accumulate) {
  static_assert(BSTEP == 2, "only the K-major B form is instantiated");
  asm volatile(
      "{\n\t.reg .pred e, p, one;\n\t.reg .b64 ta, tb;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.eq.b32 one, 0, 0;\n\t"
      DUPL_MMA_SS_PASS("%1", "%3", "2", "4", "6", "p")     // hi * hi
      DUPL_MMA_SS_PASS("%1", "%4", "2", "4", "6", "one")   // hi * lo
      DUPL_MMA_SS_PASS("%2", "%3", "2", "4", "6", "one")   // lo * hi
      "}"
      ::"r"(d_tmem), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate ? 1u : 0u)
      : "memory");
}
#define DUPL_MMA_TS(AOFF, B, KB, ACC)                                                                      \
  "add.u32 sa, %1, " #AOFF ";\n\tadd.u64 tb, " B ", " KB ";\n\t"                                             \
  "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [sa], tb, %4, " ACC ";\n\t"
// same with A in tensor memory (hi plane at a_tmem, lo plane 32 columns further; 8 columns per k-step); BSTEP = 2 (K-major B)
// or 128 (MN-major B: 2048 B = 16 rows per k-step)
template <int BSTEP>
__device__ __forceinline__ void mma_ts_split(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_hi, uint64_t b_lo, uint32_t idesc,
                                             bool accumulate) {
  static_assert(BSTEP == 2 || BSTEP == 128, "B descriptor step: 2 (K-major) or 128 (MN-major)");
  if (BSTEP == 2) {
    asm volatile(
        "{\n\t.reg .pred e, p, one;\n\t.reg .b64 tb;\n\t.reg .b32 sa;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.b32 one, 0, 0;\n\t"
        DUPL_MMA_TS(0, "%2", "0", "p") DUPL_MMA_TS(8, "%2", "2", "one") DUPL_MMA_TS(16, "%2", "4", "one") DUPL_MMA_TS(24, "%2", "6", "one")      // hi * hi
        DUPL_MMA_TS(0, "%3", "0", "one") DUPL_MMA_TS(8, "%3", "2", "one") DUPL_MMA_TS(16, "%3", "4", "one") DUPL_MMA_TS(24, "%3", "6", "one")  // hi * lo
        DUPL_MMA_TS(32, "%2", "0", "one") DUPL_MMA_TS(40, "%2", "2", "one") DUPL_MMA_TS(48, "%2", "4", "one") DUPL_MMA_TS(56, "%2", "6", "one")  // lo * hi
        "}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate ? 1u : 0u)
        : "memory");
    return;
  }
  asm volatile(
      "{\n\t.reg .pred e, p, one;\n\t.reg .b64 tb;\n\t.reg .b32 sa;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.eq.b32 one, 0, 0;\n\t"
      DUPL_MMA_TS(0, "%2", "0", "p") DUPL_MMA_TS(8, "%2", "128", "one") DUPL_MMA_TS(16, "%2", "256", "one") DUPL_MMA_TS(24, "%2", "384", "one")      // hi * hi
      DUPL_MMA_TS(0, "%3", "0", "one") DUPL_MMA_TS(8, "%3", "128", "one") DUPL_MMA_TS(16, "%3", "256", "one") DUPL_MMA_TS(24, "%3", "384", "one")  // hi * lo
      DUPL_MMA_TS(32, "%2", "0", "one") DUPL_MMA_TS(40, "%2", "128", "one") DUPL_MMA_TS(48, "%2", "256", "one") DUPL_MMA_TS(56, "%2", "384", "one")  // lo * hi
      "}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate ? 1u : 0u)
      : "memory");
}

// ---------------------------------------------------------------- split-bf16 helpers
// x = hi + lo + O(2^-17 |x|): two bf16 planes carry ~16 mantissa bits through the tensor core.
// Experiment builds only (make emu, tools/precision_table.py): -DDUPL_EMU_TF32 rounds every value to TF32 (10 explicit
// mantissa bits, cvt.rna) BEFORE it is split, so that the 3-pass product of the planes equals a single-pass kind::tf32 MMA
// on RN-rounded operands to 2^-17.  The product library is built without it.
__device__ __forceinline__ float emu_operand(float x) {
#ifdef DUPL_EMU_TF32
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
#else
  return x;
#endif
}
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  x = emu_operand(x);
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// Two values at once: one cvt.rn.bf16x2.f32 per plane.  Element a goes to the low half (lower address).
__device__ __forceinline__ void split2_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
  a = emu_operand(a);
  b = emu_operand(b);
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
}

}  // namespace dupl
