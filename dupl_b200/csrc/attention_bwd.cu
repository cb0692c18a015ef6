// Attention backward on tcgen05 (SURVEY A14 / G5 "bwd kernel for training path"; reference: autograd
// through vit.py:120-135).  Per (image, head), with S = Q K^T, P = exp(scale S - lse), D_i = sum_d dO_id O_id:
//   dP = dO V^T,  dS = scale * P o (dP - D),  dV = P^T dO,  dK = dS^T Q,  dQ = dS K
// Two kernels, each shaped so that every operand is either K-major in shared memory, a 64-wide MN-major
// tile in shared memory, or rows held by the owning thread in tensor memory (the operand modes the
// forward kernel uses):
//   attn_bwd_dq_kernel   CTA = 128 queries.  Per 64-key tile: S = Q K^T and dP = dO V^T (A, B from smem),
//                        thread i turns its rows into dS_i (registers) and stores it split-bf16 to TMEM,
//                        dQ += dS K (A from TMEM, K consumed in place as an MN-major operand).
//   attn_bwd_dkv_kernel  CTA = 128 keys, works on TRANSPOSED tiles so that a thread owns a key row:
//                        S^T = K Q^T, dP^T = V dO^T per 64-query tile, thread j forms P^T_j and dS^T_j,
//                        dV += P^T dO and dK += dS^T Q (A from TMEM, dO / Q tiles as MN-major operands).
// All products are 3-pass split-bf16 with fp32 accumulation in TMEM; dQ, dK, dV accumulate in TMEM over
// the whole loop (the backward needs no rescaling) and are written once as fp32.
#include "common.cuh"
#include "ptx.cuh"

namespace dupl {

// 16 warps: warp w owns TMEM lane quadrant w % 4 (rows 32 (w % 4) .. +31 of the tile) and column group w / 4 (16 of the 64
// columns).  With 4 warps (one per scheduler, a whole 64-column row per thread) the exp / split arithmetic issued one
// instruction every 5 cycles (ncu: 1.0 active warp per scheduler, 0.2 eligible; 25 % tensor-pipe activity): the kernels were
// bound by the latency of a single warp's dependent instruction stream, not by the MMAs.
// A 17th warp issues every TMA load and every MMA and does no arithmetic: when warp 0 did both, its own instruction stream
// (~650 instructions per tile at one-warp issue rate) was the critical path and the other warps spent half their cycles at the
// CTA barrier waiting for it.  The arithmetic warps never synchronise with each other in the dQ kernel (mbarriers only).
constexpr int AB_GROUPS = 4, AB_COLS = 64 / AB_GROUPS;
constexpr int AB_MATH_THREADS = 128 * AB_GROUPS, AB_MATH_WARPS = AB_MATH_THREADS / 32;
constexpr int AB_THREADS = AB_MATH_THREADS + 32;
constexpr int AB_T64 = 64 * 64 * 2;    // one plane of a 64-row tile (8 KB)
constexpr int AB_T128 = 128 * 64 * 2;  // one plane of a 128-row tile (16 KB)
constexpr int AB_SMEM = 4 * AB_T128 + 4 * 4 * AB_T64 + 1024 + 2048;  // 64 KB resident tiles + 4 stages x 32 KB + alignment slack + barriers / per-tile statistics

struct AttnBwdTcParams {
  CUtensorMap tm_qkv128_hi, tm_qkv128_lo, tm_qkv64_hi, tm_qkv64_lo;  // [M, 3*heads*64] planes, boxes of 128 / 64 rows
  CUtensorMap tm_do128_hi, tm_do128_lo, tm_do64_hi, tm_do64_lo;      // [M, heads*64] planes
  const float* lse;   // [M, heads]
  const float* Dvec;  // [M, heads]
  float* dqkv;        // [M, 3*heads*64]
  int tokens, row_offset, heads;
  float scale;
};

__device__ __forceinline__ void store_split_row(uint32_t taddr, const float (&v)[64]) {
  uint32_t hi[32], lo[32];
#pragma unroll
  for (int c = 0; c < 64; c += 2) split2_bf16(v[c], v[c + 1], hi[c >> 1], lo[c >> 1]);
  tmem_st_32x32(taddr, hi);
  tmem_st_32x32(taddr + 32, lo);
}
// 16 fp32 columns of a row -> 8 packed bf16x2 words in the hi plane (columns [0,32) of the operand) and 8 in the lo plane
// (+32); `taddr` already points at this thread's word offset (8 * column group) of the hi plane.
__device__ __forceinline__ void store_split_cols16(uint32_t taddr, const float (&v)[AB_COLS]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int c = 0; c < 16; c += 2) split2_bf16(v[c], v[c + 1], hi[c >> 1], lo[c >> 1]);
  tmem_st_32x8(taddr, hi);
  tmem_st_32x8(taddr + 32, lo);
}
// arrive of one warp on an mbarrier that counts warps: all lanes' preceding tcgen05 stores are complete and fenced
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  tc_wait_st();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void math_warps_sync() { asm volatile("bar.sync 1, %0;" ::"n"(AB_MATH_THREADS) : "memory"); }
__device__ __forceinline__ void load_cols16(uint32_t taddr, float (&v)[AB_COLS]) {
  uint32_t a[16];
  tmem_ld_32x16(taddr, a);
  tc_wait_ld();
#pragma unroll
  for (int c = 0; c < 16; ++c) v[c] = __uint_as_float(a[c]);
}

__device__ __forceinline__ void load_row64(uint32_t taddr, float (&v)[64]) {
  uint32_t a[32], b[32];
  tmem_ld_32x32(taddr, a);
  tmem_ld_32x32(taddr + 32, b);
  tc_wait_ld();
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    v[c] = __uint_as_float(a[c]);
    v[32 + c] = __uint_as_float(b[c]);
  }
}

// Both kernels are software-pipelined over the tile loop: the S / dP MMAs of tile j+1 are issued BEFORE the threads turn
// tile j into dS (their accumulators are double-buffered in tensor memory, the K/V resp. Q/dO tiles triple-buffered in shared
// memory), so the tensor pipe always has the next tile's products queued while the exp / split arithmetic of the current tile
// runs; the only waits left on the critical path are for data that was issued a whole iteration earlier.  (The round-1
// kernels ran MMA -> wait -> arithmetic -> wait -> MMA strictly in sequence: 26-29 % tensor-pipe activity.)
constexpr int AB_STAGES = 4;

// ------------------------------------------------------------------------------------------------ dQ
// grid (q tiles of 128, heads, images)
__global__ void __launch_bounds__(AB_THREADS, 1) attn_bwd_dq_kernel(const __grid_constant__ AttnBwdTcParams p) {
  pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                     // hi | lo, 128 rows
  uint8_t* sdO = sQ + 2 * AB_T128;        // hi | lo, 128 rows
  uint8_t* sK = sdO + 2 * AB_T128;        // [AB_STAGES][hi | lo], 64 rows
  uint8_t* sV = sK + AB_STAGES * 2 * AB_T64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + AB_STAGES * 2 * AB_T64);
  uint64_t* bar_q = bars;          // Q, dO landed
  uint64_t* bar_kv = bars + 1;     // [AB_STAGES]
  uint64_t* bar_s = bars + 5;      // [2] S, dP of tile j complete (buffer j & 1)
  uint64_t* bar_d = bars + 7;      // [2] dQ MMAs of tile j complete (dS buffer j & 1, K/V stage of tile j free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
  const int hd = p.heads * 64;
  const int row0 = p.row_offset + img * p.tokens;
  const int n_kv = (p.tokens + 63) / 64;
  uint64_t* bar_ds = bars + 9;     // [2] dS of tile j stored by all arithmetic warps (count = AB_MATH_WARPS)

  if (tid == 0) {
    for (int i = 0; i < 9; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bar_ds[0], AB_MATH_WARPS);
    mbar_init(&bar_ds[1], AB_MATH_WARPS);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  constexpr int TM_S = 0, TM_DP = 128, TM_DQ = 256, TM_DS = 320;  // S, dP, dS: two 64-column buffers each

  if (warp == AB_MATH_WARPS) {
    // ------------------------------------------------------------------ TMA + MMA warp
    auto load_kv = [&](int j) {
      const int st = j % AB_STAGES;
      mbar_arrive_expect_tx(&bar_kv[st], 4 * AB_T64);
      tma_load_2d(sK + st * 2 * AB_T64, &p.tm_qkv64_hi, &bar_kv[st], hd + head * 64, row0 + j * 64);
      tma_load_2d(sK + st * 2 * AB_T64 + AB_T64, &p.tm_qkv64_lo, &bar_kv[st], hd + head * 64, row0 + j * 64);
      tma_load_2d(sV + st * 2 * AB_T64, &p.tm_qkv64_hi, &bar_kv[st], 2 * hd + head * 64, row0 + j * 64);
      tma_load_2d(sV + st * 2 * AB_T64 + AB_T64, &p.tm_qkv64_lo, &bar_kv[st], 2 * hd + head * 64, row0 + j * 64);
    };
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, 4 * AB_T128);
      tma_load_2d(sQ, &p.tm_qkv128_hi, bar_q, head * 64, row0 + qt * 128);
      tma_load_2d(sQ + AB_T128, &p.tm_qkv128_lo, bar_q, head * 64, row0 + qt * 128);
      tma_load_2d(sdO, &p.tm_do128_hi, bar_q, head * 64, row0 + qt * 128);
      tma_load_2d(sdO + AB_T128, &p.tm_do128_lo, bar_q, head * 64, row0 + qt * 128);
      for (int j = 0; j < AB_STAGES && j < n_kv; ++j) load_kv(j);
    }
    __syncwarp();
    constexpr uint32_t idesc_kk = umma_idesc_bf16(64, 0, 0);  // B K-major
    constexpr uint32_t idesc_mn = umma_idesc_bf16(64, 0, 1);  // B MN-major
    const uint64_t dQh = umma_desc_sw128(smem_u32(sQ)), dQl = umma_desc_sw128(smem_u32(sQ + AB_T128));
    const uint64_t dOh = umma_desc_sw128(smem_u32(sdO)), dOl = umma_desc_sw128(smem_u32(sdO + AB_T128));
    // all lanes; the issuing lane is elected inside the MMA wrappers: S = Q K^T and dP = dO V^T of tile j
    auto issue_s_dp = [&](int j) {
      const int st = j % AB_STAGES, b = j & 1;
      mbar_wait(&bar_kv[st], static_cast<uint32_t>((j / AB_STAGES) & 1));
      tc_fence_after();
      const uint32_t k0 = smem_u32(sK + st * 2 * AB_T64), v0 = smem_u32(sV + st * 2 * AB_T64);
      mma_ss_split<2>(tm + TM_S + b * 64, dQh, dQl, umma_desc_sw128(k0), umma_desc_sw128(k0 + AB_T64), idesc_kk, false);
      mma_ss_split<2>(tm + TM_DP + b * 64, dOh, dOl, umma_desc_sw128(v0), umma_desc_sw128(v0 + AB_T64), idesc_kk, false);
      tc_commit(&bar_s[b]);
    };
    mbar_wait(bar_q, 0);
    issue_s_dp(0);
    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      // S / dP of tile j+1 overwrite the buffers of tile j-1: the arithmetic warps are done with those (their bar_ds arrival
      // for tile j-1 was waited for below, one turn ago)
      if (j + 1 < n_kv) issue_s_dp(j + 1);
      // tile j+3 goes into the stage of tile j-1, free once dQ(j-1) has completed.  The load is issued HERE, two turns before
      // its S / dP products are: with the refill at the end of the turn and one stage less, every tile paid the full TMA
      // latency on this warp's critical path (the arithmetic warps then spent half their time waiting for S / dP).
      if (j >= 1 && j + 3 < n_kv) {
        mbar_wait(&bar_d[b ^ 1], static_cast<uint32_t>(((j - 1) >> 1) & 1));
        if (lane == 0) load_kv(j + 3);
        __syncwarp();
      }
      mbar_wait(&bar_ds[b], static_cast<uint32_t>((j >> 1) & 1));
      tc_fence_after();
      const uint32_t k0 = smem_u32(sK + (j % AB_STAGES) * 2 * AB_T64);
      mma_ts_split<128>(tm + TM_DQ, tm + TM_DS + b * 64, umma_desc_sw128(k0), umma_desc_sw128(k0 + AB_T64), idesc_mn, j > 0);  // dQ += dS K
      tc_commit(&bar_d[b]);
    }
  } else {
    // ------------------------------------------------------------------ arithmetic warps: thread = (row, 16-column group)
    const int grp = warp >> 2;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;  // TMEM lane quadrant of this warp
    const int qrow = qt * 128 + (tid & 127);
    const bool q_ok = qrow < p.tokens;
    const long grow = static_cast<long>(row0) + qrow;
    const float lse2 = q_ok ? p.lse[grow * p.heads + head] * 1.44269504088896340736f : 0.0f;
    const float Di = q_ok ? p.Dvec[grow * p.heads + head] : 0.0f;
    const float c2 = p.scale * 1.44269504088896340736f;
    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      mbar_wait(&bar_s[b], static_cast<uint32_t>((j >> 1) & 1));
      tc_fence_after();
      float s[AB_COLS], dp[AB_COLS];
      load_cols16(tm + TM_S + b * 64 + grp * AB_COLS + lane_base, s);
      load_cols16(tm + TM_DP + b * 64 + grp * AB_COLS + lane_base, dp);
      const int kv_valid = p.tokens - j * 64 - grp * AB_COLS;
#pragma unroll
      for (int c = 0; c < AB_COLS; ++c) {
        const float pv = (q_ok && c < kv_valid) ? fast_exp2(fmaf(s[c], c2, -lse2)) : 0.0f;
        s[c] = p.scale * pv * (dp[c] - Di);  // dS
      }
      if (j >= 2) {  // dQ(j-2) read this dS buffer
        mbar_wait(&bar_d[b], static_cast<uint32_t>(((j - 2) >> 1) & 1));
        tc_fence_after();
      }
      store_split_cols16(tm + TM_DS + b * 64 + grp * (AB_COLS / 2) + lane_base, s);
      warp_arrive(&bar_ds[b], lane);
    }
    mbar_wait(&bar_d[(n_kv - 1) & 1], static_cast<uint32_t>(((n_kv - 1) >> 1) & 1));
    tc_fence_after();
    float dq[AB_COLS];
    load_cols16(tm + TM_DQ + grp * AB_COLS + lane_base, dq);
    if (q_ok) {
      float4* o = reinterpret_cast<float4*>(p.dqkv + grow * 3 * hd + head * 64 + grp * AB_COLS);
#pragma unroll
      for (int c = 0; c < AB_COLS / 4; ++c) o[c] = make_float4(dq[4 * c], dq[4 * c + 1], dq[4 * c + 2], dq[4 * c + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

// ------------------------------------------------------------------------------------------------ dK, dV
// grid (kv tiles of 128, heads, images)
__global__ void __launch_bounds__(AB_THREADS, 1) attn_bwd_dkv_kernel(const __grid_constant__ AttnBwdTcParams p) {
  pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                   // hi | lo, 128 rows
  uint8_t* sV = sK + 2 * AB_T128;
  uint8_t* sQ = sV + 2 * AB_T128;       // [AB_STAGES][hi | lo], 64 rows
  uint8_t* sdO = sQ + AB_STAGES * 2 * AB_T64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdO + AB_STAGES * 2 * AB_T64);
  uint64_t* bar_kv = bars;
  uint64_t* bar_q = bars + 1;  // [AB_STAGES]
  uint64_t* bar_s = bars + 5;  // [2] S^T, dP^T of tile i complete (buffer i & 1)
  uint64_t* bar_d = bars + 7;  // dV / dK MMAs of a tile complete (P^T / dS^T buffer and Q/dO stage free): one phase per tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  // per 64-query tile: lse * log2(e) (+inf on padding rows) | D, double-buffered; [tile & 1][0:64 lse2, 64:128 D]
  float* s_stat = reinterpret_cast<float*>(bars + 14);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kt = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
  const int hd = p.heads * 64;
  const int row0 = p.row_offset + img * p.tokens;
  const int n_q = (p.tokens + 63) / 64;
  uint64_t* bar_ds = bars + 8;  // P^T / dS^T of a tile stored by all arithmetic warps (count = AB_MATH_WARPS): one phase per tile

  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    mbar_init(bar_ds, AB_MATH_WARPS);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  // S^T, dP^T: two 64-column buffers each; P^T, dS^T (split: hi | lo) single-buffered — all 512 columns in use
  constexpr int TM_ST = 0, TM_DPT = 128, TM_DV = 256, TM_DK = 320, TM_PT = 384, TM_DST = 448;

  if (warp == AB_MATH_WARPS) {
    // ------------------------------------------------------------------ TMA + MMA warp
    auto load_q = [&](int i) {
      const int st = i % AB_STAGES;
      mbar_arrive_expect_tx(&bar_q[st], 4 * AB_T64);
      tma_load_2d(sQ + st * 2 * AB_T64, &p.tm_qkv64_hi, &bar_q[st], head * 64, row0 + i * 64);
      tma_load_2d(sQ + st * 2 * AB_T64 + AB_T64, &p.tm_qkv64_lo, &bar_q[st], head * 64, row0 + i * 64);
      tma_load_2d(sdO + st * 2 * AB_T64, &p.tm_do64_hi, &bar_q[st], head * 64, row0 + i * 64);
      tma_load_2d(sdO + st * 2 * AB_T64 + AB_T64, &p.tm_do64_lo, &bar_q[st], head * 64, row0 + i * 64);
    };
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_kv, 4 * AB_T128);
      tma_load_2d(sK, &p.tm_qkv128_hi, bar_kv, hd + head * 64, row0 + kt * 128);
      tma_load_2d(sK + AB_T128, &p.tm_qkv128_lo, bar_kv, hd + head * 64, row0 + kt * 128);
      tma_load_2d(sV, &p.tm_qkv128_hi, bar_kv, 2 * hd + head * 64, row0 + kt * 128);
      tma_load_2d(sV + AB_T128, &p.tm_qkv128_lo, bar_kv, 2 * hd + head * 64, row0 + kt * 128);
      for (int i = 0; i < AB_STAGES && i < n_q; ++i) load_q(i);
    }
    __syncwarp();
    constexpr uint32_t idesc_kk = umma_idesc_bf16(64, 0, 0);
    constexpr uint32_t idesc_mn = umma_idesc_bf16(64, 0, 1);
    const uint64_t dKh = umma_desc_sw128(smem_u32(sK)), dKl = umma_desc_sw128(smem_u32(sK + AB_T128));
    const uint64_t dVh = umma_desc_sw128(smem_u32(sV)), dVl = umma_desc_sw128(smem_u32(sV + AB_T128));
    // S^T = K Q^T and dP^T = V dO^T of query tile i
    auto issue_s_dp = [&](int i) {
      const int st = i % AB_STAGES, b = i & 1;
      mbar_wait(&bar_q[st], static_cast<uint32_t>((i / AB_STAGES) & 1));
      tc_fence_after();
      const uint32_t q0 = smem_u32(sQ + st * 2 * AB_T64), o0 = smem_u32(sdO + st * 2 * AB_T64);
      mma_ss_split<2>(tm + TM_ST + b * 64, dKh, dKl, umma_desc_sw128(q0), umma_desc_sw128(q0 + AB_T64), idesc_kk, false);
      mma_ss_split<2>(tm + TM_DPT + b * 64, dVh, dVl, umma_desc_sw128(o0), umma_desc_sw128(o0 + AB_T64), idesc_kk, false);
      tc_commit(&bar_s[b]);
    };
    mbar_wait(bar_kv, 0);
    issue_s_dp(0);
    for (int i = 0; i < n_q; ++i) {
      if (i + 1 < n_q) issue_s_dp(i + 1);  // queued behind dV / dK of tile i-1
      // tile i+3 goes into the stage of tile i-1, free once dV / dK of tile i-1 have completed; issued two turns before its
      // S^T / dP^T products are.  (The wait sits BEFORE this tile's commit: with one barrier for all tiles, a wait issued after
      // it could find the phase flipped twice.)
      if (i >= 1 && i + 3 < n_q) {
        mbar_wait(bar_d, static_cast<uint32_t>((i - 1) & 1));
        if (lane == 0) load_q(i + 3);
        __syncwarp();
      }
      mbar_wait(bar_ds, static_cast<uint32_t>(i & 1));
      tc_fence_after();
      const int st = i % AB_STAGES;
      const uint32_t q0 = smem_u32(sQ + st * 2 * AB_T64), o0 = smem_u32(sdO + st * 2 * AB_T64);
      mma_ts_split<128>(tm + TM_DV, tm + TM_PT, umma_desc_sw128(o0), umma_desc_sw128(o0 + AB_T64), idesc_mn, i > 0);   // dV += P^T dO
      mma_ts_split<128>(tm + TM_DK, tm + TM_DST, umma_desc_sw128(q0), umma_desc_sw128(q0 + AB_T64), idesc_mn, i > 0);  // dK += dS^T Q
      tc_commit(bar_d);
    }
  } else {
    // ------------------------------------------------------------------ arithmetic warps: thread = (key row, 16-query group)
    const int grp = warp >> 2;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;  // TMEM lane quadrant of this warp
    const int key = kt * 128 + (tid & 127);
    const bool k_ok = key < p.tokens;
    const float c2 = p.scale * 1.44269504088896340736f;
    // thread t < 64 fetches lse of query t of a tile, thread 64 + t its D (one coalesced-ish load per thread and tile,
    // issued one tile ahead; the tile loop then reads them as shared-memory broadcasts)
    auto fetch_stat = [&](int i) -> float {
      const int q = i * 64 + (tid & 63);
      if (q >= p.tokens) return tid < 64 ? __int_as_float(0x7f800000) : 0.0f;
      const long r = (static_cast<long>(row0) + q) * p.heads + head;
      return tid < 64 ? p.lse[r] * 1.44269504088896340736f : p.Dvec[r];
    };
    if (tid < 128) s_stat[tid] = fetch_stat(0);
    math_warps_sync();
    for (int i = 0; i < n_q; ++i) {
      const int b = i & 1;
      const float stat_next = (tid < 128 && i + 1 < n_q) ? fetch_stat(i + 1) : 0.0f;
      mbar_wait(&bar_s[b], static_cast<uint32_t>((i >> 1) & 1));
      tc_fence_after();
      float s[AB_COLS], dp[AB_COLS];
      load_cols16(tm + TM_ST + b * 64 + grp * AB_COLS + lane_base, s);
      load_cols16(tm + TM_DPT + b * 64 + grp * AB_COLS + lane_base, dp);
      const float4* stat4 = reinterpret_cast<const float4*>(s_stat + b * 128 + grp * AB_COLS);
      const float kscale = k_ok ? p.scale : 0.0f;
#pragma unroll
      for (int c4 = 0; c4 < AB_COLS / 4; ++c4) {
        const float4 l4 = stat4[c4], d4 = stat4[16 + c4];
        const float l[4] = {l4.x, l4.y, l4.z, l4.w}, d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = c4 * 4 + e;
          const float pv = k_ok ? fast_exp2(fmaf(s[c], c2, -l[e])) : 0.0f;  // padding queries: lse2 = +inf -> 0
          s[c] = pv;
          dp[c] = kscale * pv * (dp[c] - d[e]);
        }
      }
      if (i >= 1) {  // dV / dK of tile i-1 read the P^T / dS^T buffers
        mbar_wait(bar_d, static_cast<uint32_t>((i - 1) & 1));
        tc_fence_after();
      }
      store_split_cols16(tm + TM_PT + grp * (AB_COLS / 2) + lane_base, s);
      store_split_cols16(tm + TM_DST + grp * (AB_COLS / 2) + lane_base, dp);
      if (tid < 128) s_stat[(b ^ 1) * 128 + tid] = stat_next;  // buffer of tile i-1
      warp_arrive(bar_ds, lane);
      // the statistics of tile i+1 must be visible to every arithmetic warp, and nobody may still read buffer b^1 (tile i-1)
      // when it is overwritten two lines above: one barrier among the arithmetic warps per tile covers both
      math_warps_sync();
    }
    mbar_wait(bar_d, static_cast<uint32_t>((n_q - 1) & 1));
    tc_fence_after();
    float dv[AB_COLS];
    load_cols16(tm + TM_DV + grp * AB_COLS + lane_base, dv);
    if (k_ok) {
      float4* o = reinterpret_cast<float4*>(p.dqkv + (static_cast<long>(row0) + key) * 3 * hd + 2 * hd + head * 64 + grp * AB_COLS);
#pragma unroll
      for (int c = 0; c < AB_COLS / 4; ++c) o[c] = make_float4(dv[4 * c], dv[4 * c + 1], dv[4 * c + 2], dv[4 * c + 3]);
    }
    load_cols16(tm + TM_DK + grp * AB_COLS + lane_base, dv);
    if (k_ok) {
      float4* o = reinterpret_cast<float4*>(p.dqkv + (static_cast<long>(row0) + key) * 3 * hd + hd + head * 64 + grp * AB_COLS);
#pragma unroll
      for (int c = 0; c < AB_COLS / 4; ++c) o[c] = make_float4(dv[4 * c], dv[4 * c + 1], dv[4 * c + 2], dv[4 * c + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

// D[row][head] = sum_d dO[row][head*64 + d] * O[row][head*64 + d], both given as split planes
__global__ void __launch_bounds__(256) attn_bwd_d_kernel(const __nv_bfloat16* __restrict__ do_hi, const __nv_bfloat16* __restrict__ do_lo,
                                                         const __nv_bfloat16* __restrict__ o_hi, const __nv_bfloat16* __restrict__ o_lo,
                                                         float* __restrict__ Dvec, int row_offset, int rows, int heads) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (item >= rows * heads) return;
  const int row = row_offset + item / heads, h = item % heads;
  const long o = static_cast<long>(row) * heads * 64 + h * 64;
  float s = 0.0f;
  for (int d = lane; d < 64; d += 32)
    s += (__bfloat162float(do_hi[o + d]) + __bfloat162float(do_lo[o + d])) * (__bfloat162float(o_hi[o + d]) + __bfloat162float(o_lo[o + d]));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) Dvec[static_cast<long>(row) * heads + h] = s;
}

}  // namespace dupl

extern "C" int dupl_attention_bwd(const dupl_attention_bwd_args* a, void* stream) {
  using namespace dupl;
  DUPL_CHECK_ARG(a != nullptr, "dupl_attention_bwd: args is NULL");
  DUPL_CHECK_ARG(a->qkv_hi && a->qkv_lo && a->o_hi && a->o_lo && a->do_hi && a->do_lo && a->lse && a->Dvec && a->dqkv,
                 "dupl_attention_bwd: NULL pointer");
  DUPL_CHECK_ARG(a->batch > 0 && a->tokens > 0 && a->heads > 0 && a->M >= a->row_offset + a->batch * a->tokens,
                 "dupl_attention_bwd: bad shape");
  AttnBwdTcParams P;
  memset(&P, 0, sizeof(P));
  const int hd = a->heads * 64;
  int rc;
  if ((rc = make_tmap_bf16_2d(&P.tm_qkv128_hi, a->qkv_hi, a->M, 3 * hd, 3 * hd, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&P.tm_qkv128_lo, a->qkv_lo, a->M, 3 * hd, 3 * hd, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&P.tm_qkv64_hi, a->qkv_hi, a->M, 3 * hd, 3 * hd, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&P.tm_qkv64_lo, a->qkv_lo, a->M, 3 * hd, 3 * hd, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&P.tm_do128_hi, a->do_hi, a->M, hd, hd, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&P.tm_do128_lo, a->do_lo, a->M, hd, hd, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&P.tm_do64_hi, a->do_hi, a->M, hd, hd, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&P.tm_do64_lo, a->do_lo, a->M, hd, hd, 64))) return rc;
  P.lse = a->lse; P.Dvec = a->Dvec; P.dqkv = a->dqkv;
  P.tokens = a->tokens; P.row_offset = a->row_offset; P.heads = a->heads; P.scale = a->scale;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool attr_set = false;
  if (!attr_set) {
    DUPL_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    DUPL_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    attr_set = true;
  }
  const int rows = a->batch * a->tokens;
  DUPL_CUDA_OK(launch_pdl(attn_bwd_d_kernel, dim3(cdiv(rows * a->heads, 8)), dim3(256), 0, st,
                          static_cast<const __nv_bfloat16*>(a->do_hi), static_cast<const __nv_bfloat16*>(a->do_lo),
                          static_cast<const __nv_bfloat16*>(a->o_hi), static_cast<const __nv_bfloat16*>(a->o_lo), a->Dvec,
                          a->row_offset, rows, a->heads));
  count_launch();
  dim3 grid(cdiv(a->tokens, 128), a->heads, a->batch);
  DUPL_CUDA_OK(launch_pdl(attn_bwd_dkv_kernel, grid, dim3(AB_THREADS), AB_SMEM, st, P));
  count_launch();
  DUPL_CUDA_OK(launch_pdl(attn_bwd_dq_kernel, grid, dim3(AB_THREADS), AB_SMEM, st, P));
  count_launch();
  return DUPL_OK;
}
