// Fused attention forward for the ViT blocks (reference vit.py:120-135 materialises
// the [B,12,N,N] matrix): softmax(Q K^T * scale) V per (image, head), flash-style on tcgen05.
//
// One CTA = one (image, head) and TWO 128-query tiles A and B that share every K/V tile
// ("ping-pong"): while the softmax warps of tile A work on S_A(j), the tensor core computes
// S_B(j), P_A V, S_A(j+1) ... so neither pipe waits for the other in steady state.
//   warps 0-3   softmax of tile A      thread t owns query row t end to end: tcgen05.ld hands it the
//   warps 4-7   softmax of tile B      whole row of S -> running max / sum in registers, no shuffles;
//                                      P is split to bf16 hi/lo and written back to TENSOR MEMORY
//                                      (tcgen05.st), where the P V MMA reads it as its A operand;
//                                      O = O*alpha + (P V) is accumulated in registers from the fresh
//                                      per-tile product, so TMEM is never rescaled
//   warp 8      TMA producer: K(j) / V(j) tiles, 3 stages each, straight out of the qkv GEMM's
//               [M, 2304] split-bf16 planes (V is consumed in place as an MN-major operand)
//   warp 9      MMA issuer: S = Q K^T and P V as 3-pass split-bf16 tcgen05.mma 128x64x16 with the A
//               operand (Q resp. P) in tensor memory.  With both operands in shared memory a
//               128x64x16 MMA needs 192 B/clk of smem reads (the SM delivers 128): the first version
//               of this kernel was shared-memory bound at 30 % tensor-pipe utilisation.
// TMEM (512 columns): S_A S_B O_A O_B (64 fp32 columns each) | Q_A Q_B | P_A P_B (hi 32 + lo 32 columns
// each, two bf16 per column).  Q is loaded once by the softmax threads (global -> registers -> TMEM).
#include "common.cuh"
#include "ptx.cuh"

namespace dupl {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 64;
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 320;
constexpr int ATT_KV_BYTES = ATT_BKV * ATT_D * 2;  // one plane of one K or V tile: 8 KB
constexpr int ATT_STAGES = 3;
// K: stages x 2 planes | V: stages x 2 planes
constexpr int ATT_OFF_K = 0;
constexpr int ATT_OFF_V = ATT_OFF_K + ATT_STAGES * 2 * ATT_KV_BYTES;
constexpr int ATT_OFF_BAR = ATT_OFF_V + ATT_STAGES * 2 * ATT_KV_BYTES;
constexpr int ATT_SMEM = ATT_OFF_BAR + 256 + 1024;
// tensor-memory columns
constexpr int TM_S = 0;      // + t*64
constexpr int TM_O = 128;    // + t*64
constexpr int TM_Q = 256;    // + t*64 : hi 32 columns, lo 32 columns
constexpr int TM_P = 384;    // + t*64 : hi 32 columns, lo 32 columns

struct AttnParamsDev {
  CUtensorMap tm_kv_hi, tm_kv_lo;
  dupl_segment seg[DUPL_MAX_SEGMENTS];
  int cta_start[DUPL_MAX_SEGMENTS + 1];
  int nseg, heads, M;
  float scale_log2e;
  const __nv_bfloat16* q_hi;  // the qkv planes (Q rows are read directly by the softmax threads)
  const __nv_bfloat16* q_lo;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  float* lse;  // optional [M, heads]: ln sum_j exp(scale * s_ij), kept for the backward pass
};

enum {  // mbarrier slots
  BAR_Q = 0,
  BAR_K_FULL = 1,                       // [stages]
  BAR_K_EMPTY = BAR_K_FULL + ATT_STAGES,
  BAR_V_FULL = BAR_K_EMPTY + ATT_STAGES,
  BAR_V_EMPTY = BAR_V_FULL + ATT_STAGES,
  BAR_S_FULL = BAR_V_EMPTY + ATT_STAGES,  // [2 tiles]
  BAR_S_FREE = BAR_S_FULL + 2,
  BAR_P_READY = BAR_S_FREE + 2,
  BAR_O_FULL = BAR_P_READY + 2,
  BAR_COUNT = BAR_O_FULL + 2
};

__global__ void __launch_bounds__(ATT_THREADS, 1) attention_fwd_kernel(const __grid_constant__ AttnParamsDev p) {
  pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  // ---- which (segment, image, head, pair of q-tiles) is this CTA?
  int si = 0;
  for (int s = 1; s < p.nseg; ++s)
    if (static_cast<int>(blockIdx.x) >= p.cta_start[s]) si = s;
  const dupl_segment sg = p.seg[si];
  const int q_pairs = (sg.tokens + 2 * ATT_BQ - 1) / (2 * ATT_BQ);
  int local = blockIdx.x - p.cta_start[si];
  const int qp = local % q_pairs;
  local /= q_pairs;
  const int head = local % p.heads;
  const int img = local / p.heads;
  const int img_row0 = sg.row_offset + img * sg.tokens;
  const int n_kv = (sg.tokens + ATT_BKV - 1) / ATT_BKV;
  const int hd = p.heads * ATT_D;  // 768

  if (tid == 0) {
    for (int i = 0; i < BAR_COUNT; ++i) {
      const bool wg_arrives = (i >= BAR_S_FREE && i < BAR_O_FULL);  // s_free / p_ready: all 128 softmax threads
      mbar_init(&bars[i], i == BAR_Q ? 256u : (wg_arrives ? 128u : 1u));  // Q: both warpgroups store their tile
    }
    fence_mbar_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // =========================================================== TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&p.tm_kv_hi);
      tma_prefetch_desc(&p.tm_kv_lo);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % ATT_STAGES;
        const uint32_t ph = static_cast<uint32_t>((j / ATT_STAGES) & 1);
        const int row = img_row0 + j * ATT_BKV;
        uint8_t* sk = smem + ATT_OFF_K + st * 2 * ATT_KV_BYTES;
        uint8_t* sv = smem + ATT_OFF_V + st * 2 * ATT_KV_BYTES;
        mbar_wait(&bars[BAR_K_EMPTY + st], ph ^ 1);
        mbar_arrive_expect_tx(&bars[BAR_K_FULL + st], 2 * ATT_KV_BYTES);
        tma_load_2d(sk, &p.tm_kv_hi, &bars[BAR_K_FULL + st], hd + head * ATT_D, row);
        tma_load_2d(sk + ATT_KV_BYTES, &p.tm_kv_lo, &bars[BAR_K_FULL + st], hd + head * ATT_D, row);
        mbar_wait(&bars[BAR_V_EMPTY + st], ph ^ 1);
        mbar_arrive_expect_tx(&bars[BAR_V_FULL + st], 2 * ATT_KV_BYTES);
        tma_load_2d(sv, &p.tm_kv_hi, &bars[BAR_V_FULL + st], 2 * hd + head * ATT_D, row);
        tma_load_2d(sv + ATT_KV_BYTES, &p.tm_kv_lo, &bars[BAR_V_FULL + st], 2 * hd + head * ATT_D, row);
      }
    }
  } else if (warp == 9) {
    // =========================================================== MMA issuer (whole warp, leader elected per op)
    {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BKV, 0, 0);  // A = Q (K-major), B = K (K-major), N = 64 keys
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_D, 0, 1);    // A = P (K-major), B = V (MN-major), N = 64 dims
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t sK = smem_u32(smem + ATT_OFF_K), sV = smem_u32(smem + ATT_OFF_V);
      uint64_t dK[ATT_STAGES][2], dV[ATT_STAGES][2];
#pragma unroll
      for (int st = 0; st < ATT_STAGES; ++st)
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          dK[st][pl] = umma_desc_sw128(sK + (2 * st + pl) * ATT_KV_BYTES);
          dV[st][pl] = umma_desc_sw128(sV + (2 * st + pl) * ATT_KV_BYTES);
        }
      auto qk = [&](int t, int st) {  // S_t = Q_t K^T
        mma_ts_split<2>(tm + TM_S + t * 64, tm + TM_Q + t * 64, dK[st][0], dK[st][1], idesc_qk, false);
        tc_commit(&bars[BAR_S_FULL + t]);
      };
      auto pv = [&](int t, int st) {  // O_t = P_t V
        mma_ts_split<128>(tm + TM_O + t * 64, tm + TM_P + t * 64, dV[st][0], dV[st][1], idesc_pv, false);
        tc_commit(&bars[BAR_O_FULL + t]);
      };
      mbar_wait(&bars[BAR_Q], 0);
      mbar_wait(&bars[BAR_K_FULL + 0], 0);
      tc_fence_after();
      qk(0, 0);
      qk(1, 0);
      tc_commit(&bars[BAR_K_EMPTY + 0]);
#pragma unroll 1
      for (int j2 = 0; j2 < n_kv; j2 += ATT_STAGES) {
#pragma unroll
        for (int st = 0; st < ATT_STAGES; ++st) {  // stage index is a compile-time constant inside
          const int j = j2 + st;
          if (j < n_kv) {
            const uint32_t ph = static_cast<uint32_t>(j & 1);
            const uint32_t st_ph = static_cast<uint32_t>((j / ATT_STAGES) & 1);
            const bool more = j + 1 < n_kv;
            const int st1 = (st + 1) % ATT_STAGES;
            const uint32_t st1_ph = static_cast<uint32_t>(((j + 1) / ATT_STAGES) & 1);
            if (more) {
              mbar_wait(&bars[BAR_K_FULL + st1], st1_ph);
              mbar_wait(&bars[BAR_S_FREE + 0], ph);  // tile A has pulled S_A(j) into registers
              tc_fence_after();
              qk(0, st1);
            }
            mbar_wait(&bars[BAR_V_FULL + st], st_ph);
            mbar_wait(&bars[BAR_P_READY + 0], ph);
            tc_fence_after();
            pv(0, st);
            if (more) {
              mbar_wait(&bars[BAR_S_FREE + 1], ph);
              tc_fence_after();
              qk(1, st1);
              tc_commit(&bars[BAR_K_EMPTY + st1]);
            }
            mbar_wait(&bars[BAR_P_READY + 1], ph);
            tc_fence_after();
            pv(1, st);
            tc_commit(&bars[BAR_V_EMPTY + st]);
          }
        }
      }
    }
  } else {
    // =========================================================== softmax warpgroups (tile 0: warps 0-3, tile 1: warps 4-7)
    const int t = warp >> 2;
    const int r = tid & 127;  // query row inside the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_s = tmem_base + TM_S + t * 64 + lane_base;
    const uint32_t tmem_o = tmem_base + TM_O + t * 64 + lane_base;
    const uint32_t tmem_q = tmem_base + TM_Q + t * 64 + lane_base;
    const uint32_t tmem_p = tmem_base + TM_P + t * 64 + lane_base;

    // ---- this thread's query row: global -> registers -> tensor memory (hi 32 columns | lo 32 columns)
    {
      const long grow = static_cast<long>(img_row0) + (2 * qp + t) * ATT_BQ + r;
      uint32_t qh[32], ql[32];
      if (grow < p.M) {
        const uint4* gh = reinterpret_cast<const uint4*>(p.q_hi + grow * 3 * hd + head * ATT_D);
        const uint4* gl = reinterpret_cast<const uint4*>(p.q_lo + grow * 3 * hd + head * ATT_D);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 a = __ldg(gh + i), b = __ldg(gl + i);
          qh[4 * i] = a.x; qh[4 * i + 1] = a.y; qh[4 * i + 2] = a.z; qh[4 * i + 3] = a.w;
          ql[4 * i] = b.x; ql[4 * i + 1] = b.y; ql[4 * i + 2] = b.z; ql[4 * i + 3] = b.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) qh[i] = ql[i] = 0u;
      }
      tmem_st_32x32(tmem_q, qh);
      tmem_st_32x32(tmem_q + 32, ql);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&bars[BAR_Q]);
    }

    float o[ATT_D];
#pragma unroll
    for (int d = 0; d < ATT_D; ++d) o[d] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f, alpha_prev = 0.0f;

    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = static_cast<uint32_t>(j & 1);
      // ---- S row -> registers, then hand the TMEM buffer back to the tensor core
      uint32_t sv[2][32];
      mbar_wait(&bars[BAR_S_FULL + t], ph);
      tc_fence_after();
      tmem_ld_32x32(tmem_s, sv[0]);
      tmem_ld_32x32(tmem_s + 32, sv[1]);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(&bars[BAR_S_FREE + t]);

      float s[ATT_BKV];
#pragma unroll
      for (int c = 0; c < ATT_BKV; ++c) s[c] = __uint_as_float(sv[c >> 5][c & 31]);
      const int kv_valid = sg.tokens - j * ATT_BKV;  // keys >= kv_valid belong to another image / padding
      if (kv_valid < ATT_BKV) {
#pragma unroll
        for (int c = 0; c < ATT_BKV; ++c)
          if (c >= kv_valid) s[c] = -INFINITY;
      }
      float m_tile = s[0];
#pragma unroll
      for (int c = 1; c < ATT_BKV; ++c) m_tile = fmaxf(m_tile, s[c]);
      const float m_new = fmaxf(m_run, m_tile);
      const float alpha = fast_exp2((m_run - m_new) * p.scale_log2e);
      const float mb = m_new * p.scale_log2e;
      m_run = m_new;

      // ---- P = exp2(s*scale*log2e - mb), split to bf16 hi/lo (packed), row sum from the fp32 values
      uint32_t phi[ATT_BKV / 2], plo[ATT_BKV / 2];
      float l_tile = 0.0f;
#pragma unroll
      for (int c = 0; c < ATT_BKV; c += 2) {
        const float p0 = fast_exp2(fmaf(s[c], p.scale_log2e, -mb));
        const float p1 = fast_exp2(fmaf(s[c + 1], p.scale_log2e, -mb));
        l_tile += p0 + p1;
        split2_bf16(p0, p1, phi[c >> 1], plo[c >> 1]);
      }
      l_run = fmaf(l_run, alpha, l_tile);

      // ---- fold in the previous tile's P V (its MMAs have long finished; also frees the P buffer)
      if (j > 0) {
        mbar_wait(&bars[BAR_O_FULL + t], ph ^ 1);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {  // 32 columns at a time keeps the register peak below the cap
          uint32_t ov[32];
          tmem_ld_32x32(tmem_o + half * 32, ov);
          tc_wait_ld();
#pragma unroll
          for (int d = 0; d < 32; ++d) o[half * 32 + d] = fmaf(o[half * 32 + d], alpha_prev, __uint_as_float(ov[d]));
        }
      }
      alpha_prev = alpha;

      // ---- publish P(j): tensor memory, two bf16 per column (the A operand of the P V MMA)
      tmem_st_32x32(tmem_p, phi);
      tmem_st_32x32(tmem_p + 32, plo);
      tc_wait_st();
      tc_fence_before();  // also orders the O-tile TMEM reads above before the next P V is issued
      mbar_arrive(&bars[BAR_P_READY + t]);
    }

    // ---- last tile's P V, normalise, store this thread's row as split bf16
    mbar_wait(&bars[BAR_O_FULL + t], static_cast<uint32_t>((n_kv - 1) & 1));
    tc_fence_after();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t ov[32];
      tmem_ld_32x32(tmem_o + half * 32, ov);
      tc_wait_ld();
#pragma unroll
      for (int d = 0; d < 32; ++d) o[half * 32 + d] = fmaf(o[half * 32 + d], alpha_prev, __uint_as_float(ov[d]));
    }
    const int qrow = (2 * qp + t) * ATT_BQ + r;
    if (qrow < sg.tokens) {
      if (p.lse != nullptr)
        p.lse[static_cast<long>(img_row0 + qrow) * p.heads + head] =
            (m_run * p.scale_log2e + log2f(l_run)) * 0.69314718055994530942f;
      const float inv = 1.0f / l_run;
      const long off = static_cast<long>(img_row0 + qrow) * hd + head * ATT_D;
      __nv_bfloat16* oh = p.out_hi + off;
      __nv_bfloat16* ol = p.out_lo + off;
#pragma unroll
      for (int d8 = 0; d8 < ATT_D / 8; ++d8) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2_bf16(o[d8 * 8 + 2 * e] * inv, o[d8 * 8 + 2 * e + 1] * inv, hw[e], lw[e]);
        *reinterpret_cast<uint4*>(oh + d8 * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(ol + d8 * 8) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace dupl

extern "C" int dupl_attention_fwd(const dupl_attention_args* a, void* stream) {
  using namespace dupl;
  DUPL_CHECK_ARG(a != nullptr, "dupl_attention_fwd: args is NULL");
  DUPL_CHECK_ARG(a->nseg >= 1 && a->nseg <= DUPL_MAX_SEGMENTS, "dupl_attention_fwd: nseg=%d", a->nseg);
  DUPL_CHECK_ARG(a->heads >= 1 && a->M > 0, "dupl_attention_fwd: heads=%d M=%d", a->heads, a->M);
  DUPL_CHECK_ARG(a->qkv_hi && a->qkv_lo && a->out_hi && a->out_lo, "dupl_attention_fwd: NULL plane");
  AttnParamsDev P;
  memset(&P, 0, sizeof(P));
  const int hd = a->heads * ATT_D;
  int rc;
  if ((rc = make_tmap_bf16_2d(&P.tm_kv_hi, a->qkv_hi, a->M, 3 * hd, 3 * hd, ATT_BKV))) return rc;
  if ((rc = make_tmap_bf16_2d(&P.tm_kv_lo, a->qkv_lo, a->M, 3 * hd, 3 * hd, ATT_BKV))) return rc;
  // Longest sequences first so the tail of the grid is made of short CTAs.
  int order[DUPL_MAX_SEGMENTS];
  for (int s = 0; s < a->nseg; ++s) order[s] = s;
  for (int i = 0; i < a->nseg; ++i)
    for (int j = i + 1; j < a->nseg; ++j)
      if (a->seg[order[j]].tokens > a->seg[order[i]].tokens) {
        int t = order[i]; order[i] = order[j]; order[j] = t;
      }
  int total = 0;
  for (int s = 0; s < a->nseg; ++s) {
    const dupl_segment& sg = a->seg[order[s]];
    DUPL_CHECK_ARG(sg.batch > 0 && sg.tokens > 0 && sg.row_offset >= 0 &&
                       sg.row_offset + sg.batch * sg.tokens <= a->M,
                   "dupl_attention_fwd: segment %d out of range", order[s]);
    P.seg[s] = sg;
    P.cta_start[s] = total;
    total += sg.batch * a->heads * cdiv(sg.tokens, 2 * ATT_BQ);
  }
  P.cta_start[a->nseg] = total;
  P.nseg = a->nseg;
  P.heads = a->heads;
  P.scale_log2e = a->scale * 1.44269504088896340736f;
  P.M = a->M;
  P.q_hi = static_cast<const __nv_bfloat16*>(a->qkv_hi);
  P.q_lo = static_cast<const __nv_bfloat16*>(a->qkv_lo);
  P.out_hi = static_cast<__nv_bfloat16*>(a->out_hi);
  P.out_lo = static_cast<__nv_bfloat16*>(a->out_lo);
  P.lse = a->lse;
  static bool attr_set = false;
  if (!attr_set) {
    DUPL_CUDA_OK(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  DUPL_CUDA_OK(launch_pdl(attention_fwd_kernel, dim3(total), dim3(ATT_THREADS), ATT_SMEM, static_cast<cudaStream_t>(stream), P));
  count_launch();
  return DUPL_OK;
}
