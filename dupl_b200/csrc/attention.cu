// Fused attention forward for the ViT blocks (reference vit.py:120-135 materialises
// the [B,12,N,N] matrix): softmax(Q K^T * scale) V per (image, head), flash-style on tcgen05.
//
// One CTA = one (image, head, 128-query tile); 128 threads, thread t owns query row t end to end:
//   S = Q K^T        tcgen05.mma 128x64x16, split-bf16 (hi*hi + hi*lo + lo*hi), fp32 in TMEM
//   softmax          tcgen05.ld gives each thread its whole row -> running max / sum in registers,
//                    no shuffles; P is split to bf16 hi/lo and written to shared memory in the
//                    128-byte-swizzled K-major layout the tensor core reads
//   O_tile = P V     V tile used in place as an MN-major operand (no transpose pass), fp32 in TMEM,
//                    then O = O*alpha + O_tile in registers
// Q/K/V tiles are staged by TMA straight out of the qkv GEMM's [M, 2304] planes.  K and V are
// single-buffered but each load is issued as soon as its buffer drains (K(j+1) after S(j) is
// complete, V(j+1) after P V(j)), so the copies hide behind the softmax; shared memory is 96 KB so
// two CTAs share an SM and one CTA's softmax overlaps the other's MMAs.
#include "common.cuh"
#include "ptx.cuh"

namespace dupl {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 64;
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 128;
constexpr int ATT_Q_BYTES = ATT_BQ * ATT_D * 2;    // one plane 16 KB
constexpr int ATT_KV_BYTES = ATT_BKV * ATT_D * 2;  // one plane 8 KB
constexpr int ATT_P_BYTES = ATT_BQ * ATT_BKV * 2;  // one plane 16 KB
constexpr int ATT_SMEM = 2 * ATT_Q_BYTES + 4 * ATT_KV_BYTES + 2 * ATT_P_BYTES + 1024 + 128;

struct AttnParamsDev {
  CUtensorMap tm_q_hi, tm_q_lo, tm_kv_hi, tm_kv_lo;
  dupl_segment seg[DUPL_MAX_SEGMENTS];
  int cta_start[DUPL_MAX_SEGMENTS + 1];
  int nseg, heads;
  float scale_log2e;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
};

__global__ void __launch_bounds__(ATT_THREADS, 2) attention_fwd_kernel(const __grid_constant__ AttnParamsDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // hi | lo
  uint8_t* sK = sQ + 2 * ATT_Q_BYTES;       // hi | lo
  uint8_t* sV = sK + 2 * ATT_KV_BYTES;      // hi | lo
  uint8_t* sP = sV + 2 * ATT_KV_BYTES;      // hi | lo
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_P_BYTES);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;

  // ---- which (segment, image, head, q-tile) is this CTA?
  int si = 0;
  for (int s = 1; s < p.nseg; ++s)
    if (static_cast<int>(blockIdx.x) >= p.cta_start[s]) si = s;
  const dupl_segment sg = p.seg[si];
  const int q_tiles = (sg.tokens + ATT_BQ - 1) / ATT_BQ;
  int local = blockIdx.x - p.cta_start[si];
  const int qt = local % q_tiles;
  local /= q_tiles;
  const int head = local % p.heads;
  const int img = local / p.heads;
  const int img_row0 = sg.row_offset + img * sg.tokens;
  const int n_kv = (sg.tokens + ATT_BKV - 1) / ATT_BKV;
  const int hd = p.heads * ATT_D;  // 768

  if (tid == 0) {
    tma_prefetch_desc(&p.tm_q_hi);
    tma_prefetch_desc(&p.tm_q_lo);
    tma_prefetch_desc(&p.tm_kv_hi);
    tma_prefetch_desc(&p.tm_kv_lo);
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;        // 64 fp32 columns
  const uint32_t tmem_o = tmem_base + 64;   // 64 fp32 columns
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, 2 * ATT_Q_BYTES);
    tma_load_2d(sQ, &p.tm_q_hi, bar_q, head * ATT_D, img_row0 + qt * ATT_BQ);
    tma_load_2d(sQ + ATT_Q_BYTES, &p.tm_q_lo, bar_q, head * ATT_D, img_row0 + qt * ATT_BQ);
    mbar_arrive_expect_tx(bar_k, 2 * ATT_KV_BYTES);
    tma_load_2d(sK, &p.tm_kv_hi, bar_k, hd + head * ATT_D, img_row0);
    tma_load_2d(sK + ATT_KV_BYTES, &p.tm_kv_lo, bar_k, hd + head * ATT_D, img_row0);
    mbar_arrive_expect_tx(bar_v, 2 * ATT_KV_BYTES);
    tma_load_2d(sV, &p.tm_kv_hi, bar_v, 2 * hd + head * ATT_D, img_row0);
    tma_load_2d(sV + ATT_KV_BYTES, &p.tm_kv_lo, bar_v, 2 * hd + head * ATT_D, img_row0);
    mbar_wait(bar_q, 0);
  }

  constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BKV, 0, 0);  // A = Q (K-major), B = K (K-major), N = 64 keys
  constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_D, 0, 1);    // A = P (K-major), B = V (MN-major), N = 64 dims

  float o[ATT_D];
#pragma unroll
  for (int d = 0; d < ATT_D; ++d) o[d] = 0.0f;
  float m_run = -INFINITY, l_run = 0.0f;

  const uint32_t sQ_u = smem_u32(sQ), sK_u = smem_u32(sK), sV_u = smem_u32(sV), sP_u = smem_u32(sP);
  const uint32_t sw = static_cast<uint32_t>(tid & 7);
  uint8_t* p_row_hi = sP + tid * 128;
  uint8_t* p_row_lo = sP + ATT_P_BYTES + tid * 128;

  for (int j = 0; j < n_kv; ++j) {
    const uint32_t ph = static_cast<uint32_t>(j & 1);
    // ---- S = Q K^T
    if (tid == 0) {
      mbar_wait(bar_k, ph);
      tc_fence_after();
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = sQ_u + ((pass == 2) ? ATT_Q_BYTES : 0);
        const uint32_t b = sK_u + ((pass == 1) ? ATT_KV_BYTES : 0);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          tc_mma_f16(tmem_s, umma_desc_sw128(a + k * 32), umma_desc_sw128(b + k * 32), idesc_qk,
                     (pass | k) != 0 ? 1u : 0u);
      }
      tc_commit(bar_s);
    }
    mbar_wait(bar_s, ph);
    tc_fence_after();
    if (tid == 0 && j + 1 < n_kv) {  // K buffer drained: prefetch the next K tile under the softmax
      mbar_arrive_expect_tx(bar_k, 2 * ATT_KV_BYTES);
      tma_load_2d(sK, &p.tm_kv_hi, bar_k, hd + head * ATT_D, img_row0 + (j + 1) * ATT_BKV);
      tma_load_2d(sK + ATT_KV_BYTES, &p.tm_kv_lo, bar_k, hd + head * ATT_D, img_row0 + (j + 1) * ATT_BKV);
    }

    // ---- softmax on this thread's row
    uint32_t sv[2][32];
    tmem_ld_32x32(tmem_s + lane_base, sv[0]);
    tmem_ld_32x32(tmem_s + lane_base + 32, sv[1]);
    tc_wait_ld();
    const int kv_valid = sg.tokens - j * ATT_BKV;  // keys >= kv_valid belong to another image / padding
    float s[ATT_BKV];
    float m_tile = -INFINITY;
#pragma unroll
    for (int c = 0; c < ATT_BKV; ++c) {
      const float x = __uint_as_float(sv[c >> 5][c & 31]);
      s[c] = (c < kv_valid) ? x : -INFINITY;
      m_tile = fmaxf(m_tile, s[c]);
    }
    const float m_new = fmaxf(m_run, m_tile);
    const float alpha = exp2f((m_run - m_new) * p.scale_log2e);
    const float mb = m_new * p.scale_log2e;
    float l_tile = 0.0f;
#pragma unroll
    for (int c8 = 0; c8 < ATT_BKV / 8; ++c8) {
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p0 = exp2f(fmaf(s[c8 * 8 + 2 * e], p.scale_log2e, -mb));
        const float p1 = exp2f(fmaf(s[c8 * 8 + 2 * e + 1], p.scale_log2e, -mb));
        l_tile += p0 + p1;
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(p0, h0, l0);
        split_bf16(p1, h1, l1);
        hw[e] = pack_bf16(h0, h1);
        lw[e] = pack_bf16(l0, l1);
      }
      const uint32_t off = (static_cast<uint32_t>(c8) ^ sw) * 16;  // Swizzle<3,4,3>: 16-byte chunk ^= row % 8
      *reinterpret_cast<uint4*>(p_row_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(p_row_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
    l_run = l_run * alpha + l_tile;
    m_run = m_new;
    fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core's async proxy
    tc_fence_before();
    __syncthreads();

    // ---- O_tile = P V
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(bar_v, ph);
      tc_fence_after();
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = sP_u + ((pass == 2) ? ATT_P_BYTES : 0);
        const uint32_t b = sV_u + ((pass == 1) ? ATT_KV_BYTES : 0);
#pragma unroll
        for (int k = 0; k < ATT_BKV / 16; ++k)
          tc_mma_f16(tmem_o, umma_desc_sw128(a + k * 32), umma_desc_sw128(b + k * 2048), idesc_pv,
                     (pass | k) != 0 ? 1u : 0u);
      }
      tc_commit(bar_o);
    }
    mbar_wait(bar_o, ph);
    tc_fence_after();
    if (tid == 0 && j + 1 < n_kv) {  // V and P buffers drained
      mbar_arrive_expect_tx(bar_v, 2 * ATT_KV_BYTES);
      tma_load_2d(sV, &p.tm_kv_hi, bar_v, 2 * hd + head * ATT_D, img_row0 + (j + 1) * ATT_BKV);
      tma_load_2d(sV + ATT_KV_BYTES, &p.tm_kv_lo, bar_v, 2 * hd + head * ATT_D, img_row0 + (j + 1) * ATT_BKV);
    }
    uint32_t ov[2][32];
    tmem_ld_32x32(tmem_o + lane_base, ov[0]);
    tmem_ld_32x32(tmem_o + lane_base + 32, ov[1]);
    tc_wait_ld();
#pragma unroll
    for (int d = 0; d < ATT_D; ++d) o[d] = fmaf(o[d], alpha, __uint_as_float(ov[d >> 5][d & 31]));
    tc_fence_before();  // orders these TMEM reads before the next iteration's MMAs (issued after a __syncthreads)
  }

  // ---- normalise and store this thread's row as split bf16
  const int qrow = qt * ATT_BQ + tid;
  if (qrow < sg.tokens) {
    const float inv = 1.0f / l_run;
    const long off = static_cast<long>(img_row0 + qrow) * hd + head * ATT_D;
    __nv_bfloat16* oh = p.out_hi + off;
    __nv_bfloat16* ol = p.out_lo + off;
#pragma unroll
    for (int d8 = 0; d8 < ATT_D / 8; ++d8) {
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(o[d8 * 8 + 2 * e] * inv, h0, l0);
        split_bf16(o[d8 * 8 + 2 * e + 1] * inv, h1, l1);
        hw[e] = pack_bf16(h0, h1);
        lw[e] = pack_bf16(l0, l1);
      }
      *reinterpret_cast<uint4*>(oh + d8 * 8) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(ol + d8 * 8) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
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
  if ((rc = make_tmap_bf16_2d(&P.tm_q_hi, a->qkv_hi, a->M, 3 * hd, 3 * hd, ATT_BQ))) return rc;
  if ((rc = make_tmap_bf16_2d(&P.tm_q_lo, a->qkv_lo, a->M, 3 * hd, 3 * hd, ATT_BQ))) return rc;
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
    total += sg.batch * a->heads * cdiv(sg.tokens, ATT_BQ);
  }
  P.cta_start[a->nseg] = total;
  P.nseg = a->nseg;
  P.heads = a->heads;
  P.scale_log2e = a->scale * 1.44269504088896340736f;
  P.out_hi = static_cast<__nv_bfloat16*>(a->out_hi);
  P.out_lo = static_cast<__nv_bfloat16*>(a->out_lo);
  static bool attr_set = false;
  if (!attr_set) {
    DUPL_CUDA_OK(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  attention_fwd_kernel<<<total, ATT_THREADS, ATT_SMEM, static_cast<cudaStream_t>(stream)>>>(P);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}
