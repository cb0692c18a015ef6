// Persistent, warp-specialised tcgen05 GEMM for the ViT linear layers:
//   C[M,N] = A[M,K] * W[N,K]^T (+ fused epilogue), fp32 result.
//
// Precision: every operand is carried as two bf16 planes (hi, lo) and the product is formed as
// hi*hi + hi*lo + lo*hi on the tensor cores with fp32 accumulation in TMEM ("bf16x3").  A single
// bf16/fp16/tf32 pass moves the normalised CAMs by 2e-3 (tools/precision_study.py), above the 1e-3
// parity bar of the path; the split keeps the error at ~2e-5 for 3 MMAs per k-step.
//
// Structure (one CTA per SM, 192 threads, CTA PAIRS = clusters of 2, tcgen05 cta_group::2):
//   a pair computes a 256 x BN output tile; each CTA stages its 128 rows of A and HALF of the W tile
//   (BN/2 rows) in its own shared memory, and owns 128 rows of the accumulator in its own TMEM.  With
//   both operands in shared memory a single-CTA 128 x 256 x 16 MMA reads 12 KB of shared memory per
//   128 clocks (94 B/clk with the 3 passes) while TMA writes another 64 B/clk: the single-CTA version
//   sat at ~60 % tensor-pipe utilisation on the 128 B/clk shared-memory port.  The pair halves the W
//   reads per SM (62 B/clk) and the smaller stage leaves room for 3 pipeline stages.
//   warp 0     TMA producer (both CTAs): A_hi/A_lo [128 x BK], W_hi/W_lo [BN/2 x BK]; the bytes of both
//              CTAs are accounted on the leader's full barrier
//   warp 1     MMA issuer (leader CTA only): 3 x BK/16 tcgen05.mma.cta_group::2 (256 x BN x 16) per
//              k-block; commits are multicast to both CTAs; accumulators double-buffered in TMEM so
//              the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2-5  epilogue: tcgen05.ld (32 lanes x 32 columns per instruction), bias / GELU /
//              residual / split / patch-embed row remap, 128-bit global stores
// The M dimension concatenates every image of every scale (and the grid covers both students), so
// tile-quantisation loss stays below 1 % although 148 is an awkward SM count.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"

namespace dupl {

constexpr int GEMM_BM = 128;
constexpr int GEMM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-5 and 6-9 epilogue (two warps per TMEM lane quadrant)

struct GemmGroupDev {
  CUtensorMap tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo;
  const float* bias;
  const float* resid;
  float* out_f32;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  const float* pos[DUPL_MAX_SEGMENTS];
};

struct GemmParamsDev {
  GemmGroupDev g[DUPL_MAX_GROUPS];
  dupl_segment seg[DUPL_MAX_SEGMENTS];
  int groups, M, N, K, ldo, epilogue, nseg;
  int ksplit;    // split-K factor: work item (group, k-split, tile); partial s lands at out_f32 + s*M*ldo (EPI_F32 only)
  int kper;      // k-blocks per split
  int f32_rows;  // GELU_SPLIT: rows below this get the fp32 pre-activation side output
  int passes;    // 3: hi*hi + hi*lo + lo*hi;  4: + lo*lo (products whose SIGN is consumed downstream: PTC Gram)
  int a_mn;      // A planes stored [K, M]: MN-major operand, staged as 64-column blocks of [BK rows x 128 B] (BK = 64 only)
  int b_mn;      // W planes stored [K, N]: same for B (BN = 256 or 128: each CTA stages BN/2 = 128 or 64 columns)
  int epi_halves;  // 2: warps 6-9 take every other 32-column chunk of the epilogue; 1: warps 2-5 alone (RESID always)
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// BK = k-block depth in bf16 elements: 64 (128-byte swizzled rows) or 32 (64-byte rows, twice the stages).
template <int BN, int BK>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * BK * 2;  // one plane
  static constexpr int B_BYTES = (BN / 2) * BK * 2;  // this CTA's half of the W tile, one plane
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int MN_BLOCK = BK * 128;           // MN-major staging: [BK rows x 64 columns] per block
  static constexpr int EPI_STAGE_BYTES = 4 * 32 * 33 * 4;  // per epilogue warp: 32 x 33 fp32 tile for the residual rows
  static constexpr int MAX_STAGES = (227 * 1024 - 2048 - EPI_STAGE_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : ((2 * BN <= 64) ? 64 : ((2 * BN <= 128) ? 128 : ((2 * BN <= 256) ? 256 : 512)));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_STAGE_BYTES;
};

template <int BN, int BK>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_bf16x3_kernel(const __grid_constant__ GemmParamsDev p) {
  using Cfg = GemmCfg<BN, BK>;
  pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* stage_buf = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < p.groups; ++g) {
      tma_prefetch_desc(&p.g[g].tm_a_hi);
      tma_prefetch_desc(&p.g[g].tm_a_lo);
      tma_prefetch_desc(&p.g[g].tm_b_hi);
      tma_prefetch_desc(&p.g[g].tm_b_lo);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::STAGES; ++s) {
        mbar_init(&full_bar[s], 1);   // used in the leader CTA only (its producer's arrive.expect_tx)
        mbar_init(&empty_bar[s], 1);  // one multicast commit per k-block
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tmem_full[a], 1);
        mbar_init(&tmem_empty[a], 256 * p.epi_halves);  // leader's copy collects the epilogue threads of BOTH CTAs
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Work item = (group, pair of M tiles, N tile); CTA `rank` of the cluster takes M tile 2*pair + rank.
  const int rank = static_cast<int>(cluster_ctarank());
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int tiles_per_group = m_pairs * n_tiles;
  const int total_tiles = tiles_per_group * p.groups * p.ksplit;
  const int k_blocks = (p.K + BK - 1) / BK;  // a ragged last block exists only with two MN-major operands: TMA fills rows >= K with zeros
  // work item t -> (group g, k-split ks, tile r): gs = t / tiles_per_group, g = gs / ksplit, ks = gs % ksplit

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        const int gs = t / tiles_per_group;
        const int r = t - gs * tiles_per_group;
        const int g = gs / p.ksplit, ks = gs - g * p.ksplit;
        const int m0 = (2 * (r / n_tiles) + rank) * GEMM_BM;
        const int n0 = (r % n_tiles) * BN;
        const GemmGroupDev& G = p.g[g];
        const int kb_end = min(k_blocks, (ks + 1) * p.kper);
        for (int kb = ks * p.kper; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* s = smem + stage * Cfg::STAGE_BYTES;
          const uint32_t lead_full = mapa_u32(&full_bar[stage], 0);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);  // both CTAs' bytes
          if (!p.a_mn) {
            tma_load_2d_pair(s, &G.tm_a_hi, lead_full, kb * BK, m0);
            tma_load_2d_pair(s + Cfg::A_BYTES, &G.tm_a_lo, lead_full, kb * BK, m0);
          } else {
            // MN-major: the stored matrix is [K, M]; a box is 64 columns (128 B) x BK rows = one 8 KB block per 64 tile rows
#pragma unroll
            for (int j = 0; j < GEMM_BM / 64; ++j) {
              tma_load_2d_pair(s + j * Cfg::MN_BLOCK, &G.tm_a_hi, lead_full, m0 + 64 * j, kb * BK);
              tma_load_2d_pair(s + Cfg::A_BYTES + j * Cfg::MN_BLOCK, &G.tm_a_lo, lead_full, m0 + 64 * j, kb * BK);
            }
          }
          const int nb0 = n0 + rank * (BN / 2);
          if (!p.b_mn) {
            tma_load_2d_pair(s + 2 * Cfg::A_BYTES, &G.tm_b_hi, lead_full, kb * BK, nb0);
            tma_load_2d_pair(s + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &G.tm_b_lo, lead_full, kb * BK, nb0);
          } else {
#pragma unroll
            for (int j = 0; j < (BN / 2) / 64; ++j) {
              tma_load_2d_pair(s + 2 * Cfg::A_BYTES + j * Cfg::MN_BLOCK, &G.tm_b_hi, lead_full, nb0 + 64 * j, kb * BK);
              tma_load_2d_pair(s + 2 * Cfg::A_BYTES + Cfg::B_BYTES + j * Cfg::MN_BLOCK, &G.tm_b_lo, lead_full, nb0 + 64 * j, kb * BK);
            }
          }
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA; whole warp, lane elected per op)
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(BN, 0, 0, 256) | (p.a_mn ? (1u << 15) : 0u) | (p.b_mn ? (1u << 16) : 0u);
      // descriptor step per UMMA_K = 16: 32 B along a K-major row, 16 rows x 128 B of an MN-major block
      const uint32_t a_step = p.a_mn ? 2048u : 32u, b_step = p.b_mn ? 2048u : 32u;
      const uint32_t tmem_acc = __shfl_sync(0xffffffffu, tmem_base, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_acc + acc * BN;
        const int ks = (t / tiles_per_group) % p.ksplit;
        const int kb_count = min(k_blocks, (ks + 1) * p.kper) - ks * p.kper;
        for (int kb = 0; kb < kb_count; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t s = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          auto desc = [](uint32_t addr, int mn) {
            return mn ? umma_desc_sw128_mn(addr, Cfg::MN_BLOCK) : (BK == 64 ? umma_desc_sw128(addr) : umma_desc_sw64(addr));
          };
          const uint64_t a_hi = desc(s, p.a_mn), a_lo = desc(s + Cfg::A_BYTES, p.a_mn);
          const uint64_t b_hi = desc(s + 2 * Cfg::A_BYTES, p.b_mn), b_lo = desc(s + 2 * Cfg::A_BYTES + Cfg::B_BYTES, p.b_mn);
#pragma unroll
          for (int pass = 0; pass < 4; ++pass) {
            if (pass >= p.passes) break;
            const uint64_t a = (pass >= 2) ? a_lo : a_hi;
            const uint64_t b = (pass == 1 || pass == 3) ? b_lo : b_hi;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc_mma_f16_pair(d_tmem, umma_desc_advance(a, k * a_step), umma_desc_advance(b, k * b_step), idesc,
                              (kb | pass | k) != 0 ? 1u : 0u);
          }
          tc_commit_pair(&empty_bar[stage]);  // stage drained in both CTAs: tell both producers
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit_pair(&tmem_full[acc]);  // accumulator complete in both CTAs
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (((warp - 2) >> 2) < p.epi_halves) {
    // ------------------------------------------------------------------ epilogue (warps 2..5, and 6..9 when epi_halves == 2)
    // With 4 warps the GELU(erf) + split epilogue of a 256-wide tile (256 elements x ~50 instructions per thread) takes
    // ~3/4 of the tile's MMA time at K = 768 and every hiccup shows: ncu saw the tensor pipe 63 % (fc1) / 76 % (qkv) busy
    // on those launches against 93 % with the light residual epilogue.  Two warps per TMEM lane quadrant split the
    // 32-column chunks between them.
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = cluster_id; t < total_tiles; t += num_clusters) {
      const int gs = t / tiles_per_group;
      const int r = t - gs * tiles_per_group;
      const int g = gs / p.ksplit, ks = gs - g * p.ksplit;
      const int m0 = (2 * (r / n_tiles) + rank) * GEMM_BM;
      const int n0 = (r % n_tiles) * BN;
      const GemmGroupDev& G = p.g[g];
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;

      // PATCH epilogue: patch row -> token row (+1 for the cls token of each image) and pos-embed row.
      long out_row = row;
      const float* pos_row = nullptr;
      if (p.epilogue == DUPL_EPI_PATCH && row_ok) {
        int si = 0;
        for (int s = 1; s < p.nseg; ++s)
          if (row >= p.seg[s].patch_row_offset) si = s;
        const int np = p.seg[si].tokens - 1;
        const int local = row - p.seg[si].patch_row_offset;
        const int img = local / np;
        const int pidx = local - img * np;
        out_row = p.seg[si].row_offset + static_cast<long>(img) * p.seg[si].tokens + 1 + pidx;
        pos_row = G.pos[si] + static_cast<long>(1 + pidx) * p.N;
      }

      // RESID: the accumulator comes out of TMEM one row per thread, so a thread-per-row read of the residual touches 32
      // rows (16 B each) per instruction with 8 loads in flight: measured, that read doubled the N = 768, K = 768 launches
      // (proj 129 us vs 67 us without the residual).  The rows are therefore read with lane = column (32 independent
      // 128-byte loads per thread in flight) and transposed through a per-warp shared-memory tile.
      float* stg = stage_buf + (warp - 2) * (32 * 33);
      const int rows_here = min(32, p.M - (m0 + q * 32));  // warp-uniform; <= 0 when the quadrant is padding
      float rv[32];  // residual rows of the chunk being fetched (software pipeline: chunk c+1 loads while chunk c is written)
      auto fetch_resid = [&](int c0n) {
        const int coln = n0 + c0n;
        if (coln < p.N && rows_here > 0) {
          const float* rbase = G.resid + static_cast<long>(m0 + q * 32) * p.ldo + coln + lane;
          const bool lane_ok = lane < p.N - coln;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) rv[rr] = (rr < rows_here && lane_ok) ? rbase[static_cast<long>(rr) * p.ldo] : 0.0f;
        }
      };
      if (p.epilogue == DUPL_EPI_RESID) fetch_resid(0);  // in flight while the MMAs of this tile finish
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 32 * half; c0 < BN; c0 += 32 * p.epi_halves) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c0, v);
        const int col0 = n0 + c0;
        if (p.epilogue == DUPL_EPI_RESID) {
          __syncwarp();
          if (col0 < p.N && rows_here > 0) {
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) stg[rr * 33 + lane] = rv[rr];
          }
          __syncwarp();
          if (c0 + 32 < BN) fetch_resid(c0 + 32);
        }
        tc_wait_ld();
        if (row_ok && col0 < p.N) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (G.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(G.bias + col0 + j));
              f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
            }
          }
          const int ncols = min(32, p.N - col0);  // multiple of 16 by contract
          if (p.epilogue == DUPL_EPI_F32 || p.epilogue == DUPL_EPI_RESID || p.epilogue == DUPL_EPI_PATCH) {
            float* o = G.out_f32 + (static_cast<long>(ks) * p.M + out_row) * p.ldo + col0;
            const float* add = nullptr;
            if (p.epilogue == DUPL_EPI_PATCH) add = pos_row + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j < ncols) {
                float4 o4 = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                if (p.epilogue == DUPL_EPI_RESID) {
                  const float* sr = stg + lane * 33 + j;
                  o4.x += sr[0]; o4.y += sr[1]; o4.z += sr[2]; o4.w += sr[3];
                } else if (add != nullptr) {
                  const float4 a4 = *reinterpret_cast<const float4*>(add + j);
                  o4.x += a4.x; o4.y += a4.y; o4.z += a4.z; o4.w += a4.w;
                }
                *reinterpret_cast<float4*>(o + j) = o4;
              }
            }
          } else {
            if (p.epilogue == DUPL_EPI_GELU_SPLIT) {
              if (G.out_f32 != nullptr && row < p.f32_rows) {  // training: keep the pre-activation for the GELU backward
                float* o = G.out_f32 + static_cast<long>(row) * p.ldo + col0;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  if (j < ncols) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
            }
            if (p.epilogue == DUPL_EPI_RELU_SPLIT) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
            }
            __nv_bfloat16* oh = G.out_hi + static_cast<long>(row) * p.ldo + col0;
            __nv_bfloat16* ol = G.out_lo + static_cast<long>(row) * p.ldo + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (j < ncols) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split2_bf16(f[j + 2 * e], f[j + 2 * e + 1], hw[e], lw[e]);
                *reinterpret_cast<uint4*>(oh + j) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(ol + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(mapa_u32(&tmem_empty[acc], 0));  // the leader's MMA warp waits for both CTAs' epilogues
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();  // no CTA leaves while its peer may still read its shared memory or arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

// Deterministic split-K tail: out[i] = sum_s ws[s][i] in fixed order.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float4* __restrict__ ws, float4* __restrict__ out, int ks,
                                                            long n4) {
  pdl_sync();
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= n4) return;
  float4 a = __ldcs(ws + i);
  for (int s = 1; s < ks; ++s) {
    const float4 b = __ldcs(ws + s * n4 + i);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  out[i] = a;
}

template <int BN, int BK>
static int launch_gemm(const GemmParamsDev& P, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, BK>;
  static bool attr_set = false;
  if (!attr_set) {
    DUPL_CUDA_OK(cudaFuncSetAttribute(gemm_bf16x3_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int m_pairs = cdiv(cdiv(P.M, GEMM_BM), 2), n_tiles = cdiv(P.N, BN);
  const int total = m_pairs * n_tiles * P.groups * P.ksplit;  // work items, one per cluster of 2 CTAs
  const int clusters = total < gemm_sms() / 2 ? total : gemm_sms() / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  DUPL_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_bf16x3_kernel<BN, BK>, P));
  count_launch();
  return DUPL_OK;
}

// Tile width and split-K factor by a small cost model (microseconds; constants from the round-1 profiles: one
// 256-wide k-block of 3 x 4 MMAs takes ~0.9 us, a 256-wide fp32 epilogue ~3 us and is hidden behind the next
// tile's main loop except for the last tile of a CTA pair).  The persistent grid has sm_count/2 pairs; 148 SMs
// make 74 = 2 x 37 slots, so the M=3140 training shapes (39 / 117 / 156 tiles of 256) fit badly without this.
static void choose_tiling(int M, int N, int k_blocks, int groups, int max_split, bool allow_192, int& bn_out, int& ks_out,
                          int& kper_out) {
  const int slots = gemm_sms() / 2;
  const int m_pairs = cdiv(cdiv(M, GEMM_BM), 2);
  double best = 1e30;
  bn_out = 256; ks_out = 1; kper_out = k_blocks;
  const int cand[3] = {256, 192, 128};
  // smaller tiles re-read more operand bytes per MMA from shared memory (DUPL_GEMM_PEN192 / _PEN128 override for tuning)
  static const double pen192 = getenv("DUPL_GEMM_PEN192") ? atof(getenv("DUPL_GEMM_PEN192")) : 1.08;
  static const double pen128 = getenv("DUPL_GEMM_PEN128") ? atof(getenv("DUPL_GEMM_PEN128")) : 1.25;
  const double pen[3] = {1.0, pen192, pen128};
  for (int c = 0; c < 3; ++c) {
    const int bn = cand[c];
    if (bn == 192 && !allow_192) continue;
    const int n_tiles = cdiv(N, bn);
    for (int ks = 1; ks <= max_split && ks <= k_blocks; ++ks) {
      const int kper = cdiv(k_blocks, ks);
      if ((ks - 1) * kper >= k_blocks) continue;  // every split needs at least one k-block
      const long tiles = static_cast<long>(m_pairs) * n_tiles * groups * ks;
      const long waves = (tiles + slots - 1) / slots;
      double t = waves * kper * 0.9 * (bn / 256.0) * pen[c] + 3.0 * (bn / 256.0);
      if (ks > 1) t += 3.0 + (ks + 1.0) * M * static_cast<double>(N) * 4.0 * groups / 5e6;
      if (t < best - 1e-9) {
        best = t; bn_out = bn; ks_out = ks; kper_out = kper;
      }
    }
  }
}

}  // namespace dupl

extern "C" int dupl_gemm_plan(int32_t M, int32_t N, int32_t K, int32_t groups, int32_t max_ksplit, int32_t* tile_n,
                              int32_t* ksplit, int32_t* work_items) {
  using namespace dupl;
  DUPL_CHECK_ARG(M > 0 && N > 0 && K > 0 && K % 64 == 0 && groups >= 1 && tile_n && ksplit && work_items,
                 "dupl_gemm_plan: bad arguments");
  int bn, ks = 1, kper = K / 64;
  if (N < 256) bn = (N > 64) ? 128 : 64;
  else choose_tiling(M, N, K / 64, groups, max_ksplit > 1 ? max_ksplit : 1, true, bn, ks, kper);
  *tile_n = bn;
  *ksplit = ks;
  *work_items = cdiv(cdiv(M, GEMM_BM), 2) * cdiv(N, bn) * groups * ks;
  return DUPL_OK;
}

extern "C" int dupl_gemm_bf16x3(const dupl_gemm_args* a, void* stream) {
  using namespace dupl;
  DUPL_CHECK_ARG(a != nullptr, "dupl_gemm_bf16x3: args is NULL");
  DUPL_CHECK_ARG(a->groups >= 1 && a->groups <= DUPL_MAX_GROUPS, "dupl_gemm_bf16x3: groups=%d", a->groups);
  DUPL_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "dupl_gemm_bf16x3: empty problem %dx%dx%d", a->M, a->N, a->K);
  const bool a_mn = a->a_mn_major != 0, b_mn = a->b_mn_major != 0;
  DUPL_CHECK_ARG(a->K % 64 == 0 || (a_mn && b_mn), "dupl_gemm_bf16x3: K=%d must be a multiple of 64 (unless both operands are MN-major)",
                 a->K);
  // k-block depth: 64 (128-byte swizzle, 3 stages) and 32 (64-byte swizzle, 7 stages) measure the same on
  // B200 (the kernel is not latency bound); DUPL_GEMM_BK=32 selects the deeper pipeline for experiments.
  static const int bk = [] {
    const char* e = getenv("DUPL_GEMM_BK");
    return (e != nullptr && atoi(e) == 32) ? 32 : 64;
  }();
  // DUPL_GEMM_TILING=fixed restores the round-1 behaviour (256-wide tiles, no split-K) for A/B measurements.
  static const bool fixed_tiling = [] {
    const char* e = getenv("DUPL_GEMM_TILING");
    return e != nullptr && strcmp(e, "fixed") == 0;
  }();
  DUPL_CHECK_ARG(a->N % 16 == 0, "dupl_gemm_bf16x3: N=%d must be a multiple of 16", a->N);
  DUPL_CHECK_ARG(a->lda % 8 == 0 && a->lda >= (a_mn ? a->M : a->K), "dupl_gemm_bf16x3: lda=%d", a->lda);
  DUPL_CHECK_ARG(a->ldw == 0 || (a->ldw % 8 == 0 && a->ldw >= (b_mn ? a->N : a->K)), "dupl_gemm_bf16x3: ldw=%d", a->ldw);
  DUPL_CHECK_ARG(!(a_mn || b_mn) || bk == 64, "dupl_gemm_bf16x3: MN-major operands need the 64-deep k-block");
  DUPL_CHECK_ARG(!b_mn || a->N >= 128, "dupl_gemm_bf16x3: an MN-major W needs N >= 128 (N=%d)", a->N);
  DUPL_CHECK_ARG(!(a_mn || b_mn) || a->epilogue != DUPL_EPI_PATCH, "dupl_gemm_bf16x3: the PATCH epilogue takes K-major operands");
  DUPL_CHECK_ARG(a->ldo % 8 == 0 && a->ldo >= a->N, "dupl_gemm_bf16x3: ldo=%d", a->ldo);
  DUPL_CHECK_ARG(a->epilogue >= DUPL_EPI_F32 && a->epilogue <= DUPL_EPI_RELU_SPLIT, "dupl_gemm_bf16x3: epilogue=%d",
                 a->epilogue);
  DUPL_CHECK_ARG(a->max_ksplit >= 0 && a->max_ksplit <= DUPL_MAX_KSPLIT, "dupl_gemm_bf16x3: max_ksplit=%d (0..%d)",
                 a->max_ksplit, DUPL_MAX_KSPLIT);
  const bool can_split = a->max_ksplit > 1 && a->epilogue == DUPL_EPI_F32;
  if (a->max_ksplit > 1) {
    DUPL_CHECK_ARG(a->epilogue == DUPL_EPI_F32, "dupl_gemm_bf16x3: split-K needs the plain F32 epilogue");
    for (int g = 0; g < a->groups; ++g)
      DUPL_CHECK_ARG(a->g[g].splitk_ws != nullptr && a->g[g].bias == nullptr,
                     "dupl_gemm_bf16x3: split-K needs splitk_ws and no bias (group %d)", g);
  }
  GemmParamsDev P;
  memset(&P, 0, sizeof(P));
  P.groups = a->groups; P.M = a->M; P.N = a->N; P.K = a->K; P.ldo = a->ldo; P.epilogue = a->epilogue;
  P.f32_rows = a->f32_rows > 0 ? a->f32_rows : a->M;
  DUPL_CHECK_ARG(a->passes == 0 || a->passes == 3 || a->passes == 4, "dupl_gemm_bf16x3: passes=%d (3 or 4)", a->passes);
  P.passes = a->passes == 4 ? 4 : 3;
  P.a_mn = a_mn ? 1 : 0;
  P.b_mn = b_mn ? 1 : 0;
  // the residual epilogue stages its rows through 4 per-warp shared-memory tiles and is light anyway: 4 warps
  static const int epi_halves = (getenv("DUPL_GEMM_EPI_WARPS") && atoi(getenv("DUPL_GEMM_EPI_WARPS")) == 4) ? 1 : 2;
  P.epi_halves = a->epilogue == DUPL_EPI_RESID ? 1 : epi_halves;
  P.nseg = 0;
  if (a->epilogue == DUPL_EPI_PATCH) {
    DUPL_CHECK_ARG(a->nseg >= 1 && a->nseg <= DUPL_MAX_SEGMENTS, "dupl_gemm_bf16x3: nseg=%d", a->nseg);
    P.nseg = a->nseg;
    for (int s = 0; s < a->nseg; ++s) P.seg[s] = a->seg[s];
  }
  // Wide tiles for wide outputs; N <= 128 (CAM-sized heads) uses the narrow instantiations.
  const int k_blocks = cdiv(a->K, bk);
  int bn, ks = 1, kper = k_blocks;
  if (a->N < 256 || fixed_tiling || bk != 64) {
    bn = (a->N >= 256) ? 256 : ((a->N > 64) ? 128 : 64);
  } else {
    // an MN-major W is staged in 64-column blocks per CTA: 256- and 128-wide tiles only
    choose_tiling(a->M, a->N, k_blocks, a->groups, can_split ? a->max_ksplit : 1, !b_mn, bn, ks, kper);
  }
  P.ksplit = ks;
  P.kper = kper;
  const int ldw = a->ldw > 0 ? a->ldw : (b_mn ? a->N : a->K);
  for (int g = 0; g < a->groups; ++g) {
    const dupl_gemm_group& G = a->g[g];
    DUPL_CHECK_ARG(G.a_hi && G.a_lo && G.w_hi && G.w_lo, "dupl_gemm_bf16x3: NULL operand plane in group %d", g);
    const bool f32_out = a->epilogue == DUPL_EPI_F32 || a->epilogue == DUPL_EPI_RESID || a->epilogue == DUPL_EPI_PATCH;
    DUPL_CHECK_ARG(!f32_out || G.out_f32, "dupl_gemm_bf16x3: out_f32 is NULL in group %d", g);
    DUPL_CHECK_ARG(f32_out || (G.out_hi && G.out_lo), "dupl_gemm_bf16x3: out_hi/out_lo is NULL in group %d", g);
    DUPL_CHECK_ARG(a->epilogue != DUPL_EPI_RESID || G.resid, "dupl_gemm_bf16x3: resid is NULL in group %d", g);
    int rc;
    if (!a_mn) {
      if ((rc = make_tmap_bf16_2d(&P.g[g].tm_a_hi, G.a_hi, a->M, a->K, a->lda, GEMM_BM, bk))) return rc;
      if ((rc = make_tmap_bf16_2d(&P.g[g].tm_a_lo, G.a_lo, a->M, a->K, a->lda, GEMM_BM, bk))) return rc;
    } else {  // stored [K, M]: boxes of 64 columns x bk rows; rows >= K and columns >= M read as zeros
      if ((rc = make_tmap_bf16_2d(&P.g[g].tm_a_hi, G.a_hi, a->K, a->M, a->lda, bk, 64))) return rc;
      if ((rc = make_tmap_bf16_2d(&P.g[g].tm_a_lo, G.a_lo, a->K, a->M, a->lda, bk, 64))) return rc;
    }
    if (!b_mn) {
      if ((rc = make_tmap_bf16_2d(&P.g[g].tm_b_hi, G.w_hi, a->N, a->K, ldw, bn / 2, bk))) return rc;
      if ((rc = make_tmap_bf16_2d(&P.g[g].tm_b_lo, G.w_lo, a->N, a->K, ldw, bn / 2, bk))) return rc;
    } else {
      if ((rc = make_tmap_bf16_2d(&P.g[g].tm_b_hi, G.w_hi, a->K, a->N, ldw, bk, 64))) return rc;
      if ((rc = make_tmap_bf16_2d(&P.g[g].tm_b_lo, G.w_lo, a->K, a->N, ldw, bk, 64))) return rc;
    }
    P.g[g].bias = G.bias; P.g[g].resid = G.resid;
    P.g[g].out_f32 = ks > 1 ? G.splitk_ws : G.out_f32;
    P.g[g].out_hi = static_cast<__nv_bfloat16*>(G.out_hi);
    P.g[g].out_lo = static_cast<__nv_bfloat16*>(G.out_lo);
    for (int s = 0; s < DUPL_MAX_SEGMENTS; ++s) P.g[g].pos[s] = G.pos[s];
    if (a->epilogue == DUPL_EPI_PATCH)
      for (int s = 0; s < a->nseg; ++s)
        DUPL_CHECK_ARG(G.pos[s] != nullptr, "dupl_gemm_bf16x3: pos[%d] is NULL in group %d", s, g);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if (bk == 64) {
    if (bn == 256) rc = launch_gemm<256, 64>(P, st);
    else if (bn == 192) rc = launch_gemm<192, 64>(P, st);
    else if (bn == 128) rc = launch_gemm<128, 64>(P, st);
    else rc = launch_gemm<64, 64>(P, st);
  } else {
    if (bn == 256) rc = launch_gemm<256, 32>(P, st);
    else if (bn == 128) rc = launch_gemm<128, 32>(P, st);
    else rc = launch_gemm<64, 32>(P, st);
  }
  if (rc) return rc;
  if (ks > 1) {
    const long n4 = static_cast<long>(a->M) * a->ldo / 4;
    for (int g = 0; g < a->groups; ++g) {
      DUPL_CUDA_OK(launch_pdl(splitk_reduce_kernel, dim3(static_cast<unsigned>((n4 + 255) / 256)), dim3(256), 0, st,
                              reinterpret_cast<const float4*>(a->g[g].splitk_ws), reinterpret_cast<float4*>(a->g[g].out_f32), ks, n4));
      count_launch();
    }
  }
  return DUPL_OK;
}
