// CAM post-processing and pseudo-label casts — HBM-bound, one pass over
// the output, low-resolution sources staged in shared memory.
#include <stdlib.h>

#include "common.cuh"
#include "resample.cuh"

namespace dupl {

// ---------------------------------------------------------------------------------------------
// multi_scale_cam2_siamese post-processing (utils/cam_helper.py:173-202).
// value(b,k,y,x) = sum_s relu(max(up(A_s)(y,x), up(B_s)(y, W-1-x)))      A = image, B = flipped twin
// out = (value - min) / ((max - min) + 1e-5)   with min/max over the (b,k) plane.
// Two launches: PASS 0 reduces min/max per plane (atomics on the non-negative floats' bit patterns),
// PASS 1 recomputes the value from the shared-memory copy of the low-res maps and writes the
// normalised output once (128-bit stores).  The 64 MB (VOC) output is written exactly once and never
// read back.
// ---------------------------------------------------------------------------------------------
struct MscamParams {
  const float* lowres[DUPL_MAX_SEGMENTS];
  int gh[DUPL_MAX_SEGMENTS], gw[DUPL_MAX_SEGMENTS];
  int nscale, b, K, H, W;
  float* out;
  unsigned int* minmax;  // [b*K][2] : min bits, max bits
};

constexpr int MSCAM_ROWS = 32;  // output rows per block
constexpr int MSCAM_THREADS = 128;

template <int PASS>
__global__ void __launch_bounds__(MSCAM_THREADS) mscam_generic_kernel(MscamParams p) {
  extern __shared__ float sm[];
  const int plane = blockIdx.x;  // img*K + k
  const int img = plane / p.K, k = plane % p.K;
  const int y_begin = blockIdx.y * MSCAM_ROWS;
  const int y_end = min(y_begin + MSCAM_ROWS, p.H);

  // Stage the low-res planes of the image and of its twin (twin stored pre-flipped along x, which
  // commutes with the symmetric align_corners=False sampling grid).
  int off[DUPL_MAX_SEGMENTS + 1];
  off[0] = 0;
  for (int s = 0; s < p.nscale; ++s) off[s + 1] = off[s] + 2 * p.gh[s] * p.gw[s];
  for (int s = 0; s < p.nscale; ++s) {
    const int n = p.gh[s] * p.gw[s];
    const float* a = p.lowres[s] + (static_cast<long>(img) * p.K + k) * n;
    const float* bt = p.lowres[s] + (static_cast<long>(img + p.b) * p.K + k) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      sm[off[s] + i] = __ldg(a + i);
      const int yy = i / p.gw[s], xx = i % p.gw[s];
      sm[off[s] + n + yy * p.gw[s] + (p.gw[s] - 1 - xx)] = __ldg(bt + i);
    }
  }
  __syncthreads();

  float mn = INFINITY, mx = 0.0f, shift = 0.0f, denom = 1.0f;
  if (PASS == 1) {
    const float lo = __uint_as_float(p.minmax[2 * plane]);
    const float hi = __uint_as_float(p.minmax[2 * plane + 1]);
    shift = -lo;                       // cam + max(-cam)
    denom = (hi + shift) + 1e-5f;      // max(cam) + 1e-5
  }
  const int groups = (p.W + 3) / 4;
  for (int g = threadIdx.x; g < groups * (y_end - y_begin); g += blockDim.x) {
    const int y = y_begin + g / groups;
    const int x0 = (g % groups) * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < p.nscale; ++s) {
      const int gh = p.gh[s], gw = p.gw[s];
      const Lin ly = lin_coord(y, gh, static_cast<float>(gh) / p.H);
      const float* A = sm + off[s];
      const float* B = A + gh * gw;
      const float sx = static_cast<float>(gw) / p.W;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const Lin lx = lin_coord(min(x0 + e, p.W - 1), gw, sx);
        const float va = bilerp_smem(A, gw, ly, lx);
        const float vb = bilerp_smem(B, gw, ly, lx);
        acc[e] += fmaxf(fmaxf(va, vb), 0.0f);
      }
    }
    if (PASS == 0) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (x0 + e < p.W) {
          mn = fminf(mn, acc[e]);
          mx = fmaxf(mx, acc[e]);
        }
    } else {
      float* o = p.out + (static_cast<long>(plane) * p.H + y) * p.W + x0;
      float r[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) r[e] = __fdiv_rn(acc[e] + shift, denom);
      if (x0 + 3 < p.W && (p.W & 3) == 0) {
        *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
      } else {
        for (int e = 0; e < 4; ++e)
          if (x0 + e < p.W) o[e] = r[e];
      }
    }
  }
  if (PASS == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {  // values are >= 0 => unsigned bit patterns order like the floats
      atomicMin(&p.minmax[2 * plane], __float_as_uint(mn));
      atomicMax(&p.minmax[2 * plane + 1], __float_as_uint(mx));
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Column-per-thread variant (NS = number of scales, 1..4; the reference uses 3).
// A thread owns one output column of a (MSCAM_TILE_ROWS x blockDim.x) tile: its horizontal sampling
// coordinates are loop invariants, and the horizontally interpolated values of the two source rows
// (t0, t1 of bilerp4) only change when the source row index advances, i.e. every H/gh output rows.  The
// inner loop is then one vertical blend per scale and twin: ~25 instructions per output pixel instead of
// ~120, which is what moves this stage from LDS/ALU-bound towards the 64 MB store stream it should be.
// Same arithmetic (same fma order) as bilerp4: results are bit-identical to the generic kernel.
// ---------------------------------------------------------------------------------------------
constexpr int MSCAM_TILE_ROWS = 64;

struct MscamRow {  // vertical sampling of one output row for one scale, row indices relative to the staged window
  int i0, i1;
  float l0, l1;
};

// a / d correctly rounded for normal-range operands (Markstein): q0 = RN(a * RN(1/d)), r = a - q0*d (exact in fma),
// q = RN(q0 + r * RN(1/d)).  The plain `/` costs ~15 instructions per pixel for its special-case handling; here
// a is in [0, max] and d = max + 1e-5, far from overflow / denormals.
__device__ __forceinline__ float div_rn_normal(float a, float d, float rcp_d) {
  const float q0 = __fmul_rn(a, rcp_d);
  const float r = __fmaf_rn(-q0, d, a);
  return __fmaf_rn(r, rcp_d, q0);
}

template <int PASS, int NS>
__global__ void __launch_bounds__(256) mscam_kernel(MscamParams p) {
  extern __shared__ float sm[];
  __shared__ int g_rlo[NS], g_nr[NS], g_clo[NS], g_nc[NS], g_off[NS + 1];
  const int plane = blockIdx.x;  // img*K + k
  const int img = plane / p.K, k = plane % p.K;
  const int y_begin = blockIdx.y * MSCAM_TILE_ROWS;
  const int y_end = min(y_begin + MSCAM_TILE_ROWS, p.H);
  const int x_begin = blockIdx.z * blockDim.x;
  const int x_last = min(x_begin + static_cast<int>(blockDim.x), p.W) - 1;

  if (threadIdx.x == 0) {
    int off = 0;
    for (int s = 0; s < NS; ++s) {
      const int gh = p.gh[s], gw = p.gw[s];
      const int rlo = lin_coord(y_begin, gh, static_cast<float>(gh) / p.H).i0;
      const int rhi = lin_coord(y_end - 1, gh, static_cast<float>(gh) / p.H).i1;
      const int clo = lin_coord(x_begin, gw, static_cast<float>(gw) / p.W).i0;
      const int chi = lin_coord(x_last, gw, static_cast<float>(gw) / p.W).i1;
      g_rlo[s] = rlo; g_nr[s] = rhi - rlo + 1; g_clo[s] = clo; g_nc[s] = chi - clo + 1;
      g_off[s] = off;
      off += 2 * g_nr[s] * g_nc[s];
    }
    g_off[NS] = off;
  }
  __syncthreads();
  // Stage the window of the low-res planes of the image (A) and of its twin (B, stored pre-flipped along x, which
  // commutes with the symmetric align_corners=False sampling grid).
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int gw = p.gw[s], n = p.gh[s] * gw;
    const int nr = g_nr[s], nc = g_nc[s], rlo = g_rlo[s], clo = g_clo[s];
    const float* a = p.lowres[s] + (static_cast<long>(img) * p.K + k) * n;
    const float* bt = p.lowres[s] + (static_cast<long>(img + p.b) * p.K + k) * n;
    float* A = sm + g_off[s];
    float* B = A + nr * nc;
    for (int i = threadIdx.x; i < nr * nc; i += blockDim.x) {
      const int r = i / nc, c = i - r * nc;
      A[i] = __ldg(a + (rlo + r) * gw + clo + c);
      B[i] = __ldg(bt + (rlo + r) * gw + (gw - 1 - (clo + c)));
    }
  }
  MscamRow* rows = reinterpret_cast<MscamRow*>(sm + ((g_off[NS] + 3) & ~3));
  float2* wts = reinterpret_cast<float2*>(rows + MSCAM_TILE_ROWS * NS);
  int* flags = reinterpret_cast<int*>(wts + MSCAM_TILE_ROWS * NS);
  for (int i = threadIdx.x; i < (y_end - y_begin) * NS; i += blockDim.x) {
    const int r = i / NS, s = i - r * NS;
    const Lin ly = lin_coord(y_begin + r, p.gh[s], static_cast<float>(p.gh[s]) / p.H);
    rows[i] = MscamRow{ly.i0 - g_rlo[s], ly.i1 - g_rlo[s], ly.l0, ly.l1};
    wts[i] = make_float2(ly.l0, ly.l1);
  }
  __syncthreads();
  for (int r = threadIdx.x; r < y_end - y_begin; r += blockDim.x) {
    int f = 0;
    for (int s = 0; s < NS; ++s)
      if (r == 0 || rows[r * NS + s].i0 != rows[(r - 1) * NS + s].i0 || rows[r * NS + s].i1 != rows[(r - 1) * NS + s].i1) f |= 1 << s;
    flags[r] = f;
  }
  __syncthreads();

  const int x = x_begin + threadIdx.x;
  const bool active = x < p.W;
  const int xc = min(x, p.W - 1);
  int c0[NS], c1[NS], nc[NS];
  float lx0[NS], lx1[NS], t0a[NS], t1a[NS], t0b[NS], t1b[NS];
  const float* Ab[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const Lin lx = lin_coord(xc, p.gw[s], static_cast<float>(p.gw[s]) / p.W);
    c0[s] = lx.i0 - g_clo[s]; c1[s] = lx.i1 - g_clo[s]; lx0[s] = lx.l0; lx1[s] = lx.l1;
    nc[s] = g_nc[s];
    Ab[s] = sm + g_off[s];
    t0a[s] = t1a[s] = t0b[s] = t1b[s] = 0.0f;
  }
  float mn = INFINITY, mx = 0.0f, shift = 0.0f, denom = 1.0f, rcp = 1.0f;
  if (PASS == 1) {
    const float lo = __uint_as_float(p.minmax[2 * plane]);
    const float hi = __uint_as_float(p.minmax[2 * plane + 1]);
    shift = -lo;                       // cam + max(-cam)
    denom = (hi + shift) + 1e-5f;      // max(cam) + 1e-5
    rcp = __frcp_rn(denom);
  }
  float* o = p.out + (static_cast<long>(plane) * p.H + y_begin) * p.W + xc;
  for (int r = 0; r < y_end - y_begin; ++r) {
    const int changed = flags[r];      // bit s: the source rows of scale s differ from the previous output row's
    if (changed != 0) {                // block-uniform and rare: every H/gh output rows
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if ((changed >> s) & 1) {
          const MscamRow rw = rows[r * NS + s];
          const float* A0 = Ab[s] + rw.i0 * nc[s];
          const float* A1 = Ab[s] + rw.i1 * nc[s];
          const float* B0 = A0 + g_nr[s] * nc[s];
          const float* B1 = A1 + g_nr[s] * nc[s];
          t0a[s] = __fmaf_rn(lx0[s], A0[c0[s]], __fmul_rn(lx1[s], A0[c1[s]]));
          t1a[s] = __fmaf_rn(lx0[s], A1[c0[s]], __fmul_rn(lx1[s], A1[c1[s]]));
          t0b[s] = __fmaf_rn(lx0[s], B0[c0[s]], __fmul_rn(lx1[s], B0[c1[s]]));
          t1b[s] = __fmaf_rn(lx0[s], B1[c0[s]], __fmul_rn(lx1[s], B1[c1[s]]));
        }
      }
    }
    float acc = 0.0f;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const float2 l = wts[r * NS + s];
      const float va = __fmaf_rn(l.x, t0a[s], __fmul_rn(l.y, t1a[s]));
      const float vb = __fmaf_rn(l.x, t0b[s], __fmul_rn(l.y, t1b[s]));
      acc += fmaxf(fmaxf(va, vb), 0.0f);
    }
    if (PASS == 0) {
      mn = fminf(mn, acc);
      mx = fmaxf(mx, acc);
    } else if (active) {
      __stcs(o + static_cast<long>(r) * p.W, div_rn_normal(acc + shift, denom, rcp));
    }
  }
  if (PASS == 0) {
    if (!active) { mn = INFINITY; mx = 0.0f; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    if ((threadIdx.x & 31) == 0) {  // values are >= 0 => unsigned bit patterns order like the floats
      atomicMin(&p.minmax[2 * plane], __float_as_uint(mn));
      atomicMax(&p.minmax[2 * plane + 1], __float_as_uint(mx));
    }
  }
}

template <int NS>
static int launch_mscam_columns(const MscamParams& p, cudaStream_t st) {
  const int bw = (p.W % 224 == 0) ? 224 : ((p.W >= 256) ? 256 : 128);
  size_t floats = 0;
  for (int s = 0; s < NS; ++s) {
    const int nr = min(p.gh[s], cdiv(MSCAM_TILE_ROWS * p.gh[s], p.H) + 3);
    const int nc = min(p.gw[s], cdiv(bw * p.gw[s], p.W) + 3);
    floats += 2ull * nr * nc;
  }
  const size_t smem = (floats + 4) * sizeof(float) +
                      static_cast<size_t>(MSCAM_TILE_ROWS) * (NS * (sizeof(MscamRow) + sizeof(float2)) + sizeof(int));
  if (smem > 48 * 1024) return -1;  // unusual geometry: let the generic kernel handle it
  dim3 grid(p.b * p.K, cdiv(p.H, MSCAM_TILE_ROWS), cdiv(p.W, bw));
  mscam_kernel<0, NS><<<grid, bw, smem, st>>>(p);
  DUPL_LAUNCH_OK();
  mscam_kernel<1, NS><<<grid, bw, smem, st>>>(p);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

// ---------------------------------------------------------------------------------------------
// Single-launch variant: a CLUSTER of CL CTAs owns one (image, class) plane, CTA `rank` its rows
// [rank*rpc, (rank+1)*rpc).  Each CTA computes its un-normalised values ONCE into shared memory (a thread owns one
// output column, as in mscam_kernel), the plane's min / max go through distributed shared memory (two cluster
// barriers, no atomics, no second launch that recomputes every value), then the tile is normalised out of shared
// memory and streamed to HBM.  Same arithmetic in the same order as mscam_kernel: bit-identical output.
// CL = 8: 80 planes x 8 CTAs of 110 KB are 2.16 waves of the 296 CTA slots (measured 61 us per call, three waves; the
// two-pass kernels took 85 us).  CL = 16 (non-portable cluster size; 55 KB per CTA, 2.9 waves of 444 slots at half the work
// each) measured SLOWER (81 us: 16-CTA clusters place badly) and stays an experiment switch, DUPL_MSCAM_CLUSTER=16.
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ float ld_cluster_f32(uint32_t cluster_addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t mscam_mapa(const void* p, uint32_t rank) {
  uint32_t r;
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void mscam_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int NS, int CL>
__global__ void __launch_bounds__(512) mscam_cluster_kernel(MscamParams p, int rpc) {
  extern __shared__ float sm[];
  __shared__ int g_rlo[NS], g_nr[NS], g_off[NS + 1];
  __shared__ float red_mn[16], red_mx[16];
  __shared__ float cta_mm[2];  // this CTA's min / max, read by its peers
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int plane = blockIdx.x / CL;  // img*K + k
  const int img = plane / p.K, k = plane % p.K;
  const int y_begin = min(static_cast<int>(rank) * rpc, p.H);
  const int y_end = min(y_begin + rpc, p.H);
  const int nrows = y_end - y_begin;  // 0 for the trailing CTAs of a short plane: they only take part in the barriers
  const int nt = blockDim.x;

  if (threadIdx.x == 0) {
    int off = 0;
    for (int s = 0; s < NS; ++s) {
      const int gh = p.gh[s], gw = p.gw[s];
      int rlo = 0, rhi = -1;
      if (nrows > 0) {
        rlo = lin_coord(y_begin, gh, static_cast<float>(gh) / p.H).i0;
        rhi = lin_coord(y_end - 1, gh, static_cast<float>(gh) / p.H).i1;
      }
      g_rlo[s] = rlo; g_nr[s] = rhi - rlo + 1;
      g_off[s] = off;
      off += 2 * g_nr[s] * gw;
    }
    g_off[NS] = off;
  }
  __syncthreads();
  // full-width window of the low-res rows this CTA samples: image (A) and twin (B, stored pre-flipped along x)
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int gw = p.gw[s], n = p.gh[s] * gw;
    const int nr = g_nr[s], rlo = g_rlo[s];
    const float* a = p.lowres[s] + (static_cast<long>(img) * p.K + k) * n;
    const float* bt = p.lowres[s] + (static_cast<long>(img + p.b) * p.K + k) * n;
    float* A = sm + g_off[s];
    float* B = A + nr * gw;
    for (int i = threadIdx.x; i < nr * gw; i += nt) {
      const int r = i / gw, c = i - r * gw;
      A[i] = __ldg(a + (rlo + r) * gw + c);
      B[i] = __ldg(bt + (rlo + r) * gw + (gw - 1 - c));
    }
  }
  MscamRow* rows = reinterpret_cast<MscamRow*>(sm + ((g_off[NS] + 3) & ~3));
  float2* wts = reinterpret_cast<float2*>(rows + rpc * NS);
  int* flags = reinterpret_cast<int*>(wts + rpc * NS);
  float* vals = reinterpret_cast<float*>(flags + ((rpc + 3) & ~3));  // [rpc][nt]
  for (int i = threadIdx.x; i < nrows * NS; i += nt) {
    const int r = i / NS, s = i - r * NS;
    const Lin ly = lin_coord(y_begin + r, p.gh[s], static_cast<float>(p.gh[s]) / p.H);
    rows[i] = MscamRow{ly.i0 - g_rlo[s], ly.i1 - g_rlo[s], ly.l0, ly.l1};
    wts[i] = make_float2(ly.l0, ly.l1);
  }
  __syncthreads();
  for (int r = threadIdx.x; r < nrows; r += nt) {
    int f = 0;
    for (int s = 0; s < NS; ++s)
      if (r == 0 || rows[r * NS + s].i0 != rows[(r - 1) * NS + s].i0 || rows[r * NS + s].i1 != rows[(r - 1) * NS + s].i1) f |= 1 << s;
    flags[r] = f;
  }
  __syncthreads();

  const int x = threadIdx.x;
  const bool active = x < p.W;
  const int xc = min(x, p.W - 1);
  int c0[NS], c1[NS], nc[NS];
  float lx0[NS], lx1[NS], t0a[NS], t1a[NS], t0b[NS], t1b[NS];
  const float* Ab[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const Lin lx = lin_coord(xc, p.gw[s], static_cast<float>(p.gw[s]) / p.W);
    c0[s] = lx.i0; c1[s] = lx.i1; lx0[s] = lx.l0; lx1[s] = lx.l1;
    nc[s] = p.gw[s];
    Ab[s] = sm + g_off[s];
    t0a[s] = t1a[s] = t0b[s] = t1b[s] = 0.0f;
  }
  float mn = INFINITY, mx = 0.0f;
  float* vcol = vals + threadIdx.x;
#pragma unroll 4
  for (int r = 0; r < nrows; ++r) {
    const int changed = flags[r];      // bit s: the source rows of scale s differ from the previous output row's
    if (changed != 0) {                // block-uniform and rare: every H/gh output rows
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if ((changed >> s) & 1) {
          const MscamRow rw = rows[r * NS + s];
          const float* A0 = Ab[s] + rw.i0 * nc[s];
          const float* A1 = Ab[s] + rw.i1 * nc[s];
          const float* B0 = A0 + g_nr[s] * nc[s];
          const float* B1 = A1 + g_nr[s] * nc[s];
          t0a[s] = __fmaf_rn(lx0[s], A0[c0[s]], __fmul_rn(lx1[s], A0[c1[s]]));
          t1a[s] = __fmaf_rn(lx0[s], A1[c0[s]], __fmul_rn(lx1[s], A1[c1[s]]));
          t0b[s] = __fmaf_rn(lx0[s], B0[c0[s]], __fmul_rn(lx1[s], B0[c1[s]]));
          t1b[s] = __fmaf_rn(lx0[s], B1[c0[s]], __fmul_rn(lx1[s], B1[c1[s]]));
        }
      }
    }
    float acc = 0.0f;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const float2 l = wts[r * NS + s];
      const float va = __fmaf_rn(l.x, t0a[s], __fmul_rn(l.y, t1a[s]));
      const float vb = __fmaf_rn(l.x, t0b[s], __fmul_rn(l.y, t1b[s]));
      acc += fmaxf(fmaxf(va, vb), 0.0f);
    }
    vcol[r * nt] = acc;
    mn = fminf(mn, acc);
    mx = fmaxf(mx, acc);
  }
  if (!active) { mn = INFINITY; mx = 0.0f; }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  }
  if ((threadIdx.x & 31) == 0) {
    red_mn[threadIdx.x >> 5] = mn;
    red_mx[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int wi = 1; wi < (nt >> 5); ++wi) {
      mn = fminf(mn, red_mn[wi]);
      mx = fmaxf(mx, red_mx[wi]);
    }
    cta_mm[0] = mn;
    cta_mm[1] = mx;
  }
  mscam_cluster_sync();  // every CTA's min / max is published (release / acquire at cluster scope)
  float lo = INFINITY, hi = 0.0f;
#pragma unroll
  for (uint32_t c = 0; c < CL; ++c) {
    lo = fminf(lo, ld_cluster_f32(mscam_mapa(&cta_mm[0], c)));
    hi = fmaxf(hi, ld_cluster_f32(mscam_mapa(&cta_mm[1], c)));
  }
  if (rank == 0 && threadIdx.x == 0) {  // the plane's extrema, as the two-pass kernels leave them
    p.minmax[2 * plane] = __float_as_uint(lo);
    p.minmax[2 * plane + 1] = __float_as_uint(hi);
  }
  const float shift = -lo;                   // cam + max(-cam)
  const float denom = (hi + shift) + 1e-5f;  // max(cam) + 1e-5
  const float rcp = __frcp_rn(denom);
  if (active) {
    float* o = p.out + (static_cast<long>(plane) * p.H + y_begin) * p.W + x;
#pragma unroll 4
    for (int r = 0; r < nrows; ++r) __stcs(o + static_cast<long>(r) * p.W, div_rn_normal(vcol[r * nt] + shift, denom, rcp));
  }
  mscam_cluster_sync();  // no CTA leaves while a peer may still read its cta_mm
}

template <int NS, int CL>
static int launch_mscam_cluster_cl(const MscamParams& p, cudaStream_t st) {
  const int nt = (p.W + 31) / 32 * 32;
  if (nt > 512) return -1;
  const int rpc = cdiv(p.H, CL);
  size_t floats = 0;
  for (int s = 0; s < NS; ++s) {
    const int nr = min(p.gh[s], cdiv(rpc * p.gh[s], p.H) + 3);
    floats += 2ull * nr * p.gw[s];
  }
  const size_t smem = (floats + 4) * sizeof(float) + static_cast<size_t>(rpc) * NS * (sizeof(MscamRow) + sizeof(float2)) +
                      static_cast<size_t>((rpc + 3) & ~3) * sizeof(int) + static_cast<size_t>(rpc) * nt * sizeof(float);
  if (smem > 200 * 1024) return -1;  // tall planes: the two-pass kernels
  static size_t smem_set = 0;
  static int usable = -1;            // CL > 8: can the device co-schedule such a cluster at all?
  if (smem > smem_set) {
    DUPL_CUDA_OK(cudaFuncSetAttribute(mscam_cluster_kernel<NS, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    if (CL > 8) DUPL_CUDA_OK(cudaFuncSetAttribute(mscam_cluster_kernel<NS, CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    smem_set = smem;
    usable = -1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(p.b * p.K * CL));
  cfg.blockDim = dim3(nt);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (CL > 8 && usable < 0) {
    int n = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&n, mscam_cluster_kernel<NS, CL>, &cfg);
    if (e != cudaSuccess) (void)cudaGetLastError();
    usable = (e == cudaSuccess && n > 0) ? 1 : 0;
  }
  if (CL > 8 && usable == 0) return -1;
  DUPL_CUDA_OK(cudaLaunchKernelEx(&cfg, mscam_cluster_kernel<NS, CL>, p, rpc));
  count_launch();
  return DUPL_OK;
}

template <int NS>
static int launch_mscam_cluster(const MscamParams& p, cudaStream_t st) {
  static const int want = getenv("DUPL_MSCAM_CLUSTER") ? atoi(getenv("DUPL_MSCAM_CLUSTER")) : 8;
  if (want >= 16 && p.H >= 64) {
    const int rc = launch_mscam_cluster_cl<NS, 16>(p, st);
    if (rc >= 0) return rc;
  }
  return launch_mscam_cluster_cl<NS, 8>(p, st);
}

__global__ void mscam_init_minmax(unsigned int* mm, int planes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < planes) {
    mm[2 * i] = 0x7f800000u;  // +inf
    mm[2 * i + 1] = 0u;       // 0.0
  }
}

// ---------------------------------------------------------------------------------------------
// cam_to_label / cam_to_label_dynamic_cls (utils/cam_helper.py:8-55)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void py_slice(int a, int b, int n, int& lo, int& hi) {  // Python a:b on length n
  lo = a < 0 ? max(a + n, 0) : min(a, n);
  hi = b < 0 ? max(b + n, 0) : min(b, n);
}

__global__ void __launch_bounds__(256) cam_to_label_kernel(dupl_cam_to_label_args a) {
  const long hw = static_cast<long>(a.h) * a.w;
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= a.b * hw) return;
  const int img = static_cast<int>(idx / hw);
  const long pix = idx % hw;
  const int y = static_cast<int>(pix / a.w), x = static_cast<int>(pix % a.w);
  const float* c = a.cam + static_cast<long>(img) * a.K * hw + pix;
  const float* cl = a.cls_label + static_cast<long>(img) * a.K;
  float best = 0.0f;
  int arg = 0;
  for (int k = 0; k < a.K; ++k) {
    const float v = __fmul_rn(__ldg(cl + k), __ldg(c + k * hw));
    if (a.valid_cam != nullptr) a.valid_cam[static_cast<long>(img) * a.K * hw + k * hw + pix] = v;
    if (k == 0 || v > best) {  // strict > keeps the first index on ties, like torch.max
      best = v;
      arg = k;
    }
  }
  long lab = arg + 1;
  if (best <= a.bkg_thre) lab = 0;
  if (a.img_box != nullptr) {
    if (a.ignore_mid) {
      const float ht = a.high_thre != nullptr ? __ldg(a.high_thre + img) : a.high_thre_scalar;
      if (best <= ht) lab = a.ignore_index;
      if (best <= a.low_thre) lab = 0;
    }
    const int* bx = a.img_box + 4 * img;
    int y0, y1, x0, x1;
    py_slice(bx[0], bx[1], a.h, y0, y1);
    py_slice(bx[2], bx[3], a.w, x0, x1);
    if (!(y >= y0 && y < y1 && x >= x0 && x < x1)) lab = a.ignore_index;
  }
  a.label[idx] = lab;
}

// ---------------------------------------------------------------------------------------------
// label_to_aff_mask (utils/cam_helper.py:323-335)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) aff_mask_kernel(const long long* __restrict__ label, long long* __restrict__ aff,
                                                       int b, int n, long long ignore) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  const int img = blockIdx.z;
  if (j >= n) return;
  const long long li = __ldg(label + static_cast<long>(img) * n + i);
  const long long lj = __ldg(label + static_cast<long>(img) * n + j);
  long long v = (li == lj) ? 1 : 0;
  if (li == ignore || lj == ignore || i == j) v = ignore;
  aff[(static_cast<long>(img) * n + i) * n + j] = v;
}

}  // namespace dupl

using namespace dupl;

extern "C" int dupl_mscam_post(const dupl_mscam_args* a, void* stream) {
  DUPL_CHECK_ARG(a != nullptr, "dupl_mscam_post: args is NULL");
  DUPL_CHECK_ARG(a->nscale >= 1 && a->nscale <= DUPL_MAX_SEGMENTS, "dupl_mscam_post: nscale=%d", a->nscale);
  DUPL_CHECK_ARG(a->b > 0 && a->K > 0 && a->H > 0 && a->W > 0 && a->out && a->minmax, "dupl_mscam_post: bad arguments");
  MscamParams p;
  p.nscale = a->nscale; p.b = a->b; p.K = a->K; p.H = a->H; p.W = a->W;
  p.out = a->out;
  p.minmax = reinterpret_cast<unsigned int*>(a->minmax);
  size_t smem = 0;
  for (int s = 0; s < a->nscale; ++s) {
    DUPL_CHECK_ARG(a->lowres[s] && a->gh[s] > 0 && a->gw[s] > 0, "dupl_mscam_post: scale %d is empty", s);
    p.lowres[s] = a->lowres[s]; p.gh[s] = a->gh[s]; p.gw[s] = a->gw[s];
    smem += 2ull * a->gh[s] * a->gw[s] * sizeof(float);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int planes_all = a->b * a->K;
  if (a->nscale <= 4 && planes_all <= 65535 * 32 && getenv("DUPL_MSCAM_GENERIC") == nullptr) {
    int rc = -1;
    if (getenv("DUPL_MSCAM_2PASS") == nullptr) {  // one launch: a cluster of 8 CTAs per plane, values computed once
      switch (a->nscale) {
        case 1: rc = launch_mscam_cluster<1>(p, st); break;
        case 2: rc = launch_mscam_cluster<2>(p, st); break;
        case 3: rc = launch_mscam_cluster<3>(p, st); break;
        case 4: rc = launch_mscam_cluster<4>(p, st); break;
      }
      if (rc >= 0) return rc;
    }
    mscam_init_minmax<<<cdiv(planes_all, 256), 256, 0, st>>>(p.minmax, planes_all);
    DUPL_LAUNCH_OK();
    switch (a->nscale) {
      case 1: rc = launch_mscam_columns<1>(p, st); break;
      case 2: rc = launch_mscam_columns<2>(p, st); break;
      case 3: rc = launch_mscam_columns<3>(p, st); break;
      case 4: rc = launch_mscam_columns<4>(p, st); break;
    }
    if (rc >= 0) return rc;
  }
  DUPL_CHECK_ARG(smem <= 200 * 1024, "dupl_mscam_post: low-res maps need %zu B of shared memory", smem);
  static size_t smem_set0 = 0, smem_set1 = 0;
  if (smem > 48 * 1024) {
    if (smem > smem_set0) {
      DUPL_CUDA_OK(cudaFuncSetAttribute(mscam_generic_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      smem_set0 = smem;
    }
    if (smem > smem_set1) {
      DUPL_CUDA_OK(cudaFuncSetAttribute(mscam_generic_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      smem_set1 = smem;
    }
  }
  const int planes = a->b * a->K;
  mscam_init_minmax<<<cdiv(planes, 256), 256, 0, st>>>(p.minmax, planes);
  DUPL_LAUNCH_OK();
  dim3 grid(planes, cdiv(a->H, MSCAM_ROWS));
  mscam_generic_kernel<0><<<grid, MSCAM_THREADS, smem, st>>>(p);
  DUPL_LAUNCH_OK();
  mscam_generic_kernel<1><<<grid, MSCAM_THREADS, smem, st>>>(p);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_cam_to_label(const dupl_cam_to_label_args* a, void* stream) {
  DUPL_CHECK_ARG(a != nullptr, "dupl_cam_to_label: args is NULL");
  DUPL_CHECK_ARG(a->cam && a->cls_label && a->label, "dupl_cam_to_label: NULL pointer");
  DUPL_CHECK_ARG(a->b > 0 && a->K > 0 && a->h > 0 && a->w > 0, "dupl_cam_to_label: bad shape");
  const long total = static_cast<long>(a->b) * a->h * a->w;
  cam_to_label_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_label_to_aff_mask(const int64_t* label, int64_t* aff, int32_t b, int32_t n, int64_t ignore_index,
                                      void* stream) {
  DUPL_CHECK_ARG(label && aff && b > 0 && n > 0, "dupl_label_to_aff_mask: bad arguments");
  DUPL_CHECK_ARG(n <= 65535 && b <= 65535, "dupl_label_to_aff_mask: n=%d b=%d too large for the launch grid", n, b);
  dim3 grid(cdiv(n, 256), n, b);
  aff_mask_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(label), reinterpret_cast<long long*>(aff), b, n, ignore_index);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}
