// Backward-pass kernels of the training step: everything around the tcgen05 GEMMs that
// compute dgrad / wgrad (dupl_gemm_bf16x3 on transposed operand planes).
//   split_transpose   fp32 rows -> split-bf16 planes and/or their transpose (zero-padded contraction dim)
//   transpose_u16     transpose of one bf16 plane (saved activations / weights -> wgrad / dgrad operands)
//   colsum            bias gradients
//   layernorm_bwd     dx accumulated into the residual-stream gradient, dgamma/dbeta via block partials
//   gelu_bwd, relu_bwd
//   col2im3x3         gradient of the dilated im2col (gather form, deterministic)
//   gmp_classify_bwd  global-max-pool + 1x1 classifier
// (the attention backward lives in attention_bwd.cu)
// All reductions have a fixed order: results are bit-reproducible.
#include "common.cuh"
#include "ptx.cuh"

namespace dupl {

// ------------------------------------------------------------------------------------------------
// split + transpose.  src row r (r < R) lives at src + row_map(r) * ld, row_map(r) = r when tokens == 0,
// else (r / np) * tokens + first + r % np  (skips cls rows).
// ------------------------------------------------------------------------------------------------
struct RowMap {
  int tokens, np, first;
  __device__ __forceinline__ long operator()(int r) const {
    return tokens == 0 ? r : static_cast<long>(r / np) * tokens + first + r % np;
  }
};

// 64-row x 32-column tile per block (8 warps x 8 rows).  Optionally also leaves the tile's column sums in
// colsum_ws[tile_row][c] (bias gradients: summed over tile rows by colsum_finish_kernel, fixed order).
// derivative of GELU(erf) at the saved pre-activation (same expression as gelu_bwd_kernel: bit-identical products)
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// gelu_pre != nullptr: src is the gradient w.r.t. GELU's OUTPUT and every element is first multiplied by GELU'(pre) — the
// separate element-wise pass over the [M, 3072] hidden gradient (read + read + write of 38.6 MB each) disappears.
// colsum_out != nullptr: the column sums are finished in THIS launch — every CTA of a 32-column strip takes a ticket after
// publishing its partial sums, and the CTA that takes the last one adds the strip's partials in row-tile order (the order
// colsum_finish_kernel uses: same bits, whichever CTA comes last) and resets the ticket for the next launch.  98 launches
// of a 4 us kernel per training step disappear.
constexpr int COLSUM_TICKET_SETS = 32, COLSUM_TICKET_STRIPS = 128;
__device__ unsigned int g_colsum_tickets[COLSUM_TICKET_SETS * COLSUM_TICKET_STRIPS];  // zero at load, self-resetting

__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ src, int R, int Cc, int ld, RowMap map,
                                                              __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                              __nv_bfloat16* __restrict__ t_hi, __nv_bfloat16* __restrict__ t_lo,
                                                              int Rpad, float* __restrict__ colsum_ws,
                                                              const float* __restrict__ gelu_pre, float* __restrict__ colsum_out,
                                                              unsigned int* __restrict__ tickets) {
  pdl_sync();
  __shared__ float tile[64][33];
  __shared__ float part[8][32];
  __shared__ bool last_of_strip;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 32;
  const int c = c0 + lane;
  float acc = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = warp * 8 + i, r = r0 + rr;
    float v = 0.0f;
    if (r < R && c < Cc) {
      v = src[map(r) * ld + c];
      if (gelu_pre != nullptr) v *= gelu_grad(gelu_pre[static_cast<long>(r) * Cc + c]);
    }
    tile[rr][lane] = v;
    acc += v;
    if (hi != nullptr && r < R && c < Cc) {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      hi[static_cast<long>(r) * Cc + c] = h;
      lo[static_cast<long>(r) * Cc + c] = l;
    }
  }
  if (colsum_ws != nullptr) part[warp][lane] = acc;
  __syncthreads();
  if (colsum_ws != nullptr && warp == 0 && c < Cc) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][lane];
    colsum_ws[static_cast<long>(blockIdx.x) * Cc + c] = t;
  }
  if (t_hi != nullptr) {
    const int rr = 2 * lane, r = r0 + rr;  // Rpad is even: r < Rpad implies r + 1 < Rpad; rows >= R were loaded as 0
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cc = warp * 4 + i, co = c0 + cc;
      if (co < Cc && r < Rpad) {
        uint32_t h2, l2;
        split2_bf16(tile[rr][cc], tile[rr + 1][cc], h2, l2);
        *reinterpret_cast<uint32_t*>(t_hi + static_cast<long>(co) * Rpad + r) = h2;
        *reinterpret_cast<uint32_t*>(t_lo + static_cast<long>(co) * Rpad + r) = l2;
      }
    }
  }
  if (colsum_out != nullptr) {
    if (warp == 0) {
      __threadfence();  // this CTA's partial row of colsum_ws is visible device-wide before its ticket is
      __syncwarp();
      if (lane == 0) {
        const unsigned int t = atomicAdd(tickets + blockIdx.y, 1u);
        last_of_strip = (t == gridDim.x - 1);
        if (last_of_strip) tickets[blockIdx.y] = 0;  // every CTA of the strip has taken its ticket
      }
    }
    __syncthreads();
    if (last_of_strip && warp == 0 && c < Cc) {
      __threadfence();
      float t = 0.0f;
      for (int k = 0; k < static_cast<int>(gridDim.x); ++k) t += __ldcg(colsum_ws + static_cast<long>(k) * Cc + c);
      colsum_out[c] = t;
    }
  }
}

// Same split without transposed outputs (what the MN-major GEMM operands leave to do), vectorised: a lane owns 4 consecutive
// columns of a 64-row x 128-column tile (float4 loads, 8-byte stores per plane, the 8 rows of a warp in flight together).
// The transposing kernel above moved 2 bytes per lane and store: 25 us per launch on average against ~8 us of HBM time.
// Row -> warp assignment, partial sums and their order are those of split_transpose_kernel: bit-identical planes and sums.
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ src, int R, int Cc, int ld,
                                                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                         float* __restrict__ colsum_ws, const float* __restrict__ gelu_pre,
                                                         float* __restrict__ colsum_out, unsigned int* __restrict__ tickets) {
  pdl_sync();
  __shared__ float part[8][128];
  __shared__ bool last_of_strip;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 64 + warp * 8, c0 = blockIdx.y * 128;
  const int c = c0 + 4 * lane;
  const bool col_ok = c < Cc;  // Cc % 4 == 0: a lane's 4 columns are all inside or all outside
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_ok && r0 + i < R) v[i] = *reinterpret_cast<const float4*>(src + static_cast<long>(r0 + i) * ld + c);
  }
  if (gelu_pre != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (col_ok && r0 + i < R) {
        const float4 g = *reinterpret_cast<const float4*>(gelu_pre + static_cast<long>(r0 + i) * Cc + c);
        v[i].x *= gelu_grad(g.x); v[i].y *= gelu_grad(g.y); v[i].z *= gelu_grad(g.z); v[i].w *= gelu_grad(g.w);
      }
    }
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w;
    if (hi != nullptr && col_ok && r0 + i < R) {
      uint32_t h0, l0, h1, l1;
      split2_bf16(v[i].x, v[i].y, h0, l0);
      split2_bf16(v[i].z, v[i].w, h1, l1);
      *reinterpret_cast<uint2*>(hi + static_cast<long>(r0 + i) * Cc + c) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(lo + static_cast<long>(r0 + i) * Cc + c) = make_uint2(l0, l1);
    }
  }
  if (colsum_ws == nullptr) return;
  *reinterpret_cast<float4*>(&part[warp][4 * lane]) = acc;
  __syncthreads();
  const int t = threadIdx.x, col = c0 + t;
  if (t < 128 && col < Cc) {
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += part[k][t];
    colsum_ws[static_cast<long>(blockIdx.x) * Cc + col] = sum;
  }
  if (colsum_out != nullptr) {
    __threadfence();  // this CTA's partial row of colsum_ws is visible device-wide before its ticket is taken
    __syncthreads();
    if (t == 0) {
      const unsigned int tk = atomicAdd(tickets + blockIdx.y, 1u);
      last_of_strip = (tk == gridDim.x - 1);
      if (last_of_strip) tickets[blockIdx.y] = 0;  // every CTA of the strip has taken its ticket
    }
    __syncthreads();
    if (last_of_strip && t < 128 && col < Cc) {
      __threadfence();
      float sum = 0.0f;
      for (int k = 0; k < static_cast<int>(gridDim.x); ++k) sum += __ldcg(colsum_ws + static_cast<long>(k) * Cc + col);
      colsum_out[col] = sum;
    }
  }
}

__global__ void __launch_bounds__(256) colsum_finish_kernel(const float* __restrict__ ws, int ntiles, int Cc,
                                                            float* __restrict__ out) {
  pdl_sync();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= Cc) return;
  float t = 0.0f;
  for (int k = 0; k < ntiles; ++k) t += ws[static_cast<long>(k) * Cc + c];
  out[c] = t;
}

// transposes one or two bf16 planes (blockIdx.z); 64 x 64 tiles, 32-bit global accesses on both sides
__global__ void __launch_bounds__(256) transpose_u16_kernel(const unsigned short* __restrict__ in0,
                                                            const unsigned short* __restrict__ in1, int R, int Cc, int ld,
                                                            RowMap map, unsigned short* __restrict__ out0,
                                                            unsigned short* __restrict__ out1, int Rpad) {
  __shared__ unsigned short tile_t[64][66];  // [column][row]
  const unsigned short* __restrict__ in = blockIdx.z ? in1 : in0;
  unsigned short* __restrict__ out = blockIdx.z ? out1 : out0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = warp * 8 + i, r = r0 + rr, c = c0 + 2 * lane;
    uint32_t v = 0;
    if (r < R && c < Cc) v = *reinterpret_cast<const uint32_t*>(in + map(r) * ld + c);
    tile_t[2 * lane][rr] = static_cast<unsigned short>(v & 0xffffu);
    tile_t[2 * lane + 1][rr] = static_cast<unsigned short>(v >> 16);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int cc = warp * 8 + i, c = c0 + cc, r = r0 + 2 * lane;
    if (c < Cc && r < Rpad)
      *reinterpret_cast<uint32_t*>(out + static_cast<long>(c) * Rpad + r) = *reinterpret_cast<const uint32_t*>(&tile_t[cc][2 * lane]);
  }
}

// Several dense plane pairs transposed by ONE launch (blockIdx.z = 2 * item + plane): the per-layer activations and weights the
// wgrad / dgrad GEMMs of one encoder block consume.  ~200 single launches per step cost more in launch gaps inside the
// captured graph than in copy time.
struct TransposeItems {
  const unsigned short* in[2 * DUPL_MAX_TRANSPOSE_ITEMS];
  unsigned short* out[2 * DUPL_MAX_TRANSPOSE_ITEMS];
  int R[DUPL_MAX_TRANSPOSE_ITEMS], Cc[DUPL_MAX_TRANSPOSE_ITEMS], ld[DUPL_MAX_TRANSPOSE_ITEMS], Rpad[DUPL_MAX_TRANSPOSE_ITEMS];
};

__global__ void __launch_bounds__(256) transpose_multi_kernel(const TransposeItems it) {
  __shared__ unsigned short tile_t[64][66];  // [column][row]
  const int item = blockIdx.z >> 1;
  const int R = it.R[item], Cc = it.Cc[item], ld = it.ld[item], Rpad = it.Rpad[item];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  if (r0 >= Rpad || c0 >= Cc) return;
  const unsigned short* __restrict__ in = it.in[blockIdx.z];
  unsigned short* __restrict__ out = it.out[blockIdx.z];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = warp * 8 + i, r = r0 + rr, c = c0 + 2 * lane;
    uint32_t v = 0;
    if (r < R && c < Cc) v = *reinterpret_cast<const uint32_t*>(in + static_cast<long>(r) * ld + c);
    tile_t[2 * lane][rr] = static_cast<unsigned short>(v & 0xffffu);
    tile_t[2 * lane + 1][rr] = static_cast<unsigned short>(v >> 16);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int cc = warp * 8 + i, c = c0 + cc, r = r0 + 2 * lane;
    if (c < Cc && r < Rpad)
      *reinterpret_cast<uint32_t*>(out + static_cast<long>(c) * Rpad + r) = *reinterpret_cast<const uint32_t*>(&tile_t[cc][2 * lane]);
  }
}

// out[c] = sum_r x[map(r)][c].  Block = 32 columns x 8 row lanes; fixed summation order.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int R, int Cc, int ld, RowMap map,
                                                     float* __restrict__ out) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.0f;
  if (c < Cc)
    for (int r = threadIdx.y; r < R; r += 8) s += x[map(r) * ld + c];
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < Cc) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
    out[c] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward: y = (x - mean) * rstd * gamma + beta.
// dres[row] += rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
// partial[blk][0][c] = sum_rows dy * xhat, partial[blk][1][c] = sum_rows dy   (rows of this block)
// ------------------------------------------------------------------------------------------------
// LNB_ROWS rows per block: 3140 training rows -> 197 blocks with 16 (32 rows gave 99 blocks on 148 SMs).  8 rows (393 blocks,
// one row per warp) measured no better in the captured step (45.22 vs 45.05 ms): 16 stays, DUPL_LNB_ROWS=8 selects the other
// instantiation (the scratch holds either).
template <int V, int LNB_ROWS>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ gamma, float* __restrict__ dres,
                                                            float* __restrict__ partial, int rows, float eps) {
  constexpr int COLS = V * 128;
  pdl_sync();
  __shared__ float sbuf[8][COLS];  // cross-warp reduction buffer, used twice (dgamma then dbeta partials)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 ag[V], ab[V];
#pragma unroll
  for (int i = 0; i < V; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int rr = warp; rr < LNB_ROWS; rr += 8) {
    const int row = blockIdx.x * LNB_ROWS + rr;
    if (row >= rows) break;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long>(row) * COLS);
    const float4* dr = reinterpret_cast<const float4*>(dy + static_cast<long>(row) * COLS);
    float4 v[V], d[V];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i] = xr[lane + 32 * i];
      d[i] = dr[lane + 32 * i];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / COLS);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q * (1.0f / COLS) + eps);
    float s1 = 0.0f, s2 = 0.0f;  // sum g, sum g * xhat
    float4 g[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
      g[i] = make_float4(d[i].x * gm.x, d[i].y * gm.y, d[i].z * gm.z, d[i].w * gm.w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
      ag[i].x += d[i].x * v[i].x; ag[i].y += d[i].y * v[i].y; ag[i].z += d[i].z * v[i].z; ag[i].w += d[i].w * v[i].w;
      ab[i].x += d[i].x; ab[i].y += d[i].y; ab[i].z += d[i].z; ab[i].w += d[i].w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 * (1.0f / COLS), m2 = s2 * (1.0f / COLS);
    float4* out = reinterpret_cast<float4*>(dres + static_cast<long>(row) * COLS);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 o4 = out[lane + 32 * i];
      o4.x += rstd * (g[i].x - m1 - v[i].x * m2);
      o4.y += rstd * (g[i].y - m1 - v[i].y * m2);
      o4.z += rstd * (g[i].z - m1 - v[i].z * m2);
      o4.w += rstd * (g[i].w - m1 - v[i].w * m2);
      out[lane + 32 * i] = o4;
    }
  }
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = (lane + 32 * i) * 4;
      const float4 t = which == 0 ? ag[i] : ab[i];
      sbuf[warp][c] = t.x; sbuf[warp][c + 1] = t.y; sbuf[warp][c + 2] = t.z; sbuf[warp][c + 3] = t.w;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < COLS; c += 256) {
      float a = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w) a += sbuf[w][c];
      partial[(static_cast<long>(blockIdx.x) * 2 + which) * COLS + c] = a;
    }
  }
}

// 32 columns x 8 row groups per block: each thread sums every 8th block partial of its column, the 8 group sums are then
// added in a fixed order (deterministic).  The one-thread-per-column version walked ~200 partials serially (13 us).
__global__ void __launch_bounds__(256) layernorm_bwd_finish_kernel(const float* __restrict__ partial, int nblocks, int cols,
                                                                   float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_sync();
  __shared__ float sa[8][32], sb[8][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float a = 0.0f, b = 0.0f;
  if (c < cols) {
    for (int k = grp; k < nblocks; k += 8) {
      a += partial[(static_cast<long>(k) * 2 + 0) * cols + c];
      b += partial[(static_cast<long>(k) * 2 + 1) * cols + c];
    }
  }
  sa[grp][lane] = a;
  sb[grp][lane] = b;
  __syncthreads();
  if (grp == 0 && c < cols) {
    float ta = 0.0f, tb = 0.0f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      ta += sa[g][lane];
      tb += sb[g][lane];
    }
    dgamma[c] = ta;
    dbeta[c] = tb;
  }
}

// ------------------------------------------------------------------------------------------------
// elementwise backward of the fused activations
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gelu_bwd_kernel(float* __restrict__ d, const float* __restrict__ pre, long n) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  d[i] *= gelu_grad(pre[i]);
}

// d *= (act_hi > 0 || act_lo > 0): the saved split planes of relu(y)
__global__ void __launch_bounds__(256) relu_bwd_kernel(float* __restrict__ d, const __nv_bfloat16* __restrict__ act_hi,
                                                       const __nv_bfloat16* __restrict__ act_lo, long n) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float a = __bfloat162float(act_hi[i]) + __bfloat162float(act_lo[i]);
  if (!(a > 0.0f)) d[i] = 0.0f;
}

// ------------------------------------------------------------------------------------------------
// col2im of the dilated 3x3 im2col: din[row(b,y,x)][c] += sum_tap dcol[(b, y - dy, x - dx)][tap*Cin + c]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) col2im3x3_kernel(const float* __restrict__ dcol, float* __restrict__ din, int B, int gh,
                                                        int gw, int Cin, int dil, int ld_in, RowMap map, int accumulate) {
  const long total = static_cast<long>(B) * gh * gw * Cin;
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % Cin);
  long t = idx / Cin;
  const int x = static_cast<int>(t % gw);
  t /= gw;
  const int y = static_cast<int>(t % gh);
  const int b = static_cast<int>(t / gh);
  float s = 0.0f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    // output position (yo, xo) read input (yo + dy, xo + dx): this input pixel feeds yo = y - dy
    const int yo = y - (tap / 3 - 1) * dil, xo = x - (tap % 3 - 1) * dil;
    if (yo >= 0 && yo < gh && xo >= 0 && xo < gw)
      s += dcol[(static_cast<long>(b) * gh * gw + yo * gw + xo) * (9L * Cin) + tap * Cin + c];
  }
  float* o = din + map(static_cast<int>(static_cast<long>(b) * gh * gw + y * gw + x)) * ld_in + c;
  *o = accumulate ? *o + s : s;
}

// dst[map(b*np + p)][c] += src[b][c][p]
__global__ void __launch_bounds__(256) nchw_to_rows_add_kernel(const float* __restrict__ src, float* __restrict__ dst, int np,
                                                               int Cc, int ld, RowMap map) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < Cc && p < np) ? src[(static_cast<long>(b) * Cc + c) * np + p] : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < np && c < Cc) dst[map(b * np + p) * ld + c] += tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------------------------------------
// global max pool + classifier backward.  One block per image.
// dx[row(b, argmax[b][d])][d] += sum_k dlogits[b][k] w[k][d];  dw_partial[b][k][d] = dlogits[b][k] * pooled[b][d]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gmp_classify_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ dlogits, const int* __restrict__ argmax,
                                                               float* __restrict__ dx, float* __restrict__ dw_partial, int np,
                                                               int D, int K, int ld, RowMap map) {
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const int p = argmax[static_cast<long>(b) * D + d];
    const long row = map(b * np + p);
    const float pooled = x[row * ld + d];
    float g = 0.0f;
    for (int k = 0; k < K; ++k) {
      const float dl = dlogits[b * K + k];
      g = fmaf(dl, __ldg(w + static_cast<long>(k) * D + d), g);
      dw_partial[(static_cast<long>(b) * K + k) * D + d] = dl * pooled;
    }
    dx[row * ld + d] += g;  // one (row, d) per thread within the image: no race
  }
}

__global__ void __launch_bounds__(256) sum_over_first_kernel(const float* __restrict__ in, int n_first, long inner,
                                                             float* __restrict__ out) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= inner) return;
  float s = 0.0f;
  for (int b = 0; b < n_first; ++b) s += in[b * inner + i];
  out[i] = s;
}

static RowMap make_map(int tokens, int np, int first) {
  RowMap m;
  m.tokens = tokens; m.np = np > 0 ? np : 1; m.first = first;
  return m;
}

}  // namespace dupl

using namespace dupl;

static int split_transpose_impl(const float* src, int32_t R, int32_t Cc, int32_t ld, int32_t tokens, int32_t np, int32_t first,
                                void* hi, void* lo, void* t_hi, void* t_lo, int32_t Rpad, float* colsum_ws, float* colsum,
                                const float* gelu_pre, void* stream) {
  DUPL_CHECK_ARG(src && R > 0 && Cc > 0 && ld >= Cc, "dupl_split_transpose: bad arguments");
  DUPL_CHECK_ARG((hi == nullptr) == (lo == nullptr) && (t_hi == nullptr) == (t_lo == nullptr) && (hi || t_hi),
                 "dupl_split_transpose: planes must come in hi/lo pairs");
  DUPL_CHECK_ARG(t_hi == nullptr || (Rpad >= R && Rpad % 2 == 0), "dupl_split_transpose: Rpad=%d must be even and >= R=%d", Rpad, R);
  DUPL_CHECK_ARG((colsum_ws == nullptr) == (colsum == nullptr), "dupl_split_transpose: colsum needs its workspace");
  const int rows = t_hi ? Rpad : R;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(cdiv(rows, 64), cdiv(Cc, 32));
  // column sums finished by the last CTA of every 32-column strip (DUPL_COLSUM_2PASS=1: the separate finish kernel)
  static const bool two_pass = getenv("DUPL_COLSUM_2PASS") != nullptr;
  static unsigned int* ticket_base = nullptr;
  static unsigned int next_set = 0;
  const bool fused = colsum != nullptr && !two_pass && grid.y <= static_cast<unsigned>(COLSUM_TICKET_STRIPS);
  unsigned int* tickets = nullptr;
  if (fused) {
    if (ticket_base == nullptr) DUPL_CUDA_OK(cudaGetSymbolAddress(reinterpret_cast<void**>(&ticket_base), g_colsum_tickets));
    tickets = ticket_base + (next_set++ % COLSUM_TICKET_SETS) * COLSUM_TICKET_STRIPS;  // launches in flight never share a set
  }
  // no transposed outputs, identity row map, 16-byte aligned rows: the vectorised kernel (DUPL_SPLIT_ROWS=0: the tile kernel)
  static const bool rows_kernel = !(getenv("DUPL_SPLIT_ROWS") && atoi(getenv("DUPL_SPLIT_ROWS")) == 0);
  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (rows_kernel && t_hi == nullptr && tokens == 0 && Cc % 4 == 0 && ld % 4 == 0 && aligned16(src) && aligned16(hi) && aligned16(lo) &&
      (gelu_pre == nullptr || aligned16(gelu_pre)) && cdiv(Cc, 128) <= COLSUM_TICKET_STRIPS) {
    dim3 g2(cdiv(R, 64), cdiv(Cc, 128));
    DUPL_CUDA_OK(launch_pdl(split_rows_kernel, g2, dim3(256), 0, st, src, R, Cc, ld, static_cast<__nv_bfloat16*>(hi),
                            static_cast<__nv_bfloat16*>(lo), colsum_ws, gelu_pre, fused ? colsum : nullptr, tickets));
    count_launch();
    if (colsum != nullptr && !fused) {
      DUPL_CUDA_OK(launch_pdl(colsum_finish_kernel, dim3(cdiv(Cc, 256)), dim3(256), 0, st, static_cast<const float*>(colsum_ws),
                              static_cast<int>(g2.x), Cc, colsum));
      count_launch();
    }
    return DUPL_OK;
  }
  DUPL_CUDA_OK(launch_pdl(split_transpose_kernel, grid, dim3(256), 0, st, src, R, Cc, ld, make_map(tokens, np, first),
                          static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), static_cast<__nv_bfloat16*>(t_hi),
                          static_cast<__nv_bfloat16*>(t_lo), Rpad, colsum_ws, gelu_pre, fused ? colsum : nullptr, tickets));
  count_launch();
  if (colsum != nullptr && !fused) {
    DUPL_CUDA_OK(launch_pdl(colsum_finish_kernel, dim3(cdiv(Cc, 256)), dim3(256), 0, st, static_cast<const float*>(colsum_ws),
                            static_cast<int>(grid.x), Cc, colsum));
    count_launch();
  }
  return DUPL_OK;
}

extern "C" int dupl_split_transpose(const float* src, int32_t R, int32_t Cc, int32_t ld, int32_t tokens, int32_t np,
                                    int32_t first, void* hi, void* lo, void* t_hi, void* t_lo, int32_t Rpad,
                                    float* colsum_ws, float* colsum, void* stream) {
  return split_transpose_impl(src, R, Cc, ld, tokens, np, first, hi, lo, t_hi, t_lo, Rpad, colsum_ws, colsum, nullptr, stream);
}

extern "C" int dupl_split_transpose_gelu(const float* src, const float* gelu_pre, int32_t R, int32_t Cc, void* hi, void* lo, void* t_hi,
                                         void* t_lo, int32_t Rpad, float* colsum_ws, float* colsum, void* stream) {
  DUPL_CHECK_ARG(gelu_pre != nullptr, "dupl_split_transpose_gelu: pre-activation is NULL");
  return split_transpose_impl(src, R, Cc, Cc, 0, 0, 0, hi, lo, t_hi, t_lo, Rpad, colsum_ws, colsum, gelu_pre, stream);
}

extern "C" int dupl_transpose_planes(const void* in_hi, const void* in_lo, int32_t R, int32_t Cc, int32_t ld, int32_t tokens,
                                     int32_t np, int32_t first, void* out_hi, void* out_lo, int32_t Rpad, void* stream) {
  DUPL_CHECK_ARG(in_hi && out_hi && (in_lo == nullptr) == (out_lo == nullptr) && R > 0 && Cc > 0 && ld >= Cc && Rpad >= R,
                 "dupl_transpose_planes: bad arguments");
  DUPL_CHECK_ARG(Cc % 2 == 0 && ld % 2 == 0 && Rpad % 2 == 0, "dupl_transpose_planes: Cc=%d, ld=%d, Rpad=%d must be even", Cc, ld, Rpad);
  dim3 grid(cdiv(Rpad, 64), cdiv(Cc, 64), in_lo ? 2 : 1);
  transpose_u16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const unsigned short*>(in_hi), static_cast<const unsigned short*>(in_lo), R, Cc, ld, make_map(tokens, np, first),
      static_cast<unsigned short*>(out_hi), static_cast<unsigned short*>(out_lo), Rpad);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_transpose_planes_multi(const dupl_transpose_item* items, int32_t n_items, void* stream) {
  DUPL_CHECK_ARG(items != nullptr && n_items >= 1 && n_items <= DUPL_MAX_TRANSPOSE_ITEMS, "dupl_transpose_planes_multi: 1..%d items",
                 DUPL_MAX_TRANSPOSE_ITEMS);
  TransposeItems T;
  int max_r = 0, max_c = 0;
  for (int i = 0; i < n_items; ++i) {
    const dupl_transpose_item& e = items[i];
    DUPL_CHECK_ARG(e.in_hi && e.in_lo && e.out_hi && e.out_lo && e.R > 0 && e.Cc > 0 && e.ld >= e.Cc && e.Rpad >= e.R,
                   "dupl_transpose_planes_multi: bad item %d", i);
    DUPL_CHECK_ARG(e.Cc % 2 == 0 && e.ld % 2 == 0 && e.Rpad % 2 == 0, "dupl_transpose_planes_multi: item %d: Cc, ld, Rpad must be even", i);
    T.in[2 * i] = static_cast<const unsigned short*>(e.in_hi);
    T.in[2 * i + 1] = static_cast<const unsigned short*>(e.in_lo);
    T.out[2 * i] = static_cast<unsigned short*>(e.out_hi);
    T.out[2 * i + 1] = static_cast<unsigned short*>(e.out_lo);
    T.R[i] = e.R; T.Cc[i] = e.Cc; T.ld[i] = e.ld; T.Rpad[i] = e.Rpad;
    max_r = e.Rpad > max_r ? e.Rpad : max_r;
    max_c = e.Cc > max_c ? e.Cc : max_c;
  }
  dim3 grid(cdiv(max_r, 64), cdiv(max_c, 64), 2 * n_items);
  transpose_multi_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(T);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_colsum(const float* x, int32_t R, int32_t Cc, int32_t ld, int32_t tokens, int32_t np, int32_t first,
                           float* out, void* stream) {
  DUPL_CHECK_ARG(x && out && R > 0 && Cc > 0, "dupl_colsum: bad arguments");
  colsum_kernel<<<cdiv(Cc, 32), dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(x, R, Cc, ld, make_map(tokens, np, first), out);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_layernorm_bwd(const float* dy, const float* x, const float* gamma, float* dres, float* partial,
                                  float* dgamma, float* dbeta, int32_t rows, int32_t cols, float eps, void* stream) {
  DUPL_CHECK_ARG(dy && x && gamma && dres && partial && dgamma && dbeta, "dupl_layernorm_bwd: NULL pointer");
  DUPL_CHECK_ARG(rows > 0 && cols == 768, "dupl_layernorm_bwd: rows=%d cols=%d (768 columns are built)", rows, cols);
  static const int lnb_rows = (getenv("DUPL_LNB_ROWS") && atoi(getenv("DUPL_LNB_ROWS")) == 8) ? 8 : 16;
  const int nb = cdiv(rows, lnb_rows);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (lnb_rows == 16) DUPL_CUDA_OK(launch_pdl(layernorm_bwd_kernel<6, 16>, dim3(nb), dim3(256), 0, st, dy, x, gamma, dres, partial, rows, eps));
  else DUPL_CUDA_OK(launch_pdl(layernorm_bwd_kernel<6, 8>, dim3(nb), dim3(256), 0, st, dy, x, gamma, dres, partial, rows, eps));
  count_launch();
  DUPL_CUDA_OK(launch_pdl(layernorm_bwd_finish_kernel, dim3(cdiv(cols, 32)), dim3(256), 0, st, static_cast<const float*>(partial), nb,
                          cols, dgamma, dbeta));
  count_launch();
  return DUPL_OK;
}

extern "C" int dupl_gelu_bwd(float* d, const float* pre, int64_t n, void* stream) {
  DUPL_CHECK_ARG(d && pre && n > 0, "dupl_gelu_bwd: bad arguments");
  gelu_bwd_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d, pre, n);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_relu_bwd(float* d, const void* act_hi, const void* act_lo, int64_t n, void* stream) {
  DUPL_CHECK_ARG(d && act_hi && act_lo && n > 0, "dupl_relu_bwd: bad arguments");
  relu_bwd_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d, static_cast<const __nv_bfloat16*>(act_hi), static_cast<const __nv_bfloat16*>(act_lo), n);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_col2im3x3(const float* dcol, float* din, int32_t B, int32_t gh, int32_t gw, int32_t Cin, int32_t dilation,
                              int32_t ld_in, int32_t tokens, int32_t first, int32_t accumulate, void* stream) {
  DUPL_CHECK_ARG(dcol && din && B > 0 && gh > 0 && gw > 0 && Cin > 0, "dupl_col2im3x3: bad arguments");
  const long total = static_cast<long>(B) * gh * gw * Cin;
  col2im3x3_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dcol, din, B, gh, gw, Cin, dilation, ld_in, make_map(tokens, gh * gw, first), accumulate);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_nchw_to_rows_add(const float* src, float* dst, int32_t B, int32_t np, int32_t Cc, int32_t ld,
                                     int32_t tokens, int32_t first, void* stream) {
  DUPL_CHECK_ARG(src && dst && B > 0 && np > 0 && Cc > 0, "dupl_nchw_to_rows_add: bad arguments");
  dim3 grid(cdiv(np, 32), cdiv(Cc, 32), B), block(32, 8);
  nchw_to_rows_add_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, np, Cc, ld, make_map(tokens, np, first));
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_gmp_classify_bwd(const float* x, const float* w, const float* dlogits, const int32_t* argmax, float* dx,
                                     float* dw_partial, float* dw, int32_t B, int32_t np, int32_t D, int32_t K, int32_t ld,
                                     int32_t tokens, int32_t first, void* stream) {
  DUPL_CHECK_ARG(x && w && dlogits && argmax && dx && dw_partial && dw, "dupl_gmp_classify_bwd: NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  gmp_classify_bwd_kernel<<<B, 256, 0, st>>>(x, w, dlogits, argmax, dx, dw_partial, np, D, K, ld, make_map(tokens, np, first));
  DUPL_LAUNCH_OK();
  const long inner = static_cast<long>(K) * D;
  sum_over_first_kernel<<<static_cast<int>((inner + 255) / 256), 256, 0, st>>>(dw_partial, B, inner, dw);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}
