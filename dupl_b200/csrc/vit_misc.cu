// HBM-bound companions of the tcgen05 GEMMs in the ViT encoder: operand splitting, LayerNorm,
// patch extraction (fused with the multi-scale resize + flip), pos-embed resize, cls rows and
// the CAM class contraction.  All are one-pass, 128-bit vectorised, warp-shuffle reduced.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "resample.cuh"

namespace dupl {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void store_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, float a, float b, float c, float d) {
  uint32_t h0, l0, h1, l1;
  split2_bf16(a, b, h0, l0);
  split2_bf16(c, d, h1, l1);
  *reinterpret_cast<uint2*>(hi) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(lo) = make_uint2(l0, l1);
}

// ------------------------------------------------------------------------------ split
__global__ void split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long n) {
  const long n4 = n >> 2;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    store_split4(hi + 4 * i, lo + 4 * i, v.x, v.y, v.z, v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long i = (n4 << 2) + threadIdx.x;
    split_bf16(x[i], hi[i], lo[i]);
  }
}

// ------------------------------------------------------------------------------ LayerNorm -> split
// One warp per row; the row lives in registers (V float4 per lane), two-pass mean / variance.
template <int V>
__global__ void __launch_bounds__(256) layernorm_split_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              __nv_bfloat16* __restrict__ hi,
                                                              __nv_bfloat16* __restrict__ lo, float* __restrict__ out_f32,
                                                              int rows, float eps) {
  constexpr int COLS = V * 128;
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long>(row) * COLS);
  float4 v[V];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / COLS);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / COLS) + eps);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
    const long o = static_cast<long>(row) * COLS + c;
    const float4 y = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                                 (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
    if (hi != nullptr) store_split4(hi + o, lo + o, y.x, y.y, y.z, y.w);
    if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + o) = y;
  }
}

// ------------------------------------------------------------------------------ patchify
// One thread = 8 consecutive kx of one (patch row, c, ky): 16 B to each plane.
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ images, int b, int H, int W,
                                                       dupl_segment sg, int hs, int ws, int flip_twin,
                                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int np = sg.gh * sg.gw;
  const long total = static_cast<long>(sg.batch) * np * 96;
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int chunk = static_cast<int>(idx % 96);
  const long prow = idx / 96;
  const int img = static_cast<int>(prow / np);
  const int pidx = static_cast<int>(prow % np);
  const int py = pidx / sg.gw, px = pidx % sg.gw;
  const int c = chunk / 32, ky = (chunk % 32) / 2, kx0 = (chunk % 2) * 8;
  const bool flipped = flip_twin && img >= b;
  const int src_img = flipped ? img - b : img;
  const float* plane = images + (static_cast<long>(src_img) * 3 + c) * H * W;
  const int y = py * 16 + ky;
  float v[8];
  if (hs == H && ws == W) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int x = px * 16 + kx0 + e;
      v[e] = __ldg(plane + static_cast<long>(y) * W + (flipped ? ws - 1 - x : x));
    }
  } else {
    const float sy = static_cast<float>(H) / hs, sx = static_cast<float>(W) / ws;
    const Lin ly = lin_coord(y, H, sy);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int x = px * 16 + kx0 + e;
      const Lin lx = lin_coord(flipped ? ws - 1 - x : x, W, sx);
      v[e] = bilerp(plane, W, ly, lx);
    }
  }
  const long o = (sg.patch_row_offset + prow) * 768 + chunk * 8;
  __nv_bfloat16 h[8], l[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) split_bf16(v[e], h[e], l[e]);
  *reinterpret_cast<uint4*>(hi + o) =
      make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
  *reinterpret_cast<uint4*>(lo + o) =
      make_uint4(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]), pack_bf16(l[4], l[5]), pack_bf16(l[6], l[7]));
}

// ------------------------------------------------------------------------------ pos-embed bicubic
__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.0f, x1 = t, x2 = 1.0f - t, x3 = 2.0f - t;
  w[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  w[1] = ((A + 2.0f) * x1 - (A + 3.0f)) * x1 * x1 + 1.0f;
  w[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
  w[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}

__global__ void pos_embed_resize_kernel(const float* __restrict__ pos, float* __restrict__ out, int S, int gh, int gw,
                                        int D) {
  const long total = static_cast<long>(1 + gh * gw) * D;
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int d = static_cast<int>(idx % D);
  const int tok = static_cast<int>(idx / D);
  if (tok == 0) {
    out[idx] = pos[d];
    return;
  }
  const int oy = (tok - 1) / gw, ox = (tok - 1) % gw;
  if (gh == S && gw == S) {
    out[idx] = pos[static_cast<long>(tok) * D + d];
    return;
  }
  const float sy = static_cast<float>(S) / gh, sx = static_cast<float>(S) / gw;
  const float fy = (oy + 0.5f) * sy - 0.5f, fx = (ox + 0.5f) * sx - 0.5f;
  const float iyf = floorf(fy), ixf = floorf(fx);
  float wy[4], wx[4];
  cubic_coeffs(fy - iyf, wy);
  cubic_coeffs(fx - ixf, wx);
  const int iy = static_cast<int>(iyf), ix = static_cast<int>(ixf);
  float acc = 0.0f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int yy = min(max(iy - 1 + a, 0), S - 1);
    float r = 0.0f;
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int xx = min(max(ix - 1 + bb, 0), S - 1);
      r += wx[bb] * __ldg(pos + static_cast<long>(1 + yy * S + xx) * D + d);
    }
    acc += wy[a] * r;
  }
  out[idx] = acc;
}

// ------------------------------------------------------------------------------ cls rows
struct ClsRowsParams {
  dupl_segment seg[DUPL_MAX_SEGMENTS];
  const float* pos[DUPL_MAX_SEGMENTS];
  int nseg;
};
__global__ void cls_rows_kernel(float* __restrict__ tok, const float* __restrict__ cls, ClsRowsParams p, int D) {
  const int s = blockIdx.y;
  const dupl_segment sg = p.seg[s];
  const long total = static_cast<long>(sg.batch) * D;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int img = static_cast<int>(i / D), d = static_cast<int>(i % D);
    tok[(sg.row_offset + static_cast<long>(img) * sg.tokens) * D + d] = cls[d] + p.pos[s][d];
  }
}

// ------------------------------------------------------------------------------ CAM contraction
struct CamParams {
  dupl_segment seg[DUPL_MAX_SEGMENTS];
  long out_offset[DUPL_MAX_SEGMENTS];
  int nseg, total_patch_rows;
};
// One warp per TOK consecutive patch tokens: optional final LayerNorm in registers, then K dot products of length D=768.
// ncu (profiles/r02_ncu_hbm2.md): with one token per warp the kernel sits at 62-68 % of the LSU wavefront peak — the
// K x 768 classifier rows are re-read from L1 for every token (60 KB per 3 KB token row for VOC) — and takes 70-85 us for
// 21 952 tokens, 7x its HBM time.  With TOK tokens in registers a weight row serves all of them.  Measured in the captured
// step: TOK = 1: 77.3 us, TOK = 2 (128 registers, two CTAs per SM): 61.3 us, TOK = 4 (194 registers, one CTA per SM: too few
// warps to cover the load latency): 65.6 us.  Default 2 (DUPL_CAM_TOK overrides).
// Per token the arithmetic and its order do not depend on TOK (bit-identical results).
template <int TOK>
__global__ void __launch_bounds__(256, TOK <= 2 ? 2 : 1) cam_contract_kernel(const float* __restrict__ tok, const float* __restrict__ gamma,
                                                                             const float* __restrict__ beta, float eps,
                                                                             const float* __restrict__ w, int K, CamParams p,
                                                                             float* __restrict__ out) {
  constexpr int V = 6, D = 768;
  const int lane = threadIdx.x & 31;
  const int prow0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * TOK;
  if (prow0 >= p.total_patch_rows) return;
  float4 v[TOK][V];
  float* o[TOK];
  int np_t[TOK];
#pragma unroll
  for (int t = 0; t < TOK; ++t) {
    const int prow = min(prow0 + t, p.total_patch_rows - 1);  // a ragged last group repeats its last token (not stored)
    int si = 0;
    for (int s = 1; s < p.nseg; ++s)
      if (prow >= p.seg[s].patch_row_offset) si = s;
    const dupl_segment sg = p.seg[si];
    const int np = sg.tokens - 1;
    const int local = prow - sg.patch_row_offset;
    const int img = local / np, pidx = local % np;
    const long trow = sg.row_offset + static_cast<long>(img) * sg.tokens + 1 + pidx;
    const float4* xr = reinterpret_cast<const float4*>(tok + trow * D);
#pragma unroll
    for (int i = 0; i < V; ++i) v[t][i] = xr[lane + 32 * i];
    o[t] = out + p.out_offset[si] + static_cast<long>(img) * K * np + pidx;
    np_t[t] = np;
  }
  if (gamma != nullptr) {
#pragma unroll
    for (int t = 0; t < TOK; ++t) {
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < V; ++i) s += (v[t][i].x + v[t][i].y) + (v[t][i].z + v[t][i].w);
      const float mean = warp_sum(s) * (1.0f / D);
      float q = 0.0f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float a = v[t][i].x - mean, b = v[t][i].y - mean, c = v[t][i].z - mean, d = v[t][i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c = (lane + 32 * i) * 4;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
        v[t][i].x = (v[t][i].x - mean) * rstd * g.x + b.x;
        v[t][i].y = (v[t][i].y - mean) * rstd * g.y + b.y;
        v[t][i].z = (v[t][i].z - mean) * rstd * g.z + b.z;
        v[t][i].w = (v[t][i].w - mean) * rstd * g.w + b.w;
      }
    }
  }
  const int nvalid = min(TOK, p.total_patch_rows - prow0);
  for (int k = 0; k < K; ++k) {
    const float4* wr = reinterpret_cast<const float4*>(w + static_cast<long>(k) * D);
    float acc[TOK];
#pragma unroll
    for (int t = 0; t < TOK; ++t) acc[t] = 0.0f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 ww = __ldg(wr + lane + 32 * i);
#pragma unroll
      for (int t = 0; t < TOK; ++t) {
        acc[t] = fmaf(v[t][i].x, ww.x, acc[t]);
        acc[t] = fmaf(v[t][i].y, ww.y, acc[t]);
        acc[t] = fmaf(v[t][i].z, ww.z, acc[t]);
        acc[t] = fmaf(v[t][i].w, ww.w, acc[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < TOK; ++t) {
      const float r = warp_sum(acc[t]);
      if (lane == 0 && t < nvalid) o[t][static_cast<long>(k) * np_t[t]] = r;
    }
  }
}

}  // namespace dupl

using namespace dupl;

extern "C" int dupl_split_bf16(const float* x, void* hi, void* lo, int64_t n, void* stream) {
  DUPL_CHECK_ARG(x && hi && lo && n >= 0, "dupl_split_bf16: bad arguments");
  if (n == 0) return DUPL_OK;
  const long n4 = n >> 2;
  long blocks = (n4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148L * 16) blocks = 148L * 16;
  split_bf16_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), n);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_layernorm_split(const float* x, const float* gamma, const float* beta, void* out_hi, void* out_lo,
                                    float* out_f32, int32_t rows, int32_t cols, float eps, void* stream) {
  DUPL_CHECK_ARG(x && gamma && beta && ((out_hi && out_lo) || out_f32) && ((out_hi == nullptr) == (out_lo == nullptr)),
                 "dupl_layernorm_split: NULL pointer");
  DUPL_CHECK_ARG(rows > 0 && cols % 128 == 0 && cols >= 128 && cols <= 1024, "dupl_layernorm_split: rows=%d cols=%d",
                 rows, cols);
  const int grid = cdiv(rows, 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(out_hi);
  __nv_bfloat16* lo = static_cast<__nv_bfloat16*>(out_lo);
  switch (cols / 128) {
#define LN_CASE(V) \
  case V: DUPL_CUDA_OK(launch_pdl(layernorm_split_kernel<V>, dim3(grid), dim3(256), 0, st, x, gamma, beta, hi, lo, out_f32, rows, eps)); break;
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
#undef LN_CASE
  }
  count_launch();
  return DUPL_OK;
}

extern "C" int dupl_patchify(const float* images, int32_t b, int32_t H, int32_t W, const dupl_segment* seg,
                             int32_t hs, int32_t ws, int32_t flip_twin, void* out_hi, void* out_lo, void* stream) {
  DUPL_CHECK_ARG(images && seg && out_hi && out_lo, "dupl_patchify: NULL pointer");
  DUPL_CHECK_ARG(b > 0 && H > 0 && W > 0 && seg->gh > 0 && seg->gw > 0, "dupl_patchify: bad shape");
  DUPL_CHECK_ARG(hs / 16 == seg->gh && ws / 16 == seg->gw, "dupl_patchify: resized size %dx%d does not give a %dx%d patch grid",
                 hs, ws, seg->gh, seg->gw);
  DUPL_CHECK_ARG(seg->batch == (flip_twin ? 2 * b : b), "dupl_patchify: segment batch %d != %d", seg->batch,
                 flip_twin ? 2 * b : b);
  const long total = static_cast<long>(seg->batch) * seg->gh * seg->gw * 96;
  patchify_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      images, b, H, W, *seg, hs, ws, flip_twin, static_cast<__nv_bfloat16*>(out_hi), static_cast<__nv_bfloat16*>(out_lo));
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_pos_embed_resize(const float* pos_embed, float* out, int32_t src, int32_t gh, int32_t gw, int32_t D,
                                     void* stream) {
  DUPL_CHECK_ARG(pos_embed && out && src > 0 && gh > 0 && gw > 0 && D > 0, "dupl_pos_embed_resize: bad arguments");
  const long total = static_cast<long>(1 + gh * gw) * D;
  pos_embed_resize_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pos_embed, out, src, gh, gw, D);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_cls_rows(float* tok, const float* cls_token, const float* const* pos, const dupl_segment* seg,
                             int32_t nseg, int32_t D, void* stream) {
  DUPL_CHECK_ARG(tok && cls_token && pos && seg && nseg >= 1 && nseg <= DUPL_MAX_SEGMENTS, "dupl_cls_rows: bad arguments");
  ClsRowsParams p;
  p.nseg = nseg;
  for (int s = 0; s < nseg; ++s) {
    p.seg[s] = seg[s];
    p.pos[s] = pos[s];
  }
  cls_rows_kernel<<<dim3(4, nseg), 256, 0, static_cast<cudaStream_t>(stream)>>>(tok, cls_token, p, D);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_cam_contract(const float* tok, const float* gamma, const float* beta, float eps, const float* w,
                                 int32_t K, int32_t D, const dupl_segment* seg, int32_t nseg, float* out,
                                 const int64_t* out_offset_host, void* stream) {
  DUPL_CHECK_ARG(tok && w && seg && out && out_offset_host, "dupl_cam_contract: NULL pointer");
  DUPL_CHECK_ARG(D == 768, "dupl_cam_contract: D=%d (only 768 is built)", D);
  DUPL_CHECK_ARG(nseg >= 1 && nseg <= DUPL_MAX_SEGMENTS && K > 0, "dupl_cam_contract: nseg=%d K=%d", nseg, K);
  DUPL_CHECK_ARG((gamma == nullptr) == (beta == nullptr), "dupl_cam_contract: gamma/beta must both be set or NULL");
  CamParams p;
  p.nseg = nseg;
  int total = 0;
  for (int s = 0; s < nseg; ++s) {
    p.seg[s] = seg[s];
    p.out_offset[s] = out_offset_host[s];
    DUPL_CHECK_ARG(seg[s].patch_row_offset == total, "dupl_cam_contract: patch rows must be packed in segment order");
    total += seg[s].batch * (seg[s].tokens - 1);
  }
  p.total_patch_rows = total;
  static const int tok_per_warp = getenv("DUPL_CAM_TOK") ? atoi(getenv("DUPL_CAM_TOK")) : 2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (tok_per_warp >= 4) cam_contract_kernel<4><<<cdiv(total, 32), 256, 0, st>>>(tok, gamma, beta, eps, w, K, p, out);
  else if (tok_per_warp == 2) cam_contract_kernel<2><<<cdiv(total, 16), 256, 0, st>>>(tok, gamma, beta, eps, w, K, p, out);
  else cam_contract_kernel<1><<<cdiv(total, 8), 256, 0, st>>>(tok, gamma, beta, eps, w, K, p, out);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}
