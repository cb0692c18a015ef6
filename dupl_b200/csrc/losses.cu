// Losses of model/losses.py with fused forward and backward kernels:
//   get_seg_loss         (losses.py:24-39)  background / foreground balanced cross-entropy
//   get_masked_ptc_loss  (losses.py:6-21)   abs-cosine Gram of the feature map against the affinity mask
// Reductions are two-stage (per-block partials, then one block in a fixed order) so results are
// bit-reproducible.  The Gram is computed in fp32 on the CUDA cores (0.94 GFLOP per image): exact
// rather than fast; it is < 1 % of the step.
#include <cuda_bf16.h>

#include "common.cuh"
#include "resample.cuh"

namespace dupl {

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = 0.0f;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x >> 5); ++w) r += sh[w];
  return r;  // valid in thread 0
}

// ------------------------------------------------------------------------------------------------
// seg loss
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seg_ce_fwd_kernel(const float* __restrict__ pred, const long long* __restrict__ label,
                                                         int C, long hw, long total, long long ignore,
                                                         float* __restrict__ lse_out, float* __restrict__ partials) {
  __shared__ float sh[8];
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  float bg_loss = 0.0f, fg_loss = 0.0f, bg_cnt = 0.0f, fg_cnt = 0.0f;
  if (i < total) {
    const long b = i / hw, p = i % hw;
    const float* x = pred + b * C * hw + p;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(x + c * hw));
    float s = 0.0f;
    for (int c = 0; c < C; ++c) s += expf(__ldg(x + c * hw) - mx);
    const float lse = mx + logf(s);
    lse_out[i] = lse;
    const long long lab = label[i];
    if (lab != ignore && lab >= 0 && lab < C) {
      const float nll = lse - __ldg(x + lab * hw);
      if (lab == 0) {
        bg_loss = nll;
        bg_cnt = 1.0f;
      } else {
        fg_loss = nll;
        fg_cnt = 1.0f;
      }
    }
  }
  float r;
  r = block_sum(bg_loss, sh); if (threadIdx.x == 0) partials[4 * blockIdx.x + 0] = r;
  r = block_sum(fg_loss, sh); if (threadIdx.x == 0) partials[4 * blockIdx.x + 1] = r;
  r = block_sum(bg_cnt, sh);  if (threadIdx.x == 0) partials[4 * blockIdx.x + 2] = r;
  r = block_sum(fg_cnt, sh);  if (threadIdx.x == 0) partials[4 * blockIdx.x + 3] = r;
}

// stats[0..3] = sums; stats[4] = loss
__global__ void __launch_bounds__(256) seg_ce_finish_kernel(const float* __restrict__ partials, int nblocks,
                                                            float* __restrict__ stats) {
  __shared__ double sh[4][256];
  double a[4] = {0, 0, 0, 0};
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x)
    for (int k = 0; k < 4; ++k) a[k] += partials[4 * i + k];
  for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] = a[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[4] = {0, 0, 0, 0};
    for (int i = 0; i < 256; ++i)
      for (int k = 0; k < 4; ++k) t[k] += sh[k][i];
    for (int k = 0; k < 4; ++k) stats[k] = static_cast<float>(t[k]);
    const float bg = stats[0] / (stats[2] + 1e-6f), fg = stats[1] / (stats[3] + 1e-6f);
    stats[4] = (bg + fg) * 0.5f;
  }
}

__global__ void __launch_bounds__(256) seg_ce_bwd_kernel(const float* __restrict__ pred, const long long* __restrict__ label,
                                                         const float* __restrict__ lse, const float* __restrict__ stats,
                                                         const float* __restrict__ grad_out, int C, long hw, long total,
                                                         long long ignore, float* __restrict__ dpred) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const long b = i / hw, p = i % hw;
  const long long lab = label[i];
  float w = 0.0f;
  if (lab != ignore && lab >= 0 && lab < C) w = 0.5f / ((lab == 0 ? stats[2] : stats[3]) + 1e-6f) * grad_out[0];
  const float* x = pred + b * C * hw + p;
  float* d = dpred + b * C * hw + p;
  const float l = lse[i];
  for (int c = 0; c < C; ++c) {
    float g = 0.0f;
    if (w != 0.0f) g = w * (expf(__ldg(x + c * hw) - l) - (c == lab ? 1.0f : 0.0f));
    d[c * hw] = g;
  }
}

// ------------------------------------------------------------------------------------------------
// seg loss on bilinearly up-sampled logits (train_final_voc.py:345-352: F.interpolate(segs, size=label.shape[1:],
// mode='bilinear', align_corners=False) followed by get_seg_loss) without materialising the [b,C,H,W] logits:
// forward samples the low-resolution logits per label pixel (online soft-max), backward is a gather per
// low-resolution cell over the label pixels whose bilinear footprint contains it (deterministic, no atomics).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seg_up_fwd_kernel(const float* __restrict__ pred, const long long* __restrict__ label,
                                                         int C, int h, int w, int H, int W, float sy, float sx,
                                                         long long ignore, float* __restrict__ lse_out,
                                                         float* __restrict__ partials) {
  __shared__ float sh[8];
  const int b = blockIdx.z;
  const int X = blockIdx.x * 32 + (threadIdx.x & 31);
  float bg_loss = 0.0f, fg_loss = 0.0f, bg_cnt = 0.0f, fg_cnt = 0.0f;
  const float* base = pred + static_cast<long>(b) * C * h * w;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const int Y = blockIdx.y * 32 + (threadIdx.x >> 5) + 8 * k;
    if (X >= W || Y >= H) continue;
    const Lin ly = lin_coord(Y, h, sy), lx = lin_coord(X, w, sx);
    const long i = (static_cast<long>(b) * H + Y) * W + X;
    const long long lab = label[i];
    float mx = -INFINITY, sum = 0.0f, picked = 0.0f;
    for (int c = 0; c < C; ++c) {
      const float v = bilerp(base + static_cast<long>(c) * h * w, w, ly, lx);
      if (c == lab) picked = v;
      const float m2 = fmaxf(mx, v);
      sum = sum * expf(mx - m2) + expf(v - m2);
      mx = m2;
    }
    const float lse = mx + logf(sum);
    lse_out[i] = lse;
    if (lab != ignore && lab >= 0 && lab < C) {
      if (lab == 0) {
        bg_loss += lse - picked;
        bg_cnt += 1.0f;
      } else {
        fg_loss += lse - picked;
        fg_cnt += 1.0f;
      }
    }
  }
  const int blk = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  float r;
  r = block_sum(bg_loss, sh); if (threadIdx.x == 0) partials[4 * blk + 0] = r;
  r = block_sum(fg_loss, sh); if (threadIdx.x == 0) partials[4 * blk + 1] = r;
  r = block_sum(bg_cnt, sh);  if (threadIdx.x == 0) partials[4 * blk + 2] = r;
  r = block_sum(fg_cnt, sh);  if (threadIdx.x == 0) partials[4 * blk + 3] = r;
}

constexpr int SEGUP_PPT = 4;       // label pixels per thread and chunk in the backward gather
constexpr int SEGUP_MAXC = 96;
// one block per low-resolution cell (x, y, b)
__global__ void __launch_bounds__(256) seg_up_bwd_kernel(const float* __restrict__ pred, const long long* __restrict__ label,
                                                         const float* __restrict__ lse, const float* __restrict__ stats,
                                                         const float* __restrict__ grad_out, int C, int h, int w, int H, int W,
                                                         float sy, float sx, int fy, int fx, long long ignore,
                                                         float* __restrict__ dpred) {
  __shared__ float nb[SEGUP_MAXC * 9];     // logits of the 3x3 neighbourhood [c][dy][dx] (clamped at the border)
  __shared__ float red[SEGUP_MAXC * 8];
  const int cx = blockIdx.x, cy = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* base = pred + static_cast<long>(b) * C * h * w;
  for (int i = tid; i < C * 9; i += 256) {
    const int c = i / 9, dy = (i % 9) / 3, dx = i % 3;
    const int yy = min(max(cy - 1 + dy, 0), h - 1), xx = min(max(cx - 1 + dx, 0), w - 1);
    nb[i] = __ldg(base + (static_cast<long>(c) * h + yy) * w + xx);
  }
  // candidate label pixels: rows [Y0, Y0 + fy), columns [X0, X0 + fx) (a superset of the footprint; tested below)
  const int Y0 = max(static_cast<int>(floorf((cy - 1 + 0.5f) / sy - 0.5f)) - 1, 0);
  const int X0 = max(static_cast<int>(floorf((cx - 1 + 0.5f) / sx - 0.5f)) - 1, 0);
  const float gw_bg = 0.5f / (stats[2] + 1e-6f) * grad_out[0], gw_fg = 0.5f / (stats[3] + 1e-6f) * grad_out[0];

  for (int i = tid; i < C * 8; i += 256) red[i] = 0.0f;
  __syncthreads();
  for (int q0 = 0; q0 < fy * fx; q0 += 256 * SEGUP_PPT) {
    float wgt[SEGUP_PPT], l_y0[SEGUP_PPT], l_y1[SEGUP_PPT], l_x0[SEGUP_PPT], l_x1[SEGUP_PPT], lsev[SEGUP_PPT];
    int idx[SEGUP_PPT], lab[SEGUP_PPT];  // idx packs the neighbourhood offsets y0 | y1 << 2 | x0 << 4 | x1 << 6
#pragma unroll
    for (int k = 0; k < SEGUP_PPT; ++k) {
      wgt[k] = 0.0f; l_y0[k] = l_y1[k] = l_x0[k] = l_x1[k] = lsev[k] = 0.0f; idx[k] = 0; lab[k] = -1;
      const int q = q0 + tid + 256 * k;
      if (q >= fy * fx) continue;
      const int Y = Y0 + q / fx, X = X0 + q % fx;
      if (Y >= H || X >= W) continue;
      const Lin ly = lin_coord(Y, h, sy), lx = lin_coord(X, w, sx);
      if ((ly.i0 != cy && ly.i1 != cy) || (lx.i0 != cx && lx.i1 != cx)) continue;
      float wy = 0.0f, wx = 0.0f;
      if (ly.i0 == cy) wy += ly.l0;
      if (ly.i1 == cy) wy += ly.l1;
      if (lx.i0 == cx) wx += lx.l0;
      if (lx.i1 == cx) wx += lx.l1;
      const long i = (static_cast<long>(b) * H + Y) * W + X;
      const long long lb = label[i];
      if (lb == ignore || lb < 0 || lb >= C) continue;
      wgt[k] = wy * wx * (lb == 0 ? gw_bg : gw_fg);
      lab[k] = static_cast<int>(lb);
      lsev[k] = lse[i];
      l_y0[k] = ly.l0; l_y1[k] = ly.l1; l_x0[k] = lx.l0; l_x1[k] = lx.l1;
      idx[k] = (ly.i0 - cy + 1) | ((ly.i1 - cy + 1) << 2) | ((lx.i0 - cx + 1) << 4) | ((lx.i1 - cx + 1) << 6);
    }
    for (int c = 0; c < C; ++c) {
      const float* n = nb + c * 9;
      float acc = 0.0f;
#pragma unroll
      for (int k = 0; k < SEGUP_PPT; ++k) {
        if (lab[k] < 0) continue;
        const int y0 = idx[k] & 3, y1 = (idx[k] >> 2) & 3, x0 = (idx[k] >> 4) & 3, x1 = (idx[k] >> 6) & 3;
        const float t0 = __fmaf_rn(l_x0[k], n[y0 * 3 + x0], __fmul_rn(l_x1[k], n[y0 * 3 + x1]));
        const float t1 = __fmaf_rn(l_x0[k], n[y1 * 3 + x0], __fmul_rn(l_x1[k], n[y1 * 3 + x1]));
        const float v = __fmaf_rn(l_y0[k], t0, __fmul_rn(l_y1[k], t1));
        acc += wgt[k] * (expf(v - lsev[k]) - (c == lab[k] ? 1.0f : 0.0f));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) red[c * 8 + warp] += acc;  // slot owned by this warp's lane 0
    }
  }
  __syncthreads();
  if (tid < C) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[tid * 8 + k];
    dpred[((static_cast<long>(b) * C + tid) * h + cy) * w + cx] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// PTC loss.  x: [b][C][n] (n = h*w), mask: int64 [b][n][n]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ptc_invnorm_kernel(const float* __restrict__ x, int C, int n, int total,
                                                          float* __restrict__ inv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = i / n, p = i % n;
  const float* col = x + static_cast<long>(b) * C * n + p;
  float s = 0.0f;
  for (int c = 0; c < C; ++c) {
    const float v = __ldg(col + static_cast<long>(c) * n);
    s = fmaf(v, v, s);
  }
  inv[i] = 1.0f / fmaxf(sqrtf(s), 1e-8f);  // F.normalize(eps=1e-8): x / max(||x||, eps)
}

constexpr int PT = 64;  // Gram tile
constexpr int PK = 16;  // channel chunk

// Gs[b][p][q] = cos(x_p, x_q); partial masked sums of |Gs|.
__global__ void __launch_bounds__(256) ptc_gram_kernel(const float* __restrict__ x, const float* __restrict__ inv,
                                                       const long long* __restrict__ mask, int C, int n,
                                                       float* __restrict__ Gs, float* __restrict__ partials) {
  __shared__ float sa[PK][PT + 1], sb[PK][PT + 1];
  __shared__ float sh[8];
  const int b = blockIdx.z;
  const int p0 = blockIdx.y * PT, q0 = blockIdx.x * PT;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;  // 16 x 16 threads, 4 x 4 outputs each
  const float* xb = x + static_cast<long>(b) * C * n;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int c0 = 0; c0 < C; c0 += PK) {
    for (int e = threadIdx.x; e < PK * PT; e += 256) {
      const int cc = e / PT, pp = e % PT;
      const int c = c0 + cc;
      sa[cc][pp] = (c < C && p0 + pp < n) ? __ldg(xb + static_cast<long>(c) * n + p0 + pp) : 0.0f;
      sb[cc][pp] = (c < C && q0 + pp < n) ? __ldg(xb + static_cast<long>(c) * n + q0 + pp) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < PK; ++cc) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sa[cc][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = sb[cc][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float pos_sum = 0.0f, neg_sum = 0.0f, pos_cnt = 0.0f, neg_cnt = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + ty * 4 + i;
    if (p >= n) continue;
    const float ip = inv[b * n + p];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = q0 + tx * 4 + j;
      if (q >= n) continue;
      const float g = acc[i][j] * ip * inv[b * n + q];
      const long o = (static_cast<long>(b) * n + p) * n + q;
      Gs[o] = g;
      const long long m = mask[o];
      if (m == 1) {
        pos_sum += fabsf(g);
        pos_cnt += 1.0f;
      } else if (m == 0) {
        neg_sum += fabsf(g);
        neg_cnt += 1.0f;
      }
    }
  }
  const int blk = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  float r;
  r = block_sum(pos_sum, sh); if (threadIdx.x == 0) partials[4 * blk + 0] = r;
  r = block_sum(neg_sum, sh); if (threadIdx.x == 0) partials[4 * blk + 1] = r;
  r = block_sum(pos_cnt, sh); if (threadIdx.x == 0) partials[4 * blk + 2] = r;
  r = block_sum(neg_cnt, sh); if (threadIdx.x == 0) partials[4 * blk + 3] = r;
}

__global__ void __launch_bounds__(256) ptc_finish_kernel(const float* __restrict__ partials, int nblocks,
                                                         float* __restrict__ stats) {
  __shared__ double sh[4][256];
  double a[4] = {0, 0, 0, 0};
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x)
    for (int k = 0; k < 4; ++k) a[k] += partials[4 * i + k];
  for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] = a[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[4] = {0, 0, 0, 0};
    for (int i = 0; i < 256; ++i)
      for (int k = 0; k < 4; ++k) t[k] += sh[k][i];
    for (int k = 0; k < 4; ++k) stats[k] = static_cast<float>(t[k]);
    // 0.5 (1 - sum_pos / (n_pos + 1)) + 0.5 sum_neg / (n_neg + 1)
    stats[4] = 0.5f * (1.0f - stats[0] / (stats[2] + 1.0f)) + 0.5f * stats[1] / (stats[3] + 1.0f);
  }
}

// dXh[b][c][p] = sum_q T[p][q] xh[c][q],  T[p][q] = S[p][q] + S[q][p],
// S[p][q] = g * sign(Gs[p][q]) * (mask==1 ? -0.5/(n_pos+1) : mask==0 ? 0.5/(n_neg+1) : 0)
__global__ void __launch_bounds__(256) ptc_bwd_gemm_kernel(const float* __restrict__ x, const float* __restrict__ inv,
                                                           const long long* __restrict__ mask, const float* __restrict__ Gs,
                                                           const float* __restrict__ stats, const float* __restrict__ grad_out,
                                                           int C, int n, float* __restrict__ dxh) {
  __shared__ float st[PK][PT + 1];   // T[q chunk][p tile]
  __shared__ float sx[PK][PT + 1];   // xh[q chunk][c tile]
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * PT, c0 = blockIdx.y * PT;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;  // outputs: c = c0 + ty*4 + i, p = p0 + tx*4 + j
  const float g = grad_out[0];
  const float wpos = -0.5f / (stats[2] + 1.0f) * g, wneg = 0.5f / (stats[3] + 1.0f) * g;
  const float* xb = x + static_cast<long>(b) * C * n;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  auto sfun = [&](int p, int q) -> float {
    const long o = (static_cast<long>(b) * n + p) * n + q;
    const long long m = mask[o];
    if (m != 0 && m != 1) return 0.0f;
    const float gs = Gs[o];
    const float sg = gs > 0.0f ? 1.0f : (gs < 0.0f ? -1.0f : 0.0f);
    return sg * (m == 1 ? wpos : wneg);
  };
  for (int q0 = 0; q0 < n; q0 += PK) {
    for (int e = threadIdx.x; e < PK * PT; e += 256) {
      const int qq = e / PT, pp = e % PT;
      const int q = q0 + qq, p = p0 + pp;
      st[qq][pp] = (q < n && p < n) ? sfun(p, q) + sfun(q, p) : 0.0f;
      const int c = c0 + pp;
      sx[qq][pp] = (q < n && c < C) ? __ldg(xb + static_cast<long>(c) * n + q) * inv[b * n + q] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int qq = 0; qq < PK; ++qq) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sx[qq][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = st[qq][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty * 4 + i;
    if (c >= C) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = p0 + tx * 4 + j;
      if (p < n) dxh[(static_cast<long>(b) * C + c) * n + p] = acc[i][j];
    }
  }
}

// x_hat = x * inv  =>  dx = inv * (dxh - x_hat * <x_hat, dxh>)   (columns whose norm was clamped: dx = inv * dxh)
__global__ void __launch_bounds__(256) ptc_bwd_norm_kernel(const float* __restrict__ x, const float* __restrict__ inv,
                                                           const float* __restrict__ dxh, int C, int n, int total,
                                                           float* __restrict__ dx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = i / n, p = i % n;
  const long base = static_cast<long>(b) * C * n + p;
  const float iv = inv[i];
  float dot = 0.0f;
  for (int c = 0; c < C; ++c) dot = fmaf(__ldg(x + base + static_cast<long>(c) * n) * iv, __ldg(dxh + base + static_cast<long>(c) * n), dot);
  const bool clamped = iv >= 1e8f;  // ||x|| <= eps
  for (int c = 0; c < C; ++c) {
    const long o = base + static_cast<long>(c) * n;
    const float xh = __ldg(x + o) * iv;
    dx[o] = iv * (dxh[o] - (clamped ? 0.0f : xh * dot));
  }
}

// ------------------------------------------------------------------------------------------------
// PTC on the tensor cores: the two contractions of the loss (Gram x_hat x_hat^T per image and dX_hat = T x_hat) are
// [784 x 768] x [768 x 784] / [784 x 784] x [784 x 768] GEMMs -> dupl_gemm_bf16x3 (split-bf16 operands, fp32 accumulate).
// The kernels below only produce its operands and consume its results.
// ------------------------------------------------------------------------------------------------
// x [b][C][n] fp32, inv [b][n]  ->  x_hat = x * inv as split planes in both layouts:
//   rows_* [b*n][C]      (token-major: A and W operand of the Gram)
//   cm_*   [b][C][npad]  (channel-major, zero padded to npad = pad64(n): W operand of dX_hat = T x_hat)
__global__ void __launch_bounds__(256) ptc_normalize_split_kernel(const float* __restrict__ x, const float* __restrict__ inv,
                                                                  int C, int n, int npad, __nv_bfloat16* __restrict__ rows_hi,
                                                                  __nv_bfloat16* __restrict__ rows_lo,
                                                                  __nv_bfloat16* __restrict__ cm_hi,
                                                                  __nv_bfloat16* __restrict__ cm_lo) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int p = p0 + tx;
  const float iv = p < n ? inv[b * n + p] : 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k;
    float v = 0.0f;
    if (c < C && p < n) v = __ldg(x + (static_cast<long>(b) * C + c) * n + p) * iv;
    t[ty + 8 * k][tx] = v;
    if (c < C && p < npad) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const long o = (static_cast<long>(b) * C + c) * npad + p;
      cm_hi[o] = h;
      cm_lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int pp = p0 + ty + 8 * k, c = c0 + tx;
    if (pp < n && c < C) {
      const float v = t[tx][ty + 8 * k];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const long o = (static_cast<long>(b) * n + pp) * C + c;
      rows_hi[o] = h;
      rows_lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

// masked sums of |G| -> 4 partials per block (same layout as ptc_gram_kernel's, finished by ptc_finish_kernel).
// The backward consumes sign(G): cosines closer to zero than the split-bf16 operand rounding (2^-17 per element,
// ~1e-6 on a 768-term sum) are recomputed here in fp32 from x and written back, so that near-orthogonal token pairs get
// the sign an fp32 Gram would give them (about one pair per image otherwise flips and moves that image's gradient by 0.5 %).
constexpr float PTC_RECOMPUTE_BELOW = 1e-5f;

__global__ void __launch_bounds__(256) ptc_mask_reduce_kernel(float* __restrict__ G, const long long* __restrict__ mask,
                                                              const float* __restrict__ x, const float* __restrict__ inv,
                                                              int C, int n, long total, float* __restrict__ partials) {
  __shared__ float sh[8];
  float pos_sum = 0.0f, neg_sum = 0.0f, pos_cnt = 0.0f, neg_cnt = 0.0f;
  const long nn = static_cast<long>(n) * n;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += gridDim.x * 256L) {
    const long long m = mask[i];
    float g = G[i];
    if (fabsf(g) < PTC_RECOMPUTE_BELOW) {
      const int b = static_cast<int>(i / nn);
      const long r = i - b * nn;
      const int p = static_cast<int>(r / n), q = static_cast<int>(r - static_cast<long>(p) * n);
      const float* xb = x + static_cast<long>(b) * C * n;
      float acc = 0.0f;
      for (int c = 0; c < C; ++c) acc = fmaf(__ldg(xb + static_cast<long>(c) * n + p), __ldg(xb + static_cast<long>(c) * n + q), acc);
      g = acc * inv[b * n + p] * inv[b * n + q];
      G[i] = g;
    }
    g = fabsf(g);
    if (m == 1) {
      pos_sum += g;
      pos_cnt += 1.0f;
    } else if (m == 0) {
      neg_sum += g;
      neg_cnt += 1.0f;
    }
  }
  float r;
  r = block_sum(pos_sum, sh); if (threadIdx.x == 0) partials[4 * blockIdx.x + 0] = r;
  r = block_sum(neg_sum, sh); if (threadIdx.x == 0) partials[4 * blockIdx.x + 1] = r;
  r = block_sum(pos_cnt, sh); if (threadIdx.x == 0) partials[4 * blockIdx.x + 2] = r;
  r = block_sum(neg_cnt, sh); if (threadIdx.x == 0) partials[4 * blockIdx.x + 3] = r;
}

// T[p][q] = S[p][q] + S[q][p] (S as in ptc_bwd_gemm_kernel) as split planes [b*n][npad], pad columns zero.
__global__ void __launch_bounds__(256) ptc_dg_split_kernel(const float* __restrict__ G, const long long* __restrict__ mask,
                                                           const float* __restrict__ stats, const float* __restrict__ grad_out,
                                                           int n, int npad, __nv_bfloat16* __restrict__ t_hi,
                                                           __nv_bfloat16* __restrict__ t_lo) {
  __shared__ float tr[32][33];
  const int b = blockIdx.z, p0 = blockIdx.y * 32, q0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float g = grad_out[0];
  const float wpos = -0.5f / (stats[2] + 1.0f) * g, wneg = 0.5f / (stats[3] + 1.0f) * g;
  auto sfun = [&](int r, int c) -> float {
    if (r >= n || c >= n) return 0.0f;
    const long o = (static_cast<long>(b) * n + r) * n + c;
    const long long m = mask[o];
    if (m != 0 && m != 1) return 0.0f;
    const float gs = G[o];
    const float sg = gs > 0.0f ? 1.0f : (gs < 0.0f ? -1.0f : 0.0f);
    return sg * (m == 1 ? wpos : wneg);
  };
#pragma unroll
  for (int k = 0; k < 4; ++k) tr[ty + 8 * k][tx] = sfun(q0 + ty + 8 * k, p0 + tx);  // S[q][p], coalesced over p
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int pp = p0 + ty + 8 * k, q = q0 + tx;
    if (pp < n && q < npad) {
      const float v = sfun(pp, q) + tr[tx][ty + 8 * k];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const long o = (static_cast<long>(b) * n + pp) * npad + q;
      t_hi[o] = h;
      t_lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

// dxh rows [b*n][C] (token-major result of the GEMM) -> dx [b][C][n]:  dx = inv * (dxh - x_hat * <x_hat, dxh>)
__global__ void __launch_bounds__(256) ptc_norm_bwd_rows_kernel(const float* __restrict__ x, const float* __restrict__ inv,
                                                                const float* __restrict__ dxh, int C, int n,
                                                                float* __restrict__ dx) {
  __shared__ float tr[32][33];
  __shared__ float part[8][32];
  __shared__ float dots[32];
  const int b = blockIdx.y, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int p = p0 + tx;
  const float iv = p < n ? inv[b * n + p] : 0.0f;
  float dot = 0.0f;
  for (int c0 = 0; c0 < C; c0 += 32) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // dxh tile, coalesced over c
      const int pp = p0 + ty + 8 * k, c = c0 + tx;
      tr[ty + 8 * k][tx] = (pp < n && c < C) ? __ldg(dxh + (static_cast<long>(b) * n + pp) * C + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + ty + 8 * k;
      if (c < C && p < n) dot = fmaf(__ldg(x + (static_cast<long>(b) * C + c) * n + p) * iv, tr[tx][ty + 8 * k], dot);
    }
    __syncthreads();
  }
  part[ty][tx] = dot;
  __syncthreads();
  if (ty == 0) {
    float d = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) d += part[k][tx];
    dots[tx] = d;
  }
  __syncthreads();
  const float d = dots[tx];
  const bool clamped = iv >= 1e8f;  // ||x|| <= eps
  for (int c0 = 0; c0 < C; c0 += 32) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int pp = p0 + ty + 8 * k, c = c0 + tx;
      tr[ty + 8 * k][tx] = (pp < n && c < C) ? __ldg(dxh + (static_cast<long>(b) * n + pp) * C + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + ty + 8 * k;
      if (c < C && p < n) {
        const long o = (static_cast<long>(b) * C + c) * n + p;
        const float xh = __ldg(x + o) * iv;
        dx[o] = iv * (tr[tx][ty + 8 * k] - (clamped ? 0.0f : xh * d));
      }
    }
    __syncthreads();
  }
}

}  // namespace dupl

using namespace dupl;

extern "C" int dupl_seg_loss_fwd(const float* pred, const int64_t* label, int32_t b, int32_t C, int32_t H, int32_t W,
                                 int64_t ignore_index, float* lse, float* partials, float* stats, void* stream) {
  DUPL_CHECK_ARG(pred && label && lse && partials && stats && b > 0 && C > 0 && H > 0 && W > 0, "dupl_seg_loss_fwd: bad arguments");
  const long hw = static_cast<long>(H) * W, total = hw * b;
  const int blocks = static_cast<int>((total + 255) / 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  seg_ce_fwd_kernel<<<blocks, 256, 0, st>>>(pred, reinterpret_cast<const long long*>(label), C, hw, total, ignore_index, lse, partials);
  DUPL_LAUNCH_OK();
  seg_ce_finish_kernel<<<1, 256, 0, st>>>(partials, blocks, stats);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_seg_loss_bwd(const float* pred, const int64_t* label, const float* lse, const float* stats,
                                 const float* grad_out, int32_t b, int32_t C, int32_t H, int32_t W, int64_t ignore_index,
                                 float* dpred, void* stream) {
  DUPL_CHECK_ARG(pred && label && lse && stats && grad_out && dpred, "dupl_seg_loss_bwd: NULL pointer");
  const long hw = static_cast<long>(H) * W, total = hw * b;
  seg_ce_bwd_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, reinterpret_cast<const long long*>(label), lse, stats, grad_out, C, hw, total, ignore_index, dpred);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_seg_loss_up_fwd(const float* pred, const int64_t* label, int32_t b, int32_t C, int32_t h, int32_t w,
                                    int32_t H, int32_t W, int64_t ignore_index, float* lse, float* partials, float* stats,
                                    void* stream) {
  DUPL_CHECK_ARG(pred && label && lse && partials && stats && b > 0 && C > 0 && h > 0 && w > 0 && H > 0 && W > 0,
                 "dupl_seg_loss_up_fwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(cdiv(W, 32), cdiv(H, 32), b);
  seg_up_fwd_kernel<<<grid, 256, 0, st>>>(pred, reinterpret_cast<const long long*>(label), C, h, w, H, W,
                                          static_cast<float>(h) / H, static_cast<float>(w) / W, ignore_index, lse, partials);
  DUPL_LAUNCH_OK();
  seg_ce_finish_kernel<<<1, 256, 0, st>>>(partials, static_cast<int>(grid.x * grid.y * grid.z), stats);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_seg_loss_up_bwd(const float* pred, const int64_t* label, const float* lse, const float* stats,
                                    const float* grad_out, int32_t b, int32_t C, int32_t h, int32_t w, int32_t H, int32_t W,
                                    int64_t ignore_index, float* dpred, void* stream) {
  DUPL_CHECK_ARG(pred && label && lse && stats && grad_out && dpred, "dupl_seg_loss_up_bwd: NULL pointer");
  DUPL_CHECK_ARG(C <= SEGUP_MAXC, "dupl_seg_loss_up_bwd: C=%d > %d", C, SEGUP_MAXC);
  // label pixels that can touch one low-resolution cell: source coordinate within (c-1, c+1) -> 2/scale (+ slack)
  const int fy = min(H, static_cast<int>(2.0f * H / h) + 4), fx = min(W, static_cast<int>(2.0f * W / w) + 4);
  seg_up_bwd_kernel<<<dim3(w, h, b), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, reinterpret_cast<const long long*>(label), lse, stats, grad_out, C, h, w, H, W, static_cast<float>(h) / H,
      static_cast<float>(w) / W, fy, fx, ignore_index, dpred);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_ptc_loss_fwd(const float* x, const int64_t* mask, int32_t b, int32_t C, int32_t n, float* inv,
                                 float* Gs, float* partials, float* stats, void* stream) {
  DUPL_CHECK_ARG(x && mask && inv && Gs && partials && stats && b > 0 && C > 0 && n > 0, "dupl_ptc_loss_fwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ptc_invnorm_kernel<<<cdiv(b * n, 256), 256, 0, st>>>(x, C, n, b * n, inv);
  DUPL_LAUNCH_OK();
  dim3 grid(cdiv(n, PT), cdiv(n, PT), b);
  ptc_gram_kernel<<<grid, 256, 0, st>>>(x, inv, reinterpret_cast<const long long*>(mask), C, n, Gs, partials);
  DUPL_LAUNCH_OK();
  ptc_finish_kernel<<<1, 256, 0, st>>>(partials, static_cast<int>(grid.x * grid.y * grid.z), stats);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_ptc_loss_bwd(const float* x, const int64_t* mask, const float* inv, const float* Gs, const float* stats,
                                 const float* grad_out, int32_t b, int32_t C, int32_t n, float* dxh_scratch, float* dx,
                                 void* stream) {
  DUPL_CHECK_ARG(x && mask && inv && Gs && stats && grad_out && dxh_scratch && dx, "dupl_ptc_loss_bwd: NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(cdiv(n, PT), cdiv(C, PT), b);
  ptc_bwd_gemm_kernel<<<grid, 256, 0, st>>>(x, inv, reinterpret_cast<const long long*>(mask), Gs, stats, grad_out, C, n, dxh_scratch);
  DUPL_LAUNCH_OK();
  ptc_bwd_norm_kernel<<<cdiv(b * n, 256), 256, 0, st>>>(x, inv, dxh_scratch, C, n, b * n, dx);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_ptc_prepare(const float* x, int32_t b, int32_t C, int32_t n, int32_t npad, float* inv, void* rows_hi,
                                void* rows_lo, void* cm_hi, void* cm_lo, void* stream) {
  DUPL_CHECK_ARG(x && inv && rows_hi && rows_lo && cm_hi && cm_lo && b > 0 && C > 0 && n > 0 && npad >= n,
                 "dupl_ptc_prepare: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ptc_invnorm_kernel<<<cdiv(b * n, 256), 256, 0, st>>>(x, C, n, b * n, inv);
  DUPL_LAUNCH_OK();
  ptc_normalize_split_kernel<<<dim3(cdiv(npad, 32), cdiv(C, 32), b), 256, 0, st>>>(
      x, inv, C, n, npad, static_cast<__nv_bfloat16*>(rows_hi), static_cast<__nv_bfloat16*>(rows_lo),
      static_cast<__nv_bfloat16*>(cm_hi), static_cast<__nv_bfloat16*>(cm_lo));
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_ptc_mask_reduce(float* G, const int64_t* mask, const float* x, const float* inv, int32_t b, int32_t C,
                                    int32_t n, float* partials, int32_t nblocks, float* stats, void* stream) {
  DUPL_CHECK_ARG(G && mask && x && inv && partials && stats && b > 0 && C > 0 && n > 0 && nblocks > 0,
                 "dupl_ptc_mask_reduce: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ptc_mask_reduce_kernel<<<nblocks, 256, 0, st>>>(G, reinterpret_cast<const long long*>(mask), x, inv, C, n,
                                                  static_cast<long>(b) * n * n, partials);
  DUPL_LAUNCH_OK();
  ptc_finish_kernel<<<1, 256, 0, st>>>(partials, nblocks, stats);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_ptc_dg(const float* G, const int64_t* mask, const float* stats, const float* grad_out, int32_t b,
                           int32_t n, int32_t npad, void* t_hi, void* t_lo, void* stream) {
  DUPL_CHECK_ARG(G && mask && stats && grad_out && t_hi && t_lo && b > 0 && n > 0 && npad >= n, "dupl_ptc_dg: bad arguments");
  ptc_dg_split_kernel<<<dim3(cdiv(npad, 32), cdiv(n, 32), b), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      G, reinterpret_cast<const long long*>(mask), stats, grad_out, n, npad, static_cast<__nv_bfloat16*>(t_hi),
      static_cast<__nv_bfloat16*>(t_lo));
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_ptc_norm_bwd_rows(const float* x, const float* inv, const float* dxh_rows, int32_t b, int32_t C, int32_t n,
                                      float* dx, void* stream) {
  DUPL_CHECK_ARG(x && inv && dxh_rows && dx && b > 0 && C > 0 && n > 0, "dupl_ptc_norm_bwd_rows: bad arguments");
  ptc_norm_bwd_rows_kernel<<<dim3(cdiv(n, 32), b), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, inv, dxh_rows, C, n, dx);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}
