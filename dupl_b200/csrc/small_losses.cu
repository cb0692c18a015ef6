// The two script-side losses of the training loop that round 1 left to stock torch ops (kernels G11 / G20):
//   classification  train_final_voc.py:299-305   4 x F.multilabel_soft_margin_loss(logits[b,K], cls_label)
//   discrepancy     train_final_voc.py:440-447   (1 + mean cos(f1.detach(), f2)) + (1 + mean cos(f2.detach(), f1)),
//                                                cosine along the SPATIAL axis of [b, 768, 784], eps 1e-6
// Each is one forward kernel (value) and one backward kernel (gradients scaled by the upstream gradient read on the device:
// no host synchronisation); reductions run in a fixed order (bit-reproducible).
#include <math.h>

#include "common.cuh"

namespace dupl {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// logsigmoid(x) = min(x, 0) - log1p(exp(-|x|))  (torch's formula)
__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.0f) - log1pf(expf(-fabsf(x))); }

struct ClsPtrs {   // by value in the kernel parameters: no device-side pointer table, nothing to copy before a (captured) launch
  const float* x[8];
  float* g[8];
};

// loss = sum_t mean_{b,k} -[y logsig(x_t) + (1 - y) logsig(-x_t)] over the T logit tensors; one block.
__global__ void __launch_bounds__(256) cls_loss_fwd_kernel(const ClsPtrs ptrs, int T,
                                                           const float* __restrict__ label, int n, float* __restrict__ loss) {
  __shared__ float sh[8];
  float total = 0.0f;
  for (int t = 0; t < T; ++t) {           // one mean per tensor, then their sum: the order of the script
    float acc = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float x = ptrs.x[t][i], y = label[i];
      acc -= y * log_sigmoid(x) + (1.0f - y) * log_sigmoid(-x);
    }
    acc = warp_sum_f(acc);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.0f;
      for (int w = 0; w < 8; ++w) s += sh[w];
      total += s / static_cast<float>(n);
    }
  }
  if (threadIdx.x == 0) *loss = total;
}

// d loss / d x_t[i] = gout * (sigmoid(x) - y) / n
__global__ void __launch_bounds__(256) cls_loss_bwd_kernel(const ClsPtrs ptrs, int T, const float* __restrict__ label, int n,
                                                           const float* __restrict__ gout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (i >= n || t >= T) return;
  const float x = ptrs.x[t][i];
  const float s = 1.0f / (1.0f + expf(-x));
  ptrs.g[t][i] = __ldg(gout) * (s - label[i]) / static_cast<float>(n);
}

// One warp per (image, channel) row of n spatial positions: cos = (x . y) / (max(|x|, eps) max(|y|, eps))
// (torch.nn.functional.cosine_similarity: both vectors are normalised first, norms clamped at eps).
template <bool BWD>
__global__ void __launch_bounds__(256) sim_rows_kernel(const float* __restrict__ f1, const float* __restrict__ f2, int rows, int n,
                                                       float eps, float* __restrict__ cos_out, const float* __restrict__ gout,
                                                       float* __restrict__ d1, float* __restrict__ d2) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = f1 + static_cast<long>(row) * n;
  const float* y = f2 + static_cast<long>(row) * n;
  float sxy = 0.0f, sxx = 0.0f, syy = 0.0f;
  for (int i = lane; i < n; i += 32) {
    const float a = x[i], b = y[i];
    sxy = fmaf(a, b, sxy);
    sxx = fmaf(a, a, sxx);
    syy = fmaf(b, b, syy);
  }
  sxy = warp_sum_f(sxy);
  sxx = warp_sum_f(sxx);
  syy = warp_sum_f(syy);
  const float nx = sqrtf(sxx), ny = sqrtf(syy);
  const float cx = fmaxf(nx, eps), cy = fmaxf(ny, eps);
  const float c = sxy / (cx * cy);
  if (!BWD) {
    if (lane == 0) cos_out[row] = c;
    return;
  }
  // loss = 2 + (2 / rows) sum_r cos_r, split as in the script: the first term differentiates w.r.t. f2 only, the second w.r.t. f1
  // only; each contributes gout / rows * d cos / d(.).  d cos/dy = x / (cx cy) - [ny > eps] cos * y / ny^2 (clamped norm: constant)
  const float g = __ldg(gout) / static_cast<float>(rows);
  const float inv = 1.0f / (cx * cy);
  const float kx = nx > eps ? c / (nx * nx) : 0.0f, ky = ny > eps ? c / (ny * ny) : 0.0f;
  float* o1 = d1 + static_cast<long>(row) * n;
  float* o2 = d2 + static_cast<long>(row) * n;
  for (int i = lane; i < n; i += 32) {
    const float a = x[i], b = y[i];
    o1[i] = g * (b * inv - kx * a);
    o2[i] = g * (a * inv - ky * b);
  }
}

// loss = 2 + 2 * mean(cos): fixed-order sum of the per-row cosines by one block
__global__ void __launch_bounds__(256) sim_finish_kernel(const float* __restrict__ cosv, int rows, float* __restrict__ loss) {
  __shared__ float sh[8];
  float acc = 0.0f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) acc += cosv[i];
  acc = warp_sum_f(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int w = 0; w < 8; ++w) s += sh[w];
    const float m = s / static_cast<float>(rows);
    *loss = (1.0f + m) + (1.0f + m);
  }
}

}  // namespace dupl

using namespace dupl;

extern "C" int dupl_cls_loss_fwd(const float* const* logits, int32_t T, const float* label, int32_t n, float* loss, void* stream) {
  DUPL_CHECK_ARG(logits && label && loss && T >= 1 && T <= 8 && n > 0, "dupl_cls_loss_fwd: bad arguments");
  ClsPtrs P = {};
  for (int t = 0; t < T; ++t) {
    DUPL_CHECK_ARG(logits[t] != nullptr, "dupl_cls_loss_fwd: logits[%d] is NULL", t);
    P.x[t] = logits[t];
  }
  cls_loss_fwd_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(P, T, label, n, loss);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_cls_loss_bwd(const float* const* logits, float* const* grads, int32_t T, const float* label, int32_t n,
                                 const float* grad_out, void* stream) {
  DUPL_CHECK_ARG(logits && grads && label && grad_out && T >= 1 && T <= 8 && n > 0, "dupl_cls_loss_bwd: bad arguments");
  ClsPtrs P = {};
  for (int t = 0; t < T; ++t) {
    DUPL_CHECK_ARG(logits[t] != nullptr && grads[t] != nullptr, "dupl_cls_loss_bwd: NULL tensor %d", t);
    P.x[t] = logits[t];
    P.g[t] = grads[t];
  }
  cls_loss_bwd_kernel<<<dim3(cdiv(n, 256), T), 256, 0, static_cast<cudaStream_t>(stream)>>>(P, T, label, n, grad_out);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_sim_loss_fwd(const float* f1, const float* f2, int32_t rows, int32_t n, float eps, float* cos_rows, float* loss,
                                 void* stream) {
  DUPL_CHECK_ARG(f1 && f2 && cos_rows && loss && rows > 0 && n > 0, "dupl_sim_loss_fwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  sim_rows_kernel<false><<<cdiv(rows, 8), 256, 0, st>>>(f1, f2, rows, n, eps, cos_rows, nullptr, nullptr, nullptr);
  DUPL_LAUNCH_OK();
  sim_finish_kernel<<<1, 256, 0, st>>>(cos_rows, rows, loss);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_sim_loss_bwd(const float* f1, const float* f2, int32_t rows, int32_t n, float eps, const float* grad_out, float* d1,
                                 float* d2, void* stream) {
  DUPL_CHECK_ARG(f1 && f2 && grad_out && d1 && d2 && rows > 0 && n > 0, "dupl_sim_loss_bwd: bad arguments");
  sim_rows_kernel<true><<<cdiv(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(f1, f2, rows, n, eps, nullptr, grad_out, d1, d2);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}
