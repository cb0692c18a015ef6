// Fused multi-tensor PolyWarmupAdamW (reference utils/optimizer.py:38-68 = torch.optim.AdamW whose step()
// first sets the warm-up / polynomial learning rate; utils/train_helper.py:21-52: 4 parameter groups, heads and decoders at
// 10x the learning rate, weight decay on everything).
//
// One launch updates EVERY parameter tensor of both students from a table of raw pointers (parameter, gradient = view of the
// student's flat gradient arena, exp_avg, exp_avg_sq) and, for the weights the tcgen05 GEMMs consume, writes the split-bf16
// (hi, lo) planes in the same pass — the 98 per-tensor re-split launches and torch's multi_tensor_apply AdamW (14 launches,
// 1.2 ms) become one HBM-bound kernel: 16 B read + 12..16 B written per element.
// Per-parameter step counts (a parameter the loss does not reach in a phase is skipped entirely, like torch.optim.AdamW
// skips p.grad is None: no decay, no moment update, no step increment) live on the device, so the launch is CUDA-graph
// replayable; the schedule multiplier is a device scalar the host sets before each step.
#include <math.h>

#include "common.cuh"
#include "ptx.cuh"

namespace dupl {

constexpr int OPT_THREADS = 256, OPT_PER_THREAD = 8, OPT_CHUNK = OPT_THREADS * OPT_PER_THREAD;

// update of one parameter tensor this step (bias corrections of ITS step count)
__global__ void __launch_bounds__(320) adamw_advance_kernel(const dupl_adamw_param* __restrict__ params, int n_params,
                                                            const int32_t* __restrict__ active, int32_t* __restrict__ steps,
                                                            float* __restrict__ coef, double beta1, double beta2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_params) return;
  if (!active[i]) return;
  const int t = steps[i] + 1;
  steps[i] = t;
  // torch/optim/adamw.py (_single_tensor_adamw / _multi_tensor_adamw): Python doubles
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(t));
  const double bc2_sqrt = sqrt(1.0 - pow(beta2, static_cast<double>(t)));
  coef[2 * i] = static_cast<float>(1.0 / bc1);   // step_size = lr / bias_correction1 (lr applied in the update kernel)
  coef[2 * i + 1] = static_cast<float>(bc2_sqrt);
}

__global__ void __launch_bounds__(OPT_THREADS) adamw_update_kernel(const dupl_adamw_param* __restrict__ params,
                                                                   const int2* __restrict__ items,
                                                                   const int32_t* __restrict__ active,
                                                                   const float* __restrict__ coef,
                                                                   const float* __restrict__ lr_scale, float beta1, float beta2,
                                                                   float eps, float weight_decay) {
  const int2 it = items[blockIdx.x];
  if (!active[it.x]) return;
  const dupl_adamw_param P = params[it.x];
  const float lr = P.lr * __ldg(lr_scale);
  const float decay = 1.0f - lr * weight_decay;
  const float step_size = lr * coef[2 * it.x];
  const float bc2_sqrt = coef[2 * it.x + 1];
  const float w1 = 1.0f - beta1, w2 = 1.0f - beta2;
  float* __restrict__ p = static_cast<float*>(P.param);
  const float* __restrict__ g = static_cast<const float*>(P.grad);
  float* __restrict__ m = static_cast<float*>(P.exp_avg);
  float* __restrict__ v = static_cast<float*>(P.exp_avg_sq);
  __nv_bfloat16* __restrict__ hi = static_cast<__nv_bfloat16*>(P.plane_hi);
  __nv_bfloat16* __restrict__ lo = static_cast<__nv_bfloat16*>(P.plane_lo);
  const long base = static_cast<long>(it.y) * OPT_CHUNK + threadIdx.x * 4;
#pragma unroll
  for (int r = 0; r < OPT_PER_THREAD / 4; ++r) {
    const long i = base + static_cast<long>(r) * OPT_THREADS * 4;
    if (i >= P.numel) break;
    // every tensor starts on a 16-byte boundary and numel % 4 == 0 (checked on the host)
    float4 p4 = *reinterpret_cast<const float4*>(p + i);
    const float4 g4 = *reinterpret_cast<const float4*>(g + i);
    float4 m4 = *reinterpret_cast<const float4*>(m + i);
    float4 v4 = *reinterpret_cast<const float4*>(v + i);
    float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      pp[e] = __fmul_rn(pp[e], decay);                                   // param.mul_(1 - lr * weight_decay)
      mm[e] = __fmaf_rn(w1, __fsub_rn(gg[e], mm[e]), mm[e]);             // exp_avg.lerp_(grad, 1 - beta1)
      vv[e] = __fmaf_rn(__fmul_rn(w2, gg[e]), gg[e], __fmul_rn(vv[e], beta2));  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv[e]), bc2_sqrt), eps);
      pp[e] = __fmaf_rn(-step_size, __fdiv_rn(mm[e], denom), pp[e]);    // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
    *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (hi != nullptr) {
      uint32_t h0, l0, h1, l1;
      split2_bf16(pp[0], pp[1], h0, l0);
      split2_bf16(pp[2], pp[3], h1, l1);
      *reinterpret_cast<uint2*>(hi + i) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(lo + i) = make_uint2(l0, l1);
    }
  }
}

}  // namespace dupl

extern "C" int dupl_adamw_items(const int64_t* numel, int32_t n_params, int32_t* items_xy, int64_t capacity, int64_t* n_items) {
  using namespace dupl;
  DUPL_CHECK_ARG(numel != nullptr && n_params > 0 && n_items != nullptr, "dupl_adamw_items: bad arguments");
  int64_t n = 0;
  for (int i = 0; i < n_params; ++i) {
    DUPL_CHECK_ARG(numel[i] > 0 && numel[i] % 4 == 0, "dupl_adamw_items: parameter %d has %lld elements (must be a positive multiple of 4)", i,
                   static_cast<long long>(numel[i]));
    const int64_t chunks = (numel[i] + OPT_CHUNK - 1) / OPT_CHUNK;
    for (int64_t c = 0; c < chunks; ++c, ++n)
      if (items_xy != nullptr && n < capacity) {
        items_xy[2 * n] = i;
        items_xy[2 * n + 1] = static_cast<int32_t>(c);
      }
  }
  *n_items = n;
  return DUPL_OK;
}

extern "C" int dupl_adamw_step(const dupl_adamw_args* a, void* stream) {
  using namespace dupl;
  DUPL_CHECK_ARG(a != nullptr && a->params && a->items && a->active && a->steps && a->coef && a->lr_scale,
                 "dupl_adamw_step: NULL pointer");
  DUPL_CHECK_ARG(a->n_params > 0 && a->n_items > 0, "dupl_adamw_step: empty work list");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  adamw_advance_kernel<<<cdiv(a->n_params, 320), 320, 0, st>>>(a->params, a->n_params, a->active, a->steps, a->coef,
                                                              static_cast<double>(a->beta1), static_cast<double>(a->beta2));
  DUPL_LAUNCH_OK();
  adamw_update_kernel<<<static_cast<unsigned>(a->n_items), OPT_THREADS, 0, st>>>(
      a->params, reinterpret_cast<const int2*>(a->items), a->active, a->coef, a->lr_scale, a->beta1, a->beta2, a->eps, a->weight_decay);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}
