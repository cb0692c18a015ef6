// Bilinear sampling with torch's align_corners=False convention (aten UpSampleBilinear2d):
//   src = max((dst + 0.5) * in/out - 0.5, 0),  i0 = floor(src), i1 = min(i0 + 1, in - 1),
//   l1 = src - i0, l0 = 1 - l1
// and the fma ordering that reproduces torch-CPU's results bit for bit on this path
// (SURVEY.md §8(c)):  t = fma(lx0, v0, lx1*v1) per row, out = fma(ly0, t0, ly1*t1).
// Explicit _rn intrinsics stop the compiler from re-contracting the expression.
#pragma once
#include <cuda_runtime.h>

namespace dupl {

struct Lin {
  int i0, i1;
  float l0, l1;
};

__device__ __forceinline__ Lin lin_coord(int dst, int in_size, float scale) {
  const float src = fmaxf(__fsub_rn(__fmul_rn(static_cast<float>(dst) + 0.5f, scale), 0.5f), 0.0f);
  Lin r;
  r.i0 = min(static_cast<int>(src), in_size - 1);
  r.i1 = min(r.i0 + 1, in_size - 1);
  r.l1 = src - static_cast<float>(r.i0);
  r.l0 = 1.0f - r.l1;
  return r;
}

__device__ __forceinline__ float bilerp4(float v00, float v01, float v10, float v11, const Lin& y, const Lin& x) {
  const float t0 = __fmaf_rn(x.l0, v00, __fmul_rn(x.l1, v01));
  const float t1 = __fmaf_rn(x.l0, v10, __fmul_rn(x.l1, v11));
  return __fmaf_rn(y.l0, t0, __fmul_rn(y.l1, t1));
}

__device__ __forceinline__ float bilerp(const float* __restrict__ plane, int W, const Lin& y, const Lin& x) {
  const float* r0 = plane + static_cast<long>(y.i0) * W;
  const float* r1 = plane + static_cast<long>(y.i1) * W;
  return bilerp4(__ldg(r0 + x.i0), __ldg(r0 + x.i1), __ldg(r1 + x.i0), __ldg(r1 + x.i1), y, x);
}

__device__ __forceinline__ float bilerp_smem(const float* plane, int W, const Lin& y, const Lin& x) {
  const float* r0 = plane + y.i0 * W;
  const float* r1 = plane + y.i1 * W;
  return bilerp4(r0[x.i0], r0[x.i1], r1[x.i0], r1[x.i1], y, x);
}

}  // namespace dupl
