// Kernels of the decoder / classification heads: data movement around the tcgen05
// GEMM that executes the LargeFOV convolutions as implicit GEMMs (model/decoder/conv_head.py:33-41),
// and the global-max-pool + 1x1 classifier of model_dupl.py:88-95.
#include "common.cuh"
#include "ptx.cuh"

namespace dupl {

// ---------------------------------------------------------------------------------------------
// im2col for a 3x3 convolution with dilation `dil`, zero padding `dil`, stride 1, on TOKEN-MAJOR
// (NHWC) split-bf16 planes: out[(b*gh*gw + y*gw + x)][tap*Cin + c] = in[row(b, y+dy, x+dx)][c].
// Pure 16-byte copies (8 bf16 channels per thread); rows outside the image are zero.
// row(b, y, x) = row_offset + b*row_stride + first + y*gw + x  (first = 1 skips a cls token).
// ---------------------------------------------------------------------------------------------
struct Im2colParams {
  const uint4* in_hi;
  const uint4* in_lo;
  uint4* out_hi;
  uint4* out_lo;
  int B, gh, gw, cin8 /* Cin/8 */, dil, row_offset, row_stride, first;
};

__global__ void __launch_bounds__(256) im2col3x3_kernel(Im2colParams p) {
  const long total = static_cast<long>(p.B) * p.gh * p.gw * 9 * p.cin8;
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c8 = static_cast<int>(idx % p.cin8);
  long t = idx / p.cin8;
  const int tap = static_cast<int>(t % 9);
  t /= 9;
  const int x = static_cast<int>(t % p.gw);
  t /= p.gw;
  const int y = static_cast<int>(t % p.gh);
  const int b = static_cast<int>(t / p.gh);
  const int yy = y + (tap / 3 - 1) * p.dil, xx = x + (tap % 3 - 1) * p.dil;
  uint4 h = make_uint4(0, 0, 0, 0), l = h;
  if (yy >= 0 && yy < p.gh && xx >= 0 && xx < p.gw) {
    const long row = p.row_offset + static_cast<long>(b) * p.row_stride + p.first + yy * p.gw + xx;
    h = __ldg(p.in_hi + row * p.cin8 + c8);
    l = __ldg(p.in_lo + row * p.cin8 + c8);
  }
  p.out_hi[idx] = h;
  p.out_lo[idx] = l;
}

// ---------------------------------------------------------------------------------------------
// rows -> NCHW: out[b][c][p] = src[row(b, p)][c] for c < C (src row stride ld), via 32x32 smem tiles.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rows_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ out, int np,
                                                           int C, int ld, int row_offset, int row_stride, int first) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    const long row = row_offset + static_cast<long>(b) * row_stride + first + p;
    tile[r][threadIdx.x] = (p < np && c < C) ? src[row * ld + c] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (c < C && p < np) out[(static_cast<long>(b) * C + c) * np + p] = tile[threadIdx.x][r];
  }
}

// ---------------------------------------------------------------------------------------------
// Global max pool over the patch tokens of each image + bias-free 1x1 classifier:
// logits[b][k] = sum_d (max_p x[row(b,p)][d]) * w[k][d]      (model_dupl.py:88-95)
// Two launches: max-pool with grid (D/32, B) into `pooled` [B][D] (8 row lanes per column, first maximum wins like
// F.adaptive_max_pool2d; arg-max rows kept for the backward pass when `argmax` != NULL), then one block per image
// for the K dot products.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gmp_pool_kernel(const float* __restrict__ x, float* __restrict__ pooled,
                                                       int* __restrict__ argmax, int np, int D, int row_offset, int row_stride,
                                                       int first) {
  __shared__ float s_best[8][33];
  __shared__ int s_arg[8][33];
  const int b = blockIdx.y, d = blockIdx.x * 32 + threadIdx.x;
  const long row0 = row_offset + static_cast<long>(b) * row_stride + first;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  if (d < D)
    for (int p = threadIdx.y; p < np; p += 8) {
      const float v = __ldg(x + (row0 + p) * D + d);
      if (v > best || arg == 0x7fffffff) {
        best = v;
        arg = p;
      }
    }
  s_best[threadIdx.y][threadIdx.x] = best;
  s_arg[threadIdx.y][threadIdx.x] = arg;
  __syncthreads();
  if (threadIdx.y == 0 && d < D) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float v = s_best[k][threadIdx.x];
      const int a = s_arg[k][threadIdx.x];
      if (v > best || (v == best && a < arg)) {
        best = v;
        arg = a;
      }
    }
    pooled[static_cast<long>(b) * D + d] = best;
    if (argmax != nullptr) argmax[static_cast<long>(b) * D + d] = arg;
  }
}

__global__ void __launch_bounds__(256) gmp_logits_kernel(const float* __restrict__ pooled, const float* __restrict__ w,
                                                         float* __restrict__ logits, int D, int K) {
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int k = warp; k < K; k += nwarps) {
    float acc = 0.0f;
    for (int d = lane; d < D; d += 32) acc = fmaf(pooled[static_cast<long>(b) * D + d], __ldg(w + static_cast<long>(k) * D + d), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) logits[static_cast<long>(b) * K + k] = acc;
  }
}

}  // namespace dupl

using namespace dupl;

extern "C" int dupl_im2col3x3(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int32_t B, int32_t gh,
                              int32_t gw, int32_t Cin, int32_t dilation, int32_t row_offset, int32_t row_stride,
                              int32_t first, void* stream) {
  DUPL_CHECK_ARG(in_hi && in_lo && out_hi && out_lo, "dupl_im2col3x3: NULL pointer");
  DUPL_CHECK_ARG(B > 0 && gh > 0 && gw > 0 && Cin > 0 && Cin % 8 == 0 && dilation > 0, "dupl_im2col3x3: bad shape");
  Im2colParams p;
  p.in_hi = static_cast<const uint4*>(in_hi);
  p.in_lo = static_cast<const uint4*>(in_lo);
  p.out_hi = static_cast<uint4*>(out_hi);
  p.out_lo = static_cast<uint4*>(out_lo);
  p.B = B; p.gh = gh; p.gw = gw; p.cin8 = Cin / 8; p.dil = dilation;
  p.row_offset = row_offset; p.row_stride = row_stride; p.first = first;
  const long total = static_cast<long>(B) * gh * gw * 9 * p.cin8;
  im2col3x3_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_rows_to_nchw(const float* src, float* out, int32_t B, int32_t np, int32_t C, int32_t ld,
                                 int32_t row_offset, int32_t row_stride, int32_t first, void* stream) {
  DUPL_CHECK_ARG(src && out && B > 0 && np > 0 && C > 0 && ld >= C, "dupl_rows_to_nchw: bad arguments");
  dim3 grid(cdiv(np, 32), cdiv(C, 32), B), block(32, 8);
  rows_to_nchw_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(src, out, np, C, ld, row_offset, row_stride, first);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_gmp_classify(const float* x, const float* w, float* logits, float* pooled, int32_t* argmax, int32_t B,
                                 int32_t np, int32_t D, int32_t K, int32_t row_offset, int32_t row_stride, int32_t first,
                                 void* stream) {
  DUPL_CHECK_ARG(x && w && logits && pooled && B > 0 && np > 0 && D > 0 && K > 0, "dupl_gmp_classify: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  gmp_pool_kernel<<<dim3(cdiv(D, 32), B), dim3(32, 8), 0, st>>>(x, pooled, argmax, np, D, row_offset, row_stride, first);
  DUPL_LAUNCH_OK();
  gmp_logits_kernel<<<B, 256, 0, st>>>(pooled, w, logits, D, K);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}
