// Strong augmentation of the training loop on the GPU: imutils.augment_data_strong
// (utils/imutils.py:305-317) = per image ToPILImage -> RandAugment(n, m) (utils/randomaug.py:161-262: n operations drawn from
// AutoContrast, Equalize, Posterize, Color, Contrast, Brightness, Sharpness) -> ToTensor -> Normalize -> horizontal flip.
// The reference does this on the host with Pillow in EVERY iteration (GPU -> CPU -> GPU round trip, :190-191).  Here the
// images never leave the device; the operation indices are drawn on the host with the same `random.choices` call (no device
// data needed for that) and every operation reproduces Pillow's integer / float arithmetic bit for bit
// (oracle/randaug_ref.py, pinned against Pillow; tests/test_gpu_augment.py compares the kernels with Pillow itself).
//
// Layout: uint8 HWC images [B][H][W][3], two ping-pong buffers.  Per operation step: statistics (per-band histograms and the
// luma sum: the inputs of AutoContrast / Equalize / Contrast), a one-block-per-image kernel that turns them into look-up tables
// and the contrast mean exactly as ImageOps / ImageEnhance do (double arithmetic where Python uses floats), and one
// element-wise (3x3 for Sharpness) pass.  HBM-bound: 3 B read + 3 B written per pixel and step.
#include <stdint.h>

#include "common.cuh"

namespace dupl {

enum { AUG_AUTOCONTRAST = 0, AUG_EQUALIZE, AUG_POSTERIZE, AUG_COLOR, AUG_CONTRAST, AUG_BRIGHTNESS, AUG_SHARPNESS, AUG_NOPS };

struct AugWs {
  uint8_t* img[2];        // [B][H][W][3]
  unsigned int* hist;     // [B][3][256]
  unsigned long long* lsum;  // [B]
  uint8_t* lut;           // [B][3][256]
  int* mean;              // [B]
};

struct AugVals {
  float v[AUG_NOPS];      // magnitude of every operation (utils/randomaug.py:262)
};

// ToPILImage of a float tensor: pic.mul(255).byte()  (truncation), CHW -> HWC
__global__ void __launch_bounds__(256) aug_from_float_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, int B, int H,
                                                             int W) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const long hw = static_cast<long>(H) * W;
  if (i >= B * hw) return;
  const long b = i / hw, p = i % hw;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = __fmul_rn(in[(b * 3 + c) * hw + p], 255.0f);
    out[i * 3 + c] = static_cast<uint8_t>(__float2int_rz(v));
  }
}

__device__ __forceinline__ int luma_u8(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// per-band histograms + sum of the luma image ("L" conversion) of every image
__global__ void __launch_bounds__(256) aug_stats_kernel(const uint8_t* __restrict__ img, unsigned int* __restrict__ hist,
                                                        unsigned long long* __restrict__ lsum, long hw) {
  __shared__ unsigned int sh[3 * 256];
  __shared__ unsigned long long sl;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 768; i += blockDim.x) sh[i] = 0;
  if (threadIdx.x == 0) sl = 0;
  __syncthreads();
  const uint8_t* px = img + static_cast<long>(b) * hw * 3;
  unsigned long long mine = 0;
  for (long p = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; p < hw; p += static_cast<long>(gridDim.x) * blockDim.x) {
    const int r = px[p * 3], g = px[p * 3 + 1], bl = px[p * 3 + 2];
    atomicAdd(&sh[r], 1u);
    atomicAdd(&sh[256 + g], 1u);
    atomicAdd(&sh[512 + bl], 1u);
    mine += luma_u8(r, g, bl);
  }
  atomicAdd(&sl, mine);
  __syncthreads();
  for (int i = threadIdx.x; i < 768; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[b * 768 + i], sh[i]);
  if (threadIdx.x == 0 && sl) atomicAdd(&lsum[b], sl);
}

// one block per image: the look-up tables of ImageOps.autocontrast / equalize / posterize and ImageEnhance.Contrast's mean
__global__ void __launch_bounds__(256) aug_params_kernel(const int* __restrict__ ops, const unsigned int* __restrict__ hist,
                                                         const unsigned long long* __restrict__ lsum, uint8_t* __restrict__ lut,
                                                         int* __restrict__ mean, long hw, AugVals vals) {
  const int b = blockIdx.x;
  const int op = ops[b];
  const int ix = threadIdx.x;  // 256 threads: LUT entry
  if (op == AUG_CONTRAST && ix == 0) {
    // mean = int(ImageStat.Stat(image.convert("L")).mean[0] + 0.5): Python floats
    mean[b] = static_cast<int>(__dadd_rn(__ddiv_rn(static_cast<double>(lsum[b]), static_cast<double>(hw)), 0.5));
  }
  if (op > AUG_POSTERIZE) return;
  for (int c = 0; c < 3; ++c) {
    const unsigned int* h = hist + (b * 3 + c) * 256;
    int out = ix;
    if (op == AUG_POSTERIZE) {
      int bits = static_cast<int>(vals.v[AUG_POSTERIZE]);          // Posterize: v = int(v); v = max(1, v)
      bits = bits < 1 ? 1 : bits;
      out = ix & (~((1 << (8 - bits)) - 1) & 0xFF);
    } else if (op == AUG_AUTOCONTRAST) {
      int lo = 0, hi = 255;
      while (lo < 256 && !h[lo]) ++lo;
      while (hi >= 0 && !h[hi]) --hi;
      if (hi > lo) {
        const double scale = __ddiv_rn(255.0, static_cast<double>(hi - lo));
        const double offset = __dmul_rn(-static_cast<double>(lo), scale);
        const int v = static_cast<int>(__dadd_rn(__dmul_rn(static_cast<double>(ix), scale), offset));  // int(): toward zero
        out = v < 0 ? 0 : (v > 255 ? 255 : v);
      }
    } else {  // AUG_EQUALIZE
      long total = 0, before = 0;
      int nonzero = 0, last = 0;
      for (int i = 0; i < 256; ++i) {
        const unsigned int f = h[i];
        if (f) {
          ++nonzero;
          last = f;
        }
        total += f;
        if (i < ix) before += f;
      }
      const long step = (total - last) / 255;
      if (nonzero > 1 && step) {
        const long v = (step / 2 + before) / step;
        out = v > 255 ? 255 : static_cast<int>(v);               // Image.point() clips list entries
      }
    }
    lut[(b * 3 + c) * 256 + ix] = static_cast<uint8_t>(out);
  }
}

// Image.blend(degenerate, image, alpha): (UINT8)(d + alpha * (x - d)) in C float arithmetic (no contraction); clipped outside [0,1]
__device__ __forceinline__ uint8_t blend_u8(int d, int x, float alpha, bool in_range) {
  const float t = __fadd_rn(static_cast<float>(d), __fmul_rn(alpha, static_cast<float>(x - d)));
  if (in_range) return static_cast<uint8_t>(__float2int_rz(t));
  return t <= 0.0f ? 0 : (t >= 255.0f ? 255 : static_cast<uint8_t>(__float2int_rz(t)));
}

__global__ void __launch_bounds__(256) aug_apply_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                        const int* __restrict__ ops, const uint8_t* __restrict__ lut,
                                                        const int* __restrict__ mean, int H, int W, AugVals vals) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const int op = ops[b];
  const long base = static_cast<long>(b) * H * W * 3;
  const uint8_t* px = in + base + (static_cast<long>(y) * W + x) * 3;
  uint8_t* o = out + base + (static_cast<long>(y) * W + x) * 3;
  const int r = px[0], g = px[1], bl = px[2];
  if (op <= AUG_POSTERIZE) {
    const uint8_t* l = lut + b * 768;
    o[0] = l[r];
    o[1] = l[256 + g];
    o[2] = l[512 + bl];
    return;
  }
  const float alpha = vals.v[op];
  const bool in_range = alpha >= 0.0f && alpha <= 1.0f;
  int d[3];
  if (op == AUG_COLOR) {
    d[0] = d[1] = d[2] = luma_u8(r, g, bl);
  } else if (op == AUG_CONTRAST) {
    d[0] = d[1] = d[2] = mean[b];
  } else if (op == AUG_BRIGHTNESS) {
    d[0] = d[1] = d[2] = 0;
  } else {  // AUG_SHARPNESS: degenerate = ImageFilter.SMOOTH (1 1 1 / 1 5 1 / 1 1 1) / 13, border pixels copied
    if (x == 0 || y == 0 || x == W - 1 || y == H - 1) {
      d[0] = r; d[1] = g; d[2] = bl;
    } else {
      const float k1 = static_cast<float>(1.0 / 13.0), k5 = static_cast<float>(5.0 / 13.0);
      const uint8_t* up = px - static_cast<long>(W) * 3;
      const uint8_t* dn = px + static_cast<long>(W) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // Pillow's ImagingFilter3x3: ss = offset + 0.5; ss += row(y+1); ss += row(y); ss += row(y-1), each row a*k0 + b*k1 + c*k2
        float ss = 0.5f;
        ss = __fadd_rn(ss, __fadd_rn(__fadd_rn(__fmul_rn(dn[c - 3], k1), __fmul_rn(dn[c], k1)), __fmul_rn(dn[c + 3], k1)));
        ss = __fadd_rn(ss, __fadd_rn(__fadd_rn(__fmul_rn(px[c - 3], k1), __fmul_rn(px[c], k5)), __fmul_rn(px[c + 3], k1)));
        ss = __fadd_rn(ss, __fadd_rn(__fadd_rn(__fmul_rn(up[c - 3], k1), __fmul_rn(up[c], k1)), __fmul_rn(up[c + 3], k1)));
        d[c] = ss <= 0.0f ? 0 : (ss >= 255.0f ? 255 : __float2int_rz(ss));
      }
    }
  }
  o[0] = blend_u8(d[0], r, alpha, in_range);
  o[1] = blend_u8(d[1], g, alpha, in_range);
  o[2] = blend_u8(d[2], bl, alpha, in_range);
}

// ToTensor (/255) -> Normalize((x - mean) / std) -> torch.flip(dims=[2]) (width), HWC uint8 -> CHW float
__global__ void __launch_bounds__(256) aug_finish_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int B, int H, int W) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const long hw = static_cast<long>(H) * W;
  if (i >= B * hw) return;
  const long b = i / hw, p = i % hw;
  const int y = static_cast<int>(p / W), x = static_cast<int>(p % W);
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = __fdiv_rn(static_cast<float>(in[i * 3 + c]), 255.0f);
    out[(b * 3 + c) * hw + static_cast<long>(y) * W + (W - 1 - x)] = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
  }
}

static size_t aug_align(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace dupl

using namespace dupl;

extern "C" int dupl_randaug_workspace_bytes(int32_t B, int32_t H, int32_t W, size_t* bytes) {
  DUPL_CHECK_ARG(B > 0 && H > 2 && W > 2 && bytes != nullptr, "dupl_randaug_workspace_bytes: bad arguments");
  const size_t img = aug_align(static_cast<size_t>(B) * H * W * 3);
  *bytes = 2 * img + aug_align(static_cast<size_t>(B) * 768 * 4) + aug_align(static_cast<size_t>(B) * 8) +
           aug_align(static_cast<size_t>(B) * 768) + aug_align(static_cast<size_t>(B) * 4);
  return DUPL_OK;
}

extern "C" int dupl_randaug(const float* images, float* out, int32_t B, int32_t H, int32_t W, const int32_t* ops_dev, int32_t n_ops,
                            const float* magnitudes7, void* workspace, size_t workspace_bytes, void* stream) {
  DUPL_CHECK_ARG(images && out && ops_dev && magnitudes7 && workspace, "dupl_randaug: NULL pointer");
  DUPL_CHECK_ARG(B > 0 && H > 2 && W > 2 && n_ops >= 0 && n_ops <= 64, "dupl_randaug: bad shape");
  size_t need = 0;
  dupl_randaug_workspace_bytes(B, H, W, &need);
  DUPL_CHECK_ARG(workspace_bytes >= need, "dupl_randaug: workspace too small (%zu < %zu)", workspace_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(workspace);
  AugWs ws;
  const size_t img = aug_align(static_cast<size_t>(B) * H * W * 3);
  ws.img[0] = reinterpret_cast<uint8_t*>(w);
  ws.img[1] = reinterpret_cast<uint8_t*>(w + img);
  w += 2 * img;
  ws.hist = reinterpret_cast<unsigned int*>(w);
  w += aug_align(static_cast<size_t>(B) * 768 * 4);
  ws.lsum = reinterpret_cast<unsigned long long*>(w);
  w += aug_align(static_cast<size_t>(B) * 8);
  ws.lut = reinterpret_cast<uint8_t*>(w);
  w += aug_align(static_cast<size_t>(B) * 768);
  ws.mean = reinterpret_cast<int*>(w);
  AugVals vals;
  for (int i = 0; i < AUG_NOPS; ++i) vals.v[i] = magnitudes7[i];
  const long hw = static_cast<long>(H) * W;
  const int blocks = static_cast<int>((static_cast<long>(B) * hw + 255) / 256);
  aug_from_float_kernel<<<blocks, 256, 0, st>>>(images, ws.img[0], B, H, W);
  DUPL_LAUNCH_OK();
  int cur = 0;
  for (int s = 0; s < n_ops; ++s) {
    const int32_t* ops = ops_dev + static_cast<long>(s) * B;
    DUPL_CUDA_OK(cudaMemsetAsync(ws.hist, 0, static_cast<size_t>(B) * 768 * 4, st));
    DUPL_CUDA_OK(cudaMemsetAsync(ws.lsum, 0, static_cast<size_t>(B) * 8, st));
    aug_stats_kernel<<<dim3(64, B), 256, 0, st>>>(ws.img[cur], ws.hist, ws.lsum, hw);
    DUPL_LAUNCH_OK();
    aug_params_kernel<<<B, 256, 0, st>>>(ops, ws.hist, ws.lsum, ws.lut, ws.mean, hw, vals);
    DUPL_LAUNCH_OK();
    aug_apply_kernel<<<dim3(cdiv(W, 32), cdiv(H, 8), B), 256, 0, st>>>(ws.img[cur], ws.img[cur ^ 1], ops, ws.lut, ws.mean, H, W, vals);
    DUPL_LAUNCH_OK();
    cur ^= 1;
  }
  aug_finish_kernel<<<blocks, 256, 0, st>>>(ws.img[cur], out, B, H, W);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}
