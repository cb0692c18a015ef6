// CAM post-processing and pseudo-label casts — HBM-bound, one pass over
// the output, low-resolution sources staged in shared memory.
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
__global__ void __launch_bounds__(MSCAM_THREADS) mscam_kernel(MscamParams p) {
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
  DUPL_CHECK_ARG(smem <= 200 * 1024, "dupl_mscam_post: low-res maps need %zu B of shared memory", smem);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static size_t smem_set0 = 0, smem_set1 = 0;
  if (smem > 48 * 1024) {
    if (smem > smem_set0) {
      DUPL_CUDA_OK(cudaFuncSetAttribute(mscam_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      smem_set0 = smem;
    }
    if (smem > smem_set1) {
      DUPL_CUDA_OK(cudaFuncSetAttribute(mscam_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      smem_set1 = smem;
    }
  }
  const int planes = a->b * a->K;
  mscam_init_minmax<<<cdiv(planes, 256), 256, 0, st>>>(p.minmax, planes);
  DUPL_LAUNCH_OK();
  dim3 grid(planes, cdiv(a->H, MSCAM_ROWS));
  mscam_kernel<0><<<grid, MSCAM_THREADS, smem, st>>>(p);
  DUPL_LAUNCH_OK();
  mscam_kernel<1><<<grid, MSCAM_THREADS, smem, st>>>(p);
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
