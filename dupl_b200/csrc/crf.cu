// DenseCRF mean-field inference on the GPU (reference utils/dcrf.py:42-69 ->
// pydensecrf, CPU).  Restates the published algorithm (Krähenbühl & Koltun 2011; permutohedral
// lattice of Adams et al. 2010) exactly as oracle/densecrf_ref.c does — parity is pinned to that C
// restatement, NOT to pydensecrf (absent): "parity unpinned" (DESIGN.md §5).
//
// Per pairwise kernel (Gaussian d=2, bilateral d=5) the lattice is built ONCE per image:
//   points   : elevate / round / rank / barycentric per pixel, (d+1) vertex keys packed into 64 bits
//              and inserted in an open-addressing table (atomicCAS);
//   vertices : occupied slots -> dense ids by an exclusive scan (no host round trip);
//   CSR      : entries (pixel, remainder) counting-sorted by vertex, so the splat is a GATHER;
//   blur     : the two neighbours of every vertex along each of the d+1 axes, by table lookup;
//   norm     : 1/sqrt(K 1 + 1e-20) (symmetric normalisation).
// Each mean-field iteration then runs, per kernel, splat -> (d+1) blurs -> slice, all coalesced over
// the class dimension (Q is kept pixel-major [N][C]), and the slice of the last kernel is fused with the
// exp-normalise.  The splat accumulates in 64-bit fixed point (2^-32), which makes the result
// independent of summation order: bit-reproducible run to run although the build uses atomics.
#include <math.h>

#include <stdlib.h>

#include "common.cuh"

namespace dupl {

constexpr unsigned long long CRF_EMPTY = ~0ull;
constexpr int CRF_CHUNK = 128;  // csr positions walked by one warp in the splat
constexpr int CRF_BATCH = 8;    // entries whose Q rows are in flight together
constexpr float CRF_FIX = 4294967296.0f;
constexpr float CRF_UNFIX = 1.0f / 4294967296.0f;

struct CrfWs {  // device pointers of one pairwise kernel
  int d, E, T;
  unsigned long long* keys;
  int* slot_id;
  int* ent_vertex;
  float* ent_w;
  unsigned long long* vkey;
  int* nb;
  int* row_ptr;
  int* cursor;
  int* csr;
  float* norm;
  float* val_a;
  float* val_b;
  long long* acc1;
};

struct CrfLayout {
  CrfWs k[2];
  int* meta;        // [4]: M_gauss, M_bilateral, key overflow flag, unused
  int* scan_tmp;    // block sums
  size_t bytes;
};

static size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

static CrfLayout crf_layout(void* base, int W, int H) {
  CrfLayout L;
  const size_t N = static_cast<size_t>(W) * H;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += align_up(bytes);
    return p;
  };
  L.meta = static_cast<int*>(take(4 * sizeof(int)));
  size_t maxT = 0;
  for (int i = 0; i < 2; ++i) {
    CrfWs& w = L.k[i];
    w.d = i == 0 ? 2 : 5;
    w.E = static_cast<int>(N * (w.d + 1));
    w.T = 1;
    while (w.T < 2 * w.E) w.T <<= 1;
    if (static_cast<size_t>(w.T) > maxT) maxT = w.T;
    w.keys = static_cast<unsigned long long*>(take(sizeof(unsigned long long) * w.T));
    w.slot_id = static_cast<int*>(take(sizeof(int) * w.T));
    w.ent_vertex = static_cast<int*>(take(sizeof(int) * w.E));
    w.ent_w = static_cast<float*>(take(sizeof(float) * w.E));
    w.vkey = static_cast<unsigned long long*>(take(sizeof(unsigned long long) * w.E));
    w.nb = static_cast<int*>(take(sizeof(int) * 2 * (w.d + 1) * static_cast<size_t>(w.E)));
    w.row_ptr = static_cast<int*>(take(sizeof(int) * (static_cast<size_t>(w.E) + 1)));
    w.cursor = static_cast<int*>(take(sizeof(int) * w.E));
    w.csr = static_cast<int*>(take(sizeof(int) * w.E));
    w.norm = static_cast<float*>(take(sizeof(float) * N));
    w.val_a = static_cast<float*>(take(sizeof(float) * w.E));
    w.val_b = static_cast<float*>(take(sizeof(float) * w.E));
    w.acc1 = static_cast<long long*>(take(sizeof(long long) * w.E));
  }
  L.scan_tmp = static_cast<int*>(take(sizeof(int) * (maxT / 1024 + 2)));
  L.bytes = off;
  return L;
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of int32 (in place), 1024 elements per block
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

__global__ void __launch_bounds__(256) scan_blocks_kernel(int* data, int n, int* block_sums) {
  __shared__ int warp_tot[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long base = static_cast<long>(blockIdx.x) * 1024 + threadIdx.x * 4;
  int v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = base + i < n ? data[base + i] : 0;
  const int tsum = v[0] + v[1] + v[2] + v[3];
  const int incl = warp_incl_scan(tsum, lane);
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < wid; ++w) woff += warp_tot[w];
  int run = woff + incl - tsum;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (base + i < n) data[base + i] = run;
    run += v[i];
  }
  if (threadIdx.x == 255) block_sums[blockIdx.x] = run;
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(int* block_sums, int nb, int* total) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? block_sums[i] : 0;
    const int incl = warp_incl_scan(v, lane);
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int woff = carry_s;
    for (int w = 0; w < wid; ++w) woff += warp_tot[w];
    if (i < nb) block_sums[i] = woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = woff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total != nullptr) *total = carry_s;
}

__global__ void __launch_bounds__(256) scan_add_kernel(int* data, int n, const int* block_sums) {
  const long base = static_cast<long>(blockIdx.x) * 1024 + threadIdx.x * 4;
  const int add = block_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (base + i < n) data[base + i] += add;
}

static int exclusive_scan(int* data, int n, int* total, int* tmp, cudaStream_t st) {
  const int nb = cdiv(n, 1024);
  scan_blocks_kernel<<<nb, 256, 0, st>>>(data, n, tmp);
  DUPL_LAUNCH_OK();
  scan_sums_kernel<<<1, 1024, 0, st>>>(tmp, nb, total);
  DUPL_LAUNCH_OK();
  scan_add_kernel<<<nb, 256, 0, st>>>(data, n, tmp);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

// ------------------------------------------------------------------------------------------------
// lattice construction
// ------------------------------------------------------------------------------------------------
template <int D>
struct KeyPack {
  static constexpr int BITS = D <= 3 ? 16 : 12;
  static constexpr int BIAS = 1 << (BITS - 1);
  __device__ static bool pack(const int* c, unsigned long long& key) {
    key = 0;
    bool ok = true;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const int b = c[i] + BIAS;
      ok = ok && b >= 0 && b < (1 << BITS);
      key |= static_cast<unsigned long long>(b & ((1 << BITS) - 1)) << (i * BITS);
    }
    return ok;
  }
  __device__ static void unpack(unsigned long long key, int* c) {
#pragma unroll
    for (int i = 0; i < D; ++i) c[i] = static_cast<int>((key >> (i * BITS)) & ((1 << BITS) - 1)) - BIAS;
  }
};

__device__ __forceinline__ unsigned int mix64(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return static_cast<unsigned int>(k);
}

template <int D>
__global__ void __launch_bounds__(256) crf_points_kernel(const unsigned char* __restrict__ img, int W, int H, float sxy,
                                                         float srgb, CrfWs ws, int* __restrict__ meta) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= W * H) return;
  const int x = pix % W, y = pix / W;
  float f[D];
  f[0] = static_cast<float>(x) / sxy;
  f[1] = static_cast<float>(y) / sxy;
  if constexpr (D == 5) {
#pragma unroll
    for (int c = 0; c < 3; ++c) f[2 + c] = static_cast<float>(img[static_cast<long>(pix) * 3 + c]) / srgb;
  }
  // elevate onto the hyperplane sum = 0 (scale_factor[i] = (d+1) sqrt(2/3) / sqrt((i+1)(i+2)))
  float elevated[D + 1];
  const float inv_std_dev = sqrtf(2.0f / 3.0f) * static_cast<float>(D + 1);
  float sm = 0.0f;
#pragma unroll
  for (int j = D; j > 0; --j) {
    const float cf = f[j - 1] * (1.0f / sqrtf(static_cast<float>((j + 1) * j)) * inv_std_dev);
    elevated[j] = sm - static_cast<float>(j) * cf;
    sm += cf;
  }
  elevated[0] = sm;
  // closest remainder-0 point
  const float down_factor = 1.0f / static_cast<float>(D + 1), up_factor = static_cast<float>(D + 1);
  float rem0[D + 1];
  int rank[D + 1];
  int sum = 0;
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    const float v = down_factor * elevated[i];
    const float up = ceilf(v) * up_factor, down = floorf(v) * up_factor;
    const int rd2 = (up - elevated[i] < elevated[i] - down) ? static_cast<int>(up) : static_cast<int>(down);
    rem0[i] = static_cast<float>(rd2);
    sum += static_cast<int>(static_cast<float>(rd2) * down_factor);
    rank[i] = 0;
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const float di = elevated[i] - rem0[i];
#pragma unroll
    for (int j = i + 1; j <= D; ++j) {
      if (di < elevated[j] - rem0[j]) rank[i]++;
      else rank[j]++;
    }
  }
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    rank[i] += sum;
    if (rank[i] < 0) {
      rank[i] += D + 1;
      rem0[i] += static_cast<float>(D + 1);
    } else if (rank[i] > D) {
      rank[i] -= D + 1;
      rem0[i] -= static_cast<float>(D + 1);
    }
  }
  float bary[D + 2];
#pragma unroll
  for (int i = 0; i <= D + 1; ++i) bary[i] = 0.0f;
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    const float v = (elevated[i] - rem0[i]) * down_factor;
#pragma unroll
    for (int s = 0; s <= D + 1; ++s) {  // static indexing keeps bary[] in registers
      if (s == D - rank[i]) bary[s] += v;
      if (s == D - rank[i] + 1) bary[s] -= v;
    }
  }
  bary[0] += 1.0f + bary[D + 1];

#pragma unroll
  for (int r = 0; r <= D; ++r) {
    int key_c[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      // canonical[r][rank] = r if rank <= D - r else r - (D+1)
      const int canon = rank[i] <= D - r ? r : r - (D + 1);
      key_c[i] = static_cast<int>(rem0[i]) + canon;
    }
    unsigned long long key;
    if (!KeyPack<D>::pack(key_c, key)) atomicExch(&meta[2], 1);
    unsigned int slot = mix64(key) & static_cast<unsigned int>(ws.T - 1);
    for (;;) {
      const unsigned long long prev = atomicCAS(&ws.keys[slot], CRF_EMPTY, key);
      if (prev == CRF_EMPTY || prev == key) break;
      slot = (slot + 1) & static_cast<unsigned int>(ws.T - 1);
    }
    const long e = static_cast<long>(pix) * (D + 1) + r;
    ws.ent_vertex[e] = static_cast<int>(slot);
    ws.ent_w[e] = bary[r];
  }
}

__global__ void crf_flags_kernel(const unsigned long long* __restrict__ keys, int* __restrict__ flags, int T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < T) flags[i] = keys[i] != CRF_EMPTY ? 1 : 0;
}

__global__ void crf_vertices_kernel(CrfWs ws) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ws.T) return;
  const unsigned long long k = ws.keys[i];
  if (k != CRF_EMPTY) ws.vkey[ws.slot_id[i]] = k;
}

__global__ void crf_entries_kernel(CrfWs ws) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ws.E) return;
  const int v = ws.slot_id[ws.ent_vertex[e]];
  ws.ent_vertex[e] = v;
  atomicAdd(&ws.row_ptr[v], 1);
}

__global__ void crf_fill_kernel(CrfWs ws) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ws.E) return;
  const int v = ws.ent_vertex[e];
  const int pos = ws.row_ptr[v] + atomicAdd(&ws.cursor[v], 1);
  ws.csr[pos] = e;
}

template <int D>
__device__ __forceinline__ int crf_lookup(const CrfWs& ws, const int* c) {
  unsigned long long key;
  if (!KeyPack<D>::pack(c, key)) return -1;
  unsigned int slot = mix64(key) & static_cast<unsigned int>(ws.T - 1);
  for (;;) {
    const unsigned long long k = ws.keys[slot];
    if (k == CRF_EMPTY) return -1;
    if (k == key) return ws.slot_id[slot];
    slot = (slot + 1) & static_cast<unsigned int>(ws.T - 1);
  }
}

template <int D>
__global__ void __launch_bounds__(256) crf_neighbors_kernel(CrfWs ws, const int* __restrict__ M_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int M = *M_dev;
  if (i >= M) return;
  int c[D];
  KeyPack<D>::unpack(ws.vkey[i], c);
#pragma unroll
  for (int j = 0; j <= D; ++j) {
    int a[D], b[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      a[k] = c[k] - 1;
      b[k] = c[k] + 1;
    }
    if (j < D) {  // the (d+1)-th coordinate is implicit
#pragma unroll
      for (int k = 0; k < D; ++k)
        if (k == j) {
          a[k] = c[k] + D;
          b[k] = c[k] - D;
        }
    }
    const long o = (static_cast<long>(j) * ws.E + i) * 2;
    ws.nb[o] = crf_lookup<D>(ws, a);
    ws.nb[o + 1] = crf_lookup<D>(ws, b);
  }
}

// ------------------------------------------------------------------------------------------------
// filter stages.  values: [M][C] fp32; acc: [M][C] 64-bit fixed point
// ------------------------------------------------------------------------------------------------
// Gather-splat: each warp walks CRF_CHUNK consecutive csr positions (sorted by vertex), lanes = classes.
template <int CPL>
__global__ void __launch_bounds__(256) crf_splat_kernel(CrfWs ws, const float* __restrict__ Q, int C, int use_norm,
                                                        long long* __restrict__ acc, int k0, int ldc) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long start = static_cast<long>(warp) * CRF_CHUNK;
  if (start >= ws.E) return;
  const long end = min(static_cast<long>(ws.E), start + CRF_CHUNK);
  const int d1 = ws.d + 1;
  long long a[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) a[c] = 0;
  int cur = -1;
  for (long base = start; base < end; base += 32) {
    int v = -1, pix = 0;
    float wgt = 0.0f;
    if (base + lane < end) {
      const int e = ws.csr[base + lane];
      v = ws.ent_vertex[e];
      pix = e / d1;
      wgt = ws.ent_w[e] * (use_norm ? ws.norm[pix] : 1.0f);
    }
    const int cnt = static_cast<int>(min(32L, end - base));
    // The Q rows of CRF_BATCH entries are requested before the first one is consumed: walked one entry at a time, every
    // entry paid a full L2 round trip (~0.6 us per entry and warp: the kernel was latency-bound at 150 us for 0.9 M
    // entries).  The sums are 64-bit fixed point, so the grouping does not change a bit of the result.
    for (int i0 = 0; i0 < cnt; i0 += CRF_BATCH) {
      int vb[CRF_BATCH];
      float wb[CRF_BATCH], qb[CRF_BATCH][CPL];
#pragma unroll
      for (int j = 0; j < CRF_BATCH; ++j) {
        const int i = (i0 + j) & 31;
        vb[j] = __shfl_sync(0xffffffffu, v, i);
        const int pi = __shfl_sync(0xffffffffu, pix, i);
        wb[j] = __shfl_sync(0xffffffffu, wgt, i);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const int k = k0 + lane + 32 * c;
          qb[j][c] = (Q != nullptr && k < C && i0 + j < cnt) ? __ldg(Q + static_cast<long>(pi) * ldc + k) : 1.0f;
        }
      }
#pragma unroll
      for (int j = 0; j < CRF_BATCH; ++j) {
        if (i0 + j < cnt) {
          if (vb[j] != cur) {
            if (cur >= 0) {
#pragma unroll
              for (int c = 0; c < CPL; ++c) {
                const int k = k0 + lane + 32 * c;
                if (k < C && a[c] != 0)
                  atomicAdd(reinterpret_cast<unsigned long long*>(acc + static_cast<long>(cur) * ldc + k),
                            static_cast<unsigned long long>(a[c]));
                a[c] = 0;
              }
            }
            cur = vb[j];
          }
#pragma unroll
          for (int c = 0; c < CPL; ++c) {
            const int k = k0 + lane + 32 * c;
            if (k < C) a[c] += __float2ll_rn(wb[j] * qb[j][c] * CRF_FIX);
          }
        }
      }
    }
  }
  if (cur >= 0) {
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int k = k0 + lane + 32 * c;
      if (k < C && a[c] != 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(acc + static_cast<long>(cur) * ldc + k),
                  static_cast<unsigned long long>(a[c]));
    }
  }
}

// out = v + 0.5 (v[n1] + v[n2]) along axis j.  FIRST: input is the fixed-point splat accumulator.
template <bool FIRST>
__global__ void __launch_bounds__(256) crf_blur_kernel(CrfWs ws, int j, int C, const long long* __restrict__ acc,
                                                       const float* __restrict__ in, float* __restrict__ out,
                                                       const int* __restrict__ M_dev, int M_host) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const int M = M_dev != nullptr ? *M_dev : M_host;
  if (idx >= static_cast<long>(M) * C) return;
  const int i = static_cast<int>(idx / C), k = static_cast<int>(idx % C);
  const long o = (static_cast<long>(j) * ws.E + i) * 2;
  const int n1 = ws.nb[o], n2 = ws.nb[o + 1];
  auto val = [&](int v) -> float {
    if (v < 0) return 0.0f;
    const long p = static_cast<long>(v) * C + k;
    return FIRST ? static_cast<float>(acc[p]) * CRF_UNFIX : in[p];
  };
  out[idx] = val(i) + 0.5f * (val(n1) + val(n2));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_add(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Slice (+ symmetric normalisation + Potts weight) for one kernel; warp per pixel, lanes = classes.
// MODE 0: msg = s       MODE 1: msg += s       MODE 2: Q = softmax(-U + msg_in + s)  (msg_in optional)
// MODE 3: norm[pix] = 1/sqrt(s + 1e-20) (build time, C == 1)
template <int CPL, int MODE>
__global__ void __launch_bounds__(256) crf_slice_kernel(CrfWs ws, const float* __restrict__ values, int C, int N,
                                                        float weight, float* __restrict__ msg, const float* __restrict__ U,
                                                        float* __restrict__ Q, int have_msg) {
  const int lane = threadIdx.x & 31;
  const int pix = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= N) return;
  const int d1 = ws.d + 1;
  const float alpha = 1.0f / (1.0f + exp2f(-static_cast<float>(ws.d)));
  float s[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) s[c] = 0.0f;
  for (int r = 0; r < d1; ++r) {
    const long e = static_cast<long>(pix) * d1 + r;
    const int v = ws.ent_vertex[e];
    const float w = ws.ent_w[e];
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int k = lane + 32 * c;
      if (k < C) s[c] += w * values[static_cast<long>(v) * C + k] * alpha;
    }
  }
  if (MODE == 3) {
    if (lane == 0) ws.norm[pix] = 1.0f / sqrtf(s[0] + 1e-20f);
    return;
  }
  const float nrm = ws.norm[pix] * weight;
  float x[CPL];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int k = lane + 32 * c;
    x[c] = -INFINITY;
    if (k < C) {
      const long p = static_cast<long>(pix) * C + k;
      const float m = s[c] * nrm;
      if (MODE == 0) msg[p] = m;
      if (MODE == 1) msg[p] += m;
      if (MODE == 2) {
        x[c] = -U[p] + (have_msg ? msg[p] : 0.0f) + m;
        mx = fmaxf(mx, x[c]);
      }
    }
  }
  if (MODE == 2) {
    mx = warp_max(mx);
    float sum = 0.0f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int k = lane + 32 * c;
      x[c] = k < C ? expf(x[c] - mx) : 0.0f;
      sum += x[c];
    }
    sum = warp_add(sum);
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int k = lane + 32 * c;
      if (k < C) Q[static_cast<long>(pix) * C + k] = x[c] / sum;
    }
  }
}

// U[pix][k] = energy, Q = softmax(-U).  probs/unary given class-major [C][N].
template <int CPL>
__global__ void __launch_bounds__(256) crf_init_kernel(const float* __restrict__ in, int is_energy, int C, int N,
                                                       float* __restrict__ U, float* __restrict__ Q) {
  const int lane = threadIdx.x & 31;
  const int pix = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= N) return;
  float x[CPL];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int k = lane + 32 * c;
    x[c] = -INFINITY;
    if (k < C) {
      const float v = in[static_cast<long>(k) * N + pix];
      const float u = is_energy ? v : -logf(fminf(fmaxf(v, 1e-5f), 1.0f));  // unary_from_softmax
      U[static_cast<long>(pix) * C + k] = u;
      x[c] = -u;
      mx = fmaxf(mx, x[c]);
    }
  }
  mx = warp_max(mx);
  float sum = 0.0f;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int k = lane + 32 * c;
    x[c] = k < C ? expf(x[c] - mx) : 0.0f;
    sum += x[c];
  }
  sum = warp_add(sum);
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int k = lane + 32 * c;
    if (k < C) Q[static_cast<long>(pix) * C + k] = x[c] / sum;
  }
}

// [N][C] -> [C][N]
__global__ void __launch_bounds__(256) crf_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int N,
                                                            int C) {
  __shared__ float tile[32][33];
  const int p0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int p = p0 + r, k = k0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < N && k < C) ? in[static_cast<long>(p) * C + k] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int k = k0 + r, p = p0 + threadIdx.x;
    if (k < C && p < N) out[static_cast<long>(k) * N + p] = tile[threadIdx.x][r];
  }
}

template <int D>
static int build_one(const CrfWs& ws, const unsigned char* img, int W, int H, float sxy, float srgb, int* meta, int which,
                     int* scan_tmp, cudaStream_t st) {
  const int N = W * H;
  DUPL_CUDA_OK(cudaMemsetAsync(ws.keys, 0xFF, sizeof(unsigned long long) * ws.T, st));
  DUPL_CUDA_OK(cudaMemsetAsync(ws.row_ptr, 0, sizeof(int) * (static_cast<size_t>(ws.E) + 1), st));
  DUPL_CUDA_OK(cudaMemsetAsync(ws.cursor, 0, sizeof(int) * ws.E, st));
  crf_points_kernel<D><<<cdiv(N, 256), 256, 0, st>>>(img, W, H, sxy, srgb, ws, meta);
  DUPL_LAUNCH_OK();
  crf_flags_kernel<<<cdiv(ws.T, 256), 256, 0, st>>>(ws.keys, ws.slot_id, ws.T);
  DUPL_LAUNCH_OK();
  int rc = exclusive_scan(ws.slot_id, ws.T, meta + which, scan_tmp, st);
  if (rc) return rc;
  crf_vertices_kernel<<<cdiv(ws.T, 256), 256, 0, st>>>(ws);
  DUPL_LAUNCH_OK();
  crf_entries_kernel<<<cdiv(ws.E, 256), 256, 0, st>>>(ws);
  DUPL_LAUNCH_OK();
  rc = exclusive_scan(ws.row_ptr, ws.E + 1, nullptr, scan_tmp, st);
  if (rc) return rc;
  crf_fill_kernel<<<cdiv(ws.E, 256), 256, 0, st>>>(ws);
  DUPL_LAUNCH_OK();
  crf_neighbors_kernel<D><<<cdiv(ws.E, 256), 256, 0, st>>>(ws, meta + which);
  DUPL_LAUNCH_OK();
  // norm = 1/sqrt(K 1 + 1e-20): the filter applied to a vector of ones (C = 1), M read on the device
  DUPL_CUDA_OK(cudaMemsetAsync(ws.acc1, 0, sizeof(long long) * ws.E, st));
  const int warps = cdiv(ws.E, CRF_CHUNK);
  crf_splat_kernel<1><<<cdiv(warps, 8), 256, 0, st>>>(ws, nullptr, 1, 0, ws.acc1, 0, 1);
  DUPL_LAUNCH_OK();
  float* a = ws.val_a;
  float* b = ws.val_b;
  for (int j = 0; j <= D; ++j) {
    if (j == 0) crf_blur_kernel<true><<<cdiv(ws.E, 256), 256, 0, st>>>(ws, j, 1, ws.acc1, nullptr, a, meta + which, 0);
    else crf_blur_kernel<false><<<cdiv(ws.E, 256), 256, 0, st>>>(ws, j, 1, nullptr, a, b, meta + which, 0);
    DUPL_LAUNCH_OK();
    if (j > 0) {
      float* t = a; a = b; b = t;
    }
  }
  crf_slice_kernel<1, 3><<<cdiv(N, 8), 256, 0, st>>>(ws, a, 1, N, 1.0f, nullptr, nullptr, nullptr, 0);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

template <int CPL>
static int infer_impl(const dupl_crf_args* a, const CrfLayout& L, const int* M, cudaStream_t st) {
  const int N = a->W * a->H, C = a->C;
  const size_t nc = static_cast<size_t>(N) * C;
  float* U = static_cast<float*>(a->values);
  float* Q = U + nc;
  float* msg = Q + nc;
  size_t off = align_up(3 * nc * sizeof(float));
  const bool use[2] = {a->pos_w != 0.0f, a->bi_w != 0.0f};
  const float weight[2] = {a->pos_w, a->bi_w};
  long long* acc[2];
  float* va[2];
  float* vb[2];
  for (int k = 0; k < 2; ++k) {
    const size_t mc = static_cast<size_t>(M[k]) * C;
    acc[k] = reinterpret_cast<long long*>(static_cast<char*>(a->values) + off);
    off += align_up(mc * sizeof(long long));
    va[k] = reinterpret_cast<float*>(static_cast<char*>(a->values) + off);
    off += align_up(mc * sizeof(float));
    vb[k] = reinterpret_cast<float*>(static_cast<char*>(a->values) + off);
    off += align_up(mc * sizeof(float));
  }
  DUPL_CHECK_ARG(off <= a->values_bytes, "dupl_crf_infer: values buffer too small (%zu < %zu)", a->values_bytes, off);
  crf_init_kernel<CPL><<<cdiv(N, 8), 256, 0, st>>>(a->unary_or_probs, a->input_is_energy, C, N, U, Q);
  DUPL_LAUNCH_OK();
  const int last = use[1] ? 1 : 0;
  for (int it = 0; it < a->iters; ++it) {
    bool have_msg = false;
    for (int k = 0; k < 2; ++k) {
      if (!use[k]) continue;
      const CrfWs& ws = L.k[k];
      const size_t mc = static_cast<size_t>(M[k]) * C;
      DUPL_CUDA_OK(cudaMemsetAsync(acc[k], 0, mc * sizeof(long long), st));
      // The gather reads every pixel's Q row once per lattice vertex it touches (d+1 times, in vertex order = random
      // in pixel order).  With C > 32 the [N][C] matrix (100 MB at 640x480x81) does not stay in L2 and every one of those
      // reads went to DRAM (ncu: 624 MB per launch against 153 MB algorithmic).  32 classes at a time, the 128-byte
      // slices of all rows (39 MB) are L2-resident across the d+1 visits.  (DUPL_CRF_SPLAT_WIDE=1: single wide pass.)
      static const bool wide = getenv("DUPL_CRF_SPLAT_WIDE") != nullptr;
      if (CPL > 1 && !wide) {
        for (int k0 = 0; k0 < C; k0 += 32) {
          crf_splat_kernel<1><<<cdiv(cdiv(ws.E, CRF_CHUNK), 8), 256, 0, st>>>(ws, Q, min(C, k0 + 32), 1, acc[k], k0, C);
          DUPL_LAUNCH_OK();
        }
      } else {
        crf_splat_kernel<CPL><<<cdiv(cdiv(ws.E, CRF_CHUNK), 8), 256, 0, st>>>(ws, Q, C, 1, acc[k], 0, C);
        DUPL_LAUNCH_OK();
      }
      float* x = va[k];
      float* y = vb[k];
      const int blocks = static_cast<int>((mc + 255) / 256);
      for (int j = 0; j <= ws.d; ++j) {
        if (j == 0) crf_blur_kernel<true><<<blocks, 256, 0, st>>>(ws, j, C, acc[k], nullptr, x, nullptr, M[k]);
        else crf_blur_kernel<false><<<blocks, 256, 0, st>>>(ws, j, C, nullptr, x, y, nullptr, M[k]);
        DUPL_LAUNCH_OK();
        if (j > 0) {
          float* t = x; x = y; y = t;
        }
      }
      if (k == last) crf_slice_kernel<CPL, 2><<<cdiv(N, 8), 256, 0, st>>>(ws, x, C, N, weight[k], msg, U, Q, have_msg ? 1 : 0);
      else crf_slice_kernel<CPL, 0><<<cdiv(N, 8), 256, 0, st>>>(ws, x, C, N, weight[k], msg, U, Q, 0);
      DUPL_LAUNCH_OK();
      have_msg = true;
    }
  }
  dim3 grid(cdiv(N, 32), cdiv(C, 32)), block(32, 8);
  crf_transpose_kernel<<<grid, block, 0, st>>>(Q, a->out, N, C);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

}  // namespace dupl

using namespace dupl;

extern "C" int dupl_crf_workspace_bytes(int32_t W, int32_t H, size_t* bytes) {
  DUPL_CHECK_ARG(W > 0 && H > 0 && bytes != nullptr, "dupl_crf_workspace_bytes: bad arguments");
  DUPL_CHECK_ARG(static_cast<long>(W) * H * 6 < (1L << 30), "dupl_crf_workspace_bytes: image too large");
  *bytes = crf_layout(nullptr, W, H).bytes;
  return DUPL_OK;
}

extern "C" int dupl_crf_values_bytes(int32_t W, int32_t H, int32_t C, int32_t M_gauss, int32_t M_bilateral, size_t* bytes) {
  DUPL_CHECK_ARG(W > 0 && H > 0 && C > 0 && M_gauss >= 0 && M_bilateral >= 0 && bytes != nullptr,
                 "dupl_crf_values_bytes: bad arguments");
  size_t off = align_up(3 * static_cast<size_t>(W) * H * C * sizeof(float));
  const int M[2] = {M_gauss, M_bilateral};
  for (int k = 0; k < 2; ++k) {
    const size_t mc = static_cast<size_t>(M[k]) * C;
    off += align_up(mc * sizeof(long long)) + 2 * align_up(mc * sizeof(float));
  }
  *bytes = off;
  return DUPL_OK;
}

extern "C" int dupl_crf_build(const dupl_crf_args* a, void* stream) {
  DUPL_CHECK_ARG(a != nullptr, "dupl_crf_build: args is NULL");
  DUPL_CHECK_ARG(a->W > 0 && a->H > 0 && a->image && a->workspace && a->meta, "dupl_crf_build: bad arguments");
  DUPL_CHECK_ARG((a->pos_w == 0.0f || a->pos_xy_std > 0.0f) && (a->bi_w == 0.0f || (a->bi_xy_std > 0.0f && a->bi_rgb_std > 0.0f)),
                 "dupl_crf_build: standard deviations must be positive");
  CrfLayout L = crf_layout(a->workspace, a->W, a->H);
  DUPL_CHECK_ARG(L.bytes <= a->workspace_bytes, "dupl_crf_build: workspace too small (%zu < %zu)", a->workspace_bytes, L.bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DUPL_CUDA_OK(cudaMemsetAsync(L.meta, 0, 4 * sizeof(int), st));
  int rc = DUPL_OK;
  if (a->pos_w != 0.0f) rc = build_one<2>(L.k[0], a->image, a->W, a->H, a->pos_xy_std, 1.0f, L.meta, 0, L.scan_tmp, st);
  if (rc) return rc;
  if (a->bi_w != 0.0f) rc = build_one<5>(L.k[1], a->image, a->W, a->H, a->bi_xy_std, a->bi_rgb_std, L.meta, 1, L.scan_tmp, st);
  if (rc) return rc;
  DUPL_CUDA_OK(cudaMemcpyAsync(a->meta, L.meta, 4 * sizeof(int), cudaMemcpyDeviceToDevice, st));
  return DUPL_OK;
}

extern "C" int dupl_crf_infer(const dupl_crf_args* a, int32_t M_gauss, int32_t M_bilateral, void* stream) {
  DUPL_CHECK_ARG(a != nullptr, "dupl_crf_infer: args is NULL");
  DUPL_CHECK_ARG(a->W > 0 && a->H > 0 && a->C > 0 && a->C <= 96 && a->iters >= 0, "dupl_crf_infer: bad shape (C <= 96)");
  DUPL_CHECK_ARG(a->unary_or_probs && a->out && a->workspace && a->values, "dupl_crf_infer: NULL pointer");
  DUPL_CHECK_ARG(a->pos_w != 0.0f || a->bi_w != 0.0f, "dupl_crf_infer: both pairwise weights are zero");
  CrfLayout L = crf_layout(a->workspace, a->W, a->H);
  const int M[2] = {M_gauss, M_bilateral};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->C <= 32) return infer_impl<1>(a, L, M, st);
  if (a->C <= 64) return infer_impl<2>(a, L, M, st);
  return infer_impl<3>(a, L, M, st);
}
