// Pixel-adaptive refinement (model/PAR.py:64-91) and the refine_cams_with_* wrappers around it
// (utils/cam_helper.py:338-440).  The reference gathers neighbours with 66 one-hot dilated
// convolutions per call and materialises [1,c,48,h,w] tensors; here neighbours are gathered
// directly (replicate border = index clamp), the 48-way affinity is computed once per image and
// shared by every mask set that is propagated over that image, and all channel planes of all
// images advance together in one launch per iteration.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "resample.cuh"

namespace dupl {

constexpr int PAR_MAX_N = 8 * DUPL_PAR_MAX_DIL;

struct ParGeom {
  int dil[DUPL_PAR_MAX_DIL];
  int ndil;
};

// offsets in the order of PAR.get_kernel (PAR.py:10-24): (-1,-1) (-1,0) (-1,1) (0,-1) (0,1) (1,-1) (1,0) (1,1)
__device__ __constant__ int c_par_dy[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
__device__ __constant__ int c_par_dx[8] = {-1, 0, 1, -1, 1, -1, 0, 1};

struct ParPos {
  float v[PAR_MAX_N];
};

// ---------------------------------------------------------------------------------------------
// Affinity: one thread per pixel, the 8*ndil neighbour values of one channel live in registers.
// ---------------------------------------------------------------------------------------------
template <int NDIL>
__global__ void __launch_bounds__(256) par_affinity_kernel(const float* __restrict__ imgs, float* __restrict__ aff,
                                                           int C, int h, int w, ParGeom g, ParPos pos, float w1,
                                                           float w2) {
  constexpr int N = 8 * NDIL;
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  const int b = blockIdx.z;
  if (x >= w || y >= h) return;
  const long hw = static_cast<long>(h) * w;
  float a[N];
#pragma unroll
  for (int n = 0; n < N; ++n) a[n] = 0.0f;
  for (int c = 0; c < C; ++c) {
    const float* plane = imgs + (static_cast<long>(b) * C + c) * hw;
    const float ctr = __ldg(plane + static_cast<long>(y) * w + x);
    float v[N];
    float sum = 0.0f;
#pragma unroll
    for (int di = 0; di < NDIL; ++di) {
      const int d = g.dil[di];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int yy = min(max(y + c_par_dy[j] * d, 0), h - 1);
        const int xx = min(max(x + c_par_dx[j] * d, 0), w - 1);
        v[di * 8 + j] = __ldg(plane + static_cast<long>(yy) * w + xx);
        sum += v[di * 8 + j];
      }
    }
    const float mean = sum / N;
    float ss = 0.0f;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const float dlt = v[n] - mean;
      ss = fmaf(dlt, dlt, ss);
    }
    const float sd = sqrtf(ss / (N - 1)) + 1e-8f;  // torch.std is unbiased
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const float t = __fdiv_rn(__fdiv_rn(fabsf(v[n] - ctr), sd), w1);
      a[n] -= t * t;
    }
  }
  float mx = -INFINITY;
#pragma unroll
  for (int n = 0; n < N; ++n) {
    a[n] = a[n] / C;  // mean over channels
    mx = fmaxf(mx, a[n]);
  }
  float s = 0.0f;
#pragma unroll
  for (int n = 0; n < N; ++n) {
    a[n] = expf(a[n] - mx);
    s += a[n];
  }
  float* o = aff + static_cast<long>(b) * N * hw + static_cast<long>(y) * w + x;
#pragma unroll
  for (int n = 0; n < N; ++n) o[n * hw] = __fdiv_rn(a[n], s) + w2 * pos.v[n];
}

// ---------------------------------------------------------------------------------------------
// One propagation step for up to CH planes of one image per thread block column (blockIdx.z).
// nactive (device, optional): number of live planes of each image; chunks past it exit.
// ---------------------------------------------------------------------------------------------
template <int NDIL, int CH>
__global__ void __launch_bounds__(256) par_propagate_kernel(const float* __restrict__ aff, const float* __restrict__ src,
                                                            float* __restrict__ dst, const int* __restrict__ nactive,
                                                            int P, int chunks, int h, int w, ParGeom g) {
  constexpr int N = 8 * NDIL;
  const int b = blockIdx.z / chunks;
  const int chunk = blockIdx.z % chunks;
  const int live = nactive != nullptr ? min(__ldg(nactive + b), P) : P;
  const int p0 = chunk * CH;
  if (p0 >= live) return;
  const int np = min(CH, live - p0);
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= w || y >= h) return;
  const long hw = static_cast<long>(h) * w;
  const float* A = aff + static_cast<long>(b) * N * hw + static_cast<long>(y) * w + x;
  const float* S = src + (static_cast<long>(b) * P + p0) * hw;
  float acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = 0.0f;
#pragma unroll
  for (int di = 0; di < NDIL; ++di) {
    const int d = g.dil[di];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int yy = min(max(y + c_par_dy[j] * d, 0), h - 1);
      const int xx = min(max(x + c_par_dx[j] * d, 0), w - 1);
      const float a = __ldg(A + (di * 8 + j) * hw);
      const float* q = S + static_cast<long>(yy) * w + xx;
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (c < np) acc[c] = fmaf(a, __ldg(q + c * hw), acc[c]);
    }
  }
  float* D = dst + (static_cast<long>(b) * P + p0) * hw + static_cast<long>(y) * w + x;
#pragma unroll
  for (int c = 0; c < CH; ++c)
    if (c < np) D[c * hw] = acc[c];
}

// ---------------------------------------------------------------------------------------------
// Tiled propagation step: a CTA owns a PT_W x PT_H pixel tile of one image and up to PT_CH mask planes.
// The planes' tile + halo (max dilation, replicate border baked in while loading) is staged in shared
// memory, so the 48-neighbour gather is `LDS [tile + constant offset]` with no index clamping in the inner
// loop, and the affinity (the only per-iteration stream that does not fit in shared memory) is read exactly
// once per pixel and plane chunk with 128-byte coalesced loads.  Two CTAs fit on an SM (<= 5 planes x 20 KB),
// so one CTA's staging overlaps the other's gather.  Same products and the same summation order as
// par_propagate_kernel: results are bit-identical.
// ---------------------------------------------------------------------------------------------
constexpr int PT_W = 32, PT_H = 16, PT_CH = 5, PT_MAX_HALO = 24;

template <int NDIL, int NP>
__device__ __forceinline__ void par_tile_gather(const float* __restrict__ A, long hw, const float* tile, int pitch,
                                                int plane_stride, int center, const ParGeom& g, float (&acc)[NP]) {
#pragma unroll
  for (int c = 0; c < NP; ++c) acc[c] = 0.0f;
#pragma unroll
  for (int di = 0; di < NDIL; ++di) {
    const int d = g.dil[di];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = __ldg(A + (di * 8 + j) * hw);
      const float* q = tile + center + (c_par_dy[j] * d) * pitch + c_par_dx[j] * d;
#pragma unroll
      for (int c = 0; c < NP; ++c) acc[c] = fmaf(a, q[c * plane_stride], acc[c]);
    }
  }
}

// The reference's only configuration (train_final_voc.py:160: dilations [1,2,4,8,12,24]): every neighbour offset inside the
// staged tile is a compile-time constant, so the gather is `LDS [center + imm]` + FFMA and nothing else.
constexpr int PS_HALO = 24, PS_PITCH = PT_W + 2 * PS_HALO, PS_ROWS = PT_H + 2 * PS_HALO, PS_PLANE = PS_PITCH * PS_ROWS;
__host__ __device__ constexpr int ps_dil(int di) { return di == 0 ? 1 : di == 1 ? 2 : di == 2 ? 4 : di == 3 ? 8 : di == 4 ? 12 : 24; }
__host__ __device__ constexpr int ps_dy(int j) { return j < 3 ? -1 : (j < 5 ? 0 : 1); }
__host__ __device__ constexpr int ps_dx(int j) { return (j == 0 || j == 3 || j == 5) ? -1 : ((j == 1 || j == 6) ? 0 : 1); }

template <int NP>
__device__ __forceinline__ void par_std_gather(const float* __restrict__ A, long hw, const float* q0, float* __restrict__ D) {
  float acc[NP];
#pragma unroll
  for (int c = 0; c < NP; ++c) acc[c] = 0.0f;
#pragma unroll
  for (int di = 0; di < 6; ++di) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = __ldg(A + (di * 8 + j) * hw);
      const int off = (ps_dy(j) * ps_dil(di)) * PS_PITCH + ps_dx(j) * ps_dil(di);  // folds to an immediate after unrolling
#pragma unroll
      for (int c = 0; c < NP; ++c) acc[c] = fmaf(a, q0[c * PS_PLANE + off], acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < NP; ++c) D[c * hw] = acc[c];
}

__global__ void __launch_bounds__(PT_W * PT_H, 2) par_propagate_std_kernel(const float* __restrict__ aff,
                                                                            const float* __restrict__ src,
                                                                            float* __restrict__ dst,
                                                                            const int* __restrict__ nactive, int P,
                                                                            int max_chunks, int h, int w) {
  extern __shared__ float tile[];
  const int b = blockIdx.z / max_chunks;
  const int chunk = blockIdx.z % max_chunks;
  const int live = nactive != nullptr ? min(__ldg(nactive + b), P) : P;
  const int chunks = (live + PT_CH - 1) / PT_CH;
  if (chunk >= chunks) return;
  const int per = (live + chunks - 1) / chunks;
  const int p0 = chunk * per;
  const int np = min(per, live - p0);
  if (np <= 0) return;
  const int x0 = blockIdx.x * PT_W, y0 = blockIdx.y * PT_H;
  const long hw = static_cast<long>(h) * w;
  const float* S = src + (static_cast<long>(b) * P + p0) * hw;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // stage rows wid, wid+16, ..: tile[c][r][cx] = plane_c[clamp(y0 - 24 + r)][clamp(x0 - 24 + cx)]
  for (int r = wid; r < PS_ROWS; r += PT_H) {
    const long row = static_cast<long>(min(max(y0 - PS_HALO + r, 0), h - 1)) * w;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int cx = lane + 32 * k;
      if (cx < PS_PITCH) {
        const long o = row + min(max(x0 - PS_HALO + cx, 0), w - 1);
        for (int c = 0; c < np; ++c) tile[c * PS_PLANE + r * PS_PITCH + cx] = __ldg(S + c * hw + o);
      }
    }
  }
  __syncthreads();
  const int x = x0 + lane, y = y0 + wid;
  if (x >= w || y >= h) return;
  const float* A = aff + static_cast<long>(b) * 48 * hw + static_cast<long>(y) * w + x;
  const float* q0 = tile + (wid + PS_HALO) * PS_PITCH + lane + PS_HALO;
  float* D = dst + (static_cast<long>(b) * P + p0) * hw + static_cast<long>(y) * w + x;
  switch (np) {
    case 1: par_std_gather<1>(A, hw, q0, D); break;
    case 2: par_std_gather<2>(A, hw, q0, D); break;
    case 3: par_std_gather<3>(A, hw, q0, D); break;
    case 4: par_std_gather<4>(A, hw, q0, D); break;
    case 5: par_std_gather<5>(A, hw, q0, D); break;
  }
}

// Same tile kernel with the affinity stream staged by TMA: the per-pixel affinity loads were the exposed latency of
// par_propagate_std_kernel (ncu: issue slots 45 % busy, L1 hit rate 5 %).  Thread 0 keeps PA_STAGES box loads in flight
// (3-D tensor map over [B*48][h][w], box = 3 neighbour planes x 16 rows x 32 pixels = 6 KB) behind an mbarrier ring;
// the gather then reads affinity AND masks from shared memory.  <= 4 planes x 20 KB + 24 KB ring: two CTAs per SM.
constexpr int PA_CH = 4, PA_NG = 3, PA_STAGES = 4, PA_GROUPS = 48 / PA_NG;
constexpr int PA_STAGE_FLOATS = PA_NG * PT_H * PT_W;
constexpr int PA_SMEM_BYTES = (PA_CH * PS_PLANE + PA_STAGES * PA_STAGE_FLOATS) * 4 + (PA_STAGES + 1) * 8;

template <int NP>
__device__ __forceinline__ void par_tma_gather(const float* q0, const float* ring, uint64_t* full, const CUtensorMap* tm,
                                               int x0, int y0, int plane0, int lane, int wid, bool issuer, float* __restrict__ D,
                                               long hw, bool store) {
  float acc[NP];
#pragma unroll
  for (int c = 0; c < NP; ++c) acc[c] = 0.0f;
#pragma unroll
  for (int g = 0; g < PA_GROUPS; ++g) {
    const int s = g % PA_STAGES;
    mbar_wait(&full[s], (g / PA_STAGES) & 1);
    const float* a_s = ring + s * PA_STAGE_FLOATS + wid * PT_W + lane;
#pragma unroll
    for (int jj = 0; jj < PA_NG; ++jj) {
      const int n = g * PA_NG + jj, di = n / 8, j = n % 8;
      const float a = a_s[jj * PT_H * PT_W];
      const int off = (ps_dy(j) * ps_dil(di)) * PS_PITCH + ps_dx(j) * ps_dil(di);
#pragma unroll
      for (int c = 0; c < NP; ++c) acc[c] = fmaf(a, q0[c * PS_PLANE + off], acc[c]);
    }
    if (g + PA_STAGES < PA_GROUPS) {
      __syncthreads();  // every thread is done with ring stage s
      if (issuer) {
        mbar_arrive_expect_tx(&full[s], PA_STAGE_FLOATS * 4);
        tma_load_3d(const_cast<float*>(ring) + s * PA_STAGE_FLOATS, tm, &full[s], x0, y0, plane0 + (g + PA_STAGES) * PA_NG);
      }
    }
  }
  if (store) {
#pragma unroll
    for (int c = 0; c < NP; ++c) D[c * hw] = acc[c];
  }
}

// Persistent: 2 CTAs per SM walk the list of live work items (image, plane chunk, tile).  The number of live planes per
// image is only known on the device (nactive, written by the refine prologue: no host sync), so a plain grid would have to
// cover all ceil(P / 4) = 11 chunks of every image and ~85 % of its CTAs would exit at once.
// The mask tiles (+ halo) are TMA box loads as well: out-of-image cells arrive as zeros and only those cells are then
// overwritten with the replicate-border values by the threads (interior tiles: no staging instructions at all; the
// register-staged version spent as many instructions on staging as on the gather).
__global__ void __launch_bounds__(PT_W * PT_H, 2) par_propagate_tma_kernel(const __grid_constant__ CUtensorMap tm_aff,
                                                                            const __grid_constant__ CUtensorMap tm_src,
                                                                            const float* __restrict__ src,
                                                                            float* __restrict__ dst,
                                                                            const int* __restrict__ nactive, int P, int B,
                                                                            int h, int w, int tiles_x, int tiles_y) {
  extern __shared__ __align__(128) float tile[];
  float* ring = tile + PA_CH * PS_PLANE;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + PA_STAGES * PA_STAGE_FLOATS);
  uint64_t* tile_full = full + PA_STAGES;
  const long hw = static_cast<long>(h) * w;
  const bool issuer = threadIdx.x == 0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int tiles = tiles_x * tiles_y;
  if (issuer) {
    for (int s = 0; s < PA_STAGES; ++s) mbar_init(&full[s], 1);
    mbar_init(tile_full, 1);
    fence_mbar_init();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  int total = 0;
  for (int i = 0; i < B; ++i) {
    const int live = nactive != nullptr ? min(__ldg(nactive + i), P) : P;
    total += ((live + PA_CH - 1) / PA_CH) * tiles;
  }
  uint32_t tile_parity = 0;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    // decode item -> (image b, plane chunk, tile); every stage of the affinity ring completes an even number of phases per
    // item, so the barrier parities of par_tma_gather repeat from item to item
    int b = 0, rem = item, live = 0, chunks = 0;
    for (; b < B; ++b) {
      live = nactive != nullptr ? min(__ldg(nactive + b), P) : P;
      chunks = (live + PA_CH - 1) / PA_CH;
      if (rem < chunks * tiles) break;
      rem -= chunks * tiles;
    }
    const int chunk = rem / tiles, t = rem - chunk * tiles;
    const int per = (live + chunks - 1) / chunks;  // balanced split: 6 live planes -> 3 + 3
    const int p0 = chunk * per;
    const int np = min(per, live - p0);
    const int x0 = (t % tiles_x) * PT_W, y0 = (t / tiles_x) * PT_H;
    if (issuer) {
      mbar_arrive_expect_tx(tile_full, np * PS_PLANE * 4);
      for (int c = 0; c < np; ++c) tma_load_3d(tile + c * PS_PLANE, &tm_src, tile_full, x0 - PS_HALO, y0 - PS_HALO, b * P + p0 + c);
      for (int s = 0; s < PA_STAGES; ++s) {
        mbar_arrive_expect_tx(&full[s], PA_STAGE_FLOATS * 4);
        tma_load_3d(ring + s * PA_STAGE_FLOATS, &tm_aff, &full[s], x0, y0, b * 48 + s * PA_NG);
      }
    }
    mbar_wait(tile_full, tile_parity);
    tile_parity ^= 1;
    // replicate border: cells whose source coordinate lies outside the image (TMA wrote zeros there)
    if (x0 < PS_HALO || y0 < PS_HALO || x0 + PT_W + PS_HALO > w || y0 + PT_H + PS_HALO > h) {
      const float* S = src + (static_cast<long>(b) * P + p0) * hw;
      for (int r = wid; r < PS_ROWS; r += PT_H) {
        const int sy = y0 - PS_HALO + r;
        const int gy = min(max(sy, 0), h - 1);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int cx = lane + 32 * k;
          const int sx = x0 - PS_HALO + cx;
          if (cx < PS_PITCH && (sy != gy || sx < 0 || sx >= w)) {
            const long o = static_cast<long>(gy) * w + min(max(sx, 0), w - 1);
            for (int c = 0; c < np; ++c) tile[c * PS_PLANE + r * PS_PITCH + cx] = __ldg(S + c * hw + o);
          }
        }
      }
      __syncthreads();
    }
    const int x = x0 + lane, y = y0 + wid;
    const bool store = x < w && y < h && np > 0;
    const float* q0 = tile + (wid + PS_HALO) * PS_PITCH + lane + PS_HALO;
    float* D = dst + (static_cast<long>(b) * P + p0) * hw + static_cast<long>(min(y, h - 1)) * w + min(x, w - 1);
    switch (np) {
      case 1: par_tma_gather<1>(q0, ring, full, &tm_aff, x0, y0, b * 48, lane, wid, issuer, D, hw, store); break;
      case 2: par_tma_gather<2>(q0, ring, full, &tm_aff, x0, y0, b * 48, lane, wid, issuer, D, hw, store); break;
      case 3: par_tma_gather<3>(q0, ring, full, &tm_aff, x0, y0, b * 48, lane, wid, issuer, D, hw, store); break;
      default: par_tma_gather<4>(q0, ring, full, &tm_aff, x0, y0, b * 48, lane, wid, issuer, D, hw, store); break;
    }
    __syncthreads();  // everyone is done with the tiles and the ring before the next item overwrites them
  }
}

template <int NDIL>
__global__ void __launch_bounds__(PT_W * PT_H, 2) par_propagate_tiled_kernel(const float* __restrict__ aff,
                                                                              const float* __restrict__ src,
                                                                              float* __restrict__ dst,
                                                                              const int* __restrict__ nactive, int P,
                                                                              int max_chunks, int h, int w, int halo,
                                                                              ParGeom g) {
  extern __shared__ float tile[];
  constexpr int N = 8 * NDIL;
  const int b = blockIdx.z / max_chunks;
  const int chunk = blockIdx.z % max_chunks;
  const int live = nactive != nullptr ? min(__ldg(nactive + b), P) : P;
  const int chunks = (live + PT_CH - 1) / PT_CH;
  if (chunk >= chunks) return;
  const int per = (live + chunks - 1) / chunks;  // balanced split: 6 live planes -> 3 + 3, not 5 + 1
  const int p0 = chunk * per;
  const int np = min(per, live - p0);
  if (np <= 0) return;
  const int x0 = blockIdx.x * PT_W, y0 = blockIdx.y * PT_H;
  const int pitch = PT_W + 2 * halo, rows = PT_H + 2 * halo;
  const int plane_stride = pitch * rows;
  const long hw = static_cast<long>(h) * w;
  const float* S = src + (static_cast<long>(b) * P + p0) * hw;
  // stage: tile[c][r][cx] = plane_c[clamp(y0 - halo + r)][clamp(x0 - halo + cx)]
  for (int i = threadIdx.x; i < plane_stride; i += PT_W * PT_H) {
    const int r = i / pitch, cx = i - r * pitch;
    const int gy = min(max(y0 - halo + r, 0), h - 1);
    const int gx = min(max(x0 - halo + cx, 0), w - 1);
    const long o = static_cast<long>(gy) * w + gx;
    for (int c = 0; c < np; ++c) tile[c * plane_stride + i] = __ldg(S + c * hw + o);
  }
  __syncthreads();
  const int tx = threadIdx.x % PT_W, ty = threadIdx.x / PT_W;
  const int x = x0 + tx, y = y0 + ty;
  if (x >= w || y >= h) return;
  const float* A = aff + static_cast<long>(b) * N * hw + static_cast<long>(y) * w + x;
  const int center = (ty + halo) * pitch + tx + halo;
  float* D = dst + (static_cast<long>(b) * P + p0) * hw + static_cast<long>(y) * w + x;
  switch (np) {
#define PT_CASE(NP)                                                                 \
  case NP: {                                                                        \
    float acc[NP];                                                                  \
    par_tile_gather<NDIL, NP>(A, hw, tile, pitch, plane_stride, center, g, acc);    \
    _Pragma("unroll") for (int c = 0; c < NP; ++c) D[c * hw] = acc[c];              \
  } break;
    PT_CASE(1) PT_CASE(2) PT_CASE(3) PT_CASE(4) PT_CASE(5)
#undef PT_CASE
  }
}

// ---------------------------------------------------------------------------------------------
// Refine prologue: one thread per half-resolution pixel.
// Live planes of image i: a = v*nch + slot, v in {high, low}, slot 0 = background, slot s>0 = s-th
// present class in ascending order (== valid_key of the reference), nch = 1 + #present.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float mean2x2(const float* __restrict__ p, int W) {
  // bilinear(align_corners=False) to exactly half size: l0 = l1 = 0.5 on both axes, torch op order
  const float2 r0 = __ldg(reinterpret_cast<const float2*>(p));
  const float2 r1 = __ldg(reinterpret_cast<const float2*>(p + W));
  const float t0 = __fmaf_rn(0.5f, r0.x, __fmul_rn(0.5f, r0.y));
  const float t1 = __fmaf_rn(0.5f, r1.x, __fmul_rn(0.5f, r1.y));
  return __fmaf_rn(0.5f, t0, __fmul_rn(0.5f, t1));
}

__global__ void __launch_bounds__(256) refine_prologue_kernel(dupl_refine_prologue_args a, int* __restrict__ nactive) {
  const int h2 = a.H / 2, w2 = a.W / 2;
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  const int b = blockIdx.z;
  if (x >= w2 || y >= h2) return;
  const long HW = static_cast<long>(a.H) * a.W;
  const long hw = static_cast<long>(h2) * w2;
  const long src_off = static_cast<long>(2 * y) * a.W + 2 * x;
  const long dst_off = static_cast<long>(y) * w2 + x;
  const int P = 2 * (a.K + 1);

#pragma unroll
  for (int c = 0; c < 3; ++c)
    a.images_ds[(static_cast<long>(b) * 3 + c) * hw + dst_off] = mean2x2(a.images + (static_cast<long>(b) * 3 + c) * HW + src_off, a.W);

  const float* cl = a.cls_label + static_cast<long>(b) * a.K;
  int nch = 1;
  for (int k = 0; k < a.K; ++k) nch += (__ldg(cl + k) != 0.0f);
  if (x == 0 && y == 0) nactive[b] = 2 * nch;

  const float bh = a.bkg_h != nullptr ? mean2x2(a.bkg_h + static_cast<long>(b) * HW + src_off, a.W) : a.bkg_h_scalar;
  const float bl = a.bkg_l_scalar;
  float* mh = a.masks + static_cast<long>(b) * P * hw + dst_off;  // variant high: planes [0, nch)
  float* ml = mh + static_cast<long>(nch) * hw;                   // variant low : planes [nch, 2 nch)

  // pass 1: down-sampled class scores -> planes (raw), running max
  float mx = -INFINITY;
  int slot = 1;
  for (int k = 0; k < a.K; ++k) {
    if (__ldg(cl + k) != 0.0f) {
      const float v = mean2x2(a.cams + (static_cast<long>(b) * a.K + k) * HW + src_off, a.W);
      mh[slot * hw] = v;
      mx = fmaxf(mx, v);
      ++slot;
    }
  }
  const float mxh = fmaxf(mx, bh), mxl = fmaxf(mx, bl);
  // pass 2: exp and sums
  const float eh0 = expf(bh - mxh), el0 = expf(bl - mxl);
  float sh = eh0, sl = el0;
  for (int s = 1; s < nch; ++s) {
    const float v = mh[s * hw];
    const float eh = expf(v - mxh), el = expf(v - mxl);
    mh[s * hw] = eh;
    ml[s * hw] = el;
    sh += eh;
    sl += el;
  }
  // pass 3: normalise
  mh[0] = __fdiv_rn(eh0, sh);
  ml[0] = __fdiv_rn(el0, sl);
  for (int s = 1; s < nch; ++s) {
    mh[s * hw] = __fdiv_rn(mh[s * hw], sh);
    ml[s * hw] = __fdiv_rn(ml[s * hw], sl);
  }
}

// ---------------------------------------------------------------------------------------------
// Refine epilogue: one thread per full-resolution pixel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void py_slice2(int a, int b, int n, int& lo, int& hi) {
  lo = a < 0 ? max(a + n, 0) : min(a, n);
  hi = b < 0 ? max(b + n, 0) : min(b, n);
}

__global__ void __launch_bounds__(256) refine_epilogue_kernel(dupl_refine_epilogue_args a) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  const int b = blockIdx.z;
  if (x >= a.W || y >= a.H) return;
  const int h2 = a.H / 2, w2 = a.W / 2;
  const long hw = static_cast<long>(h2) * w2;
  const int P = 2 * (a.K + 1);
  const float* cl = a.cls_label + static_cast<long>(b) * a.K;
  int nch = 1;
  for (int k = 0; k < a.K; ++k) nch += (__ldg(cl + k) != 0.0f);

  const Lin ly = lin_coord(y, h2, static_cast<float>(h2) / a.H);
  const Lin lx = lin_coord(x, w2, static_cast<float>(w2) / a.W);
  const float* base = a.masks + static_cast<long>(b) * P * hw;
  int arg[2];
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    float best = 0.0f;
    int bi = 0;
    for (int s = 0; s < nch; ++s) {
      const float val = bilerp(base + static_cast<long>(v * nch + s) * hw, w2, ly, lx);
      if (s == 0 || val > best) {  // first index on ties (torch.argmax)
        best = val;
        bi = s;
      }
    }
    arg[v] = bi;
  }
  // valid_key lookup: slot s > 0 -> (index of the s-th present class) + 1
  float key[2] = {0.0f, 0.0f};
  int slot = 0;
  for (int k = 0; k < a.K; ++k) {
    if (__ldg(cl + k) != 0.0f) {
      ++slot;
      if (arg[0] == slot) key[0] = static_cast<float>(k + 1);
      if (arg[1] == slot) key[1] = static_cast<float>(k + 1);
    }
  }
  const int* bx = a.img_box + 4 * b;
  int y0, y1, x0, x1;
  py_slice2(bx[0], bx[1], a.H, y0, y1);
  py_slice2(bx[2], bx[3], a.W, x0, x1);
  const bool inside = y >= y0 && y < y1 && x >= x0 && x < x1;
  const float lh = inside ? key[0] : a.ignore_index;
  const float ll = inside ? key[1] : a.ignore_index;
  float out = lh;
  if (lh == 0.0f) out = a.ignore_index;
  if (lh + ll == 0.0f) out = 0.0f;
  const long o = (static_cast<long>(b) * a.H + y) * a.W + x;
  a.label[o] = out;
  if (a.label_h != nullptr) a.label_h[o] = lh;
  if (a.label_l != nullptr) a.label_l[o] = ll;
}

static int fill_geom(const int32_t* dil, int ndil, ParGeom& g) {
  if (dil == nullptr || ndil < 1 || ndil > DUPL_PAR_MAX_DIL) {
    set_error("PAR: ndil=%d (1..%d supported)", ndil, DUPL_PAR_MAX_DIL);
    return DUPL_ERR_INVALID_ARGUMENT;
  }
  g.ndil = ndil;
  for (int i = 0; i < DUPL_PAR_MAX_DIL; ++i) g.dil[i] = i < ndil ? dil[i] : 0;
  return DUPL_OK;
}

// softmax_n(-(pos/(std+1e-8)/w1)^2) of PAR.get_pos (PAR.py:51-62, 84-87): a constant vector.
static void pos_prior(const ParGeom& g, float w1, ParPos& out) {
  const int N = 8 * g.ndil;
  float pos[PAR_MAX_N];
  const float r2 = static_cast<float>(sqrt(2.0));
  static const int diag[8] = {1, 0, 1, 0, 0, 1, 0, 1};
  for (int di = 0; di < g.ndil; ++di)
    for (int j = 0; j < 8; ++j) pos[di * 8 + j] = (diag[j] ? r2 : 1.0f) * static_cast<float>(g.dil[di]);
  double mean = 0.0;
  for (int n = 0; n < N; ++n) mean += pos[n];
  mean /= N;
  double ss = 0.0;
  for (int n = 0; n < N; ++n) ss += (pos[n] - mean) * (pos[n] - mean);
  const float sd = static_cast<float>(sqrt(ss / (N - 1)));
  float e[PAR_MAX_N], mx = -INFINITY;
  for (int n = 0; n < N; ++n) {
    const float t = pos[n] / (sd + 1e-8f) / w1;
    e[n] = -(t * t);
    mx = fmaxf(mx, e[n]);
  }
  float s = 0.0f;
  for (int n = 0; n < N; ++n) {
    e[n] = expf(e[n] - mx);
    s += e[n];
  }
  for (int n = 0; n < PAR_MAX_N; ++n) out.v[n] = n < N ? e[n] / s : 0.0f;
}

}  // namespace dupl

using namespace dupl;

extern "C" int dupl_par_affinity(const float* imgs, float* aff, int32_t B, int32_t C, int32_t h, int32_t w,
                                 const int32_t* dilations_host, int32_t ndil, float w1, float w2, void* stream) {
  DUPL_CHECK_ARG(imgs && aff && B > 0 && C > 0 && h > 0 && w > 0, "dupl_par_affinity: bad arguments");
  ParGeom g;
  int rc = fill_geom(dilations_host, ndil, g);
  if (rc) return rc;
  ParPos pos;
  pos_prior(g, w1, pos);
  dim3 grid(cdiv(w, 32), cdiv(h, 8), B), block(32, 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (ndil) {
#define AFF_CASE(ND) case ND: par_affinity_kernel<ND><<<grid, block, 0, st>>>(imgs, aff, C, h, w, g, pos, w1, w2); break;
    AFF_CASE(1) AFF_CASE(2) AFF_CASE(3) AFF_CASE(4) AFF_CASE(5) AFF_CASE(6) AFF_CASE(7) AFF_CASE(8)
#undef AFF_CASE
  }
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_par_propagate(const float* aff, float* masks, float* scratch, const int32_t* nactive, int32_t B,
                                  int32_t P, int32_t h, int32_t w, const int32_t* dilations_host, int32_t ndil,
                                  int32_t num_iter, int32_t* result_in_scratch_host, void* stream) {
  DUPL_CHECK_ARG(aff && masks && scratch && B > 0 && P > 0 && h > 0 && w > 0 && num_iter >= 0,
                 "dupl_par_propagate: bad arguments");
  ParGeom g;
  int rc = fill_geom(dilations_host, ndil, g);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int halo = 0;
  for (int i = 0; i < ndil; ++i) halo = g.dil[i] > halo ? g.dil[i] : halo;
  const bool force_simple = getenv("DUPL_PAR_SIMPLE") != nullptr;  // A/B switch for tests and measurements
  if (halo <= PT_MAX_HALO && !force_simple && static_cast<long>(B) * cdiv(P, PT_CH) <= 65535) {
    const int max_chunks = cdiv(P, PT_CH);
    const size_t smem = static_cast<size_t>(PT_CH) * (PT_W + 2 * halo) * (PT_H + 2 * halo) * sizeof(float);
    dim3 tgrid(cdiv(w, PT_W), cdiv(h, PT_H), B * max_chunks), tblock(PT_W * PT_H);
    float* tsrc = masks;
    float* tdst = scratch;
    // standard dilations + 16-byte aligned rows: affinity staged by TMA (DUPL_PAR_NO_TMA=1: plain loads, for A/B runs)
    if (ndil == 6 && g.dil[0] == 1 && g.dil[1] == 2 && g.dil[2] == 4 && g.dil[3] == 8 && g.dil[4] == 12 && g.dil[5] == 24 &&
        w % 4 == 0 && getenv("DUPL_PAR_NO_TMA") == nullptr && getenv("DUPL_PAR_GENERIC_TILE") == nullptr) {
      CUtensorMap tm, tm_a, tm_b;  // affinity; the two mask buffers the iterations ping-pong between
      int rc2 = make_tmap_f32_3d(&tm, aff, w, h, static_cast<uint64_t>(B) * 48, PT_W, PT_H, PA_NG);
      if (rc2) return rc2;
      if ((rc2 = make_tmap_f32_3d(&tm_a, masks, w, h, static_cast<uint64_t>(B) * P, PS_PITCH, PS_ROWS, 1))) return rc2;
      if ((rc2 = make_tmap_f32_3d(&tm_b, scratch, w, h, static_cast<uint64_t>(B) * P, PS_PITCH, PS_ROWS, 1))) return rc2;
      static bool attr_tma = false;
      if (!attr_tma) {
        DUPL_CUDA_OK(cudaFuncSetAttribute(par_propagate_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PA_SMEM_BYTES));
        attr_tma = true;
      }
      const int tiles_x = cdiv(w, PT_W), tiles_y = cdiv(h, PT_H);
      const long max_items = static_cast<long>(B) * cdiv(P, PA_CH) * tiles_x * tiles_y;
      const int ctas = static_cast<int>(max_items < 2L * sm_count() ? max_items : 2L * sm_count());
      for (int it = 0; it < num_iter; ++it) {
        par_propagate_tma_kernel<<<ctas, tblock, PA_SMEM_BYTES, st>>>(tm, tsrc == masks ? tm_a : tm_b, tsrc, tdst, nactive, P, B, h, w,
                                                                      tiles_x, tiles_y);
        DUPL_LAUNCH_OK();
        float* t = tsrc; tsrc = tdst; tdst = t;
      }
      if (result_in_scratch_host != nullptr) *result_in_scratch_host = (tsrc == scratch) ? 1 : 0;
      return DUPL_OK;
    }
    const bool standard = ndil == 6 && g.dil[0] == 1 && g.dil[1] == 2 && g.dil[2] == 4 && g.dil[3] == 8 && g.dil[4] == 12 &&
                          g.dil[5] == 24 && getenv("DUPL_PAR_GENERIC_TILE") == nullptr;
    if (standard) {
      static bool attr_std = false;
      if (!attr_std) {
        DUPL_CUDA_OK(cudaFuncSetAttribute(par_propagate_std_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          PT_CH * PS_PLANE * 4));
        attr_std = true;
      }
    }
    for (int it = 0; it < num_iter; ++it) {
      if (standard) {
        par_propagate_std_kernel<<<tgrid, tblock, PT_CH * PS_PLANE * sizeof(float), st>>>(aff, tsrc, tdst, nactive, P, max_chunks, h, w);
        DUPL_LAUNCH_OK();
        float* t = tsrc; tsrc = tdst; tdst = t;
        continue;
      }
      switch (ndil) {
#define TILED_CASE(ND)                                                                                                     \
  case ND: {                                                                                                               \
    static bool attr = false;                                                                                              \
    if (!attr) {                                                                                                           \
      DUPL_CUDA_OK(cudaFuncSetAttribute(par_propagate_tiled_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                        PT_CH * (PT_W + 2 * PT_MAX_HALO) * (PT_H + 2 * PT_MAX_HALO) * 4));                  \
      attr = true;                                                                                                         \
    }                                                                                                                      \
    par_propagate_tiled_kernel<ND><<<tgrid, tblock, smem, st>>>(aff, tsrc, tdst, nactive, P, max_chunks, h, w, halo, g);   \
  } break;
        TILED_CASE(1) TILED_CASE(2) TILED_CASE(3) TILED_CASE(4) TILED_CASE(5) TILED_CASE(6) TILED_CASE(7) TILED_CASE(8)
#undef TILED_CASE
      }
      DUPL_LAUNCH_OK();
      float* t = tsrc; tsrc = tdst; tdst = t;
    }
    if (result_in_scratch_host != nullptr) *result_in_scratch_host = (tsrc == scratch) ? 1 : 0;
    return DUPL_OK;
  }
  constexpr int CH = 8;
  const int chunks = cdiv(P, CH);
  DUPL_CHECK_ARG(static_cast<long>(B) * chunks <= 65535, "dupl_par_propagate: B*chunks=%ld exceeds the launch grid",
                 static_cast<long>(B) * chunks);
  dim3 grid(cdiv(w, 32), cdiv(h, 8), B * chunks), block(32, 8);
  float* src = masks;
  float* dst = scratch;
  for (int it = 0; it < num_iter; ++it) {
    switch (ndil) {
#define PROP_CASE(ND) \
  case ND: par_propagate_kernel<ND, CH><<<grid, block, 0, st>>>(aff, src, dst, nactive, P, chunks, h, w, g); break;
      PROP_CASE(1) PROP_CASE(2) PROP_CASE(3) PROP_CASE(4) PROP_CASE(5) PROP_CASE(6) PROP_CASE(7) PROP_CASE(8)
#undef PROP_CASE
    }
    DUPL_LAUNCH_OK();
    float* t = src; src = dst; dst = t;
  }
  if (result_in_scratch_host != nullptr) *result_in_scratch_host = (src == scratch) ? 1 : 0;
  return DUPL_OK;
}

extern "C" int dupl_refine_prologue(const dupl_refine_prologue_args* a, void* stream) {
  DUPL_CHECK_ARG(a != nullptr, "dupl_refine_prologue: args is NULL");
  DUPL_CHECK_ARG(a->images && a->cams && a->cls_label && a->images_ds && a->masks && a->nactive,
                 "dupl_refine_prologue: NULL pointer");
  DUPL_CHECK_ARG(a->b > 0 && a->K > 0 && a->H > 0 && a->W > 0 && a->H % 2 == 0 && a->W % 2 == 0,
                 "dupl_refine_prologue: bad shape b=%d K=%d H=%d W=%d (H, W must be even)", a->b, a->K, a->H, a->W);
  dim3 grid(cdiv(a->W / 2, 32), cdiv(a->H / 2, 8), a->b), block(32, 8);
  refine_prologue_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(*a, a->nactive);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}

extern "C" int dupl_refine_epilogue(const dupl_refine_epilogue_args* a, void* stream) {
  DUPL_CHECK_ARG(a != nullptr, "dupl_refine_epilogue: args is NULL");
  DUPL_CHECK_ARG(a->masks && a->cls_label && a->img_box && a->label, "dupl_refine_epilogue: NULL pointer");
  DUPL_CHECK_ARG(a->b > 0 && a->K > 0 && a->H > 0 && a->W > 0 && a->H % 2 == 0 && a->W % 2 == 0,
                 "dupl_refine_epilogue: bad shape");
  dim3 grid(cdiv(a->W, 32), cdiv(a->H, 8), a->b), block(32, 8);
  refine_epilogue_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}
