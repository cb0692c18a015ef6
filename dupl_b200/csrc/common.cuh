// Host-side helpers shared by every translation unit of libdupl.so:
// error convention of the C-ABI (include/dupl.h) and TMA tensor-map construction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dupl.h"

namespace dupl {

// Thread-local message returned by dupl_last_error().
void set_error(const char* fmt, ...);

#define DUPL_CHECK_ARG(cond, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      ::dupl::set_error(__VA_ARGS__);    \
      return DUPL_ERR_INVALID_ARGUMENT;  \
    }                                    \
  } while (0)

#define DUPL_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::dupl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DUPL_ERR_CUDA;                                                             \
    }                                                                                   \
  } while (0)

// Checks the launch itself (not asynchronous execution errors: the library never synchronises).
#define DUPL_LAUNCH_OK()            \
  do {                              \
    ::dupl::count_launch();         \
    DUPL_CUDA_OK(cudaGetLastError()); \
  } while (0)

// Number of kernels this library has launched in the process (dupl_launch_count()).
void count_launch();

// 2-D bf16 tensor map: `rows` x `cols` elements, row stride `ld` elements, box = box_rows x box_cols
// columns (64 -> 128-byte rows with the 128-byte swizzle, 32 -> 64-byte rows with the 64-byte swizzle); out-of-bounds elements read as zero.
// Returns 0 on success (DUPL_ERR_* otherwise).
int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols = 64);

// 3-D fp32 tensor map over a dense [d2][d1][d0] array (d0 innermost), box = b2 x b1 x b0, no swizzle; out-of-bounds
// elements read as zero.  d0 must be a multiple of 4 (16-byte global strides).
int make_tmap_f32_3d(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                     uint32_t b2);

int sm_count();
int gemm_sms();  // SMs the persistent GEMM grid may occupy (dupl_set_gemm_sm_limit)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Programmatic dependent launch (sm_90+): the ~800 kernels of a captured training step are short (median ~10 us) and each
// boundary costs a launch latency once the previous grid has drained.  Kernels launched through launch_pdl() may be
// SCHEDULED while their predecessor in the stream is still running: every such kernel starts with pdl_sync(), which
// (1) lets its own successor be scheduled early and (2) blocks until the predecessor has completed and its writes are
// visible — so no kernel touches global memory before the grid it depends on is done, and the order of all memory
// operations is the stream order.  MEASURED (B200, captured phase-B step, round 2): 47.31 ms with the attribute vs 46.46 ms
// without — the early-scheduled CTAs wait inside the kernels (colsum_finish 4.0 -> 8.9 us, split_transpose 17.9 -> 35 us of
// recorded duration) and cost more than the launch latency they hide.  OFF by default; DUPL_PDL=1 enables it (without
// the attribute pdl_sync() is a no-op).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

}  // namespace dupl
