#include <stdlib.h>

#include "common.cuh"

#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>

namespace dupl {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is not linked: the entry point is fetched from the driver that is already loaded in
// the process (lazily, on first use in the calling process — never at import / fork time).
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return DUPL_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0) {
    set_error("TMA operand must be 16-byte aligned (base %p, row stride %llu elements)", base,
              static_cast<unsigned long long>(ld));
    return DUPL_ERR_INVALID_ARGUMENT;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %llu cols %llu ld %llu box_rows %u)",
              static_cast<int>(r), static_cast<unsigned long long>(rows), static_cast<unsigned long long>(cols),
              static_cast<unsigned long long>(ld), box_rows);
    return DUPL_ERR_CUDA;
  }
  return DUPL_OK;
}

int make_tmap_f32_3d(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                     uint32_t b2) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return DUPL_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (d0 * 4) % 16 != 0) {
    set_error("TMA operand must be 16-byte aligned (base %p, row of %llu floats)", base, static_cast<unsigned long long>(d0));
    return DUPL_ERR_INVALID_ARGUMENT;
  }
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstride[2] = {d0 * 4, d0 * d1 * 4};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3-D fp32) failed with CUresult %d", static_cast<int>(r));
    return DUPL_ERR_CUDA;
  }
  return DUPL_OK;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("DUPL_PDL");
    return e != nullptr && atoi(e) != 0;
  }();
  return on;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0)
      n = 148;
  }
  return n;
}

static std::atomic<int> g_gemm_sm_limit{0};
int gemm_sms() {
  const int lim = g_gemm_sm_limit.load(std::memory_order_relaxed), n = sm_count();
  return (lim >= 2 && lim < n) ? (lim & ~1) : n;
}

}  // namespace dupl

extern "C" int dupl_set_gemm_sm_limit(int32_t sms) { return dupl::g_gemm_sm_limit.exchange(sms < 0 ? 0 : sms); }
extern "C" int dupl_version(void) { return DUPL_ABI_VERSION; }
extern "C" const char* dupl_last_error(void) { return dupl::g_err; }
extern "C" int64_t dupl_launch_count(void) { return dupl::g_launches.load(); }
