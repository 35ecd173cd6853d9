#include "common.h"

#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>

namespace gap {

static thread_local char g_err[512] = "no error";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

static std::mutex g_dbg_mu;
static std::map<std::string, int>& dbg_map() {
  static std::map<std::string, int> m;
  return m;
}
// Knobs exist for bring-up A/B runs only.  Until gap_debug_set is called for the first time (i.e. always, in
// production) a lookup is one relaxed atomic load: no mutex, no string map on the launch path.
static std::atomic<int> g_dbg_any{0};
int debug_get(const char* key, int dflt) {
  if (g_dbg_any.load(std::memory_order_relaxed) == 0) return dflt;
  std::lock_guard<std::mutex> lk(g_dbg_mu);
  auto it = dbg_map().find(key);
  return it == dbg_map().end() ? dflt : it->second;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                     bool swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return GAP_ERR_DRIVER;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("TMA base address %p is not 16-byte aligned", base);
    return GAP_ERR_ALIGNMENT;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
    if (i + 1 < rank) {
      gstr[i] = strides_bytes[i];
      if (gstr[i] % 16 != 0) {
        set_error("TMA stride %d = %llu bytes is not a multiple of 16", i,
                  (unsigned long long)gstr[i]);
        return GAP_ERR_ALIGNMENT;
      }
    }
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                  const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu,%llu box "
              "%u,%u,%u,%u)",
              (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
              (unsigned long long)(rank > 2 ? gdim[2] : 0),
              (unsigned long long)(rank > 3 ? gdim[3] : 0), bx[0], rank > 1 ? bx[1] : 0,
              rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0);
    return GAP_ERR_DRIVER;
  }
  return 0;
}

}  // namespace gap

extern "C" {

const char* gap_last_error_string(void) { return gap::g_err; }
int gap_version(void) { return 100; }
int gap_sm_count(void) { return gap::sm_count(); }
int gap_debug_set(const char* key, int value) {
  std::lock_guard<std::mutex> lk(gap::g_dbg_mu);
  gap::dbg_map()[key] = value;
  gap::g_dbg_any.store(1, std::memory_order_relaxed);
  return 0;
}
}
