// Host-side helpers shared by the translation units of libgap_b200: error reporting, the
// driver-entry-point lookup for cuTensorMapEncodeTiled (so the library links against cudart only),
// and device properties.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <utility>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gap_b200.h"

namespace gap {

void set_error(const char* fmt, ...);
int sm_count();
int debug_get(const char* key, int dflt);

// Encode a tiled TMA descriptor for a bf16 tensor.  dims/strides are innermost-first; strides[i]
// is the byte stride of dimension i+1 (dimension 0 is contiguous).  Returns 0 or a gap_status.
int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                     bool swizzle128);

// Launch with the programmatic-dependent-launch attribute (see ptx.cuh: pdl_wait / pdl_trigger): the kernel's prologue
// (barrier init, TMEM allocation, descriptor prefetch, launch latency) overlaps the tail of the previous kernel in the
// stream.  Only for kernels that call pdl_wait() before their first global-memory access.
// MEASURED (tools/ab_knob.py pdl 0 1, ABBA-ordered blocks of the eager training step): 8.47 ms without, 8.66 ms with
// the attribute -- early-scheduled dependents disturb the balance between the three streams of the step -- so it is
// OFF unless gap_debug_set("pdl", 1).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = debug_get("pdl", 0) != 0 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Launch as clusters of two CTAs (a CTA pair on the two SMs of one TPC: tcgen05 cta_group::2 kernels).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pair(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = debug_get("pdl", 0) != 0 ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#define GAP_CHECK_ARG(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      gap::set_error(__VA_ARGS__);    \
      return GAP_ERR_BAD_ARG;         \
    }                                 \
  } while (0)

#define GAP_CUDA(call)                                                             \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      gap::set_error("%s failed: %s", #call, cudaGetErrorString(e__));             \
      return static_cast<int>(e__);                                                \
    }                                                                              \
  } while (0)

}  // namespace gap
