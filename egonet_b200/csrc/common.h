// Shared host-side helpers for the egonet_b200 native library.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>

#include "egonet_b200.h"

namespace egn {

// thread-local last-error message (egn_last_error)
void set_error(const char* fmt, ...);
const char* get_error();

// returns EGN_OK or sets the error and returns EGN_ERR_NO_DEVICE
int require_device();

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define EGN_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::egn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                       __LINE__);                                                         \
      return EGN_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define EGN_LAUNCH_CHECK(what)                                                            \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      ::egn::set_error("launch of %s failed: %s", what, cudaGetErrorString(_e));          \
      return EGN_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define EGN_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::egn::set_error(__VA_ARGS__);                                                      \
      return EGN_ERR_INVALID;                                                             \
    }                                                                                     \
  } while (0)


// ---------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the HC forward is launched with the
// programmatic-stream-serialisation attribute: its CTAs may be scheduled while the previous kernel
// of the stream is still draining, run their independent prologue (barrier init, TMEM allocation,
// tensor-map prefetch, weight loads) and then block in pdl_wait() until the previous grid has
// completed and flushed its writes.  Rules kept by every such kernel:
//   * pdl_wait() is executed by every thread that reads or writes activation memory, before it does;
//   * at least the threads that produce the kernel's output wait, so "this grid completed" always
//     implies "its predecessor completed" (dependencies reach back more than one launch);
//   * pdl_trigger() right at the top: dependents only become schedulable once every CTA of this grid
//     has started, so they never compete with its own unscheduled CTAs.
// EGN_PDL=0 launches everything fully serialised (the instructions are then no-ops).
// ---------------------------------------------------------------------------
bool pdl_enabled();

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
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
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Same, with thread-block clusters of `cluster_x` CTAs along x (CTA pairs for tcgen05 cta_group::2).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                      unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace egn
