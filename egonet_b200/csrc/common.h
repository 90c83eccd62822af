// Shared host-side helpers for the egonet_b200 native library.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

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

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace egn
