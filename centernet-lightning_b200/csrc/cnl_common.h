// Shared helpers for the cnl_b200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include "../../include/cnl_b200.h"

namespace cnl {

// Thread-local error text returned by cnl_last_error().
char* error_buffer();
int fail(int code, const char* fmt, ...);

#define CNL_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::cnl::fail(CNL_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                             \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#ifdef __CUDACC__
// THE logistic of this library (cnl_sigmoid, the from_logits decode, the heat-map epilogue of the forward engine): three
// separately rounded fp32 operations, the arithmetic torch.sigmoid runs on a CUDA tensor - i.e. what the reference's
// `.sigmoid()` (centernet_lightning/models/centernet.py:205) computes on the device.
__device__ __forceinline__ float sigmoid32(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }
#endif

}  // namespace cnl
