// Shared helpers of libtzddpc.so (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/tzddpc.h"

namespace tz {

char* last_error_buf();                                  // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);                // records the message, returns code

#define TZ_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t err__ = (call);                                                              \
    if (err__ != cudaSuccess)                                                                \
      return ::tz::fail(TZ_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
  } while (0)

#define TZ_REQUIRE(cond, ...)                                \
  do {                                                       \
    if (!(cond)) return ::tz::fail(TZ_EINVAL, __VA_ARGS__);  \
  } while (0)

constexpr int kMaxN = 8;   // dim_x supported by the fused path
constexpr int kMaxM = 4;   // dim_u

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace tz
