// Shared helpers of libtzddpc.so (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
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

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// Small batches (the batch-1 closed loop: 12 back-to-back launches of ~25 us) are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel of the stream may be scheduled (and run its prologue)
// as soon as every CTA of this one has called pdl_trigger() or exited, and blocks in pdl_wait() until this grid has
// completed and its writes are visible.  Everything a kernel reads that an earlier kernel of the stream may have written
// comes after pdl_wait().  Both are no-ops in a launch without the attribute.
// Measured (B200): batch 1 26.8 -> 23.6 us per step; full batches get SLOWER with the attribute (65,536 scenarios:
// 0.0623-0.0631 ms against 0.0598-0.0608; 8,192: 0.040 against 0.030 -- the early-resident CTAs of the next two kernels
// compete with the running one), so launches of kPdlMaxBatch scenarios or more do not use it.
constexpr int64_t kPdlMaxBatch = 256;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// TZDDPC_PDL = 0 turns the attribute off (read per launch, no state)
inline bool pdl_enabled(int64_t batch) {
  const char* e = getenv("TZDDPC_PDL");
  if (e && e[0] == '0') return false;
  if (e && e[0] == '2') return true;           // (experiments: force it on for every batch size)
  return batch < kPdlMaxBatch;
}

template <class... KArgs, class... Args>
cudaError_t launch_kernel(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool pdl,
                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace tz
