// Shared helpers for the dmt_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/dmt_b200.h"

namespace dmt {

// thread-local last-error message (the only mutable global state of the library)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define DMT_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::dmt::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

#define DMT_CUDA_LAUNCH_CHECK(what)                              \
  do {                                                           \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return ::dmt::cuda_fail(_e, (what));  \
  } while (0)

constexpr float kLnEps = 1e-8f;  // TransformerModel_util.py:58

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

// streaming (read-once) 128-bit load that does not allocate in L1
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// ---- programmatic dependent launch (sm_90+): a kernel launched with launch_pdl may begin -- CTAs scheduled, prologue
// run -- as soon as every CTA of the kernel before it in the stream has called griddep_launch() (or exited) and an SM
// has room; it must call griddep_wait() before it touches anything that kernel writes (the wait returns when the
// previous kernel has completed and its memory is visible).  Both instructions are no-ops in a plain launch.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool off = getenv("DMT_PDL") && atoi(getenv("DMT_PDL")) == 0;      // diagnostic switch: plain launches
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = off ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);
}

inline int sm_count_cached() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace dmt
