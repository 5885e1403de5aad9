// Device side of the compact host->device batch format (data.py: PackedBatch(compact=True)).
//
// The reference's input pipeline hands the model int64 ids and fp32 features (tfrecord_mask.py:23-84); over PCIe
// that is 18.4 MB per 4096-sample batch and, with 8 ranks sharing the host's memory system, the end-to-end limiter
// (SCALE_r01: 0.58 efficiency at 8 GPUs).  The compact format ships
//   * id arrays of small-vocabulary features (category / time-bucket ids < 65536) as uint16,
//   * the dense `features` block as bf16 (the bf16 tensor-core path rounds it to bf16 before its first GEMM anyway),
// and these two kernels restore what the compute kernels consume: int32 id arrays (dmt_widen_u16, all arrays of a
// batch in ONE launch) and the fp32 feature columns of the MMoE input (dmt_copy_dense_features_bf16).
// Pure byte movement, HBM/L2 bound: bytes = n * (2 + 4) per widened id, B * dim * (2 + 4) for the features.
#include <cuda_bf16.h>

#include "dmt_common.cuh"

namespace dmt {

struct WidenArgs {
  dmt_widen_desc d[DMT_MAX_WIDEN];
  int32_t n;
};

// blockIdx.y = array; a thread widens 8 consecutive ids (one 16-byte load, two 16-byte stores); arrays start on
// 256-byte boundaries of the packed buffer, so the vector accesses are aligned; the ragged tail is scalar.
__global__ void __launch_bounds__(256) widen_u16_kernel(const __grid_constant__ WidenArgs a) {
  const dmt_widen_desc& d = a.d[blockIdx.y];
  const int64_t groups = d.n >> 3;
  const uint4* __restrict__ src = reinterpret_cast<const uint4*>(d.src);
  int4* __restrict__ dst = reinterpret_cast<int4*>(d.dst);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const uint4 v = __ldcs(src + g);
    dst[2 * g] = make_int4((int)(v.x & 0xffffu), (int)(v.x >> 16), (int)(v.y & 0xffffu), (int)(v.y >> 16));
    dst[2 * g + 1] = make_int4((int)(v.z & 0xffffu), (int)(v.z >> 16), (int)(v.w & 0xffffu), (int)(v.w >> 16));
  }
  if (blockIdx.x == 0) {
    const int64_t t = (groups << 3) + threadIdx.x;
    if (t < d.n) d.dst[t] = (int32_t)d.src[t];
  }
}

// one warp per row; a lane converts 2 consecutive columns per trip (bf16x2 in, two fp32 out)
__global__ void __launch_bounds__(256)
copy_dense_bf16_kernel(const __nv_bfloat16* __restrict__ src, int batch, int dim, float* __restrict__ dst, int64_t ld) {
  const int lane = threadIdx.x & 31;
  for (int b = blockIdx.x * 8 + (threadIdx.x >> 5); b < batch; b += gridDim.x * 8) {
    const __nv_bfloat16* __restrict__ s = src + (int64_t)b * dim;
    float* __restrict__ d = dst + (int64_t)b * ld;
    for (int c = lane; c < dim; c += 32) d[c] = __bfloat162float(s[c]);
  }
}

// dense features -> bf16 columns [0, dim) of the bf16 MMoE input: one warp per row, a lane converts two consecutive
// columns per trip (two coalesced scalar loads -- a 615-wide source row is only element-aligned -- one 4-byte bf16x2
// store: the destination rows are 16-byte aligned), four trips in flight
template <typename T>
__device__ __forceinline__ float dense_ld(const T* p) {
  if constexpr (sizeof(T) == 2) return __bfloat162float(*p);
  else return __ldcs(p);
}
template <typename T>
__global__ void __launch_bounds__(256)
stage_dense_bf16_kernel(const T* __restrict__ src, int batch, int dim, __nv_bfloat16* __restrict__ dst, int64_t ld) {
  const int lane = threadIdx.x & 31;
  for (int b = blockIdx.x * 8 + (threadIdx.x >> 5); b < batch; b += gridDim.x * 8) {
    const T* __restrict__ s = src + (int64_t)b * dim;
    __nv_bfloat16* __restrict__ d = dst + (int64_t)b * ld;
    int c = 2 * lane;
    for (; c + 192 + 1 < dim; c += 256) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[2 * u] = dense_ld(s + c + 64 * u);
        v[2 * u + 1] = dense_ld(s + c + 64 * u + 1);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) *reinterpret_cast<__nv_bfloat162*>(d + c + 64 * u) = __floats2bfloat162_rn(v[2 * u], v[2 * u + 1]);
    }
    for (; c < dim; c += 64) {
      if (c + 1 < dim) *reinterpret_cast<__nv_bfloat162*>(d + c) = __floats2bfloat162_rn(dense_ld(s + c), dense_ld(s + c + 1));
      else d[c] = __float2bfloat16(dense_ld(s + c));
    }
  }
}

}  // namespace dmt

extern "C" {

int dmt_stage_dense_features_bf16(const void* features, int32_t features_are_bf16, int32_t batch, int32_t dim,
                                  void* out_bf16, int64_t out_ld, void* stream) {
  DMT_REQUIRE(features && out_bf16 && batch >= 0 && dim > 0 && out_ld >= dim, DMT_ERR_INVALID_ARGUMENT,
              "dmt_stage_dense_features_bf16: bad arguments");
  if (batch == 0) return DMT_OK;
  DMT_REQUIRE(out_ld % 2 == 0 && ((uintptr_t)out_bf16 & 3) == 0, DMT_ERR_INVALID_ARGUMENT,
              "dmt_stage_dense_features_bf16: out_ld must be even and the output 4-byte aligned");
  int64_t blocks = ((int64_t)batch + 7) / 8;           // one warp per row
  const int64_t cap = (int64_t)dmt::sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  if (features_are_bf16)
    dmt::stage_dense_bf16_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)features, batch, dim, (__nv_bfloat16*)out_bf16, out_ld);
  else
    dmt::stage_dense_bf16_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const float*)features, batch, dim, (__nv_bfloat16*)out_bf16, out_ld);
  DMT_CUDA_LAUNCH_CHECK("stage_dense_bf16_kernel");
  return DMT_OK;
}

int dmt_widen_u16(int32_t n_arrays, const dmt_widen_desc* arrays, void* stream) {
  DMT_REQUIRE(n_arrays >= 0 && (arrays || n_arrays == 0), DMT_ERR_INVALID_ARGUMENT, "dmt_widen_u16: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int base = 0; base < n_arrays; base += DMT_MAX_WIDEN) {
    dmt::WidenArgs a;
    a.n = n_arrays - base < DMT_MAX_WIDEN ? n_arrays - base : DMT_MAX_WIDEN;
    int64_t longest = 0;
    for (int i = 0; i < a.n; ++i) {
      const dmt_widen_desc& d = arrays[base + i];
      DMT_REQUIRE(d.n >= 0 && ((d.src && d.dst) || d.n == 0), DMT_ERR_INVALID_ARGUMENT,
                  "dmt_widen_u16: array %d is incomplete", base + i);
      DMT_REQUIRE((((uintptr_t)d.src | (uintptr_t)d.dst) & 15) == 0, DMT_ERR_INVALID_ARGUMENT,
                  "dmt_widen_u16: array %d is not 16-byte aligned", base + i);
      a.d[i] = d;
      if (d.n > longest) longest = d.n;
    }
    for (int i = a.n; i < DMT_MAX_WIDEN; ++i) a.d[i] = dmt_widen_desc{nullptr, nullptr, 0};
    if (longest == 0) continue;
    int64_t bx = ((longest >> 3) + 255) / 256;
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;                    // id arrays are <= a few 100 k entries: grid-stride beyond
    dmt::widen_u16_kernel<<<dim3((unsigned)bx, (unsigned)a.n), 256, 0, st>>>(a);
    DMT_CUDA_LAUNCH_CHECK("widen_u16_kernel");
  }
  return DMT_OK;
}

int dmt_copy_dense_features_bf16(const void* features_bf16, int32_t batch, int32_t dim, float* out, int64_t out_ld,
                                 void* stream) {
  DMT_REQUIRE(features_bf16 && out && batch >= 0 && dim > 0 && out_ld >= dim, DMT_ERR_INVALID_ARGUMENT,
              "dmt_copy_dense_features_bf16: bad arguments");
  if (batch == 0) return DMT_OK;
  int64_t blocks = ((int64_t)batch + 7) / 8;           // one warp per row
  const int64_t cap = (int64_t)dmt::sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  dmt::copy_dense_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)features_bf16, batch, dim, out, out_ld);
  DMT_CUDA_LAUNCH_CHECK("copy_dense_bf16_kernel");
  return DMT_OK;
}

}  // extern "C"
