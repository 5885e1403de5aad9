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

struct WidenIdsArgs {
  dmt_widen_ids_desc d[DMT_MAX_WIDEN];
  int32_t n;
};

// 1- / 2- / 3-byte little-endian ids -> int32: blockIdx.y = array; a thread widens 16 consecutive ids (16 / 32 / 48
// source bytes as 16-byte loads, four 16-byte stores); arrays start on 256-byte boundaries; the ragged tail is bytewise.
__global__ void __launch_bounds__(256) widen_ids_kernel(const __grid_constant__ WidenIdsArgs a) {
  const dmt_widen_ids_desc& d = a.d[blockIdx.y];
  const int nb = d.bytes;
  const int64_t groups = d.n >> 4;
  const uint4* __restrict__ src = reinterpret_cast<const uint4*>(d.src);
  int4* __restrict__ dst = reinterpret_cast<int4*>(d.dst);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    uint32_t w[12];
    int id[16];
    if (nb == 1) {
      const uint4 v = __ldcs(src + g);
      w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
#pragma unroll
      for (int i = 0; i < 16; ++i) id[i] = (int)((w[i >> 2] >> (8 * (i & 3))) & 0xffu);
    } else if (nb == 2) {
      const uint4 v0 = __ldcs(src + 2 * g), v1 = __ldcs(src + 2 * g + 1);
      w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
#pragma unroll
      for (int i = 0; i < 16; ++i) id[i] = (int)((w[i >> 1] >> (16 * (i & 1))) & 0xffffu);
    } else {
      const uint4 v0 = __ldcs(src + 3 * g), v1 = __ldcs(src + 3 * g + 1), v2 = __ldcs(src + 3 * g + 2);
      w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
      w[8] = v2.x; w[9] = v2.y; w[10] = v2.z; w[11] = v2.w;
#pragma unroll
      for (int q = 0; q < 4; ++q) {                      // 4 ids = 3 words
        const uint32_t a0 = w[3 * q], a1 = w[3 * q + 1], a2 = w[3 * q + 2];
        id[4 * q + 0] = (int)(a0 & 0xffffffu);
        id[4 * q + 1] = (int)((a0 >> 24) | ((a1 & 0xffffu) << 8));
        id[4 * q + 2] = (int)((a1 >> 16) | ((a2 & 0xffu) << 16));
        id[4 * q + 3] = (int)(a2 >> 8);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[4 * g + q] = make_int4(id[4 * q], id[4 * q + 1], id[4 * q + 2], id[4 * q + 3]);
  }
  if (blockIdx.x == 0) {
    const uint8_t* __restrict__ sb = reinterpret_cast<const uint8_t*>(d.src);
    for (int64_t t = (groups << 4) + threadIdx.x; t < d.n; t += blockDim.x) {
      uint32_t v = 0;
      for (int k = 0; k < nb; ++k) v |= (uint32_t)sb[t * nb + k] << (8 * k);
      d.dst[t] = (int32_t)v;
    }
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

int dmt_widen_ids(int32_t n_arrays, const dmt_widen_ids_desc* arrays, void* stream) {
  DMT_REQUIRE(n_arrays >= 0 && (arrays || n_arrays == 0), DMT_ERR_INVALID_ARGUMENT, "dmt_widen_ids: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int base = 0; base < n_arrays; base += DMT_MAX_WIDEN) {
    dmt::WidenIdsArgs a;
    a.n = n_arrays - base < DMT_MAX_WIDEN ? n_arrays - base : DMT_MAX_WIDEN;
    int64_t longest = 0;
    for (int i = 0; i < a.n; ++i) {
      const dmt_widen_ids_desc& d = arrays[base + i];
      DMT_REQUIRE(d.n >= 0 && ((d.src && d.dst) || d.n == 0), DMT_ERR_INVALID_ARGUMENT,
                  "dmt_widen_ids: array %d is incomplete", base + i);
      DMT_REQUIRE(d.bytes >= 1 && d.bytes <= 3, DMT_ERR_INVALID_ARGUMENT, "dmt_widen_ids: array %d: %d bytes per id",
                  base + i, d.bytes);
      DMT_REQUIRE((((uintptr_t)d.src | (uintptr_t)d.dst) & 15) == 0, DMT_ERR_INVALID_ARGUMENT,
                  "dmt_widen_ids: array %d is not 16-byte aligned", base + i);
      a.d[i] = d;
      if (d.n > longest) longest = d.n;
    }
    for (int i = a.n; i < DMT_MAX_WIDEN; ++i) a.d[i] = dmt_widen_ids_desc{nullptr, nullptr, 0, 1, 0};
    if (longest == 0) continue;
    int64_t bx = ((longest >> 4) + 255) / 256;
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    dmt::widen_ids_kernel<<<dim3((unsigned)bx, (unsigned)a.n), 256, 0, st>>>(a);
    DMT_CUDA_LAUNCH_CHECK("widen_ids_kernel");
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
