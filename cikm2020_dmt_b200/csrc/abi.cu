// Library-level exports + the raw row gather (dmt_embed_gather).
#include <stdarg.h>
#include <string.h>

#include "dmt_common.cuh"

namespace dmt {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return DMT_ERR_CUDA;
}

// ---------------------------------------------------------------------------------
// Row gather.  One "slot" = one 16-byte chunk of one output row; a warp covers
// consecutive chunks, so a 128-byte Sku row is one fully coalesced 128-byte request
// and the output is written as contiguous 512-byte warp stores.  Pure HBM-bound
// byte movement: bytes = n*(D*4 + 4) read + n*D*4 written.
template <int VEC>
__global__ void __launch_bounds__(256)
embed_gather_kernel(const float* __restrict__ table, int64_t rows, int dim, const int32_t* __restrict__ ids,
                    int64_t n_ids, int zero_pad, float* __restrict__ out) {
  const int chunks = dim / VEC;
  const int64_t total = n_ids * chunks;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += stride) {
    const int64_t n = s / chunks;
    const int c = (int)(s - n * chunks);
    int64_t row = (int64_t)__ldg(ids + n) - (zero_pad ? 1 : 0);
    const bool valid = row >= 0 && row < rows;
    if (VEC == 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) v = ld_stream4(table + row * dim + c * 4);
      __stcs(reinterpret_cast<float4*>(out + n * dim + c * 4), v);
    } else {
      float v = valid ? __ldg(table + row * dim + c) : 0.f;
      __stcs(out + n * dim + c, v);
    }
  }
}


// 32-byte version (dim % 8 == 0, 32-byte aligned buffers): one lane moves one 32-byte sector with a single
// 256-bit load and a single 256-bit store, and keeps kGatherRows rows in flight (ids of the whole batch first,
// then all row sectors, then the stores) -- the kernel is a random-access stream, so bytes in flight per SM are
// what set the achieved bandwidth.
constexpr int kGatherRows = 4;
__global__ void __launch_bounds__(256)
embed_gather32_kernel(const float* __restrict__ table, int64_t rows, int dim, int lanes_log2,
                      const int32_t* __restrict__ ids, int64_t n_ids, int zero_pad, float* __restrict__ out) {
  // lanes = dim/8 sectors per row, rounded up to a power of two (lanes beyond the row idle)
  const int lanes = 1 << lanes_log2, sectors = dim >> 3;
  const int64_t rows_per_pass = ((int64_t)gridDim.x * blockDim.x) >> lanes_log2;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> lanes_log2;
  const int c = threadIdx.x & (lanes - 1);
  if (c >= sectors) return;
  for (int64_t n0 = g; n0 < n_ids; n0 += rows_per_pass * kGatherRows) {
    int64_t row[kGatherRows];
#pragma unroll
    for (int u = 0; u < kGatherRows; ++u) {
      const int64_t n = n0 + u * rows_per_pass;
      row[u] = n < n_ids ? (int64_t)__ldg(ids + n) - zero_pad : -1;
    }
    float v[kGatherRows][8];
#pragma unroll
    for (int u = 0; u < kGatherRows; ++u) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[u][e] = 0.f;
      if (row[u] >= 0 && row[u] < rows)
        asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[u][0]), "=f"(v[u][1]), "=f"(v[u][2]), "=f"(v[u][3]), "=f"(v[u][4]), "=f"(v[u][5]),
                       "=f"(v[u][6]), "=f"(v[u][7])
                     : "l"(table + row[u] * dim + c * 8));
    }
#pragma unroll
    for (int u = 0; u < kGatherRows; ++u) {
      const int64_t n = n0 + u * rows_per_pass;
      if (n < n_ids)
        asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out + n * dim + c * 8), "f"(v[u][0]),
                     "f"(v[u][1]), "f"(v[u][2]), "f"(v[u][3]), "f"(v[u][4]), "f"(v[u][5]), "f"(v[u][6]), "f"(v[u][7])
                     : "memory");
    }
  }
}

}  // namespace dmt

extern "C" {

int dmt_abi_version(void) { return DMT_ABI_VERSION; }

const char* dmt_last_error(void) { return dmt::g_err; }

#ifndef DMT_BUILD_DIGEST
#define DMT_BUILD_DIGEST "unknown"
#endif
const char* dmt_build_digest(void) { return DMT_BUILD_DIGEST; }

int dmt_device_sm_count(void) {
  int dev = 0, n = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return dmt::cuda_fail(e, "cudaGetDevice");
  e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return dmt::cuda_fail(e, "cudaDeviceGetAttribute");
  return n;
}

int dmt_embed_gather(const float* table, int64_t rows, int32_t dim, const int32_t* ids, int64_t n_ids,
                     int32_t zero_pad, float* out, void* stream) {
  DMT_REQUIRE(table && ids && out, DMT_ERR_INVALID_ARGUMENT, "dmt_embed_gather: null pointer");
  DMT_REQUIRE(rows > 0 && dim > 0 && n_ids >= 0, DMT_ERR_INVALID_ARGUMENT,
              "dmt_embed_gather: rows=%lld dim=%d n_ids=%lld", (long long)rows, dim, (long long)n_ids);
  if (n_ids == 0) return DMT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (dim % 4 == 0) && (((uintptr_t)table | (uintptr_t)out) % 16 == 0);
  const int64_t total = n_ids * (vec ? dim / 4 : dim);
  const int sms = dmt::sm_count_cached();
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sms * 16;   // 8 resident CTAs/SM x 2 waves; grid-stride beyond
  if (blocks > cap) blocks = cap;
  if (dim % 8 == 0 && dim <= 256 && (((uintptr_t)table | (uintptr_t)out) % 32 == 0)) {
    int lg = 0;
    while ((8 << lg) < dim) ++lg;
    const int64_t lanes_total = n_ids << lg;
    int64_t b32 = (lanes_total / dmt::kGatherRows + 255) / 256;
    const int64_t cap32 = (int64_t)sms * 8;          // 8 resident CTAs per SM, grid-stride beyond
    if (b32 > cap32) b32 = cap32;
    if (b32 < 1) b32 = 1;
    dmt::embed_gather32_kernel<<<(unsigned)b32, 256, 0, st>>>(table, rows, dim, lg, ids, n_ids, zero_pad ? 1 : 0, out);
    DMT_CUDA_LAUNCH_CHECK("dmt_embed_gather");
    return DMT_OK;
  }
  if (vec)
    dmt::embed_gather_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(table, rows, dim, ids, n_ids, zero_pad, out);
  else
    dmt::embed_gather_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(table, rows, dim, ids, n_ids, zero_pad, out);
  DMT_CUDA_LAUNCH_CHECK("dmt_embed_gather");
  return DMT_OK;
}

}  // extern "C"
