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

}  // namespace dmt

extern "C" {

int dmt_abi_version(void) { return DMT_ABI_VERSION; }

const char* dmt_last_error(void) { return dmt::g_err; }

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
  if (vec)
    dmt::embed_gather_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(table, rows, dim, ids, n_ids, zero_pad, out);
  else
    dmt::embed_gather_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(table, rows, dim, ids, n_ids, zero_pad, out);
  DMT_CUDA_LAUNCH_CHECK("dmt_embed_gather");
  return DMT_OK;
}

}  // extern "C"
