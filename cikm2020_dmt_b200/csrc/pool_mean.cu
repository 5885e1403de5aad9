// dmt_pool_mean_fwd: weighted-mean pooled embedding lookups (A9 / A11) + dense feature copy.
//
// One CTA per sample; thread c owns output column c of the concatenated pooled block, so the
// D_f threads of one feature read one table row as a single coalesced segment per token and
// the ids / weights are warp-broadcast loads.  HBM-bound: bytes = sum_f nnz_f*(D_f*4 + 4) read
// + B*W*4 written.
#include "dmt_common.cuh"

namespace dmt {

struct PoolArgs {
  dmt_pool_feat f[DMT_MAX_POOL_FEATS];
  int32_t col_first[DMT_MAX_POOL_FEATS + 1];  // prefix sum of dims (thread -> feature map)
  int32_t n_feats;
  int32_t width;      // sum of dims
  int32_t batch;
  float* out;
  int64_t out_ld;
};

__global__ void __launch_bounds__(256) pool_mean_kernel(const __grid_constant__ PoolArgs a) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < a.width; c += blockDim.x) {
    int f = 0;
    while (f + 1 < a.n_feats && c >= a.col_first[f + 1]) ++f;
    const dmt_pool_feat& pf = a.f[f];
    const int j = c - a.col_first[f];
    const int beg = __ldg(pf.offsets + b), end = __ldg(pf.offsets + b + 1);
    float num = 0.f, den = 0.f;
    int t = beg;
    // four tokens per trip: the id -> row dependent loads of different tokens overlap (the kernel is
    // latency-bound otherwise); accumulation order stays token order
    for (; t + 4 <= end; t += 4) {
      int64_t row[4];
      float w[4], e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        row[u] = __ldg(pf.ids + t + u);
        w[u] = pf.weights ? __ldg(pf.weights + t + u) : 1.0f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        e[u] = (row[u] >= 0 && row[u] < pf.rows) ? __ldg(pf.table + row[u] * pf.dim + j) : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        num = fmaf(w[u], e[u], num);
        den += w[u];
      }
    }
    for (; t < end; ++t) {
      const int64_t row = __ldg(pf.ids + t);
      const float w = pf.weights ? __ldg(pf.weights + t) : 1.0f;
      const float e = (row >= 0 && row < pf.rows) ? __ldg(pf.table + row * pf.dim + j) : 0.f;
      num = fmaf(w, e, num);
      den += w;
    }
    // tf.nn.embedding_lookup_sparse(combiner='mean'): sum(w*row)/sum(w); an empty row is absent
    // from the SparseTensor and comes out as 0
    a.out[(int64_t)b * a.out_ld + pf.out_col + j] = (end > beg) ? num / den : 0.f;
  }
}

__global__ void __launch_bounds__(256)
copy_dense_kernel(const float* __restrict__ src, int batch, int dim, float* __restrict__ dst, int64_t ld) {
  const int64_t total = (int64_t)batch * dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / dim;
    const int c = (int)(i - b * dim);
    dst[b * ld + c] = __ldcs(src + i);
  }
}

}  // namespace dmt

extern "C" {

int dmt_pool_mean_fwd(int32_t batch, int32_t n_feats, const dmt_pool_feat* feats, float* out, int64_t out_ld,
                      void* stream) {
  DMT_REQUIRE(feats && out, DMT_ERR_INVALID_ARGUMENT, "dmt_pool_mean_fwd: null pointer");
  DMT_REQUIRE(batch >= 0 && n_feats > 0 && n_feats <= DMT_MAX_POOL_FEATS, DMT_ERR_INVALID_ARGUMENT,
              "dmt_pool_mean_fwd: batch=%d n_feats=%d (max %d)", batch, n_feats, DMT_MAX_POOL_FEATS);
  if (batch == 0) return DMT_OK;
  dmt::PoolArgs a;
  int col = 0;
  for (int f = 0; f < n_feats; ++f) {
    DMT_REQUIRE(feats[f].table && feats[f].ids && feats[f].offsets && feats[f].dim > 0 && feats[f].rows > 0,
                DMT_ERR_INVALID_ARGUMENT, "dmt_pool_mean_fwd: feature %d is incomplete", f);
    a.f[f] = feats[f];
    a.col_first[f] = col;
    col += feats[f].dim;
  }
  for (int f = n_feats; f <= DMT_MAX_POOL_FEATS; ++f) a.col_first[f] = col;
  a.n_feats = n_feats;
  a.width = col;
  a.batch = batch;
  a.out = out;
  a.out_ld = out_ld;
  const int threads = col >= 256 ? 256 : ((col + 31) / 32) * 32;
  dmt::pool_mean_kernel<<<batch, threads, 0, (cudaStream_t)stream>>>(a);
  DMT_CUDA_LAUNCH_CHECK("pool_mean_kernel");
  return DMT_OK;
}

int dmt_copy_dense_features(const float* features, int32_t batch, int32_t dim, float* out, int64_t out_ld,
                            void* stream) {
  DMT_REQUIRE(features && out && batch >= 0 && dim > 0, DMT_ERR_INVALID_ARGUMENT,
              "dmt_copy_dense_features: bad arguments");
  if (batch == 0) return DMT_OK;
  const int64_t total = (int64_t)batch * dim;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)dmt::sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  dmt::copy_dense_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(features, batch, dim, out, out_ld);
  DMT_CUDA_LAUNCH_CHECK("copy_dense_kernel");
  return DMT_OK;
}

}  // extern "C"
