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

__global__ void __launch_bounds__(1024) pool_mean_kernel(const __grid_constant__ PoolArgs a) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < a.width; c += blockDim.x) {
    int f = 0;
    while (f + 1 < a.n_feats && c >= a.col_first[f + 1]) ++f;
    const dmt_pool_feat& pf = a.f[f];
    const int j = c - a.col_first[f];
    const int beg = __ldg(pf.offsets + b), end = __ldg(pf.offsets + b + 1);
    float num = 0.f, den = 0.f;
    int t = beg;
    // eight (then four) tokens per trip: the id -> row dependent loads of different tokens overlap (the kernel
    // is latency-bound otherwise); accumulation order stays token order
    for (; t + 8 <= end; t += 8) {
      int64_t row[8];
      float w[8], e[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        row[u] = __ldg(pf.ids + t + u);
        w[u] = pf.weights ? __ldg(pf.weights + t + u) : 1.0f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        e[u] = (row[u] >= 0 && row[u] < pf.rows) ? __ldg(pf.table + row[u] * pf.dim + j) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        num = fmaf(w[u], e[u], num);
        den += w[u];
      }
    }
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


// Throughput version (every dim a power of two in [4, 128]): one WARP per (sample, feature) job.  A row is read as
// dim/4 float4 lanes, so a warp covers 32/(dim/4) tokens per step and keeps up to 8 steps in flight: the ids of a
// whole round are requested first, then all of its rows -- one dependent load pair per round instead of one per
// few tokens -- and every lane has work no matter how long the sample's other features are.  The token lanes are
// reduced in a fixed xor-shuffle order (deterministic).
// one round of STEPS steps: ids (and weights) of the whole round first, then all rows, then the accumulation
template <int STEPS, bool WTS>
__device__ __forceinline__ void pool_round(const dmt_pool_feat& pf, int t0, int end, int TPI, int tl, int v4,
                                           float4& num, float& den) {
  int row[STEPS];
  float w[STEPS];
#pragma unroll
  for (int u = 0; u < STEPS; ++u) {
    const int t = t0 + u * TPI + tl;
    row[u] = -1;
    w[u] = 0.f;
    if (t < end) {
      row[u] = __ldg(pf.ids + t);
      w[u] = WTS ? __ldg(pf.weights + t) : 1.0f;
    }
  }
  float4 e[STEPS];
#pragma unroll
  for (int u = 0; u < STEPS; ++u) {
    e[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row[u] >= 0 && row[u] < pf.rows) e[u] = ldg4(pf.table + (int64_t)row[u] * pf.dim + v4);
  }
#pragma unroll
  for (int u = 0; u < STEPS; ++u) {
    num.x = fmaf(w[u], e[u].x, num.x);
    num.y = fmaf(w[u], e[u].y, num.y);
    num.z = fmaf(w[u], e[u].z, num.z);
    num.w = fmaf(w[u], e[u].w, num.w);
    den += w[u];
  }
}

// grid = (sample groups, features): the feature -- table, dim, lane split -- is uniform per CTA, so its descriptor
// is read once and the only per-job index math is shifts; warps stride over the samples.
__global__ void __launch_bounds__(256) pool_mean_warp_kernel(const __grid_constant__ PoolArgs a) {
  const int lane = threadIdx.x & 31;
  const dmt_pool_feat pf = a.f[blockIdx.y];
  const int V = pf.dim >> 2;                          // float4 lanes per row (power of two)
  const int lgV = 31 - __clz(V);
  const int TPI = 32 >> lgV;                          // tokens per step
  const int tl = lane >> lgV, v4 = (lane & (V - 1)) * 4;
  float* const out = a.out + pf.out_col + v4;
  for (int b = blockIdx.x * 8 + (threadIdx.x >> 5); b < a.batch; b += gridDim.x * 8) {
    const int beg = __ldg(pf.offsets + b), end = __ldg(pf.offsets + b + 1);
    float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
    float den = 0.f;
    int t0 = beg;
    // (warp-uniform branches: the issue slots of steps that have no tokens are not spent)
    if (pf.weights) {
      for (; end - t0 > 2 * TPI; t0 += 8 * TPI) pool_round<8, true>(pf, t0, end, TPI, tl, v4, num, den);
      if (end - t0 > TPI) pool_round<2, true>(pf, t0, end, TPI, tl, v4, num, den);
      else if (end > t0) pool_round<1, true>(pf, t0, end, TPI, tl, v4, num, den);
    } else {
      for (; end - t0 > 2 * TPI; t0 += 8 * TPI) pool_round<8, false>(pf, t0, end, TPI, tl, v4, num, den);
      if (end - t0 > TPI) pool_round<2, false>(pf, t0, end, TPI, tl, v4, num, den);
      else if (end > t0) pool_round<1, false>(pf, t0, end, TPI, tl, v4, num, den);
    }
    for (int o = 16; o >= V; o >>= 1) {               // token lanes: lane bits above the float4 index
      num.x += __shfl_xor_sync(0xffffffffu, num.x, o);
      num.y += __shfl_xor_sync(0xffffffffu, num.y, o);
      num.z += __shfl_xor_sync(0xffffffffu, num.z, o);
      num.w += __shfl_xor_sync(0xffffffffu, num.w, o);
      den += __shfl_xor_sync(0xffffffffu, den, o);
    }
    if (tl == 0) {
      // tf.nn.embedding_lookup_sparse(combiner='mean'): sum(w*row)/sum(w); an absent row comes out as 0
      const float inv = (end > beg) ? 1.0f / den : 0.f;
      float* o = out + (int64_t)b * a.out_ld;
      o[0] = num.x * inv;
      o[1] = num.y * inv;
      o[2] = num.z * inv;
      o[3] = num.w * inv;
    }
  }
}

__global__ void __launch_bounds__(256)
copy_dense_kernel(const float* __restrict__ src, int batch, int dim, float* __restrict__ dst, int64_t ld) {
  // one warp per row, lanes stride the columns (no per-element division); rows are independent
  const int lane = threadIdx.x & 31;
  for (int b = blockIdx.x * 8 + (threadIdx.x >> 5); b < batch; b += gridDim.x * 8) {
    const float* __restrict__ s = src + (int64_t)b * dim;
    float* __restrict__ d = dst + (int64_t)b * ld;
    int c = lane;
    for (; c + 96 < dim; c += 128) {
      const float v0 = __ldcs(s + c), v1 = __ldcs(s + c + 32), v2 = __ldcs(s + c + 64), v3 = __ldcs(s + c + 96);
      d[c] = v0; d[c + 32] = v1; d[c + 64] = v2; d[c + 96] = v3;
    }
    for (; c < dim; c += 32) d[c] = __ldcs(s + c);
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
  // one thread per output column (a second pass over the columns would double the dependent-load chain)
  bool vec = true;                                    // float4 rows: dim a power of two in [4, 128], aligned tables
  for (int f = 0; f < n_feats; ++f) {
    const int d = feats[f].dim;
    vec = vec && d >= 4 && d <= 128 && (d & (d - 1)) == 0 && ((uintptr_t)feats[f].table & 15) == 0;
  }
  if (vec) {
    int gx = (batch + 7) / 8;                          // 8 warps per CTA, one sample per warp and trip
    const int cap = (dmt::sm_count_cached() * 16 + n_feats - 1) / n_feats;
    if (gx > cap) gx = cap < 1 ? 1 : cap;
    dmt::pool_mean_warp_kernel<<<dim3(gx, n_feats), 256, 0, (cudaStream_t)stream>>>(a);
    DMT_CUDA_LAUNCH_CHECK("pool_mean_warp_kernel");
    return DMT_OK;
  }
  const int threads = col >= 1024 ? 1024 : ((col + 31) / 32) * 32;
  dmt::pool_mean_kernel<<<batch, threads, 0, (cudaStream_t)stream>>>(a);
  DMT_CUDA_LAUNCH_CHECK("pool_mean_kernel");
  return DMT_OK;
}

int dmt_copy_dense_features(const float* features, int32_t batch, int32_t dim, float* out, int64_t out_ld,
                            void* stream) {
  DMT_REQUIRE(features && out && batch >= 0 && dim > 0, DMT_ERR_INVALID_ARGUMENT,
              "dmt_copy_dense_features: bad arguments");
  if (batch == 0) return DMT_OK;
  int64_t blocks = ((int64_t)batch + 7) / 8;           // one warp per row
  const int64_t cap = (int64_t)dmt::sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  dmt::copy_dense_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(features, batch, dim, out, out_ld);
  DMT_CUDA_LAUNCH_CHECK("copy_dense_kernel");
  return DMT_OK;
}

}  // extern "C"
