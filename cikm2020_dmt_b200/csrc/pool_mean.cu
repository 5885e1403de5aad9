// dmt_pool_mean_fwd: weighted-mean pooled embedding lookups (A9 / A11) + dense feature copy.
//
// One CTA per sample; thread c owns output column c of the concatenated pooled block, so the
// D_f threads of one feature read one table row as a single coalesced segment per token and
// the ids / weights are warp-broadcast loads.  HBM-bound: bytes = sum_f nnz_f*(D_f*4 + 4) read
// + B*W*4 written.
#include <cuda_bf16.h>

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


// Grouped version (every row width a multiple of 4): features that share one CSR offsets array (the 6 lookups of one behaviour sequence;
// the 5 single-id item features) are ONE job per sample.  Lane l of the warp owns one float4 of the group's
// concatenated row (feature, float4 index) and walks the sample's tokens in order, eight tokens in flight: no
// shuffle reduction, no idle token lanes for the 8-wide tables, 1/6 of the per-job index math of the
// warp-per-(sample, feature) kernel of round 1 -- that one was issue-bound (68 % of the issue slots busy).
// The sum runs in token order like the scalar kernel (and the oracle).  Output fp32 or bf16 (the bf16 tensor-core
// MMoE reads its input as bf16: the columns are written in that type directly, no conversion pass).
constexpr int kPoolMaxGroups = DMT_MAX_POOL_FEATS;    // worst case: no two features share their offsets
struct PoolGroupArgs {
  dmt_pool_feat f[DMT_MAX_POOL_FEATS];
  uint8_t lane_feat[kPoolMaxGroups][32];              // feature of lane l (0xff: idle)
  uint8_t lane_part[kPoolMaxGroups][32];              // float4 index inside that feature's row
  int32_t batch;
  void* out;
  int64_t out_ld;
  int32_t out_bf16;
};

template <bool WTS>
__device__ __forceinline__ void pool_group_walk(const int32_t* __restrict__ ids, const float* __restrict__ wts,
                                                const float* __restrict__ tab, int64_t rows, int dim, int beg, int end,
                                                float4& num, float& den) {
  // software pipeline: the ids (and weights) of round r + 1 are requested before the rows of round r are consumed,
  // so a round costs ONE dependent-load latency instead of two
  int nrow[8];
  float nw[8];
  auto fetch_ids = [&](int t0) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      nrow[u] = -1;
      nw[u] = 0.f;
      if (t0 + u < end) {
        nrow[u] = __ldg(ids + t0 + u);
        nw[u] = WTS ? __ldg(wts + t0 + u) : 1.0f;
      }
    }
  };
  fetch_ids(beg);
  for (int t0 = beg; t0 < end; t0 += 8) {
    int row[8];
    float w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      row[u] = nrow[u];
      w[u] = nw[u];
    }
    float4 e[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      e[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row[u] >= 0 && row[u] < rows) e[u] = ldg4(tab + (int64_t)row[u] * dim);
    }
    if (t0 + 8 < end) fetch_ids(t0 + 8);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      num.x = fmaf(w[u], e[u].x, num.x);
      num.y = fmaf(w[u], e[u].y, num.y);
      num.z = fmaf(w[u], e[u].z, num.z);
      num.w = fmaf(w[u], e[u].w, num.w);
      den += w[u];
    }
  }
}

__global__ void __launch_bounds__(256, 3) pool_group_kernel(const __grid_constant__ PoolGroupArgs a) {
  const int lane = threadIdx.x & 31, g = blockIdx.y;
  const int fi = a.lane_feat[g][lane];
  if (fi == 0xff) return;
  const int part = a.lane_part[g][lane];
  // this lane's lookup, read once (lanes of different features read different descriptors)
  const int32_t* __restrict__ ids = a.f[fi].ids;
  const int32_t* __restrict__ offs = a.f[fi].offsets;
  const float* __restrict__ wts = a.f[fi].weights;
  const float* __restrict__ tab = a.f[fi].table + part * 4;
  const int64_t rows = a.f[fi].rows;
  const int dim = a.f[fi].dim;
  const int col = a.f[fi].out_col + part * 4;
  for (int b = blockIdx.x * 8 + (threadIdx.x >> 5); b < a.batch; b += gridDim.x * 8) {
    const int beg = __ldg(offs + b), end = __ldg(offs + b + 1);
    float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
    float den = 0.f;
    if (wts) pool_group_walk<true>(ids, wts, tab, rows, dim, beg, end, num, den);
    else pool_group_walk<false>(ids, wts, tab, rows, dim, beg, end, num, den);
    // tf.nn.embedding_lookup_sparse(combiner='mean'): sum(w*row)/sum(w); an absent row comes out as 0
    const float inv = (end > beg) ? 1.0f / den : 0.f;
    if (a.out_bf16) {
      __nv_bfloat16* o = static_cast<__nv_bfloat16*>(a.out) + (int64_t)b * a.out_ld + col;
      o[0] = __float2bfloat16(num.x * inv);
      o[1] = __float2bfloat16(num.y * inv);
      o[2] = __float2bfloat16(num.z * inv);
      o[3] = __float2bfloat16(num.w * inv);
    } else {
      float* o = static_cast<float*>(a.out) + (int64_t)b * a.out_ld + col;
      o[0] = num.x * inv;
      o[1] = num.y * inv;
      o[2] = num.z * inv;
      o[3] = num.w * inv;
    }
  }
}

// features -> groups (same offsets array, <= 32 float4 lanes per group); false when the grouped kernel does not apply
static bool pool_build_groups(int n_feats, const dmt_pool_feat* feats, PoolGroupArgs& a, int* n_groups) {
  int ng = 0;
  int lanes[kPoolMaxGroups];
  const int32_t* goffs[kPoolMaxGroups];
  for (int f = 0; f < n_feats; ++f) {
    const int d = feats[f].dim;
    if (d % 4 != 0 || ((uintptr_t)feats[f].table & 15) != 0 || d / 4 > 32) return false;
    int g = -1;
    for (int k = 0; k < ng; ++k)
      if (goffs[k] == feats[f].offsets && lanes[k] + d / 4 <= 32) { g = k; break; }
    if (g < 0) {
      if (ng == kPoolMaxGroups) return false;
      g = ng++;
      goffs[g] = feats[f].offsets;
      lanes[g] = 0;
      for (int l = 0; l < 32; ++l) a.lane_feat[g][l] = 0xff, a.lane_part[g][l] = 0;
    }
    for (int p = 0; p < d / 4; ++p) {
      a.lane_feat[g][lanes[g]] = (uint8_t)f;
      a.lane_part[g][lanes[g]] = (uint8_t)p;
      ++lanes[g];
    }
  }
  for (int g = ng; g < kPoolMaxGroups; ++g)
    for (int l = 0; l < 32; ++l) a.lane_feat[g][l] = 0xff, a.lane_part[g][l] = 0;
  *n_groups = ng;
  return ng > 0;
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

static int pool_mean_impl(const char* fn, int32_t batch, int32_t n_feats, const dmt_pool_feat* feats, void* out,
                          int64_t out_ld, int out_bf16, void* stream) {
  DMT_REQUIRE(feats && out, DMT_ERR_INVALID_ARGUMENT, "%s: null pointer", fn);
  DMT_REQUIRE(batch >= 0 && n_feats > 0 && n_feats <= DMT_MAX_POOL_FEATS, DMT_ERR_INVALID_ARGUMENT,
              "%s: batch=%d n_feats=%d (max %d)", fn, batch, n_feats, DMT_MAX_POOL_FEATS);
  if (batch == 0) return DMT_OK;
  int col = 0;
  for (int f = 0; f < n_feats; ++f) {
    DMT_REQUIRE(feats[f].table && feats[f].ids && feats[f].offsets && feats[f].dim > 0 && feats[f].rows > 0,
                DMT_ERR_INVALID_ARGUMENT, "%s: feature %d is incomplete", fn, f);
    col += feats[f].dim;
  }
  {
    // features that share their CSR offsets walk the sample's tokens together (pool_group_kernel)
    dmt::PoolGroupArgs ga;
    int ng = 0;
    if (dmt::pool_build_groups(n_feats, feats, ga, &ng)) {
      for (int f = 0; f < n_feats; ++f) ga.f[f] = feats[f];
      ga.batch = batch;
      ga.out = out;
      ga.out_ld = out_ld;
      ga.out_bf16 = out_bf16;
      int gx = (batch + 7) / 8;
      const int cap = (dmt::sm_count_cached() * 16 + ng - 1) / ng;
      if (gx > cap) gx = cap < 1 ? 1 : cap;
      dmt::pool_group_kernel<<<dim3(gx, ng), 256, 0, (cudaStream_t)stream>>>(ga);
      DMT_CUDA_LAUNCH_CHECK("pool_group_kernel");
      return DMT_OK;
    }
  }
  DMT_REQUIRE(!out_bf16, DMT_ERR_UNSUPPORTED_SHAPE,
              "%s: bf16 output needs row widths that are multiples of 4 and 16-byte aligned tables", fn);
  dmt::PoolArgs a;
  col = 0;
  for (int f = 0; f < n_feats; ++f) {
    a.f[f] = feats[f];
    a.col_first[f] = col;
    col += feats[f].dim;
  }
  for (int f = n_feats; f <= DMT_MAX_POOL_FEATS; ++f) a.col_first[f] = col;
  a.n_feats = n_feats;
  a.width = col;
  a.batch = batch;
  a.out = static_cast<float*>(out);
  a.out_ld = out_ld;
  // row widths that are not multiples of 4 (the 5-wide bias tables): one CTA per sample, one thread per column
  const int threads = col >= 1024 ? 1024 : ((col + 31) / 32) * 32;
  dmt::pool_mean_kernel<<<batch, threads, 0, (cudaStream_t)stream>>>(a);
  DMT_CUDA_LAUNCH_CHECK("pool_mean_kernel");
  return DMT_OK;
}

int dmt_pool_mean_fwd(int32_t batch, int32_t n_feats, const dmt_pool_feat* feats, float* out, int64_t out_ld,
                      void* stream) {
  return pool_mean_impl("dmt_pool_mean_fwd", batch, n_feats, feats, out, out_ld, 0, stream);
}

int dmt_pool_mean_fwd_bf16(int32_t batch, int32_t n_feats, const dmt_pool_feat* feats, void* out_bf16, int64_t out_ld,
                           void* stream) {
  return pool_mean_impl("dmt_pool_mean_fwd_bf16", batch, n_feats, feats, out_bf16, out_ld, 1, stream);
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
