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


// Grouped version (every row width a multiple of 4): features that share one CSR offsets array (the 6 lookups of one
// behaviour sequence; the 5 single-id item features) are walked together, one warp per (sample, group).  A group has
// P float4 "parts" per token (a 32-wide Sku row = 8 parts, five 8-wide rows = 10 parts); the warp processes
// 3 tokens per step (2 / 1 for rows wider than 40 / 64 floats) -- lane = (token slot, part) -- and keeps eight steps
// in flight, ids requested one round ahead of the rows, the next sample's first round during this sample's last.
// Every lane sums the tokens of its slot in token order; the slot sums are added in slot order (fixed order that
// depends on the feature's row width only: bit-identical run to run, in any batch, with any grouping).  Round 1's warp-per-(sample, feature) kernel spent most of its
// issue slots on per-job index math (68 % busy, 39 us at B = 4096); a first grouped version with ONE token per step
// had too few loads in flight (38 us).  Output fp32 or bf16 (the bf16 tensor-core MMoE reads its input as bf16: the
// columns are written in that type directly, no conversion pass).
constexpr int kPoolMaxGroups = DMT_MAX_POOL_FEATS;    // worst case: no two features share their offsets
struct PoolGroupArgs {
  dmt_pool_feat f[DMT_MAX_POOL_FEATS];
  uint8_t part_feat[kPoolMaxGroups][32];              // part p of a token of group g -> feature
  uint8_t part_idx[kPoolMaxGroups][32];               //                              -> float4 index inside its row
  uint8_t n_parts[kPoolMaxGroups];                    // P
  uint8_t n_slots[kPoolMaxGroups];                    // token slots per step (pool_slots of the group's features)
  int32_t batch;
  void* out;
  int64_t out_ld;
  int32_t out_bf16;
};

template <bool WTS>
__device__ __forceinline__ void pool_group_samples(const PoolGroupArgs& a, const int32_t* __restrict__ ids,
                                                   const int32_t* __restrict__ offs, const float* __restrict__ wts,
                                                   const float* __restrict__ tab, int64_t rows, int dim, int col, int P,
                                                   int tps, bool active, int slot) {
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * 8;
  // Software pipeline over tokens AND samples: the ids (and weights) of the next round -- the first round of the NEXT
  // sample after the last round of this one -- are requested before the rows of the current round are consumed, and
  // the next sample's offsets one sample ahead: a round costs ONE dependent-load latency, a new sample none.
  int nrow[8];
  float nw[8];
  auto fetch_ids = [&](int t0, int end) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int t = t0 + u * tps;
      nrow[u] = -1;
      nw[u] = 0.f;
      if (active && t < end) {
        nrow[u] = __ldg(ids + t);
        nw[u] = WTS ? __ldg(wts + t) : 1.0f;
      }
    }
  };
  int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  int beg = 0, end = 0;
  if (b < a.batch) {
    beg = __ldg(offs + b);
    end = __ldg(offs + b + 1);
  }
  fetch_ids(beg + slot, end);
  while (b < a.batch) {
    const int nb = b + stride;
    int nbeg = 0, nend = 0;
    if (nb < a.batch) {
      nbeg = __ldg(offs + nb);
      nend = __ldg(offs + nb + 1);
    }
    float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
    float den = 0.f;
    int t0 = beg + slot;
    do {                                               // (an empty sample runs one round of nothing)
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
      t0 += 8 * tps;
      if (t0 - slot < end) fetch_ids(t0, end);          // (warp-uniform: t0 - slot is the round's first token)
      else fetch_ids(nbeg + slot, nend);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        num.x = fmaf(w[u], e[u].x, num.x);
        num.y = fmaf(w[u], e[u].y, num.y);
        num.z = fmaf(w[u], e[u].z, num.z);
        num.w = fmaf(w[u], e[u].w, num.w);
        den += w[u];
      }
    } while (t0 - slot < end);
    for (int sl = 1; sl < tps; ++sl) {                 // slot sums -> the slot-0 lanes, in slot order
      const int src = lane + sl * P;
      const float vx = __shfl_sync(0xffffffffu, num.x, src), vy = __shfl_sync(0xffffffffu, num.y, src);
      const float vz = __shfl_sync(0xffffffffu, num.z, src), vw = __shfl_sync(0xffffffffu, num.w, src);
      const float vd = __shfl_sync(0xffffffffu, den, src);
      if (slot == 0) {
        num.x += vx; num.y += vy; num.z += vz; num.w += vw;
        den += vd;
      }
    }
    if (active && slot == 0) {
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
    b = nb;
    beg = nbeg;
    end = nend;
  }
}

__global__ void __launch_bounds__(256, 2) pool_group_kernel(const __grid_constant__ PoolGroupArgs a) {
  const int lane = threadIdx.x & 31, g = blockIdx.y;
  const int P = a.n_parts[g], tps = a.n_slots[g];
  const bool active = lane < tps * P;
  const int slot = active ? lane / P : 0, pi = active ? lane - slot * P : 0;
  const int fi = a.part_feat[g][pi], part = a.part_idx[g][pi];
  // this lane's lookup, read once (lanes of different features read different descriptors)
  const int32_t* __restrict__ ids = a.f[fi].ids;
  const int32_t* __restrict__ offs = a.f[fi].offsets;
  const float* __restrict__ wts = a.f[fi].weights;
  const float* __restrict__ tab = a.f[fi].table + part * 4;
  const int64_t rows = a.f[fi].rows;
  const int dim = a.f[fi].dim;
  const int col = a.f[fi].out_col + part * 4;
  // (lanes of one warp may differ in having weights: both instantiations run, each with its lanes)
  const bool any_w = __any_sync(0xffffffffu, wts != nullptr), all_w = __all_sync(0xffffffffu, wts != nullptr || !active);
  if (!any_w) pool_group_samples<false>(a, ids, offs, wts, tab, rows, dim, col, P, tps, active, slot);
  else if (all_w) pool_group_samples<true>(a, ids, offs, wts, tab, rows, dim, col, P, tps, active, slot);
  else {
    pool_group_samples<true>(a, ids, offs, wts, tab, rows, dim, col, P, tps, active && wts != nullptr, slot);
    pool_group_samples<false>(a, ids, offs, wts, tab, rows, dim, col, P, tps, active && wts == nullptr, slot);
  }
}

// Token slots per step of a feature: a function of ITS row width only, so that the order in which a sample's rows
// are summed -- tokens t = slot (mod slots) in token order per slot, then the slots in order -- does not depend on
// which other features share the group (a sample's output is bit-identical in any batch and any grouping).
static inline int pool_slots(int parts) { return parts <= 10 ? 3 : (parts <= 16 ? 2 : 1); }

// features -> groups: same offsets array, same slot count, slots x parts <= 32 lanes; a row of >= 8 float4 parts
// (Sku) is a group of its own.  false when the grouped kernel does not apply.
static bool pool_build_groups(int n_feats, const dmt_pool_feat* feats, PoolGroupArgs& a, int* n_groups) {
  int ng = 0;
  int parts[kPoolMaxGroups];
  const int32_t* goffs[kPoolMaxGroups];
  bool wide[kPoolMaxGroups];
  for (int f = 0; f < n_feats; ++f) {
    const int d = feats[f].dim;
    if (d % 4 != 0 || ((uintptr_t)feats[f].table & 15) != 0 || d / 4 > 32) return false;
    const int np = d / 4, sl = pool_slots(np);
    int g = -1;
    if (np < 8)
      for (int k = 0; k < ng; ++k)
        if (goffs[k] == feats[f].offsets && !wide[k] && a.n_slots[k] == sl && (parts[k] + np) * sl <= 32) { g = k; break; }
    if (g < 0) {
      if (ng == kPoolMaxGroups) return false;
      g = ng++;
      goffs[g] = feats[f].offsets;
      parts[g] = 0;
      wide[g] = np >= 8;
      a.n_slots[g] = (uint8_t)sl;
      for (int l = 0; l < 32; ++l) a.part_feat[g][l] = 0, a.part_idx[g][l] = 0;
    }
    for (int p = 0; p < np; ++p) {
      a.part_feat[g][parts[g]] = (uint8_t)f;
      a.part_idx[g][parts[g]] = (uint8_t)p;
      ++parts[g];
    }
  }
  for (int g = 0; g < kPoolMaxGroups; ++g) {
    a.n_parts[g] = g < ng ? (uint8_t)parts[g] : 1;
    if (g >= ng) a.n_slots[g] = 1;
  }
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
      // ~two waves of 2 CTAs per SM over all groups: every warp walks several samples (the cross-sample prefetch needs
      // a next sample), the cheap single-token groups retire early and make room for the sequence groups
      int gx = (batch + 7) / 8;
      const int cap = (dmt::sm_count_cached() * 4 + ng - 1) / ng;
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
