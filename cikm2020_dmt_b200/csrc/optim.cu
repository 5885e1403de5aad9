// K9 + K10: segmented sparse-gradient scatter-add fused with TF-1 Adam, and the dense Adam pass.
//
// tf.train.AdamOptimizer (inference_mlp.py:272-273) has DENSE semantics even for embedding tables: the
// reference densifies IndexedSlices in average_gradients (run_dnn.py:63-72), so every row decays its
// moments and moves every step.  The update is
//     lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//     theta -= lr_t * m / (sqrt(v) + eps)
//
// Tables never see a dense gradient here.  Per table and step:
//   1. dmt_embed_grad_expand      every lookup of the step (sequence tokens, target items, pooled features)
//                                 becomes a (row key, gradient-row reference, scale) triple;
//   2. the caller sorts the keys (stable; any device sort -- plumbing);
//   3. dmt_embed_adam_sorted      one warp per run of equal keys sums the referenced gradient rows in sorted
//                                 order (deterministic, no atomics) and applies the Adam update to that row,
//                                 marking it touched;
//   4. dmt_adam_rows_untouched    one streaming pass applies the g = 0 update to every other row (pure HBM:
//                                 6 * V * D * 4 bytes) and clears the marks.
#include "dmt_common.cuh"

namespace dmt {

struct AdamScalars {
  float lr_t, b1, b2, eps, one_minus_b1, one_minus_b2;
};

static AdamScalars adam_scalars(const dmt_adam_cfg* c) {
  AdamScalars s;
  const double t = (double)c->step;
  s.lr_t = (float)((double)c->lr * sqrt(1.0 - pow((double)c->beta2, t)) / (1.0 - pow((double)c->beta1, t)));
  s.b1 = c->beta1;
  s.b2 = c->beta2;
  s.eps = c->epsilon;
  s.one_minus_b1 = 1.0f - c->beta1;
  s.one_minus_b2 = 1.0f - c->beta2;
  return s;
}

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const AdamScalars& s) {
  m = fmaf(s.b1, m, s.one_minus_b1 * g);
  v = fmaf(s.b2, v, s.one_minus_b2 * g * g);
  p -= s.lr_t * m / (sqrtf(v) + s.eps);
}

__global__ void __launch_bounds__(256) adam_dense_kernel(float* __restrict__ p, float* __restrict__ m,
                                                         float* __restrict__ v, const float* __restrict__ g,
                                                         int64_t n, float gscale, AdamScalars s) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      float4 pv = *reinterpret_cast<float4*>(p + i), mv = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      const float4 gv = ld_stream4(g + i);
      adam_update(pv.x, mv.x, vv.x, gv.x * gscale, s);
      adam_update(pv.y, mv.y, vv.y, gv.y * gscale, s);
      adam_update(pv.z, mv.z, vv.z, gv.z * gscale, s);
      adam_update(pv.w, mv.w, vv.w, gv.w * gscale, s);
      *reinterpret_cast<float4*>(p + i) = pv;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (int64_t j = i; j < n; ++j) adam_update(p[j], m[j], v[j], g[j] * gscale, s);
    }
  }
}

struct ExpandArgs {
  dmt_grad_source src[DMT_MAX_GRAD_SOURCES];
  int64_t base[DMT_MAX_GRAD_SOURCES + 1];   // prefix sum of src[i].n
  int32_t n_sources;
  int64_t rows;
  int32_t* keys;
  int64_t* refs;
  float* scale;
};

// ref = (source << 40) | row of the gradient matrix that holds this lookup's gradient
__global__ void __launch_bounds__(256) grad_expand_kernel(const __grid_constant__ ExpandArgs a) {
  const int64_t total = a.base[a.n_sources];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int s = 0;
    while (s + 1 < a.n_sources && i >= a.base[s + 1]) ++s;
    const dmt_grad_source& g = a.src[s];
    const int64_t j = i - a.base[s];
    const int64_t row = (int64_t)__ldg(g.ids + j) + g.id_offset;
    int64_t grow = j;
    float sc = 1.0f;
    if (g.offsets) {   // gradient rows are per SAMPLE: find the sample that owns lookup j
      int lo = 0, hi = g.batch;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(g.offsets + mid) <= j) lo = mid; else hi = mid;
      }
      grow = lo;
      if (g.mean) {    // tf.nn.embedding_lookup_sparse(combiner='mean'): d/d row = w_j / sum_w
        const int beg = __ldg(g.offsets + lo), end = __ldg(g.offsets + lo + 1);
        float sum = 0.f;
        if (g.weights) for (int t = beg; t < end; ++t) sum += __ldg(g.weights + t);
        else sum = (float)(end - beg);
        sc = (g.weights ? __ldg(g.weights + j) : 1.0f) / sum;
      }
    }
    const bool ok = row >= 0 && row < a.rows;
    a.keys[i] = ok ? (int32_t)row : INT32_MAX;   // invalid / zero-pad index 0 sorts to the end
    a.refs[i] = ((int64_t)s << 40) | grow;
    a.scale[i] = sc;
  }
}

struct SortedAdamArgs {
  dmt_grad_source src[DMT_MAX_GRAD_SOURCES];
  float *table, *m, *v;
  int64_t rows;
  int32_t dim;
  const int32_t* keys;      // sorted ascending
  const int64_t* perm;      // sorted position -> expanded position
  const int64_t* refs;      // expanded position -> (source, gradient row)
  const float* scale;
  int64_t n;
  uint8_t* touched;
  float gscale;
  AdamScalars s;
};

// One warp per position; the warp whose position starts a run of equal keys owns that row.
__global__ void __launch_bounds__(256) adam_sorted_kernel(const __grid_constant__ SortedAdamArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= a.n) return;
  const int32_t key = __ldg(a.keys + w);
  if (key == INT32_MAX) return;
  if (w > 0 && __ldg(a.keys + w - 1) == key) return;   // not the head of its run
  const int D = a.dim;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};                  // columns lane, lane+32, ... (D <= 128)
  for (int64_t i = w; i < a.n && __ldg(a.keys + i) == key; ++i) {
    const int64_t e = __ldg(a.perm + i);
    const int64_t ref = __ldg(a.refs + e);
    const dmt_grad_source& g = a.src[ref >> 40];
    const float sc = __ldg(a.scale + e);
    const float* gr = g.grad + (ref & 0xFFFFFFFFFFll) * g.grad_ld + g.grad_col;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (lane + 32 * c < D) acc[c] = fmaf(sc, __ldg(gr + lane + 32 * c), acc[c]);
  }
  const int64_t base = (int64_t)key * D;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int col = lane + 32 * c;
    if (col < D) {
      float p = a.table[base + col], m = a.m[base + col], v = a.v[base + col];
      adam_update(p, m, v, acc[c] * a.gscale, a.s);
      a.table[base + col] = p;
      a.m[base + col] = m;
      a.v[base + col] = v;
    }
  }
  if (lane == 0) a.touched[key] = 1;
}

__global__ void __launch_bounds__(256) adam_untouched_kernel(float* __restrict__ table, float* __restrict__ m,
                                                             float* __restrict__ v, int64_t rows, int dim,
                                                             uint8_t* __restrict__ touched, AdamScalars s) {
  const int64_t total = rows * dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dim;
    if (touched[r]) continue;
    float p = table[i], mm = m[i], vv = v[i];
    adam_update(p, mm, vv, 0.f, s);
    table[i] = p;
    m[i] = mm;
    v[i] = vv;
  }
}

}  // namespace dmt

extern "C" {

int dmt_adam_dense(const dmt_adam_cfg* cfg, float* param, float* m, float* v, const float* grad, int64_t n,
                   float grad_scale, void* stream) {
  DMT_REQUIRE(cfg && param && m && v && grad, DMT_ERR_INVALID_ARGUMENT, "dmt_adam_dense: null pointer");
  DMT_REQUIRE(cfg->step >= 1 && n >= 0, DMT_ERR_INVALID_ARGUMENT, "dmt_adam_dense: step=%d n=%lld", cfg->step,
              (long long)n);
  DMT_REQUIRE((((uintptr_t)param | (uintptr_t)m | (uintptr_t)v | (uintptr_t)grad) & 15) == 0, DMT_ERR_INVALID_ARGUMENT,
              "dmt_adam_dense: buffers must be 16-byte aligned");
  if (n == 0) return DMT_OK;
  int64_t blocks = (n / 4 + 255) / 256 + 1;
  const int64_t cap = (int64_t)dmt::sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  dmt::adam_dense_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, m, v, grad, n, grad_scale,
                                                                            dmt::adam_scalars(cfg));
  DMT_CUDA_LAUNCH_CHECK("adam_dense_kernel");
  return DMT_OK;
}

int dmt_embed_grad_expand(int32_t n_sources, const dmt_grad_source* sources, int64_t rows, int32_t* keys,
                          int64_t* refs, float* scale, void* stream) {
  DMT_REQUIRE(sources && keys && refs && scale, DMT_ERR_INVALID_ARGUMENT, "dmt_embed_grad_expand: null pointer");
  DMT_REQUIRE(n_sources > 0 && n_sources <= DMT_MAX_GRAD_SOURCES, DMT_ERR_INVALID_ARGUMENT,
              "dmt_embed_grad_expand: n_sources=%d (max %d)", n_sources, DMT_MAX_GRAD_SOURCES);
  dmt::ExpandArgs a;
  a.base[0] = 0;
  for (int s = 0; s < n_sources; ++s) {
    DMT_REQUIRE(sources[s].ids && sources[s].n >= 0, DMT_ERR_INVALID_ARGUMENT, "dmt_embed_grad_expand: source %d", s);
    a.src[s] = sources[s];
    a.base[s + 1] = a.base[s] + sources[s].n;
  }
  a.n_sources = n_sources;
  a.rows = rows;
  a.keys = keys;
  a.refs = refs;
  a.scale = scale;
  const int64_t total = a.base[n_sources];
  if (total == 0) return DMT_OK;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)dmt::sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  dmt::grad_expand_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  DMT_CUDA_LAUNCH_CHECK("grad_expand_kernel");
  return DMT_OK;
}

int dmt_embed_adam_sorted(const dmt_adam_cfg* cfg, float* table, float* m, float* v, int64_t rows, int32_t dim,
                          int32_t n_sources, const dmt_grad_source* sources, const int32_t* sorted_keys,
                          const int64_t* perm, const int64_t* refs, const float* scale, int64_t n, float grad_scale,
                          uint8_t* touched, void* stream) {
  DMT_REQUIRE(cfg && table && m && v && sources && sorted_keys && perm && refs && scale && touched,
              DMT_ERR_INVALID_ARGUMENT, "dmt_embed_adam_sorted: null pointer");
  DMT_REQUIRE(dim > 0 && dim <= 128 && n_sources > 0 && n_sources <= DMT_MAX_GRAD_SOURCES && cfg->step >= 1,
              DMT_ERR_UNSUPPORTED_SHAPE, "dmt_embed_adam_sorted: dim=%d n_sources=%d step=%d", dim, n_sources,
              cfg->step);
  if (n == 0) return DMT_OK;
  dmt::SortedAdamArgs a;
  for (int s = 0; s < n_sources; ++s) a.src[s] = sources[s];
  a.table = table; a.m = m; a.v = v;
  a.rows = rows; a.dim = dim;
  a.keys = sorted_keys; a.perm = perm; a.refs = refs; a.scale = scale;
  a.n = n; a.touched = touched; a.gscale = grad_scale;
  a.s = dmt::adam_scalars(cfg);
  const int64_t blocks = (n * 32 + 255) / 256;
  dmt::adam_sorted_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  DMT_CUDA_LAUNCH_CHECK("adam_sorted_kernel");
  return DMT_OK;
}

int dmt_adam_rows_untouched(const dmt_adam_cfg* cfg, float* table, float* m, float* v, int64_t rows, int32_t dim,
                            uint8_t* touched, void* stream) {
  DMT_REQUIRE(cfg && table && m && v && touched && cfg->step >= 1, DMT_ERR_INVALID_ARGUMENT,
              "dmt_adam_rows_untouched: bad arguments");
  if (rows == 0) return DMT_OK;
  int64_t blocks = (rows * dim + 255) / 256;
  const int64_t cap = (int64_t)dmt::sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  dmt::adam_untouched_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(table, m, v, rows, dim, touched,
                                                                                dmt::adam_scalars(cfg));
  DMT_CUDA_LAUNCH_CHECK("adam_untouched_kernel");
  cudaError_t e = cudaMemsetAsync(touched, 0, (size_t)rows, (cudaStream_t)stream);
  if (e != cudaSuccess) return dmt::cuda_fail(e, "cudaMemsetAsync(touched)");
  return DMT_OK;
}

}  // extern "C"
