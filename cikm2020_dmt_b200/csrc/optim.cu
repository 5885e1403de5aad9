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
//   3. dmt_embed_adam_sorted      chunked two-pass segmented reduction of the referenced gradient rows in
//                                 sorted order (deterministic, no atomics, hot rows do not serialise) + the
//                                 Adam update of each touched row;
//   4. dmt_adam_rows_untouched    one streaming pass applies the g = 0 update to every other row (pure HBM:
//                                 6 * V * D * 4 bytes) and clears the marks.
#include "dmt_common.cuh"

namespace dmt {

struct AdamScalars {
  float lr_t, b1, b2, eps, one_minus_b1, one_minus_b2;
  int kind;       // DMT_OPT_ADAM | DMT_OPT_SGD | DMT_OPT_ADAGRAD
};

static AdamScalars adam_scalars(const dmt_adam_cfg* c) {
  AdamScalars s;
  const double t = (double)c->step;
  s.kind = c->kind;
  s.lr_t = c->kind == DMT_OPT_ADAM
               ? (float)((double)c->lr * sqrt(1.0 - pow((double)c->beta2, t)) / (1.0 - pow((double)c->beta1, t)))
               : c->lr;
  s.b1 = c->beta1;
  s.b2 = c->beta2;
  s.eps = c->epsilon;
  s.one_minus_b1 = 1.0f - c->beta1;
  s.one_minus_b2 = 1.0f - c->beta2;
  return s;
}

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const AdamScalars& s) {
  if (s.kind == DMT_OPT_SGD) {               // tf.train.GradientDescentOptimizer
    p -= s.lr_t * g;
    return;
  }
  if (s.kind == DMT_OPT_ADAGRAD) {           // tf.train.AdagradOptimizer: the accumulator lives in m
    m = fmaf(g, g, m);
    p -= s.lr_t * g / sqrtf(m);
    return;
  }
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
  float* dense_out;         // non-null: write the summed gradient row here instead of applying Adam
  float* carry;             // [chunks][2][dim] partial sums of the sub-runs that cross a chunk boundary
};

// Deterministic segmented reduction over the sorted lookups, robust to hot rows (a 23-row time table or the
// 302 OOV bucket rows of Sku receive 10^4..10^5 lookups each): the sorted positions are cut into fixed chunks
// of kChunk; pass 1 (one warp per chunk) sums every sub-run of equal keys inside its chunk in sorted order --
// a run that lies inside one chunk is finished on the spot, a sub-run that continues from / into a neighbour
// chunk is parked in `carry`; pass 2 (the warp of the chunk where such a run starts) adds the parked partials
// chunk by chunk.  Fixed chunking + fixed order => run-to-run identical sums, no atomics; the serial depth of a
// run of length r is kChunk + r / kChunk instead of r.
constexpr int kChunk = 128;
constexpr int kMaxCols = 4;   // D <= 128: columns lane, lane+32, ...

__device__ __forceinline__ void finish_row(const SortedAdamArgs& a, int32_t key, const float (&acc)[kMaxCols], int lane) {
  const int D = a.dim;
  const int64_t base = (int64_t)key * D;
  if (a.dense_out) {
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c)
      if (lane + 32 * c < D) a.dense_out[base + lane + 32 * c] = acc[c] * a.gscale;
    return;
  }
#pragma unroll
  for (int c = 0; c < kMaxCols; ++c) {
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

__global__ void __launch_bounds__(256) adam_sorted_pass1_kernel(const __grid_constant__ SortedAdamArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t chunk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t beg = chunk * kChunk;
  if (beg >= a.n) return;
  const int64_t end = min(a.n, beg + (int64_t)kChunk);
  const int D = a.dim;
  const int32_t prev_key = beg > 0 ? __ldg(a.keys + beg - 1) : -1;
  const int32_t next_key = end < a.n ? __ldg(a.keys + end) : -1;
  float acc[kMaxCols] = {0.f, 0.f, 0.f, 0.f};
  int32_t cur = __ldg(a.keys + beg);
  bool starts_here = cur != prev_key;
  for (int64_t i0 = beg; i0 < end; i0 += 32) {
    // lanes fetch 32 (key, gradient-row address, scale) triples at once, then the warp walks them in order
    const int64_t i = i0 + lane;
    int32_t k = -1;
    const float* gr = nullptr;
    float sc = 0.f;
    if (i < end) {
      k = __ldg(a.keys + i);
      const int64_t e = __ldg(a.perm + i);
      const int64_t ref = __ldg(a.refs + e);
      const dmt_grad_source& g = a.src[ref >> 40];
      sc = __ldg(a.scale + e);
      gr = g.grad + (ref & 0xFFFFFFFFFFll) * g.grad_ld + g.grad_col;
    }
    const int cnt = (int)min((int64_t)32, end - i0);
    // four lookups per trip: their gradient-row loads are independent and issued together (the walk itself is
    // sequential because the sums must be formed in sorted order)
    for (int j0 = 0; j0 < cnt; j0 += 4) {
      int32_t kq[4];
      float sq[4], gq[4][kMaxCols];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = min(j0 + u, 31);
        kq[u] = __shfl_sync(0xffffffffu, k, j);
        sq[u] = __shfl_sync(0xffffffffu, sc, j);
        const float* grj = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)gr, j));
        const bool live = j0 + u < cnt && kq[u] != INT32_MAX;
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c)
          gq[u][c] = (live && lane + 32 * c < D) ? __ldg(grj + lane + 32 * c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j0 + u >= cnt) break;
        const int32_t kj = kq[u];
        if (kj != cur) {
          if (cur != INT32_MAX) {
            if (starts_here) {
              finish_row(a, cur, acc, lane);      // the run began and ended inside this chunk
            } else {
#pragma unroll
              for (int c = 0; c < kMaxCols; ++c)
                if (lane + 32 * c < D) a.carry[(chunk * 2 + 0) * D + lane + 32 * c] = acc[c];
            }
          }
#pragma unroll
          for (int c = 0; c < kMaxCols; ++c) acc[c] = 0.f;
          cur = kj;
          starts_here = true;
        }
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c) acc[c] = fmaf(sq[u], gq[u][c], acc[c]);
      }
    }
  }
  if (cur == INT32_MAX) return;
  const bool ends_here = cur != next_key;
  if (starts_here && ends_here) {
    finish_row(a, cur, acc, lane);
  } else {
    const int slot = starts_here ? 1 : 0;   // 1: head of a run that continues; 0: continuation (maybe whole chunk)
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c)
      if (lane + 32 * c < D) a.carry[(chunk * 2 + slot) * D + lane + 32 * c] = acc[c];
  }
}

__global__ void __launch_bounds__(256) adam_sorted_pass2_kernel(const __grid_constant__ SortedAdamArgs a) {
  const int lane = threadIdx.x & 31;
  int64_t chunk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t beg = chunk * kChunk;
  if (beg >= a.n) return;
  int64_t end = min(a.n, beg + (int64_t)kChunk);
  if (end >= a.n) return;                                     // nothing continues past the last chunk
  const int32_t key = __ldg(a.keys + end - 1);
  if (key == INT32_MAX || __ldg(a.keys + end) != key) return;  // the trailing run ends here
  // this chunk owns the run iff the run starts inside it
  if (__ldg(a.keys + beg) == key && beg > 0 && __ldg(a.keys + beg - 1) == key) return;
  const int D = a.dim;
  float acc[kMaxCols];
#pragma unroll
  for (int c = 0; c < kMaxCols; ++c) acc[c] = lane + 32 * c < D ? a.carry[(chunk * 2 + 1) * D + lane + 32 * c] : 0.f;
  for (;;) {
    ++chunk;
    const int64_t b2 = chunk * kChunk;
    const int64_t e2 = min(a.n, b2 + (int64_t)kChunk);
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c)
      if (lane + 32 * c < D) acc[c] += a.carry[(chunk * 2 + 0) * D + lane + 32 * c];
    if (e2 >= a.n || __ldg(a.keys + e2 - 1) != key || __ldg(a.keys + e2) != key) break;
  }
  finish_row(a, key, acc, lane);
}

// out[key, :] = scale * gradient row, for lookups whose keys are UNIQUE (the compact table of a row-sharded
// embedding: one compact row per lookup).  One warp per lookup.
struct ScatterRowsArgs {
  dmt_grad_source src[DMT_MAX_GRAD_SOURCES];
  const int32_t* keys;
  const int64_t* refs;
  const float* scale;
  int64_t n;
  int32_t dim;
  float* out;
};

__global__ void __launch_bounds__(256) grad_scatter_rows_kernel(const __grid_constant__ ScatterRowsArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= a.n) return;
  const int32_t key = __ldg(a.keys + w);
  if (key == INT32_MAX) return;
  const int64_t ref = __ldg(a.refs + w);
  const dmt_grad_source& g = a.src[ref >> 40];
  const float sc = __ldg(a.scale + w);
  const float* gr = g.grad + (ref & 0xFFFFFFFFFFll) * g.grad_ld + g.grad_col;
  for (int c = lane; c < a.dim; c += 32) a.out[(int64_t)key * a.dim + c] = sc * __ldg(gr + c);
}

// VEC floats per thread (4 when dim % 4 == 0): one 16-byte load/store per array, one `touched` byte per chunk.
template <int VEC>
__global__ void __launch_bounds__(256) adam_untouched_kernel(float* __restrict__ table, float* __restrict__ m,
                                                             float* __restrict__ v, int64_t rows, int dim,
                                                             const uint8_t* __restrict__ touched, AdamScalars s) {
  const int chunks = dim / VEC;
  const int64_t total = rows * chunks;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = total < (1ll << 32) ? (int64_t)((uint32_t)i / (uint32_t)chunks) : i / chunks;
    if (touched[r]) continue;
    if (VEC == 4) {
      float4 p = *reinterpret_cast<float4*>(table + i * 4), mm = *reinterpret_cast<float4*>(m + i * 4);
      float4 vv = *reinterpret_cast<float4*>(v + i * 4);
      adam_update(p.x, mm.x, vv.x, 0.f, s);
      adam_update(p.y, mm.y, vv.y, 0.f, s);
      adam_update(p.z, mm.z, vv.z, 0.f, s);
      adam_update(p.w, mm.w, vv.w, 0.f, s);
      *reinterpret_cast<float4*>(table + i * 4) = p;
      *reinterpret_cast<float4*>(m + i * 4) = mm;
      *reinterpret_cast<float4*>(v + i * 4) = vv;
    } else {
      float p = table[i], mm = m[i], vv = v[i];
      adam_update(p, mm, vv, 0.f, s);
      table[i] = p;
      m[i] = mm;
      v[i] = vv;
    }
  }
}

// =====================================================================================================================
// Multi-table variants: ALL embedding tables of a step in one expand / one key sort / one segmented-Adam pair / one
// untouched-rows pass (ten tables used to mean ten rounds of five launches plus ten device sorts -- the launch and
// allocator overhead of that loop was a third of the Adam stage).  Keys carry the table: (table << 24) | row.
constexpr int kMultiRowBits = 24;
constexpr int kMaxMultiSources = 64;

struct ExpandMultiArgs {
  dmt_grad_source src[kMaxMultiSources];
  int64_t base[kMaxMultiSources + 1];
  int8_t tid[kMaxMultiSources];
  int64_t rows[DMT_MAX_ADAM_TABLES];
  int32_t n_sources;
  int32_t* keys;
  int64_t* refs;
  float* scale;
};

__global__ void __launch_bounds__(256) grad_expand_multi_kernel(const __grid_constant__ ExpandMultiArgs a) {
  const int64_t total = a.base[a.n_sources];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int lo_s = 0, hi_s = a.n_sources;          // last source whose base <= i
    while (hi_s - lo_s > 1) {
      const int mid = (lo_s + hi_s) >> 1;
      if (a.base[mid] <= i) lo_s = mid; else hi_s = mid;
    }
    const int s = lo_s;
    const dmt_grad_source& g = a.src[s];
    const int64_t j = i - a.base[s];
    const int64_t row = (int64_t)__ldg(g.ids + j) + g.id_offset;
    int64_t grow = j;
    float sc = 1.0f;
    if (g.offsets) {
      int lo = 0, hi = g.batch;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(g.offsets + mid) <= j) lo = mid; else hi = mid;
      }
      grow = lo;
      if (g.mean) {
        const int beg = __ldg(g.offsets + lo), end = __ldg(g.offsets + lo + 1);
        float sum = 0.f;
        if (g.weights) for (int t = beg; t < end; ++t) sum += __ldg(g.weights + t);
        else sum = (float)(end - beg);
        sc = (g.weights ? __ldg(g.weights + j) : 1.0f) / sum;
      }
    }
    const int t = a.tid[s];
    const bool ok = row >= 0 && row < a.rows[t];
    a.keys[i] = ok ? (int32_t)(((int64_t)t << kMultiRowBits) | row) : INT32_MAX;
    a.refs[i] = ((int64_t)s << 40) | grow;
    a.scale[i] = sc;
  }
}

struct SortedMultiArgs {
  dmt_grad_source src[kMaxMultiSources];
  dmt_adam_table tab[DMT_MAX_ADAM_TABLES];
  const int32_t* keys;
  const int64_t* perm;
  const int64_t* refs;
  const float* scale;
  int64_t n;
  float gscale;
  AdamScalars s;
  float* carry;             // [chunks][2][kMultiMaxDim]
};
constexpr int kMultiMaxDim = 128;

__device__ __forceinline__ void finish_row_multi(const SortedMultiArgs& a, int32_t key, const float (&acc)[kMaxCols], int lane) {
  const dmt_adam_table& t = a.tab[key >> kMultiRowBits];
  const int D = t.dim;
  const int64_t row = key & ((1 << kMultiRowBits) - 1);
  const int64_t base = row * D;
  if (t.dense_out) {
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c)
      if (lane + 32 * c < D) t.dense_out[base + lane + 32 * c] = acc[c] * a.gscale;
    return;
  }
#pragma unroll
  for (int c = 0; c < kMaxCols; ++c) {
    const int col = lane + 32 * c;
    if (col < D) {
      float p = t.table[base + col], m = t.m[base + col], v = t.v[base + col];
      adam_update(p, m, v, acc[c] * a.gscale, a.s);
      t.table[base + col] = p;
      t.m[base + col] = m;
      t.v[base + col] = v;
    }
  }
  if (lane == 0) t.touched[row] = 1;
}

// same chunked two-pass segmented reduction as adam_sorted_pass1/2, the table taken from the key of each run
__global__ void __launch_bounds__(256) adam_sorted_multi_pass1_kernel(const __grid_constant__ SortedMultiArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t chunk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t beg = chunk * kChunk;
  if (beg >= a.n) return;
  const int64_t end = min(a.n, beg + (int64_t)kChunk);
  const int32_t prev_key = beg > 0 ? __ldg(a.keys + beg - 1) : -1;
  const int32_t next_key = end < a.n ? __ldg(a.keys + end) : -1;
  float acc[kMaxCols] = {0.f, 0.f, 0.f, 0.f};
  int32_t cur = __ldg(a.keys + beg);
  bool starts_here = cur != prev_key;
  for (int64_t i0 = beg; i0 < end; i0 += 32) {
    const int64_t i = i0 + lane;
    int32_t k = -1;
    const float* gr = nullptr;
    float sc = 0.f;
    if (i < end) {
      k = __ldg(a.keys + i);
      const int64_t e = __ldg(a.perm + i);
      const int64_t ref = __ldg(a.refs + e);
      const dmt_grad_source& g = a.src[ref >> 40];
      sc = __ldg(a.scale + e);
      gr = g.grad + (ref & 0xFFFFFFFFFFll) * g.grad_ld + g.grad_col;
    }
    const int cnt = (int)min((int64_t)32, end - i0);
    for (int j0 = 0; j0 < cnt; j0 += 4) {
      int32_t kq[4];
      float sq[4], gq[4][kMaxCols];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = min(j0 + u, 31);
        kq[u] = __shfl_sync(0xffffffffu, k, j);
        sq[u] = __shfl_sync(0xffffffffu, sc, j);
        const float* grj = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)gr, j));
        const bool live = j0 + u < cnt && kq[u] != INT32_MAX;
        const int D = live ? a.tab[kq[u] >> kMultiRowBits].dim : 0;
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c) gq[u][c] = (lane + 32 * c < D) ? __ldg(grj + lane + 32 * c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j0 + u >= cnt) break;
        const int32_t kj = kq[u];
        if (kj != cur) {
          if (cur != INT32_MAX) {
            if (starts_here) {
              finish_row_multi(a, cur, acc, lane);
            } else {
#pragma unroll
              for (int c = 0; c < kMaxCols; ++c) a.carry[(chunk * 2 + 0) * kMultiMaxDim + lane + 32 * c] = acc[c];
            }
          }
#pragma unroll
          for (int c = 0; c < kMaxCols; ++c) acc[c] = 0.f;
          cur = kj;
          starts_here = true;
        }
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c) acc[c] = fmaf(sq[u], gq[u][c], acc[c]);
      }
    }
  }
  if (cur == INT32_MAX) return;
  const bool ends_here = cur != next_key;
  if (starts_here && ends_here) {
    finish_row_multi(a, cur, acc, lane);
  } else {
    const int slot = starts_here ? 1 : 0;
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c) a.carry[(chunk * 2 + slot) * kMultiMaxDim + lane + 32 * c] = acc[c];
  }
}

__global__ void __launch_bounds__(256) adam_sorted_multi_pass2_kernel(const __grid_constant__ SortedMultiArgs a) {
  const int lane = threadIdx.x & 31;
  int64_t chunk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t beg = chunk * kChunk;
  if (beg >= a.n) return;
  int64_t end = min(a.n, beg + (int64_t)kChunk);
  if (end >= a.n) return;
  const int32_t key = __ldg(a.keys + end - 1);
  if (key == INT32_MAX || __ldg(a.keys + end) != key) return;
  if (__ldg(a.keys + beg) == key && beg > 0 && __ldg(a.keys + beg - 1) == key) return;
  float acc[kMaxCols];
#pragma unroll
  for (int c = 0; c < kMaxCols; ++c) acc[c] = a.carry[(chunk * 2 + 1) * kMultiMaxDim + lane + 32 * c];
  for (;;) {
    ++chunk;
    const int64_t b2 = chunk * kChunk;
    const int64_t e2 = min(a.n, b2 + (int64_t)kChunk);
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c) acc[c] += a.carry[(chunk * 2 + 0) * kMultiMaxDim + lane + 32 * c];
    if (e2 >= a.n || __ldg(a.keys + e2 - 1) != key || __ldg(a.keys + e2) != key) break;
  }
  finish_row_multi(a, key, acc, lane);
}

struct UntouchedMultiArgs {
  dmt_adam_table tab[DMT_MAX_ADAM_TABLES];
  AdamScalars s;
};

// blockIdx.y = table; the g = 0 update of every row the step did not touch (dense TF-1 semantics), float4 when the
// row width allows it
__global__ void __launch_bounds__(256) adam_untouched_multi_kernel(const __grid_constant__ UntouchedMultiArgs a) {
  const dmt_adam_table& t = a.tab[blockIdx.y];
  const bool vec = (t.dim % 4 == 0) && ((((uintptr_t)t.table | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0);
  if (vec) {
    const int chunks = t.dim / 4;
    const int64_t total = t.rows * chunks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = total < (1ll << 32) ? (int64_t)((uint32_t)i / (uint32_t)chunks) : i / chunks;
      if (t.touched[r]) continue;
      float4 p = *reinterpret_cast<float4*>(t.table + i * 4), mm = *reinterpret_cast<float4*>(t.m + i * 4);
      float4 vv = *reinterpret_cast<float4*>(t.v + i * 4);
      adam_update(p.x, mm.x, vv.x, 0.f, a.s);
      adam_update(p.y, mm.y, vv.y, 0.f, a.s);
      adam_update(p.z, mm.z, vv.z, 0.f, a.s);
      adam_update(p.w, mm.w, vv.w, 0.f, a.s);
      *reinterpret_cast<float4*>(t.table + i * 4) = p;
      *reinterpret_cast<float4*>(t.m + i * 4) = mm;
      *reinterpret_cast<float4*>(t.v + i * 4) = vv;
    }
  } else {
    const int64_t total = t.rows * t.dim;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      if (t.touched[i / t.dim]) continue;
      float p = t.table[i], mm = t.m[i], vv = t.v[i];
      adam_update(p, mm, vv, 0.f, a.s);
      t.table[i] = p;
      t.m[i] = mm;
      t.v[i] = vv;
    }
  }
}

static int sorted_launch(const SortedAdamArgs& a, cudaStream_t st) {
  const int64_t chunks = (a.n + kChunk - 1) / kChunk;
  const int64_t blocks = (chunks * 32 + 255) / 256;
  adam_sorted_pass1_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  DMT_CUDA_LAUNCH_CHECK("adam_sorted_pass1_kernel");
  adam_sorted_pass2_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  DMT_CUDA_LAUNCH_CHECK("adam_sorted_pass2_kernel");
  return DMT_OK;
}

}  // namespace dmt

extern "C" {

size_t dmt_embed_sorted_workspace_bytes(int64_t n, int32_t dim) {
  if (n <= 0 || dim <= 0) return 256;
  return (size_t)((n + dmt::kChunk - 1) / dmt::kChunk) * 2 * dim * sizeof(float) + 256;
}


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
                          uint8_t* touched, void* workspace, size_t workspace_bytes, void* stream) {
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
  a.dense_out = nullptr;
  DMT_REQUIRE(workspace && workspace_bytes >= dmt_embed_sorted_workspace_bytes(n, dim), DMT_ERR_WORKSPACE_TOO_SMALL,
              "dmt_embed_adam_sorted: workspace %zu < %zu bytes", workspace_bytes, dmt_embed_sorted_workspace_bytes(n, dim));
  a.carry = (float*)workspace;
  return dmt::sorted_launch(a, (cudaStream_t)stream);
}

int dmt_embed_grad_densify_sorted(int64_t rows, int32_t dim, int32_t n_sources, const dmt_grad_source* sources,
                                  const int32_t* sorted_keys, const int64_t* perm, const int64_t* refs,
                                  const float* scale, int64_t n, float grad_scale, float* dense_out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  DMT_REQUIRE(sources && sorted_keys && perm && refs && scale && dense_out, DMT_ERR_INVALID_ARGUMENT,
              "dmt_embed_grad_densify_sorted: null pointer");
  DMT_REQUIRE(dim > 0 && dim <= 128 && n_sources > 0 && n_sources <= DMT_MAX_GRAD_SOURCES, DMT_ERR_UNSUPPORTED_SHAPE,
              "dmt_embed_grad_densify_sorted: dim=%d n_sources=%d", dim, n_sources);
  if (n == 0) return DMT_OK;
  dmt::SortedAdamArgs a{};
  for (int s = 0; s < n_sources; ++s) a.src[s] = sources[s];
  a.rows = rows; a.dim = dim;
  a.keys = sorted_keys; a.perm = perm; a.refs = refs; a.scale = scale;
  a.n = n; a.gscale = grad_scale;
  a.dense_out = dense_out;
  DMT_REQUIRE(workspace && workspace_bytes >= dmt_embed_sorted_workspace_bytes(n, dim), DMT_ERR_WORKSPACE_TOO_SMALL,
              "dmt_embed_grad_densify_sorted: workspace %zu < %zu bytes", workspace_bytes,
              dmt_embed_sorted_workspace_bytes(n, dim));
  a.carry = (float*)workspace;
  return dmt::sorted_launch(a, (cudaStream_t)stream);
}

int dmt_embed_grad_scatter_rows(int32_t n_sources, const dmt_grad_source* sources, const int32_t* keys,
                                const int64_t* refs, const float* scale, int64_t n, int32_t dim, float* out,
                                void* stream) {
  DMT_REQUIRE(sources && keys && refs && scale && out, DMT_ERR_INVALID_ARGUMENT,
              "dmt_embed_grad_scatter_rows: null pointer");
  DMT_REQUIRE(dim > 0 && n_sources > 0 && n_sources <= DMT_MAX_GRAD_SOURCES, DMT_ERR_UNSUPPORTED_SHAPE,
              "dmt_embed_grad_scatter_rows: dim=%d n_sources=%d", dim, n_sources);
  if (n == 0) return DMT_OK;
  dmt::ScatterRowsArgs a{};
  for (int s = 0; s < n_sources; ++s) a.src[s] = sources[s];
  a.keys = keys; a.refs = refs; a.scale = scale; a.n = n; a.dim = dim; a.out = out;
  const int64_t blocks = (n * 32 + 255) / 256;
  dmt::grad_scatter_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  DMT_CUDA_LAUNCH_CHECK("grad_scatter_rows_kernel");
  return DMT_OK;
}

int dmt_adam_rows_untouched(const dmt_adam_cfg* cfg, float* table, float* m, float* v, int64_t rows, int32_t dim,
                            uint8_t* touched, void* stream) {
  DMT_REQUIRE(cfg && table && m && v && touched && cfg->step >= 1, DMT_ERR_INVALID_ARGUMENT,
              "dmt_adam_rows_untouched: bad arguments");
  if (rows == 0) return DMT_OK;
  if (cfg->kind != DMT_OPT_ADAM) {             // a row without gradient does not move: only clear the marks
    cudaError_t e0 = cudaMemsetAsync(touched, 0, (size_t)rows, (cudaStream_t)stream);
    if (e0 != cudaSuccess) return dmt::cuda_fail(e0, "cudaMemsetAsync(touched)");
    return DMT_OK;
  }
  const bool vec = dim % 4 == 0 && (((uintptr_t)table | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
  int64_t blocks = (rows * (vec ? dim / 4 : dim) + 255) / 256;
  const int64_t cap = (int64_t)dmt::sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  if (vec)
    dmt::adam_untouched_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(table, m, v, rows, dim, touched,
                                                                                     dmt::adam_scalars(cfg));
  else
    dmt::adam_untouched_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(table, m, v, rows, dim, touched,
                                                                                     dmt::adam_scalars(cfg));
  DMT_CUDA_LAUNCH_CHECK("adam_untouched_kernel");
  cudaError_t e = cudaMemsetAsync(touched, 0, (size_t)rows, (cudaStream_t)stream);
  if (e != cudaSuccess) return dmt::cuda_fail(e, "cudaMemsetAsync(touched)");
  return DMT_OK;
}

static int check_tables(const char* who, int32_t n_tables, const dmt_adam_table* tables) {
  DMT_REQUIRE(tables && n_tables > 0 && n_tables <= DMT_MAX_ADAM_TABLES, DMT_ERR_INVALID_ARGUMENT,
              "%s: n_tables=%d (max %d)", who, n_tables, DMT_MAX_ADAM_TABLES);
  for (int t = 0; t < n_tables; ++t) {
    DMT_REQUIRE((tables[t].dense_out || (tables[t].table && tables[t].m && tables[t].v && tables[t].touched)) &&
                    tables[t].dim > 0 && tables[t].dim <= dmt::kMultiMaxDim && tables[t].rows > 0,
                DMT_ERR_INVALID_ARGUMENT, "%s: table %d is incomplete", who, t);
    DMT_REQUIRE(tables[t].rows <= ((int64_t)1 << dmt::kMultiRowBits), DMT_ERR_UNSUPPORTED_SHAPE,
                "%s: table %d has %lld rows (the packed keys hold 2^%d; use the per-table entry points)", who, t,
                (long long)tables[t].rows, dmt::kMultiRowBits);
  }
  return DMT_OK;
}

int dmt_embed_grad_expand_multi(int32_t n_tables, const dmt_adam_table* tables, int32_t n_sources,
                                const dmt_grad_source* sources, const int32_t* source_table, int32_t* keys,
                                int64_t* refs, float* scale, void* stream) {
  int rc = check_tables("dmt_embed_grad_expand_multi", n_tables, tables);
  if (rc != DMT_OK) return rc;
  DMT_REQUIRE(sources && source_table && keys && refs && scale && n_sources > 0 && n_sources <= dmt::kMaxMultiSources,
              DMT_ERR_INVALID_ARGUMENT, "dmt_embed_grad_expand_multi: n_sources=%d (max %d)", n_sources,
              dmt::kMaxMultiSources);
  dmt::ExpandMultiArgs a{};
  a.base[0] = 0;
  for (int s = 0; s < n_sources; ++s) {
    DMT_REQUIRE(sources[s].ids && sources[s].n >= 0 && source_table[s] >= 0 && source_table[s] < n_tables,
                DMT_ERR_INVALID_ARGUMENT, "dmt_embed_grad_expand_multi: source %d", s);
    a.src[s] = sources[s];
    a.tid[s] = (int8_t)source_table[s];
    a.base[s + 1] = a.base[s] + sources[s].n;
  }
  for (int t = 0; t < n_tables; ++t) a.rows[t] = tables[t].rows;
  a.n_sources = n_sources;
  a.keys = keys;
  a.refs = refs;
  a.scale = scale;
  const int64_t total = a.base[n_sources];
  if (total == 0) return DMT_OK;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)dmt::sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  dmt::grad_expand_multi_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  DMT_CUDA_LAUNCH_CHECK("grad_expand_multi_kernel");
  return DMT_OK;
}

size_t dmt_embed_sorted_multi_workspace_bytes(int64_t n) {
  if (n <= 0) return 256;
  return (size_t)((n + dmt::kChunk - 1) / dmt::kChunk) * 2 * dmt::kMultiMaxDim * sizeof(float) + 256;
}

int dmt_embed_adam_sorted_multi(const dmt_adam_cfg* cfg, int32_t n_tables, const dmt_adam_table* tables,
                                int32_t n_sources, const dmt_grad_source* sources, const int32_t* sorted_keys,
                                const int64_t* perm, const int64_t* refs, const float* scale, int64_t n,
                                float grad_scale, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_tables("dmt_embed_adam_sorted_multi", n_tables, tables);
  if (rc != DMT_OK) return rc;
  bool any_adam = false;
  for (int t = 0; t < n_tables; ++t) any_adam = any_adam || tables[t].dense_out == nullptr;
  DMT_REQUIRE((!any_adam || (cfg && cfg->step >= 1)) && sources && sorted_keys && perm && refs && scale &&
                  n_sources > 0 && n_sources <= dmt::kMaxMultiSources,
              DMT_ERR_INVALID_ARGUMENT, "dmt_embed_adam_sorted_multi: bad arguments");
  if (n == 0) return DMT_OK;
  DMT_REQUIRE(workspace && workspace_bytes >= dmt_embed_sorted_multi_workspace_bytes(n), DMT_ERR_WORKSPACE_TOO_SMALL,
              "dmt_embed_adam_sorted_multi: workspace %zu < %zu bytes", workspace_bytes,
              dmt_embed_sorted_multi_workspace_bytes(n));
  dmt::SortedMultiArgs a{};
  for (int s = 0; s < n_sources; ++s) a.src[s] = sources[s];
  for (int t = 0; t < n_tables; ++t) a.tab[t] = tables[t];
  a.keys = sorted_keys; a.perm = perm; a.refs = refs; a.scale = scale;
  a.n = n; a.gscale = grad_scale;
  if (any_adam) a.s = dmt::adam_scalars(cfg);
  a.carry = (float*)workspace;
  const int64_t chunks = (n + dmt::kChunk - 1) / dmt::kChunk;
  const int64_t blocks = (chunks * 32 + 255) / 256;
  dmt::adam_sorted_multi_pass1_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  DMT_CUDA_LAUNCH_CHECK("adam_sorted_multi_pass1_kernel");
  dmt::adam_sorted_multi_pass2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  DMT_CUDA_LAUNCH_CHECK("adam_sorted_multi_pass2_kernel");
  return DMT_OK;
}

int dmt_adam_rows_untouched_multi(const dmt_adam_cfg* cfg, int32_t n_tables, const dmt_adam_table* tables,
                                  void* stream) {
  int rc = check_tables("dmt_adam_rows_untouched_multi", n_tables, tables);
  if (rc != DMT_OK) return rc;
  DMT_REQUIRE(cfg && cfg->step >= 1, DMT_ERR_INVALID_ARGUMENT, "dmt_adam_rows_untouched_multi: bad arguments");
  if (cfg->kind != DMT_OPT_ADAM) return DMT_OK;   // a row without gradient does not move
  dmt::UntouchedMultiArgs a{};
  int64_t biggest = 1;
  for (int t = 0; t < n_tables; ++t) {
    a.tab[t] = tables[t];
    const int64_t work = tables[t].rows * (tables[t].dim % 4 == 0 ? tables[t].dim / 4 : tables[t].dim);
    if (work > biggest) biggest = work;
  }
  a.s = dmt::adam_scalars(cfg);
  int64_t blocks = (biggest + 255) / 256;
  const int64_t cap = (int64_t)dmt::sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  dmt::adam_untouched_multi_kernel<<<dim3((unsigned)blocks, (unsigned)n_tables), 256, 0, (cudaStream_t)stream>>>(a);
  DMT_CUDA_LAUNCH_CHECK("adam_untouched_multi_kernel");
  return DMT_OK;    // the caller clears its (contiguous) `touched` marks with one memset
}

}  // extern "C"
