// Grouped fp32 SIMT GEMM (see gemm_f32.cuh): 16*TM x 64 x 16 tiles, 256 threads, TM x 4 register tile.
#include "gemm_f32.cuh"

namespace dmt {

namespace {

constexpr int BN = 64, BK = 16;

__device__ __forceinline__ float epilogue(const GemmProb& P, float v, int m, int n) {
  if (P.addend) v += __ldg(P.addend + (int64_t)m * P.ld_add + n);
  v *= P.alpha;
  if (P.bias) v += __ldg(P.bias + n);
  if (P.relu) v = fmaxf(v, 0.f);
  if (P.mask && !(__ldg(P.mask + (int64_t)m * P.ld_mask + n) > 0.f)) v = 0.f;
  return v;
}

template <int TM>
__global__ void __launch_bounds__(256) gemm_group_kernel(const __grid_constant__ GemmGroup g) {
  constexpr int BM = 16 * TM;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  int pi = 0;
  while (pi + 1 < g.n && (int)blockIdx.x >= g.p[pi + 1].cta0) ++pi;
  const GemmProb& P = g.p[pi];
  int local = blockIdx.x - P.cta0;
  const int tiles = P.tiles_m * P.tiles_n;
  const int split = local / tiles;
  local -= split * tiles;
  const int tmi = local / P.tiles_n, tni = local - tmi * P.tiles_n;
  const int m0 = tmi * BM, n0 = tni * BN;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int M = P.M, N = P.N;
  const bool do_colsum = P.colsum != nullptr && tmi == 0;

  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  float cs[4] = {0.f, 0.f, 0.f, 0.f};

  for (int part = 0; part < P.n_parts; ++part) {
    const float* __restrict__ A = P.part[part].A;
    const float* __restrict__ B = P.part[part].B;
    const int64_t lda = P.part[part].lda, ldb = P.part[part].ldb;
    const int K = P.part[part].K;
    int kbeg = 0, kend = K;
    if (P.splits > 1) {
      const int chunk = (((K + P.splits - 1) / P.splits) + BK - 1) / BK * BK;
      kbeg = min(K, split * chunk);
      kend = min(K, kbeg + chunk);
    }
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
      for (int i = 0; i < (BM * BK) / 256; ++i) {
        const int idx = tid + i * 256;
        int r, kk;
        if (P.transA) { kk = idx / BM; r = idx - kk * BM; } else { r = idx / BK; kk = idx - r * BK; }
        const int m = m0 + r, k = k0 + kk;
        float v = 0.f;
        if (m < M && k < kend) v = __ldg(P.transA ? A + (int64_t)k * lda + m : A + (int64_t)m * lda + k);
        As[kk][r] = v;
      }
#pragma unroll
      for (int i = 0; i < (BK * BN) / 256; ++i) {
        const int idx = tid + i * 256;
        int c, kk;
        if (P.transB) { c = idx / BK; kk = idx - c * BK; } else { kk = idx / BN; c = idx - kk * BN; }
        const int n = n0 + c, k = k0 + kk;
        float v = 0.f;
        if (n < N && k < kend) v = __ldg(P.transB ? B + (int64_t)n * ldb + k : B + (int64_t)k * ldb + n);
        Bs[kk][c] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 w = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        float av[TM];
#pragma unroll
        for (int i = 0; i < TM; i += 4) {
          const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
          av[i] = a4.x; av[i + 1] = a4.y; av[i + 2] = a4.z; av[i + 3] = a4.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          acc[i][0] = fmaf(av[i], w.x, acc[i][0]);
          acc[i][1] = fmaf(av[i], w.y, acc[i][1]);
          acc[i][2] = fmaf(av[i], w.z, acc[i][2]);
          acc[i][3] = fmaf(av[i], w.w, acc[i][3]);
        }
        if (do_colsum) { cs[0] += w.x; cs[1] += w.y; cs[2] += w.z; cs[3] += w.w; }
      }
      __syncthreads();
    }
  }

  if (P.splits > 1) {
    const int rows = M + (P.colsum ? 1 : 0);
    float* __restrict__ part = P.partial + (int64_t)split * rows * N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty * TM + i;
      if (m >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n < N) part[(int64_t)m * N + n] = acc[i][j];
      }
    }
    if (do_colsum && ty == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n < N) part[(int64_t)M * N + n] = cs[j];
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = epilogue(P, acc[i][j], m, n);
      float* c = P.C + (int64_t)m * P.ldc + n;
      if (P.accumulate) v += *c;
      *c = v;
    }
  }
  if (do_colsum && ty == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) P.colsum[n] = (P.colsum_accumulate ? P.colsum[n] : 0.f) + cs[j];
    }
  }
}

// Fixed-order reduction of the split-K partials + epilogue.
__global__ void __launch_bounds__(256) gemm_group_reduce_kernel(const __grid_constant__ GemmGroup g) {
  int pi = -1;
  for (int i = 0; i < g.n; ++i)
    if (g.p[i].splits > 1 && (int)blockIdx.x >= g.p[i].red0) pi = i;
  if (pi < 0) return;
  const GemmProb& P = g.p[pi];
  const int M = P.M, N = P.N;
  const int rows = M + (P.colsum ? 1 : 0);
  const int64_t idx = (int64_t)(blockIdx.x - P.red0) * 256 + threadIdx.x;
  if (idx >= (int64_t)rows * N) return;
  const int m = (int)(idx / N), n = (int)(idx - (int64_t)m * N);
  float s = 0.f;
  for (int k = 0; k < P.splits; ++k) s += P.partial[((int64_t)k * rows + m) * N + n];
  if (m < M) {
    float v = epilogue(P, s, m, n);
    float* c = P.C + (int64_t)m * P.ldc + n;
    if (P.accumulate) v += *c;
    *c = v;
  } else {
    P.colsum[n] = (P.colsum_accumulate ? P.colsum[n] : 0.f) + s;
  }
}

}  // namespace

size_t gemm_partial_bytes(const GemmProb& p) {
  if (p.splits <= 1) return 0;
  return (size_t)p.splits * (p.M + (p.colsum ? 1 : 0)) * p.N * sizeof(float);
}

int gemm_pick_splits_tc(int M, int N, int64_t K, bool colsum);
int gemm_tc_group_launch(GemmGroup& g, cudaStream_t st, int (*reduce)(GemmGroup&, cudaStream_t));

static int launch_reduce(GemmGroup& g, cudaStream_t st) {
  gemm_group_reduce_kernel<<<g.total_red, 256, 0, st>>>(g);
  DMT_CUDA_LAUNCH_CHECK("gemm_group_reduce_kernel");
  return DMT_OK;
}

int gemm_pick_splits(int M, int N, int64_t K, bool use_tc, bool colsum) {
  if (use_tc) return gemm_pick_splits_tc(M, N, K, colsum);
  const int tiles = ((M + 63) / 64) * ((N + BN - 1) / BN);
  const int target = 2 * sm_count_cached();
  int64_t s = target / (tiles > 0 ? tiles : 1);
  const int64_t kmax = K / 256;   // at least 256 contraction rows per split
  if (s > kmax) s = kmax;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return (int)s;
}

int gemm_group_launch(GemmGroup& g, cudaStream_t st) {
  if (g.n <= 0) return DMT_OK;
  DMT_REQUIRE(g.n <= kGemmMaxProbs, DMT_ERR_INVALID_ARGUMENT, "gemm group: %d problems (max %d)", g.n, kGemmMaxProbs);
  int min_m = 1 << 30;
  bool empty = true;
  for (int i = 0; i < g.n; ++i) {
    const GemmProb& p = g.p[i];
    if (p.M > 0 && p.N > 0) {
      empty = false;
      if (p.M < min_m) min_m = p.M;
    }
    DMT_REQUIRE(p.n_parts >= 1 && p.n_parts <= kGemmMaxParts, DMT_ERR_INVALID_ARGUMENT, "gemm group: n_parts=%d", p.n_parts);
    DMT_REQUIRE(p.splits == 1 || (p.n_parts == 1 && p.partial), DMT_ERR_INVALID_ARGUMENT,
                "gemm group: split-K needs a single part and a partial buffer");
    DMT_REQUIRE(!p.colsum || p.n_parts == 1, DMT_ERR_INVALID_ARGUMENT, "gemm group: colsum needs a single part");
  }
  if (empty) return DMT_OK;
  if (g.use_tc) return gemm_tc_group_launch(g, st, launch_reduce);
  const int TM = min_m >= 2048 ? 8 : 4;
  const int BM = 16 * TM;
  int cta = 0, red = 0;
  for (int i = 0; i < g.n; ++i) {
    GemmProb& p = g.p[i];
    p.tiles_m = (p.M + BM - 1) / BM;
    p.tiles_n = (p.N + BN - 1) / BN;
    p.cta0 = cta;
    p.red0 = red;
    cta += p.tiles_m * p.tiles_n * p.splits;
    if (p.splits > 1) red += (int)(((int64_t)(p.M + (p.colsum ? 1 : 0)) * p.N + 255) / 256);
  }
  g.total_ctas = cta;
  g.total_red = red;
  if (cta > 0) {
    if (TM == 8)
      gemm_group_kernel<8><<<cta, 256, 0, st>>>(g);
    else
      gemm_group_kernel<4><<<cta, 256, 0, st>>>(g);
    DMT_CUDA_LAUNCH_CHECK("gemm_group_kernel");
  }
  if (red > 0) {
    gemm_group_reduce_kernel<<<red, 256, 0, st>>>(g);
    DMT_CUDA_LAUNCH_CHECK("gemm_group_reduce_kernel");
  }
  return DMT_OK;
}

}  // namespace dmt
