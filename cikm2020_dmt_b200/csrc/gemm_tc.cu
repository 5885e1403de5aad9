// Grouped GEMM on tcgen05 tensor cores for the bf16 training path: same problem descriptors and epilogues
// as the fp32 SIMT kernel (gemm_f32.cuh), operands converted fp32 -> bf16 while they are staged.
//
//   * one CTA = one 128-row accumulator tile x the problem's whole N (<= 256) x one K range; accumulator in
//     TMEM (fp32), tcgen05.mma 128 x Npad x 16 issued by one thread, two shared-memory stages of K = 64 so the
//     threads stage chunk i+1 while the tensor core works on chunk i (mbarrier per stage);
//   * staging writes the NO-SWIZZLE canonical K-major image [k/8][row][8] for BOTH operands whatever their
//     orientation in memory: a thread owns (row, 8 consecutive k), reads them as two float4 (K contiguous) or
//     as 8 warp-coalesced scalars (K strided, the weight-gradient contractions over tokens) and stores ONE
//     16-byte chunk -- consecutive lanes hit consecutive 16-byte slots, bank-conflict free;
//   * use_tc == 3 ("bf16x3"): every operand is split x = hi + lo (two bf16 images) and each K step issues
//     hi*hi + hi*lo + lo*hi -- ~2^-16 relative operand precision, i.e. fp32-grade gradients at tensor-core
//     speed (these GEMMs are bound by staging / HBM, not by the MMA rate); use_tc == 1 is plain bf16;
//   * the bias gradient (column sums of B) rides along as an extra all-ones row of A;
//   * split-K partials go through the same fixed-order reduce kernel as the fp32 path (deterministic).
#include "gemm_f32.cuh"
#include "umma.cuh"

namespace dmt {

using namespace umma;

namespace {

constexpr int kTcThreads = 256;
constexpr int BM = 128, BN = 256;                 // tile rows, max tile columns
// K per shared-memory stage: 64, or 32 in the split (x3) mode whose two images per operand would otherwise
// double the footprint and halve the CTAs per SM (these GEMMs live on occupancy, not on MMA rate)
__host__ __device__ constexpr int stage_k(bool x3) { return x3 ? 32 : 64; }
// one bf16 image of a B chunk is Nmax * 128 bytes, Nmax = the widest (16-padded) tile of the group

__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]);
  v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]);
  v.w = pack_bf16x2(f[6], f[7]);
  return v;
}

// 8 consecutive-k elements of row `r` of an operand: contiguous (ld = row stride) or strided (ld = k stride).
__device__ __forceinline__ void load8(const float* __restrict__ base, int64_t ld, bool k_contig, int64_t r, int k,
                                      int kend, bool vec_ok, float* f) {
  if (k_contig) {
    const float* p = base + r * ld + k;
    if (vec_ok && k + 8 <= kend) {
      const float4 a = ldg4(p), b = ldg4(p + 4);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = k + e < kend ? __ldg(p + e) : 0.f;
    }
  } else {
    const float* p = base + (int64_t)k * ld + r;
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = k + e < kend ? __ldg(p + (int64_t)e * ld) : 0.f;
  }
}

// hi / lo split of 8 values: hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void split8(const float* f, uint4& hi, uint4& lo) {
  float l[8];
  __nv_bfloat16 h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    h[e] = __float2bfloat16(f[e]);
    l[e] = f[e] - __bfloat162float(h[e]);
  }
  uint32_t* hp = reinterpret_cast<uint32_t*>(&hi);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 v = __halves2bfloat162(h[2 * e], h[2 * e + 1]);
    hp[e] = *reinterpret_cast<uint32_t*>(&v);
  }
  lo = pack8(l);
}

// one 16-byte chunk (8 bf16) of an operand image; with X3 also the low-order image `lo_off` bytes behind it
template <int X3>
__device__ __forceinline__ void store_chunk(uint8_t* dst, int lo_off, const float* f) {
  if (X3) {
    uint4 hi, lo;
    split8(f, hi, lo);
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + lo_off) = lo;
  } else {
    *reinterpret_cast<uint4*>(dst) = pack8(f);
  }
}

template <int X3>
__global__ void __launch_bounds__(kTcThreads, 4) gemm_tc_group_kernel(const __grid_constant__ GemmGroup g) {
  constexpr int BKC = stage_k(X3 != 0);
  constexpr int kStageA = BM * BKC * 2;                 // one bf16 image of an A chunk
  const int kStage = kStageA + g.tc_nmax * (BKC * 2);   // bytes of one (A, B) image pair
  const int kStageAll = (X3 ? 2 : 1) * kStage;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mbar[2];
  __shared__ uint32_t tmem_base_s;
  int pi = 0;
  while (pi + 1 < g.n && (int)blockIdx.x >= g.p[pi + 1].cta0) ++pi;
  const GemmProb& P = g.p[pi];
  int local = blockIdx.x - P.cta0;
  const int tiles = P.tiles_m * P.tiles_n;
  const int split = local / tiles;
  local -= split * tiles;
  const int tmi = local / P.tiles_n, tni = local - tmi * P.tiles_n;
  const int m0 = tmi * BM, n0 = tni * P.tc_bn;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = P.M, NF = P.N;                   // NF: the problem's full N (row stride of the partials)
  const int N = min(P.tc_bn, NF - n0);           // columns of this tile
  const int Npad = (N + 15) & ~15;
  const bool ones_row = P.colsum != nullptr;
  uint32_t ncols = 32;
  while ((int)ncols < Npad) ncols <<= 1;

  if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  // operand images: K contiguous in memory -> canonical K-major [k/8][row][8] (LBO = rows*16 to the next 8 k,
  // SBO = 128 to the next 8 rows); row index contiguous in memory (the token-contracting weight-gradient
  // operands, the [K,N] weights of a forward GEMM) -> MN-major [row/8][k][8] (LBO = 128 to the next 8 k,
  // SBO = 64*16 to the next 8 rows), staged without any transposition.
  const bool a_mn = P.transA != 0, b_mn = P.transB == 0;
  const uint32_t idesc = make_idesc_bf16(BM, Npad, a_mn, b_mn);
  const uint32_t dHiA = a_mn ? desc_hi(BKC * 16, kLayoutNone) : desc_hi(128, kLayoutNone);
  const uint32_t dHiB = b_mn ? desc_hi(BKC * 16, kLayoutNone) : desc_hi(128, kLayoutNone);

  int it = 0;
  for (int part = 0; part < P.n_parts; ++part) {
    const float* __restrict__ A = P.part[part].A;
    const float* __restrict__ Bm = P.part[part].B;
    const int64_t lda = P.part[part].lda, ldb = P.part[part].ldb;
    const int K = P.part[part].K;
    int kbeg = 0, kend = K;
    if (P.splits > 1) {
      const int chunk = (((K + P.splits - 1) / P.splits) + BKC - 1) / BKC * BKC;
      kbeg = min(K, split * chunk);
      kend = min(K, kbeg + chunk);
    }
    // 16-byte loads need: aligned base, row stride a multiple of 4 floats and a 4-aligned first element
    const bool a_vec = (lda % 4 == 0) && (((uintptr_t)A & 15) == 0) && ((a_mn ? m0 : kbeg) % 4 == 0);
    const bool b_vec = (ldb % 4 == 0) && (((uintptr_t)Bm & 15) == 0) && ((b_mn ? n0 : kbeg) % 4 == 0);
    for (int k0 = kbeg; k0 < kend; k0 += BKC, ++it) {
      const int s = g.tc_stages > 1 ? (it & 1) : 0;
      if (g.tc_stages > 1) {
        if (it >= 2) mbar_wait(&mbar[s], ((it >> 1) - 1) & 1);   // the MMAs that read this stage have retired
      } else if (it >= 1) {
        mbar_wait(&mbar[0], (it - 1) & 1);
      }
      uint8_t* sA = smem + s * kStageAll;      // [A hi | B hi | A lo | B lo]
      uint8_t* sB = sA + kStageA;
      // ---- A: 128 rows x 64 k
      if (!a_mn) {      // K contiguous in memory -> K-major image [k/8][row][8]
#pragma unroll
        for (int j = 0; j < BM * (BKC / 8) / kTcThreads; ++j) {
          const int i = tid + j * kTcThreads, row = i % BM, kc = i / BM;
          const int m = m0 + row, k = k0 + kc * 8;
          float f[8];
          if (m < M && k < kend) {
            load8(A, lda, true, m, k, kend, a_vec, f);
          } else {
            const bool one = ones_row && m == M;       // the bias-gradient row
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = (one && k + e < kend) ? 1.0f : 0.f;
          }
          store_chunk<X3>(sA + kc * (BM * 16) + row * 16, kStage, f);
        }
      } else {          // M contiguous in memory -> MN-major image [m/8][k][8], no transposition needed
#pragma unroll
        for (int j = 0; j < (BM / 8) * BKC / kTcThreads; ++j) {
          const int blk = warp + j * (kTcThreads / 32);              // 4 row-octets x 8 k per warp-instruction
          const int o = (blk & 3) * 4 + (lane >> 3), kk = (blk >> 2) * 8 + (lane & 7);
          const int m = m0 + o * 8, k = k0 + kk;
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = 0.f;
          if (k < kend) {
            if (m + 8 <= M && a_vec) {
              const float4 x = ldg4(A + (int64_t)k * lda + m), y = ldg4(A + (int64_t)k * lda + m + 4);
              f[0] = x.x; f[1] = x.y; f[2] = x.z; f[3] = x.w; f[4] = y.x; f[5] = y.y; f[6] = y.z; f[7] = y.w;
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                if (m + e < M) f[e] = __ldg(A + (int64_t)k * lda + m + e);
                else if (ones_row && m + e == M) f[e] = 1.0f;
              }
            }
          }
          store_chunk<X3>(sA + o * (BKC * 16) + kk * 16, kStage, f);
        }
      }
      // ---- B: Npad rows x 64 k.  Four items per trip: all their (independent) global loads are issued before
      //      the first conversion, otherwise every trip of this runtime-bounded loop costs one DRAM latency.
      if (!b_mn) {
        const int total = Npad * (BKC / 8);
        for (int i0 = tid; i0 < total; i0 += 4 * kTcThreads) {
          float f[4][8];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kTcThreads;
            const int n = i % Npad, kc = i / Npad, k = k0 + kc * 8;
            if (i < total && n < N && k < kend) {
              load8(Bm, ldb, true, n0 + n, k, kend, b_vec, f[u]);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[u][e] = 0.f;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kTcThreads;
            if (i < total) store_chunk<X3>(sB + (i / Npad) * (Npad * 16) + (i % Npad) * 16, kStage, f[u]);
          }
        }
      } else {
        const int n_oct = Npad / 8, groups = (n_oct + 3) / 4, total = groups * (BKC / 8);
        for (int b0 = warp; b0 < total; b0 += 4 * (kTcThreads / 32)) {
          float f[4][8];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int blk = b0 + u * (kTcThreads / 32);
            const int o = (blk % groups) * 4 + (lane >> 3), kk = (blk / groups) * 8 + (lane & 7);
            const int n = o * 8, k = k0 + kk;
#pragma unroll
            for (int e = 0; e < 8; ++e) f[u][e] = 0.f;
            if (blk < total && o < n_oct && k < kend) {
              const float* p = Bm + (int64_t)k * ldb + n0 + n;
              if (n + 8 <= N && b_vec) {
                const float4 x = ldg4(p), y = ldg4(p + 4);
                f[u][0] = x.x; f[u][1] = x.y; f[u][2] = x.z; f[u][3] = x.w;
                f[u][4] = y.x; f[u][5] = y.y; f[u][6] = y.z; f[u][7] = y.w;
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  if (n + e < N) f[u][e] = __ldg(p + e);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int blk = b0 + u * (kTcThreads / 32);
            const int o = (blk % groups) * 4 + (lane >> 3), kk = (blk / groups) * 8 + (lane & 7);
            if (blk < total && o < n_oct) store_chunk<X3>(sB + o * (BKC * 16) + kk * 16, kStage, f[u]);
          }
        }
      }
      fence_proxy_async();
      fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        fence_after_sync();
        const uint32_t dA = desc_lo(smem_u32(sA), a_mn ? 128 : BM * 16);
        const uint32_t dB = desc_lo(smem_u32(sB), b_mn ? 128 : Npad * 16);
        const uint32_t stepA = a_mn ? 16 : 2 * BM, stepB = b_mn ? 16 : 2 * Npad;   // 16 k further, in 16-byte units
#pragma unroll
        for (int ks = 0; ks < BKC / 16; ++ks) {
          const uint64_t ah = desc_join(dA + ks * stepA, dHiA), bh = desc_join(dB + ks * stepB, dHiB);
          mma_bf16_ss(tbase, ah, bh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          if (X3) {   // + hi*lo + lo*hi (the lo images sit kStage bytes behind the hi images)
            const uint64_t al = desc_join(dA + kStage / 16 + ks * stepA, dHiA);
            const uint64_t bl = desc_join(dB + kStage / 16 + ks * stepB, dHiB);
            mma_bf16_ss(tbase, ah, bl, idesc, 1u);
            mma_bf16_ss(tbase, al, bh, idesc, 1u);
          }
        }
        commit(&mbar[s]);
      }
    }
  }
  if (it > 0) {
    const int last = it - 1;
    if (g.tc_stages > 1)
      mbar_wait(&mbar[last & 1], (last >> 1) & 1);   // tcgen05.commit covers every MMA issued before it
    else
      mbar_wait(&mbar[0], last & 1);
    fence_after_sync();
  }

  // ---- epilogue.  TMEM hands a thread 32 consecutive columns of ITS row; global memory wants a warp on 32
  //      consecutive columns of ONE row.  Each warp transposes its 32 x 32 block through shared memory (the
  //      staging buffers are free now; stride 33 floats: conflict-free both ways), then every load of the
  //      addend / mask and every store is one coalesced 128-byte row segment.
  //      warps w and w+4 own the same 32 rows (TMEM lane quarter) and take alternate 32-column blocks.
  const int half = warp >> 2;
  const int mw = m0 + (warp & 3) * 32;                 // first row of this warp's lane quarter
  const int rows_out = M + (ones_row ? 1 : 0);
  float* tb = reinterpret_cast<float*>(smem) + warp * (32 * 33);
  for (int c0 = half * 32; c0 < Npad; c0 += 64) {
    uint32_t r[32];
    if (it > 0) {
      tmem_ld32(tmem_addr(tbase, c0), r);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = 0u;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) tb[lane * 33 + j] = __uint_as_float(r[j]);
    __syncwarp();
    const int n = c0 + lane;                            // this lane's column inside the tile
    if (n < N) {
      if (P.splits > 1) {
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
          const int m = mw + rr;
          if (m < rows_out) P.partial[((int64_t)split * rows_out + m) * NF + n0 + n] = tb[rr * 33 + lane];
        }
      } else {
        // 8 rows per trip: all the (independent) addend / mask / C loads of the trip are issued first
        const float bias = P.bias ? __ldg(P.bias + n0 + n) : 0.f;
        for (int r0 = 0; r0 < 32; r0 += 8) {
          float ad[8], mk[8], old[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int m = mw + r0 + u;
            const bool ok = m < M;
            ad[u] = (ok && P.addend) ? __ldg(P.addend + (int64_t)m * P.ld_add + n0 + n) : 0.f;
            mk[u] = (ok && P.mask) ? __ldg(P.mask + (int64_t)m * P.ld_mask + n0 + n) : 1.f;
            old[u] = (ok && P.accumulate) ? P.C[(int64_t)m * P.ldc + n0 + n] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int m = mw + r0 + u;
            const float acc = tb[(r0 + u) * 33 + lane];
            if (m < M) {
              float v = (acc + ad[u]) * P.alpha + bias;
              if (P.relu) v = fmaxf(v, 0.f);
              if (!(mk[u] > 0.f)) v = 0.f;
              P.C[(int64_t)m * P.ldc + n0 + n] = v + old[u];
            } else if (m == M && ones_row) {            // the all-ones row: column sums of B
              float* cs = P.colsum + n0 + n;
              *cs = (P.colsum_accumulate ? *cs : 0.f) + acc;
            }
          }
        }
      }
    }
    __syncwarp();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, ncols);
}

}  // namespace

// Tile width: a row-parallel problem (activations x weights^T over all tokens) takes 64-column tiles -- small
// CTAs, several per SM, so one tile's epilogue overlaps another's staging; a weight-gradient contraction
// (few tiles, K = all tokens, split-K) takes the full 256 so the token operand is staged once.
static inline int pick_bn(const GemmProb& p) { return (!p.transA && p.M >= 4096 && p.N > 64) ? 64 : BN; }

int gemm_pick_splits_tc(int M, int N, int64_t K, bool colsum) {
  const int tiles = ((M + (colsum ? 1 : 0) + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int target = 2 * sm_count_cached();
  int64_t s = target / (tiles > 0 ? tiles : 1);
  const int64_t kmax = K / 512;   // at least 512 contraction rows per split
  if (s > kmax) s = kmax;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return (int)s;
}

// Launches the group on tensor cores.  `reduce` runs the shared fixed-order split-K reduction afterwards.
int gemm_tc_group_launch(GemmGroup& g, cudaStream_t st, int (*reduce)(GemmGroup&, cudaStream_t)) {
  int cta = 0, red = 0;
  bool any = false;
  for (int i = 0; i < g.n; ++i) {
    GemmProb& p = g.p[i];
    const int rows = p.M + (p.colsum ? 1 : 0);
    p.tc_bn = pick_bn(p);
    p.tiles_m = (p.M > 0 && p.N > 0) ? (rows + BM - 1) / BM : 0;
    p.tiles_n = (p.N + p.tc_bn - 1) / p.tc_bn;
    p.cta0 = cta;
    p.red0 = red;
    cta += p.tiles_m * p.tiles_n * p.splits;
    if (p.tiles_m > 0) any = true;
    if (p.splits > 1 && p.tiles_m > 0) red += (int)(((int64_t)rows * p.N + 255) / 256);
  }
  g.total_ctas = cta;
  g.total_red = red;
  if (!any) return DMT_OK;
  const bool x3 = g.use_tc == 3;
  const int kc = stage_k(x3);
  int nmax = 16, stages = 1;
  for (int i = 0; i < g.n; ++i) {
    const GemmProb& p = g.p[i];
    if (p.tiles_m == 0) continue;
    const int w = p.N < p.tc_bn ? p.N : p.tc_bn;
    if (((w + 15) & ~15) > nmax) nmax = (w + 15) & ~15;
    int64_t chunks = 0;                      // K chunks one CTA walks
    for (int q = 0; q < p.n_parts; ++q) chunks += (p.part[q].K / (p.splits > 1 ? p.splits : 1) + kc - 1) / kc;
    if (chunks > 1) stages = 2;
  }
  g.tc_nmax = nmax;
  g.tc_stages = stages;
  int smem = stages * (x3 ? 2 : 1) * (BM * kc * 2 + nmax * (kc * 2));
  if (smem < 8 * 32 * 33 * 4) smem = 8 * 32 * 33 * 4;     // the epilogue's transposition buffers
  cudaError_t e = cudaFuncSetAttribute(x3 ? gemm_tc_group_kernel<1> : gemm_tc_group_kernel<0>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_tc_group_kernel)");
  if (x3)
    gemm_tc_group_kernel<1><<<cta, kTcThreads, smem, st>>>(g);
  else
    gemm_tc_group_kernel<0><<<cta, kTcThreads, smem, st>>>(g);
  DMT_CUDA_LAUNCH_CHECK("gemm_tc_group_kernel");
  if (red > 0) return reduce(g, st);
  return DMT_OK;
}

}  // namespace dmt
