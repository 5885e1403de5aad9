// TMA-fed tf32 tcgen05 GEMMs for the row-batched training pipeline (see gemm_tf32.cuh).
//
// Data path of one operand element: HBM --TMA (cp.async.bulk.tensor, SWIZZLE_128B)--> shared memory --tcgen05.mma
// kind::tf32--> TMEM (fp32 accumulate) --tcgen05.ld--> registers --> HBM.  Warp-specialised: warp 0 = TMA producer,
// warp 1 = MMA issuer (one elected thread each), the other warps = epilogue.
//
// Shared-memory operand images (what the descriptors below describe; cute/atom/mma_traits_sm100.hpp):
//   K-major,  SW128: a TMA box {32 fp32 (k), R rows}: row r at r * 128 B, 16-byte chunks XOR-swizzled by (r % 8);
//                    8-row groups are 1024 B apart (SBO); one MMA (K = 8) reads 32 B of every row, so successive
//                    K steps advance the start address by 32 B inside the swizzle span.
//   MN-major 32-bit operands have ONE legal swizzled layout, "128B swizzle with 32-byte atoms"
//   (UMMA LayoutType::SWIZZLE_128B_BASE32B = 1, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): a TMA box {32 fp32 (m or
//                    n), KT rows (k = tokens)}: row r at r * 128 B with its four 32-byte chunks XOR-swizzled by
//                    (r % 4); 4 k-rows form one 512-byte atom (SBO = 512 B to the next 4 k), the next 32 m/n
//                    elements are one box (LBO = KT * 128 B) further; one MMA (K = 8) consumes two atoms per 32 m/n,
//                    successive K steps advance the start address by 1024 B.
#include <cuda.h>

#include "gemm_tf32.cuh"
#include "umma.cuh"

namespace dmt {

using namespace umma;

namespace {

constexpr int RBM = 128;                 // rows of an accumulator tile (TMEM lanes)
constexpr int RKC = 32;                  // fp32 per 128-byte swizzle row
constexpr int kRowStage = RBM * 128;     // one A stage: 128 rows x 128 B
constexpr int kRowsThreads = 320;        // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int kRowsEpiWarps = 8;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

struct RowsArgs {
  CUtensorMap tmA;      // [M, K] fp32, box {32, 128}
  CUtensorMap tmB;      // [N, K] fp32, box {32, N}
  Tf32Rows p;
  int n_tiles, nkc, stages;
};

__global__ void __launch_bounds__(kRowsThreads, 1) tf32_rows_kernel(const __grid_constant__ RowsArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8], bfull, tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[256];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = g.p.N, K = g.p.K, S = g.stages, nkc = g.nkc;
  const uint32_t b_chunk = (uint32_t)N * 128u;             // one resident K chunk of Bt: N rows x 128 B
  const uint32_t sB = smem_u32(smem), sA = sB + (uint32_t)nkc * b_chunk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&bfull, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], kRowsEpiWarps);
    }
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmB) : "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  for (int i = threadIdx.x; i < 256; i += kRowsThreads) sbias[i] = (g.p.bias && i < N) ? __ldg(g.p.bias + i) : 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: Bt once, then the A tiles of this CTA's row tiles =====
      mbar_expect_tx(&bfull, (uint32_t)nkc * b_chunk);
      for (int kc = 0; kc < nkc; ++kc) tma_load_2d(sB + kc * b_chunk, &g.tmB, kc * RKC, 0, &bfull);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x)
        for (int kc = 0; kc < nkc; ++kc, ++it) {
          const int s = it % S;
          mbar_wait(&empty_bar[s], ((it / S) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], kRowStage);
          tma_load_2d(sA + s * kRowStage, &g.tmA, kc * RKC, tile * RBM, &full_bar[s]);
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = make_idesc_tf32(RBM, N);
      mbar_wait(&bfull, 0);
      fence_after_sync();
      uint32_t it = 0, j = 0;
      for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++j) {
        const uint32_t buf = j & 1;
        mbar_wait(&tempty[buf], ((j >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator
        fence_after_sync();
        for (int kc = 0; kc < nkc; ++kc, ++it) {
          const int s = it % S;
          mbar_wait(&full_bar[s], (it / S) & 1);
          fence_after_sync();
          const uint32_t a0 = sA + s * kRowStage, b0 = sB + kc * b_chunk;
          const int rem = K - kc * RKC;
          const int ksteps = rem >= RKC ? RKC / 8 : (rem + 7) / 8;
          for (int k = 0; k < ksteps; ++k)
            mma_tf32_ss(tbase + buf * 256, make_smem_desc(a0 + k * 32, 16, 1024, kLayoutSW128),
                        make_smem_desc(b0 + k * 32, 16, 1024, kLayoutSW128), idesc, (kc | k) != 0);
          commit(&empty_bar[s]);                           // frees the stage once these MMAs have read it
        }
        commit(&tfull[buf]);                               // accumulator of this tile complete
      }
    }
  } else {
    // ===== epilogue: TMEM -> (+addend) * alpha + bias, relu, mask, (+C) -> global.  Warps q and q + 4 share TMEM
    //       lane quarter q and take alternate 32-column chunks; the addend / mask / old-C values of a chunk are
    //       requested BEFORE the accumulator wait so their latency overlaps the MMAs. =====
    const int ew = warp - 2, q = warp & 3, half = ew >> 2;
    const Tf32Rows& p = g.p;
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++j) {
      const uint32_t buf = j & 1;
      const int64_t m = (int64_t)tile * RBM + q * 32 + lane;
      const bool ok = m < p.M;
      bool waited = false;
      for (int c0 = half * 32; c0 < N; c0 += 64) {
        const int w = N - c0 < 32 ? N - c0 : 32;          // 32, or a 16-column tail
        float4 ad[8], mk[8], old[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          ad[v] = make_float4(0.f, 0.f, 0.f, 0.f);
          mk[v] = make_float4(1.f, 1.f, 1.f, 1.f);
          old[v] = ad[v];
          if (ok && v * 4 < w) {
            if (p.addend) ad[v] = ld_stream4(p.addend + m * p.ld_add + c0 + v * 4);
            if (p.mask) mk[v] = ld_stream4(p.mask + m * p.ld_mask + c0 + v * 4);
            if (p.accumulate) old[v] = *reinterpret_cast<const float4*>(p.C + m * p.ldc + c0 + v * 4);
          }
        }
        if (!waited) {
          mbar_wait(&tfull[buf], (j >> 1) & 1);
          fence_after_sync();
          waited = true;
        }
        uint32_t r[32];
        if (w == 32) {
          tmem_ld32(tmem_addr(tbase, buf * 256 + c0), r);
        } else {
          tmem_ld16(tmem_addr(tbase, buf * 256 + c0), r);
#pragma unroll
          for (int e = 16; e < 32; ++e) r[e] = 0u;
        }
        tmem_ld_wait();
        if (ok) {
          float* crow = p.C + m * p.ldc + c0;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            if (v * 4 < w) {
              const float a4[4] = {ad[v].x, ad[v].y, ad[v].z, ad[v].w};
              const float m4[4] = {mk[v].x, mk[v].y, mk[v].z, mk[v].w};
              const float o4[4] = {old[v].x, old[v].y, old[v].z, old[v].w};
              float y[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float t = (__uint_as_float(r[v * 4 + e]) + a4[e]) * p.alpha + sbias[c0 + v * 4 + e];
                if (p.relu) t = fmaxf(t, 0.f);
                if (!(m4[e] > 0.f)) t = 0.f;
                y[e] = t + o4[e];
              }
              *reinterpret_cast<float4*>(crow + v * 4) = make_float4(y[0], y[1], y[2], y[3]);
            }
          }
        }
      }
      if (!waited) {                                       // a half without columns still takes part in the hand-over
        mbar_wait(&tfull[buf], (j >> 1) & 1);
        fence_after_sync();
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradients: D[128 x NB] (one M tile) over one token range per CTA.
constexpr int WKT = 64;                  // tokens per stage (8 MMAs of K = 8)
constexpr int kWgradThreads = 192;       // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue

struct WgradArgs {
  CUtensorMap tmP;      // [T, MA] fp32, box {32, WKT}
  CUtensorMap tmQ;      // [T, NB] fp32, box {32, WKT}
  int64_t T, tps;       // tokens per split (multiple of WKT)
  int MA, NB, m_tiles, splits, stages;
  float* partial;       // [splits][m_tiles * 128][NB]
  float* partial_cs;    // non-null: [splits][m_tiles * 128] column sums of P (one extra MMA per K step)
};

__global__ void __launch_bounds__(kWgradThreads, 1) tf32_wgrad_kernel(const __grid_constant__ WgradArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8], accum_bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, mt = blockIdx.y, S = g.stages, NB = g.NB;
  const int nbq = (NB + 31) / 32;
  constexpr uint32_t kBox = WKT * 128;                     // one {32, WKT} box
  const uint32_t stage_bytes = (4 + nbq) * kBox;
  const uint32_t s0 = smem_u32(smem);
  const int64_t t0 = (int64_t)split * g.tps;
  int64_t t1 = t0 + g.tps;
  if (t1 > g.T) t1 = g.T;
  const int nkb = t1 > t0 ? (int)((t1 - t0 + WKT - 1) / WKT) : 0;
  const bool want_cs = g.partial_cs != nullptr;
  const int cs_col = nbq * 32;                             // TMEM column of the column-sum accumulator
  uint32_t ncols = 32;
  while ((int)ncols < (want_cs ? cs_col + 16 : NB)) ncols <<= 1;
  const uint32_t ones = s0 + S * stage_bytes;              // 1 KB of 1.0f: 8 k-rows x 128 B (any swizzle of ones is ones)
  if (want_cs) {
    for (int i = threadIdx.x; i < 256; i += kWgradThreads) reinterpret_cast<float*>(smem + S * stage_bytes)[i] = 1.0f;
    fence_proxy_async();
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmP) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmQ) : "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, ncols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        mbar_wait(&empty_bar[s], ((kb / S) & 1) ^ 1);
        const uint32_t sp = s0 + s * stage_bytes, sq = sp + 4 * kBox;
        const int tok = (int)(t0 + (int64_t)kb * WKT);
        mbar_expect_tx(&full_bar[s], stage_bytes);
        for (int i = 0; i < 4; ++i) tma_load_2d(sp + i * kBox, &g.tmP, mt * RBM + i * 32, tok, &full_bar[s]);
        for (int i = 0; i < nbq; ++i) tma_load_2d(sq + i * kBox, &g.tmQ, i * 32, tok, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(RBM, NB, true, true), idesc_cs = make_idesc_tf32(RBM, 16, true, true);
      const uint64_t d_ones = make_smem_desc(ones, kBox, 512, kLayoutSW128Base32B);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        mbar_wait(&full_bar[s], (kb / S) & 1);
        fence_after_sync();
        const uint32_t sp = s0 + s * stage_bytes, sq = sp + 4 * kBox;
#pragma unroll
        for (int k = 0; k < WKT / 8; ++k) {
          const uint64_t dp = make_smem_desc(sp + k * 1024, kBox, 512, kLayoutSW128Base32B);
          mma_tf32_ss(tbase, dp, make_smem_desc(sq + k * 1024, kBox, 512, kLayoutSW128Base32B), idesc, (kb | k) != 0);
          if (want_cs) mma_tf32_ss(tbase + cs_col, dp, d_ones, idesc_cs, (kb | k) != 0);   // sum_t P[t, m] * 1
        }
        commit(&empty_bar[s]);
      }
      commit(&accum_bar);
    }
  } else {
    mbar_wait(&accum_bar, 0);
    fence_after_sync();
    const int row = (warp & 3) * 32 + lane;
    const int m = mt * RBM + row;
    float* prow = g.partial + ((int64_t)split * g.m_tiles * RBM + m) * NB;
    if (want_cs) {
      uint32_t r8[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      if (nkb > 0) {
        tmem_ld8(tmem_addr(tbase, cs_col), r8);
        tmem_ld_wait();
      }
      if (m < g.MA) g.partial_cs[(int64_t)split * g.m_tiles * RBM + m] = __uint_as_float(r8[0]);
    }
    for (int c0 = 0; c0 < NB; c0 += 32) {
      const int w = NB - c0 < 32 ? NB - c0 : 32;
      uint32_t r[32];
      if (nkb > 0) {
        if (w == 32) {
          tmem_ld32(tmem_addr(tbase, c0), r);
        } else {
          tmem_ld16(tmem_addr(tbase, c0), r);
#pragma unroll
          for (int e = 16; e < 32; ++e) r[e] = 0u;
        }
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = 0u;
      }
      if (m < g.MA) {
#pragma unroll
        for (int v = 0; v < 8; ++v)
          if (v * 4 < w)
            *reinterpret_cast<float4*>(prow + c0 + v * 4) =
                make_float4(__uint_as_float(r[v * 4]), __uint_as_float(r[v * 4 + 1]), __uint_as_float(r[v * 4 + 2]),
                            __uint_as_float(r[v * 4 + 3]));
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, ncols);
}

// ---------------------------------------------------------------------------------------------------------------
// General tiled GEMM: one 128 x BN tile per CTA, K streamed in chunks of 32 through a TMA ring.
constexpr int kGemmThreadsG = 192;       // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int GKC = 32;                  // K per stage

struct GemmArgsG {
  CUtensorMap tmA[4], tmB[4];
  Tf32Gemm p;
  int BN, stages, nkc;
};

__global__ void __launch_bounds__(kGemmThreadsG, 1) tf32_gemm_kernel(const __grid_constant__ GemmArgsG g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8], accum_bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Tf32Gemm& p = g.p;
  const int BN = g.BN, S = g.stages, nkc = g.nkc;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * RBM;
  const int z0 = p.reduce_z ? 0 : blockIdx.z, nzl = p.reduce_z ? p.nz : 1;   // problems this CTA walks
  const int total = nzl * nkc;
  const uint32_t a_bytes = RBM * 128, b_bytes = (uint32_t)BN * 128, stage_bytes = a_bytes + b_bytes;
  const uint32_t s0 = smem_u32(smem);
  uint32_t ncols = 32;
  while ((int)ncols < BN) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, ncols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < total; ++it) {
        const int z = z0 + it / nkc, kc = it % nkc, s = it % S;
        mbar_wait(&empty_bar[s], ((it / S) & 1) ^ 1);
        const uint32_t sa = s0 + s * stage_bytes, sb = sa + a_bytes;
        mbar_expect_tx(&full_bar[s], stage_bytes);
        if (p.a_mn) {                       // [K, M] storage: four {32 m, 32 k} boxes
          for (int i = 0; i < 4; ++i) tma_load_2d(sa + i * 4096, &g.tmA[z], m0 + i * 32, kc * GKC, &full_bar[s]);
        } else {                            // [M, K] storage: one {32 k, 128 m} box
          tma_load_2d(sa, &g.tmA[z], kc * GKC, m0, &full_bar[s]);
        }
        if (p.b_mn) {
          for (int i = 0; i < BN / 32; ++i) tma_load_2d(sb + i * 4096, &g.tmB[z], n0 + i * 32, kc * GKC, &full_bar[s]);
        } else {
          tma_load_2d(sb, &g.tmB[z], kc * GKC, n0, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(RBM, BN, p.a_mn != 0, p.b_mn != 0);
      for (int it = 0; it < total; ++it) {
        const int kc = it % nkc, s = it % S;
        mbar_wait(&full_bar[s], (it / S) & 1);
        fence_after_sync();
        const uint32_t sa = s0 + s * stage_bytes, sb = sa + a_bytes;
        const int rem = p.K - kc * GKC;
        const int ksteps = rem >= GKC ? GKC / 8 : (rem + 7) / 8;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t da = p.a_mn ? make_smem_desc(sa + k * 1024, 4096, 512, kLayoutSW128Base32B)
                                     : make_smem_desc(sa + k * 32, 16, 1024, kLayoutSW128);
          const uint64_t db = p.b_mn ? make_smem_desc(sb + k * 1024, 4096, 512, kLayoutSW128Base32B)
                                     : make_smem_desc(sb + k * 32, 16, 1024, kLayoutSW128);
          mma_tf32_ss(tbase, da, db, idesc, (it | k) != 0);
        }
        commit(&empty_bar[s]);
      }
      commit(&accum_bar);
    }
  } else {
    mbar_wait(&accum_bar, 0);
    fence_after_sync();
    const int zc = p.reduce_z ? 0 : blockIdx.z;
    const int64_t m = (int64_t)m0 + (warp & 3) * 32 + lane;
    const bool ok = m < p.M;
    float* crow = p.C[zc] + m * p.ldc;
    const float* mrow = p.mask[zc] ? p.mask[zc] + m * p.ld_mask : nullptr;
    const float* bias = p.bias[zc];
    const bool vec = ((((uintptr_t)p.C[zc]) & 15) == 0) && (p.ldc % 4 == 0) &&
                     (!p.mask[zc] || (((((uintptr_t)p.mask[zc]) & 15) == 0) && p.ld_mask % 4 == 0));
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= p.N) break;                               // warp-uniform
      uint32_t r[32];
      tmem_ld32(tmem_addr(tbase, c0), r);
      tmem_ld_wait();
      if (!ok) continue;
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const int n = n0 + c0 + v * 4;
        if (n >= p.N) break;
        float y[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float t = __uint_as_float(r[v * 4 + e]) + ((bias && n + e < p.N) ? __ldg(bias + n + e) : 0.f);
          if (p.relu) t = fmaxf(t, 0.f);
          y[e] = t;
        }
        if (vec && n + 4 <= p.N) {
          if (mrow) {
            const float4 mk = ld_stream4(mrow + n);
            if (!(mk.x > 0.f)) y[0] = 0.f;
            if (!(mk.y > 0.f)) y[1] = 0.f;
            if (!(mk.z > 0.f)) y[2] = 0.f;
            if (!(mk.w > 0.f)) y[3] = 0.f;
          }
          float4 o = make_float4(y[0], y[1], y[2], y[3]);
          if (p.accumulate) {
            const float4 old = *reinterpret_cast<const float4*>(crow + n);
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *reinterpret_cast<float4*>(crow + n) = o;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (n + e < p.N) {
              float t = y[e];
              if (mrow && !(__ldg(mrow + n + e) > 0.f)) t = 0.f;
              crow[n + e] = (p.accumulate ? crow[n + e] : 0.f) + t;
            }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, ncols);
}

struct WgradReduceArgs {
  const float* partial;
  const float* partial_cs;
  Tf32WgradSeg seg[4];
  int n_seg, MA, NB, m_tiles, splits, transposed, accumulate;
};

// fixed-order sum over the splits (deterministic), scattered into the weight tensors
__global__ void __launch_bounds__(256) tf32_wgrad_reduce_kernel(const __grid_constant__ WgradReduceArgs a) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (a.partial_cs && i >= a.MA * a.NB && i < a.MA * a.NB + a.MA) {      // the column sums (bias gradients)
    const int m = i - a.MA * a.NB;
    float s = 0.f;
    for (int k = 0; k < a.splits; ++k) s += a.partial_cs[(int64_t)k * a.m_tiles * RBM + m];
    for (int q = 0; q < a.n_seg; ++q) {
      const Tf32WgradSeg& sg = a.seg[q];
      if (m >= sg.m0 && m < sg.m1 && sg.colsum) sg.colsum[m - sg.m0] = (a.accumulate ? sg.colsum[m - sg.m0] : 0.f) + s;
    }
    return;
  }
  if (i >= a.MA * a.NB) return;
  const int m = i / a.NB, n = i - m * a.NB;
  const int64_t stride = (int64_t)a.m_tiles * RBM * a.NB;
  const float* p = a.partial + (int64_t)m * a.NB + n;
  float s = 0.f;
  for (int k = 0; k < a.splits; ++k) s += p[k * stride];
  for (int q = 0; q < a.n_seg; ++q) {
    const Tf32WgradSeg& sg = a.seg[q];
    if (m >= sg.m0 && m < sg.m1) {
      float* dst = a.transposed ? sg.C + (int64_t)n * sg.ldc + (m - sg.m0) : sg.C + (int64_t)(m - sg.m0) * sg.ldc + n;
      *dst = (a.accumulate ? *dst : 0.f) + s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Column sums (bias gradients): stage 1 = per-CTA sums over a contiguous row range, stage 2 = fixed-order reduce.
constexpr int kColsumCtas = 296;

__global__ void __launch_bounds__(256) tf32_colsum_kernel(const float* __restrict__ X, int64_t ldx, int64_t T, int W,
                                                          float* __restrict__ partial) {
  __shared__ float4 red[256];
  const int w4 = W >> 2;                       // float4 columns (W % 4 == 0)
  const int rows_par = 256 / w4;               // rows handled in parallel by one CTA
  const int c = threadIdx.x % w4, rsub = threadIdx.x / w4;
  const int64_t per = (T + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per;
  int64_t r1 = r0 + per;
  if (r1 > T) r1 = T;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rsub < rows_par)
    for (int64_t r = r0 + rsub; r < r1; r += rows_par) {
      const float4 v = ld_stream4(X + r * ldx + c * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < w4) {
    float4 s = red[threadIdx.x];
    for (int k = 1; k < rows_par; ++k) {
      const float4 v = red[k * w4 + threadIdx.x];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(partial + (int64_t)blockIdx.x * W + threadIdx.x * 4) = s;
  }
}

// one warp per column: lanes stride the per-CTA partials, then a fixed-order shuffle tree (deterministic)
__global__ void __launch_bounds__(256) tf32_colsum_reduce_kernel(const float* __restrict__ partial, int n_part, int W,
                                                                 float* __restrict__ out, int accumulate) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= W) return;
  float s = 0.f;
  for (int k = lane; k < n_part; k += 32) s += partial[(int64_t)k * W + c];
  s = warp_sum(s);
  if (lane == 0) out[c] = (accumulate ? out[c] : 0.f) + s;
}

struct PackArgs {
  Tf32PackMat m[4];
  Tf32PackVec v[4];
  int n_mats, n_vecs;
  float* dst;
  int64_t ldd;
  float* dstv;
};

// blockIdx.y = source; consecutive threads walk the SOURCE's contiguous dimension (coalesced reads; the destination
// is a few hundred KB at most and L2-resident)
__global__ void __launch_bounds__(256) tf32_pack_kernel(const __grid_constant__ PackArgs a) {
  const int which = blockIdx.y;
  if (which < a.n_mats) {
    const Tf32PackMat& s = a.m[which];
    const int total = s.rows * s.cols;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
      int r, c;
      if (s.transpose) {            // source is [cols, rows]: i = c * rows + r
        c = i / s.rows;
        r = i - c * s.rows;
        a.dst[(int64_t)(s.dst_row + r) * a.ldd + s.dst_col + c] = __ldg(s.W + (int64_t)c * s.ldw + r);
      } else {
        r = i / s.cols;
        c = i - r * s.cols;
        a.dst[(int64_t)(s.dst_row + r) * a.ldd + s.dst_col + c] = __ldg(s.W + (int64_t)r * s.ldw + c);
      }
    }
  } else if (which - a.n_mats < a.n_vecs) {
    const Tf32PackVec& s = a.v[which - a.n_mats];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < s.n; i += gridDim.x * 256)
      a.dstv[s.dst + i] = s.v ? __ldg(s.v + i) : 0.f;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 row-major [rows, cols] (row stride ld floats), box {32 cols, box_rows}, 128-byte swizzle, OOB -> 0
int make_map_f32(CUtensorMap* tm, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                 bool mn_major = false) {
  EncodeTiledFn fn = encode_fn();
  DMT_REQUIRE(fn, DMT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  DMT_REQUIRE(((uintptr_t)base & 15) == 0 && ld % 4 == 0 && rows > 0 && cols > 0 && box_rows > 0 && box_rows <= 256,
              DMT_ERR_INVALID_ARGUMENT, "tf32 GEMM operand: base %p ld %lld rows %lld cols %lld box_rows %d (16-byte "
              "aligned base and row stride required)", (const void*)base, (long long)ld, (long long)rows,
              (long long)cols, box_rows);
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {RKC, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DMT_REQUIRE(r == CUDA_SUCCESS, DMT_ERR_CUDA, "cuTensorMapEncodeTiled(f32) failed (%d) rows=%lld cols=%lld ld=%lld",
              (int)r, (long long)rows, (long long)cols, (long long)ld);
  return DMT_OK;
}

}  // namespace

int make_map_f32_public(CUtensorMap* tm, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                        bool mn_major) {
  return make_map_f32(tm, base, rows, cols, ld, box_rows, mn_major);
}

int tf32_pack(const Tf32PackMat* mats, int n_mats, float* dst, int64_t ldd, const Tf32PackVec* vecs, int n_vecs,
              float* dstv, cudaStream_t st) {
  DMT_REQUIRE(n_mats >= 0 && n_mats <= 4 && n_vecs >= 0 && n_vecs <= 4 && (n_mats == 0 || dst) && (n_vecs == 0 || dstv),
              DMT_ERR_INVALID_ARGUMENT, "tf32_pack: %d matrices, %d vectors", n_mats, n_vecs);
  if (n_mats + n_vecs == 0) return DMT_OK;
  PackArgs a{};
  int biggest = 1;
  for (int i = 0; i < n_mats; ++i) {
    a.m[i] = mats[i];
    if (mats[i].rows * mats[i].cols > biggest) biggest = mats[i].rows * mats[i].cols;
  }
  for (int i = 0; i < n_vecs; ++i) a.v[i] = vecs[i];
  a.n_mats = n_mats;
  a.n_vecs = n_vecs;
  a.dst = dst;
  a.ldd = ldd;
  a.dstv = dstv;
  int bx = (biggest + 255) / 256;
  if (bx > 64) bx = 64;
  tf32_pack_kernel<<<dim3(bx, n_mats + n_vecs), 256, 0, st>>>(a);
  DMT_CUDA_LAUNCH_CHECK("tf32_pack_kernel");
  return DMT_OK;
}

int tf32_gemm(const Tf32Gemm& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || p.nz <= 0) return DMT_OK;
  DMT_REQUIRE(p.nz <= 4 && p.K > 0 && p.lda % 4 == 0 && p.ldb % 4 == 0, DMT_ERR_INVALID_ARGUMENT,
              "tf32_gemm: nz=%d K=%d lda=%lld ldb=%lld", p.nz, p.K, (long long)p.lda, (long long)p.ldb);
  DMT_REQUIRE(p.M < ((int64_t)1 << 31) - RBM, DMT_ERR_UNSUPPORTED_SHAPE, "tf32_gemm: M=%lld", (long long)p.M);
  GemmArgsG g;
  g.p = p;
  int BN = ((p.N + 31) / 32) * 32;           // a tile is a whole number of 32-column boxes
  if (BN > 256) BN = 256;
  if (BN == 96 || BN == 160 || BN == 224) BN += 32;   // keep N of the MMA a multiple of 64 / power-of-two friendly
  g.BN = BN;
  g.nkc = (p.K + GKC - 1) / GKC;
  const int stage_bytes = RBM * 128 + BN * 128;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  g.stages = stages;
  for (int z = 0; z < p.nz; ++z) {
    DMT_REQUIRE(p.A[z] && p.B[z] && (p.reduce_z ? (z > 0 || p.C[0]) : (p.C[z] != nullptr)), DMT_ERR_INVALID_ARGUMENT,
                "tf32_gemm: problem %d is incomplete", z);
    int rc = p.a_mn ? make_map_f32(&g.tmA[z], p.A[z], p.K, p.M, p.lda, GKC, true)
                    : make_map_f32(&g.tmA[z], p.A[z], p.M, p.K, p.lda, RBM, false);
    if (rc != DMT_OK) return rc;
    rc = p.b_mn ? make_map_f32(&g.tmB[z], p.B[z], p.K, p.N, p.ldb, GKC, true)
                : make_map_f32(&g.tmB[z], p.B[z], p.N, p.K, p.ldb, BN, false);
    if (rc != DMT_OK) return rc;
  }
  for (int z = p.nz; z < 4; ++z) {
    g.tmA[z] = g.tmA[0];
    g.tmB[z] = g.tmB[0];
  }
  const int smem = stages * stage_bytes + 1024;
  cudaError_t e = cudaFuncSetAttribute(tf32_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tf32_gemm_kernel)");
  dim3 grid((p.N + BN - 1) / BN, (unsigned)((p.M + RBM - 1) / RBM), p.reduce_z ? 1 : p.nz);
  tf32_gemm_kernel<<<grid, kGemmThreadsG, smem, st>>>(g);
  DMT_CUDA_LAUNCH_CHECK("tf32_gemm_kernel");
  return DMT_OK;
}

int tf32_rows(const Tf32Rows& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0) return DMT_OK;
  if (p.N > 256) {            // wider outputs run as column blocks of <= 256 (each re-reads A)
    DMT_REQUIRE(p.N % 16 == 0, DMT_ERR_UNSUPPORTED_SHAPE, "tf32_rows: N=%d must be a multiple of 16", p.N);
    const int blocks = (p.N + 255) / 256;
    const int per = ((p.N / 16 + blocks - 1) / blocks) * 16;
    for (int n0 = 0; n0 < p.N; n0 += per) {
      Tf32Rows q = p;
      q.N = p.N - n0 < per ? p.N - n0 : per;
      q.Bt = p.Bt + (int64_t)n0 * p.ldb;
      q.C = p.C + n0;
      if (p.bias) q.bias = p.bias + n0;
      if (p.addend) q.addend = p.addend + n0;
      if (p.mask) q.mask = p.mask + n0;
      const int rc = tf32_rows(q, st);
      if (rc != DMT_OK) return rc;
    }
    return DMT_OK;
  }
  DMT_REQUIRE(p.N % 16 == 0 && p.N <= 256 && p.K > 0 && p.K % 4 == 0, DMT_ERR_UNSUPPORTED_SHAPE,
              "tf32_rows: N=%d (multiple of 16, <= 256) K=%d (multiple of 4)", p.N, p.K);
  DMT_REQUIRE(p.A && p.Bt && p.C && p.ldc % 4 == 0 && ((uintptr_t)p.C & 15) == 0, DMT_ERR_INVALID_ARGUMENT,
              "tf32_rows: bad output / operands");
  DMT_REQUIRE((!p.addend || (p.ld_add % 4 == 0 && ((uintptr_t)p.addend & 15) == 0)) &&
                  (!p.mask || (p.ld_mask % 4 == 0 && ((uintptr_t)p.mask & 15) == 0)),
              DMT_ERR_INVALID_ARGUMENT, "tf32_rows: addend / mask must be 16-byte aligned with ld %% 4 == 0");
  DMT_REQUIRE(p.M < ((int64_t)1 << 31) - RBM, DMT_ERR_UNSUPPORTED_SHAPE, "tf32_rows: M=%lld", (long long)p.M);
  RowsArgs g;
  g.p = p;
  g.nkc = (p.K + RKC - 1) / RKC;
  g.n_tiles = (int)((p.M + RBM - 1) / RBM);
  const int b_bytes = g.nkc * p.N * 128;
  int stages = (220 * 1024 - b_bytes) / kRowStage;
  if (stages > 8) stages = 8;
  DMT_REQUIRE(stages >= 2, DMT_ERR_UNSUPPORTED_SHAPE, "tf32_rows: N=%d K=%d weights do not fit shared memory", p.N, p.K);
  g.stages = stages;
  int rc = make_map_f32(&g.tmA, p.A, p.M, p.K, p.lda, RBM);
  if (rc != DMT_OK) return rc;
  rc = make_map_f32(&g.tmB, p.Bt, p.N, p.K, p.ldb, p.N);
  if (rc != DMT_OK) return rc;
  const int smem = b_bytes + stages * kRowStage + 1024;
  cudaError_t e = cudaFuncSetAttribute(tf32_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tf32_rows_kernel)");
  const int sms = sm_count_cached();
  const int grid = g.n_tiles < sms ? g.n_tiles : sms;
  tf32_rows_kernel<<<grid, kRowsThreads, smem, st>>>(g);
  DMT_CUDA_LAUNCH_CHECK("tf32_rows_kernel");
  return DMT_OK;
}

static void wgrad_plan(int64_t T, int MA, int NB, int* m_tiles, int* splits, int64_t* tps) {
  const int mt = (MA + RBM - 1) / RBM;
  int s = sm_count_cached() / (mt > 0 ? mt : 1);
  const int64_t kmax = (T + 4 * WKT - 1) / (4 * WKT);        // at least 4 stages of tokens per split
  if (s > kmax) s = (int)kmax;
  if (s < 1) s = 1;
  int64_t per = (T + s - 1) / s;
  per = (per + WKT - 1) / WKT * WKT;
  if (per < WKT) per = WKT;
  s = (int)((T + per - 1) / per);
  if (s < 1) s = 1;
  *m_tiles = mt;
  *splits = s;
  *tps = per;
  (void)NB;
}

size_t tf32_wgrad_partial_bytes(int64_t T, int MA, int NB) {
  int mt, s;
  int64_t per;
  wgrad_plan(T, MA, NB, &mt, &s, &per);
  return (((size_t)s * mt * RBM * (NB + 1) * sizeof(float) + 255) & ~(size_t)255) + 256;   // + column-sum partials
}

int tf32_wgrad(const Tf32Wgrad& p, cudaStream_t st) {
  if (p.MA <= 0 || p.NB <= 0) return DMT_OK;
  DMT_REQUIRE(p.NB % 16 == 0 && p.NB <= 256 && p.MA % 4 == 0, DMT_ERR_UNSUPPORTED_SHAPE,
              "tf32_wgrad: NB=%d (multiple of 16, <= 256) MA=%d (multiple of 4)", p.NB, p.MA);
  DMT_REQUIRE(p.P && p.Q && p.partial && p.n_seg >= 1 && p.n_seg <= 4 && p.T >= 0 && p.T < ((int64_t)1 << 31) - WKT,
              DMT_ERR_INVALID_ARGUMENT, "tf32_wgrad: bad arguments");
  WgradArgs g;
  wgrad_plan(p.T, p.MA, p.NB, &g.m_tiles, &g.splits, &g.tps);
  g.T = p.T;
  g.MA = p.MA;
  g.NB = p.NB;
  g.partial = p.partial;
  bool want_cs = false;
  for (int i = 0; i < p.n_seg; ++i) want_cs = want_cs || p.seg[i].colsum != nullptr;
  g.partial_cs = want_cs ? p.partial + (((size_t)g.splits * g.m_tiles * RBM * p.NB + 63) & ~(size_t)63) : nullptr;
  const int nbq = (p.NB + 31) / 32;
  DMT_REQUIRE(!want_cs || nbq * 32 + 16 <= 512, DMT_ERR_UNSUPPORTED_SHAPE, "tf32_wgrad: NB=%d with column sums", p.NB);
  const int stage_bytes = (4 + nbq) * WKT * 128;
  int stages = (219 * 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  DMT_REQUIRE(stages >= 2, DMT_ERR_UNSUPPORTED_SHAPE, "tf32_wgrad: NB=%d stage does not fit", p.NB);
  g.stages = stages;
  if (p.T > 0) {
    int rc = make_map_f32(&g.tmP, p.P, p.T, p.MA, p.ldp, WKT, true);
    if (rc != DMT_OK) return rc;
    rc = make_map_f32(&g.tmQ, p.Q, p.T, p.NB, p.ldq, WKT, true);
    if (rc != DMT_OK) return rc;
    const int smem = stages * stage_bytes + 1024 + 1024;     // + the block of ones
    cudaError_t e = cudaFuncSetAttribute(tf32_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tf32_wgrad_kernel)");
    tf32_wgrad_kernel<<<dim3(g.splits, g.m_tiles), kWgradThreads, smem, st>>>(g);
    DMT_CUDA_LAUNCH_CHECK("tf32_wgrad_kernel");
  } else {
    g.splits = 0;                                          // no tokens: the reduction writes / keeps zeros
  }
  WgradReduceArgs r;
  r.partial = p.partial;
  r.partial_cs = g.partial_cs;
  for (int i = 0; i < 4; ++i) r.seg[i] = p.seg[i < p.n_seg ? i : 0];
  r.n_seg = p.n_seg;
  r.MA = p.MA;
  r.NB = p.NB;
  r.m_tiles = g.m_tiles;
  r.splits = g.splits;
  r.transposed = p.transposed;
  r.accumulate = p.accumulate;
  tf32_wgrad_reduce_kernel<<<(p.MA * p.NB + (want_cs ? p.MA : 0) + 255) / 256, 256, 0, st>>>(r);
  DMT_CUDA_LAUNCH_CHECK("tf32_wgrad_reduce_kernel");
  return DMT_OK;
}

size_t tf32_colsum_scratch_bytes(int W) { return ((size_t)kColsumCtas * W * sizeof(float) + 255) & ~(size_t)255; }

int tf32_colsum(const float* X, int64_t ldx, int64_t T, int W, float* out, int accumulate, float* scratch,
                cudaStream_t st) {
  DMT_REQUIRE((X || T == 0) && out && scratch && W > 0 && W % 4 == 0 && W <= 1024 && ldx % 4 == 0 &&
                  ((uintptr_t)X & 15) == 0 && T >= 0,
              DMT_ERR_INVALID_ARGUMENT, "tf32_colsum: W=%d ldx=%lld", W, (long long)ldx);
  int ctas = kColsumCtas;
  if (T < ctas) ctas = T > 0 ? (int)T : 1;
  tf32_colsum_kernel<<<ctas, 256, 0, st>>>(X, ldx, T, W, scratch);
  DMT_CUDA_LAUNCH_CHECK("tf32_colsum_kernel");
  tf32_colsum_reduce_kernel<<<(W + 7) / 8, 256, 0, st>>>(scratch, ctas, W, out, accumulate);
  DMT_CUDA_LAUNCH_CHECK("tf32_colsum_reduce_kernel");
  return DMT_OK;
}

}  // namespace dmt

// ---- self-tests of the engine (tests/test_gpu_tf32.py): plain device pointers in, no model state ------------------
extern "C" {

int dmt_selftest_tf32_rows(const float* A, int64_t lda, const float* Bt, int64_t ldb, int64_t M, int32_t N, int32_t K,
                           float* C, int64_t ldc, const float* bias, const float* addend, int64_t ld_add,
                           const float* mask, int64_t ld_mask, float alpha, int32_t relu, int32_t accumulate,
                           void* stream) {
  dmt::Tf32Rows p{A, lda, Bt, ldb, M, N, K, C, ldc, bias, addend, ld_add, mask, ld_mask, alpha, relu, accumulate};
  return dmt::tf32_rows(p, (cudaStream_t)stream);
}

int dmt_selftest_tf32_gemm(const float* A, int64_t lda, int32_t a_mn, const float* B, int64_t ldb, int32_t b_mn,
                           int64_t M, int32_t N, int32_t K, float* C, int64_t ldc, const float* bias, const float* mask,
                           int64_t ld_mask, int32_t relu, int32_t accumulate, void* stream) {
  dmt::Tf32Gemm p{};
  p.A[0] = A; p.B[0] = B; p.lda = lda; p.ldb = ldb; p.a_mn = a_mn; p.b_mn = b_mn; p.nz = 1;
  p.M = M; p.N = N; p.K = K; p.C[0] = C; p.ldc = ldc; p.bias[0] = bias; p.mask[0] = mask; p.ld_mask = ld_mask;
  p.relu = relu; p.accumulate = accumulate;
  return dmt::tf32_gemm(p, (cudaStream_t)stream);
}

size_t dmt_selftest_tf32_wgrad_bytes(int64_t T, int32_t MA, int32_t NB) { return dmt::tf32_wgrad_partial_bytes(T, MA, NB); }

int dmt_selftest_tf32_wgrad(const float* P, int64_t ldp, const float* Q, int64_t ldq, int64_t T, int32_t MA, int32_t NB,
                            float* C, int64_t ldc, int32_t transposed, int32_t accumulate, float* colsum,
                            void* workspace, void* stream) {
  dmt::Tf32Wgrad p{};
  p.P = P; p.ldp = ldp; p.Q = Q; p.ldq = ldq; p.T = T; p.MA = MA; p.NB = NB;
  p.seg[0] = dmt::Tf32WgradSeg{C, ldc, 0, MA, colsum};
  p.n_seg = 1;
  p.transposed = transposed;
  p.accumulate = accumulate;
  p.partial = (float*)workspace;
  return dmt::tf32_wgrad(p, (cudaStream_t)stream);
}

int dmt_selftest_tf32_colsum(const float* X, int64_t ldx, int64_t T, int32_t W, float* out, int32_t accumulate,
                             void* scratch, void* stream) {
  return dmt::tf32_colsum(X, ldx, T, W, out, accumulate, (float*)scratch, (cudaStream_t)stream);
}

}  // extern "C"
