// dmt_mmoe_fwd, DMT_PRECISION_BF16: the expert MLPs as TMA-fed tcgen05 GEMMs.
//
//   cast+gate kernel : x fp32 -> bf16 (TMA-able, ld padded to 8 elements) and, in the same pass over x,
//                      the 2 x 4 gate logits + softmax (mmoe_transformer_unbias.py:85-94)
//   gemm_tc_kernel   : C[z] = relu(A[z] Bt[z]^T + bias[z]) per expert z, bf16 in / fp32 accumulate in
//                      TMEM / bf16 out.  Warp-specialised: warp 0 = TMA producer (128B-swizzled 128x64
//                      tiles, 3-stage mbarrier ring), warp 1 = MMA issuer, warps 2-5 = epilogue
//   mmoe_head_kernel : gate mixture + task towers (shared with the fp32 path, mmoe_f32.cu)
#include <cuda.h>
#include <stdlib.h>

#include "dmt_common.cuh"
#include "umma.cuh"

namespace dmt {

using namespace umma;

constexpr int GBM = 128, GBN = 128, GBK = 64, GSTAGES = 3;   // 3 stages = 2 CTAs per SM (6 stages, 1 CTA: slower)
constexpr int kGemmThreads = 192;
constexpr int kStageBytes = (GBM * GBK + GBN * GBK) * 2;

struct GemmTcArgs {
  CUtensorMap tmA, tmB;        // A: [rows, K] bf16 row-major; Bt: [rows, K] bf16 row-major (K-major B operand)
  const float* bias;           // [z][N]
  __nv_bfloat16* C;            // [z][M][ldc]
  int32_t M, N, K;
  int32_t a_rows_per_z;        // row offset of expert z inside the A tensor map (0: shared input)
  int32_t b_rows_per_z;        // row offset of expert z inside the Bt tensor map
  int64_t ldc, c_stride_z;
  int32_t relu;
  // layer 0 of dmt_mmoe_fwd_bf16in: slab z == gate_z holds the n_tasks x n_experts gate kernels as extra B rows
  // (mmoe_transformer_unbias.py:85-94); its epilogue writes the gate softmaxes instead of an activation tile
  int32_t gate_z, n_tasks, n_experts;
  float* gates;                // [n_tasks][M][n_experts]
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__global__ void __launch_bounds__(kGemmThreads) gemm_tc_kernel(const __grid_constant__ GemmTcArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[GSTAGES], empty_bar[GSTAGES], accum_bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // the gate slab (if any) is scheduled FIRST (blockIdx.z == 0) so that its 128-row CTAs are not a tail wave of their
  // own; z = slab index in the B / bias / C layouts (experts 0 .. E-1, gate slab = E)
  const int n0 = blockIdx.x * GBN, m0 = blockIdx.y * GBM;
  const int z = g.gate_z < 0 ? (int)blockIdx.z : (blockIdx.z == 0 ? g.gate_z : (int)blockIdx.z - 1);
  const int nkb = (g.K + GBK - 1) / GBK;
  if (z == g.gate_z && n0 > 0) return;           // the gate slab is one (mostly empty) N tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < GSTAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmB) : "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, GBN);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  griddep_launch();      // (a dependent launch behind this kernel -- the fused tail -- may start its prologue)
  const uint32_t tbase = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GSTAGES;
        mbar_wait(&empty_bar[s], ((kb / GSTAGES) & 1) ^ 1);
        uint8_t* sa = smem + s * kStageBytes;
        uint8_t* sb = sa + GBM * GBK * 2;
        mbar_expect_tx(&full_bar[s], kStageBytes);
        tma_load_2d(sa, &g.tmA, kb * GBK, z * g.a_rows_per_z + m0, &full_bar[s]);
        tma_load_2d(sb, &g.tmB, kb * GBK, z * g.b_rows_per_z + n0, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc_bf16(GBM, GBN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GSTAGES;
        mbar_wait(&full_bar[s], (kb / GSTAGES) & 1);
        fence_after_sync();
        const uint32_t sa = smem_u32(smem + s * kStageBytes), sb = sa + GBM * GBK * 2;
#pragma unroll
        for (int k = 0; k < GBK / 16; ++k)
          mma_bf16_ss(tbase, make_smem_desc(sa + k * 32, 16, 1024, kLayoutSW128),
                      make_smem_desc(sb + k * 32, 16, 1024, kLayoutSW128), idesc, (kb | k) != 0);
        commit(&empty_bar[s]);                 // frees the stage once these MMAs have read it
      }
      commit(&accum_bar);                      // accumulator complete
    }
  } else {
    // ===== epilogue: TMEM -> +bias, relu -> bf16 -> global =====
    mbar_wait(&accum_bar, 0);
    fence_after_sync();
    const int row = (warp & 3) * 32 + lane;
    const int m = m0 + row;
    const float* __restrict__ bias = g.bias + (int64_t)z * g.N;
    if (z == g.gate_z) {
      // gates: columns [t * E, t * E + E) of this row are task t's logits
      uint32_t r[32];
      tmem_ld32(tmem_addr(tbase, 0), r);
      tmem_ld_wait();
      if (m < g.M) {
        const int E = g.n_experts;
        for (int t = 0; t < g.n_tasks; ++t) {
          float v[DMT_MAX_EXPERTS];
          float mx = -INFINITY;
#pragma unroll
          for (int e = 0; e < DMT_MAX_EXPERTS; ++e) {
            v[e] = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j)               // (register arrays want compile-time indices)
              if (e < E && j == t * E + e) v[e] = __uint_as_float(r[j]) + __ldg(bias + j);
            mx = fmaxf(mx, v[e]);
          }
          float den = 0.f;
#pragma unroll
          for (int e = 0; e < DMT_MAX_EXPERTS; ++e) {
            v[e] = e < E ? expf(v[e] - mx) : 0.f;
            den += v[e];
          }
#pragma unroll
          for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
            if (e < E) g.gates[((int64_t)t * g.M + m) * E + e] = v[e] / den;
        }
      }
    } else {
    __nv_bfloat16* crow = g.C + (int64_t)z * g.c_stride_z + (int64_t)m * g.ldc;
#pragma unroll
    for (int c0 = 0; c0 < GBN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_addr(tbase, c0), r);
      tmem_ld_wait();
      if (m < g.M) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const int n = n0 + c0 + j;
          if (n + 8 <= g.N) {
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              y[e] = __uint_as_float(r[j + e]) + __ldg(bias + n + e);
              if (g.relu) y[e] = fmaxf(y[e], 0.f);
            }
            uint4 v;
            v.x = pack_bf16x2(y[0], y[1]);
            v.y = pack_bf16x2(y[2], y[3]);
            v.z = pack_bf16x2(y[4], y[5]);
            v.w = pack_bf16x2(y[6], y[7]);
            *reinterpret_cast<uint4*>(crow + n) = v;
          } else {
            for (int e = 0; e < 8 && n + e < g.N; ++e) {
              float yv = __uint_as_float(r[j + e]) + __ldg(bias + n + e);
              if (g.relu) yv = fmaxf(yv, 0.f);
              crow[n + e] = __float2bfloat16(yv);
            }
          }
        }
      }
    }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, GBN);
}

// One warp per sample: x fp32 -> bf16 row (zero-padded to ldx) + gate softmax for every task.
__global__ void __launch_bounds__(256) mmoe_cast_gate_kernel(const float* __restrict__ x, int64_t x_ld, int B, int K,
                                                             __nv_bfloat16* __restrict__ xb, int ldxb,
                                                             dmt_dense g0, dmt_dense g1, dmt_dense g2, dmt_dense g3,
                                                             int n_tasks, int E, float* __restrict__ gates) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  if (b >= B) return;
  const dmt_dense gate[DMT_MAX_TASKS] = {g0, g1, g2, g3};
  const float* __restrict__ xr = x + (int64_t)b * x_ld;
  float acc[DMT_MAX_TASKS][DMT_MAX_EXPERTS];
#pragma unroll
  for (int t = 0; t < DMT_MAX_TASKS; ++t)
#pragma unroll
    for (int e = 0; e < DMT_MAX_EXPERTS; ++e) acc[t][e] = 0.f;
  for (int k = lane; k < ldxb; k += 32) {
    const float xv = k < K ? __ldg(xr + k) : 0.f;
    xb[(int64_t)b * ldxb + k] = __float2bfloat16(xv);
    if (k < K) {
#pragma unroll
      for (int t = 0; t < DMT_MAX_TASKS; ++t)
        if (t < n_tasks) {
#pragma unroll
          for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
            if (e < E) acc[t][e] = fmaf(xv, __ldg(gate[t].w + (int64_t)k * E + e), acc[t][e]);
        }
    }
  }
#pragma unroll
  for (int t = 0; t < DMT_MAX_TASKS; ++t)
    if (t < n_tasks) {
      float mx = -INFINITY;
#pragma unroll
      for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
        if (e < E) {
          acc[t][e] = warp_sum(acc[t][e]) + __ldg(gate[t].b + e);
          mx = fmaxf(mx, acc[t][e]);
        }
      float den = 0.f;
#pragma unroll
      for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
        if (e < E) {
          acc[t][e] = expf(acc[t][e] - mx);
          den += acc[t][e];
        }
      if (lane == 0)
        for (int e = 0; e < E; ++e) gates[((int64_t)t * B + b) * E + e] = acc[t][e] / den;
    }
}


// Fast path for 4 experts (dmt.conf): the gate kernels [K, 4] of every task are staged in shared memory once per
// CTA, warps stride over the samples, every lane converts 4 consecutive inputs per trip (float4 in, 4 x bf16 out,
// one float4 of gate weights per (task, input)).
constexpr int kCastWarps = 16;
__global__ void __launch_bounds__(kCastWarps * 32) mmoe_cast_gate4_kernel(const float* __restrict__ x, int64_t x_ld, int B,
                                                                          int K, __nv_bfloat16* __restrict__ xb, int ldxb,
                                                                          dmt_dense g0, dmt_dense g1, dmt_dense g2,
                                                                          dmt_dense g3, int n_tasks,
                                                                          float* __restrict__ gates) {
  // [task][K128] float4 = the 4 expert weights of input k, permuted inside every block of 128 inputs so that the
  // lanes of a warp (lane owns inputs 4*lane .. 4*lane+3) read consecutive float4s: slot(k) = (k & ~127) |
  // ((k & 3) << 5) | ((k & 127) >> 2)
  extern __shared__ float4 sg[];
  const dmt_dense gate[DMT_MAX_TASKS] = {g0, g1, g2, g3};
  const int K128 = (K + 127) & ~127;
  for (int t = 0; t < n_tasks; ++t)
    for (int i = threadIdx.x; i < K; i += blockDim.x)
      sg[t * K128 + ((i & ~127) | ((i & 3) << 5) | ((i & 127) >> 2))] = __ldg(reinterpret_cast<const float4*>(gate[t].w) + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = blockIdx.x * kCastWarps + warp; b < B; b += gridDim.x * kCastWarps) {
    const float* __restrict__ xr = x + (int64_t)b * x_ld;
    float acc[DMT_MAX_TASKS][4];
#pragma unroll
    for (int t = 0; t < DMT_MAX_TASKS; ++t)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[t][e] = 0.f;
#pragma unroll 3
    for (int k0 = lane * 4; k0 < ldxb; k0 += 128) {
      float xv[4] = {0.f, 0.f, 0.f, 0.f};
      if (k0 + 4 <= K) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xr + k0));
        xv[0] = v.x; xv[1] = v.y; xv[2] = v.z; xv[3] = v.w;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (k0 + u < K) xv[u] = __ldg(xr + k0 + u);
      }
      uint2 pk;
      __nv_bfloat162 p01 = __floats2bfloat162_rn(xv[0], xv[1]), p23 = __floats2bfloat162_rn(xv[2], xv[3]);
      pk.x = *reinterpret_cast<uint32_t*>(&p01);
      pk.y = *reinterpret_cast<uint32_t*>(&p23);
      *reinterpret_cast<uint2*>(xb + (int64_t)b * ldxb + k0) = pk;
#pragma unroll
      for (int t = 0; t < DMT_MAX_TASKS; ++t)
        if (t < n_tasks) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (k0 + u < K) {
              const float4 w = sg[t * K128 + (k0 & ~127) + (u << 5) + lane];
              acc[t][0] = fmaf(xv[u], w.x, acc[t][0]);
              acc[t][1] = fmaf(xv[u], w.y, acc[t][1]);
              acc[t][2] = fmaf(xv[u], w.z, acc[t][2]);
              acc[t][3] = fmaf(xv[u], w.w, acc[t][3]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < DMT_MAX_TASKS; ++t)
      if (t < n_tasks) {
        float mx = -INFINITY;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[t][e] = warp_sum(acc[t][e]) + __ldg(gate[t].b + e);
          mx = fmaxf(mx, acc[t][e]);
        }
        float den = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[t][e] = expf(acc[t][e] - mx);
          den += acc[t][e];
        }
        if (lane == 0)
          *reinterpret_cast<float4*>(gates + ((int64_t)t * B + b) * 4) =
              make_float4(acc[t][0] / den, acc[t][1] / den, acc[t][2] / den, acc[t][3] / den);
      }
  }
}

// `gate_rows` (layer 0 only) = n_tasks * E extra rows after the experts': row t * E + e = gate t's kernel column e
__global__ void mmoe_prepare_kernel(dmt_mmoe_weights w, int layer, int E, int K, int N, int ldk, int gate_rows,
                                    __nv_bfloat16* __restrict__ out) {
  const int64_t total = ((int64_t)E * N + gate_rows) * ldk;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = i % ldk;
    const int64_t row = i / ldk;
    float v = 0.f;
    if (k < K) {
      if (row < (int64_t)E * N) {
        const int n = row % N, z = row / N;
        v = w.expert[z][layer].w[(int64_t)k * N + n];
      } else {
        const int j = (int)(row - (int64_t)E * N), t = j / E, e = j % E;
        v = w.gate[t].w[(int64_t)k * E + e];
      }
    }
    out[i] = __float2bfloat16(v);
  }
}
__global__ void mmoe_prepare_bias_kernel(dmt_mmoe_weights w, int layer, int E, int N, int gate_rows,
                                         float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E * N) out[i] = w.expert[i / N][layer].b[i % N];
  else if (i < E * N + gate_rows) out[i] = w.gate[(i - E * N) / E].b[(i - E * N) % E];
}

// =====================================================================================================================
// Expert layers 2 + 3 + the first tower layer in ONE kernel (hidden_units_bottom = x, 256, 128; one tower layer):
// a CTA owns 128 samples of ONE expert.  h2 = relu(h1 W2 + b2) never leaves the SM: the layer-2 accumulator
// (256 TMEM columns) is packed to bf16 IN PLACE and is the tensor-memory A operand of layer 3; h3 likewise feeds the
// tower contraction u_e = h3_e [Wt_0 | Wt_1 ...] (linear, so it commutes with the gate mixture:
// z_t Wt = sum_e g_te (h3_e Wt), mmoe_transformer_unbias.py:99-126).  Only u [E][B][64] fp32 goes back to HBM; the
// mixture + bias + ReLU + output unit are mmoe_mix_kernel.  Replaces two GEMM launches + the head kernel
// (15 + 13 + 26 us at B = 4096) and the round trip of h2 / h3 through HBM.
// Warp roles as gemm_tc_kernel: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (thread = sample row).
// =====================================================================================================================
constexpr int kT2N2 = 256, kT2N3 = 128, kT2NT = 64, kT2Stages = 3;
constexpr int kT2StageBytes = (GBM * GBK + kT2N2 * GBK) * 2;        // 48 KB: A 128 x 64 | B 256 x 64
constexpr int kT2W3Bytes = kT2N3 * kT2N2 * 2;                        // 64 KB: 4 k-blocks of 128 rows x 64
constexpr int kT2WtBytes = kT2NT * kT2N3 * 2;                        // 16 KB: 2 k-blocks of 64 rows x 64
constexpr int kT2Smem = kT2Stages * kT2StageBytes + kT2W3Bytes + kT2WtBytes + 1024;

struct MmoeTailArgs {
  CUtensorMap tmA;             // h1  [E * B, K1]     box {64, 128}
  CUtensorMap tmB2;            // W2t [E * 256, K1]   box {64, 128}
  CUtensorMap tmB3;            // W3t [E * 128, 256]  box {64, 128}
  CUtensorMap tmBt;            // Wt  [64, 128]       box {64, 64}    rows t * U + j, zero rows beyond n_tasks * U
  const float* b2;             // [E][256]
  const float* b3;             // [E][128]
  float* u;                    // [E][B][64]
  int32_t B, K1;
};

__global__ void __launch_bounds__(kGemmThreads) mmoe_tail_kernel(const __grid_constant__ MmoeTailArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kT2Stages], empty_bar[kT2Stages], wbar, acc2_bar, a3_bar, acc3_bar, a4_bar, acc4_bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW3 = smem + kT2Stages * kT2StageBytes;
  uint8_t* sWt = sW3 + kT2W3Bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * GBM, z = blockIdx.y;
  const int nkb = (g.K1 + GBK - 1) / GBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kT2Stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&wbar, 1);
    mbar_init(&acc2_bar, 1);
    mbar_init(&acc3_bar, 1);
    mbar_init(&acc4_bar, 1);
    mbar_init(&a3_bar, 128);
    mbar_init(&a4_bar, 128);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmB2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmB3) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmBt) : "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  constexpr int tH2 = 0, tH3 = 256, tU = 384;          // accumulators; the packed A operands reuse [0, 128) / [0, 64)
  griddep_launch();                                    // (the mixture kernel behind this one)

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: layer-3 / tower kernels once, then the layer-2 operand ring =====
      mbar_expect_tx(&wbar, kT2W3Bytes + kT2WtBytes);
      for (int kb = 0; kb < kT2N2 / GBK; ++kb) tma_load_2d(sW3 + kb * (kT2N3 * GBK * 2), &g.tmB3, kb * GBK, z * kT2N3, &wbar);
      for (int kb = 0; kb < kT2N3 / GBK; ++kb) tma_load_2d(sWt + kb * (kT2NT * GBK * 2), &g.tmBt, kb * GBK, 0, &wbar);
      griddep_wait();    // dependent launch: h1 is the layer-0 GEMM's output (the images above are not)
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kT2Stages;
        mbar_wait(&empty_bar[s], ((kb / kT2Stages) & 1) ^ 1);
        uint8_t* sa = smem + s * kT2StageBytes;
        uint8_t* sb = sa + GBM * GBK * 2;
        mbar_expect_tx(&full_bar[s], kT2StageBytes);
        tma_load_2d(sa, &g.tmA, kb * GBK, z * g.B + m0, &full_bar[s]);
        tma_load_2d(sb, &g.tmB2, kb * GBK, z * kT2N2, &full_bar[s]);
        tma_load_2d(sb + GBM * GBK * 2, &g.tmB2, kb * GBK, z * kT2N2 + GBM, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      {
        constexpr uint32_t idesc = make_idesc_bf16(GBM, kT2N2);
        for (int kb = 0; kb < nkb; ++kb) {
          const int s = kb % kT2Stages;
          mbar_wait(&full_bar[s], (kb / kT2Stages) & 1);
          fence_after_sync();
          const uint32_t sa = smem_u32(smem + s * kT2StageBytes), sb = sa + GBM * GBK * 2;
#pragma unroll
          for (int k = 0; k < GBK / 16; ++k)
            mma_bf16_ss(tbase + tH2, make_smem_desc(sa + k * 32, 16, 1024, kLayoutSW128),
                        make_smem_desc(sb + k * 32, 16, 1024, kLayoutSW128), idesc, (kb | k) != 0);
          commit(&empty_bar[s]);
        }
        commit(&acc2_bar);
      }
      {  // layer 3: A = relu(h2) packed in tensor memory columns [0, 128)
        mbar_wait(&wbar, 0);
        mbar_wait(&a3_bar, 0);
        fence_after_sync();
        constexpr uint32_t idesc = make_idesc_bf16(GBM, kT2N3);
        const uint32_t sw = smem_u32(sW3);
#pragma unroll
        for (int ks = 0; ks < kT2N2 / 16; ++ks)
          mma_bf16_ts(tbase + tH3, tbase + ks * 8,
                      make_smem_desc(sw + (ks / 4) * (kT2N3 * GBK * 2) + (ks % 4) * 32, 16, 1024, kLayoutSW128), idesc, ks > 0);
        commit(&acc3_bar);
      }
      {  // tower contraction: A = relu(h3) packed in columns [0, 64)
        mbar_wait(&a4_bar, 0);
        fence_after_sync();
        constexpr uint32_t idesc = make_idesc_bf16(GBM, kT2NT);
        const uint32_t sw = smem_u32(sWt);
#pragma unroll
        for (int ks = 0; ks < kT2N3 / 16; ++ks)
          mma_bf16_ts(tbase + tU, tbase + ks * 8,
                      make_smem_desc(sw + (ks / 4) * (kT2NT * GBK * 2) + (ks % 4) * 32, 16, 1024, kLayoutSW128), idesc, ks > 0);
        commit(&acc4_bar);
      }
    }
  } else {
    // ===== epilogue warps: thread = sample row = TMEM lane =====
    const int row = (warp & 3) * 32 + lane;
    const int m = m0 + row;
    auto relu_pack = [&](int acc_col, int ncols, const float* __restrict__ bias) {
      // fp32 accumulator columns [acc_col, acc_col + ncols) -> relu(+bias) -> bf16 pairs, packed into columns
      // [0, ncols / 2): the packed words trail the columns already read (same thread = same lane)
#pragma unroll 1
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_addr(tbase, acc_col + c0), r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0) + q);
          pk[q * 2] = pack_bf16x2(fmaxf(__uint_as_float(r[q * 4]) + bb.x, 0.f), fmaxf(__uint_as_float(r[q * 4 + 1]) + bb.y, 0.f));
          pk[q * 2 + 1] = pack_bf16x2(fmaxf(__uint_as_float(r[q * 4 + 2]) + bb.z, 0.f), fmaxf(__uint_as_float(r[q * 4 + 3]) + bb.w, 0.f));
        }
        tmem_st16(tmem_addr(tbase, c0 / 2), pk);
      }
      tmem_st_wait();
      fence_before_sync();
    };
    mbar_wait(&acc2_bar, 0);
    fence_after_sync();
    relu_pack(tH2, kT2N2, g.b2 + z * kT2N2);
    mbar_arrive(&a3_bar);
    mbar_wait(&acc3_bar, 0);
    fence_after_sync();
    relu_pack(tH3, kT2N3, g.b3 + z * kT2N3);
    mbar_arrive(&a4_bar);
    mbar_wait(&acc4_bar, 0);
    fence_after_sync();
    float* urow = g.u + ((int64_t)z * g.B + m) * kT2NT;
#pragma unroll
    for (int c0 = 0; c0 < kT2NT; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_addr(tbase, tU + c0), r);
      tmem_ld_wait();
      if (m < g.B) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(urow + c0 + q * 4) =
              make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]), __uint_as_float(r[q * 4 + 2]),
                          __uint_as_float(r[q * 4 + 3]));
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, 512);
}

// logits[t][b] = relu(sum_e g[t][b][e] u[e][b][t U + :] + b_t) . w_out_t + b_out_t   (mmoe_transformer_unbias.py:99-126)
struct MmoeMixArgs {
  const float* u;              // [E][B][64]
  const float* gates;          // [T][B][E]
  dmt_dense tower[DMT_MAX_TASKS], tower_out[DMT_MAX_TASKS];
  float* logits;               // [T][B]
  int32_t B, E, T, U;
};
__global__ void __launch_bounds__(256) mmoe_mix_kernel(const __grid_constant__ MmoeMixArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  griddep_wait();                                      // dependent launch: u comes from the fused tail kernel (read
                                                       // with coherent loads, not through the read-only path)
  if (b >= a.B) return;
  for (int t = 0; t < a.T; ++t) {
    float part = 0.f;
    for (int j = lane; j < a.U; j += 32) {
      float acc = 0.f;
      for (int e = 0; e < a.E; ++e)
        acc = fmaf(__ldcg(a.gates + ((int64_t)t * a.B + b) * a.E + e), __ldcg(a.u + ((int64_t)e * a.B + b) * kT2NT + t * a.U + j), acc);
      part = fmaf(fmaxf(acc + __ldg(a.tower[t].b + j), 0.f), __ldg(a.tower_out[t].w + j), part);
    }
    part = warp_sum(part);
    if (lane == 0) a.logits[(int64_t)t * a.B + b] = part + __ldg(a.tower_out[t].b);
  }
}

// tower image for the tail kernel: Wt[n = t U + j][k] = tower_t kernel [k][j]  (bf16, zero rows up to 64)
__global__ void mmoe_prepare_tower_kernel(dmt_mmoe_weights w, int T, int U, int Hd, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kT2NT * Hd) return;
  const int k = i % Hd, n = i / Hd, t = n / U, j = n % U;
  out[i] = __float2bfloat16(t < T ? w.tower[t][0].w[(int64_t)k * U + j] : 0.f);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
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

// 2-D bf16 row-major [rows, cols] (row stride ld elements), box 64 x 128, 128-byte swizzle, OOB -> 0.
static int make_map(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows = GBM) {
  EncodeTiledFn fn = encode_fn();
  DMT_REQUIRE(fn, DMT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {GBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DMT_REQUIRE(r == CUDA_SUCCESS, DMT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld",
              (int)r, (long long)rows, (long long)cols, (long long)ld);
  return DMT_OK;
}

static inline int pad8(int k) { return (k + 7) & ~7; }

// the gate kernels ride along as extra rows of layer 0's B operand (dmt_mmoe_fwd_bf16in reads them; the fp32-input
// entry computes its gates in fp32 while it converts x)
static inline int gate_rows(const dmt_mmoe_cfg* cfg, int layer) {
  static_assert(DMT_MAX_TASKS * DMT_MAX_EXPERTS <= 32, "the gate logits must fit one 32-column accumulator chunk");
  return layer == 0 ? cfg->n_tasks * cfg->n_experts : 0;
}

// prepared = [Bt_0 (+ gate rows) | Bt_1 | ... ] bf16 then [bias_0 (+ gate biases) | bias_1 | ...] fp32 (256-byte aligned
// sections)
static size_t prepared_layout(const dmt_mmoe_cfg* cfg, size_t* w_off, size_t* b_off) {
  size_t off = 0;
  int K = cfg->in_dim;
  for (int l = 0; l < cfg->n_layers; ++l) {
    if (w_off) w_off[l] = off;
    off += (((size_t)cfg->n_experts * cfg->units[l] + gate_rows(cfg, l)) * pad8(K) * 2 + 255) & ~(size_t)255;
    K = cfg->units[l];
  }
  for (int l = 0; l < cfg->n_layers; ++l) {
    if (b_off) b_off[l] = off;
    off += (((size_t)cfg->n_experts * cfg->units[l] + gate_rows(cfg, l)) * 4 + 255) & ~(size_t)255;
  }
  return off;
}

// hidden_units_bottom = (x, 256, 128), one tower layer of <= 64 / n_tasks units: layers 2 + 3 + tower fused
static bool tail_fused(const dmt_mmoe_cfg* cfg) {
  return cfg->n_layers == 3 && cfg->units[1] == kT2N2 && cfg->units[2] == kT2N3 && cfg->n_tower_layers == 1 &&
         cfg->tower_units[0] > 0 && cfg->n_tasks * cfg->tower_units[0] <= kT2NT;
}

// (+ the bf16 tower image of the fused tail kernel after the sections of prepared_layout)
size_t mmoe_tc_prepared_bytes(const dmt_mmoe_cfg* cfg) {
  return prepared_layout(cfg, nullptr, nullptr) + (tail_fused(cfg) ? kT2WtBytes : 0) + 256;
}

// activations: xb | h_0 | h_1 | ... (bf16) | gates (fp32)
static size_t workspace_layout(const dmt_mmoe_cfg* cfg, size_t* h_off, size_t* gate_off) {
  size_t off = ((size_t)cfg->batch * pad8(cfg->in_dim) * 2 + 255) & ~(size_t)255;
  for (int l = 0; l < cfg->n_layers; ++l) {
    if (h_off) h_off[l] = off;
    off += ((size_t)cfg->n_experts * cfg->batch * cfg->units[l] * 2 + 255) & ~(size_t)255;
  }
  if (gate_off) *gate_off = off;
  off += ((size_t)cfg->n_tasks * cfg->batch * cfg->n_experts * 4 + 255) & ~(size_t)255;
  return off;
}

size_t mmoe_tc_workspace_bytes(const dmt_mmoe_cfg* cfg) { return workspace_layout(cfg, nullptr, nullptr) + 256; }

bool mmoe_tc_supported(const dmt_mmoe_cfg* cfg, const char** why) {
  *why = nullptr;
  for (int l = 0; l < cfg->n_layers; ++l)
    if (cfg->units[l] % 8) *why = "bf16 MMoE path needs hidden_units_bottom that are multiples of 8";
  return *why == nullptr;
}

int mmoe_tc_prepare(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, void* prepared, cudaStream_t st) {
  size_t w_off[DMT_MAX_LAYERS], b_off[DMT_MAX_LAYERS];
  prepared_layout(cfg, w_off, b_off);
  uint8_t* base = (uint8_t*)prepared;
  int K = cfg->in_dim;
  for (int l = 0; l < cfg->n_layers; ++l) {
    const int N = cfg->units[l], E = cfg->n_experts, ldk = pad8(K), gr = gate_rows(cfg, l);
    const int64_t total = ((int64_t)E * N + gr) * ldk;
    mmoe_prepare_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(*w, l, E, K, N, ldk, gr,
                                                                         (__nv_bfloat16*)(base + w_off[l]));
    mmoe_prepare_bias_kernel<<<(E * N + gr + 255) / 256, 256, 0, st>>>(*w, l, E, N, gr, (float*)(base + b_off[l]));
    K = N;
  }
  if (tail_fused(cfg)) {
    const size_t wt_off = prepared_layout(cfg, nullptr, nullptr);
    mmoe_prepare_tower_kernel<<<(kT2NT * kT2N3 + 255) / 256, 256, 0, st>>>(*w, cfg->n_tasks, cfg->tower_units[0], kT2N3,
                                                                          (__nv_bfloat16*)(base + wt_off));
  }
  DMT_CUDA_LAUNCH_CHECK("mmoe_prepare_kernel");
  return DMT_OK;
}

int mmoe_head_launch(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const float* x, int64_t x_ld,
                     const void* h_last, int h_is_bf16, const float* gates, float* logits, cudaStream_t st);

// x: fp32 input (converted to bf16 into the workspace here) -- or xb_in: the input is bf16 already
int mmoe_tc_launch(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const float* x, int64_t x_ld,
                   const void* xb_in, int64_t xb_ld, float* logits, void* workspace, const void* prepared,
                   cudaStream_t st) {
  size_t w_off[DMT_MAX_LAYERS], b_off[DMT_MAX_LAYERS], h_off[DMT_MAX_LAYERS], gate_off;
  prepared_layout(cfg, w_off, b_off);
  workspace_layout(cfg, h_off, &gate_off);
  uint8_t* ws = (uint8_t*)workspace;
  const uint8_t* pw = (const uint8_t*)prepared;
  const int B = cfg->batch, E = cfg->n_experts;
  __nv_bfloat16* xb = (__nv_bfloat16*)ws;
  float* gates = (float*)(ws + gate_off);
  const int ldx = pad8(cfg->in_dim);
  const size_t gsm = (size_t)cfg->n_tasks * ((cfg->in_dim + 127) & ~127) * sizeof(float4);
  bool gate_al = true;
  for (int t = 0; t < cfg->n_tasks; ++t) gate_al = gate_al && ((uintptr_t)w->gate[t].w & 15) == 0;
  const bool fold_gates = xb_in != nullptr;                  // gates computed by the layer-0 GEMM itself
  if (fold_gates) {
    // (nothing to launch: slab z = E of the layer-0 GEMM below)
  } else if (E == 4 && gate_al && x_ld % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)gates & 15) == 0 && gsm <= 160 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mmoe_cast_gate4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mmoe_cast_gate4_kernel)");
    int grid = (B + kCastWarps - 1) / kCastWarps;
    const int cap = 2 * sm_count_cached();
    if (grid > cap) grid = cap;
    mmoe_cast_gate4_kernel<<<grid, kCastWarps * 32, gsm, st>>>(x, x_ld, B, cfg->in_dim, xb, ldx, w->gate[0], w->gate[1],
                                                                w->gate[2], w->gate[3], cfg->n_tasks, gates);
    DMT_CUDA_LAUNCH_CHECK("mmoe_cast_gate4_kernel");
  } else {
    mmoe_cast_gate_kernel<<<(B + 7) / 8, 256, 0, st>>>(x, x_ld, B, cfg->in_dim, xb, ldx, w->gate[0], w->gate[1],
                                                       w->gate[2], w->gate[3], cfg->n_tasks, E, gates);
    DMT_CUDA_LAUNCH_CHECK("mmoe_cast_gate_kernel");
  }

  const int smem_bytes = GSTAGES * kStageBytes + 1024;
  {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_tc_kernel)");
  }
  const __nv_bfloat16* a = xb_in ? (const __nv_bfloat16*)xb_in : xb;
  int64_t a_rows = B;          // rows in the A tensor map
  int a_rows_per_z = 0, K = cfg->in_dim;
  int64_t lda = xb_in ? xb_ld : ldx;
  static const bool fuse_env = !(getenv("DMT_MMOE_FUSED_TAIL") && atoi(getenv("DMT_MMOE_FUSED_TAIL")) == 0);
  const bool fused = tail_fused(cfg) && fuse_env;
  for (int l = 0; l < (fused ? 1 : cfg->n_layers); ++l) {
    const int N = cfg->units[l];
    GemmTcArgs g;
    int rc = make_map(&g.tmA, a, a_rows, K, lda);
    if (rc != DMT_OK) return rc;
    rc = make_map(&g.tmB, pw + w_off[l], (int64_t)E * N + gate_rows(cfg, l), K, pad8(K));
    if (rc != DMT_OK) return rc;
    g.bias = (const float*)(pw + b_off[l]);
    g.C = (__nv_bfloat16*)(ws + h_off[l]);
    g.M = B; g.N = N; g.K = K;
    g.a_rows_per_z = a_rows_per_z;
    g.b_rows_per_z = N;
    g.ldc = N;
    g.c_stride_z = (int64_t)B * N;
    g.relu = 1;
    const bool gz = fold_gates && l == 0;
    g.gate_z = gz ? E : -1;
    g.n_tasks = cfg->n_tasks;
    g.n_experts = E;
    g.gates = gates;
    dim3 grid((N + GBN - 1) / GBN, (B + GBM - 1) / GBM, gz ? E + 1 : E);
    gemm_tc_kernel<<<grid, kGemmThreads, smem_bytes, st>>>(g);
    DMT_CUDA_LAUNCH_CHECK("gemm_tc_kernel");
    a = g.C;
    a_rows = (int64_t)E * B;
    a_rows_per_z = B;
    K = N;
    lda = N;
  }
  if (fused) {
    const int K1 = cfg->units[0];
    MmoeTailArgs t;
    int rc = make_map(&t.tmA, ws + h_off[0], (int64_t)E * B, K1, K1);
    if (rc != DMT_OK) return rc;
    rc = make_map(&t.tmB2, pw + w_off[1], (int64_t)E * kT2N2, K1, pad8(K1));
    if (rc != DMT_OK) return rc;
    rc = make_map(&t.tmB3, pw + w_off[2], (int64_t)E * kT2N3, kT2N2, kT2N2);
    if (rc != DMT_OK) return rc;
    rc = make_map(&t.tmBt, pw + prepared_layout(cfg, nullptr, nullptr), kT2NT, kT2N3, kT2N3, kT2NT);
    if (rc != DMT_OK) return rc;
    t.b2 = (const float*)(pw + b_off[1]);
    t.b3 = (const float*)(pw + b_off[2]);
    t.u = (float*)(ws + h_off[1]);                       // h2 never exists in HBM: its slot holds u [E][B][64] fp32
    t.B = B;
    t.K1 = K1;
    cudaError_t e = cudaFuncSetAttribute(mmoe_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kT2Smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mmoe_tail_kernel)");
    // programmatic dependent launches: each kernel's prologue (TMEM allocation, barrier setup, the W3 / tower images)
    // runs while the kernel before it drains
    // (336.7 -> 324.3 us per step)
    e = launch_pdl(mmoe_tail_kernel, dim3((B + GBM - 1) / GBM, E), dim3(kGemmThreads), kT2Smem, st, t);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(mmoe_tail_kernel)");
    DMT_CUDA_LAUNCH_CHECK("mmoe_tail_kernel");
    MmoeMixArgs mx;
    mx.u = t.u;
    mx.gates = gates;
    for (int tt = 0; tt < DMT_MAX_TASKS; ++tt) {
      mx.tower[tt] = w->tower[tt][0];
      mx.tower_out[tt] = w->tower_out[tt];
    }
    mx.logits = logits;
    mx.B = B; mx.E = E; mx.T = cfg->n_tasks; mx.U = cfg->tower_units[0];
    e = launch_pdl(mmoe_mix_kernel, dim3((B + 7) / 8), dim3(256), 0, st, mx);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(mmoe_mix_kernel)");
    DMT_CUDA_LAUNCH_CHECK("mmoe_mix_kernel");
    return DMT_OK;
  }
  return mmoe_head_launch(cfg, w, x, x_ld, ws + h_off[cfg->n_layers - 1], 1, gates, logits, st);
}

}  // namespace dmt
