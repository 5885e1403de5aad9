// dmt_mmoe_fwd, DMT_PRECISION_F32: expert MLPs as batched fp32 SIMT GEMMs with a fused
// bias+ReLU epilogue, then one warp per sample for gates -> mixture -> task towers.
#include <cuda_bf16.h>

#include "dmt_common.cuh"

namespace dmt {

struct GemmBatch {
  const float* A[DMT_MAX_EXPERTS];
  const float* W[DMT_MAX_EXPERTS];
  const float* bias[DMT_MAX_EXPERTS];
  float* C[DMT_MAX_EXPERTS];
  int64_t lda, ldc;
  int M, N, K;
  int relu;
};

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;

// C[z] = act(A[z] W[z] + bias[z]); A row-major [M, K] (lda), W row-major [K, N] (TF layout).
__global__ void __launch_bounds__(256) gemm_bias_act_f32_kernel(const __grid_constant__ GemmBatch g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN];
  const int z = blockIdx.z;
  const float* __restrict__ A = g.A[z];
  const float* __restrict__ W = g.W[z];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tm0 = (tid >> 4) * TM, tn0 = (tid & 15) * TN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / 256; ++i) {
      const int idx = tid + i * 256;
      const int r = idx / BK, kk = idx - r * BK;
      const int m = m0 + r, k = k0 + kk;
      As[kk][r] = (m < g.M && k < g.K) ? __ldg(A + (int64_t)m * g.lda + k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < (BK * BN) / 256; ++i) {
      const int idx = tid + i * 256;
      const int kk = idx / BN, c = idx - kk * BN;
      const int k = k0 + kk, n = n0 + c;
      Ws[kk][c] = (k < g.K && n < g.N) ? __ldg(W + (int64_t)k * g.N + n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][tm0]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][tm0 + 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[kk][tn0]);
      const float av[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wv[TN] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float* __restrict__ bias = g.bias[z];
  float* __restrict__ C = g.C[z];
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + tm0 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tn0 + j;
      if (n >= g.N) continue;
      float y = acc[i][j] + __ldg(bias + n);
      if (g.relu) y = fmaxf(y, 0.f);
      C[(int64_t)m * g.ldc + n] = y;
    }
  }
}

struct HeadArgs {
  dmt_mmoe_cfg cfg;
  dmt_dense gate[DMT_MAX_TASKS];
  dmt_dense tower[DMT_MAX_TASKS][DMT_MAX_LAYERS];
  dmt_dense tower_out[DMT_MAX_TASKS];
  const float* x;
  int64_t x_ld;
  const void* h_last;    // [E][B][H] fp32, or bf16 when h_is_bf16
  const float* gates;    // optional precomputed softmax gates [T][B][E] (NULL: computed here from x)
  int32_t h_is_bf16;
  float* logits;         // [T][B]
  int32_t hdim;          // units of the last expert layer
  int32_t vec_floats;    // per-warp scratch floats (2 buffers)
};

constexpr int kHeadWarps = 8;

// One warp per sample: gate softmax (mmoe_transformer_unbias.py:85-94), expert mixture
// (:99-104), tower MLP (:107-126).
__global__ void __launch_bounds__(kHeadWarps * 32) mmoe_head_kernel(const __grid_constant__ HeadArgs a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kHeadWarps + warp;
  if (b >= a.cfg.batch) return;
  float* y0 = sm + warp * a.vec_floats;
  float* y1 = y0 + a.vec_floats / 2;
  const int E = a.cfg.n_experts, K = a.cfg.in_dim, Hd = a.hdim, B = a.cfg.batch;
  const float* __restrict__ xr = a.x + (int64_t)b * a.x_ld;
  for (int t = 0; t < a.cfg.n_tasks; ++t) {
    float gl[DMT_MAX_EXPERTS];
#pragma unroll
    for (int e = 0; e < DMT_MAX_EXPERTS; ++e) gl[e] = 0.f;
    float inv = 1.0f;
    if (a.gates) {
#pragma unroll
      for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
        if (e < E) gl[e] = __ldg(a.gates + ((int64_t)t * B + b) * E + e);
    } else {
      const float* __restrict__ Wg = a.gate[t].w;
      for (int k = lane; k < K; k += 32) {
        const float xv = __ldg(xr + k);
#pragma unroll
        for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
          if (e < E) gl[e] = fmaf(xv, __ldg(Wg + (int64_t)k * E + e), gl[e]);
      }
      float mx = -INFINITY;
#pragma unroll
      for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
        if (e < E) {
          gl[e] = warp_sum(gl[e]) + __ldg(a.gate[t].b + e);
          mx = fmaxf(mx, gl[e]);
        }
      float den = 0.f;
#pragma unroll
      for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
        if (e < E) {
          gl[e] = expf(gl[e] - mx);
          den += gl[e];
        }
      inv = 1.0f / den;
    }
    for (int c = lane; c < Hd; c += 32) {
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
        if (e < E) {
          const int64_t idx = ((int64_t)e * B + b) * Hd + c;
          const float hv = a.h_is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.h_last)[idx])
                                       : __ldg(reinterpret_cast<const float*>(a.h_last) + idx);
          acc = fmaf(gl[e] * inv, hv, acc);
        }
      y0[c] = acc;
    }
    __syncwarp();
    int in_dim = Hd;
    float* cur = y0;
    float* nxt = y1;
    for (int l = 0; l < a.cfg.n_tower_layers; ++l) {
      const int units = a.cfg.tower_units[l];
      const float* __restrict__ Wt = a.tower[t][l].w;
      for (int n = lane; n < units; n += 32) {
        float acc = 0.f;
        for (int k = 0; k < in_dim; ++k) acc = fmaf(cur[k], __ldg(Wt + (int64_t)k * units + n), acc);
        nxt[n] = fmaxf(acc + __ldg(a.tower[t][l].b + n), 0.f);
      }
      __syncwarp();
      float* tmp = cur; cur = nxt; nxt = tmp;
      in_dim = units;
    }
    float acc = 0.f;
    for (int k = lane; k < in_dim; k += 32) acc = fmaf(cur[k], __ldg(a.tower_out[t].w + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) a.logits[(int64_t)t * B + b] = acc + __ldg(a.tower_out[t].b);
    __syncwarp();
  }
}


// Same computation with precomputed gates (bf16 path) for ONE tower layer (dmt.conf: 128 -> 32 -> 1), organised as a
// small batched GEMM per CTA: 32 samples x all tasks.  The tower kernels of every task are staged in shared memory
// once, the gate mixtures z[s][t][:] are built with coalesced 8-byte loads of the bf16 expert outputs, then
// thread (s, t, q) computes units/4 tower outputs from shared memory and the 4 threads of (s, t) reduce
// relu(out + b) . w_out by shuffles.
constexpr int kHeadTileS = 32;
__global__ void __launch_bounds__(256) mmoe_head_tile_kernel(const __grid_constant__ HeadArgs a) {
  extern __shared__ float sm[];
  const int E = a.cfg.n_experts, Hd = a.hdim, B = a.cfg.batch, T = a.cfg.n_tasks, U = a.cfg.tower_units[0];
  float* sW = sm;                                   // [T][Hd][U]
  float* sB = sW + T * Hd * U;                      // [T][U] tower bias | [T][U] output kernel
  float* sZ = sB + 2 * T * U;                       // [kHeadTileS][T][Hd + 4]
  const int zld = Hd + 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int t = 0; t < T; ++t) {
    for (int i = tid; i < Hd * U; i += 256) sW[t * Hd * U + i] = __ldg(a.tower[t][0].w + i);
    for (int i = tid; i < U; i += 256) {
      sB[t * U + i] = __ldg(a.tower[t][0].b + i);
      sB[(T + t) * U + i] = __ldg(a.tower_out[t].w + i);
    }
  }
  const int b0 = blockIdx.x * kHeadTileS;
  // mixtures: warp w handles samples w, w + 8, ...; a lane covers 4 consecutive columns per trip
  for (int s = warp; s < kHeadTileS; s += 8) {
    const int b = min(b0 + s, B - 1);
    float g[DMT_MAX_TASKS][DMT_MAX_EXPERTS];
#pragma unroll
    for (int t = 0; t < DMT_MAX_TASKS; ++t)
#pragma unroll
      for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
        g[t][e] = (t < T && e < E) ? __ldg(a.gates + ((int64_t)t * B + b) * E + e) : 0.f;
    for (int c = lane * 4; c < Hd; c += 128) {
      float z[DMT_MAX_TASKS][4];
#pragma unroll
      for (int t = 0; t < DMT_MAX_TASKS; ++t) z[t][0] = z[t][1] = z[t][2] = z[t][3] = 0.f;
#pragma unroll
      for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
        if (e < E) {
          const int64_t idx = ((int64_t)e * B + b) * Hd + c;
          float h[4];
          if (a.h_is_bf16) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(a.h_last) + idx));
            const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
            const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
            h[0] = lo.x; h[1] = lo.y; h[2] = hi.x; h[3] = hi.y;
          } else {
            const float4 v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.h_last) + idx));
            h[0] = v.x; h[1] = v.y; h[2] = v.z; h[3] = v.w;
          }
#pragma unroll
          for (int t = 0; t < DMT_MAX_TASKS; ++t)
            if (t < T) {
#pragma unroll
              for (int j = 0; j < 4; ++j) z[t][j] = fmaf(g[t][e], h[j], z[t][j]);
            }
        }
#pragma unroll
      for (int t = 0; t < DMT_MAX_TASKS; ++t)
        if (t < T) *reinterpret_cast<float4*>(sZ + (s * T + t) * zld + c) = make_float4(z[t][0], z[t][1], z[t][2], z[t][3]);
    }
  }
  __syncthreads();
  // tower: item = (s, t, q); q-th quarter of the units.  4 consecutive lanes share (s, t).
  const int UQ = U / 4;                              // units per thread (<= 16, checked by the launcher)
  for (int item = tid; item < kHeadTileS * T * 4; item += 256) {
    const int q = item & 3, st = item >> 2, t = st % T, s = st / T;
    const float* zr = sZ + (s * T + t) * zld;
    const float* W = sW + t * Hd * U + q * UQ;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int k = 0; k < Hd; k += 4) {
      const float4 zv = *reinterpret_cast<const float4*>(zr + k);
      const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float* wr = W + (k + kk) * U;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          if (j < UQ) {
            const float4 w = *reinterpret_cast<const float4*>(wr + j);
            acc[j] = fmaf(zz[kk], w.x, acc[j]);
            acc[j + 1] = fmaf(zz[kk], w.y, acc[j + 1]);
            acc[j + 2] = fmaf(zz[kk], w.z, acc[j + 2]);
            acc[j + 3] = fmaf(zz[kk], w.w, acc[j + 3]);
          }
      }
    }
    float part = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < UQ) part = fmaf(fmaxf(acc[j] + sB[t * U + q * UQ + j], 0.f), sB[(T + t) * U + q * UQ + j], part);
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if (q == 0 && b0 + s < B) a.logits[(int64_t)t * B + b0 + s] = part + __ldg(a.tower_out[t].b);
  }
}

int mmoe_head_launch(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const float* x, int64_t x_ld,
                     const void* h_last, int h_is_bf16, const float* gates, float* logits, cudaStream_t st) {
  const int B = cfg->batch;
  const int in_dim = cfg->units[cfg->n_layers - 1];
  HeadArgs h;
  h.cfg = *cfg;
  for (int t = 0; t < cfg->n_tasks; ++t) {
    h.gate[t] = w->gate[t];
    for (int l = 0; l < cfg->n_tower_layers; ++l) h.tower[t][l] = w->tower[t][l];
    h.tower_out[t] = w->tower_out[t];
  }
  h.x = x;
  h.x_ld = x_ld;
  h.h_last = h_last;
  h.gates = gates;
  h.h_is_bf16 = h_is_bf16;
  h.logits = logits;
  h.hdim = in_dim;
  int mx = in_dim;
  for (int l = 0; l < cfg->n_tower_layers; ++l) mx = cfg->tower_units[l] > mx ? cfg->tower_units[l] : mx;
  h.vec_floats = 2 * ((mx + 31) / 32 * 32);
  const int U0 = cfg->tower_units[0];
  if (gates && cfg->n_tower_layers == 1 && in_dim % 4 == 0 && U0 % 16 == 0 && U0 <= 64 &&
      (kHeadTileS * cfg->n_tasks * 4) % 256 == 0 && ((uintptr_t)h_last & 15) == 0) {   // bf16 path (gates precomputed)
    const size_t fsmem = ((size_t)cfg->n_tasks * in_dim * U0 + 2 * (size_t)cfg->n_tasks * U0 +
                          (size_t)kHeadTileS * cfg->n_tasks * (in_dim + 4)) * sizeof(float);
    if (fsmem <= 200 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(mmoe_head_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mmoe_head_tile_kernel)");
      mmoe_head_tile_kernel<<<(B + kHeadTileS - 1) / kHeadTileS, 256, fsmem, st>>>(h);
      DMT_CUDA_LAUNCH_CHECK("mmoe_head_tile_kernel");
      return DMT_OK;
    }
  }
  const size_t smem = (size_t)kHeadWarps * h.vec_floats * sizeof(float);
  DMT_REQUIRE(smem <= 48 * 1024, DMT_ERR_UNSUPPORTED_SHAPE, "dmt_mmoe_fwd: tower width %d too large", mx);
  mmoe_head_kernel<<<(B + kHeadWarps - 1) / kHeadWarps, kHeadWarps * 32, smem, st>>>(h);
  DMT_CUDA_LAUNCH_CHECK("mmoe_head_kernel");
  return DMT_OK;
}

int mmoe_f32_launch(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const float* x, int64_t x_ld,
                    float* logits, float* ws, cudaStream_t st) {
  const int B = cfg->batch, E = cfg->n_experts;
  const float* in = x;
  int64_t in_ld = x_ld;
  int in_dim = cfg->in_dim;
  int64_t in_stride = 0;   // experts share x for layer 0
  float* layer_out = ws;
  for (int l = 0; l < cfg->n_layers; ++l) {
    const int units = cfg->units[l];
    GemmBatch g;
    for (int e = 0; e < E; ++e) {
      g.A[e] = in + in_stride * e;
      g.W[e] = w->expert[e][l].w;
      g.bias[e] = w->expert[e][l].b;
      g.C[e] = layer_out + (int64_t)e * B * units;
    }
    g.lda = in_ld;
    g.ldc = units;
    g.M = B;
    g.N = units;
    g.K = in_dim;
    g.relu = 1;
    dim3 grid((units + BN - 1) / BN, (B + BM - 1) / BM, E);
    gemm_bias_act_f32_kernel<<<grid, 256, 0, st>>>(g);
    DMT_CUDA_LAUNCH_CHECK("gemm_bias_act_f32_kernel");
    in = layer_out;
    in_ld = units;
    in_dim = units;
    in_stride = (int64_t)B * units;
    layer_out += (int64_t)E * B * units;
  }
  return mmoe_head_launch(cfg, w, x, x_ld, in, 0, nullptr, logits, st);
}

size_t mmoe_tc_prepared_bytes(const dmt_mmoe_cfg* cfg);
size_t mmoe_tc_workspace_bytes(const dmt_mmoe_cfg* cfg);
bool mmoe_tc_supported(const dmt_mmoe_cfg* cfg, const char** why);
int mmoe_tc_prepare(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, void* prepared, cudaStream_t st);
int mmoe_tc_launch(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const float* x, int64_t x_ld,
                   const void* xb_in, int64_t xb_ld, float* logits, void* workspace, const void* prepared,
                   cudaStream_t st);

}  // namespace dmt

extern "C" {

size_t dmt_mmoe_workspace_bytes(const dmt_mmoe_cfg* cfg) {
  if (!cfg) return 0;
  if (cfg->precision == DMT_PRECISION_BF16) return dmt::mmoe_tc_workspace_bytes(cfg);
  size_t floats = 0;
  for (int l = 0; l < cfg->n_layers && l < DMT_MAX_LAYERS; ++l)
    floats += (size_t)cfg->n_experts * cfg->batch * cfg->units[l];
  return floats * sizeof(float) + 256;
}

size_t dmt_mmoe_prepared_bytes(const dmt_mmoe_cfg* cfg) {
  if (!cfg || cfg->precision != DMT_PRECISION_BF16) return 0;
  return dmt::mmoe_tc_prepared_bytes(cfg);
}

int dmt_mmoe_prepare_weights(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, void* prepared, size_t prepared_bytes,
                             void* stream) {
  DMT_REQUIRE(cfg && w && prepared, DMT_ERR_INVALID_ARGUMENT, "dmt_mmoe_prepare_weights: null pointer");
  DMT_REQUIRE(cfg->precision == DMT_PRECISION_BF16, DMT_ERR_INVALID_ARGUMENT,
              "dmt_mmoe_prepare_weights: only the bf16 path has prepared weights");
  DMT_REQUIRE(cfg->n_experts > 0 && cfg->n_experts <= DMT_MAX_EXPERTS && cfg->n_layers > 0 &&
                  cfg->n_layers <= DMT_MAX_LAYERS,
              DMT_ERR_INVALID_ARGUMENT, "dmt_mmoe_prepare_weights: configuration out of range");
  DMT_REQUIRE(prepared_bytes >= dmt::mmoe_tc_prepared_bytes(cfg), DMT_ERR_WORKSPACE_TOO_SMALL,
              "dmt_mmoe_prepare_weights: buffer %zu < %zu bytes", prepared_bytes, dmt::mmoe_tc_prepared_bytes(cfg));
  DMT_REQUIRE(((uintptr_t)prepared & 255) == 0, DMT_ERR_INVALID_ARGUMENT, "dmt_mmoe_prepare_weights: unaligned buffer");
  return dmt::mmoe_tc_prepare(cfg, w, prepared, (cudaStream_t)stream);
}

int dmt_mmoe_fwd(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const float* x, int64_t x_ld, float* logits,
                 void* workspace, size_t workspace_bytes, const void* prepared, void* stream) {
  DMT_REQUIRE(cfg && w && x && logits, DMT_ERR_INVALID_ARGUMENT, "dmt_mmoe_fwd: null pointer");
  DMT_REQUIRE(cfg->batch >= 0 && cfg->in_dim > 0 && cfg->n_experts > 0 && cfg->n_experts <= DMT_MAX_EXPERTS &&
                  cfg->n_layers > 0 && cfg->n_layers <= DMT_MAX_LAYERS && cfg->n_tasks > 0 &&
                  cfg->n_tasks <= DMT_MAX_TASKS && cfg->n_tower_layers >= 0 && cfg->n_tower_layers <= DMT_MAX_LAYERS,
              DMT_ERR_INVALID_ARGUMENT, "dmt_mmoe_fwd: configuration out of range");
  DMT_REQUIRE(x_ld >= cfg->in_dim, DMT_ERR_INVALID_ARGUMENT, "dmt_mmoe_fwd: x_ld < in_dim");
  DMT_REQUIRE(workspace && workspace_bytes >= dmt_mmoe_workspace_bytes(cfg), DMT_ERR_WORKSPACE_TOO_SMALL,
              "dmt_mmoe_fwd: workspace %zu < %zu bytes", workspace_bytes, dmt_mmoe_workspace_bytes(cfg));
  if (cfg->batch == 0) return DMT_OK;
  if (cfg->precision == DMT_PRECISION_F32)
    return dmt::mmoe_f32_launch(cfg, w, x, x_ld, logits, (float*)workspace, (cudaStream_t)stream);
  DMT_REQUIRE(cfg->precision == DMT_PRECISION_BF16, DMT_ERR_INVALID_ARGUMENT, "dmt_mmoe_fwd: precision %d",
              cfg->precision);
  const char* why = nullptr;
  DMT_REQUIRE(dmt::mmoe_tc_supported(cfg, &why), DMT_ERR_UNSUPPORTED_SHAPE, "dmt_mmoe_fwd: %s", why);
  DMT_REQUIRE(prepared, DMT_ERR_INVALID_ARGUMENT,
              "dmt_mmoe_fwd(bf16): pass the buffer written by dmt_mmoe_prepare_weights");
  DMT_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)prepared & 255) == 0, DMT_ERR_INVALID_ARGUMENT,
              "dmt_mmoe_fwd(bf16): workspace / prepared must be 256-byte aligned");
  return dmt::mmoe_tc_launch(cfg, w, x, x_ld, nullptr, 0, logits, workspace, prepared, (cudaStream_t)stream);
}

int dmt_mmoe_fwd_bf16in(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const void* xb, int64_t xb_ld, float* logits,
                        void* workspace, size_t workspace_bytes, const void* prepared, void* stream) {
  DMT_REQUIRE(cfg && w && xb && logits, DMT_ERR_INVALID_ARGUMENT, "dmt_mmoe_fwd_bf16in: null pointer");
  DMT_REQUIRE(cfg->batch >= 0 && cfg->in_dim > 0 && cfg->n_experts > 0 && cfg->n_experts <= DMT_MAX_EXPERTS &&
                  cfg->n_layers > 0 && cfg->n_layers <= DMT_MAX_LAYERS && cfg->n_tasks > 0 &&
                  cfg->n_tasks <= DMT_MAX_TASKS && cfg->n_tower_layers >= 0 && cfg->n_tower_layers <= DMT_MAX_LAYERS,
              DMT_ERR_INVALID_ARGUMENT, "dmt_mmoe_fwd_bf16in: configuration out of range");
  DMT_REQUIRE(cfg->precision == DMT_PRECISION_BF16, DMT_ERR_INVALID_ARGUMENT,
              "dmt_mmoe_fwd_bf16in: only the bf16 path reads a bf16 input (precision %d)", cfg->precision);
  DMT_REQUIRE(xb_ld >= cfg->in_dim && xb_ld % 8 == 0 && ((uintptr_t)xb & 15) == 0, DMT_ERR_INVALID_ARGUMENT,
              "dmt_mmoe_fwd_bf16in: xb_ld=%lld must be a multiple of 8 and >= in_dim, xb 16-byte aligned",
              (long long)xb_ld);
  DMT_REQUIRE(workspace && workspace_bytes >= dmt_mmoe_workspace_bytes(cfg), DMT_ERR_WORKSPACE_TOO_SMALL,
              "dmt_mmoe_fwd_bf16in: workspace %zu < %zu bytes", workspace_bytes, dmt_mmoe_workspace_bytes(cfg));
  if (cfg->batch == 0) return DMT_OK;
  const char* why = nullptr;
  DMT_REQUIRE(dmt::mmoe_tc_supported(cfg, &why), DMT_ERR_UNSUPPORTED_SHAPE, "dmt_mmoe_fwd_bf16in: %s", why);
  DMT_REQUIRE(prepared, DMT_ERR_INVALID_ARGUMENT,
              "dmt_mmoe_fwd_bf16in: pass the buffer written by dmt_mmoe_prepare_weights");
  DMT_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)prepared & 255) == 0, DMT_ERR_INVALID_ARGUMENT,
              "dmt_mmoe_fwd_bf16in: workspace / prepared must be 256-byte aligned");
  return dmt::mmoe_tc_launch(cfg, w, nullptr, 0, xb, xb_ld, logits, workspace, prepared, (cudaStream_t)stream);
}

}  // extern "C"
