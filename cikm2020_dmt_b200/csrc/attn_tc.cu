// Self-attention of the row-batched pipeline on tcgen05 (kind::tf32), forward: softmax(Q_h K_h^T / sqrt(dk)) V_h over
// the valid keys of each sample, + residual, LayerNorm (TransformerModel_util.py:11-56,160-209).  Replaces the
// one-CTA-per-sample SIMT kernel (attn_fwd_kernel) for d_model = H * 32 when the caller guarantees that no sequence
// is longer than LP = min(transformer_maxlen_k, 64) (DMT_SEQ_LEN_EXACT).
//
// A tile = 128 / SLOT consecutive samples, each in its own SLOT-row slot (SLOT = 16 / 32 / 64 >= the longest sequence
// of the batch): no preprocessing pass, every index is a shift, and a row's key window is the 32-aligned, warp-uniform
// column range of its slot (tcgen05.ld / st are warp-collective and take aligned chunks).  One TMA box per (sample,
// operand, head) fetches the sample's first SLOT token rows from the CSR-packed qkv matrix into its slot: Q_h / K_h as
// K-major SWIZZLE_128B images ({32 fp32 = dk, SLOT rows}), V_h as an MN-major SWIZZLE_128B_ATOM_32B image.  Rows past
// a sample's length are its neighbours' (or zero-filled) rows: masked as keys, never stored as queries.
//
// Per tile and head:  S = Q_h K_h^T  (tcgen05.mma, 128 x 128 x 32, accumulator in TMEM columns [0,128))
//                     one thread = one query row = one TMEM lane: it reads the columns of its slot, does the masked
//                     softmax (+ dropout) in registers and writes the fp32 probabilities back IN PLACE, unnormalised
//                     (zeros outside the slot): P_h is the A operand of the next MMA
//                     O_h = P_h V_h  (A from tensor memory, 128 x 32 x 128, columns [128 + 32 h, +32)), x 1 / sum
// then z1 = O + residual, LayerNorm -> a (both saved for the backward).  128 threads, 96 KB of operands: two CTAs per
// SM overlap one tile's softmax with the other's MMAs / loads.
#include <cuda.h>
#include <stdlib.h>

#include "dropout.cuh"
#include "gemm_tf32.cuh"
#include "umma.cuh"

namespace dmt {

using namespace umma;

int make_map_f32_public(CUtensorMap* tm, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                        bool mn_major);

namespace {

constexpr int kDK = 32, kBox = 128 * 128;   // the operand image of one head: 128 rows x 128 B = 16 KB

struct AttnTcArgs {
  CUtensorMap tmQK;   // qkv [T, 3D], SWIZZLE_128B, box {32, SLOT}
  CUtensorMap tmV;    // qkv [T, 3D], SWIZZLE_128B_ATOM_32B, box {32, SLOT}
  const float* h;     // [T, D] block input (residual)
  const float* gamma;
  const float* beta;
  float* z1;          // [T, D]
  float* a;           // [T, D]
  const int32_t* offsets;
  int32_t B, D, H, LP, n_tiles;
  int64_t T;
  Dropout drop;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem], tf32: A = fp32 values of row i in TMEM lane i, one K element per column
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int H, int SLOT>
__global__ void __launch_bounds__(128, 2) attn_fwd_tc_kernel(const __grid_constant__ AttnTcArgs g) {
  constexpr int D = H * kDK, NS = 128 / SLOT;
  constexpr int CW = SLOT < 32 ? 32 : SLOT;          // columns a warp reads: the slot(s) its 32 rows belong to
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_off[NS + 1];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sQ = smem_u32(smem), sK = sQ + H * kBox, sV = sK + H * kBox;
  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmQK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmV) : "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t idesc_s = make_idesc_tf32(128, 128, false, false), idesc_o = make_idesc_tf32(128, kDK, false, true);
  const float sl2 = (1.0f / sqrtf((float)kDK)) * 1.4426950408889634f;   // softmax(s / sqrt(dk)) through exp2
  uint32_t ph_load = 0, ph_mma = 0;
  const int slot = tid / SLOT, pos = tid % SLOT;
  const int wcol = (tid & ~31) / CW * CW;            // first column of the chunk(s) this warp reads (32-aligned)

  for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
    const int b0 = tile * NS;
    if (tid <= NS) s_off[tid] = __ldg(g.offsets + min(b0 + tid, g.B));
    __syncthreads();
    if (tid < NS) {
      // one lane per sample: its SLOT rows of Q_h, K_h, V_h for every head (rows past the tensor are zero-filled)
      if (tid == 0) mbar_expect_tx(&bar_load, 3 * H * kBox);
      __syncwarp((1u << NS) - 1u);
      const int row = s_off[tid];
      const uint32_t so = (uint32_t)tid * SLOT * 128;
      for (int hh = 0; hh < H; ++hh) {
        tma_load_2d(sQ + hh * kBox + so, &g.tmQK, hh * kDK, row, &bar_load);
        tma_load_2d(sK + hh * kBox + so, &g.tmQK, D + hh * kDK, row, &bar_load);
        tma_load_2d(sV + hh * kBox + so, &g.tmV, 2 * D + hh * kDK, row, &bar_load);
      }
    }
    const int b = b0 + slot;
    int L = (b < g.B) ? s_off[slot + 1] - s_off[slot] : 0;
    if (L > g.LP) L = g.LP;
    if (L > SLOT) L = SLOT;
    const bool valid = pos < L;
    const int kbeg = slot * SLOT - wcol;               // my window inside the warp's chunk: [kbeg, kbeg + L)
    float y[D];
    mbar_wait(&bar_load, ph_load);
    ph_load ^= 1;
#pragma unroll
    for (int hh = 0; hh < H; ++hh) {
      if (tid == 0) {
        fence_after_sync();
#pragma unroll
        for (int k = 0; k < kDK / 8; ++k)
          mma_tf32_ss(tbase, make_smem_desc(sQ + hh * kBox + k * 32, 16, 1024, kLayoutSW128),
                      make_smem_desc(sK + hh * kBox + k * 32, 16, 1024, kLayoutSW128), idesc_s, k != 0);
        commit(&bar_mma);
      }
      mbar_wait(&bar_mma, ph_mma);
      ph_mma ^= 1;
      fence_after_sync();
      // ---- masked softmax over my sample's keys (TransformerModel_util.py:36-51): columns [kbeg, kbeg + L) of the
      //      warp's chunk; probabilities go back unnormalised, the denominator is taken before dropout
      uint32_t p[CW];
#pragma unroll
      for (int c = 0; c < CW; c += 32) tmem_ld32(tmem_addr(tbase, wcol + c), p + c);
      tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < CW; ++j)
        if (j >= kbeg && j < kbeg + L) mx = fmaxf(mx, __uint_as_float(p[j]));
      float sum = 0.f;
      const uint32_t didx = (uint32_t)(((b * H + hh) * g.LP + pos) * g.LP);
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        float e = 0.f;
        if (valid && j >= kbeg && j < kbeg + L) {
          e = ex2_approx((__uint_as_float(p[j]) - mx) * sl2);
          sum += e;
          if (g.drop.on) e *= g.drop.mult(didx + (uint32_t)(j - kbeg));
        }
        p[j] = __float_as_uint(e);
      }
      const float inv = valid ? 1.0f / sum : 0.f;
      {
        uint32_t z[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) z[j] = 0u;
#pragma unroll
        for (int c = 0; c < 128; c += 32) {              // warp-uniform choice per 32-column chunk
          if (c >= wcol && c < wcol + CW) tmem_st32(tmem_addr(tbase, c), p + (c % CW));   // wcol is a multiple of CW
          else tmem_st32(tmem_addr(tbase, c), z);
        }
      }
      tmem_st_wait();
      fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        fence_after_sync();
#pragma unroll
        for (int k = 0; k < 128 / 8; ++k)
          mma_tf32_ts(tbase + 128 + hh * kDK, tbase + k * 8,
                      make_smem_desc(sV + hh * kBox + k * 1024, kBox, 512, kLayoutSW128Base32B), idesc_o, k != 0);
        commit(&bar_mma);
      }
      mbar_wait(&bar_mma, ph_mma);
      ph_mma ^= 1;
      fence_after_sync();
      uint32_t o[32];
      tmem_ld32(tmem_addr(tbase, 128 + hh * kDK), o);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < kDK; ++e) y[hh * kDK + e] = __uint_as_float(o[e]) * inv;
      fence_before_sync();                               // the next head's S MMA overwrites columns [0,128)
      __syncthreads();
    }
    // ---- + residual, LayerNorm (TransformerModel_util.py:204-207)
    if (valid) {
      const int64_t row = (int64_t)s_off[slot] + pos;
      const float4* hr = reinterpret_cast<const float4*>(g.h + row * D);
      float4* zr = reinterpret_cast<float4*>(g.z1 + row * D);
      float s0 = 0.f;
#pragma unroll
      for (int v = 0; v < D / 4; ++v) {
        const float4 t = ld_stream4(reinterpret_cast<const float*>(hr + v));
        y[4 * v] += t.x; y[4 * v + 1] += t.y; y[4 * v + 2] += t.z; y[4 * v + 3] += t.w;
        zr[v] = make_float4(y[4 * v], y[4 * v + 1], y[4 * v + 2], y[4 * v + 3]);
        s0 += (y[4 * v] + y[4 * v + 1]) + (y[4 * v + 2] + y[4 * v + 3]);
      }
      const float mean = s0 * (1.0f / D);
      float q0 = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const float dl = y[c] - mean;
        q0 = fmaf(dl, dl, q0);
      }
      const float rstd = 1.0f / sqrtf(q0 * (1.0f / D) + kLnEps);
      float4* ar = reinterpret_cast<float4*>(g.a + row * D);
#pragma unroll
      for (int v = 0; v < D / 4; ++v) {
        const float4 gg = ldg4(g.gamma + 4 * v), bb = ldg4(g.beta + 4 * v);
        ar[v] = make_float4(fmaf(gg.x, (y[4 * v] - mean) * rstd, bb.x), fmaf(gg.y, (y[4 * v + 1] - mean) * rstd, bb.y),
                            fmaf(gg.z, (y[4 * v + 2] - mean) * rstd, bb.z), fmaf(gg.w, (y[4 * v + 3] - mean) * rstd, bb.w));
      }
    }
    __syncthreads();                                     // s_off / operand images are rewritten by the next tile
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

template <int SLOT>
int launch_attn(AttnTcArgs& g, const float* qkv, int D, cudaStream_t st) {
  int rc = make_map_f32_public(&g.tmQK, qkv, g.T, 3 * D, 3 * D, SLOT, false);
  if (rc != DMT_OK) return rc;
  rc = make_map_f32_public(&g.tmV, qkv, g.T, 3 * D, 3 * D, SLOT, true);
  if (rc != DMT_OK) return rc;
  constexpr int NS = 128 / SLOT;
  g.n_tiles = (g.B + NS - 1) / NS;
  const int smem = 3 * 2 * kBox + 1024;
  cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<2, SLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn_fwd_tc_kernel)");
  const int cap = 2 * sm_count_cached();
  const int grid = g.n_tiles < cap ? g.n_tiles : cap;
  attn_fwd_tc_kernel<2, SLOT><<<grid, 128, smem, st>>>(g);
  DMT_CUDA_LAUNCH_CHECK("attn_fwd_tc_kernel");
  return DMT_OK;
}

}  // namespace

bool attn_fwd_tc_supported(const dmt_seq_cfg& c) {
  static int off = -1;                       // diagnostics: DMT_ATTN_SIMT=1 keeps the per-sample kernels
  if (off < 0) {
    const char* e = getenv("DMT_ATTN_SIMT");
    off = (e && e[0] == '1') ? 1 : 0;
  }
  if (off) return false;
  const int LP = c.maxlen < DMT_MAX_SEQ_LEN ? c.maxlen : DMT_MAX_SEQ_LEN;
  return (c.flags & DMT_SEQ_LEN_EXACT) && c.num_heads == 2 && c.d_model == 64 && LP <= 64;
}

int attn_fwd_tc_launch(const dmt_seq_cfg& c, const float* qkv, const float* h, const float* gamma, const float* beta,
                       float* z1, float* a, const int32_t* offsets, int64_t T, int LP, const Dropout& drop,
                       cudaStream_t st) {
  if (T <= 0 || c.batch <= 0) return DMT_OK;
  AttnTcArgs g;
  g.h = h; g.gamma = gamma; g.beta = beta; g.z1 = z1; g.a = a; g.offsets = offsets;
  g.B = c.batch; g.D = c.d_model; g.H = c.num_heads; g.LP = LP; g.T = T;
  g.drop = drop;
  int bound = c.slot_len > 0 ? c.slot_len : LP;        // DMT_SEQ_LEN_EXACT: no sequence is longer than this
  if (bound > LP) bound = LP;
  if (bound <= 16) return launch_attn<16>(g, qkv, c.d_model, st);
  if (bound <= 32) return launch_attn<32>(g, qkv, c.d_model, st);
  return launch_attn<64>(g, qkv, c.d_model, st);
}

}  // namespace dmt
