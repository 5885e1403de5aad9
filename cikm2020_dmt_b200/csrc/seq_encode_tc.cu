// dmt_seq_encode_fwd, DMT_PRECISION_BF16: persistent fused tile kernel on tcgen05 tensor cores.
//
// One CTA per SM loops over 128-row tiles.  A tile packs NS = 128/SLOT samples of one behaviour
// sequence into fixed SLOT-row slots (SLOT = 16/32/64 >= longest sequence in the batch), so every
// index is a shift and cross-sample attention is a block-diagonal mask.  Per tile, everything
// between the embedding rows and the [d_model] interest vectors stays on chip:
//
//   gather+concat+sqrt(d)+pos -> X (bf16, smem, canonical K-major image)
//   tcgen05.mma  X  x Wqkv            -> TMEM -> +bias -> Q,K,V images (smem)
//   tcgen05.mma  Q_h x K_h^T          -> TMEM -> masked softmax in registers (1 thread = 1 row) -> P_h
//   tcgen05.mma  P_h x V_h (MN-major) -> TMEM -> +X, LayerNorm -> A (smem, over X)
//   tcgen05.mma  A x W1 -> relu -> H ; tcgen05.mma H x W2 -> +A, LayerNorm -> memory (fp32, smem)
//   decoder (single query per sample): algebraically folded K/V projections, CUDA cores
//
// HBM traffic per (sample, sequence) is the algorithmic minimum: ids + embedding rows in, one
// interest vector out.  Weights live in shared memory for the lifetime of the CTA (bf16 images
// produced once by dmt_seq_prepare_weights).
#include <limits.h>
#include <stdlib.h>

#include "dmt_common.cuh"
#include "seq_tc.cuh"
#include "umma.cuh"

namespace dmt {

using namespace umma;

constexpr int kTcThreads = 256;

// Weight images (all bf16), "image(N, K)" = [k/8][n][8] with element (n, k) = W_tf[k][n] unless noted:
//   wqkv  image(3D, D)   columns n = [Q | K | V]            (tcgen05 B operand, K-major)
//   w1    image(DFF, D)                                      (tcgen05 B operand + decoder FF mat-vec)
//   w2    image(D, DFF)                                      (tcgen05 B operand + decoder FF mat-vec)
//   G     image(H*D, D)  G[(h,k)][j] = sum_{c in head h} Wq[j][c] Wk[k][c]   -- the decoder's query and
//                        key projections folded:  score_t = M_t . (dvec G_h + g_h)  (+ a constant per
//                        (sample, head) that softmax ignores)
//   dv    image(D, D)    decoder Wv                          (mat-vec  o = ctx Wv)
//   g     fp32 [H*D]     g[(h,k)] = sum_{c in head h} bq[c] Wk[k][c]
//   wvbd  image(D, H*D)  element (n = c, k = (h, j)) = Wv_dec[j][c] if column c belongs to head h, else 0 (v2 tail)
__global__ void seq_prepare_kernel(const float* __restrict__ wq, const float* __restrict__ wk,
                                   const float* __restrict__ wv, const float* __restrict__ w1,
                                   const float* __restrict__ w2, const float* __restrict__ dq,
                                   const float* __restrict__ dbq, const float* __restrict__ dk,
                                   const float* __restrict__ dv, __nv_bfloat16* __restrict__ out, int D, int DFF,
                                   int H) {
  const size_t n_qkv = prep_wqkv(D), n_w1 = prep_w1(D, DFF), n_dd = (size_t)D * D, n_g = (size_t)H * D * D;
  const size_t total_bf16 = n_qkv + 2 * n_w1 + n_g + n_dd;
  const int DK = D / H;
  const size_t off_wvbd = prep_off_wvbd(D, DFF, H), n_wvbd = prep_wvbd(D, H);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_bf16 + (size_t)H * D + n_wvbd;
       i += (size_t)gridDim.x * blockDim.x) {
    float v;
    size_t j = i;
    if (i >= total_bf16 + (size_t)H * D) {  // wvbd image(D, H*D): [k/8][n][8]
      const size_t q = i - total_bf16 - (size_t)H * D;
      const int e = q % 8, n = (q / 8) % D, kc = q / (8 * D), k = kc * 8 + e, h = k / D, jj = k % D;
      out[off_wvbd + q] = __float2bfloat16((n / DK == h) ? dv[(size_t)jj * D + n] : 0.f);
      continue;
    }
    if (i >= total_bf16) {                 // g[(h,k)] fp32, stored right after the bf16 images
      const int hk = (int)(i - total_bf16), h = hk / D, k = hk % D;
      float acc = 0.f;
      for (int c = h * DK; c < (h + 1) * DK; ++c) acc = fmaf(dbq[c], dk[(size_t)k * D + c], acc);
      reinterpret_cast<float*>(out + total_bf16)[hk] = acc;
      continue;
    }
    if (j < n_qkv) {                       // image(3D, D)
      const int e = j % 8, n = (j / 8) % (3 * D), kc = j / (8 * 3 * D), k = kc * 8 + e;
      const float* src = n < D ? wq : (n < 2 * D ? wk : wv);
      v = src[(size_t)k * D + (n % D)];
    } else if ((j -= n_qkv) < n_w1) {      // image(DFF, D): W1 [D, DFF]
      const int e = j % 8, n = (j / 8) % DFF, kc = j / (8 * DFF), k = kc * 8 + e;
      v = w1[(size_t)k * DFF + n];
    } else if ((j -= n_w1) < n_w1) {       // image(D, DFF): W2 [DFF, D]
      const int e = j % 8, n = (j / 8) % D, kc = j / (8 * D), k = kc * 8 + e;
      v = w2[(size_t)k * D + n];
    } else if ((j -= n_w1) < n_g) {        // image(H*D, D): G[(h,k)][jj]
      const int e = j % 8, n = (j / 8) % (H * D), jc = j / (8 * H * D), jj = jc * 8 + e;
      const int h = n / D, k = n % D;
      float acc = 0.f;
      for (int c = h * DK; c < (h + 1) * DK; ++c) acc = fmaf(dq[(size_t)jj * D + c], dk[(size_t)k * D + c], acc);
      v = acc;
    } else {                               // image(D, D): decoder Wv
      j -= n_g;
      const int e = j % 8, n = (j / 8) % D, kc = j / (8 * D), k = kc * 8 + e;
      v = dv[(size_t)k * D + n];
    }
    out[i] = __float2bfloat16(v);
  }
}

__device__ __forceinline__ void bf16x8_to_float(const uint4& v, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ uint4 float8_to_bf16(const float* f) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]);
  v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]);
  v.w = pack_bf16x2(f[6], f[7]);
  return v;
}

// Row-local LayerNorm over N values held in registers (TransformerModel_util.py:58-78).
template <int N>
__device__ __forceinline__ void ln_inplace(float* y, const float* __restrict__ g, const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) s += y[i];
  const float mean = s * (1.0f / N);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float d = y[i] - mean;
    q += d * d;
  }
  const float rstd = 1.0f / sqrtf(q * (1.0f / N) + kLnEps);
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = g[i] * ((y[i] - mean) * rstd) + b[i];
}

template <int D, int DFF, int H, int SLOT>
struct TcLayout {
  static constexpr int NS = 128 / SLOT;             // samples per tile
  static constexpr int DK = D / H;
  static constexpr int KC = D / 8;                  // 16-byte chunks per token row
  static constexpr int ROWB = 128 * 16;             // bytes of one chunk column of a 128-row image
  static constexpr int MLD = D + 4;                 // fp32 row stride of the encoder memory
  // shared-memory byte offsets
  static constexpr int oWqkv = 0;
  static constexpr int oW1 = oWqkv + 3 * D * D * 2;
  static constexpr int oW2 = oW1 + D * DFF * 2;
  static constexpr int oXA = oW2 + DFF * D * 2;               // X, then A (attention output)
  static constexpr int szR2a = 3 * 128 * D * 2, szR2b = 128 * DFF * 2, szR2c = 128 * MLD * 4;
  static constexpr int szR2 = szR2a > szR2b ? (szR2a > szR2c ? szR2a : szR2c) : (szR2b > szR2c ? szR2b : szR2c);
  static constexpr int oR2 = oXA + 128 * D * 2;               // Q|K|V images -> H image -> memory fp32
  static constexpr int oP0 = oR2 + szR2;                      // P of head 0 (head 1 reuses the Q|K images)
  static constexpr int oFV = oP0 + 128 * 128 * 2;             // fp32 vectors
  // fp32 vector slots (floats)
  static constexpr int vBQKV = 0, vB1 = 3 * D, vB2 = vB1 + DFF, vLN = vB2 + D;   // 6 LN vectors
  static constexpr int vDB = vLN + 6 * D;                     // decoder biases bq|bk|bv
  static constexpr int vDVEC = vDB + 3 * D;                   // [NS][D] scaled target embeddings
  static constexpr int vDEC = vDVEC + NS * D;                 // decoder scratch for 2 samples
  static constexpr int dQT = 0, dSC = dQT + 2 * H * D, dCTX = dSC + 2 * H * SLOT, dOV = dCTX + 2 * H * D,
                       dAV = dOV + 2 * D, dHV = dAV + 2 * D, dPART = dHV + 2 * DFF, dEND = dPART + 4 * D;
  // decoder weights are bulk-copied per tile into the P0 buffer once the PV MMAs have consumed it
  static constexpr int decBytes = (H * D * D + D * D) * 2 + H * D * 4;
  static_assert(decBytes <= 128 * 128 * 2 && decBytes % 16 == 0, "decoder weights must fit the P0 buffer");
  static constexpr int nFV = vDEC + dEND;
  static constexpr int oLen = oFV + nFV * 4;                  // int32 [NS] lengths
  static constexpr int total = oLen + NS * 4 + 64;
  // TMEM columns
  static constexpr int tQKV = 0, tO = 192 < 3 * D ? 3 * D : 192, tS = 256, tFF1 = 0, tFF2 = 256;
};

#define DMT_TICK(idx)                                                   \
  do {                                                                  \
    if (a.dbg && tid == 0) {                                            \
      const long long _now = clock64();                                 \
      atomicAdd(a.dbg + (idx), (unsigned long long)(_now - t_last));    \
      t_last = _now;                                                    \
    }                                                                   \
  } while (0)

template <int D, int DFF, int H, int SLOT>
__global__ void __launch_bounds__(kTcThreads, 1) seq_encode_tc_kernel(const __grid_constant__ SeqTcArgs a) {
  using L = TcLayout<D, DFF, H, SLOT>;
  static_assert(H == 2, "v1 tensor-core path: two heads");
  static_assert(L::DK % 16 == 0 && D % 16 == 0 && DFF % 16 == 0 && 3 * D <= 256 && DFF <= 256, "tile shape");
  static_assert(L::tO + H * L::DK <= 256 && L::tS + H * 128 <= 512, "TMEM plan");
  constexpr int NS = L::NS, DK = L::DK, KC = L::KC, ROWB = L::ROWB, MLD = L::MLD;
  constexpr int CW = SLOT < 32 ? 32 : SLOT;        // score columns a warp loads (covers its rows' slots)

  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, wbar;
  __shared__ uint32_t tmem_base_s;
  float* fv = reinterpret_cast<float*>(smem + L::oFV);
  int* slen = reinterpret_cast<int*>(smem + L::oLen);
  uint8_t* sXA = smem + L::oXA;
  uint8_t* sR2 = smem + L::oR2;
  uint8_t* sP0 = smem + L::oP0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2;                       // warps w and w+4 share TMEM lanes; they split columns/heads
  const int row = (warp & 3) * 32 + lane;           // accumulator row == TMEM lane of this thread
  const int B = a.cfg.batch;

  // ---- one-time setup: TMEM, barrier, resident weights and vectors ----
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&wbar, 1);
    mbar_fence_init();
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.prepared);
    uint4* dst = reinterpret_cast<uint4*>(smem + L::oWqkv);
    constexpr int n16 = (3 * D * D + 2 * D * DFF) * 2 / 16;
    for (int i = tid; i < n16; i += kTcThreads) dst[i] = __ldg(src + i);
    for (int i = tid; i < D; i += kTcThreads) {
      fv[L::vBQKV + i] = a.bq[i];
      fv[L::vBQKV + D + i] = a.bk[i];
      fv[L::vBQKV + 2 * D + i] = a.bv[i];
      fv[L::vB2 + i] = a.b2[i];
      fv[L::vLN + 0 * D + i] = a.ln1_g[i];
      fv[L::vLN + 1 * D + i] = a.ln1_b[i];
      fv[L::vLN + 2 * D + i] = a.ln2_g[i];
      fv[L::vLN + 3 * D + i] = a.ln2_b[i];
      fv[L::vLN + 4 * D + i] = a.ln3_g[i];
      fv[L::vLN + 5 * D + i] = a.ln3_b[i];
      fv[L::vDB + i] = a.dbq[i];
      fv[L::vDB + D + i] = a.dbk[i];
      fv[L::vDB + 2 * D + i] = a.dbv[i];
    }
    for (int i = tid; i < DFF; i += kTcThreads) fv[L::vB1 + i] = a.b1[i];
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t aWqkv = smem_u32(smem + L::oWqkv), aW1 = smem_u32(smem + L::oW1), aW2 = smem_u32(smem + L::oW2);
  const uint32_t aXA = smem_u32(sXA), aR2 = smem_u32(sR2), aP0 = smem_u32(sP0);
  // loop-invariant descriptor words: K-major images use LBO = one chunk column, SBO = 128 B
  const uint32_t dHi = desc_hi(128, kLayoutNone), dHiV = desc_hi(ROWB, kLayoutNone);
  const uint32_t dXA = desc_lo(aXA, ROWB), dR2 = desc_lo(aR2, ROWB), dP0 = desc_lo(aP0, ROWB);
  const uint32_t dWqkv = desc_lo(aWqkv, 3 * D * 16), dW1 = desc_lo(aW1, DFF * 16), dW2 = desc_lo(aW2, D * 16);
  const __nv_bfloat16* gDec = a.prepared + prep_wqkv(D) + 2 * prep_w1(D, DFF);   // G | Wv | g
  const uint8_t* sG = sP0;                                  // image(H*D, D)
  const uint8_t* sDv = sP0 + H * D * D * 2;                 // image(D, D)
  const float* sGb = reinterpret_cast<const float*>(sP0 + (H * D * D + D * D) * 2);
  const float sqrt_d = sqrtf((float)D);
  const float scale = 1.0f / sqrtf((float)DK);
  const int nf = a.cfg.n_feats;
  const int lmax = a.cfg.maxlen < SLOT ? a.cfg.maxlen : SLOT;
  uint32_t phase = 0, wphase = 0;

  // ---- software-pipelined gather: the ids / embedding rows of tile t+1 are requested while tile t is
  //      being computed (three dependent load levels -- offsets, ids, rows -- each issued just before one
  //      of the MMA waits of the current tile), so P0 only converts registers that have already landed.
  constexpr int NI = 128 * KC / kTcThreads;          // (row, chunk) items per thread; tile-invariant mapping
  static_assert(128 * KC % kTcThreads == 0 && NS * KC <= kTcThreads, "gather item mapping");
  constexpr int kInvalid = INT_MIN;
  int pf_len = 0;                                    // tid < NS: length of slot tid in the next tile
  int pf_o0[NI], pf_o1[NI], pf_l0[NI], pf_l1[NI];    // offsets of the chunk's feature / of the last feature
  int pf_id[NI];
  float4 pf_e0[NI], pf_e1[NI];
  int pf_tid = 0;                                    // target item id (tid < NS*KC)
  float4 pf_t0 = make_float4(0.f, 0.f, 0.f, 0.f), pf_t1 = pf_t0;
  uint32_t posr[NI][4];                              // learned positions of this thread's items, packed bf16
#pragma unroll
  for (int k = 0; k < NI; ++k) {
    const int i = tid + k * kTcThreads, c = i % KC, t = (i / KC) % SLOT;
    float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
    if (t < a.cfg.maxlen) {
      p0 = ldg4(a.pos + t * D + c * 8);
      p1 = ldg4(a.pos + t * D + c * 8 + 4);
    }
    posr[k][0] = pack_bf16x2(p0.x, p0.y); posr[k][1] = pack_bf16x2(p0.z, p0.w);
    posr[k][2] = pack_bf16x2(p1.x, p1.y); posr[k][3] = pack_bf16x2(p1.z, p1.w);
  }

  auto stage_offsets = [&](int nt) {                 // level 1: CSR offsets
    if (nt >= a.n_tiles) return;
    const int nb0 = nt * NS;
    if (tid < NS) {
      const int b = nb0 + tid;
      pf_len = 0;
      if (b < B) pf_len = min(__ldg(a.in.offsets[nf - 1] + b + 1) - __ldg(a.in.offsets[nf - 1] + b), lmax);
    }
#pragma unroll
    for (int k = 0; k < NI; ++k) {
      const int i = tid + k * kTcThreads, c = i % KC, r = i / KC, b = nb0 + r / SLOT;
      pf_o0[k] = pf_o1[k] = pf_l0[k] = pf_l1[k] = 0;
      if (b < B) {
        const int f = a.chunk_feat[c];
        pf_o0[k] = __ldg(a.in.offsets[f] + b);
        pf_o1[k] = __ldg(a.in.offsets[f] + b + 1);
        pf_l0[k] = __ldg(a.in.offsets[nf - 1] + b);
        pf_l1[k] = __ldg(a.in.offsets[nf - 1] + b + 1);
      }
    }
    if (tid < NS * KC) {
      const int c = tid % KC, b = nb0 + tid / KC;
      pf_tid = kInvalid;
      if (b < B) pf_tid = __ldg(a.in.item_ids[a.chunk_feat[c]] + b);
    }
  };
  auto stage_ids = [&](int nt) {                     // level 2: token ids
    if (nt >= a.n_tiles) return;
#pragma unroll
    for (int k = 0; k < NI; ++k) {
      const int i = tid + k * kTcThreads, c = i % KC, t = (i / KC) % SLOT;
      pf_id[k] = kInvalid;                           // padded position: X row stays exactly zero
      if (t < min(pf_l1[k] - pf_l0[k], lmax))
        pf_id[k] = (t < pf_o1[k] - pf_o0[k]) ? __ldg(a.in.ids[a.chunk_feat[c]] + pf_o0[k] + t) : 0;
    }
  };
  auto stage_rows = [&](int nt) {                    // level 3: embedding rows (the HBM traffic)
    if (nt >= a.n_tiles) return;
#pragma unroll
    for (int k = 0; k < NI; ++k) {
      const int c = (tid + k * kTcThreads) % KC, f = a.chunk_feat[c];
      pf_e0[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      pf_e1[k] = pf_e0[k];
      const int64_t rw = (int64_t)pf_id[k] - (a.cfg.zero_pad ? 1 : 0);
      if (pf_id[k] != kInvalid && rw >= 0 && rw < a.in.rows[f]) {
        const float* src = a.in.table[f] + rw * a.in.dim[f] + a.chunk_off[c];
        pf_e0[k] = ld_stream4(src);
        pf_e1[k] = ld_stream4(src + 4);
      }
    }
    if (tid < NS * KC) {
      const int c = tid % KC, f = a.chunk_feat[c];
      pf_t0 = make_float4(0.f, 0.f, 0.f, 0.f);
      pf_t1 = pf_t0;
      const int64_t rw = (int64_t)pf_tid - (a.cfg.zero_pad ? 1 : 0);
      if (pf_tid != kInvalid && rw >= 0 && rw < a.in.rows[f]) {
        const float* src = a.in.table[f] + rw * a.in.dim[f] + a.chunk_off[c];
        pf_t0 = ld_stream4(src);
        pf_t1 = ld_stream4(src + 4);
      }
    }
  };
  stage_offsets(blockIdx.x);
  stage_ids(blockIdx.x);
  stage_rows(blockIdx.x);

  long long t_last = clock64();
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int b0 = tile * NS;
    const int next_tile = tile + gridDim.x;
    if (tid < NS) slen[tid] = pf_len;
    __syncthreads();
    DMT_TICK(0);

    // ---- P0: prefetched rows -> concat + sqrt(d) scale + learned position -> X image (bf16) ----
#pragma unroll
    for (int k = 0; k < NI; ++k) {
      const int i = tid + k * kTcThreads, c = i % KC, r = i / KC;
      float x[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = 0.f;
      if (pf_id[k] != kInvalid) {
        const __nv_bfloat162* pp = reinterpret_cast<const __nv_bfloat162*>(posr[k]);
        const float2 q0 = __bfloat1622float2(pp[0]), q1 = __bfloat1622float2(pp[1]);
        const float2 q2 = __bfloat1622float2(pp[2]), q3 = __bfloat1622float2(pp[3]);
        x[0] = fmaf(pf_e0[k].x, sqrt_d, q0.x); x[1] = fmaf(pf_e0[k].y, sqrt_d, q0.y);
        x[2] = fmaf(pf_e0[k].z, sqrt_d, q1.x); x[3] = fmaf(pf_e0[k].w, sqrt_d, q1.y);
        x[4] = fmaf(pf_e1[k].x, sqrt_d, q2.x); x[5] = fmaf(pf_e1[k].y, sqrt_d, q2.y);
        x[6] = fmaf(pf_e1[k].z, sqrt_d, q3.x); x[7] = fmaf(pf_e1[k].w, sqrt_d, q3.y);
      }
      *reinterpret_cast<uint4*>(sXA + c * ROWB + r * 16) = float8_to_bf16(x);
    }
    if (tid < NS * KC) {                               // target item rows -> decoder input (fp32)
      float* dv = fv + L::vDVEC + (tid / KC) * D + (tid % KC) * 8;
      dv[0] = pf_t0.x * sqrt_d; dv[1] = pf_t0.y * sqrt_d; dv[2] = pf_t0.z * sqrt_d; dv[3] = pf_t0.w * sqrt_d;
      dv[4] = pf_t1.x * sqrt_d; dv[5] = pf_t1.y * sqrt_d; dv[6] = pf_t1.z * sqrt_d; dv[7] = pf_t1.w * sqrt_d;
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    DMT_TICK(1);
    stage_offsets(next_tile);

    // ---- P1: [Q|K|V] = X Wqkv ----
    if (tid == 0) {
      fence_after_sync();
      constexpr uint32_t idesc = make_idesc_bf16(128, 3 * D);
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks)
        mma_bf16_ss(tbase + L::tQKV, desc_join(dXA + ks * (2 * ROWB / 16), dHi),
                    desc_join(dWqkv + ks * (2 * 3 * D), dHi), idesc, ks > 0);
      commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();

    DMT_TICK(2);
    // ---- P2: + bias, bf16, Q / K / V images ([chunk][row][8] each) ----
    {
      constexpr int colsPerHalf = 3 * D / 2;
      static_assert(colsPerHalf % 32 == 0, "QKV epilogue split");
#pragma unroll
      for (int blk = 0; blk < colsPerHalf / 32; ++blk) {
        const int n0 = half * colsPerHalf + blk * 32;
        uint32_t r[32];
        tmem_ld32(tmem_addr(tbase, L::tQKV + n0), r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int n = n0 + g * 8;
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = __uint_as_float(r[g * 8 + e]) + fv[L::vBQKV + n + e];
          const int m = n / D, ch = (n % D) / 8;
          *reinterpret_cast<uint4*>(sR2 + m * (128 * D * 2) + ch * ROWB + row * 16) = float8_to_bf16(y);
        }
      }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    DMT_TICK(3);
    stage_ids(next_tile);

    // ---- P3: S_h = Q_h K_h^T for both heads ----
    if (tid == 0) {
      fence_after_sync();
      constexpr uint32_t idesc = make_idesc_bf16(128, 128);
#pragma unroll
      for (int h = 0; h < H; ++h)
#pragma unroll
        for (int ks = 0; ks < DK / 16; ++ks) {
          const uint32_t ch = (h * DK) / 8 + ks * 2;
          mma_bf16_ss(tbase + L::tS + h * 128, desc_join(dR2 + ch * (ROWB / 16), dHi),
                      desc_join(dR2 + (128 * D * 2 + ch * ROWB) / 16, dHi), idesc, ks > 0);
        }
      commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();

    DMT_TICK(4);
    // ---- P4: masked softmax, one thread = one (row, head); P images ----
    {
      const int h = half;
      const int slot = row / SLOT;
      const int len = slen[slot];
      const int col0 = (row / CW) * CW;               // warp-uniform: rows of a warp share the CW window
      const int lo = slot * SLOT - col0;              // this row's keys are window columns [lo, lo + len)
      const float sl2 = scale * 1.4426950408889634f;  // softmax(s*scale) through exp2
      uint32_t r[CW];
#pragma unroll
      for (int blk = 0; blk < CW / 32; ++blk) tmem_ld32(tmem_addr(tbase, L::tS + h * 128 + col0 + blk * 32), r + blk * 32);
      tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        const bool ok = (unsigned)(j - lo) < (unsigned)len;
        const float v = ok ? __uint_as_float(r[j]) : -INFINITY;
        r[j] = __float_as_uint(v);
        mx = fmaxf(mx, v);
      }
      const float mxs = (mx == -INFINITY) ? 0.f : mx * sl2;
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        const float e = ex2_approx(fmaf(__uint_as_float(r[j]), sl2, -mxs));   // exp2(-inf) == 0 for masked keys
        r[j] = __float_as_uint(e);
        sum += e;
      }
      const float inv = sum > 0.f ? 1.0f / sum : 0.f;
      uint8_t* dstP = (h == 0) ? sP0 : sR2;            // head 1 reuses the (dead) Q|K images
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      const int kc0 = col0 / 8;
#pragma unroll
      for (int jj = 0; jj < CW / 8; ++jj) {           // the CW keys this row can attend to
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(r[jj * 8 + 0]) * inv, __uint_as_float(r[jj * 8 + 1]) * inv);
        v.y = pack_bf16x2(__uint_as_float(r[jj * 8 + 2]) * inv, __uint_as_float(r[jj * 8 + 3]) * inv);
        v.z = pack_bf16x2(__uint_as_float(r[jj * 8 + 4]) * inv, __uint_as_float(r[jj * 8 + 5]) * inv);
        v.w = pack_bf16x2(__uint_as_float(r[jj * 8 + 6]) * inv, __uint_as_float(r[jj * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(dstP + (kc0 + jj) * ROWB + row * 16) = v;
      }
#pragma unroll
      for (int kc = 0; kc < 16; ++kc)                 // every other key block of the tile: exact zeros
        if (kc < kc0 || kc >= kc0 + CW / 8) *reinterpret_cast<uint4*>(dstP + kc * ROWB + row * 16) = zero;
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();

    DMT_TICK(5);
    // ---- P5: O_h = P_h V_h (V read as an MN-major B operand straight from its image) ----
    if (tid == 0) {
      fence_after_sync();
      constexpr uint32_t idesc = make_idesc_bf16(128, DK, false, true);
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const uint32_t dP = (h == 0) ? dP0 : dR2;
        const uint32_t dV = desc_lo(aR2 + 2 * (128 * D * 2) + ((h * DK) / 8) * ROWB, 128);   // MN-major: LBO = next 8 keys
#pragma unroll
        for (int ks = 0; ks < 128 / 16; ++ks)
          mma_bf16_ss(tbase + L::tO + h * DK, desc_join(dP + ks * (2 * ROWB / 16), dHi),
                      desc_join(dV + ks * (256 / 16), dHiV), idesc, ks > 0);
      }
      commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();

    DMT_TICK(6);
    // ---- P6: A = LN(O + X) (self-attention LayerNorm), written over X ----
    if (tid == 0) {   // P0 is dead until the next tile's softmax: stream the decoder weights into it (async)
      mbar_expect_tx(&wbar, L::decBytes);
      bulk_g2s(sP0, gDec, L::decBytes, &wbar);
    }
    if (half == 0) {
      float y[D];
#pragma unroll
      for (int blk = 0; blk < D / 32; ++blk) {
        uint32_t r[32];
        tmem_ld32(tmem_addr(tbase, L::tO + blk * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) y[blk * 32 + e] = __uint_as_float(r[e]);
      }
#pragma unroll
      for (int c = 0; c < KC; ++c) {
        float x[8];
        bf16x8_to_float(*reinterpret_cast<const uint4*>(sXA + c * ROWB + row * 16), x);
#pragma unroll
        for (int e = 0; e < 8; ++e) y[c * 8 + e] += x[e];
      }
      ln_inplace<D>(y, fv + L::vLN + 0 * D, fv + L::vLN + 1 * D);
#pragma unroll
      for (int c = 0; c < KC; ++c) *reinterpret_cast<uint4*>(sXA + c * ROWB + row * 16) = float8_to_bf16(y + c * 8);
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    DMT_TICK(7);
    stage_rows(next_tile);

    // ---- P7: hidden = A W1 ----
    if (tid == 0) {
      fence_after_sync();
      constexpr uint32_t idesc = make_idesc_bf16(128, DFF);
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks)
        mma_bf16_ss(tbase + L::tFF1, desc_join(dXA + ks * (2 * ROWB / 16), dHi),
                    desc_join(dW1 + ks * (2 * DFF), dHi), idesc, ks > 0);
      commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();

    DMT_TICK(8);
    // ---- P8: relu(+b1) -> H image ----
    {
      constexpr int colsPerHalf = DFF / 2;
      static_assert(colsPerHalf % 32 == 0, "FF1 epilogue split");
#pragma unroll
      for (int blk = 0; blk < colsPerHalf / 32; ++blk) {
        const int n0 = half * colsPerHalf + blk * 32;
        uint32_t r[32];
        tmem_ld32(tmem_addr(tbase, L::tFF1 + n0), r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = fmaxf(__uint_as_float(r[g * 8 + e]) + fv[L::vB1 + n0 + g * 8 + e], 0.f);
          *reinterpret_cast<uint4*>(sR2 + ((n0 + g * 8) / 8) * ROWB + row * 16) = float8_to_bf16(y);
        }
      }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();

    DMT_TICK(9);
    // ---- P9: F = H W2 ----
    if (tid == 0) {
      fence_after_sync();
      constexpr uint32_t idesc = make_idesc_bf16(128, D);
#pragma unroll
      for (int ks = 0; ks < DFF / 16; ++ks)
        mma_bf16_ss(tbase + L::tFF2, desc_join(dR2 + ks * (2 * ROWB / 16), dHi),
                    desc_join(dW2 + ks * (2 * D), dHi), idesc, ks > 0);
      commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    fence_after_sync();

    DMT_TICK(10);
    // ---- P10: memory = LN(F + b2 + A) (feed-forward LayerNorm), fp32 row-major over the H image ----
    if (half == 0) {
      float y[D];
#pragma unroll
      for (int blk = 0; blk < D / 32; ++blk) {
        uint32_t r[32];
        tmem_ld32(tmem_addr(tbase, L::tFF2 + blk * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) y[blk * 32 + e] = __uint_as_float(r[e]) + fv[L::vB2 + blk * 32 + e];
      }
#pragma unroll
      for (int c = 0; c < KC; ++c) {
        float x[8];
        bf16x8_to_float(*reinterpret_cast<const uint4*>(sXA + c * ROWB + row * 16), x);
#pragma unroll
        for (int e = 0; e < 8; ++e) y[c * 8 + e] += x[e];
      }
      ln_inplace<D>(y, fv + L::vLN + 2 * D, fv + L::vLN + 3 * D);
      float* mrow = reinterpret_cast<float*>(sR2) + row * MLD;
#pragma unroll
      for (int c = 0; c < D / 4; ++c)
        *reinterpret_cast<float4*>(mrow + c * 4) = make_float4(y[c * 4], y[c * 4 + 1], y[c * 4 + 2], y[c * 4 + 3]);
    }
    fence_before_sync();
    __syncthreads();

    DMT_TICK(11);
    // ---- P11: decoder, two samples at a time (TransformerModel.py:125-171).  K/V projections of the
    //      memory are folded:  score_j = M_j . (Wk_h qd_h) + bk_h . qd_h ,  o_h = (sum_j p_j M_j) Wv_h + bv_h ----
    const float* Mem = reinterpret_cast<const float*>(sR2);
    float* dec = fv + L::vDEC;
    mbar_wait(&wbar, wphase);
    wphase ^= 1;
    for (int g0 = 0; g0 < NS; g0 += 2) {
      // a: qt[s][h][k] = dvec[s] . G[(h,k)] + g[(h,k)]   (query and key projections folded)
      for (int i = tid; i < 2 * H * D; i += kTcThreads) {
        const int n = i % (H * D), s = i / (H * D);
        const float* dv = fv + L::vDVEC + (g0 + s) * D;
        float acc = sGb[n];
#pragma unroll
        for (int jc = 0; jc < KC; ++jc) {
          float w[8];
          bf16x8_to_float(*reinterpret_cast<const uint4*>(sG + (jc * H * D + n) * 16), w);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc = fmaf(dv[jc * 8 + e], w[e], acc);
        }
        dec[L::dQT + s * H * D + n] = acc;
      }
      __syncthreads();
      // c: scores over the slot's valid keys
      for (int i = tid; i < 2 * H * SLOT; i += kTcThreads) {
        const int j = i % SLOT, h = (i / SLOT) % H, s = i / (SLOT * H);
        float sc = -INFINITY;
        if (j < slen[g0 + s]) {
          const float* m = Mem + ((g0 + s) * SLOT + j) * MLD;
          const float* qt = dec + L::dQT + (s * H + h) * D;
          float acc = 0.f;   // the per-(sample, head) constant bk_h . qd_h cancels in the softmax
#pragma unroll
          for (int k = 0; k < D; k += 4) {
            const float4 mv = *reinterpret_cast<const float4*>(m + k);
            acc = fmaf(mv.x, qt[k], fmaf(mv.y, qt[k + 1], fmaf(mv.z, qt[k + 2], fmaf(mv.w, qt[k + 3], acc))));
          }
          sc = acc * scale;
        }
        dec[L::dSC + (s * H + h) * SLOT + j] = sc;
      }
      __syncthreads();
      // d: softmax per (sample, head): one warp each
      if (warp < 2 * H) {
        float* sc = dec + L::dSC + warp * SLOT;
        float mx = -INFINITY;
        for (int j = lane; j < SLOT; j += 32) mx = fmaxf(mx, sc[j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < SLOT; j += 32) {
          const float e = (sc[j] == -INFINITY) ? 0.f : __expf(sc[j] - mx);
          sc[j] = e;
          sum += e;
        }
        sum = warp_sum(sum);
        const float inv = sum > 0.f ? 1.0f / sum : 0.f;
        for (int j = lane; j < SLOT; j += 32) sc[j] *= inv;
      }
      __syncthreads();
      // e: ctx[s][h][k] = sum_j p_j M_j[k]
      for (int i = tid; i < 2 * H * D; i += kTcThreads) {
        const int k = i % D, h = (i / D) % H, s = i / (D * H);
        const float* pr = dec + L::dSC + (s * H + h) * SLOT;
        const float* m = Mem + ((g0 + s) * SLOT) * MLD + k;
        // p_j is exactly 0 beyond the sequence length and every memory row is finite, so the loop runs
        // over the whole slot: fixed trip count, four independent accumulators
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int j = 0; j < SLOT; j += 4) {
          a0 = fmaf(pr[j], m[j * MLD], a0);
          a1 = fmaf(pr[j + 1], m[(j + 1) * MLD], a1);
          a2 = fmaf(pr[j + 2], m[(j + 2) * MLD], a2);
          a3 = fmaf(pr[j + 3], m[(j + 3) * MLD], a3);
        }
        dec[L::dCTX + (s * H + h) * D + k] = (a0 + a1) + (a2 + a3);
      }
      __syncthreads();
      // f: o = ctx_h Wv_h + bv (sum_j p_j == 1; an empty sequence contributes o = 0) ; y = o + dvec
      if (tid < 2 * D) {
        const int s = tid / D, c = tid % D, h = c / DK;
        const float* ctx = dec + L::dCTX + (s * H + h) * D;
        float acc = slen[g0 + s] > 0 ? fv[L::vDB + 2 * D + c] : 0.f;
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
          float w[8];
          bf16x8_to_float(*reinterpret_cast<const uint4*>(sDv + (kc * D + c) * 16), w);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc = fmaf(ctx[kc * 8 + e], w[e], acc);
        }
        dec[L::dOV + s * D + c] = acc + fv[L::vDVEC + (g0 + s) * D + c];
      }
      __syncthreads();
      // g: av = LN3(y) (vanilla-attention LayerNorm): one warp per sample
      if (warp < 2) {
        const float* y = dec + L::dOV + warp * D;
        float v[(D + 31) / 32], sm = 0.f;
#pragma unroll
        for (int i = 0; i < (D + 31) / 32; ++i) {
          const int c = lane + 32 * i;
          v[i] = c < D ? y[c] : 0.f;
          sm += v[i];
        }
        const float mean = warp_sum(sm) * (1.0f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < (D + 31) / 32; ++i)
          if (lane + 32 * i < D) q += (v[i] - mean) * (v[i] - mean);
        const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + kLnEps);
#pragma unroll
        for (int i = 0; i < (D + 31) / 32; ++i) {
          const int c = lane + 32 * i;
          if (c < D) dec[L::dAV + warp * D + c] = fv[L::vLN + 4 * D + c] * ((v[i] - mean) * rstd) + fv[L::vLN + 5 * D + c];
        }
      }
      __syncthreads();
      // h: hv = relu(av W1 + b1): W1 image in shared memory, thread n serves both samples
      for (int n = tid; n < DFF; n += kTcThreads) {
        float a0 = fv[L::vB1 + n], a1 = a0;
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
          float w[8];
          bf16x8_to_float(*reinterpret_cast<const uint4*>(smem + L::oW1 + (kc * DFF + n) * 16), w);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            a0 = fmaf(dec[L::dAV + kc * 8 + e], w[e], a0);
            a1 = fmaf(dec[L::dAV + D + kc * 8 + e], w[e], a1);
          }
        }
        dec[L::dHV + n] = fmaxf(a0, 0.f);
        dec[L::dHV + DFF + n] = fmaxf(a1, 0.f);
      }
      __syncthreads();
      // i: f = hv W2 (+ b2 + av), reduction over DFF split in two parts across the 256 threads
      {
        const int c = tid % D, s = (tid / D) % 2, part = tid / (2 * D);
        if (part < 2) {
          float acc = 0.f;
          const float* hv = dec + L::dHV + s * DFF;
          for (int kc = part * (DFF / 16); kc < (part + 1) * (DFF / 16); ++kc) {
            float w[8];
            bf16x8_to_float(*reinterpret_cast<const uint4*>(smem + L::oW2 + (kc * D + c) * 16), w);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc = fmaf(hv[kc * 8 + e], w[e], acc);
          }
          dec[L::dPART + (part * 2 + s) * D + c] = acc;
        }
      }
      __syncthreads();
      // j: u = LN2(f + b2 + av) (shared feed-forward LayerNorm) -> global
      if (warp < 2) {
        const int s = warp, b = b0 + g0 + s;
        float v[(D + 31) / 32], sm = 0.f;
#pragma unroll
        for (int i = 0; i < (D + 31) / 32; ++i) {
          const int c = lane + 32 * i;
          v[i] = c < D ? dec[L::dPART + s * D + c] + dec[L::dPART + (2 + s) * D + c] + fv[L::vB2 + c] +
                             dec[L::dAV + s * D + c]
                       : 0.f;
          sm += v[i];
        }
        const float mean = warp_sum(sm) * (1.0f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < (D + 31) / 32; ++i)
          if (lane + 32 * i < D) q += (v[i] - mean) * (v[i] - mean);
        const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + kLnEps);
        if (b < B) {
#pragma unroll
          for (int i = 0; i < (D + 31) / 32; ++i) {
            const int c = lane + 32 * i;
            if (c < D)
              a.out[(int64_t)b * a.out_ld + c] = fv[L::vLN + 2 * D + c] * ((v[i] - mean) * rstd) + fv[L::vLN + 3 * D + c];
          }
        }
      }
      __syncthreads();
    }
    DMT_TICK(12);
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

template <int D, int DFF, int H, int SLOT>
static int launch_tc(const SeqTcArgs& a, cudaStream_t st) {
  using L = TcLayout<D, DFF, H, SLOT>;
  static_assert(L::total <= 227 * 1024, "shared-memory plan exceeds 227 KB");
  auto kern = seq_encode_tc_kernel<D, DFF, H, SLOT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(seq_encode_tc_kernel)");
  const int sms = sm_count_cached();
  const int grid = a.n_tiles < sms ? a.n_tiles : sms;
  kern<<<grid, kTcThreads, L::total, st>>>(a);
  DMT_CUDA_LAUNCH_CHECK("seq_encode_tc_kernel");
  return DMT_OK;
}

static unsigned long long* g_seq_profile = nullptr;   // diagnostics only (dmt_debug_seq_profile)
void seq_tc_set_profile(unsigned long long* p) { g_seq_profile = p; }

size_t seq_tc_prepared_bytes(const dmt_seq_cfg* cfg) {
  return (prep_total(cfg->d_model, cfg->d_ff, cfg->num_heads) * 2 + 511) / 256 * 256;
}

// v2 (seq_encode_tc2.cu)
bool seq_tc2_supported(const dmt_seq_cfg* cfg);
size_t seq_tc2_ctx_bytes(const dmt_seq_cfg* cfg);
int seq_encode_tc2_launch(SeqTcArgs& a, cudaStream_t st);

// DMT_SEQ_TC_V1=1 selects the v1 kernel (one tile in flight per SM) -- kept for A/B measurements
static bool use_v1() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DMT_SEQ_TC_V1");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// workspace = [prepared weight images | v2: decoder contexts [B][H*D] fp32]
size_t seq_tc_workspace_bytes(const dmt_seq_cfg* cfg) {
  return seq_tc_prepared_bytes(cfg) + seq_tc2_ctx_bytes(cfg);
}

bool seq_tc_supported(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const char** why) {
  *why = nullptr;
  if (!(cfg->d_model == 64 && cfg->d_ff == 256 && cfg->num_heads == 2))
    *why = "bf16 tensor-core path is built for d_model=64, d_ff=256, 2 heads";
  else if (cfg->n_enc_blocks != 1 || cfg->n_dec_blocks != 1)
    *why = "bf16 tensor-core path is built for 1 encoder + 1 decoder block";
  else
    for (int f = 0; f < cfg->n_feats; ++f)
      if (in->dim[f] % 8) *why = "bf16 tensor-core path needs pair dims that are multiples of 8";
  return *why == nullptr;
}

int seq_tc_prepare(const dmt_seq_cfg* cfg, const dmt_seq_weights* w, void* prepared, cudaStream_t st) {
  const size_t total = prep_total(cfg->d_model, cfg->d_ff, cfg->num_heads);
  const int blocks = (int)((total + 255) / 256);
  const dmt_attn_weights& e = w->enc_attn[0];
  const dmt_attn_weights& d = w->dec_attn[0];
  seq_prepare_kernel<<<blocks, 256, 0, st>>>(e.q.w, e.k.w, e.v.w, w->ff[0].w1.w, w->ff[0].w2.w, d.q.w, d.q.b, d.k.w,
                                            d.v.w, (__nv_bfloat16*)prepared, cfg->d_model, cfg->d_ff,
                                            cfg->num_heads);
  DMT_CUDA_LAUNCH_CHECK("seq_prepare_kernel");
  return DMT_OK;
}

static void fill_args(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, float* out,
                      int64_t out_ld, void* workspace, SeqTcArgs& a) {
  const void* prepared = workspace;
  a.cfg = *cfg;
  a.in = *in;
  a.pos = w->pos;
  const dmt_attn_weights& e = w->enc_attn[0];
  const dmt_attn_weights& d = w->dec_attn[0];
  a.bq = e.q.b; a.bk = e.k.b; a.bv = e.v.b; a.ln1_g = e.ln.gamma; a.ln1_b = e.ln.beta;
  a.b1 = w->ff[0].w1.b; a.b2 = w->ff[0].w2.b; a.ln2_g = w->ff[0].ln.gamma; a.ln2_b = w->ff[0].ln.beta;
  a.dbq = d.q.b; a.dbk = d.k.b; a.dbv = d.v.b; a.ln3_g = d.ln.gamma; a.ln3_b = d.ln.beta;
  a.prepared = (const __nv_bfloat16*)prepared;
  a.dbg = g_seq_profile;
  a.out = out;
  a.out_ld = out_ld;
  a.n_tiles = 0;
  int c = 0;
  for (int f = 0; f < cfg->n_feats; ++f)
    for (int o = 0; o < in->dim[f]; o += 8, ++c) {
      a.chunk_feat[c] = f;
      a.chunk_off[c] = o;
    }
  for (; c < 32; ++c) a.chunk_feat[c] = a.chunk_off[c] = 0;
  a.ctx = static_cast<uint8_t*>(workspace) + seq_tc_prepared_bytes(cfg);
}

int seq_tails_launch(int n, const SeqTcArgs* args, cudaStream_t st);

int seq_tc_tails(int n, const dmt_seq_cfg* const* cfgs, const dmt_seq_input* const* ins, const dmt_seq_weights* const* ws,
                 float* const* outs, const int64_t* out_lds, void* const* workspaces, cudaStream_t st) {
  SeqTcArgs args[DMT_MAX_TAIL_SEQS];
  for (int i = 0; i < n; ++i) {
    DMT_REQUIRE(!use_v1() && seq_tc2_supported(cfgs[i]), DMT_ERR_UNSUPPORTED_SHAPE,
                "dmt_seq_tail_fwd: sequence %d was not run by the deferred-tail kernel", i);
    fill_args(cfgs[i], ins[i], ws[i], outs[i], out_lds[i], workspaces[i], args[i]);
  }
  return seq_tails_launch(n, args, st);
}

int seq_encode_tc_launch(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, float* out,
                         int64_t out_ld, void* workspace, cudaStream_t st) {
  SeqTcArgs a;
  fill_args(cfg, in, w, out, out_ld, workspace, a);
  if (!use_v1() && seq_tc2_supported(cfg)) return seq_encode_tc2_launch(a, st);
  DMT_REQUIRE(!(cfg->flags & DMT_SEQ_DEFER_TAIL), DMT_ERR_UNSUPPORTED_SHAPE,
              "dmt_seq_encode_fwd: DMT_SEQ_DEFER_TAIL needs the v2 kernel (maxlen <= 55)");
  int slot = cfg->slot_len > 0 ? cfg->slot_len : cfg->maxlen;
  if (slot > cfg->maxlen) slot = cfg->maxlen;
  DMT_REQUIRE(slot <= 64, DMT_ERR_UNSUPPORTED_SHAPE, "dmt_seq_encode_fwd(bf16): sequences longer than 64 (%d)", slot);
  if (slot <= 16) {
    a.n_tiles = (cfg->batch + 7) / 8;
    return launch_tc<64, 256, 2, 16>(a, st);
  }
  if (slot <= 32) {
    a.n_tiles = (cfg->batch + 3) / 4;
    return launch_tc<64, 256, 2, 32>(a, st);
  }
  a.n_tiles = (cfg->batch + 1) / 2;
  return launch_tc<64, 256, 2, 64>(a, st);
}

}  // namespace dmt
