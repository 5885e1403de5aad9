// dmt_seq_encode_fwd_train on the tensor-core GEMM engine: the training forward of one behaviour sequence as a
// row-batched pipeline whose outputs ARE the activations the backward needs (seq_train.cuh), so every byte
// written to HBM is used again:
//
//   gather+concat+sqrt(d)+pos            -> hin[0] (tokens), din[0] (targets)      HBM-bound row gather
//   per encoder block:  Q|K|V = h W + b  -> qkv                                     grouped GEMM (3 problems)
//                       self-attention + residual + LayerNorm (one CTA per sample)  -> z1, a
//                       relu(a W1 + b1) -> f1 ;  f1 W2 + b2 + a -> z2               grouped GEMMs
//                       LayerNorm rows   -> hin[blk+1]
//   per decoder block:  K|V = memory W + b -> kvd ; q = d W + b -> qd               grouped GEMM (3 problems)
//                       single-query attention + residual + LayerNorm (CTA/sample)  -> pd, z1d, ad
//                       FF as above on B rows -> f1d, z2d ; LayerNorm -> din[blk+1]
//
// Same arithmetic as TransformerModel.py:84-171 / TransformerModel_util.py:11-108,160-235 (see
// seq_encode_f32.cu for the fused fp32 statement); only the GEMM operands are rounded (bf16 or split-bf16).
#include "dropout.cuh"
#include "gemm_f32.cuh"
#include "gemm_tf32.cuh"
#include "seq_train.cuh"

namespace dmt {

bool attn_fwd_tc_supported(const dmt_seq_cfg& c);
int attn_fwd_tc_launch(const dmt_seq_cfg& c, const float* qkv, const float* h, const float* gamma, const float* beta,
                       float* z1, float* a, const int32_t* offsets, int64_t T, int LP, const Dropout& drop,
                       cudaStream_t st);

namespace {

struct GatherArgs {
  dmt_seq_cfg cfg;
  dmt_seq_input in;
  const float* pos;
  int32_t col_off[DMT_MAX_SEQ_FEATS + 1];
  float* h0;     // [T, D]
  float* d0;     // [B, D]
  int LP;
};

__device__ __forceinline__ float lookup(const float* __restrict__ table, int64_t rows, int dim, int id, int c, int zp) {
  const int64_t row = (int64_t)id - (zp ? 1 : 0);
  if (row < 0 || row >= rows) return 0.f;
  return __ldg(table + row * dim + c);
}

// One CTA per sample (mmoe_transformer_unbias.py:153-158,181; TransformerModel.py:97-100,147).
__global__ void __launch_bounds__(256) seq_gather_kernel(const __grid_constant__ GatherArgs a) {
  const int D = a.cfg.d_model, nf = a.cfg.n_feats, b = blockIdx.x;
  const int off_last = __ldg(a.in.offsets[nf - 1] + b);
  const int len_all = __ldg(a.in.offsets[nf - 1] + b + 1) - off_last;
  const int L = min(len_all, a.LP);
  const float sqrt_d = sqrtf((float)D);
  const Dropout drop_enc(a.cfg.dropout_rate, a.cfg.dropout_seed, kSiteEncIn);
  const Dropout drop_dec(a.cfg.dropout_rate, a.cfg.dropout_seed, kSiteDecIn);
  for (int i = threadIdx.x; i < len_all * D; i += 256) {
    const int t = i / D, c = i - t * D;
    float v = 0.f;
    if (t < L) {
      int f = 0;
      while (f + 1 < nf && c >= a.col_off[f + 1]) ++f;
      const int off = __ldg(a.in.offsets[f] + b);
      const int len_f = __ldg(a.in.offsets[f] + b + 1) - off;
      const int id = (t < len_f) ? __ldg(a.in.ids[f] + off + t) : 0;
      v = lookup(a.in.table[f], a.in.rows[f], a.in.dim[f], id, c - a.col_off[f], a.cfg.zero_pad) * sqrt_d +
          __ldg(a.pos + t * D + c);
      v *= drop_enc.mult((uint32_t)((off_last + t) * D + c));      // TransformerModel.py:101
    }
    a.h0[((int64_t)off_last + t) * D + c] = v;
  }
  for (int c = threadIdx.x; c < D; c += 256) {
    int f = 0;
    while (f + 1 < nf && c >= a.col_off[f + 1]) ++f;
    const int id = __ldg(a.in.item_ids[f] + b);
    a.d0[(int64_t)b * D + c] =
        lookup(a.in.table[f], a.in.rows[f], a.in.dim[f], id, c - a.col_off[f], a.cfg.zero_pad) * sqrt_d *
        drop_dec.mult((uint32_t)(b * D + c));                        // TransformerModel.py:151
  }
}

__device__ __forceinline__ void ln_row_warp(const float* v_in, int D, const float* __restrict__ gamma,
                                            const float* __restrict__ beta, float* out, int lane) {
  // v_in / out: D values of one row, strided by lane (values lane, lane+32, ...); D <= 256
  float v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < D ? v_in[c] : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (lane + 32 * i < D) {
      const float dl = v[i] - mean;
      q += dl * dl;
    }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)D + kLnEps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    if (c < D) out[c] = __ldg(gamma + c) * ((v[i] - mean) * rstd) + __ldg(beta + c);
  }
}

// rows of z -> LayerNorm -> out (one warp per row; `out2` optionally receives a strided copy)
__global__ void __launch_bounds__(256) ln_fwd_rows_kernel(const float* __restrict__ z, int64_t rows, int D,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ out,
                                                          float* __restrict__ out2, int64_t ld2) {
  __shared__ float buf[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < rows; r += (int64_t)gridDim.x * 8) {
    for (int c = lane; c < D; c += 32) buf[warp][c] = __ldg(z + r * D + c);
    __syncwarp();
    ln_row_warp(buf[warp], D, gamma, beta, buf[warp], lane);
    __syncwarp();
    for (int c = lane; c < D; c += 32) {
      out[r * D + c] = buf[warp][c];
      if (out2) out2[r * ld2 + c] = buf[warp][c];
    }
    __syncwarp();
  }
}

struct AttnFwdArgs {
  const float* qkv;   // [T, 3D]
  const float* h;     // [T, D] block input (residual)
  const float* gamma;
  const float* beta;
  float* z1;          // [T, D]
  float* a;           // [T, D]
  const int32_t* offsets;
  int D, H, LP;
  Dropout drop;       // attention-probability dropout of this block (TransformerModel_util.py:51)
};

// One CTA per sample: softmax(Q_h K_h^T / sqrt(dk)) V_h over the L valid keys, + residual, LayerNorm.
__global__ void __launch_bounds__(256) attn_fwd_kernel(const __grid_constant__ AttnFwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int D = a.D, H = a.H, LP = a.LP, dk = D / H, ld = D + 1, lds = LP + 1;
  float* Q = sm;
  float* K = Q + LP * ld;
  float* V = K + LP * ld;
  float* S = V + LP * ld;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t off = __ldg(a.offsets + b);
  const int L = min(__ldg(a.offsets + b + 1) - (int)off, LP);
  if (L == 0) return;
  for (int i = tid; i < L * D; i += 256) {
    const int t = i / D, c = i - t * D;
    const float* row = a.qkv + (off + t) * 3 * D;
    Q[t * ld + c] = __ldg(row + c);
    K[t * ld + c] = __ldg(row + D + c);
    V[t * ld + c] = __ldg(row + 2 * D + c);
  }
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)dk);
  for (int h = 0; h < H; ++h) {
    const int hc = h * dk;
    // 2 x 2 register tile per thread (rows qa, qa+hl; keys kb, kb+hl): one shared-memory load per FMA
    const int hl = (L + 1) >> 1;
    for (int i = tid; i < hl * hl; i += 256) {
      const int qa = i / hl, kb = i - qa * hl;
      const int qb = min(qa + hl, L - 1), kc = min(kb + hl, L - 1);
      const float *q0 = Q + qa * ld + hc, *q1 = Q + qb * ld + hc, *k0 = K + kb * ld + hc, *k1 = K + kc * ld + hc;
      float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
      for (int c = 0; c < dk; ++c) {
        const float a0 = q0[c], a1 = q1[c], b0 = k0[c], b1 = k1[c];
        s00 = fmaf(a0, b0, s00); s01 = fmaf(a0, b1, s01); s10 = fmaf(a1, b0, s10); s11 = fmaf(a1, b1, s11);
      }
      S[qa * lds + kb] = s00 * scale;
      if (kb + hl < L) S[qa * lds + kb + hl] = s01 * scale;
      if (qa + hl < L) {
        S[(qa + hl) * lds + kb] = s10 * scale;
        if (kb + hl < L) S[(qa + hl) * lds + kb + hl] = s11 * scale;
      }
    }
    __syncthreads();
    for (int r = warp; r < L; r += 8) {
      float m = -INFINITY;
      for (int j = lane; j < L; j += 32) m = fmaxf(m, S[r * lds + j]);
      m = warp_max(m);
      float s = 0.f;
      for (int j = lane; j < L; j += 32) {
        const float e = expf(S[r * lds + j] - m);
        S[r * lds + j] = e;
        s += e;
      }
      const float inv = 1.0f / warp_sum(s);
      for (int j = lane; j < L; j += 32)
        S[r * lds + j] *= inv * a.drop.mult((uint32_t)(((b * H + h) * LP + r) * LP + j));
    }
    __syncthreads();
    for (int i = tid; i < hl * dk; i += 256) {   // two query rows per thread share every V load
      const int qi = i / dk, c = i - qi * dk;
      const int q2 = min(qi + hl, L - 1);
      float acc = 0.f, acc2 = 0.f;
      for (int j = 0; j < L; ++j) {
        const float vx = V[j * ld + hc + c];
        acc = fmaf(S[qi * lds + j], vx, acc);
        acc2 = fmaf(S[q2 * lds + j], vx, acc2);
      }
      Q[qi * ld + hc + c] = acc;     // this head's Q columns are dead after the scores
      if (qi + hl < L) Q[(qi + hl) * ld + hc + c] = acc2;
    }
    __syncthreads();
  }
  for (int r = warp; r < L; r += 8) {
    float* row = Q + r * ld;
    for (int c = lane; c < D; c += 32) {
      const float z = row[c] + __ldg(a.h + (off + r) * D + c);
      row[c] = z;
      a.z1[(off + r) * D + c] = z;
    }
    __syncwarp();
    ln_row_warp(row, D, a.gamma, a.beta, row, lane);
    __syncwarp();
    for (int c = lane; c < D; c += 32) a.a[(off + r) * D + c] = row[c];
  }
}

struct DecAttnFwdArgs {
  const float* qd;    // [B, D]
  const float* kvd;   // [T, 2D]
  const float* din;   // [B, D] residual
  const float* gamma;
  const float* beta;
  float* pd;          // [B, H, LP]
  float* z1d;         // [B, D]
  float* ad;          // [B, D]
  const int32_t* offsets;
  int D, H, LP;
  Dropout drop;
};

__global__ void __launch_bounds__(128) dec_attn_fwd_kernel(const __grid_constant__ DecAttnFwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int D = a.D, H = a.H, LP = a.LP, dk = D / H, ld = 2 * D + 1;
  float* KV = sm;
  float* q = KV + LP * ld;
  float* p = q + D;          // [H][LP]
  float* o = p + H * LP;     // [D] (padded to 256 for the LayerNorm helper)
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t off = __ldg(a.offsets + b);
  const int L = min(__ldg(a.offsets + b + 1) - (int)off, LP);
  for (int i = tid; i < L * 2 * D; i += 128) {
    const int t = i / (2 * D), c = i - t * 2 * D;
    KV[t * ld + c] = __ldg(a.kvd + (off + t) * 2 * D + c);
  }
  for (int c = tid; c < D; c += 128) q[c] = __ldg(a.qd + (int64_t)b * D + c);
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)dk);
  for (int i = tid; i < H * L; i += 128) {
    const int h = i / L, j = i - h * L;
    float s = 0.f;
    for (int c = 0; c < dk; ++c) s = fmaf(q[h * dk + c], KV[j * ld + h * dk + c], s);
    p[h * LP + j] = s * scale;
  }
  __syncthreads();
  for (int h = warp; h < H; h += 4) {
    float m = -INFINITY;
    for (int j = lane; j < L; j += 32) m = fmaxf(m, p[h * LP + j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float e = expf(p[h * LP + j] - m);
      p[h * LP + j] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = s > 0.f ? 1.0f / s : 0.f;
    for (int j = lane; j < LP; j += 32) p[h * LP + j] = j < L ? p[h * LP + j] * inv : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < H * LP; i += 128) a.pd[(int64_t)b * H * LP + i] = p[i];
  for (int c = tid; c < D; c += 128) {
    const int h = c / dk;
    float acc = 0.f;
    for (int j = 0; j < L; ++j)
      acc = fmaf(p[h * LP + j] * a.drop.mult((uint32_t)((b * H + h) * LP + j)), KV[j * ld + D + c], acc);
    const float z = acc + __ldg(a.din + (int64_t)b * D + c);
    o[c] = z;
    a.z1d[(int64_t)b * D + c] = z;
  }
  __syncthreads();
  if (warp == 0) {
    ln_row_warp(o, D, a.gamma, a.beta, o, lane);
    __syncwarp();
    for (int c = lane; c < D; c += 32) a.ad[(int64_t)b * D + c] = o[c];
  }
}

struct ZeroCapArgs {
  SeqSaved sv;
  const int32_t* offsets;
  int D, DFF, LP, n_enc, n_dec;
};

// Tokens beyond the on-chip cap of the per-sample kernels (never in the reference's data: maxlen_k bounds the
// sequences) must be inert in every contraction of the backward: zero their rows in all saved buffers.
__global__ void __launch_bounds__(256) zero_capped_rows_kernel(const __grid_constant__ ZeroCapArgs a) {
  const int b = blockIdx.x;
  const int64_t off = __ldg(a.offsets + b);
  const int len_all = __ldg(a.offsets + b + 1) - (int)off;
  if (len_all <= a.LP) return;
  const int n = len_all - a.LP;
  const int64_t r0 = off + a.LP;
  auto zero = [&](float* p, int W) {
    for (int i = threadIdx.x; i < n * W; i += 256) p[r0 * W + i] = 0.f;
  };
  for (int k = 0; k <= a.n_enc; ++k) zero(a.sv.hin[k], a.D);
  for (int k = 0; k < a.n_enc; ++k) {
    zero(a.sv.qkv[k], 3 * a.D);
    zero(a.sv.z1[k], a.D);
    zero(a.sv.a[k], a.D);
    zero(a.sv.f1[k], a.DFF);
    zero(a.sv.z2[k], a.D);
  }
  for (int k = 0; k < a.n_dec; ++k) zero(a.sv.kvd[k], 2 * a.D);
}

// C[rows, N] = act(A[rows, K] W + b (+ addend)), W in the TF layout [K, N]
inline void fwd_prob(GemmProb& p, const float* A, int64_t lda, int K, const dmt_dense& w, int N, int64_t rows, float* C,
                     int64_t ldc, bool relu, const float* addend, int64_t ld_add) {
  gemm_prob_init(p);
  p.n_parts = 1;
  p.part[0] = GemmPart{A, w.w, lda, (int64_t)N, K, 0};
  p.M = (int)rows;
  p.N = N;
  p.C = C;
  p.ldc = ldc;
  p.bias = w.b;
  p.relu = relu ? 1 : 0;
  p.addend = addend;
  p.ld_add = ld_add;
}

// DMT_PRECISION_TF32: C[rows, N] = act(A[rows, K] W + b (+ addend)) on the TMA-fed tf32 engine.  `Bt` = the packed
// K-major operand [N, K] (row stride K), `bias` its packed bias.
inline int tf32_dense(const float* A, int64_t lda, int K, const float* Bt, const float* bias, int N, int64_t rows,
                      float* C, int64_t ldc, bool relu, const float* addend, int64_t ld_add, cudaStream_t st) {
  Tf32Rows p{};
  p.A = A; p.lda = lda; p.Bt = Bt; p.ldb = K; p.M = rows; p.N = N; p.K = K; p.C = C; p.ldc = ldc;
  p.bias = bias; p.addend = addend; p.ld_add = ld_add; p.alpha = 1.0f; p.relu = relu ? 1 : 0;
  return tf32_rows(p, st);
}

// Bt = [W_0^T ; W_1^T ; ...] (each TF kernel [K, n_i] -> rows of the K-major operand), bias = [b_0 | b_1 | ...]
inline int tf32_pack_dense(const dmt_dense* const* w, const int* n_out, int n, int K, float* Bt, float* bias,
                           cudaStream_t st) {
  Tf32PackMat m[4];
  Tf32PackVec v[4];
  int row = 0;
  for (int i = 0; i < n; ++i) {
    m[i] = Tf32PackMat{w[i]->w, (int64_t)n_out[i], n_out[i], K, 1, row, 0};
    v[i] = Tf32PackVec{w[i]->b, n_out[i], row};
    row += n_out[i];
  }
  return tf32_pack(m, n, Bt, K, v, n, bias, st);
}

}  // namespace

int seq_fwd_train_pipeline(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, float* out,
                           int64_t out_ld, int64_t T, const SeqSaved& sv, cudaStream_t st) {
  const dmt_seq_cfg& c = *cfg;
  const int d = c.d_model, dff = c.d_ff, B = c.batch, H = c.num_heads, LP = seq_lp(c);
  const int use_tc = gemm_engine(c.precision);
  const bool tf = gemm_tf32(c.precision);          // per-token GEMMs on the TMA-fed tf32 engine
  // weight-pack scratch of the tf32 route: [Q|K|V (or K|V) operand + bias | W1^T + b1 | W2^T + b2]
  float* pk_qkv = sv.pack;
  float* pk_qkv_b = pk_qkv + (size_t)3 * d * d;
  float* pk_w1 = pk_qkv_b + 4 * d;
  float* pk_w1_b = pk_w1 + (size_t)d * dff;
  float* pk_w2 = pk_w1_b + dff;
  float* pk_w2_b = pk_w2 + (size_t)d * dff;
  float* pk_q = pk_w2_b + 4 * d;                   // decoder query projection [d, d] + bias
  float* pk_q_b = pk_q + (size_t)d * d;
  if (tf)
    DMT_REQUIRE(d % 16 == 0 && dff % 16 == 0, DMT_ERR_UNSUPPORTED_SHAPE,
                "DMT_PRECISION_TF32 needs d_model and d_ff that are multiples of 16 (got %d, %d)", d, dff);
  const int32_t* offsets = in->offsets[c.n_feats - 1];
  const int ln_grid = 4 * sm_count_cached();
  int rc;
  {
    GatherArgs a{};
    a.cfg = c;
    a.in = *in;
    a.pos = w->pos;
    int col = 0;
    for (int f = 0; f < c.n_feats; ++f) {
      a.col_off[f] = col;
      col += in->dim[f];
    }
    for (int f = c.n_feats; f <= DMT_MAX_SEQ_FEATS; ++f) a.col_off[f] = col;
    DMT_REQUIRE(col == d, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_encode_fwd_train: pair dims sum to %d, d_model is %d", col, d);
    a.h0 = sv.hin[0];
    a.d0 = sv.din[0];
    a.LP = LP;
    seq_gather_kernel<<<B, 256, 0, st>>>(a);
    DMT_CUDA_LAUNCH_CHECK("seq_gather_kernel");
  }
  for (int blk = 0; blk < c.n_enc_blocks && T > 0; ++blk) {
    const dmt_attn_weights& aw = w->enc_attn[blk];
    const dmt_ff_weights& fw = w->ff[blk];
    if (tf) {   // one GEMM for Q | K | V: hin is read once
      const dmt_dense* ws3[3] = {&aw.q, &aw.k, &aw.v};
      const int n3[3] = {d, d, d};
      if ((rc = tf32_pack_dense(ws3, n3, 3, d, pk_qkv, pk_qkv_b, st))) return rc;
      if ((rc = tf32_dense(sv.hin[blk], d, d, pk_qkv, pk_qkv_b, 3 * d, T, sv.qkv[blk], 3 * d, false, nullptr, 0, st)))
        return rc;
    } else {
      GemmGroup grp{};
      grp.use_tc = use_tc;
      fwd_prob(grp.p[0], sv.hin[blk], d, d, aw.q, d, T, sv.qkv[blk], 3 * d, false, nullptr, 0);
      fwd_prob(grp.p[1], sv.hin[blk], d, d, aw.k, d, T, sv.qkv[blk] + d, 3 * d, false, nullptr, 0);
      fwd_prob(grp.p[2], sv.hin[blk], d, d, aw.v, d, T, sv.qkv[blk] + 2 * d, 3 * d, false, nullptr, 0);
      grp.n = 3;
      if ((rc = gemm_group_launch(grp, st))) return rc;
    }
    if (tf && attn_fwd_tc_supported(c)) {   // tcgen05 attention: whole-sample tiles cut from the packed token rows
      if ((rc = attn_fwd_tc_launch(c, sv.qkv[blk], sv.hin[blk], aw.ln.gamma, aw.ln.beta, sv.z1[blk], sv.a[blk], offsets,
                                   T, LP, Dropout(c.dropout_rate, c.dropout_seed, kSiteSelfProbs + blk), st)))
        return rc;
    } else {
      AttnFwdArgs a{sv.qkv[blk], sv.hin[blk], aw.ln.gamma, aw.ln.beta, sv.z1[blk], sv.a[blk], offsets, d, H, LP,
                    Dropout(c.dropout_rate, c.dropout_seed, kSiteSelfProbs + blk)};
      const size_t smem = ((size_t)3 * LP * (d + 1) + LP * (LP + 1)) * sizeof(float);
      DMT_REQUIRE(smem <= 227 * 1024, DMT_ERR_UNSUPPORTED_SHAPE, "dmt_seq_encode_fwd_train: attention tile %zu B", smem);
      cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn_fwd_kernel)");
      attn_fwd_kernel<<<B, 256, smem, st>>>(a);
      DMT_CUDA_LAUNCH_CHECK("attn_fwd_kernel");
    }
    if (tf) {
      const dmt_dense* w1p[1] = {&fw.w1};
      const dmt_dense* w2p[1] = {&fw.w2};
      const int n1[1] = {dff}, n2[1] = {d};
      if ((rc = tf32_pack_dense(w1p, n1, 1, d, pk_w1, pk_w1_b, st))) return rc;
      if ((rc = tf32_pack_dense(w2p, n2, 1, dff, pk_w2, pk_w2_b, st))) return rc;
      if ((rc = tf32_dense(sv.a[blk], d, d, pk_w1, pk_w1_b, dff, T, sv.f1[blk], dff, true, nullptr, 0, st))) return rc;
      if ((rc = tf32_dense(sv.f1[blk], dff, dff, pk_w2, pk_w2_b, d, T, sv.z2[blk], d, false, sv.a[blk], d, st)))
        return rc;
    } else {
      {
        GemmGroup grp{};
        grp.use_tc = use_tc;
        fwd_prob(grp.p[0], sv.a[blk], d, d, fw.w1, dff, T, sv.f1[blk], dff, true, nullptr, 0);
        grp.n = 1;
        if ((rc = gemm_group_launch(grp, st))) return rc;
      }
      {
        GemmGroup grp{};
        grp.use_tc = use_tc;
        fwd_prob(grp.p[0], sv.f1[blk], dff, dff, fw.w2, d, T, sv.z2[blk], d, false, sv.a[blk], d);
        grp.n = 1;
        if ((rc = gemm_group_launch(grp, st))) return rc;
      }
    }
    ln_fwd_rows_kernel<<<ln_grid, 256, 0, st>>>(sv.z2[blk], T, d, fw.ln.gamma, fw.ln.beta, sv.hin[blk + 1], nullptr, 0);
    DMT_CUDA_LAUNCH_CHECK("ln_fwd_rows_kernel");
  }
  const float* mem = sv.hin[c.n_enc_blocks];
  for (int blk = 0; blk < c.n_dec_blocks; ++blk) {
    const dmt_attn_weights& aw = w->dec_attn[blk];
    const dmt_ff_weights& fw = w->ff[blk];
    if (tf) {
      const dmt_dense* wq[1] = {&aw.q};
      const dmt_dense* wkv[2] = {&aw.k, &aw.v};
      const int nq[1] = {d}, nkv[2] = {d, d};
      if ((rc = tf32_pack_dense(wq, nq, 1, d, pk_q, pk_q_b, st))) return rc;
      if ((rc = tf32_dense(sv.din[blk], d, d, pk_q, pk_q_b, d, B, sv.qd[blk], d, false, nullptr, 0, st))) return rc;
      if (T > 0) {
        if ((rc = tf32_pack_dense(wkv, nkv, 2, d, pk_qkv, pk_qkv_b, st))) return rc;
        if ((rc = tf32_dense(mem, d, d, pk_qkv, pk_qkv_b, 2 * d, T, sv.kvd[blk], 2 * d, false, nullptr, 0, st)))
          return rc;
      }
    } else {
      GemmGroup grp{};
      grp.use_tc = use_tc;
      int n = 0;
      fwd_prob(grp.p[n++], sv.din[blk], d, d, aw.q, d, B, sv.qd[blk], d, false, nullptr, 0);
      if (T > 0) {
        fwd_prob(grp.p[n++], mem, d, d, aw.k, d, T, sv.kvd[blk], 2 * d, false, nullptr, 0);
        fwd_prob(grp.p[n++], mem, d, d, aw.v, d, T, sv.kvd[blk] + d, 2 * d, false, nullptr, 0);
      }
      grp.n = n;
      if ((rc = gemm_group_launch(grp, st))) return rc;
    }
    {
      DecAttnFwdArgs a{sv.qd[blk], sv.kvd[blk], sv.din[blk], aw.ln.gamma, aw.ln.beta, sv.pd[blk], sv.z1d[blk], sv.ad[blk],
                       offsets, d, H, LP, Dropout(c.dropout_rate, c.dropout_seed, kSiteVanillaProbs + blk)};
      const size_t smem = ((size_t)LP * (2 * d + 1) + d + H * LP + 256 + 8) * sizeof(float);
      cudaError_t e = cudaFuncSetAttribute(dec_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dec_attn_fwd_kernel)");
      dec_attn_fwd_kernel<<<B, 128, smem, st>>>(a);
      DMT_CUDA_LAUNCH_CHECK("dec_attn_fwd_kernel");
    }
    if (tf) {
      const dmt_dense* w1p[1] = {&fw.w1};
      const dmt_dense* w2p[1] = {&fw.w2};
      const int n1[1] = {dff}, n2[1] = {d};
      if ((rc = tf32_pack_dense(w1p, n1, 1, d, pk_w1, pk_w1_b, st))) return rc;
      if ((rc = tf32_pack_dense(w2p, n2, 1, dff, pk_w2, pk_w2_b, st))) return rc;
      if ((rc = tf32_dense(sv.ad[blk], d, d, pk_w1, pk_w1_b, dff, B, sv.f1d[blk], dff, true, nullptr, 0, st))) return rc;
      if ((rc = tf32_dense(sv.f1d[blk], dff, dff, pk_w2, pk_w2_b, d, B, sv.z2d[blk], d, false, sv.ad[blk], d, st)))
        return rc;
    } else {
      {
        GemmGroup grp{};
        grp.use_tc = use_tc;
        fwd_prob(grp.p[0], sv.ad[blk], d, d, fw.w1, dff, B, sv.f1d[blk], dff, true, nullptr, 0);
        grp.n = 1;
        if ((rc = gemm_group_launch(grp, st))) return rc;
      }
      {
        GemmGroup grp{};
        grp.use_tc = use_tc;
        fwd_prob(grp.p[0], sv.f1d[blk], dff, dff, fw.w2, d, B, sv.z2d[blk], d, false, sv.ad[blk], d);
        grp.n = 1;
        if ((rc = gemm_group_launch(grp, st))) return rc;
      }
    }
    const bool last = blk == c.n_dec_blocks - 1;
    ln_fwd_rows_kernel<<<ln_grid, 256, 0, st>>>(sv.z2d[blk], B, d, fw.ln.gamma, fw.ln.beta, sv.din[blk + 1],
                                                last ? out : nullptr, out_ld);
    DMT_CUDA_LAUNCH_CHECK("ln_fwd_rows_kernel");
  }
  if (c.n_dec_blocks == 0) {
    cudaError_t e = cudaMemcpy2DAsync(out, out_ld * sizeof(float), sv.din[0], d * sizeof(float), d * sizeof(float), B,
                                      cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy2DAsync(interest)");
  }
  if (T > 0) {
    ZeroCapArgs z{sv, offsets, d, dff, LP, c.n_enc_blocks, c.n_dec_blocks};
    zero_capped_rows_kernel<<<B, 256, 0, st>>>(z);
    DMT_CUDA_LAUNCH_CHECK("zero_capped_rows_kernel");
  }
  return DMT_OK;
}

}  // namespace dmt
