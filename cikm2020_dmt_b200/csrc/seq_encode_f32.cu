// dmt_seq_encode_fwd, DMT_PRECISION_F32: one CTA keeps one whole behaviour sequence on chip.
//
//   ids --gather/concat/*sqrt(d)+pos--> X --QKV--> self-attention --LN--> FF --LN--> memory
//   target item --*sqrt(d)--> q --vanilla attention over memory--> LN --> FF --> LN --> interest
//
// Nothing but the ids, the embedding rows and the [d_model] interest vector touches HBM: the
// algorithmic bytes per (sample, sequence) are L*(sum_f D_f*4 + 20) + (sum_f D_f*4 + 20) + 4 +
// d*4.  Padded positions are never materialised (they are inert in the reference).
// All arithmetic is fp32 on CUDA cores; this is the exact-parity path (tolerance 1e-4 against
// the fp64 oracle) and the arithmetic reference for the bf16 tensor-core path.
#include "dmt_common.cuh"
#include "dropout.cuh"
#include "seq_train.cuh"

namespace dmt {

struct SeqArgs {
  dmt_seq_cfg cfg;
  dmt_seq_input in;
  dmt_seq_weights w;
  float* out;
  int64_t out_ld;
  int32_t col_off[DMT_MAX_SEQ_FEATS + 1];
  int32_t lp;    // rows reserved per on-chip activation buffer (>= longest sequence kept)
  int32_t ld;    // padded row stride of [*, d_model] buffers (d_model + 4: conflict-free float4 rows)
  int32_t ldh;   // padded row stride of the [*, d_ff] buffer
  int32_t region_floats;
  SeqSaved sv;   // SAVE instantiation only: where the training forward leaves its activations
};

constexpr int kThreads = 256;

// Y[t, n] = act(sum_k A[t, k] * W[k, n] + bias[n]) for t < L.  A lives in shared memory, W/bias in
// global memory in the TF [in, out] layout.  Each thread owns a TR x 4 register tile: W is read as
// coalesced float4 across the warp, A as broadcast float4 along k.
template <int TR>
__device__ __forceinline__ void gemm_rows(const float* __restrict__ As, int lda, int L, int K,
                                          const float* __restrict__ W, const float* __restrict__ bias, int N,
                                          float* __restrict__ Ys, int ldy, bool relu) {
  const int ncg = N >> 2;
  const int nrg = (L + TR - 1) / TR;
  for (int item = threadIdx.x; item < ncg * nrg; item += kThreads) {
    const int rg = item / ncg;
    const int n0 = (item - rg * ncg) << 2;
    const int t0 = rg * TR;
    float acc[TR][4];
#pragma unroll
    for (int r = 0; r < TR; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    const float* arow[TR];
#pragma unroll
    for (int r = 0; r < TR; ++r) arow[r] = As + min(t0 + r, L - 1) * lda;
    for (int k = 0; k < K; k += 4) {
      const float4 w0 = ldg4(W + (size_t)(k + 0) * N + n0);
      const float4 w1 = ldg4(W + (size_t)(k + 1) * N + n0);
      const float4 w2 = ldg4(W + (size_t)(k + 2) * N + n0);
      const float4 w3 = ldg4(W + (size_t)(k + 3) * N + n0);
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(arow[r] + k);
        acc[r][0] = fmaf(x.x, w0.x, fmaf(x.y, w1.x, fmaf(x.z, w2.x, fmaf(x.w, w3.x, acc[r][0]))));
        acc[r][1] = fmaf(x.x, w0.y, fmaf(x.y, w1.y, fmaf(x.z, w2.y, fmaf(x.w, w3.y, acc[r][1]))));
        acc[r][2] = fmaf(x.x, w0.z, fmaf(x.y, w1.z, fmaf(x.z, w2.z, fmaf(x.w, w3.z, acc[r][2]))));
        acc[r][3] = fmaf(x.x, w0.w, fmaf(x.y, w1.w, fmaf(x.z, w2.w, fmaf(x.w, w3.w, acc[r][3]))));
      }
    }
    const float4 bv = ldg4(bias + n0);
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const int t = t0 + r;
      if (t < L) {
        float4 y = make_float4(acc[r][0] + bv.x, acc[r][1] + bv.y, acc[r][2] + bv.z, acc[r][3] + bv.w);
        if (relu) {
          y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
        }
        *reinterpret_cast<float4*>(Ys + t * ldy + n0) = y;
      }
    }
  }
}

// y[n] = act(sum_k x[k] W[k, n] + b[n]) (+ res[n]) for one row held in shared memory.
__device__ __forceinline__ void gemv_row(const float* __restrict__ x, int K, const float* __restrict__ W,
                                         const float* __restrict__ bias, int N, float* __restrict__ y, bool relu,
                                         const float* __restrict__ res) {
  for (int n = threadIdx.x; n < N; n += kThreads) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(x[k], __ldg(W + (size_t)k * N + n), acc);
    acc += __ldg(bias + n);
    if (relu) acc = fmaxf(acc, 0.f);
    if (res) acc += res[n];
    y[n] = acc;
  }
}

// out[t, :] = LN(a[t, :] + (res ? res[t, :] : 0)) * gamma + beta, one warp per row, D <= 256.
// TransformerModel_util.py:58-78: biased variance, eps inside the sqrt.  In-place safe.
__device__ __forceinline__ void ln_rows(const float* a, int lda, const float* res, int ldr, int L, int D,
                                        const float* __restrict__ gamma, const float* __restrict__ beta, float* out,
                                        int ldo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = warp; t < L; t += kThreads / 32) {
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      v[i] = 0.f;
      if (c < D) {
        v[i] = a[t * lda + c] + (res ? res[t * ldr + c] : 0.f);
        s += v[i];
      }
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      if (c < D) {
        const float dlt = v[i] - mean;
        q += dlt * dlt;
      }
    }
    const float var = warp_sum(q) / (float)D;
    const float rstd = 1.0f / sqrtf(var + kLnEps);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      if (c < D) out[t * ldo + c] = __ldg(gamma + c) * ((v[i] - mean) * rstd) + __ldg(beta + c);
    }
  }
}

__device__ __forceinline__ float lookup_elem(const float* __restrict__ table, int64_t rows, int dim, int id, int c,
                                             int zero_pad) {
  const int64_t row = (int64_t)id - (zero_pad ? 1 : 0);
  if (row < 0 || row >= rows) return 0.f;   // index 0 under zero_pad == the all-zero row (base.py:89)
  return __ldg(table + row * dim + c);
}

// rows [0, L) of an on-chip [*, ld] buffer -> global rows off + t (row stride gld floats, first column gcol)
__device__ __forceinline__ void save_rows(const float* s, int ld, int L, int W, float* g, int64_t off, int gld,
                                          int gcol) {
  for (int i = threadIdx.x; i < L * W; i += kThreads) {
    const int t = i / W, c = i - t * W;
    g[(off + t) * gld + gcol + c] = s[t * ld + c];
  }
}
__device__ __forceinline__ void zero_rows(float* g, int64_t off, int t0, int t1, int W) {
  for (int i = threadIdx.x; i < (t1 - t0) * W; i += kThreads) g[(off + t0) * W + i] = 0.f;
}

template <bool SAVE>
__global__ void __launch_bounds__(kThreads, 2) seq_encode_f32_kernel(const __grid_constant__ SeqArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int D = a.cfg.d_model, DFF = a.cfg.d_ff, H = a.cfg.num_heads, dk = D / H;
  const int LP = a.lp, ld = a.ld, ldh = a.ldh, lds = LP + 1;
  const int nf = a.cfg.n_feats;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;

  float* X = smem;                       // [LP][ld]  encoder input, later the FF output / memory
  float* A = X + LP * ld;                // [LP][ld]  attention output after LN
  float* R = A + LP * ld;                // region: Q,K,V,S  |  FF hidden
  float* Q = R;
  float* Kb = Q + LP * ld;
  float* V = Kb + LP * ld;
  float* S = V + LP * ld;                // [LP][LP+1] one head at a time
  float* Hb = R;                         // [LP][ldh]
  float* vec = R + a.region_floats;      // small per-sample vectors
  float* dvec = vec;                     // [D] decoder state
  float* qd = dvec + D;                  // [D] projected query
  float* avec = qd + D;                  // [D]
  float* ovec = avec + D;                // [D]
  float* hvec = ovec + D;                // [DFF]
  float* sc = hvec + DFF;                // [H][LP] decoder attention probabilities

  const int off_last = __ldg(a.in.offsets[nf - 1] + b);
  const int len_all = __ldg(a.in.offsets[nf - 1] + b + 1) - off_last;
  const int L = min(len_all, LP);
  const float sqrt_d = sqrtf((float)D);
  const float scale = 1.0f / sqrtf((float)dk);
  // training forward only (SAVE): the dropout sites of B12; rate 0 / eval -> every multiplier is 1
  const float drate = SAVE ? a.cfg.dropout_rate : 0.f;
  const Dropout drop_enc(drate, a.cfg.dropout_seed, kSiteEncIn), drop_dec(drate, a.cfg.dropout_seed, kSiteDecIn);

  // ---- A2/A3: gather + concat + scale + learned position (mmoe_transformer_unbias.py:153-158,181;
  //      TransformerModel.py:97-100) ----
  for (int i = tid; i < L * D; i += kThreads) {
    const int t = i / D, c = i - t * D;
    int f = 0;
    while (f + 1 < nf && c >= a.col_off[f + 1]) ++f;
    const int off = __ldg(a.in.offsets[f] + b);
    const int len_f = __ldg(a.in.offsets[f] + b + 1) - off;
    const int id = (t < len_f) ? __ldg(a.in.ids[f] + off + t) : 0;
    const float e = lookup_elem(a.in.table[f], a.in.rows[f], a.in.dim[f], id, c - a.col_off[f], a.cfg.zero_pad);
    X[t * ld + c] = (e * sqrt_d + __ldg(a.w.pos + t * D + c)) * drop_enc.mult((uint32_t)((off_last + t) * D + c));
  }
  for (int c = tid; c < D; c += kThreads) {
    int f = 0;
    while (f + 1 < nf && c >= a.col_off[f + 1]) ++f;
    const int id = __ldg(a.in.item_ids[f] + b);
    dvec[c] = lookup_elem(a.in.table[f], a.in.rows[f], a.in.dim[f], id, c - a.col_off[f], a.cfg.zero_pad) * sqrt_d *
              drop_dec.mult((uint32_t)(b * D + c));
  }
  __syncthreads();
  if (SAVE) {
    save_rows(X, ld, L, D, a.sv.hin[0], off_last, D, 0);
    if (len_all > L) {   // tokens beyond the on-chip cap: inert zero rows in every saved buffer
      for (int blk = 0; blk <= a.cfg.n_enc_blocks; ++blk) zero_rows(a.sv.hin[blk], off_last, L, len_all, D);
      for (int blk = 0; blk < a.cfg.n_enc_blocks; ++blk) {
        zero_rows(a.sv.qkv[blk], off_last, L, len_all, 3 * D);
        zero_rows(a.sv.z1[blk], off_last, L, len_all, D);
        zero_rows(a.sv.a[blk], off_last, L, len_all, D);
        zero_rows(a.sv.f1[blk], off_last, L, len_all, DFF);
        zero_rows(a.sv.z2[blk], off_last, L, len_all, D);
      }
      for (int blk = 0; blk < a.cfg.n_dec_blocks; ++blk) zero_rows(a.sv.kvd[blk], off_last, L, len_all, 2 * D);
    }
  }

  // ---- A3-A6: encoder blocks ----
  for (int blk = 0; blk < a.cfg.n_enc_blocks && L > 0; ++blk) {
    const dmt_attn_weights& aw = a.w.enc_attn[blk];
    const dmt_ff_weights& fw = a.w.ff[blk];
    gemm_rows<8>(X, ld, L, D, aw.q.w, aw.q.b, D, Q, ld, false);
    gemm_rows<8>(X, ld, L, D, aw.k.w, aw.k.b, D, Kb, ld, false);
    gemm_rows<8>(X, ld, L, D, aw.v.w, aw.v.b, D, V, ld, false);
    __syncthreads();
    if (SAVE) {
      save_rows(Q, ld, L, D, a.sv.qkv[blk], off_last, 3 * D, 0);
      save_rows(Kb, ld, L, D, a.sv.qkv[blk], off_last, 3 * D, D);
      save_rows(V, ld, L, D, a.sv.qkv[blk], off_last, 3 * D, 2 * D);
    }
    const Dropout drop_p(drate, a.cfg.dropout_seed, kSiteSelfProbs + blk);
    for (int h = 0; h < H; ++h) {
      const int hc = h * dk;
      for (int i = tid; i < L * L; i += kThreads) {
        const int qi = i / L, kj = i - qi * L;
        const float* qp = Q + qi * ld + hc;
        const float* kp = Kb + kj * ld + hc;
        float acc = 0.f;
        if ((dk & 3) == 0) {
          for (int c = 0; c < dk; c += 4) {
            const float4 qv = *reinterpret_cast<const float4*>(qp + c);
            const float4 kv = *reinterpret_cast<const float4*>(kp + c);
            acc = fmaf(qv.x, kv.x, fmaf(qv.y, kv.y, fmaf(qv.z, kv.z, fmaf(qv.w, kv.w, acc))));
          }
        } else {
          for (int c = 0; c < dk; ++c) acc = fmaf(qp[c], kp[c], acc);
        }
        S[qi * lds + kj] = acc * scale;
      }
      __syncthreads();
      // softmax over the L valid keys; masked keys contribute exp(-2^32+1 - max) == 0 in fp32
      for (int r = warp; r < L; r += kThreads / 32) {
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
          S[r * lds + j] *= inv * drop_p.mult((uint32_t)(((b * H + h) * LP + r) * LP + j));
      }
      __syncthreads();
      // context; overwrites this head's Q columns (dead after the scores)
      for (int i = tid; i < L * dk; i += kThreads) {
        const int qi = i / dk, c = i - qi * dk;
        float acc = 0.f;
        for (int j = 0; j < L; ++j) acc = fmaf(S[qi * lds + j], V[j * ld + hc + c], acc);
        Q[qi * ld + hc + c] = acc;
      }
      __syncthreads();
    }
    if (SAVE) {
      for (int i = tid; i < L * D; i += kThreads) {
        const int t = i / D, c = i - t * D;
        a.sv.z1[blk][(off_last + t) * D + c] = Q[t * ld + c] + X[t * ld + c];
      }
    }
    ln_rows(Q, ld, X, ld, L, D, aw.ln.gamma, aw.ln.beta, A, ld);   // residual = the block input
    __syncthreads();
    if (SAVE) save_rows(A, ld, L, D, a.sv.a[blk], off_last, D, 0);
    gemm_rows<8>(A, ld, L, D, fw.w1.w, fw.w1.b, DFF, Hb, ldh, true);
    __syncthreads();
    if (SAVE) save_rows(Hb, ldh, L, DFF, a.sv.f1[blk], off_last, DFF, 0);
    gemm_rows<4>(Hb, ldh, L, DFF, fw.w2.w, fw.w2.b, D, X, ld, false);
    __syncthreads();
    if (SAVE) {
      for (int i = tid; i < L * D; i += kThreads) {
        const int t = i / D, c = i - t * D;
        a.sv.z2[blk][(off_last + t) * D + c] = X[t * ld + c] + A[t * ld + c];
      }
      __syncthreads();   // the LayerNorm below rewrites X in place
    }
    ln_rows(X, ld, A, ld, L, D, fw.ln.gamma, fw.ln.beta, X, ld);
    __syncthreads();
    if (SAVE) save_rows(X, ld, L, D, a.sv.hin[blk + 1], off_last, D, 0);
  }

  // ---- A7: decoder blocks, single query over the encoder memory (TransformerModel.py:125-171) ----
  for (int blk = 0; blk < a.cfg.n_dec_blocks; ++blk) {
    const dmt_attn_weights& aw = a.w.dec_attn[blk];
    const dmt_ff_weights& fw = a.w.ff[blk];
    if (SAVE)
      for (int c = tid; c < D; c += kThreads) a.sv.din[blk][(int64_t)b * D + c] = dvec[c];
    gemv_row(dvec, D, aw.q.w, aw.q.b, D, qd, false, nullptr);
    if (L > 0) {
      gemm_rows<8>(X, ld, L, D, aw.k.w, aw.k.b, D, Kb, ld, false);
      gemm_rows<8>(X, ld, L, D, aw.v.w, aw.v.b, D, V, ld, false);
    }
    __syncthreads();
    if (SAVE) {
      for (int c = tid; c < D; c += kThreads) a.sv.qd[blk][(int64_t)b * D + c] = qd[c];
      save_rows(Kb, ld, L, D, a.sv.kvd[blk], off_last, 2 * D, 0);
      save_rows(V, ld, L, D, a.sv.kvd[blk], off_last, 2 * D, D);
    }
    for (int i = tid; i < H * L; i += kThreads) {
      const int h = i / L, j = i - h * L;
      float acc = 0.f;
      for (int c = 0; c < dk; ++c) acc = fmaf(qd[h * dk + c], Kb[j * ld + h * dk + c], acc);
      sc[h * LP + j] = acc * scale;
    }
    __syncthreads();
    for (int h = warp; h < H; h += kThreads / 32) {
      float m = -INFINITY;
      for (int j = lane; j < L; j += 32) m = fmaxf(m, sc[h * LP + j]);
      m = warp_max(m);
      float s = 0.f;
      for (int j = lane; j < L; j += 32) {
        const float e = expf(sc[h * LP + j] - m);
        sc[h * LP + j] = e;
        s += e;
      }
      s = warp_sum(s);
      const float inv = s > 0.f ? 1.0f / s : 0.f;
      for (int j = lane; j < L; j += 32) sc[h * LP + j] *= inv;
    }
    __syncthreads();
    const Dropout drop_v(drate, a.cfg.dropout_seed, kSiteVanillaProbs + blk);
    for (int c = tid; c < D; c += kThreads) {
      const int h = c / dk;
      float acc = 0.f;
      for (int j = 0; j < L; ++j)
        acc = fmaf(sc[h * LP + j] * drop_v.mult((uint32_t)((b * H + h) * LP + j)), V[j * ld + c], acc);
      ovec[c] = acc;
    }
    __syncthreads();
    if (SAVE) {
      for (int i = tid; i < H * LP; i += kThreads) {
        const int j = i % LP;
        a.sv.pd[blk][(int64_t)b * H * LP + i] = j < L ? sc[i] : 0.f;
      }
      for (int c = tid; c < D; c += kThreads) a.sv.z1d[blk][(int64_t)b * D + c] = ovec[c] + dvec[c];
    }
    ln_rows(ovec, D, dvec, D, 1, D, aw.ln.gamma, aw.ln.beta, avec, D);
    __syncthreads();
    if (SAVE)
      for (int c = tid; c < D; c += kThreads) a.sv.ad[blk][(int64_t)b * D + c] = avec[c];
    gemv_row(avec, D, fw.w1.w, fw.w1.b, DFF, hvec, true, nullptr);
    __syncthreads();
    if (SAVE)
      for (int c = tid; c < DFF; c += kThreads) a.sv.f1d[blk][(int64_t)b * DFF + c] = hvec[c];
    gemv_row(hvec, DFF, fw.w2.w, fw.w2.b, D, ovec, false, avec);
    __syncthreads();
    if (SAVE)
      for (int c = tid; c < D; c += kThreads) a.sv.z2d[blk][(int64_t)b * D + c] = ovec[c];
    ln_rows(ovec, D, nullptr, 0, 1, D, fw.ln.gamma, fw.ln.beta, dvec, D);
    __syncthreads();
  }
  if (SAVE)
    for (int c = tid; c < D; c += kThreads) a.sv.din[a.cfg.n_dec_blocks][(int64_t)b * D + c] = dvec[c];
  for (int c = tid; c < D; c += kThreads) a.out[(int64_t)b * a.out_ld + c] = dvec[c];
}

int seq_encode_f32_launch(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, float* out,
                          int64_t out_ld, const SeqSaved* saved, cudaStream_t st) {
  SeqArgs a;
  a.sv = saved ? *saved : SeqSaved{};
  a.cfg = *cfg;
  a.in = *in;
  a.w = *w;
  a.out = out;
  a.out_ld = out_ld;
  int col = 0;
  for (int f = 0; f < cfg->n_feats; ++f) {
    a.col_off[f] = col;
    col += in->dim[f];
  }
  for (int f = cfg->n_feats; f <= DMT_MAX_SEQ_FEATS; ++f) a.col_off[f] = col;
  DMT_REQUIRE(col == cfg->d_model, DMT_ERR_INVALID_ARGUMENT,
              "dmt_seq_encode_fwd: pair dims sum to %d, d_model is %d", col, cfg->d_model);
  const int D = cfg->d_model, DFF = cfg->d_ff;
  a.lp = cfg->maxlen < DMT_MAX_SEQ_LEN ? cfg->maxlen : DMT_MAX_SEQ_LEN;
  a.ld = D + 4;
  a.ldh = DFF + 4;
  const int attn_floats = 3 * a.lp * a.ld + a.lp * (a.lp + 1);
  const int ff_floats = a.lp * a.ldh;
  a.region_floats = ((attn_floats > ff_floats ? attn_floats : ff_floats) + 3) & ~3;
  const size_t smem_floats = (size_t)2 * a.lp * a.ld + a.region_floats + 4 * D + DFF + cfg->num_heads * a.lp + 8;
  const size_t smem_bytes = smem_floats * sizeof(float);
  DMT_REQUIRE(smem_bytes <= 227 * 1024, DMT_ERR_UNSUPPORTED_SHAPE,
              "dmt_seq_encode_fwd: sequence tile needs %zu B shared memory (> 227 KB)", smem_bytes);
  cudaError_t e = cudaFuncSetAttribute(saved ? seq_encode_f32_kernel<true> : seq_encode_f32_kernel<false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(seq_encode_f32_kernel)");
  if (saved)
    seq_encode_f32_kernel<true><<<cfg->batch, kThreads, smem_bytes, st>>>(a);
  else
    seq_encode_f32_kernel<false><<<cfg->batch, kThreads, smem_bytes, st>>>(a);
  DMT_CUDA_LAUNCH_CHECK("seq_encode_f32_kernel");
  return DMT_OK;
}

}  // namespace dmt
