// dmt_seq_encode_bwd (fp32): backward of one behaviour sequence's encoder/decoder
// (TransformerModel.py:84-171, TransformerModel_util.py:11-108,160-235) from the activations the training
// forward saved.  Row-batched pipeline:
//
//   d(interest) -> [LN bwd -> FF bwd (2 GEMMs) -> LN bwd -> single-query attention bwd -> dD, dMemory] per decoder block
//               -> [LN bwd -> FF bwd -> LN bwd -> self-attention bwd (per sample) -> dX GEMM]          per encoder block
//               -> token-row gradients (x sqrt(d)), position-table gradient, target-row gradients
//
// Weight gradients are  saved_activation^T x gradient  contractions over all tokens: grouped split-K
// GEMMs with a fixed-order reduction (gemm_f32.cuh).  Nothing here uses floating-point atomics.
#include "dropout.cuh"
#include "gemm_f32.cuh"
#include "gemm_tf32.cuh"
#include "seq_train.cuh"

namespace dmt {

namespace {

constexpr int kLnWarps = 8;
constexpr int kLnMaxCols = 8;   // D <= 256

struct LnBwdArgs {
  const float* z;      // LayerNorm input rows
  const float* dy;
  const float* gamma;
  float* dz;
  float* partial;      // [gridDim.x][2*D]: per-CTA sums of dy*xhat | dy
  int64_t ldz, lddy, lddz, rows;
  int D;
};

// One warp per row: dz = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma
// (backward of TransformerModel_util.py:58-78, biased variance, eps inside the sqrt).
__global__ void __launch_bounds__(kLnWarps * 32) ln_bwd_kernel(const __grid_constant__ LnBwdArgs a) {
  __shared__ float red[kLnWarps][2 * 32 * kLnMaxCols];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = a.D;
  float gam[kLnMaxCols], dg[kLnMaxCols], db[kLnMaxCols];
#pragma unroll
  for (int i = 0; i < kLnMaxCols; ++i) {
    const int c = lane + 32 * i;
    gam[i] = c < D ? __ldg(a.gamma + c) : 0.f;
    dg[i] = db[i] = 0.f;
  }
  const float invD = 1.0f / (float)D;
  for (int64_t r = (int64_t)blockIdx.x * kLnWarps + warp; r < a.rows; r += (int64_t)gridDim.x * kLnWarps) {
    float v[kLnMaxCols], dy[kLnMaxCols];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxCols; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < D ? __ldg(a.z + r * a.ldz + c) : 0.f;
      dy[i] = c < D ? __ldg(a.dy + r * a.lddy + c) : 0.f;
      s += v[i];
    }
    const float mean = warp_sum(s) * invD;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxCols; ++i) {
      const int c = lane + 32 * i;
      if (c < D) {
        const float dl = v[i] - mean;
        q += dl * dl;
      }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * invD + kLnEps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxCols; ++i) {
      const int c = lane + 32 * i;
      if (c < D) {
        v[i] = (v[i] - mean) * rstd;      // xhat
        const float g = dy[i] * gam[i];
        sg += g;
        sgx += g * v[i];
        dg[i] += dy[i] * v[i];
        db[i] += dy[i];
      }
    }
    sg = warp_sum(sg) * invD;
    sgx = warp_sum(sgx) * invD;
#pragma unroll
    for (int i = 0; i < kLnMaxCols; ++i) {
      const int c = lane + 32 * i;
      if (c < D) a.dz[r * a.lddz + c] = rstd * (dy[i] * gam[i] - sg - v[i] * sgx);
    }
  }
#pragma unroll
  for (int i = 0; i < kLnMaxCols; ++i) {
    const int c = lane + 32 * i;
    if (c < D) {
      red[warp][c] = dg[i];
      red[warp][D + c] = db[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * D; c += kLnWarps * 32) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) s += red[w][c];
    a.partial[(int64_t)blockIdx.x * 2 * D + c] = s;
  }
}

// dgamma[c] += sum_cta partial[cta][c]; dbeta[c] += sum_cta partial[cta][D + c] (fixed order).
__global__ void ln_param_reduce_kernel(const float* __restrict__ partial, int n_cta, int D, float* dgamma,
                                       float* dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * D) return;
  float s = 0.f;
  for (int i = 0; i < n_cta; ++i) s += partial[(int64_t)i * 2 * D + c];
  if (c < D)
    dgamma[c] += s;
  else
    dbeta[c - D] += s;
}

int ln_bwd_launch(const float* z, int64_t ldz, const float* dy, int64_t lddy, const float* gamma, float* dz,
                  int64_t lddz, int64_t rows, int D, float* partial, int grid, float* dgamma, float* dbeta,
                  cudaStream_t st) {
  LnBwdArgs a{z, dy, gamma, dz, partial, ldz, lddy, lddz, rows, D};
  ln_bwd_kernel<<<grid, kLnWarps * 32, 0, st>>>(a);
  DMT_CUDA_LAUNCH_CHECK("ln_bwd_kernel");
  ln_param_reduce_kernel<<<(2 * D + 127) / 128, 128, 0, st>>>(partial, grid, D, dgamma, dbeta);
  DMT_CUDA_LAUNCH_CHECK("ln_param_reduce_kernel");
  return DMT_OK;
}

// ---------------------------------------------------------------------------------------------------
// Self-attention backward, one CTA per sample (TransformerModel_util.py:11-56 + head split :193-201).
//   P = softmax(Q_h K_h^T / sqrt(dk)) over the L valid keys (recomputed), O_h = P V_h
//   dV_h = P^T dO_h ; dP = dO_h V_h^T ; dS = P * (dP - rowsum(P * dP)) / sqrt(dk)
//   dQ_h = dS K_h ; dK_h = dS^T Q_h
struct AttnBwdArgs {
  const float* qkv;        // [T, 3D]
  const float* d_o;        // [T, D] gradient of the attention context (= LayerNorm-input gradient)
  float* dqkv;             // [T, 3D]
  const int32_t* offsets;  // [B+1]
  int D, H, LP;
  Dropout drop;            // the forward's attention-probability mask M: O_h = (P * M) V_h
};

constexpr int kAttnThreads = 256;

__global__ void __launch_bounds__(kAttnThreads) attn_bwd_kernel(const __grid_constant__ AttnBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int D = a.D, H = a.H, LP = a.LP, dk = D / H, ld = D + 1, lds = LP + 1;   // odd strides: conflict-free
  float* Q = sm;
  float* K = Q + LP * ld;
  float* V = K + LP * ld;
  float* dO = V + LP * ld;
  float* P = dO + LP * ld;      // [LP][lds]
  float* dS = P + LP * lds;     // [LP][lds]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t off = __ldg(a.offsets + b);
  const int len_all = __ldg(a.offsets + b + 1) - (int)off;
  const int L = min(len_all, LP);
  for (int i = tid; i < (len_all - L) * 3 * D; i += kAttnThreads) a.dqkv[(off + L) * 3 * D + i] = 0.f;
  if (L == 0) return;
  for (int i = tid; i < L * D; i += kAttnThreads) {
    const int t = i / D, c = i - t * D;
    const float* row = a.qkv + (off + t) * 3 * D;
    Q[t * ld + c] = __ldg(row + c);
    K[t * ld + c] = __ldg(row + D + c);
    V[t * ld + c] = __ldg(row + 2 * D + c);
    dO[t * ld + c] = __ldg(a.d_o + (off + t) * D + c);
  }
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)dk);
  for (int h = 0; h < H; ++h) {
    const int hc = h * dk;
    // 2 x 2 register tile per thread (rows qa, qa+hl; keys kb, kb+hl): one shared-memory load per FMA instead of
    // two; consecutive lanes take consecutive keys, so the K / V rows (odd stride) are conflict-free
    const int hl = (L + 1) >> 1;
    for (int i = tid; i < hl * hl; i += kAttnThreads) {
      const int qa = i / hl, kb = i - qa * hl;
      const int qb = min(qa + hl, L - 1), kc = min(kb + hl, L - 1);
      const float *q0 = Q + qa * ld + hc, *q1 = Q + qb * ld + hc, *k0 = K + kb * ld + hc, *k1 = K + kc * ld + hc;
      const float *o0 = dO + qa * ld + hc, *o1 = dO + qb * ld + hc, *v0 = V + kb * ld + hc, *v1 = V + kc * ld + hc;
      float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f, p00 = 0.f, p01 = 0.f, p10 = 0.f, p11 = 0.f;
      for (int c = 0; c < dk; ++c) {
        const float a0 = q0[c], a1 = q1[c], b0 = k0[c], b1 = k1[c];
        s00 = fmaf(a0, b0, s00); s01 = fmaf(a0, b1, s01); s10 = fmaf(a1, b0, s10); s11 = fmaf(a1, b1, s11);
        const float e0 = o0[c], e1 = o1[c], w0 = v0[c], w1 = v1[c];
        p00 = fmaf(e0, w0, p00); p01 = fmaf(e0, w1, p01); p10 = fmaf(e1, w0, p10); p11 = fmaf(e1, w1, p11);
      }
      P[qa * lds + kb] = s00 * scale; dS[qa * lds + kb] = p00;
      if (kb + hl < L) { P[qa * lds + kb + hl] = s01 * scale; dS[qa * lds + kb + hl] = p01; }
      if (qa + hl < L) {
        P[(qa + hl) * lds + kb] = s10 * scale; dS[(qa + hl) * lds + kb] = p10;
        if (kb + hl < L) { P[(qa + hl) * lds + kb + hl] = s11 * scale; dS[(qa + hl) * lds + kb + hl] = p11; }
      }
    }
    __syncthreads();
    for (int r = warp; r < L; r += kAttnThreads / 32) {
      float m = -INFINITY;
      for (int j = lane; j < L; j += 32) m = fmaxf(m, P[r * lds + j]);
      m = warp_max(m);
      float s = 0.f;
      for (int j = lane; j < L; j += 32) {
        const float e = expf(P[r * lds + j] - m);
        P[r * lds + j] = e;
        s += e;
      }
      const float inv = 1.0f / warp_sum(s);
      float pd = 0.f;
      // with dropout: dS holds d(P*M); dP = d(P*M) * M, and the value path (dV) needs P*M
      for (int j = lane; j < L; j += 32) {
        const float p = P[r * lds + j] * inv;
        const float dp = dS[r * lds + j] * a.drop.mult((uint32_t)(((b * H + h) * LP + r) * LP + j));
        P[r * lds + j] = p;
        dS[r * lds + j] = dp;
        pd = fmaf(p, dp, pd);
      }
      pd = warp_sum(pd);
      for (int j = lane; j < L; j += 32) {
        const float p = P[r * lds + j];
        dS[r * lds + j] = p * (dS[r * lds + j] - pd) * scale;
        P[r * lds + j] = p * a.drop.mult((uint32_t)(((b * H + h) * LP + r) * LP + j));
      }
    }
    __syncthreads();
    // two rows (t, t+hl) per thread share the K / Q / dO loads of every j
    for (int i = tid; i < hl * dk; i += kAttnThreads) {
      const int t = i / dk, c = i - t * dk;
      const int t2 = min(t + hl, L - 1);
      float dq = 0.f, dkk = 0.f, dv = 0.f, dq2 = 0.f, dkk2 = 0.f, dv2 = 0.f;
      for (int j = 0; j < L; ++j) {
        const float kx = K[j * ld + hc + c], qx = Q[j * ld + hc + c], ox = dO[j * ld + hc + c];
        dq = fmaf(dS[t * lds + j], kx, dq);        // t = query row
        dkk = fmaf(dS[j * lds + t], qx, dkk);      // t = key row
        dv = fmaf(P[j * lds + t], ox, dv);
        dq2 = fmaf(dS[t2 * lds + j], kx, dq2);
        dkk2 = fmaf(dS[j * lds + t2], qx, dkk2);
        dv2 = fmaf(P[j * lds + t2], ox, dv2);
      }
      float* row = a.dqkv + (off + t) * 3 * D + hc + c;
      row[0] = dq;
      row[D] = dkk;
      row[2 * D] = dv;
      if (t + hl < L) {
        float* row2 = a.dqkv + (off + t + hl) * 3 * D + hc + c;
        row2[0] = dq2;
        row2[D] = dkk2;
        row2[2 * D] = dv2;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// Vanilla (decoder) attention backward: ONE query per sample over the L memory rows
// (TransformerModel.py:157-166).  od_h = sum_j p[h][j] Vd[j,h]; the probabilities were saved.
struct DecAttnBwdArgs {
  const float* qd;         // [B, D]
  const float* kvd;        // [T, 2D]
  const float* pd;         // [B, H, LP]
  const float* d_od;       // [B, D] gradient of the context (= LayerNorm-input gradient)
  float* dqd;              // [B, D]
  float* dkvd;             // [T, 2D]
  const int32_t* offsets;
  int D, H, LP;
  Dropout drop;
};

constexpr int kDecThreads = 128;

__global__ void __launch_bounds__(kDecThreads) dec_attn_bwd_kernel(const __grid_constant__ DecAttnBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int D = a.D, H = a.H, LP = a.LP, dk = D / H, ld = 2 * D + 1;
  float* KV = sm;                  // [LP][2D+1]
  float* q = KV + LP * ld;         // [D]
  float* dod = q + D;              // [D]
  float* p = dod + D;              // [H][LP]
  float* ds = p + H * LP;          // [H][LP]
  float* rsum = ds + H * LP;       // [H]
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t off = __ldg(a.offsets + b);
  const int len_all = __ldg(a.offsets + b + 1) - (int)off;
  const int L = min(len_all, LP);
  for (int i = tid; i < (len_all - L) * 2 * D; i += kDecThreads) a.dkvd[(off + L) * 2 * D + i] = 0.f;
  for (int i = tid; i < L * 2 * D; i += kDecThreads) {
    const int t = i / (2 * D), c = i - t * 2 * D;
    KV[t * ld + c] = __ldg(a.kvd + (off + t) * 2 * D + c);
  }
  for (int c = tid; c < D; c += kDecThreads) {
    q[c] = __ldg(a.qd + (int64_t)b * D + c);
    dod[c] = __ldg(a.d_od + (int64_t)b * D + c);
  }
  for (int i = tid; i < H * LP; i += kDecThreads) p[i] = __ldg(a.pd + (int64_t)b * H * LP + i);
  __syncthreads();
  for (int i = tid; i < H * L; i += kDecThreads) {
    const int h = i / L, j = i - h * L;
    float dp = 0.f;
    for (int c = 0; c < dk; ++c) dp = fmaf(dod[h * dk + c], KV[j * ld + D + h * dk + c], dp);
    ds[h * LP + j] = dp * a.drop.mult((uint32_t)((b * H + h) * LP + j));   // d(p*M) -> dp
  }
  __syncthreads();
  if (tid < H) {
    float r = 0.f;
    for (int j = 0; j < L; ++j) r = fmaf(p[tid * LP + j], ds[tid * LP + j], r);
    rsum[tid] = r;
  }
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)dk);
  for (int i = tid; i < H * L; i += kDecThreads) {
    const int h = i / L, j = i - h * L;
    ds[h * LP + j] = p[h * LP + j] * (ds[h * LP + j] - rsum[h]) * scale;
  }
  __syncthreads();
  for (int i = tid; i < L * D; i += kDecThreads) {
    const int j = i / D, c = i - j * D, h = c / dk;
    float* row = a.dkvd + (off + j) * 2 * D;
    row[c] = ds[h * LP + j] * q[c];
    row[D + c] = p[h * LP + j] * a.drop.mult((uint32_t)((b * H + h) * LP + j)) * dod[c];
  }
  for (int c = tid; c < D; c += kDecThreads) {
    const int h = c / dk;
    float acc = 0.f;
    for (int j = 0; j < L; ++j) acc = fmaf(ds[h * LP + j], KV[j * ld + c], acc);
    a.dqd[(int64_t)b * D + c] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------
// Position-table gradient (positional_encoding_learn, TransformerModel_util.py:281-316): the encoder
// input is rows * sqrt(d) + P[t], so dP[t] = sum_b dH0[b, t] = (1/sqrt(d)) * sum_b d_tokens[b, t].
// One CTA per position, fixed summation order.
constexpr int kPosSplits = 16;

struct PosGradArgs {
  const float* d_tokens;   // [T, D]
  const int32_t* offsets;
  float* partial;          // [kPosSplits][LP][D]
  float* dpos;             // [maxlen, D]  +=
  float inv_scale;
  int B, D, LP;
};

// grid (LP, kPosSplits): CTA (t, s) sums position t over samples b = s, s + kPosSplits, ... (fixed order)
__global__ void __launch_bounds__(256) pos_grad_kernel(const __grid_constant__ PosGradArgs a) {
  __shared__ float red[256];
  const int t = blockIdx.x, split = blockIdx.y, D = a.D;
  const int groups = 256 / D > 0 ? 256 / D : 1;
  const int c = threadIdx.x % D, grp = threadIdx.x / D;
  float s = 0.f;
  if (threadIdx.x < groups * D) {
    for (int b = split + grp * kPosSplits; b < a.B; b += groups * kPosSplits) {
      const int off = __ldg(a.offsets + b);
      const int len = min(__ldg(a.offsets + b + 1) - off, a.LP);
      if (t < len) s += __ldg(a.d_tokens + ((int64_t)off + t) * D + c);
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < D) {
    float tot = 0.f;
    for (int gi = 0; gi < groups; ++gi) tot += red[gi * D + threadIdx.x];
    a.partial[((int64_t)split * a.LP + t) * D + threadIdx.x] = tot;
  }
}

__global__ void pos_grad_reduce_kernel(const __grid_constant__ PosGradArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.LP * a.D) return;
  float tot = 0.f;
  for (int s = 0; s < kPosSplits; ++s) tot += a.partial[(int64_t)s * a.LP * a.D + i];
  a.dpos[i] += tot * a.inv_scale;
}

// In-place backward of an input-dropout site: g[i] *= M[i] (rows x D, element index = r * D + c).
__global__ void dropout_bwd_kernel(float* __restrict__ g, int64_t n, Dropout drop) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g[i] *= drop.mult((uint32_t)i);
}

__global__ void scale_copy_kernel(const float* __restrict__ src, int64_t lds_, float* __restrict__ dst, int64_t ldd,
                                  int64_t rows, int D, float alpha) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * D) return;
  const int64_t r = i / D;
  const int c = (int)(i - r * D);
  dst[r * ldd + c] = src[r * lds_ + c] * alpha;
}

// ---------------------------------------------------------------------------------------------------
struct BwdWs {
  float *dh[2];     // [T, d]  gradient of an encoder block output / input (ping-pong)
  float *dz2, *da, *dz1;   // [T, d]
  float *df1;       // [T, dff]
  float *dqkv;      // [T, 3d]
  float *dkvd;      // [T, 2d]
  float *dd[2];     // [B, d]  gradient of a decoder block output / input
  float *dz2d, *dad, *dz1d, *dqd;   // [B, d]
  float *df1d;      // [B, dff]
  float *ln_partial;                // [ln_grid][2d]
  float *pos_partial;               // [kPosSplits][LP][d]
  float *gpart[8];  // split-K scratch of the weight-gradient contractions (one block's worth)
  float *tf_pack;   // DMT_PRECISION_TF32: concatenated [Wq|Wk|Wv] / [Wk|Wv] operand of the dX GEMMs
  float *tf_colsum; //                     scratch of the bias-gradient column sums
  float *tf_wgrad;  //                     split partials of the token-contraction GEMMs
  int ln_grid;
  int splits_T[8], splits_B[8];
};

// weight-gradient problems of one block: index -> (M, N)
//   0 W1 [d,dff]  1 W2 [dff,d]  2..4 Wq/Wk/Wv [d,d]
inline void wg_shape(const dmt_seq_cfg& c, int i, int* M, int* N) {
  const int d = c.d_model, dff = c.d_ff;
  switch (i) {
    case 0: *M = d; *N = dff; break;
    case 1: *M = dff; *N = d; break;
    default: *M = d; *N = d; break;
  }
}

size_t bwd_carve(const dmt_seq_cfg& c, int64_t T, void* base, BwdWs* out) {
  Carver cv(base);
  BwdWs w{};
  const size_t d = c.d_model, dff = c.d_ff, B = c.batch;
  w.dh[0] = cv.take(T * d);
  w.dh[1] = cv.take(T * d);
  w.dz2 = cv.take(T * d);
  w.da = cv.take(T * d);
  w.dz1 = cv.take(T * d);
  w.df1 = cv.take(T * dff);
  w.dqkv = cv.take(T * 3 * d);
  w.dkvd = cv.take(T * 2 * d);
  w.dd[0] = cv.take(B * d);
  w.dd[1] = cv.take(B * d);
  w.dz2d = cv.take(B * d);
  w.dad = cv.take(B * d);
  w.dz1d = cv.take(B * d);
  w.dqd = cv.take(B * d);
  w.df1d = cv.take(B * dff);
  w.ln_grid = 2 * sm_count_cached();
  w.ln_partial = cv.take((size_t)w.ln_grid * 2 * d);
  w.pos_partial = cv.take((size_t)kPosSplits * seq_lp(c) * d);
  for (int i = 0; i < 5; ++i) {
    int M, N;
    wg_shape(c, i, &M, &N);
    const bool tc = gemm_engine(c.precision) != 0;
    w.splits_T[i] = gemm_pick_splits(M, N, T, tc);
    w.splits_B[i] = gemm_pick_splits(M, N, (int64_t)B, tc);
    const int s = w.splits_T[i] > w.splits_B[i] ? w.splits_T[i] : w.splits_B[i];
    w.gpart[i] = cv.take((size_t)s * (M + 1) * N);
  }
  if (gemm_tf32(c.precision)) {
    const int di = (int)d, dffi = (int)dff;
    w.tf_pack = cv.take(3 * d * d + 64);
    const int wmax = dffi > 3 * di ? dffi : 3 * di;
    w.tf_colsum = cv.take(tf32_colsum_scratch_bytes(wmax) / sizeof(float));
    size_t pb = 0;
    const int64_t rows[2] = {T, (int64_t)B};
    for (int r = 0; r < 2; ++r) {
      const size_t cand[3] = {tf32_wgrad_partial_bytes(rows[r], dffi, di), tf32_wgrad_partial_bytes(rows[r], 3 * di, di),
                              tf32_wgrad_partial_bytes(rows[r], 2 * di, di)};
      for (int k = 0; k < 3; ++k) pb = cand[k] > pb ? cand[k] : pb;
    }
    w.tf_wgrad = cv.take(pb / sizeof(float));
  }
  if (out) *out = w;
  return cv.off + 256;
}

// dW (+)= act^T grad over `rows` rows, db (+)= column sums of grad
inline void wgrad_prob(GemmProb& p, const float* act, int64_t ld_act, const float* grad, int64_t ld_grad, int64_t rows,
                       int M, int N, const dmt_dense& g, int splits, float* partial) {
  gemm_prob_init(p);
  p.n_parts = 1;
  p.part[0] = GemmPart{act, grad, ld_act, ld_grad, (int)rows, 0};
  p.M = M;
  p.N = N;
  p.transA = 1;
  p.transB = 0;
  p.C = const_cast<float*>(g.w);
  p.ldc = N;
  p.accumulate = 1;
  p.colsum = const_cast<float*>(g.b);
  p.colsum_accumulate = 1;
  p.splits = splits;
  p.partial = partial;
}

// C[rows, N] = epilogue(grad[rows, K] * W^T), W stored [N, K] (TF kernel [in = N, out = K])
inline void dgrad_prob(GemmProb& p, int n_parts, const float* const* grad, const int64_t* ld_grad,
                       const float* const* W, const int* K, int64_t rows, int N, float* C, int64_t ldc) {
  gemm_prob_init(p);
  p.n_parts = n_parts;
  for (int i = 0; i < n_parts; ++i) p.part[i] = GemmPart{grad[i], W[i], ld_grad[i], (int64_t)K[i], K[i], 0};
  p.M = (int)rows;
  p.N = N;
  p.transA = 0;
  p.transB = 1;
  p.C = C;
  p.ldc = ldc;
}


// ---- DMT_PRECISION_TF32 helpers ---------------------------------------------------------------------
// C[rows, N] (+)= mask((grad[rows, K] Bt[N, K]^T + addend) * alpha): dX of a dense layer; Bt = the TF kernel [N, K]
// itself (rows = inputs) or a packed concatenation of several
inline int tf32_dgrad(const float* grad, int64_t ld_grad, int K, const float* Bt, int64_t ldb, int N, int64_t rows,
                      float* C, int64_t ldc, const float* addend, int64_t ld_add, const float* mask, int64_t ld_mask,
                      float alpha, bool accumulate, cudaStream_t st) {
  Tf32Rows p{};
  p.A = grad; p.lda = ld_grad; p.Bt = Bt; p.ldb = ldb; p.M = rows; p.N = N; p.K = K; p.C = C; p.ldc = ldc;
  p.addend = addend; p.ld_add = ld_add; p.mask = mask; p.ld_mask = ld_mask; p.alpha = alpha;
  p.accumulate = accumulate ? 1 : 0;
  return tf32_rows(p, st);
}

// dW_i (+)= act^T grad_i for the column blocks grad_i = grad[:, i*n : (i+1)*n] (n_w weight tensors [K_act, n] that share
// the activation), db_i (+)= column sums of grad_i.  The wider matrix is the row operand of the contraction.
inline int tf32_wgrads(const float* act, int64_t ld_act, int k_act, const float* grad, int64_t ld_grad, int n,
                       int n_w, const dmt_dense* const* g, int64_t rows, const BwdWs& ws, cudaStream_t st) {
  int rc;
  Tf32Wgrad p{};
  p.T = rows;
  p.partial = ws.tf_wgrad;
  p.accumulate = 1;
  if (n_w == 1 && k_act > n) {       // e.g. W2 [dff, d]: D[m = act column][n = grad column] = dW as stored
    p.P = act; p.ldp = ld_act; p.MA = k_act;
    p.Q = grad; p.ldq = ld_grad; p.NB = n;
    p.seg[0] = Tf32WgradSeg{const_cast<float*>(g[0]->w), (int64_t)n, 0, k_act, nullptr};
    p.n_seg = 1;
    p.transposed = 0;
    if ((rc = tf32_wgrad(p, st))) return rc;
    return tf32_colsum(grad, ld_grad, rows, n, const_cast<float*>(g[0]->b), 1, ws.tf_colsum, st);
  }
  // D[m = grad column][n = act column] = dW^T, one segment per weight tensor; the bias gradients (column sums of the
  // gradient matrix = the row operand) ride along inside the same kernel
  p.P = grad; p.ldp = ld_grad; p.MA = n * n_w;
  p.Q = act; p.ldq = ld_act; p.NB = k_act;
  for (int i = 0; i < n_w; ++i)
    p.seg[i] = Tf32WgradSeg{const_cast<float*>(g[i]->w), (int64_t)n, i * n, (i + 1) * n, const_cast<float*>(g[i]->b)};
  p.n_seg = n_w;
  p.transposed = 1;
  return tf32_wgrad(p, st);
}

// Bt[n][i*K + k] = W_i[n][k]: the [N, K] kernels of several projections side by side (dX = sum_i grad_i W_i^T as ONE
// GEMM over the concatenated gradient columns)
inline int tf32_pack_cat(const dmt_dense* const* w, int n_w, int N, int K, float* Bt, cudaStream_t st) {
  Tf32PackMat m[4];
  for (int i = 0; i < n_w; ++i) m[i] = Tf32PackMat{w[i]->w, (int64_t)K, N, K, 0, 0, i * K};
  return tf32_pack(m, n_w, Bt, (int64_t)n_w * K, nullptr, 0, nullptr, st);
}

}  // namespace

size_t seq_bwd_workspace_bytes(const dmt_seq_cfg* cfg, int64_t T) { return bwd_carve(*cfg, T, nullptr, nullptr); }

int seq_bwd_launch(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, int64_t T,
                   const SeqSaved& sv, const float* d_out, int64_t d_out_ld, const dmt_seq_grads* g, float* d_tokens,
                   float* d_target, void* ws_base, cudaStream_t st) {
  const dmt_seq_cfg& c = *cfg;
  DMT_REQUIRE(T < (1ll << 31), DMT_ERR_UNSUPPORTED_SHAPE, "dmt_seq_encode_bwd: %lld tokens", (long long)T);
  BwdWs ws;
  bwd_carve(c, T, ws_base, &ws);
  const int d = c.d_model, dff = c.d_ff, B = c.batch, H = c.num_heads, LP = seq_lp(c);
  const float sqrt_d = sqrtf((float)d);
  const int use_tc = gemm_engine(c.precision);   // GEMMs on tcgen05 (bf16 operands, fp32 accumulate)
  const bool tf = gemm_tf32(c.precision);        // per-token GEMMs on the TMA-fed tf32 engine (gemm_tf32.cu)
  if (tf)
    DMT_REQUIRE(d % 16 == 0 && dff % 16 == 0, DMT_ERR_UNSUPPORTED_SHAPE,
                "DMT_PRECISION_TF32 needs d_model and d_ff that are multiples of 16 (got %d, %d)", d, dff);
  const int32_t* offsets = in->offsets[c.n_feats - 1];
  int rc;

  // ------------------------------------------------------------------ decoder blocks, last to first
  const float* dcur = d_out;       // gradient of the current block's output [B, d]
  int64_t dcur_ld = d_out_ld;
  bool dmem_written = false;
  float* dmem = ws.dh[0];          // gradient of the encoder memory [T, d]
  for (int blk = c.n_dec_blocks - 1; blk >= 0; --blk) {
    const dmt_attn_weights& aw = w->dec_attn[blk];
    const dmt_ff_weights& fw = w->ff[blk];
    const dmt_attn_weights& ag = g->dec_attn[blk];
    const dmt_ff_weights& fg = g->ff[blk];
    // FF LayerNorm
    rc = ln_bwd_launch(sv.z2d[blk], d, dcur, dcur_ld, fw.ln.gamma, ws.dz2d, d, B, d, ws.ln_partial, ws.ln_grid,
                       const_cast<float*>(fg.ln.gamma), const_cast<float*>(fg.ln.beta), st);
    if (rc) return rc;
    if (tf) {   // dF1 = (dZ2 W2^T) * (F1 > 0) ;  dA = dZ2 + dF1 W1^T   (the TF kernels are the K-major operands as stored)
      if ((rc = tf32_dgrad(ws.dz2d, d, d, fw.w2.w, d, dff, B, ws.df1d, dff, nullptr, 0, sv.f1d[blk], dff, 1.0f, false, st)))
        return rc;
      if ((rc = tf32_dgrad(ws.df1d, dff, dff, fw.w1.w, dff, d, B, ws.dad, d, ws.dz2d, d, nullptr, 0, 1.0f, false, st)))
        return rc;
    } else {
      {   // dF1 = (dZ2 W2^T) * (F1 > 0)
        GemmGroup grp{};
        grp.use_tc = use_tc;
        const float* gr[1] = {ws.dz2d};
        const int64_t lg[1] = {d};
        const float* W[1] = {fw.w2.w};
        const int K[1] = {d};
        dgrad_prob(grp.p[0], 1, gr, lg, W, K, B, dff, ws.df1d, dff);
        grp.p[0].mask = sv.f1d[blk];
        grp.p[0].ld_mask = dff;
        grp.n = 1;
        if ((rc = gemm_group_launch(grp, st))) return rc;
      }
      {   // dA = dZ2 + dF1 W1^T
        GemmGroup grp{};
        grp.use_tc = use_tc;
        const float* gr[1] = {ws.df1d};
        const int64_t lg[1] = {dff};
        const float* W[1] = {fw.w1.w};
        const int K[1] = {dff};
        dgrad_prob(grp.p[0], 1, gr, lg, W, K, B, d, ws.dad, d);
        grp.p[0].addend = ws.dz2d;
        grp.p[0].ld_add = d;
        grp.n = 1;
        if ((rc = gemm_group_launch(grp, st))) return rc;
      }
    }
    // attention LayerNorm: input z1d = context + block input
    rc = ln_bwd_launch(sv.z1d[blk], d, ws.dad, d, aw.ln.gamma, ws.dz1d, d, B, d, ws.ln_partial, ws.ln_grid,
                       const_cast<float*>(ag.ln.gamma), const_cast<float*>(ag.ln.beta), st);
    if (rc) return rc;
    {
      DecAttnBwdArgs a{sv.qd[blk], sv.kvd[blk], sv.pd[blk], ws.dz1d, ws.dqd, ws.dkvd, offsets, d, H, LP,
                       Dropout(c.dropout_rate, c.dropout_seed, kSiteVanillaProbs + blk)};
      const size_t smem = ((size_t)LP * (2 * d + 1) + 2 * d + 2 * H * LP + H + 8) * sizeof(float);
      cudaError_t e = cudaFuncSetAttribute(dec_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dec_attn_bwd_kernel)");
      dec_attn_bwd_kernel<<<B, kDecThreads, smem, st>>>(a);
      DMT_CUDA_LAUNCH_CHECK("dec_attn_bwd_kernel");
    }
    float* dnext = ws.dd[blk & 1];
    if (tf) {
      // dD_in = dZ1 + dQd Wq^T (x sqrt(d) into d_target for the first block) ; dMemory (+)= [dKd | dVd] [Wk | Wv]^T
      if ((rc = tf32_dgrad(ws.dqd, d, d, aw.q.w, d, d, B, blk == 0 ? d_target : dnext, d, ws.dz1d, d, nullptr, 0,
                           blk == 0 ? sqrt_d : 1.0f, false, st)))
        return rc;
      if (T > 0) {
        const dmt_dense* kv[2] = {&aw.k, &aw.v};
        if ((rc = tf32_pack_cat(kv, 2, d, d, ws.tf_pack, st))) return rc;
        if ((rc = tf32_dgrad(ws.dkvd, 2 * d, 2 * d, ws.tf_pack, 2 * d, d, T, dmem, d, nullptr, 0, nullptr, 0, 1.0f,
                             dmem_written, st)))
          return rc;
      }
      dmem_written = true;
      // weight gradients of this block
      const dmt_dense* g1[1] = {&fg.w1};
      const dmt_dense* g2[1] = {&fg.w2};
      const dmt_dense* gq[1] = {&ag.q};
      if ((rc = tf32_wgrads(sv.ad[blk], d, d, ws.df1d, dff, dff, 1, g1, B, ws, st))) return rc;
      if ((rc = tf32_wgrads(sv.f1d[blk], dff, dff, ws.dz2d, d, d, 1, g2, B, ws, st))) return rc;
      if ((rc = tf32_wgrads(sv.din[blk], d, d, ws.dqd, d, d, 1, gq, B, ws, st))) return rc;
      if (T > 0) {
        const dmt_dense* gkv[2] = {&ag.k, &ag.v};
        if ((rc = tf32_wgrads(sv.hin[c.n_enc_blocks], d, d, ws.dkvd, 2 * d, d, 2, gkv, T, ws, st))) return rc;
      }
    } else {
      {   // dD_in = dZ1 + dQd Wq^T (x sqrt(d) into d_target for the first block) ; dMemory (+)= dKd Wk^T + dVd Wv^T
        GemmGroup grp{};
        grp.use_tc = use_tc;
        const float* gr[1] = {ws.dqd};
        const int64_t lg[1] = {d};
        const float* W[1] = {aw.q.w};
        const int K[1] = {d};
        dgrad_prob(grp.p[0], 1, gr, lg, W, K, B, d, blk == 0 ? d_target : dnext, d);
        grp.p[0].addend = ws.dz1d;
        grp.p[0].ld_add = d;
        if (blk == 0) grp.p[0].alpha = sqrt_d;
        grp.n = 1;
        if (T > 0) {
          const float* gr2[2] = {ws.dkvd, ws.dkvd + d};
          const int64_t lg2[2] = {2 * d, 2 * d};
          const float* W2[2] = {aw.k.w, aw.v.w};
          const int K2[2] = {d, d};
          dgrad_prob(grp.p[1], 2, gr2, lg2, W2, K2, T, d, dmem, d);
          grp.p[1].accumulate = dmem_written ? 1 : 0;
          grp.n = 2;
        }
        if ((rc = gemm_group_launch(grp, st))) return rc;
        dmem_written = true;
      }
      {   // weight gradients of this block
        GemmGroup grp{};
        grp.use_tc = use_tc;
        int n = 0;
        wgrad_prob(grp.p[n++], sv.ad[blk], d, ws.df1d, dff, B, d, dff, fg.w1, ws.splits_B[0], ws.gpart[0]);
        wgrad_prob(grp.p[n++], sv.f1d[blk], dff, ws.dz2d, d, B, dff, d, fg.w2, ws.splits_B[1], ws.gpart[1]);
        wgrad_prob(grp.p[n++], sv.din[blk], d, ws.dqd, d, B, d, d, ag.q, ws.splits_B[2], ws.gpart[2]);
        if (T > 0) {
          const float* mem = sv.hin[c.n_enc_blocks];
          wgrad_prob(grp.p[n++], mem, d, ws.dkvd, 2 * d, T, d, d, ag.k, ws.splits_T[3], ws.gpart[3]);
          wgrad_prob(grp.p[n++], mem, d, ws.dkvd + d, 2 * d, T, d, d, ag.v, ws.splits_T[4], ws.gpart[4]);
        }
        grp.n = n;
        if ((rc = gemm_group_launch(grp, st))) return rc;
      }
    }
    dcur = dnext;
    dcur_ld = d;
  }
  if (c.n_dec_blocks == 0) {   // interest = target * sqrt(d): no decoder variables, memory unused
    scale_copy_kernel<<<(unsigned)(((int64_t)B * d + 255) / 256), 256, 0, st>>>(d_out, d_out_ld, d_target, d, B, d,
                                                                               sqrt_d);
    DMT_CUDA_LAUNCH_CHECK("scale_copy_kernel");
  }
  if (c.dropout_rate > 0.f) {
    dropout_bwd_kernel<<<(unsigned)(((int64_t)B * d + 255) / 256), 256, 0, st>>>(
        d_target, (int64_t)B * d, Dropout(c.dropout_rate, c.dropout_seed, kSiteDecIn));
    DMT_CUDA_LAUNCH_CHECK("dropout_bwd_kernel");
  }
  if (T == 0) return DMT_OK;
  if (!dmem_written) {
    cudaError_t e = cudaMemsetAsync(dmem, 0, (size_t)T * d * sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(dmem)");
  }

  // ------------------------------------------------------------------ encoder blocks, last to first
  float* dh = dmem;
  int ping = 0;
  for (int blk = c.n_enc_blocks - 1; blk >= 0; --blk) {
    const dmt_attn_weights& aw = w->enc_attn[blk];
    const dmt_ff_weights& fw = w->ff[blk];
    const dmt_attn_weights& ag = g->enc_attn[blk];
    const dmt_ff_weights& fg = g->ff[blk];
    rc = ln_bwd_launch(sv.z2[blk], d, dh, d, fw.ln.gamma, ws.dz2, d, T, d, ws.ln_partial, ws.ln_grid,
                       const_cast<float*>(fg.ln.gamma), const_cast<float*>(fg.ln.beta), st);
    if (rc) return rc;
    if (tf) {
      if ((rc = tf32_dgrad(ws.dz2, d, d, fw.w2.w, d, dff, T, ws.df1, dff, nullptr, 0, sv.f1[blk], dff, 1.0f, false, st)))
        return rc;
      if ((rc = tf32_dgrad(ws.df1, dff, dff, fw.w1.w, dff, d, T, ws.da, d, ws.dz2, d, nullptr, 0, 1.0f, false, st)))
        return rc;
    } else {
      {
        GemmGroup grp{};
        grp.use_tc = use_tc;
        const float* gr[1] = {ws.dz2};
        const int64_t lg[1] = {d};
        const float* W[1] = {fw.w2.w};
        const int K[1] = {d};
        dgrad_prob(grp.p[0], 1, gr, lg, W, K, T, dff, ws.df1, dff);
        grp.p[0].mask = sv.f1[blk];
        grp.p[0].ld_mask = dff;
        grp.n = 1;
        if ((rc = gemm_group_launch(grp, st))) return rc;
      }
      {
        GemmGroup grp{};
        grp.use_tc = use_tc;
        const float* gr[1] = {ws.df1};
        const int64_t lg[1] = {dff};
        const float* W[1] = {fw.w1.w};
        const int K[1] = {dff};
        dgrad_prob(grp.p[0], 1, gr, lg, W, K, T, d, ws.da, d);
        grp.p[0].addend = ws.dz2;
        grp.p[0].ld_add = d;
        grp.n = 1;
        if ((rc = gemm_group_launch(grp, st))) return rc;
      }
    }
    rc = ln_bwd_launch(sv.z1[blk], d, ws.da, d, aw.ln.gamma, ws.dz1, d, T, d, ws.ln_partial, ws.ln_grid,
                       const_cast<float*>(ag.ln.gamma), const_cast<float*>(ag.ln.beta), st);
    if (rc) return rc;
    {
      AttnBwdArgs a{sv.qkv[blk], ws.dz1, ws.dqkv, offsets, d, H, LP,
                    Dropout(c.dropout_rate, c.dropout_seed, kSiteSelfProbs + blk)};
      const size_t smem = ((size_t)4 * LP * (d + 1) + 2 * LP * (LP + 1)) * sizeof(float);
      DMT_REQUIRE(smem <= 227 * 1024, DMT_ERR_UNSUPPORTED_SHAPE, "dmt_seq_encode_bwd: attention tile needs %zu B", smem);
      cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn_bwd_kernel)");
      attn_bwd_kernel<<<B, kAttnThreads, smem, st>>>(a);
      DMT_CUDA_LAUNCH_CHECK("attn_bwd_kernel");
    }
    ping ^= 1;
    float* dhin = blk == 0 ? d_tokens : ws.dh[ping];
    if (tf) {
      // dH_in = dZ1 + [dQ | dK | dV] [Wq | Wk | Wv]^T   (x sqrt(d) for the first block: row gradients)
      const dmt_dense* qkvw[3] = {&aw.q, &aw.k, &aw.v};
      if ((rc = tf32_pack_cat(qkvw, 3, d, d, ws.tf_pack, st))) return rc;
      if ((rc = tf32_dgrad(ws.dqkv, 3 * d, 3 * d, ws.tf_pack, 3 * d, d, T, dhin, d, ws.dz1, d, nullptr, 0,
                           blk == 0 ? sqrt_d : 1.0f, false, st)))
        return rc;
      const dmt_dense* g1[1] = {&fg.w1};
      const dmt_dense* g2[1] = {&fg.w2};
      const dmt_dense* gqkv[3] = {&ag.q, &ag.k, &ag.v};
      if ((rc = tf32_wgrads(sv.a[blk], d, d, ws.df1, dff, dff, 1, g1, T, ws, st))) return rc;
      if ((rc = tf32_wgrads(sv.f1[blk], dff, dff, ws.dz2, d, d, 1, g2, T, ws, st))) return rc;
      if ((rc = tf32_wgrads(sv.hin[blk], d, d, ws.dqkv, 3 * d, d, 3, gqkv, T, ws, st))) return rc;
    } else {
      {   // dH_in = dZ1 + dQ Wq^T + dK Wk^T + dV Wv^T   (x sqrt(d) for the first block: row gradients)
        GemmGroup grp{};
        grp.use_tc = use_tc;
        const float* gr[3] = {ws.dqkv, ws.dqkv + d, ws.dqkv + 2 * d};
        const int64_t lg[3] = {3 * d, 3 * d, 3 * d};
        const float* W[3] = {aw.q.w, aw.k.w, aw.v.w};
        const int K[3] = {d, d, d};
        dgrad_prob(grp.p[0], 3, gr, lg, W, K, T, d, dhin, d);
        grp.p[0].addend = ws.dz1;
        grp.p[0].ld_add = d;
        if (blk == 0) grp.p[0].alpha = sqrt_d;
        grp.n = 1;
        if ((rc = gemm_group_launch(grp, st))) return rc;
      }
      {
        GemmGroup grp{};
        grp.use_tc = use_tc;
        int n = 0;
        wgrad_prob(grp.p[n++], sv.a[blk], d, ws.df1, dff, T, d, dff, fg.w1, ws.splits_T[0], ws.gpart[0]);
        wgrad_prob(grp.p[n++], sv.f1[blk], dff, ws.dz2, d, T, dff, d, fg.w2, ws.splits_T[1], ws.gpart[1]);
        wgrad_prob(grp.p[n++], sv.hin[blk], d, ws.dqkv, 3 * d, T, d, d, ag.q, ws.splits_T[2], ws.gpart[2]);
        wgrad_prob(grp.p[n++], sv.hin[blk], d, ws.dqkv + d, 3 * d, T, d, d, ag.k, ws.splits_T[3], ws.gpart[3]);
        wgrad_prob(grp.p[n++], sv.hin[blk], d, ws.dqkv + 2 * d, 3 * d, T, d, d, ag.v, ws.splits_T[4], ws.gpart[4]);
        grp.n = n;
        if ((rc = gemm_group_launch(grp, st))) return rc;
      }
    }
    dh = dhin;
  }
  if (c.n_enc_blocks == 0) {   // memory = encoder input
    scale_copy_kernel<<<(unsigned)((T * d + 255) / 256), 256, 0, st>>>(dmem, d, d_tokens, d, T, d, sqrt_d);
    DMT_CUDA_LAUNCH_CHECK("scale_copy_kernel");
  }
  if (c.dropout_rate > 0.f) {   // H0 = dropout(X sqrt(d) + P), D0 = dropout(q sqrt(d)): gradients pick up the masks
    dropout_bwd_kernel<<<(unsigned)((T * d + 255) / 256), 256, 0, st>>>(d_tokens, T * d,
                                                                      Dropout(c.dropout_rate, c.dropout_seed, kSiteEncIn));
    DMT_CUDA_LAUNCH_CHECK("dropout_bwd_kernel");
  }
  {
    PosGradArgs a{d_tokens, offsets, ws.pos_partial, const_cast<float*>(g->pos), 1.0f / sqrt_d, B, d, LP};
    DMT_REQUIRE(d <= 256, DMT_ERR_UNSUPPORTED_SHAPE, "dmt_seq_encode_bwd: d_model %d > 256", d);
    pos_grad_kernel<<<dim3(LP, kPosSplits), 256, 0, st>>>(a);
    DMT_CUDA_LAUNCH_CHECK("pos_grad_kernel");
    pos_grad_reduce_kernel<<<(LP * d + 255) / 256, 256, 0, st>>>(a);
    DMT_CUDA_LAUNCH_CHECK("pos_grad_reduce_kernel");
  }
  return DMT_OK;
}

}  // namespace dmt
