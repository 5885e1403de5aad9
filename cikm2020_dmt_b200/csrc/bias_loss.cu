// dmt_bias_loss_fwd: neighbouring-bias tower (A11) + unbiased two-task loss and its logit
// gradients (A12), fused: one thread per sample, the tiny MLP weights staged in shared memory.
#include "dmt_common.cuh"
#include "dropout.cuh"

namespace dmt {

constexpr int kMaxBiasWidth = 64;
constexpr float kKerasEps = 1e-7f;   // keras.backend.epsilon(), inference_mlp.py:167

struct BiasLossArgs {
  dmt_bias_loss_cfg cfg;
  dmt_bias_weights w;
  const float* bias_in;
  int64_t bias_ld;
  const float* logits;   // [2][B]
  const float* mask;     // [B,5] or null
  float* y_bias;         // [B]
  float* probs;          // [2][B] or null
  float* dlogits;        // [3][B] or null
  float* per_sample;     // [B] weighted loss terms (scratch) or null
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// Keras sparse_categorical_crossentropy([1-p, p], y, from_logits=False) and d/dp
// (inference_mlp.py:162-168): clip to [eps, 1-eps], log, softmax cross-entropy.
__device__ __forceinline__ void xent_clip(float p, int y, float& xe, float& dxe_dp) {
  const float q = 1.0f - p;
  const float c0 = fminf(fmaxf(q, kKerasEps), 1.0f - kKerasEps);
  const float c1 = fminf(fmaxf(p, kKerasEps), 1.0f - kKerasEps);
  const float d0 = (q > kKerasEps && q < 1.0f - kKerasEps) ? -1.0f : 0.0f;   // dc0/dp
  const float d1 = (p > kKerasEps && p < 1.0f - kKerasEps) ? 1.0f : 0.0f;    // dc1/dp
  const float s = c0 + c1;
  const float cy = y ? c1 : c0;
  const float dy = y ? d1 : d0;
  xe = logf(s) - logf(cy);
  dxe_dp = (d0 + d1) / s - dy / cy;
}

// kLanes lanes per sample: lane j of a sample computes the units n = j, j + kLanes, ... of every layer, the
// activations of the sample live in a private shared-memory row (stride kMaxBiasWidth + 1: conflict-free across the
// samples of a warp).  Lane 0 of the sample then evaluates the loss terms.
constexpr int kLanes = 4;
constexpr int kActLd = kMaxBiasWidth + 1;
__global__ void __launch_bounds__(128) bias_loss_kernel(const __grid_constant__ BiasLossArgs a) {
  extern __shared__ float wsm[];
  // stage all layer weights: [in,out] kernels then biases, in layer order
  const int nl = a.cfg.n_hidden + 1;
  int in_dim = a.cfg.in_dim, total = 0;
  for (int l = 0; l < nl; ++l) {
    const int units = l < a.cfg.n_hidden ? a.cfg.units[l] : 1;
    const int nw = in_dim * units;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) wsm[total + i] = __ldg(a.w.layer[l].w + i);
    for (int i = threadIdx.x; i < units; i += blockDim.x) wsm[total + nw + i] = __ldg(a.w.layer[l].b + i);
    total += nw + units;
    in_dim = units;
  }
  const int B = a.cfg.batch;
  const int sl = threadIdx.x / kLanes, sub = threadIdx.x % kLanes;       // sample slot in the CTA, lane in the sample
  const int b = blockIdx.x * (blockDim.x / kLanes) + sl;
  const int bc = b < B ? b : B - 1;                                       // out-of-range slots redo the last sample
  float* cur = wsm + ((total + 3) & ~3) + sl * 2 * kActLd;
  float* nxt = cur + kActLd;
  in_dim = a.cfg.n_hidden < 0 ? 1 : a.cfg.in_dim;   // n_hidden < 0: bias_in already holds y_bias
  for (int k = sub; k < in_dim; k += kLanes) cur[k] = __ldg(a.bias_in + (int64_t)bc * a.bias_ld + k);
  __syncthreads();
  int base = 0;
  for (int l = 0; l < nl; ++l) {
    const int units = l < a.cfg.n_hidden ? a.cfg.units[l] : 1;
    const float* W = wsm + base;
    const float* bb = W + in_dim * units;
    for (int n = sub; n < units; n += kLanes) {
      float a0 = 0.f, a1 = 0.f;
      int k = 0;
      for (; k + 2 <= in_dim; k += 2) {
        a0 = fmaf(cur[k], W[k * units + n], a0);
        a1 = fmaf(cur[k + 1], W[(k + 1) * units + n], a1);
      }
      if (k < in_dim) a0 = fmaf(cur[k], W[k * units + n], a0);
      float acc = (a0 + a1) + bb[n];
      acc = (l < a.cfg.n_hidden) ? fmaxf(acc, 0.f) : acc;      // relu hidden, identity output (:263-287)
      if (l < a.cfg.n_hidden && a.cfg.dropout_rate[l] > 0.f)   // training mode (:272,280)
        acc *= Dropout(a.cfg.dropout_rate[l], a.cfg.dropout_seed, kSiteBias + l).mult((uint32_t)(bc * units + n));
      nxt[n] = acc;
    }
    __syncwarp();
    float* tmp = cur; cur = nxt; nxt = tmp;
    base += in_dim * units + units;
    in_dim = units;
  }
  if (sub != 0 || b >= B) return;
  const float yb = cur[0];
  a.y_bias[b] = yb;
  if (!a.probs && !a.mask) return;     // inference: the tower alone (the logits may still be in flight on another stream)

  const float lc = __ldg(a.logits + b), lo = __ldg(a.logits + B + b);
  const float sc = sigmoidf_(lc), so = sigmoidf_(lo), sb = sigmoidf_(yb);
  float p_ctr, p_cvr;
  if (a.cfg.two_head_multiply) {       // run_dnn.py:92-94
    p_ctr = sc * sb;
    p_cvr = so * sb;
  } else {                              // run_dnn.py:96-98
    p_ctr = sigmoidf_(lc + yb);
    p_cvr = sigmoidf_(lo + yb);
  }
  if (a.probs) {
    a.probs[b] = p_ctr;
    a.probs[B + b] = p_cvr;
  }
  if (!a.mask) return;
  float m[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) m[c] = __ldg(a.mask + (int64_t)b * 5 + c);
  const int y_clk = (int)(m[1] + m[2] + m[3] + m[4]);   // inference_mlp.py:192
  const int y_ord = (int)(m[3] + m[4]);                 // inference_mlp.py:193
  float w_clk = 0.f, w_ord = 0.f;
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    w_clk = fmaf(m[c], a.cfg.weight_ctr[c], w_clk);
    w_ord = fmaf(m[c], a.cfg.weight_ecvr[c], w_ord);
  }
  float xe_c, dc, xe_cr, dcr, xe_o, dor, xe_or, dorr;
  xent_clip(p_ctr, y_clk, xe_c, dc);
  xent_clip(sc, y_clk, xe_cr, dcr);
  xent_clip(p_cvr, y_ord, xe_o, dor);
  xent_clip(so, y_ord, xe_or, dorr);
  const float rel = a.cfg.ctr_rel ? 1.0f : 0.0f;
  const float invB = 1.0f / (float)B;
  const float kc = a.cfg.loss_weight[0] * w_clk * invB, ko = a.cfg.loss_weight[1] * w_ord * invB;
  if (a.per_sample) a.per_sample[b] = kc * (xe_c + rel * xe_cr) + ko * (xe_o + rel * xe_or);
  if (a.dlogits) {
    float g_click, g_order, g_bias;
    if (a.cfg.two_head_multiply) {
      g_click = kc * (dc * sc * (1.f - sc) * sb + rel * dcr * sc * (1.f - sc));
      g_order = ko * (dor * so * (1.f - so) * sb + rel * dorr * so * (1.f - so));
      g_bias = (kc * dc * sc + ko * dor * so) * sb * (1.f - sb);
    } else {
      const float gc = kc * dc * p_ctr * (1.f - p_ctr), go = ko * dor * p_cvr * (1.f - p_cvr);
      g_click = gc + kc * rel * dcr * sc * (1.f - sc);
      g_order = go + ko * rel * dorr * so * (1.f - so);
      g_bias = gc + go;
    }
    a.dlogits[b] = g_click;
    a.dlogits[B + b] = g_order;
    a.dlogits[2 * (int64_t)B + b] = g_bias;
  }
}

// Fixed-order sum of the per-sample terms -> deterministic loss (no atomics).
__global__ void __launch_bounds__(1024) loss_reduce_kernel(const float* __restrict__ per_sample, int n,
                                                           float* __restrict__ loss) {
  __shared__ float part[1024];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) s += per_sample[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = part[0];
}

}  // namespace dmt

extern "C" {

size_t dmt_loss_scratch_bytes(int32_t batch) { return (size_t)(batch > 0 ? batch : 0) * sizeof(float) + 256; }

int dmt_bias_loss_fwd(const dmt_bias_loss_cfg* cfg, const dmt_bias_weights* w, const float* bias_in, int64_t bias_ld,
                      const float* logits, const float* mask, float* y_bias, float* probs, float* loss,
                      float* dlogits, void* loss_scratch, void* stream) {
  DMT_REQUIRE(cfg && w && bias_in && logits && y_bias, DMT_ERR_INVALID_ARGUMENT, "dmt_bias_loss_fwd: null pointer");
  DMT_REQUIRE(cfg->n_hidden <= DMT_MAX_LAYERS && cfg->in_dim > 0 &&
                  cfg->in_dim <= dmt::kMaxBiasWidth && bias_ld >= cfg->in_dim,
              DMT_ERR_INVALID_ARGUMENT, "dmt_bias_loss_fwd: in_dim=%d n_hidden=%d", cfg->in_dim, cfg->n_hidden);
  DMT_REQUIRE(!(loss || dlogits) || mask, DMT_ERR_INVALID_ARGUMENT, "dmt_bias_loss_fwd: loss needs mask");
  DMT_REQUIRE(!loss || loss_scratch, DMT_ERR_WORKSPACE_TOO_SMALL, "dmt_bias_loss_fwd: loss needs loss_scratch");
  int in_dim = cfg->in_dim;
  size_t wfloats = 0;
  for (int l = 0; cfg->n_hidden >= 0 && l <= cfg->n_hidden; ++l) {
    const int units = l < cfg->n_hidden ? cfg->units[l] : 1;
    DMT_REQUIRE(units > 0 && units <= dmt::kMaxBiasWidth, DMT_ERR_UNSUPPORTED_SHAPE,
                "dmt_bias_loss_fwd: hidden_units_bias[%d]=%d (max %d)", l, units, dmt::kMaxBiasWidth);
    wfloats += (size_t)in_dim * units + units;
    in_dim = units;
  }
  if (cfg->batch <= 0) return DMT_OK;
  dmt::BiasLossArgs a;
  a.cfg = *cfg;
  a.w = *w;
  a.bias_in = bias_in;
  a.bias_ld = bias_ld;
  a.logits = logits;
  a.mask = mask;
  a.y_bias = y_bias;
  a.probs = probs;
  a.dlogits = dlogits;
  a.per_sample = loss ? (float*)loss_scratch : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int kSamplesPerCta = 128 / dmt::kLanes;
  const size_t smem = (((wfloats + 3) & ~(size_t)3) + (size_t)kSamplesPerCta * 2 * dmt::kActLd) * sizeof(float);
  dmt::bias_loss_kernel<<<(cfg->batch + kSamplesPerCta - 1) / kSamplesPerCta, 128, smem, st>>>(a);
  DMT_CUDA_LAUNCH_CHECK("bias_loss_kernel");
  if (loss) {
    dmt::loss_reduce_kernel<<<1, 1024, 0, st>>>((const float*)loss_scratch, cfg->batch, loss);
    DMT_CUDA_LAUNCH_CHECK("loss_reduce_kernel");
  }
  return DMT_OK;
}

}  // extern "C"
