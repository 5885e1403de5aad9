// dmt_mmoe_bwd (fp32): backward of expert_gate + build_tower (mmoe_transformer_unbias.py:63-126,293-310).
//
//   mmoe_head_bwd_kernel   one warp per sample: recompute gates / mixture / tower activations, then
//                          d(logit) -> tower -> mixture -> {d(last expert layer) (ReLU-masked), d(gate logits)}
//   grouped GEMMs          expert layers last to first: dW = in^T dH (split-K over the batch), dH_prev =
//                          (dH W^T) * relu'(H_prev); the first layer and both gates contract into dx in ONE
//                          multi-part problem, so dx is written once.
#include "dropout.cuh"
#include "gemm_f32.cuh"
#include "gemm_tf32.cuh"
#include "seq_train.cuh"   // Carver

namespace dmt {

namespace {

constexpr int kHeadWarps = 8;

struct HeadBwdArgs {
  dmt_mmoe_cfg cfg;
  dmt_dense gate[DMT_MAX_TASKS];
  dmt_dense tower[DMT_MAX_TASKS][DMT_MAX_LAYERS];
  dmt_dense tower_out[DMT_MAX_TASKS];
  const float* x;
  int64_t x_ld;
  const float* h_last;      // [E][B][Hd]
  const float* dlogits;     // [T][B]
  float* zt;                // [T][B][Hd]           mixture (tower input)
  float* act[DMT_MAX_LAYERS];    // [T][B][units_l]   tower activations
  float* dpre[DMT_MAX_LAYERS];   // [T][B][units_l]   gradient of the tower pre-activations
  float* dh_last;           // [E][B][Hd]           gradient of the last expert layer's pre-activation
  float* dgl;               // [B][T*E]             gradient of the gate logits
  int32_t hdim, maxw;
};

__global__ void __launch_bounds__(kHeadWarps * 32) mmoe_head_bwd_kernel(const __grid_constant__ HeadBwdArgs a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kHeadWarps + warp;
  if (b >= a.cfg.batch) return;
  const int E = a.cfg.n_experts, K = a.cfg.in_dim, Hd = a.hdim, B = a.cfg.batch, NT = a.cfg.n_tasks;
  const int NL = a.cfg.n_tower_layers, W = a.maxw;
  // per-warp scratch: activations [NL+1][W] | gradient ping-pong [2][W] | dh accumulators [E][Hd]
  float* base = sm + (size_t)warp * ((NL + 3) * W + E * Hd);
  float* acts = base;
  float* gbuf = base + (NL + 1) * W;
  float* dhacc = gbuf + 2 * W;
  for (int i = lane; i < E * Hd; i += 32) dhacc[i] = 0.f;
  const float* __restrict__ xr = a.x + (int64_t)b * a.x_ld;
  for (int t = 0; t < NT; ++t) {
    // ---- forward recompute (same arithmetic as mmoe_head_kernel)
    float gl[DMT_MAX_EXPERTS];
#pragma unroll
    for (int e = 0; e < DMT_MAX_EXPERTS; ++e) gl[e] = 0.f;
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
    const float inv = 1.0f / den;
#pragma unroll
    for (int e = 0; e < DMT_MAX_EXPERTS; ++e) gl[e] *= inv;   // softmax gates
    float* z = acts;
    for (int c = lane; c < Hd; c += 32) {
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
        if (e < E) acc = fmaf(gl[e], __ldg(a.h_last + ((int64_t)e * B + b) * Hd + c), acc);
      z[c] = acc;
      a.zt[((int64_t)t * B + b) * Hd + c] = acc;
    }
    __syncwarp();
    int in_dim = Hd;
    for (int l = 0; l < NL; ++l) {
      const int units = a.cfg.tower_units[l];
      const float* __restrict__ Wt = a.tower[t][l].w;
      const float* cur = acts + l * W;
      float* nxt = acts + (l + 1) * W;
      for (int n = lane; n < units; n += 32) {
        float acc = 0.f;
        for (int k = 0; k < in_dim; ++k) acc = fmaf(cur[k], __ldg(Wt + (int64_t)k * units + n), acc);
        const float y = fmaxf(acc + __ldg(a.tower[t][l].b + n), 0.f);
        nxt[n] = y;
        a.act[l][((int64_t)t * B + b) * units + n] = y;
      }
      __syncwarp();
      in_dim = units;
    }
    // ---- backward
    const float dl = __ldg(a.dlogits + (int64_t)t * B + b);
    float* dcur = gbuf;
    float* dnxt = gbuf + W;
    for (int k = lane; k < in_dim; k += 32) dcur[k] = dl * __ldg(a.tower_out[t].w + k);
    __syncwarp();
    for (int l = NL - 1; l >= 0; --l) {
      const int units = a.cfg.tower_units[l];
      const int prev = l == 0 ? Hd : a.cfg.tower_units[l - 1];
      const float* y = acts + (l + 1) * W;
      for (int n = lane; n < units; n += 32) {
        const float dp = y[n] > 0.f ? dcur[n] : 0.f;
        dcur[n] = dp;
        a.dpre[l][((int64_t)t * B + b) * units + n] = dp;
      }
      __syncwarp();
      const float* __restrict__ Wt = a.tower[t][l].w;
      for (int k = lane; k < prev; k += 32) {
        float acc = 0.f;
        for (int n = 0; n < units; ++n) acc = fmaf(dcur[n], __ldg(Wt + (int64_t)k * units + n), acc);
        dnxt[k] = acc;
      }
      __syncwarp();
      float* tmp = dcur; dcur = dnxt; dnxt = tmp;
    }
    // dcur = d(mixture) [Hd]
    float dg[DMT_MAX_EXPERTS];
#pragma unroll
    for (int e = 0; e < DMT_MAX_EXPERTS; ++e) dg[e] = 0.f;
    for (int c = lane; c < Hd; c += 32) {
      const float dz = dcur[c];
#pragma unroll
      for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
        if (e < E) {
          const float hv = __ldg(a.h_last + ((int64_t)e * B + b) * Hd + c);
          dg[e] = fmaf(dz, hv, dg[e]);
          if (hv > 0.f) dhacc[e * Hd + c] = fmaf(gl[e], dz, dhacc[e * Hd + c]);
        }
    }
    float dot = 0.f;
#pragma unroll
    for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
      if (e < E) {
        dg[e] = warp_sum(dg[e]);
        dot = fmaf(gl[e], dg[e], dot);
      }
    if (lane == 0) {
#pragma unroll
      for (int e = 0; e < DMT_MAX_EXPERTS; ++e)
        if (e < E) a.dgl[(int64_t)b * NT * E + t * E + e] = gl[e] * (dg[e] - dot);
    }
    __syncwarp();
  }
  for (int i = lane; i < E * Hd; i += 32) {
    const int e = i / Hd, c = i - e * Hd;
    a.dh_last[((int64_t)e * B + b) * Hd + c] = dhacc[i];
  }
}

struct MmoeBwdWs {
  float* dH[DMT_MAX_LAYERS];     // [E][B][units_l]
  float* zt;
  float* act[DMT_MAX_LAYERS];
  float* dpre[DMT_MAX_LAYERS];
  float* dgl;
  float* part_expert[DMT_MAX_EXPERTS];   // split-K scratch, one per expert (largest layer)
  float* part_gate[DMT_MAX_TASKS];
  float* part_tower[DMT_MAX_TASKS][DMT_MAX_LAYERS + 1];
  float* tf_colsum;                      // DMT_PRECISION_TF32: scratch of the bias-gradient column sums
  int splits_layer[DMT_MAX_LAYERS];
  int splits_gate, splits_tower[DMT_MAX_LAYERS + 1];
};

size_t mmoe_bwd_carve(const dmt_mmoe_cfg& c, void* base, MmoeBwdWs* out) {
  Carver cv(base);
  MmoeBwdWs w{};
  const bool tc = gemm_engine(c.precision) != 0;
  const size_t B = c.batch, E = c.n_experts, T = c.n_tasks;
  const int Hd = c.units[c.n_layers - 1];
  for (int l = 0; l < c.n_layers; ++l) w.dH[l] = cv.take(E * B * c.units[l]);
  w.zt = cv.take(T * B * Hd);
  for (int l = 0; l < c.n_tower_layers; ++l) {
    w.act[l] = cv.take(T * B * c.tower_units[l]);
    w.dpre[l] = cv.take(T * B * c.tower_units[l]);
  }
  w.dgl = cv.take(B * T * E);
  size_t worst = 0;
  int in_dim = c.in_dim;
  for (int l = 0; l < c.n_layers; ++l) {
    w.splits_layer[l] = gemm_pick_splits(in_dim, c.units[l], (int64_t)B, tc);
    const size_t need = (size_t)w.splits_layer[l] * (in_dim + 1) * c.units[l];
    if (need > worst) worst = need;
    in_dim = c.units[l];
  }
  for (size_t e = 0; e < E; ++e) w.part_expert[e] = cv.take(worst);
  w.splits_gate = gemm_pick_splits(c.in_dim, (int)E, (int64_t)B, tc);
  for (size_t t = 0; t < T; ++t) w.part_gate[t] = cv.take((size_t)w.splits_gate * (c.in_dim + 1) * E);
  in_dim = Hd;
  for (int l = 0; l <= c.n_tower_layers; ++l) {
    const int units = l < c.n_tower_layers ? c.tower_units[l] : 1;
    w.splits_tower[l] = gemm_pick_splits(in_dim, units, (int64_t)B, tc);
    for (size_t t = 0; t < T; ++t) w.part_tower[t][l] = cv.take((size_t)w.splits_tower[l] * (in_dim + 1) * units);
    in_dim = units;
  }
  if (gemm_tf32(c.precision)) {
    int wmax = 4;
    for (int l = 0; l < c.n_layers; ++l) wmax = c.units[l] > wmax ? c.units[l] : wmax;
    w.tf_colsum = cv.take(tf32_colsum_scratch_bytes((wmax + 3) / 4 * 4) / sizeof(float));
  }
  if (out) *out = w;
  return cv.off + 256;
}

inline void wgrad(GemmProb& p, const float* act, int64_t ld_act, const float* grad, int64_t ld_grad, int rows, int M,
                  int N, const dmt_dense& g, int splits, float* partial) {
  gemm_prob_init(p);
  p.n_parts = 1;
  p.part[0] = GemmPart{act, grad, ld_act, ld_grad, rows, 0};
  p.M = M;
  p.N = N;
  p.transA = 1;
  p.C = const_cast<float*>(g.w);
  p.ldc = N;
  p.accumulate = 1;
  p.colsum = const_cast<float*>(g.b);
  p.colsum_accumulate = 1;
  p.splits = splits;
  p.partial = partial;
}

}  // namespace

int mmoe_bwd_launch(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const float* x, int64_t x_ld,
                    const float* fwd_ws, const float* dlogits, const dmt_mmoe_grads* g, float* dx, int64_t dx_ld,
                    int dx_col0, void* ws_base, cudaStream_t st) {
  const dmt_mmoe_cfg& c = *cfg;
  const int B = c.batch, E = c.n_experts, NT = c.n_tasks, NL = c.n_layers;
  const int Hd = c.units[NL - 1];
  MmoeBwdWs ws;
  mmoe_bwd_carve(c, ws_base, &ws);
  const int use_tc = gemm_engine(c.precision);
  // forward activations: [l][E][B][units_l]
  const float* H[DMT_MAX_LAYERS];
  {
    const float* p = fwd_ws;
    for (int l = 0; l < NL; ++l) {
      H[l] = p;
      p += (int64_t)E * B * c.units[l];
    }
  }
  int rc;
  {
    HeadBwdArgs h{};
    h.cfg = c;
    for (int t = 0; t < NT; ++t) {
      h.gate[t] = w->gate[t];
      for (int l = 0; l < c.n_tower_layers; ++l) h.tower[t][l] = w->tower[t][l];
      h.tower_out[t] = w->tower_out[t];
    }
    h.x = x;
    h.x_ld = x_ld;
    h.h_last = H[NL - 1];
    h.dlogits = dlogits;
    h.zt = ws.zt;
    for (int l = 0; l < c.n_tower_layers; ++l) {
      h.act[l] = ws.act[l];
      h.dpre[l] = ws.dpre[l];
    }
    h.dh_last = ws.dH[NL - 1];
    h.dgl = ws.dgl;
    h.hdim = Hd;
    int mx = Hd;
    for (int l = 0; l < c.n_tower_layers; ++l) mx = c.tower_units[l] > mx ? c.tower_units[l] : mx;
    h.maxw = (mx + 31) / 32 * 32;
    const size_t smem = (size_t)kHeadWarps * ((c.n_tower_layers + 3) * h.maxw + E * Hd) * sizeof(float);
    DMT_REQUIRE(smem <= 200 * 1024, DMT_ERR_UNSUPPORTED_SHAPE, "dmt_mmoe_bwd: tower scratch %zu B too large", smem);
    cudaError_t e = cudaFuncSetAttribute(mmoe_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mmoe_head_bwd_kernel)");
    mmoe_head_bwd_kernel<<<(B + kHeadWarps - 1) / kHeadWarps, kHeadWarps * 32, smem, st>>>(h);
    DMT_CUDA_LAUNCH_CHECK("mmoe_head_bwd_kernel");
  }
  // tower + gate weight gradients (contractions over the batch)
  for (int t = 0; t < NT; ++t) {
    GemmGroup grp{};
      grp.use_tc = use_tc;
    int n = 0;
    int in_dim = Hd;
    const float* in = ws.zt + (int64_t)t * B * Hd;
    for (int l = 0; l < c.n_tower_layers; ++l) {
      const int units = c.tower_units[l];
      wgrad(grp.p[n++], in, in_dim, ws.dpre[l] + (int64_t)t * B * units, units, B, in_dim, units, g->tower[t][l],
            ws.splits_tower[l], ws.part_tower[t][l]);
      in = ws.act[l] + (int64_t)t * B * units;
      in_dim = units;
    }
    wgrad(grp.p[n++], in, in_dim, dlogits + (int64_t)t * B, 1, B, in_dim, 1, g->tower_out[t],
          ws.splits_tower[c.n_tower_layers], ws.part_tower[t][c.n_tower_layers]);
    wgrad(grp.p[n++], x, x_ld, ws.dgl + t * E, (int64_t)NT * E, B, c.in_dim, E, g->gate[t], ws.splits_gate,
          ws.part_gate[t]);
    grp.n = n;
    if ((rc = gemm_group_launch(grp, st))) return rc;
  }
  // expert layers, last to first
  const bool tf = gemm_tf32(c.precision);
  bool tf_ok = tf && E <= 4;
  for (int l = 0; l < NL && tf_ok; ++l) tf_ok = c.units[l] % 4 == 0;
  for (int l = NL - 1; l >= 0; --l) {
    const int units = c.units[l];
    const int in_dim = l == 0 ? c.in_dim : c.units[l - 1];
    if (tf_ok) {
      // dW_e (+)= in_e^T dH_e: both operands MN-major straight from the row-major activations (contraction over
      // the batch), the four experts in one launch; db_e (+)= column sums of dH_e
      Tf32Gemm p{};
      p.nz = E;
      p.a_mn = 1;
      p.b_mn = 1;
      p.lda = l == 0 ? x_ld : in_dim;
      p.ldb = units;
      p.M = in_dim;
      p.N = units;
      p.K = B;
      p.ldc = units;
      p.accumulate = 1;
      for (int e = 0; e < E; ++e) {
        p.A[e] = l == 0 ? x : H[l - 1] + (int64_t)e * B * in_dim;
        p.B[e] = ws.dH[l] + (int64_t)e * B * units;
        p.C[e] = const_cast<float*>(g->expert[e][l].w);
      }
      if ((rc = tf32_gemm(p, st))) return rc;
      for (int e = 0; e < E; ++e)
        if ((rc = tf32_colsum(ws.dH[l] + (int64_t)e * B * units, units, B, units, const_cast<float*>(g->expert[e][l].b), 1,
                              ws.tf_colsum, st)))
          return rc;
    } else {
      GemmGroup grp{};
      grp.use_tc = use_tc;
      for (int e = 0; e < E; ++e) {
        const float* in = l == 0 ? x : H[l - 1] + (int64_t)e * B * in_dim;
        wgrad(grp.p[e], in, l == 0 ? x_ld : in_dim, ws.dH[l] + (int64_t)e * B * units, units, B, in_dim, units,
              g->expert[e][l], ws.splits_layer[l], ws.part_expert[e]);
      }
      grp.n = E;
      if ((rc = gemm_group_launch(grp, st))) return rc;
    }
    if (l > 0 && tf_ok) {
      // dH_{l-1,e} = (dH_{l,e} W_e^T) * (H_{l-1,e} > 0): the TF kernel [in, units] is the K-major operand as stored
      Tf32Gemm p{};
      p.nz = E;
      p.lda = units;
      p.ldb = units;
      p.M = B;
      p.N = in_dim;
      p.K = units;
      p.ldc = in_dim;
      p.ld_mask = in_dim;
      for (int e = 0; e < E; ++e) {
        p.A[e] = ws.dH[l] + (int64_t)e * B * units;
        p.B[e] = w->expert[e][l].w;
        p.C[e] = ws.dH[l - 1] + (int64_t)e * B * in_dim;
        p.mask[e] = H[l - 1] + (int64_t)e * B * in_dim;
      }
      if ((rc = tf32_gemm(p, st))) return rc;
    } else if (l > 0) {
      GemmGroup grp{};
      grp.use_tc = use_tc;
      for (int e = 0; e < E; ++e) {
        GemmProb& p = grp.p[e];
        gemm_prob_init(p);
        p.n_parts = 1;
        p.part[0] = GemmPart{ws.dH[l] + (int64_t)e * B * units, w->expert[e][l].w, units, units, units, 0};
        p.M = B;
        p.N = in_dim;
        p.transB = 1;
        p.C = ws.dH[l - 1] + (int64_t)e * B * in_dim;
        p.ldc = in_dim;
        p.mask = H[l - 1] + (int64_t)e * B * in_dim;
        p.ld_mask = in_dim;
      }
      grp.n = E;
      if ((rc = gemm_group_launch(grp, st))) return rc;
    } else if (dx && dx_col0 < c.in_dim) {
      DMT_REQUIRE(E + NT <= kGemmMaxParts, DMT_ERR_UNSUPPORTED_SHAPE, "dmt_mmoe_bwd: experts + tasks = %d > %d", E + NT,
                  kGemmMaxParts);
      if (tf_ok) {   // dx[:, col0:] = sum_e dH_{0,e} W_{0,e}[col0:, :]^T (one launch, summed over the experts) ...
        Tf32Gemm q{};
        q.nz = E;
        q.reduce_z = 1;
        q.lda = units;
        q.ldb = units;
        q.M = B;
        q.N = c.in_dim - dx_col0;
        q.K = units;
        q.C[0] = dx + dx_col0;
        q.ldc = dx_ld;
        for (int e = 0; e < E; ++e) {
          q.A[e] = ws.dH[0] + (int64_t)e * B * units;
          q.B[e] = w->expert[e][0].w + (int64_t)dx_col0 * units;
        }
        if ((rc = tf32_gemm(q, st))) return rc;
      }
      GemmGroup grp{};
      grp.use_tc = use_tc;
      GemmProb& p = grp.p[0];
      gemm_prob_init(p);
      int n = 0;
      if (!tf_ok)
        for (int e = 0; e < E; ++e)
          p.part[n++] = GemmPart{ws.dH[0] + (int64_t)e * B * units, w->expert[e][0].w + (int64_t)dx_col0 * units, units,
                                 units, units, 0};
      for (int t = 0; t < NT; ++t)      // ... + the gates' share (K = experts per task: a few columns)
        p.part[n++] = GemmPart{ws.dgl + t * E, w->gate[t].w + (int64_t)dx_col0 * E, (int64_t)NT * E, E, E, 0};
      p.n_parts = n;
      p.M = B;
      p.N = c.in_dim - dx_col0;
      p.transB = 1;
      p.C = dx + dx_col0;
      p.ldc = dx_ld;
      p.accumulate = tf_ok ? 1 : 0;
      grp.n = 1;
      if ((rc = gemm_group_launch(grp, st))) return rc;
    }
  }
  return DMT_OK;
}

size_t mmoe_bwd_workspace_bytes(const dmt_mmoe_cfg* cfg) { return mmoe_bwd_carve(*cfg, nullptr, nullptr); }

int mmoe_head_launch(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const float* x, int64_t x_ld,
                     const void* h_last, int h_is_bf16, const float* gates, float* logits, cudaStream_t st);

// Training forward: expert layers through the grouped GEMM (fp32 activations kept for the backward; bf16
// tensor-core operands when cfg->precision says so), then the shared head kernel.
int mmoe_fwd_train_launch(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const float* x, int64_t x_ld,
                          float* logits, float* ws, cudaStream_t st) {
  const dmt_mmoe_cfg& c = *cfg;
  const int B = c.batch, E = c.n_experts;
  const float* in = x;
  int64_t in_ld = x_ld, in_stride = 0;
  int in_dim = c.in_dim;
  float* out = ws;
  bool tf_ok = gemm_tf32(c.precision) && E <= 4 && x_ld % 4 == 0;
  for (int l = 0; l < c.n_layers && tf_ok; ++l) tf_ok = c.units[l] % 4 == 0;
  for (int l = 0; l < c.n_layers; ++l) {
    const int units = c.units[l];
    if (tf_ok) {   // H_{l,e} = relu(in_e W_e + b_e): the TF kernel [in, units] is the MN-major operand as stored
      Tf32Gemm p{};
      p.nz = E;
      p.b_mn = 1;
      p.lda = in_ld;
      p.ldb = units;
      p.M = B;
      p.N = units;
      p.K = in_dim;
      p.ldc = units;
      p.relu = 1;
      for (int e = 0; e < E; ++e) {
        p.A[e] = in + in_stride * e;
        p.B[e] = w->expert[e][l].w;
        p.C[e] = out + (int64_t)e * B * units;
        p.bias[e] = w->expert[e][l].b;
      }
      int rc = tf32_gemm(p, st);
      if (rc) return rc;
    } else {
      GemmGroup grp{};
      grp.use_tc = gemm_engine(c.precision);
      for (int e = 0; e < E; ++e) {
        GemmProb& p = grp.p[e];
        gemm_prob_init(p);
        p.n_parts = 1;
        p.part[0] = GemmPart{in + in_stride * e, w->expert[e][l].w, in_ld, units, in_dim, 0};
        p.M = B;
        p.N = units;
        p.C = out + (int64_t)e * B * units;
        p.ldc = units;
        p.bias = w->expert[e][l].b;
        p.relu = 1;
      }
      grp.n = E;
      int rc = gemm_group_launch(grp, st);
      if (rc) return rc;
    }
    in = out;
    in_ld = units;
    in_dim = units;
    in_stride = (int64_t)B * units;
    out += (int64_t)E * B * units;
  }
  return mmoe_head_launch(cfg, w, x, x_ld, in, 0, nullptr, logits, st);
}

// ---------------------------------------------------------------------------------------------------
// Bias tower backward (mmoe_transformer_unbias.py:259-289): one thread per sample recomputes the tiny MLP
// and writes the per-layer activations / pre-activation gradients; the weight gradients are batch
// contractions handled by the grouped GEMM.
namespace {

constexpr int kMaxBiasWidth = 64;

struct BiasBwdArgs {
  dmt_bias_loss_cfg cfg;
  dmt_bias_weights w;
  const float* bias_in;
  int64_t bias_ld;
  const float* dy;                    // [B]
  float* act[DMT_MAX_LAYERS];         // [B, units_l] hidden activations
  float* dpre[DMT_MAX_LAYERS + 1];    // [B, units_l] (last: [B, 1])
  float* d_in;                        // [B, in_dim]
  int64_t d_in_ld;
};

__global__ void __launch_bounds__(128) bias_bwd_kernel(const __grid_constant__ BiasBwdArgs a) {
  const int B = a.cfg.batch;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int nh = a.cfg.n_hidden;
  float acts[DMT_MAX_LAYERS + 1][kMaxBiasWidth];
  int dims[DMT_MAX_LAYERS + 2];
  dims[0] = a.cfg.in_dim;
  for (int k = 0; k < dims[0]; ++k) acts[0][k] = __ldg(a.bias_in + (int64_t)b * a.bias_ld + k);
  for (int l = 0; l < nh; ++l) {
    const int units = a.cfg.units[l];
    dims[l + 1] = units;
    const float* __restrict__ W = a.w.layer[l].w;
    for (int n = 0; n < units; ++n) {
      float acc = 0.f;
      for (int k = 0; k < dims[l]; ++k) acc = fmaf(acts[l][k], __ldg(W + k * units + n), acc);
      float y = fmaxf(acc + __ldg(a.w.layer[l].b + n), 0.f);
      if (a.cfg.dropout_rate[l] > 0.f)
        y *= Dropout(a.cfg.dropout_rate[l], a.cfg.dropout_seed, kSiteBias + l).mult((uint32_t)(b * units + n));
      acts[l + 1][n] = y;
      a.act[l][(int64_t)b * units + n] = y;
    }
  }
  float dcur[kMaxBiasWidth], dnxt[kMaxBiasWidth];
  const float dyb = __ldg(a.dy + b);
  a.dpre[nh][b] = dyb;
  {
    const float* __restrict__ W = a.w.layer[nh].w;   // [dims[nh], 1]
    for (int k = 0; k < dims[nh]; ++k) dcur[k] = dyb * __ldg(W + k);
  }
  for (int l = nh - 1; l >= 0; --l) {
    const int units = dims[l + 1];
    for (int n = 0; n < units; ++n) {
      // out = relu(z) * M: a dropped or inactive unit carries no gradient, a kept one scales it by M
      float dp = acts[l + 1][n] > 0.f ? dcur[n] : 0.f;
      if (a.cfg.dropout_rate[l] > 0.f)
        dp *= Dropout(a.cfg.dropout_rate[l], a.cfg.dropout_seed, kSiteBias + l).mult((uint32_t)(b * units + n));
      dcur[n] = dp;
      a.dpre[l][(int64_t)b * units + n] = dp;
    }
    const float* __restrict__ W = a.w.layer[l].w;
    for (int k = 0; k < dims[l]; ++k) {
      float acc = 0.f;
      for (int n = 0; n < units; ++n) acc = fmaf(dcur[n], __ldg(W + k * units + n), acc);
      dnxt[k] = acc;
    }
    for (int k = 0; k < dims[l]; ++k) dcur[k] = dnxt[k];
  }
  for (int k = 0; k < dims[0]; ++k) a.d_in[(int64_t)b * a.d_in_ld + k] = dcur[k];
}

struct BiasBwdWs {
  float* act[DMT_MAX_LAYERS];
  float* dpre[DMT_MAX_LAYERS + 1];
  float* part[DMT_MAX_LAYERS + 1];
  int splits[DMT_MAX_LAYERS + 1];
};

size_t bias_bwd_carve(const dmt_bias_loss_cfg& c, void* base, BiasBwdWs* out) {
  Carver cv(base);
  BiasBwdWs w{};
  const size_t B = c.batch;
  int in_dim = c.in_dim;
  for (int l = 0; l <= c.n_hidden; ++l) {
    const int units = l < c.n_hidden ? c.units[l] : 1;
    if (l < c.n_hidden) w.act[l] = cv.take(B * units);
    w.dpre[l] = cv.take(B * units);
    w.splits[l] = gemm_pick_splits(in_dim, units, (int64_t)B);
    w.part[l] = cv.take((size_t)w.splits[l] * (in_dim + 1) * units);
    in_dim = units;
  }
  if (out) *out = w;
  return cv.off + 256;
}

}  // namespace

}  // namespace dmt

extern "C" {

size_t dmt_mmoe_bwd_workspace_bytes(const dmt_mmoe_cfg* cfg) {
  if (!cfg || cfg->n_layers <= 0 || cfg->n_layers > DMT_MAX_LAYERS) return 0;
  return dmt::mmoe_bwd_workspace_bytes(cfg);
}

int dmt_mmoe_bwd(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const float* x, int64_t x_ld,
                 const void* fwd_workspace, const float* dlogits, const dmt_mmoe_grads* grads, float* dx, int64_t dx_ld,
                 int32_t dx_col0, void* workspace, size_t workspace_bytes, void* stream) {
  DMT_REQUIRE(cfg && w && x && fwd_workspace && dlogits && grads && workspace, DMT_ERR_INVALID_ARGUMENT,
              "dmt_mmoe_bwd: null pointer");
  DMT_REQUIRE(cfg->batch >= 0 && cfg->in_dim > 0 && cfg->n_experts > 0 && cfg->n_experts <= DMT_MAX_EXPERTS &&
                  cfg->n_layers > 0 && cfg->n_layers <= DMT_MAX_LAYERS && cfg->n_tasks > 0 &&
                  cfg->n_tasks <= DMT_MAX_TASKS && cfg->n_tower_layers >= 0 && cfg->n_tower_layers <= DMT_MAX_LAYERS,
              DMT_ERR_INVALID_ARGUMENT, "dmt_mmoe_bwd: configuration out of range");
  DMT_REQUIRE(x_ld >= cfg->in_dim && (!dx || dx_ld >= cfg->in_dim) && dx_col0 >= 0, DMT_ERR_INVALID_ARGUMENT,
              "dmt_mmoe_bwd: bad strides");
  DMT_REQUIRE(((uintptr_t)workspace & 255) == 0 && workspace_bytes >= dmt::mmoe_bwd_workspace_bytes(cfg),
              DMT_ERR_WORKSPACE_TOO_SMALL, "dmt_mmoe_bwd: workspace %zu < %zu bytes (256-byte aligned)", workspace_bytes,
              dmt::mmoe_bwd_workspace_bytes(cfg));
  if (cfg->batch == 0) return DMT_OK;
  return dmt::mmoe_bwd_launch(cfg, w, x, x_ld, (const float*)fwd_workspace, dlogits, grads, dx, dx_ld, dx_col0,
                              workspace, (cudaStream_t)stream);
}

size_t dmt_mmoe_train_workspace_bytes(const dmt_mmoe_cfg* cfg) {
  if (!cfg) return 0;
  size_t floats = 0;
  for (int l = 0; l < cfg->n_layers && l < DMT_MAX_LAYERS; ++l)
    floats += (size_t)cfg->n_experts * cfg->batch * cfg->units[l];
  return floats * sizeof(float) + 256;
}

int dmt_mmoe_fwd_train(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const float* x, int64_t x_ld, float* logits,
                       void* workspace, size_t workspace_bytes, void* stream) {
  DMT_REQUIRE(cfg && w && x && logits && workspace, DMT_ERR_INVALID_ARGUMENT, "dmt_mmoe_fwd_train: null pointer");
  DMT_REQUIRE(cfg->batch >= 0 && cfg->in_dim > 0 && cfg->n_experts > 0 && cfg->n_experts <= DMT_MAX_EXPERTS &&
                  cfg->n_layers > 0 && cfg->n_layers <= DMT_MAX_LAYERS && cfg->n_tasks > 0 &&
                  cfg->n_tasks <= DMT_MAX_TASKS && cfg->n_tower_layers >= 0 && cfg->n_tower_layers <= DMT_MAX_LAYERS,
              DMT_ERR_INVALID_ARGUMENT, "dmt_mmoe_fwd_train: configuration out of range");
  DMT_REQUIRE(x_ld >= cfg->in_dim, DMT_ERR_INVALID_ARGUMENT, "dmt_mmoe_fwd_train: x_ld < in_dim");
  DMT_REQUIRE(workspace_bytes >= dmt_mmoe_train_workspace_bytes(cfg), DMT_ERR_WORKSPACE_TOO_SMALL,
              "dmt_mmoe_fwd_train: workspace %zu < %zu bytes", workspace_bytes, dmt_mmoe_train_workspace_bytes(cfg));
  if (cfg->batch == 0) return DMT_OK;
  return dmt::mmoe_fwd_train_launch(cfg, w, x, x_ld, logits, (float*)workspace, (cudaStream_t)stream);
}

size_t dmt_bias_bwd_workspace_bytes(const dmt_bias_loss_cfg* cfg) {
  if (!cfg || cfg->n_hidden < 0 || cfg->n_hidden > DMT_MAX_LAYERS) return 0;
  return dmt::bias_bwd_carve(*cfg, nullptr, nullptr);
}

int dmt_bias_bwd(const dmt_bias_loss_cfg* cfg, const dmt_bias_weights* w, const float* bias_in, int64_t bias_ld,
                 const float* dy_bias, const dmt_bias_grads* grads, float* d_bias_in, int64_t d_in_ld, void* workspace,
                 size_t workspace_bytes, void* stream) {
  DMT_REQUIRE(cfg && w && bias_in && dy_bias && grads && d_bias_in && workspace, DMT_ERR_INVALID_ARGUMENT,
              "dmt_bias_bwd: null pointer");
  DMT_REQUIRE(cfg->n_hidden >= 0 && cfg->n_hidden <= DMT_MAX_LAYERS && cfg->in_dim > 0 &&
                  cfg->in_dim <= dmt::kMaxBiasWidth && bias_ld >= cfg->in_dim && d_in_ld >= cfg->in_dim,
              DMT_ERR_INVALID_ARGUMENT, "dmt_bias_bwd: in_dim=%d n_hidden=%d", cfg->in_dim, cfg->n_hidden);
  for (int l = 0; l < cfg->n_hidden; ++l)
    DMT_REQUIRE(cfg->units[l] > 0 && cfg->units[l] <= dmt::kMaxBiasWidth, DMT_ERR_UNSUPPORTED_SHAPE,
                "dmt_bias_bwd: hidden_units_bias[%d]=%d (max %d)", l, cfg->units[l], dmt::kMaxBiasWidth);
  DMT_REQUIRE(((uintptr_t)workspace & 255) == 0 && workspace_bytes >= dmt_bias_bwd_workspace_bytes(cfg),
              DMT_ERR_WORKSPACE_TOO_SMALL, "dmt_bias_bwd: workspace %zu < %zu bytes", workspace_bytes,
              dmt_bias_bwd_workspace_bytes(cfg));
  if (cfg->batch <= 0) return DMT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  dmt::BiasBwdWs ws;
  dmt::bias_bwd_carve(*cfg, workspace, &ws);
  dmt::BiasBwdArgs a{};
  a.cfg = *cfg;
  a.w = *w;
  a.bias_in = bias_in;
  a.bias_ld = bias_ld;
  a.dy = dy_bias;
  for (int l = 0; l < cfg->n_hidden; ++l) a.act[l] = ws.act[l];
  for (int l = 0; l <= cfg->n_hidden; ++l) a.dpre[l] = ws.dpre[l];
  a.d_in = d_bias_in;
  a.d_in_ld = d_in_ld;
  dmt::bias_bwd_kernel<<<(cfg->batch + 127) / 128, 128, 0, st>>>(a);
  DMT_CUDA_LAUNCH_CHECK("bias_bwd_kernel");
  dmt::GemmGroup grp{};
  int in_dim = cfg->in_dim;
  const float* in = bias_in;
  int64_t in_ld = bias_ld;
  for (int l = 0; l <= cfg->n_hidden; ++l) {
    const int units = l < cfg->n_hidden ? cfg->units[l] : 1;
    dmt::wgrad(grp.p[l], in, in_ld, ws.dpre[l], units, cfg->batch, in_dim, units, grads->layer[l], ws.splits[l],
               ws.part[l]);
    if (l < cfg->n_hidden) {
      in = ws.act[l];
      in_ld = units;
      in_dim = units;
    }
  }
  grp.n = cfg->n_hidden + 1;
  return dmt::gemm_group_launch(grp, st);
}

}  // extern "C"
