// dmt_seq_encode_fwd, DMT_PRECISION_BF16: host side of the fused tcgen05 tile kernels (seq_encode_tc3.cu) --
// which shapes they cover, the bf16 weight images (dmt_seq_prepare_weights), kernel-argument marshalling and the
// launch dispatch.
#include <limits.h>
#include <stdlib.h>

#include "dmt_common.cuh"
#include "seq_tc.cuh"
#include "umma.cuh"

namespace dmt {

using namespace umma;

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

static unsigned long long* g_seq_profile = nullptr;   // diagnostics only (dmt_debug_seq_profile)
void seq_tc_set_profile(unsigned long long* p) { g_seq_profile = p; }

size_t seq_tc_prepared_bytes(const dmt_seq_cfg* cfg) {
  return (prep_total(cfg->d_model, cfg->d_ff, cfg->num_heads) * 2 + 511) / 256 * 256;
}

// seq_encode_tc3.cu
bool seq_tc2_supported(const dmt_seq_cfg* cfg);
size_t seq_tc2_ctx_bytes(const dmt_seq_cfg* cfg);

size_t seq_tc_sched_bytes(const dmt_seq_cfg* cfg);
int seq_encode_multi_launch(int n, const SeqTcArgs* args, void* const* scheds, bool defer_tail, cudaEvent_t wait_before_encode,
                            cudaStream_t st);

// workspace = [prepared weight images | decoder-context images | length-class schedule (perm, counts)]
size_t seq_tc_workspace_bytes(const dmt_seq_cfg* cfg) {
  return seq_tc_prepared_bytes(cfg) + seq_tc2_ctx_bytes(cfg) + seq_tc_sched_bytes(cfg);
}

bool seq_tc_supported(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const char** why) {
  *why = nullptr;
  if (!(cfg->d_model == 64 && cfg->d_ff == 256 && cfg->num_heads == 2))
    *why = "bf16 tensor-core path is built for d_model=64, d_ff=256, 2 heads";
  else if (cfg->n_enc_blocks != 1 || cfg->n_dec_blocks != 1)
    *why = "bf16 tensor-core path is built for 1 encoder + 1 decoder block";
  else if (!seq_tc2_supported(cfg))
    *why = "bf16 tensor-core path: position table / feature chunks exceed the tile kernel's shared-memory plan "
           "(transformer_maxlen_k <= 55, <= 8 feature chunks)";
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
    fill_args(cfgs[i], ins[i], ws[i], outs[i], out_lds[i], workspaces[i], args[i]);
  }
  return seq_tails_launch(n, args, st);
}

int seq_tc_multi(int n, const dmt_seq_cfg* const* cfgs, const dmt_seq_input* const* ins, const dmt_seq_weights* const* ws,
                 float* const* outs, const int64_t* out_lds, void* const* workspaces, cudaEvent_t wait_before_encode,
                 cudaStream_t st) {
  SeqTcArgs args[DMT_MAX_TAIL_SEQS];
  void* scheds[DMT_MAX_TAIL_SEQS];
  for (int i = 0; i < n; ++i) {
    fill_args(cfgs[i], ins[i], ws[i], outs[i], out_lds[i], workspaces[i], args[i]);
    scheds[i] = static_cast<uint8_t*>(workspaces[i]) + seq_tc_prepared_bytes(cfgs[i]) + seq_tc2_ctx_bytes(cfgs[i]);
  }
  return seq_encode_multi_launch(n, args, scheds, false, wait_before_encode, st);
}

int seq_encode_tc_launch(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, float* out,
                         int64_t out_ld, void* workspace, cudaStream_t st) {
  // one sequence = a one-sequence launch of the multi-sequence kernel (length classes included)
  SeqTcArgs a;
  fill_args(cfg, in, w, out, out_ld, workspace, a);
  void* sched = static_cast<uint8_t*>(workspace) + seq_tc_prepared_bytes(cfg) + seq_tc2_ctx_bytes(cfg);
  return seq_encode_multi_launch(1, &a, &sched, (cfg->flags & DMT_SEQ_DEFER_TAIL) != 0, nullptr, st);
}

}  // namespace dmt
