// dmt_seq_encode_fwd: argument validation + dispatch on dmt_precision.
#include "dmt_common.cuh"
#include "seq_train.cuh"

namespace dmt {
int seq_encode_f32_launch(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, float* out,
                          int64_t out_ld, const SeqSaved* saved, cudaStream_t st);
int seq_bwd_launch(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, int64_t T,
                   const SeqSaved& sv, const float* d_out, int64_t d_out_ld, const dmt_seq_grads* g, float* d_tokens,
                   float* d_target, void* ws, cudaStream_t st);
size_t seq_bwd_workspace_bytes(const dmt_seq_cfg* cfg, int64_t T);
int seq_fwd_train_pipeline(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, float* out,
                           int64_t out_ld, int64_t T, const SeqSaved& sv, cudaStream_t st);
size_t seq_tc_prepared_bytes(const dmt_seq_cfg* cfg);
bool seq_tc_supported(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const char** why);
int seq_tc_prepare(const dmt_seq_cfg* cfg, const dmt_seq_weights* w, void* prepared, cudaStream_t st);
void seq_tc_set_profile(unsigned long long* p);
int seq_encode_tc_launch(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, float* out,
                         int64_t out_ld, void* workspace, cudaStream_t st);
size_t seq_tc_workspace_bytes(const dmt_seq_cfg* cfg);
int seq_tc_multi(int n, const dmt_seq_cfg* const* cfgs, const dmt_seq_input* const* ins, const dmt_seq_weights* const* ws,
                 float* const* outs, const int64_t* out_lds, void* const* workspaces, cudaEvent_t wait_before_encode,
                 cudaStream_t st);
int seq_timer_enable(int on);
int seq_timer_read(float* total_ms, int32_t* launches);
int seq_tc_tails(int n, const dmt_seq_cfg* const* cfgs, const dmt_seq_input* const* ins, const dmt_seq_weights* const* ws,
                 float* const* outs, const int64_t* out_lds, void* const* workspaces, cudaStream_t st);
}

static int validate_seq(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w) {
  DMT_REQUIRE(cfg && in && w, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_encode_fwd: null pointer");
  DMT_REQUIRE(cfg->batch >= 0, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_encode_fwd: batch=%d", cfg->batch);
  DMT_REQUIRE(cfg->n_feats > 0 && cfg->n_feats <= DMT_MAX_SEQ_FEATS, DMT_ERR_INVALID_ARGUMENT,
              "dmt_seq_encode_fwd: n_feats=%d (max %d)", cfg->n_feats, DMT_MAX_SEQ_FEATS);
  DMT_REQUIRE(cfg->n_enc_blocks >= 0 && cfg->n_enc_blocks <= DMT_MAX_BLOCKS && cfg->n_dec_blocks >= 0 &&
                  cfg->n_dec_blocks <= DMT_MAX_BLOCKS,
              DMT_ERR_UNSUPPORTED_SHAPE, "dmt_seq_encode_fwd: blocks enc=%d dec=%d (max %d)", cfg->n_enc_blocks,
              cfg->n_dec_blocks, DMT_MAX_BLOCKS);
  DMT_REQUIRE(cfg->d_model > 0 && cfg->d_model % 4 == 0 && cfg->d_model <= 256, DMT_ERR_UNSUPPORTED_SHAPE,
              "dmt_seq_encode_fwd: d_model=%d must be a multiple of 4 and <= 256", cfg->d_model);
  DMT_REQUIRE(cfg->d_ff > 0 && cfg->d_ff % 4 == 0, DMT_ERR_UNSUPPORTED_SHAPE,
              "dmt_seq_encode_fwd: d_ff=%d must be a multiple of 4", cfg->d_ff);
  DMT_REQUIRE(cfg->num_heads > 0 && cfg->d_model % cfg->num_heads == 0, DMT_ERR_INVALID_ARGUMENT,
              "dmt_seq_encode_fwd: d_model %d not divisible by num_heads %d", cfg->d_model, cfg->num_heads);
  DMT_REQUIRE(cfg->maxlen > 0 && cfg->maxlen <= DMT_MAX_SEQ_LEN, DMT_ERR_UNSUPPORTED_SHAPE,
              "dmt_seq_encode_fwd: maxlen=%d (max %d)", cfg->maxlen, DMT_MAX_SEQ_LEN);
  DMT_REQUIRE(w->pos, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_encode_fwd: position table missing");
  for (int f = 0; f < cfg->n_feats; ++f)
    DMT_REQUIRE(in->table[f] && in->ids[f] && in->offsets[f] && in->item_ids[f] && in->dim[f] > 0 && in->rows[f] > 0,
                DMT_ERR_INVALID_ARGUMENT, "dmt_seq_encode_fwd: feature pair %d is incomplete", f);
  return DMT_OK;
}

extern "C" {

size_t dmt_seq_encode_workspace_bytes(const dmt_seq_cfg* cfg, int64_t max_tokens) {
  (void)max_tokens;
  if (!cfg) return 0;
  if (cfg->precision == DMT_PRECISION_BF16)   // weight images + per-sample decoder contexts (grows with batch)
    return dmt::seq_tc_workspace_bytes(cfg);
  return 256;   // the fused fp32 path keeps every intermediate on chip
}

int dmt_seq_prepare_weights(const dmt_seq_cfg* cfg, const dmt_seq_weights* w, void* prepared, size_t prepared_bytes,
                            void* stream) {
  DMT_REQUIRE(cfg && w && prepared, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_prepare_weights: null pointer");
  DMT_REQUIRE(cfg->precision == DMT_PRECISION_BF16, DMT_ERR_INVALID_ARGUMENT,
              "dmt_seq_prepare_weights: only the bf16 path has prepared weights");
  DMT_REQUIRE(prepared_bytes >= dmt::seq_tc_prepared_bytes(cfg), DMT_ERR_WORKSPACE_TOO_SMALL,
              "dmt_seq_prepare_weights: buffer %zu < %zu bytes", prepared_bytes, dmt::seq_tc_prepared_bytes(cfg));
  DMT_REQUIRE(((uintptr_t)prepared & 15) == 0, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_prepare_weights: unaligned buffer");
  return dmt::seq_tc_prepare(cfg, w, prepared, (cudaStream_t)stream);
}

int dmt_seq_encode_fwd(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, float* out,
                       int64_t out_ld, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = validate_seq(cfg, in, w);
  if (rc != DMT_OK) return rc;
  DMT_REQUIRE(out && out_ld >= cfg->d_model, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_encode_fwd: bad output");
  if (cfg->batch == 0) return DMT_OK;
  if (cfg->precision == DMT_PRECISION_F32)
    return dmt::seq_encode_f32_launch(cfg, in, w, out, out_ld, nullptr, (cudaStream_t)stream);
  DMT_REQUIRE(cfg->precision == DMT_PRECISION_BF16, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_encode_fwd: precision %d",
              cfg->precision);
  const char* why = nullptr;
  DMT_REQUIRE(dmt::seq_tc_supported(cfg, in, &why), DMT_ERR_UNSUPPORTED_SHAPE, "dmt_seq_encode_fwd: %s", why);
  DMT_REQUIRE(workspace && workspace_bytes >= dmt::seq_tc_workspace_bytes(cfg), DMT_ERR_WORKSPACE_TOO_SMALL,
              "dmt_seq_encode_fwd(bf16): workspace %zu < %zu bytes (dmt_seq_encode_workspace_bytes: the images written "
              "by dmt_seq_prepare_weights + batch-sized scratch)", workspace_bytes, dmt::seq_tc_workspace_bytes(cfg));
  DMT_REQUIRE(((uintptr_t)workspace & 15) == 0, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_encode_fwd(bf16): unaligned workspace");
  return dmt::seq_encode_tc_launch(cfg, in, w, out, out_ld, workspace, (cudaStream_t)stream);
}

int dmt_seq_tail_fwd(int32_t n_seq, const dmt_seq_cfg* const* cfgs, const dmt_seq_input* const* ins,
                     const dmt_seq_weights* const* ws, float* const* outs, const int64_t* out_lds,
                     void* const* workspaces, void* stream) {
  DMT_REQUIRE(n_seq >= 0 && n_seq <= DMT_MAX_TAIL_SEQS, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_tail_fwd: n_seq=%d (max %d)",
              n_seq, DMT_MAX_TAIL_SEQS);
  if (n_seq == 0) return DMT_OK;
  DMT_REQUIRE(cfgs && ins && ws && outs && out_lds && workspaces, DMT_ERR_INVALID_ARGUMENT,
              "dmt_seq_tail_fwd: null pointer");
  for (int i = 0; i < n_seq; ++i) {
    int rc = validate_seq(cfgs[i], ins[i], ws[i]);
    if (rc != DMT_OK) return rc;
    DMT_REQUIRE(cfgs[i]->precision == DMT_PRECISION_BF16 && outs[i] && workspaces[i] && out_lds[i] >= cfgs[i]->d_model,
                DMT_ERR_INVALID_ARGUMENT, "dmt_seq_tail_fwd: sequence %d: bf16 path, output and workspace required", i);
    const char* why = nullptr;
    DMT_REQUIRE(dmt::seq_tc_supported(cfgs[i], ins[i], &why), DMT_ERR_UNSUPPORTED_SHAPE, "dmt_seq_tail_fwd: %s", why);
  }
  return dmt::seq_tc_tails(n_seq, cfgs, ins, ws, outs, out_lds, workspaces, (cudaStream_t)stream);
}

}  // extern "C"

// shared by dmt_seq_encode_multi_fwd and dmt_forward_bf16 (forward.cu), which passes the event its sequence stream
// waits for between the length-class kernel and the tile kernel
namespace dmt {
int seq_encode_multi_checked(int32_t n_seq, const dmt_seq_cfg* const* cfgs, const dmt_seq_input* const* ins,
                             const dmt_seq_weights* const* ws, float* const* outs, const int64_t* out_lds,
                             void* const* workspaces, const size_t* workspace_bytes, cudaEvent_t wait_before_encode,
                             void* stream) {
  DMT_REQUIRE(n_seq >= 0 && n_seq <= DMT_MAX_TAIL_SEQS, DMT_ERR_INVALID_ARGUMENT,
              "dmt_seq_encode_multi_fwd: n_seq=%d (max %d)", n_seq, DMT_MAX_TAIL_SEQS);
  if (n_seq == 0) return DMT_OK;
  DMT_REQUIRE(cfgs && ins && ws && outs && out_lds && workspaces && workspace_bytes, DMT_ERR_INVALID_ARGUMENT,
              "dmt_seq_encode_multi_fwd: null pointer");
  for (int i = 0; i < n_seq; ++i) {
    int rc = validate_seq(cfgs[i], ins[i], ws[i]);
    if (rc != DMT_OK) return rc;
    DMT_REQUIRE(cfgs[i]->precision == DMT_PRECISION_BF16, DMT_ERR_INVALID_ARGUMENT,
                "dmt_seq_encode_multi_fwd: sequence %d: only the bf16 path has a multi-sequence launch", i);
    DMT_REQUIRE(outs[i] && out_lds[i] >= cfgs[i]->d_model, DMT_ERR_INVALID_ARGUMENT,
                "dmt_seq_encode_multi_fwd: sequence %d: bad output", i);
    DMT_REQUIRE(cfgs[i]->batch > 0, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_encode_multi_fwd: sequence %d: empty batch", i);
    const char* why = nullptr;
    DMT_REQUIRE(dmt::seq_tc_supported(cfgs[i], ins[i], &why), DMT_ERR_UNSUPPORTED_SHAPE, "dmt_seq_encode_multi_fwd: %s", why);
    DMT_REQUIRE(workspaces[i] && workspace_bytes[i] >= dmt::seq_tc_workspace_bytes(cfgs[i]), DMT_ERR_WORKSPACE_TOO_SMALL,
                "dmt_seq_encode_multi_fwd: sequence %d: workspace %zu < %zu bytes (dmt_seq_encode_workspace_bytes)", i,
                workspace_bytes[i], dmt::seq_tc_workspace_bytes(cfgs[i]));
    DMT_REQUIRE(((uintptr_t)workspaces[i] & 15) == 0, DMT_ERR_INVALID_ARGUMENT,
                "dmt_seq_encode_multi_fwd: sequence %d: unaligned workspace", i);
  }
  return dmt::seq_tc_multi(n_seq, cfgs, ins, ws, outs, out_lds, workspaces, wait_before_encode, (cudaStream_t)stream);
}
}  // namespace dmt

extern "C" {

int dmt_seq_encode_multi_fwd(int32_t n_seq, const dmt_seq_cfg* const* cfgs, const dmt_seq_input* const* ins,
                             const dmt_seq_weights* const* ws, float* const* outs, const int64_t* out_lds,
                             void* const* workspaces, const size_t* workspace_bytes, void* stream) {
  return dmt::seq_encode_multi_checked(n_seq, cfgs, ins, ws, outs, out_lds, workspaces, workspace_bytes, nullptr, stream);
}

size_t dmt_seq_saved_bytes(const dmt_seq_cfg* cfg, int64_t n_tokens) {
  if (!cfg || n_tokens < 0) return 0;
  return dmt::seq_saved_carve(*cfg, n_tokens, nullptr, nullptr);
}

int dmt_seq_encode_fwd_train(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, float* out,
                             int64_t out_ld, int64_t n_tokens, void* saved, size_t saved_bytes, void* stream) {
  int rc = validate_seq(cfg, in, w);
  if (rc != DMT_OK) return rc;
  DMT_REQUIRE(out && out_ld >= cfg->d_model, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_encode_fwd_train: bad output");
  DMT_REQUIRE(n_tokens >= 0 && saved && ((uintptr_t)saved & 255) == 0, DMT_ERR_INVALID_ARGUMENT,
              "dmt_seq_encode_fwd_train: `saved` must be a 256-byte aligned buffer");
  DMT_REQUIRE(saved_bytes >= dmt_seq_saved_bytes(cfg, n_tokens), DMT_ERR_WORKSPACE_TOO_SMALL,
              "dmt_seq_encode_fwd_train: saved buffer %zu < %zu bytes", saved_bytes, dmt_seq_saved_bytes(cfg, n_tokens));
  if (cfg->batch == 0) return DMT_OK;
  dmt::SeqSaved sv;
  dmt::seq_saved_carve(*cfg, n_tokens, saved, &sv);
  if (cfg->precision != DMT_PRECISION_F32)   // row-batched pipeline on the tensor-core GEMM engine
    return dmt::seq_fwd_train_pipeline(cfg, in, w, out, out_ld, n_tokens, sv, (cudaStream_t)stream);
  return dmt::seq_encode_f32_launch(cfg, in, w, out, out_ld, &sv, (cudaStream_t)stream);
}

size_t dmt_seq_bwd_workspace_bytes(const dmt_seq_cfg* cfg, int64_t n_tokens) {
  if (!cfg || n_tokens < 0) return 0;
  return dmt::seq_bwd_workspace_bytes(cfg, n_tokens);
}

int dmt_seq_encode_bwd(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, int64_t n_tokens,
                       const void* saved, size_t saved_bytes, const float* d_out, int64_t d_out_ld,
                       const dmt_seq_grads* grads, float* d_tokens, float* d_target, void* workspace,
                       size_t workspace_bytes, void* stream) {
  int rc = validate_seq(cfg, in, w);
  if (rc != DMT_OK) return rc;
  DMT_REQUIRE(saved && d_out && grads && d_target && workspace && (d_tokens || n_tokens == 0),
              DMT_ERR_INVALID_ARGUMENT, "dmt_seq_encode_bwd: null pointer");
  DMT_REQUIRE(d_out_ld >= cfg->d_model && n_tokens >= 0, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_encode_bwd: bad sizes");
  DMT_REQUIRE((((uintptr_t)saved | (uintptr_t)workspace) & 255) == 0, DMT_ERR_INVALID_ARGUMENT,
              "dmt_seq_encode_bwd: saved / workspace must be 256-byte aligned");
  DMT_REQUIRE(saved_bytes >= dmt_seq_saved_bytes(cfg, n_tokens), DMT_ERR_WORKSPACE_TOO_SMALL,
              "dmt_seq_encode_bwd: saved buffer %zu < %zu bytes", saved_bytes, dmt_seq_saved_bytes(cfg, n_tokens));
  DMT_REQUIRE(workspace_bytes >= dmt::seq_bwd_workspace_bytes(cfg, n_tokens), DMT_ERR_WORKSPACE_TOO_SMALL,
              "dmt_seq_encode_bwd: workspace %zu < %zu bytes", workspace_bytes,
              dmt::seq_bwd_workspace_bytes(cfg, n_tokens));
  DMT_REQUIRE(grads->pos, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_encode_bwd: gradient descriptor incomplete");
  if (cfg->batch == 0) return DMT_OK;
  dmt::SeqSaved sv;
  dmt::seq_saved_carve(*cfg, n_tokens, const_cast<void*>(saved), &sv);
  return dmt::seq_bwd_launch(cfg, in, w, n_tokens, sv, d_out, d_out_ld, grads, d_tokens, d_target, workspace,
                             (cudaStream_t)stream);
}

int dmt_debug_seq_timer(int32_t enable) { return dmt::seq_timer_enable(enable); }

int dmt_debug_seq_timer_read(float* total_ms, int32_t* launches) {
  DMT_REQUIRE(total_ms && launches, DMT_ERR_INVALID_ARGUMENT, "dmt_debug_seq_timer_read: null pointer");
  return dmt::seq_timer_read(total_ms, launches);
}

int dmt_debug_seq_profile(void* device_counters) {
  dmt::seq_tc_set_profile((unsigned long long*)device_counters);
  return DMT_OK;
}

}  // extern "C"
