// dmt_seq_encode_fwd: argument validation + dispatch on dmt_precision.
#include "dmt_common.cuh"

namespace dmt {
int seq_encode_f32_launch(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, float* out,
                          int64_t out_ld, cudaStream_t st);
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
  return 256;   // the fused fp32 path keeps every intermediate on chip
}

int dmt_seq_encode_fwd(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w, float* out,
                       int64_t out_ld, void* workspace, size_t workspace_bytes, void* stream) {
  (void)workspace;
  (void)workspace_bytes;
  int rc = validate_seq(cfg, in, w);
  if (rc != DMT_OK) return rc;
  DMT_REQUIRE(out && out_ld >= cfg->d_model, DMT_ERR_INVALID_ARGUMENT, "dmt_seq_encode_fwd: bad output");
  if (cfg->batch == 0) return DMT_OK;
  DMT_REQUIRE(cfg->precision == DMT_PRECISION_F32, DMT_ERR_UNSUPPORTED_SHAPE,
              "dmt_seq_encode_fwd: precision %d not built", cfg->precision);
  return dmt::seq_encode_f32_launch(cfg, in, w, out, out_ld, (cudaStream_t)stream);
}

}  // extern "C"
