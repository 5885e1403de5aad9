/*
 * dmt_b200.h -- C ABI of the B200-native DMT ranking hot path.
 *
 * The reference (guyulongcs/CIKM2020_DMT) has no FFI: every operator on this path is a
 * stock TensorFlow-1.12 op reached through Python.  Each entry point below therefore
 * replaces a *group of TF op calls* in the reference's Python; the file:line it replaces
 * is cited on the declaration (paths relative to DMT_code/).  INTEGRATION.md shows the
 * ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - every export returns int: 0 = ok, < 0 = dmt_status code; dmt_last_error() returns a
 *     thread-local message for the last failure on the calling thread;
 *   - no C++ types, no torch types: plain pointers, sizes and POD structs;
 *   - every pointer inside the structs is a DEVICE pointer unless stated; the structs
 *     themselves live in HOST memory and are read during the call;
 *   - enqueue-only: work is launched on the caller's `stream` (a cudaStream_t passed as
 *     void*); no allocation, no free, no hidden synchronisation, no global mutable state;
 *   - the caller owns every buffer including the workspace (size it with the
 *     *_workspace_bytes helpers);
 *   - dense weights use the TF layout: kernel [in, out] row-major, y = x W + b;
 *   - id features are CSR: values int32 [nnz] (post-lookup index in [0, V)), offsets
 *     int32 [B+1] -- the left-packed tf.SparseTensor the reference's input pipeline
 *     produces (data_feed/tfrecord_mask.py:23-84, data_feed/index_tables.py:37-45).
 */
#ifndef DMT_B200_H
#define DMT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define DMT_API __declspec(dllexport)
#else
#define DMT_API __attribute__((visibility("default")))
#endif

#define DMT_ABI_VERSION 6

#define DMT_MAX_SEQ_FEATS 8   /* (user, item) feature pairs per behaviour sequence */
#define DMT_MAX_BLOCKS 4      /* transformer_num_blocks_{encode,decode}            */
#define DMT_MAX_POOL_FEATS 64 /* pooled features per dmt_pool_mean_* launch        */
#define DMT_MAX_EXPERTS 8
#define DMT_MAX_TASKS 4
#define DMT_MAX_LAYERS 4
#define DMT_MAX_SEQ_LEN 64    /* fused encoder keeps one whole sequence on chip     */
#define DMT_MAX_GRAD_SOURCES 16 /* lookups feeding one table per step (3 seq + 3 target + pooled) */

typedef enum dmt_status {
  DMT_OK = 0,
  DMT_ERR_INVALID_ARGUMENT = -1,
  DMT_ERR_UNSUPPORTED_SHAPE = -2,
  DMT_ERR_WORKSPACE_TOO_SMALL = -3,
  DMT_ERR_CUDA = -4,
  DMT_ERR_NO_DEVICE = -5
} dmt_status;

typedef enum dmt_precision {
  DMT_PRECISION_F32 = 0,  /* fp32 CUDA-core math end to end (parity tolerance 1e-4)  */
  DMT_PRECISION_BF16 = 1, /* bf16 operands on tcgen05 tensor cores, fp32 accumulate  */
  DMT_PRECISION_BF16X3 = 2 /* training entry points only (fwd_train / bwd): every GEMM operand is split
                              x = hi + lo into two bf16 images and each product is hi*hi + hi*lo + lo*hi on
                              tcgen05 -- fp32-grade gradients at tensor-core speed                       */
  ,
  DMT_PRECISION_TF32 = 3   /* training entry points only: the per-token GEMMs of the sequence pipeline run on the
                              TMA-fed tcgen05 kind::tf32 engine straight from the fp32 activations (no operand
                              conversion pass; 10-bit operand mantissa, fp32 accumulate); the MMoE GEMMs use the
                              bf16x3 engine                                                             */
} dmt_precision;

/* one tf.layers.dense / base.dense_layer: kernel [in,out] + bias [out] */
typedef struct dmt_dense {
  const float* w;
  const float* b;
} dmt_dense;

/* TransformerModel_util.py:58-78 */
typedef struct dmt_layernorm {
  const float* gamma;
  const float* beta;
} dmt_layernorm;

/* TransformerModel_util.py:186-207: Q/K/V projections (no W_O) + LayerNorm */
typedef struct dmt_attn_weights {
  dmt_dense q, k, v;
  dmt_layernorm ln;
} dmt_attn_weights;

/* TransformerModel_util.py:212-235; shared by encoder block i and decoder block i */
typedef struct dmt_ff_weights {
  dmt_dense w1, w2;
  dmt_layernorm ln;
} dmt_ff_weights;

/* dmt_seq_cfg.flags (bf16 path): leave the decoder tail (ctx Wv + residual -> LN -> FF -> LN) of this call to a
 * later dmt_seq_tail_fwd, which runs the tails of several sequences as ONE launch.  Without the bit,
 * dmt_seq_encode_fwd is self-contained (three launches: length classes, tile kernel, tails). */
#define DMT_SEQ_DEFER_TAIL 1
/* dmt_seq_cfg.flags: the caller guarantees that NO sequence of the batch is longer than min(slot_len, maxlen) (it
 * knows the lengths: they came through its host memory).  Lets the training pipeline cut whole-sample tiles out of
 * the packed token rows without a pass over the offsets (tensor-core attention); without it the per-sample kernels
 * run, which cap over-long sequences themselves. */
#define DMT_SEQ_LEN_EXACT 2
/* dmt_seq_cfg.flags (bf16 path: dmt_seq_encode_fwd / dmt_seq_tail_fwd / dmt_seq_encode_multi_fwd): `out` is a bf16
 * array (out_ld in bf16 elements) -- the interest vectors go straight into the bf16 MMoE input (dmt_mmoe_fwd_bf16in). */
#define DMT_SEQ_OUT_BF16 4

typedef struct dmt_seq_cfg {
  int32_t batch;        /* B                                                          */
  int32_t d_model;      /* transformer_d_model == sum of the pair dims                */
  int32_t d_ff;         /* transformer_d_ff                                           */
  int32_t num_heads;    /* transformer_num_heads; head j = columns [j*d_k,(j+1)*d_k)  */
  int32_t n_enc_blocks; /* transformer_num_blocks_encode                              */
  int32_t n_dec_blocks; /* transformer_num_blocks_decode                              */
  int32_t maxlen;       /* transformer_maxlen_k: rows of the learned position table   */
  int32_t zero_pad;     /* != 0: index i reads variable row i-1, 0 -> zero vector
                           (base.py:87-89); 0: index i reads row i                    */
  int32_t n_feats;      /* pairs in this sequence's attention_embed group             */
  int32_t precision;    /* dmt_precision                                              */
  int32_t slot_len;     /* upper bound on the sequence lengths in THIS batch (0 = maxlen): sizes the
                           slot tiles of the training pipeline's tensor-core attention (with
                           DMT_SEQ_LEN_EXACT).  The bf16 inference kernels no longer need it: they
                           pick a 16 / 32 / 64-row slot per SAMPLE on the device and truncate
                           sequences longer than maxlen, like the reference's position table  */
  int32_t flags;        /* DMT_SEQ_* bits (0 = none)                                 */
  /* training entry points only (dmt_seq_encode_fwd_train / _bwd): transformer_dropout_rate applied at the
     encoder input, the decoder input and the attention probabilities (TransformerModel.py:101,151;
     TransformerModel_util.py:51).  The keep mask is a counter-based hash of (dropout_seed, site, element),
     recomputed by the backward; pass the same cfg to both.  0 = no dropout. */
  float dropout_rate;
  uint32_t dropout_seed;
} dmt_seq_cfg;

/* per-sequence inputs: generate_data(), mmoe_transformer_unbias.py:130-186 */
typedef struct dmt_seq_input {
  const float* table[DMT_MAX_SEQ_FEATS];     /* embedding variable [rows, dim] fp32    */
  int64_t rows[DMT_MAX_SEQ_FEATS];
  int32_t dim[DMT_MAX_SEQ_FEATS];
  int32_t _pad[DMT_MAX_SEQ_FEATS];
  const int32_t* ids[DMT_MAX_SEQ_FEATS];     /* user feature CSR values                */
  const int32_t* offsets[DMT_MAX_SEQ_FEATS]; /* user feature CSR offsets [B+1]; the
                                                LAST pair's lengths are the sequence
                                                lengths (:141-146,183)                 */
  const int32_t* item_ids[DMT_MAX_SEQ_FEATS];/* target item index, one per sample [B]  */
} dmt_seq_input;

typedef struct dmt_seq_weights {
  const float* pos; /* positional_encoding_learn table [maxlen, d_model] (util:281-316) */
  dmt_attn_weights enc_attn[DMT_MAX_BLOCKS]; /* num_blocks_i/self-attention            */
  dmt_attn_weights dec_attn[DMT_MAX_BLOCKS]; /* num_blocks_i/vanilla_attention         */
  dmt_ff_weights ff[DMT_MAX_BLOCKS];         /* num_blocks_i/positionwise_feedforward  */
} dmt_seq_weights;

/* one pooled feature: tf.nn.embedding_lookup_sparse(table, ids, weights, 'mean') */
typedef struct dmt_pool_feat {
  const float* table;     /* raw variable: index i reads row i (base.py:115-116)       */
  int64_t rows;
  const int32_t* ids;     /* CSR values                                                */
  const int32_t* offsets; /* CSR offsets [B+1]                                         */
  const float* weights;   /* `<feature>Wts` values [nnz] or NULL (all ones)            */
  int32_t dim;
  int32_t out_col;        /* first output column                                       */
} dmt_pool_feat;

typedef struct dmt_mmoe_cfg {
  int32_t batch;
  int32_t in_dim;                       /* 615 + pooled + n_seq*d_model (1199)         */
  int32_t n_experts;                    /* num_experts                                 */
  int32_t n_layers;                     /* len(hidden_units_bottom)                    */
  int32_t units[DMT_MAX_LAYERS];        /* hidden_units_bottom                         */
  int32_t n_tasks;                      /* 2: click, order                             */
  int32_t n_tower_layers;               /* len(hidden_units_task)                      */
  int32_t tower_units[DMT_MAX_LAYERS];  /* hidden_units_task                           */
  int32_t precision;                    /* dmt_precision                               */
} dmt_mmoe_cfg;

typedef struct dmt_mmoe_weights {
  dmt_dense expert[DMT_MAX_EXPERTS][DMT_MAX_LAYERS]; /* expert-e/expert-layer-l        */
  dmt_dense gate[DMT_MAX_TASKS];                     /* gates-t/gates-layer-0 [in, E]  */
  dmt_dense tower[DMT_MAX_TASKS][DMT_MAX_LAYERS];    /* <task>-fc-l                    */
  dmt_dense tower_out[DMT_MAX_TASKS];                /* <task>-output [units, 1]       */
} dmt_mmoe_weights;

typedef struct dmt_bias_loss_cfg {
  int32_t batch;
  int32_t in_dim;                      /* sum of emb_bias dims (20)                    */
  int32_t n_hidden;                    /* len(hidden_units_bias)                       */
  int32_t units[DMT_MAX_LAYERS];       /* hidden_units_bias                            */
  int32_t two_head_multiply;           /* loss_unbias_method == two_head_multiply      */
  int32_t ctr_rel;                     /* loss_ctr_rel_method == ctr_rel               */
  float weight_ctr[5];                 /* [class_weight] weight_ctr by ascending label */
  float weight_ecvr[5];                /* [class_weight] weight_ecvr                   */
  float loss_weight[2];                /* [parameter] loss_weight                      */
  /* training mode: dropout_rate_bias after each hidden layer (mmoe_transformer_unbias.py:272,280); same
     hash mask as dmt_seq_cfg, pass the same cfg to dmt_bias_loss_fwd and dmt_bias_bwd.  0 = off. */
  float dropout_rate[DMT_MAX_LAYERS];
  uint32_t dropout_seed;
} dmt_bias_loss_cfg;

typedef struct dmt_bias_weights {
  dmt_dense layer[DMT_MAX_LAYERS + 1]; /* layer_bias0..n_hidden (last one -> 1 unit)   */
} dmt_bias_weights;

/* ---- library ------------------------------------------------------------------- */
DMT_API int dmt_abi_version(void);
DMT_API const char* dmt_last_error(void);
/* sha256 (hex) of the sources this library was built from (the .cu / .cuh files of csrc, this header, the nvcc flags): the
 * binding compares it with the digest of the sources it sits next to and refuses a stale library -- struct
 * layouts are part of those sources */
DMT_API const char* dmt_build_digest(void);
/* number of SMs of the current device (grid sizing); < 0 on error */
DMT_API int dmt_device_sm_count(void);

/* ---- A1/A2: raw row gather ------------------------------------------------------
 * Replaces tf.nn.embedding_lookup(embedding(..., zero_pad), ids)
 * (model/net/base.py:81-91; mmoe_transformer_unbias.py:153-158).  out[n, :] is the looked-up
 * row of ids[n]; with zero_pad index 0 -> zeros and index i -> table row i-1.  Bit-exact.
 * This is also the BASELINE config-5 vocabulary-sweep kernel. */
DMT_API int dmt_embed_gather(const float* table, int64_t rows, int32_t dim, const int32_t* ids,
                             int64_t n_ids, int32_t zero_pad, float* out, void* stream);

/* ---- A2-A8: one behaviour sequence ------------------------------------------------
 * Replaces generate_data (mmoe_transformer_unbias.py:130-186) + TransformerModel.encode_decode
 * (TransformerModel.py:51-171) + multihead_attention / ff / ln
 * (TransformerModel_util.py:11-108,160-235,281-316) for ONE sequence, eval mode
 * (dropout sites inactive).  Writes the interest vector of sample b to
 * out[b*out_ld + 0 .. d_model).  Samples are independent; padded positions are never
 * computed (they are inert in the reference). */
/* DMT_PRECISION_BF16: bytes of the per-sequence workspace = [bf16 weight images | batch-sized scratch: one
 * bf16 decoder-context image per 128 samples].  Grows with cfg->batch; 16-byte aligned; one buffer per
 * sequence (the images are that sequence's weights).  DMT_PRECISION_F32: 256 (nothing is spilled). */
DMT_API size_t dmt_seq_encode_workspace_bytes(const dmt_seq_cfg* cfg, int64_t max_tokens);
/* DMT_PRECISION_BF16 only: convert this sequence's transformer weights to the bf16 shared-memory
 * images the tensor-core kernels keep resident, at the front of the workspace (call again whenever
 * the weights change or the workspace is re-allocated).  The same buffer is then passed as
 * `workspace` to dmt_seq_encode_fwd. */
DMT_API int dmt_seq_prepare_weights(const dmt_seq_cfg* cfg, const dmt_seq_weights* w, void* prepared,
                                    size_t prepared_bytes, void* stream);
DMT_API int dmt_seq_encode_fwd(const dmt_seq_cfg* cfg, const dmt_seq_input* in,
                               const dmt_seq_weights* w, float* out, int64_t out_ld,
                               void* workspace, size_t workspace_bytes, void* stream);
/* DMT_PRECISION_BF16: the decoder tails of n_seq (<= DMT_MAX_TAIL_SEQS) sequences whose dmt_seq_encode_fwd ran with
 * DMT_SEQ_DEFER_TAIL, one launch (TransformerModel.py:157-171 for every sample of every sequence).  Element i of
 * each array is the argument the deferred call of sequence i was given. */
#define DMT_MAX_TAIL_SEQS 4
DMT_API int dmt_seq_tail_fwd(int32_t n_seq, const dmt_seq_cfg* const* cfgs, const dmt_seq_input* const* ins,
                             const dmt_seq_weights* const* ws, float* const* outs, const int64_t* out_lds,
                             void* const* workspaces, void* stream);
/* DMT_PRECISION_BF16: A2-A8 for ALL behaviour sequences of the step (trans_core x 3, mmoe_transformer_unbias.py:
 * 150-216) as three launches in total: the samples of every sequence are ordered by length class on the device
 * (> 32 | 17..32 | <= 16 tokens -> 2 / 4 / 8 samples per 128-row tile instead of padding everything to the longest
 * history), ONE persistent tile kernel walks the tiles of all (sequence, class) segments, one launch runs the decoder
 * tails.  Arguments as dmt_seq_tail_fwd; each workspace is that sequence's prepared buffer
 * (dmt_seq_encode_workspace_bytes / dmt_seq_prepare_weights).  Results per sample are those of dmt_seq_encode_fwd
 * up to the summation order of the softmax (bf16 tolerance). */
DMT_API int dmt_seq_encode_multi_fwd(int32_t n_seq, const dmt_seq_cfg* const* cfgs, const dmt_seq_input* const* ins,
                                     const dmt_seq_weights* const* ws, float* const* outs, const int64_t* out_lds,
                                     void* const* workspaces, const size_t* workspace_bytes, void* stream);

/* ---- A9/A11: pooled embeddings ----------------------------------------------------
 * Replaces embedding_combiner (model/net/base.py:93-124) and embedding_combiner_bias
 * (mmoe_transformer_unbias.py:235-257): out[b, out_col + j] = sum_t w_t row(ids_t)[j] / sum_t w_t.
 * `feats` is a HOST array of n_feats <= DMT_MAX_POOL_FEATS descriptors. */
DMT_API int dmt_pool_mean_fwd(int32_t batch, int32_t n_feats, const dmt_pool_feat* feats,
                              float* out, int64_t out_ld, void* stream);
/* same, bf16 output (out_ld in bf16 elements): the pooled columns of the bf16 MMoE input (dmt_mmoe_fwd_bf16in).
 * Needs row widths that are multiples of 4 and 16-byte aligned tables.  Features that share ONE offsets array
 * (pointer equality: the parallel id lists of one behaviour sequence) are walked together, one warp per sample. */
DMT_API int dmt_pool_mean_fwd_bf16(int32_t batch, int32_t n_feats, const dmt_pool_feat* feats,
                                   void* out_bf16, int64_t out_ld, void* stream);

/* strided 2-D copy of the dense `features` block into the MMoE input (base.py:95-96) */
DMT_API int dmt_copy_dense_features(const float* features, int32_t batch, int32_t dim,
                                    float* out, int64_t out_ld, void* stream);

/* ---- A0: compact host->device batch format ------------------------------------------
 * The reference feeds int64 ids and fp32 features (data_feed/tfrecord_mask.py:23-84,140-157).  The data loader of
 * this library may ship id arrays of small vocabularies (< 65536: Cid2, Cid3, Time*) as uint16 and the dense
 * `features` block as bf16; these two calls restore the int32 id arrays / fp32 feature columns the kernels read.
 * `arrays` is a HOST array; src/dst are device pointers, 16-byte aligned. */
#define DMT_MAX_WIDEN 64
typedef struct dmt_widen_desc {
  const uint16_t* src;
  int32_t* dst;
  int64_t n;
} dmt_widen_desc;
DMT_API int dmt_widen_u16(int32_t n_arrays, const dmt_widen_desc* arrays, void* stream);
/* the general form: every array carries its own width -- 1 byte (vocabularies < 256: the time buckets), 2 bytes
 * (< 65536) or 3 bytes little-endian (< 2^24: Sku, Brand, Shopid) per id */
typedef struct dmt_widen_ids_desc {
  const void* src;
  int32_t* dst;
  int64_t n;
  int32_t bytes;
  int32_t _pad;
} dmt_widen_ids_desc;
DMT_API int dmt_widen_ids(int32_t n_arrays, const dmt_widen_ids_desc* arrays, void* stream);
/* bf16 [batch, dim] (dense) -> fp32 columns [0, dim) of out (row stride out_ld); same role as
 * dmt_copy_dense_features (base.py:95-96) */
DMT_API int dmt_copy_dense_features_bf16(const void* features_bf16, int32_t batch, int32_t dim,
                                         float* out, int64_t out_ld, void* stream);
/* fp32 or bf16 [batch, dim] (dense) -> bf16 columns [0, dim) of the bf16 MMoE input (row stride
 * out_ld elements, even; out 4-byte aligned); base.py:95-96 for the bf16 tensor-core path */
DMT_API int dmt_stage_dense_features_bf16(const void* features, int32_t features_are_bf16, int32_t batch, int32_t dim,
                                          void* out_bf16, int64_t out_ld, void* stream);

/* ---- A10: MMoE experts + gates + task towers --------------------------------------
 * Replaces expert_gate + build_tower (mmoe_transformer_unbias.py:63-126,293-310).
 * logits is [n_tasks][batch] (task-major): logits[0] = click_logit, logits[1] = order_logit. */
DMT_API size_t dmt_mmoe_workspace_bytes(const dmt_mmoe_cfg* cfg);
/* DMT_PRECISION_BF16 only: bf16 K-major weight images of the expert layers for the tcgen05 GEMMs
 * (call again whenever the weights change); 256-byte aligned buffer of dmt_mmoe_prepared_bytes. */
DMT_API size_t dmt_mmoe_prepared_bytes(const dmt_mmoe_cfg* cfg);
DMT_API int dmt_mmoe_prepare_weights(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w,
                                     void* prepared, size_t prepared_bytes, void* stream);
/* `prepared` is NULL for DMT_PRECISION_F32. */
DMT_API int dmt_mmoe_fwd(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const float* x,
                         int64_t x_ld, float* logits, void* workspace, size_t workspace_bytes,
                         const void* prepared, void* stream);
/* DMT_PRECISION_BF16: the MMoE input is already bf16 -- [batch, xb_ld] with xb_ld a multiple of 8 and >= in_dim,
 * written in place by its producers (dmt_stage_dense_features_bf16, dmt_pool_mean_fwd_bf16, the sequence tails with
 * DMT_SEQ_OUT_BF16): no fp32 copy of x and no conversion pass.  The gate logits are extra output columns of the
 * layer-0 expert GEMM (gate kernels rounded to bf16 like the expert kernels, fp32 accumulation), softmaxed in its
 * epilogue.  Workspace / prepared as for dmt_mmoe_fwd. */
DMT_API int dmt_mmoe_fwd_bf16in(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const void* xb,
                                int64_t xb_ld, float* logits, void* workspace, size_t workspace_bytes,
                                const void* prepared, void* stream);

/* ---- the whole bf16 inference forward as one host call -------------------------------
 * Replaces mmoe_transformer_unbias.inference (mmoe_transformer_unbias.py:293-316, eval mode) for the bf16 tensor-core
 * path: the static part of every descriptor is built once (dmt_fwd_desc, all HOST memory, kept alive by the caller);
 * per call the caller passes one table of feature pointers and the dense block.  The function patches the descriptor
 * templates, runs the bias branch and the sequences on two library-owned side streams forked from / joined into
 * `stream`, and issues dmt_stage_dense_features_bf16, dmt_pool_mean_fwd(_bf16), dmt_seq_encode_multi_fwd,
 * dmt_mmoe_fwd_bf16in and dmt_bias_loss_fwd -- results are exactly those calls'.
 * One caller per device at a time: the two side streams and their events belong to the device, not to the call (two
 * host threads that drive the same GPU must serialise their calls or use the per-operator entry points).
 *   feats   [n_features] ids / offsets / weights (NULL = unit weights) of every CSR feature, device pointers
 *   scores  [n_tasks + 1][batch]: task logits, then y_bias (not written with is_predict) */
typedef struct dmt_fwd_feature {
  const int32_t* ids;
  const int32_t* offsets;
  const float* weights;
} dmt_fwd_feature;

typedef struct dmt_fwd_desc {
  int32_t batch;
  int32_t n_seq;
  int32_t n_pool;                       /* lookups of embedding_combiner (base.py:93-124)                  */
  int32_t n_bias_pool;                  /* lookups of embedding_combiner_bias (:235-257)                    */
  int32_t feature_dim;                  /* dense block width (0: none)                                      */
  int32_t is_predict;                   /* skip the bias branch                                             */
  int32_t interest_col;                 /* first column of the interest vectors in the MMoE input          */
  int32_t _pad;
  const dmt_pool_feat* pool;            /* templates: table / rows / dim / out_col filled                   */
  const int32_t* pool_feature;          /* lookup i reads feature pool_feature[i] of `feats`                */
  const dmt_pool_feat* bias_pool;
  const int32_t* bias_pool_feature;
  const dmt_seq_cfg* seq_cfg[DMT_MAX_TAIL_SEQS];        /* flags must hold DMT_SEQ_OUT_BF16                 */
  const dmt_seq_input* seq_in[DMT_MAX_TAIL_SEQS];       /* templates: table / rows / dim filled             */
  const int32_t* seq_user_feature[DMT_MAX_TAIL_SEQS];   /* pair f of sequence q: user feature index         */
  const int32_t* seq_item_feature[DMT_MAX_TAIL_SEQS];   /*                       item feature index         */
  const dmt_seq_weights* seq_w[DMT_MAX_TAIL_SEQS];
  void* seq_ws[DMT_MAX_TAIL_SEQS];                      /* prepared workspaces (dmt_seq_prepare_weights)    */
  size_t seq_ws_bytes[DMT_MAX_TAIL_SEQS];
  const dmt_mmoe_cfg* mmoe_cfg;
  const dmt_mmoe_weights* mmoe_w;
  void* mmoe_ws;
  size_t mmoe_ws_bytes;
  const void* mmoe_prepared;
  const dmt_bias_loss_cfg* bias_cfg;
  const dmt_bias_weights* bias_w;
  float* bias_in;                       /* [batch, bias_ld] scratch: pooled bias embeddings                 */
  int64_t bias_ld;
  void* xb;                             /* [batch, xb_ld] bf16 MMoE input (scratch)                         */
  int64_t xb_ld;
  void* inputs_ready;                   /* optional cudaEvent_t: the batch's arrays are on the device once it
                                           has fired (the prefetch copy).  Set per call; lets the length classes
                                           of this call be formed under the previous call's MMoE.  NULL: the
                                           inputs are ordered by `stream` like everything else                 */
} dmt_fwd_desc;

DMT_API int dmt_forward_bf16(const dmt_fwd_desc* desc, int32_t n_features, const dmt_fwd_feature* feats,
                             const void* features, int32_t features_are_bf16, float* scores, void* stream);

/* ---- A11/A12: bias tower + unbiased multi-task loss -------------------------------
 * Replaces embedding_mlp_bias (mmoe_transformer_unbias.py:259-289, eval mode),
 * cal_ctr_cvr_unibas (run_dnn.py:90-100) and logit_loss_unbias + cal_cross_entropy
 * (model/inference_mlp.py:162-223).
 *   bias_in  [B, in_dim] pooled bias embeddings (dmt_pool_mean_fwd output); with
 *            cfg->n_hidden < 0 the tower is skipped and bias_in[b*bias_ld] IS y_bias
 *            (loss on logits that were produced earlier)
 *   logits   [2][B] from dmt_mmoe_fwd
 *   mask     [B, 5] one-hot over labels (0,1,2,4,5), or NULL to skip the loss
 *   y_bias   [B]   out
 *   probs    [2][B] out: p_ctr, p_cvr (NULL to skip)
 *   loss     device scalar out (NULL to skip); needs loss_scratch of
 *            dmt_loss_scratch_bytes(batch) bytes
 *   dlogits  [3][B] out: dLoss/d{click_logit, order_logit, y_bias} (NULL to skip) */
DMT_API size_t dmt_loss_scratch_bytes(int32_t batch);
DMT_API int dmt_bias_loss_fwd(const dmt_bias_loss_cfg* cfg, const dmt_bias_weights* w,
                              const float* bias_in, int64_t bias_ld, const float* logits,
                              const float* mask, float* y_bias, float* probs, float* loss,
                              float* dlogits, void* loss_scratch, void* stream);

/* ---- A13: training forward / backward ----------------------------------------------
 * tf.gradients of the graph built by inference + loss_multi_task_unbias (run_dnn.py:181,
 * compute_gradients).  fp32 arithmetic (DMT_PRECISION_F32) or fp32 storage with the GEMMs on bf16 tensor
 * cores (DMT_PRECISION_BF16).  The training forward writes the activations the backward needs
 * into a caller-owned `saved` buffer (HBM is cheap on this part: ~3.6 KB per token); the backward is a
 * short pipeline of row-batched kernels (LayerNorm / attention backward, grouped GEMMs with fixed-order
 * split-K) that ACCUMULATES every parameter gradient into the buffers named by a `*_grads` descriptor --
 * same layout as the weight descriptor, pointing into a gradient buffer the caller zeroed at the start of
 * the step (the feed-forward variables are shared by encoder and decoder block i, so they receive two
 * contributions).  No floating-point atomics: gradients are run-to-run deterministic.
 * Dropout sites are inactive (rate 0 / eval mode), as in the forward. */
typedef dmt_seq_weights dmt_seq_grads;   /* pointers are written (+=) */
typedef dmt_mmoe_weights dmt_mmoe_grads;
typedef dmt_bias_weights dmt_bias_grads;

/* n_tokens = nnz of the sequence's id features (all pairs of one sequence have equal lengths in the data). */
DMT_API size_t dmt_seq_saved_bytes(const dmt_seq_cfg* cfg, int64_t n_tokens);
DMT_API int dmt_seq_encode_fwd_train(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w,
                                     float* out, int64_t out_ld, int64_t n_tokens, void* saved,
                                     size_t saved_bytes, void* stream);
DMT_API size_t dmt_seq_bwd_workspace_bytes(const dmt_seq_cfg* cfg, int64_t n_tokens);
/* d_out     [B, d_model] (row stride d_out_ld): dLoss/d(interest vector)
 * d_tokens  [n_tokens, d_model] out: dLoss/d(looked-up row) of every sequence token, columns in pair
 *           order (feeds dmt_grad_source with id_offset -1, grad_col = the pair's first column)
 * d_target  [B, d_model] out: dLoss/d(target-item rows) */
DMT_API int dmt_seq_encode_bwd(const dmt_seq_cfg* cfg, const dmt_seq_input* in, const dmt_seq_weights* w,
                               int64_t n_tokens, const void* saved, size_t saved_bytes, const float* d_out,
                               int64_t d_out_ld, const dmt_seq_grads* grads, float* d_tokens, float* d_target,
                               void* workspace, size_t workspace_bytes, void* stream);

/* Training forward of the MMoE block: like dmt_mmoe_fwd, but every expert layer's activations stay in
 * `workspace` as fp32 [layer][expert][B][units] for the backward; with DMT_PRECISION_BF16 the GEMMs of the
 * training path (forward and backward) run on tcgen05 with bf16 operands / fp32 accumulation. */
DMT_API size_t dmt_mmoe_train_workspace_bytes(const dmt_mmoe_cfg* cfg);
DMT_API int dmt_mmoe_fwd_train(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const float* x, int64_t x_ld,
                               float* logits, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of dmt_mmoe_fwd_train (`fwd_workspace` = its workspace).
 *   dlogits  [n_tasks][B]
 *   dx       [B, dx_ld] out: columns [dx_col0, in_dim) are written (the dense `features` block in front
 *            of dx_col0 is an input, not a variable) */
DMT_API size_t dmt_mmoe_bwd_workspace_bytes(const dmt_mmoe_cfg* cfg);
DMT_API int dmt_mmoe_bwd(const dmt_mmoe_cfg* cfg, const dmt_mmoe_weights* w, const float* x, int64_t x_ld,
                         const void* fwd_workspace, const float* dlogits, const dmt_mmoe_grads* grads, float* dx,
                         int64_t dx_ld, int32_t dx_col0, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the bias tower (mmoe_transformer_unbias.py:259-289): dy_bias [B] -> d_bias_in [B, in_dim]
 * (row stride d_in_ld) + the layer_bias* gradients. */
DMT_API size_t dmt_bias_bwd_workspace_bytes(const dmt_bias_loss_cfg* cfg);
DMT_API int dmt_bias_bwd(const dmt_bias_loss_cfg* cfg, const dmt_bias_weights* w, const float* bias_in,
                         int64_t bias_ld, const float* dy_bias, const dmt_bias_grads* grads, float* d_bias_in,
                         int64_t d_in_ld, void* workspace, size_t workspace_bytes, void* stream);

/* ---- A13: optimizer ----------------------------------------------------------------
 * tf.train.AdamOptimizer(lr) (model/inference_mlp.py:272-273) applied to the tower-averaged gradients
 * (run_dnn.py:45-80,203,207).  TF-1 form: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps),
 * DENSE over every row of every table (the reference densifies IndexedSlices, run_dnn.py:63-72). */
typedef struct dmt_adam_cfg {
  float lr;       /* piecewise_constant(global_step, step_boundary, learning_rate), run_dnn.py:125-126 */
  float beta1;    /* 0.9   */
  float beta2;    /* 0.999 */
  float epsilon;  /* 1e-8  */
  int32_t step;   /* t >= 1 */
  int32_t kind;   /* which optimizer of inference_mlp.py:264-277 the update kernels apply:
                     0 = tf.train.AdamOptimizer (dmt.conf), 1 = GradientDescentOptimizer (theta -= lr g),
                     2 = AdagradOptimizer (acc += g^2, theta -= lr g / sqrt(acc); acc lives in `m`, initial value 0.1).
                     1 and 2 leave a row without gradient unchanged, so the untouched-rows passes are no-ops. */
} dmt_adam_cfg;
#define DMT_OPT_ADAM 0
#define DMT_OPT_SGD 1
#define DMT_OPT_ADAGRAD 2

/* One group of lookups into a table whose gradient is needed: the backward of tf.nn.embedding_lookup
 * (sequence / target path, base.py:81-91: id_offset = -1 with zero_pad, index 0 has no gradient) or of
 * tf.nn.embedding_lookup_sparse(combiner='mean') (pooled path, base.py:116: id_offset = 0, mean = 1). */
typedef struct dmt_grad_source {
  const int32_t* ids;     /* [n] lookup indices                                                     */
  const int32_t* offsets; /* CSR offsets [batch+1] when `grad` holds ONE ROW PER SAMPLE (pooled and
                             target lookups); NULL when it holds one row per lookup (sequence tokens) */
  const float* weights;   /* `<feature>Wts` [n] or NULL (pooled path)                               */
  const float* grad;      /* gradient rows, row stride grad_ld floats; this table's slice starts at
                             column grad_col                                                         */
  int64_t n;
  int64_t grad_ld;
  int32_t grad_col;
  int32_t id_offset;      /* variable row = id + id_offset; rows outside [0, V) carry no gradient   */
  int32_t batch;
  int32_t mean;           /* != 0: scale each lookup by w / sum_w of its sample                     */
} dmt_grad_source;

/* theta/m/v/grad: flat fp32 buffers of n elements (16-byte aligned); g = grad * grad_scale. */
DMT_API int dmt_adam_dense(const dmt_adam_cfg* cfg, float* param, float* m, float* v, const float* grad,
                           int64_t n, float grad_scale, void* stream);
/* K9 step 1: expand every lookup into (row key | INT32_MAX if none, gradient-row reference, scale).
 * keys/refs/scale have sum(sources[i].n) entries, in source order.  `sources` is a HOST array. */
DMT_API int dmt_embed_grad_expand(int32_t n_sources, const dmt_grad_source* sources, int64_t rows,
                                  int32_t* keys, int64_t* refs, float* scale, void* stream);
/* K9 step 2 + K10: `sorted_keys` ascending with `perm` (sorted position -> expanded position) from any
 * stable device sort; the gradient rows of every run of equal keys are summed in sorted order by a chunked
 * two-pass segmented reduction (deterministic; a row hit 10^5 times does not serialise) and the Adam update
 * is applied to that row; touched[row] = 1.  workspace: dmt_embed_sorted_workspace_bytes(n, dim). */
DMT_API size_t dmt_embed_sorted_workspace_bytes(int64_t n, int32_t dim);
DMT_API int dmt_embed_adam_sorted(const dmt_adam_cfg* cfg, float* table, float* m, float* v, int64_t rows,
                                  int32_t dim, int32_t n_sources, const dmt_grad_source* sources,
                                  const int32_t* sorted_keys, const int64_t* perm, const int64_t* refs,
                                  const float* scale, int64_t n, float grad_scale, uint8_t* touched,
                                  void* workspace, size_t workspace_bytes, void* stream);
/* Data-parallel variants of K9.
 * densify: same segmented reduction as dmt_embed_adam_sorted, but the summed gradient row (x grad_scale) is
 *   written to dense_out [rows, dim] (zeroed by the caller) -- the replicated small tables join the dense
 *   allreduce bucket this way, exactly what average_gradients does (run_dnn.py:63-72).
 * scatter_rows: out[key, :] = scale * gradient row for lookups with UNIQUE keys (keys/refs/scale straight from
 *   dmt_embed_grad_expand, unsorted): the per-row gradients of a row-sharded table's compact copy, ready for
 *   the all-to-all back to the owning ranks. */
DMT_API int dmt_embed_grad_densify_sorted(int64_t rows, int32_t dim, int32_t n_sources,
                                          const dmt_grad_source* sources, const int32_t* sorted_keys,
                                          const int64_t* perm, const int64_t* refs, const float* scale, int64_t n,
                                          float grad_scale, float* dense_out, void* workspace,
                                          size_t workspace_bytes, void* stream);
DMT_API int dmt_embed_grad_scatter_rows(int32_t n_sources, const dmt_grad_source* sources, const int32_t* keys,
                                        const int64_t* refs, const float* scale, int64_t n, int32_t dim, float* out,
                                        void* stream);
/* K10 for the rows without a gradient this step (g = 0 still moves them); clears `touched`. */
DMT_API int dmt_adam_rows_untouched(const dmt_adam_cfg* cfg, float* table, float* m, float* v,
                                    int64_t rows, int32_t dim, uint8_t* touched, void* stream);

/* Multi-table variants of the three calls above: every embedding table of one optimizer step (run_dnn.py:203-207
 * applies ONE Adam op over all variables) in one expand, one key sort (caller), one segmented-reduction + Adam pair
 * and one untouched-rows pass.  Keys are (table index << 24) | row, so tables hold at most 2^24 rows here; `sources`
 * and `tables` are HOST arrays, source_table[s] = the table source s looks up.  dmt_adam_rows_untouched_multi leaves
 * the `touched` marks set: the caller clears them (one memset when they live in one buffer). */
#define DMT_MAX_ADAM_TABLES 16
#define DMT_MAX_MULTI_GRAD_SOURCES 64
typedef struct dmt_adam_table {
  float* table;
  float* m;
  float* v;
  uint8_t* touched;   /* [rows] */
  int64_t rows;
  int32_t dim;
  int32_t _pad;
  float* dense_out;   /* dmt_embed_adam_sorted_multi: non-NULL = write the summed gradient rows (x grad_scale) into this
                         zero-initialised [rows, dim] matrix instead of applying Adam (the replicated tables of a
                         data-parallel step: densified into the allreduce bucket); table / m / v / touched unused */
} dmt_adam_table;
DMT_API int dmt_embed_grad_expand_multi(int32_t n_tables, const dmt_adam_table* tables, int32_t n_sources,
                                        const dmt_grad_source* sources, const int32_t* source_table, int32_t* keys,
                                        int64_t* refs, float* scale, void* stream);
DMT_API size_t dmt_embed_sorted_multi_workspace_bytes(int64_t n);
DMT_API int dmt_embed_adam_sorted_multi(const dmt_adam_cfg* cfg, int32_t n_tables, const dmt_adam_table* tables,
                                        int32_t n_sources, const dmt_grad_source* sources,
                                        const int32_t* sorted_keys, const int64_t* perm, const int64_t* refs,
                                        const float* scale, int64_t n, float grad_scale, void* workspace,
                                        size_t workspace_bytes, void* stream);
DMT_API int dmt_adam_rows_untouched_multi(const dmt_adam_cfg* cfg, int32_t n_tables, const dmt_adam_table* tables,
                                          void* stream);

/* ---- diagnostics -------------------------------------------------------------------
 * One 128 x N x K bf16 tcgen05 GEMM through each shared-memory operand layout the tensor-core
 * kernels use (mode 0: K-major/K-major no-swizzle, A [128,K], B [N,K]; mode 1: B MN-major,
 * B [K,N]; mode 2: both SWIZZLE_128B K-major, K % 64 == 0).  C [128,N] fp32.  Exists so that the
 * descriptor encodings are pinned by a unit test; not part of the model path. */
DMT_API int dmt_selftest_umma(int32_t mode, const void* A, const void* B, float* C, int32_t N,
                              int32_t K, void* stream);

/* The TMA-fed tf32 tcgen05 GEMM engine of the training pipeline (DMT_PRECISION_TF32), exposed on raw device
 * pointers so that its shared-memory descriptors are pinned by unit tests (tests/test_gpu_tf32.py):
 *   rows  : C[M,N] (+)= mask(relu((A[M,K] Bt[N,K]^T + addend) * alpha + bias))     N % 16 == 0, N <= 256, K % 4 == 0
 *   wgrad : C (+)= P[T,MA]^T Q[T,NB]  (transposed: C[n][m], else C[m][n]); workspace of .._wgrad_bytes
 *   colsum: out[W] (+)= column sums of X[T,W]; scratch >= 296 * W floats
 * All matrices fp32 row-major, 16-byte aligned, row strides multiples of 4 floats. */
DMT_API int dmt_selftest_tf32_rows(const float* A, int64_t lda, const float* Bt, int64_t ldb, int64_t M, int32_t N,
                                   int32_t K, float* C, int64_t ldc, const float* bias, const float* addend,
                                   int64_t ld_add, const float* mask, int64_t ld_mask, float alpha, int32_t relu,
                                   int32_t accumulate, void* stream);
/* gemm: C[M,N] (+)= mask(relu(op(A) op(B) + bias)); a_mn: A stored [K,M], else [M,K]; b_mn: B stored [K,N], else [N,K] */
DMT_API int dmt_selftest_tf32_gemm(const float* A, int64_t lda, int32_t a_mn, const float* B, int64_t ldb,
                                   int32_t b_mn, int64_t M, int32_t N, int32_t K, float* C, int64_t ldc,
                                   const float* bias, const float* mask, int64_t ld_mask, int32_t relu,
                                   int32_t accumulate, void* stream);
DMT_API size_t dmt_selftest_tf32_wgrad_bytes(int64_t T, int32_t MA, int32_t NB);
DMT_API int dmt_selftest_tf32_wgrad(const float* P, int64_t ldp, const float* Q, int64_t ldq, int64_t T, int32_t MA,
                                    int32_t NB, float* C, int64_t ldc, int32_t transposed, int32_t accumulate,
                                    float* colsum /* [MA] (+)= column sums of P, or NULL */, void* workspace,
                                    void* stream);
DMT_API int dmt_selftest_tf32_colsum(const float* X, int64_t ldx, int64_t T, int32_t W, float* out,
                                     int32_t accumulate, void* scratch, void* stream);

/* Diagnostics: when `device_counters` (2048 x uint64 on the device, zeroed by the caller) is non-NULL, the bf16
 * sequence kernel runs its instrumented instantiation: [16 q + phase] SM cycles thread 0 of every CTA spends in each
 * phase of sequence q (gather-convert, QKV MMA, QKV epilogue, ... decoder; 12 = pipeline prime, 13 = drain, 14 = tiles,
 * 15 = sequence prologue), [64 + cta] cycles of the whole CTA, [320 + cta] / [576 + cta] %globaltimer at entry / exit,
 * [1024 + 128 group + 16 tile + phase] the phase timeline of both tile groups of CTA 0 (first six tiles).  Pass NULL
 * to switch it off.  Process-wide debug switch, not for production. */
DMT_API int dmt_debug_seq_profile(void* device_counters);
/* diagnostics: CUDA events around every seq_encode_multi_kernel launch (on its launch stream) while enabled;
 * _read waits for them and returns their summed duration and count (then keeps collecting). */
DMT_API int dmt_debug_seq_timer(int32_t enable);
DMT_API int dmt_debug_seq_timer_read(float* total_ms, int32_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* DMT_B200_H */
