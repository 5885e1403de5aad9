// dmt_forward_bf16: the whole bf16 inference forward (mmoe_transformer_unbias.py:293-316: generate_data + trans_core x n +
// embedding_combiner + expert_gate + build_tower + embedding_mlp_bias) as ONE host call.
//
// At 0.34 ms of GPU time per 4096-sample step the Python plugin's own work per call -- descriptor filling, twelve
// ctypes calls, stream / event bookkeeping: ~0.3 ms -- is the end-to-end limiter.  This driver is the native runtime
// of that path: the static part of every descriptor (tables, weights, workspaces, columns) lives in a dmt_fwd_desc
// built once per (model, batch size); per call the host passes one table of feature pointers (ids / offsets / weights,
// e.g. packed-buffer base + the layout offsets a PackedBatch records when it is built) and this function patches the
// descriptor templates, forks the three independent branches onto their streams and issues the same entry points a
// caller would:
//   side stream 1 : bias pooled lookups -> bias tower                      (dmt_pool_mean_fwd, dmt_bias_loss_fwd)
//   side stream 0 : length classes -> all sequences, one tile kernel -> tails   (dmt_seq_encode_multi_fwd)
//   main stream   : dense features, pooled lookups -> (join) MMoE              (dmt_stage_dense_features_bf16,
//                                                                                dmt_pool_mean_fwd_bf16, dmt_mmoe_fwd_bf16in)
// No device work of its own: results are those of the individual calls.
#include <string.h>

#include <mutex>

#include "dmt_common.cuh"

namespace dmt {
int seq_encode_multi_checked(int32_t n_seq, const dmt_seq_cfg* const* cfgs, const dmt_seq_input* const* ins,
                             const dmt_seq_weights* const* ws, float* const* outs, const int64_t* out_lds,
                             void* const* workspaces, const size_t* workspace_bytes, cudaEvent_t wait_before_encode,
                             void* stream);
namespace {

struct FwdStreams {
  bool ready = false;
  cudaStream_t side[2];
  cudaEvent_t fork, join[2];
};
constexpr int kMaxDevices = 64;
FwdStreams g_fwd[kMaxDevices];
std::mutex g_fwd_mutex;      // guards the lazy creation; the streams themselves serve ONE caller per device at a time

int fwd_streams(FwdStreams** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice(dmt_forward_bf16)");
  DMT_REQUIRE(dev >= 0 && dev < kMaxDevices, DMT_ERR_INVALID_ARGUMENT, "dmt_forward_bf16: device %d", dev);
  FwdStreams& s = g_fwd[dev];
  std::lock_guard<std::mutex> lock(g_fwd_mutex);
  if (!s.ready) {
    // The sequence stream has the highest priority: its persistent tile kernel needs whole SMs, and the short CTAs of
    // the dense / pooled / bias kernels that start beside it would otherwise keep it from becoming resident (its
    // duration is set by its LAST CTA to start); with the priority they fill in while it ramps up and as its CTAs
    // retire instead.
    int prio_lo = 0, prio_hi = 0;
    e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetStreamPriorityRange(dmt_forward_bf16)");
    for (int i = 0; i < 2; ++i) {
      e = cudaStreamCreateWithPriority(&s.side[i], cudaStreamNonBlocking, i == 0 ? prio_hi : prio_lo);
      if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreateWithPriority(dmt_forward_bf16)");
      e = cudaEventCreateWithFlags(&s.join[i], cudaEventDisableTiming);
      if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreateWithFlags(dmt_forward_bf16)");
    }
    e = cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreateWithFlags(dmt_forward_bf16)");
    s.ready = true;
  }
  *out = &s;
  return DMT_OK;
}

int patch_pool(const char* what, int n, const dmt_pool_feat* tmpl, const int32_t* feat_of, int n_features,
               const dmt_fwd_feature* feats, dmt_pool_feat* out) {
  for (int i = 0; i < n; ++i) {
    const int f = feat_of[i];
    DMT_REQUIRE(f >= 0 && f < n_features, DMT_ERR_INVALID_ARGUMENT, "dmt_forward_bf16: %s lookup %d names feature %d of %d",
                what, i, f, n_features);
    out[i] = tmpl[i];
    out[i].ids = feats[f].ids;
    out[i].offsets = feats[f].offsets;
    out[i].weights = feats[f].weights;
  }
  return DMT_OK;
}

#define DMT_CUDA_TRY(call, what)                             \
  do {                                                       \
    cudaError_t _e = (call);                                 \
    if (_e != cudaSuccess) return cuda_fail(_e, (what));     \
  } while (0)

}  // namespace
}  // namespace dmt

extern "C" int dmt_forward_bf16(const dmt_fwd_desc* d, int32_t n_features, const dmt_fwd_feature* feats,
                                const void* features, int32_t features_are_bf16, float* scores, void* stream) {
  using namespace dmt;
  DMT_REQUIRE(d && feats && scores, DMT_ERR_INVALID_ARGUMENT, "dmt_forward_bf16: null pointer");
  DMT_REQUIRE(d->batch > 0 && d->n_seq >= 0 && d->n_seq <= DMT_MAX_TAIL_SEQS && d->n_pool >= 0 &&
                  d->n_pool <= DMT_MAX_POOL_FEATS && d->n_bias_pool >= 0 && d->n_bias_pool <= DMT_MAX_POOL_FEATS,
              DMT_ERR_INVALID_ARGUMENT, "dmt_forward_bf16: batch=%d n_seq=%d n_pool=%d n_bias_pool=%d", d->batch, d->n_seq,
              d->n_pool, d->n_bias_pool);
  DMT_REQUIRE(d->mmoe_cfg && d->mmoe_w && d->xb && d->xb_ld > 0, DMT_ERR_INVALID_ARGUMENT,
              "dmt_forward_bf16: MMoE descriptors / input buffer missing");
  DMT_REQUIRE(d->feature_dim == 0 || features, DMT_ERR_INVALID_ARGUMENT, "dmt_forward_bf16: dense features missing");
  DMT_REQUIRE(d->is_predict || (d->bias_cfg && d->bias_w && d->bias_in), DMT_ERR_INVALID_ARGUMENT,
              "dmt_forward_bf16: bias-tower descriptors missing");
  for (int i = 0; i < n_features; ++i)
    DMT_REQUIRE(feats[i].ids && feats[i].offsets, DMT_ERR_INVALID_ARGUMENT, "dmt_forward_bf16: feature %d has no ids / offsets", i);
  FwdStreams* fs = nullptr;
  int rc = fwd_streams(&fs);
  if (rc != DMT_OK) return rc;
  cudaStream_t main = (cudaStream_t)stream;
  const int B = d->batch, T = d->mmoe_cfg->n_tasks;

  DMT_CUDA_TRY(cudaEventRecord(fs->fork, main), "cudaEventRecord(dmt_forward_bf16)");
  // ---- bias branch (reads nothing the rest writes) ----
  if (!d->is_predict) {
    cudaStream_t s1 = fs->side[1];
    DMT_CUDA_TRY(cudaStreamWaitEvent(s1, fs->fork, 0), "cudaStreamWaitEvent(dmt_forward_bf16)");
    dmt_pool_feat pf[DMT_MAX_POOL_FEATS];
    rc = patch_pool("bias", d->n_bias_pool, d->bias_pool, d->bias_pool_feature, n_features, feats, pf);
    if (rc != DMT_OK) return rc;
    if (d->n_bias_pool > 0) {
      rc = dmt_pool_mean_fwd(B, d->n_bias_pool, pf, d->bias_in, d->bias_ld, s1);
      if (rc != DMT_OK) return rc;
    }
    rc = dmt_bias_loss_fwd(d->bias_cfg, d->bias_w, d->bias_in, d->bias_ld, scores, nullptr, scores + (int64_t)T * B, nullptr,
                           nullptr, nullptr, nullptr, s1);
    if (rc != DMT_OK) return rc;
    DMT_CUDA_TRY(cudaEventRecord(fs->join[1], s1), "cudaEventRecord(dmt_forward_bf16)");
  }
  // ---- behaviour sequences ----
  if (d->n_seq > 0) {
    cudaStream_t s0 = fs->side[0];
    // With an `inputs_ready` event (a prefetched batch) the length-class kernel -- it only reads the batch's offsets
    // and writes this stream's own schedule buffers -- does not wait for the caller's stream: it runs as soon as the
    // previous call's tails are done, under the previous call's MMoE.  The tile kernel then waits for the fork.
    cudaEvent_t late_wait = nullptr;
    if (d->inputs_ready) {
      DMT_CUDA_TRY(cudaStreamWaitEvent(s0, (cudaEvent_t)d->inputs_ready, 0), "cudaStreamWaitEvent(dmt_forward_bf16)");
      late_wait = fs->fork;
    } else {
      DMT_CUDA_TRY(cudaStreamWaitEvent(s0, fs->fork, 0), "cudaStreamWaitEvent(dmt_forward_bf16)");
    }
    dmt_seq_input in[DMT_MAX_TAIL_SEQS];
    const dmt_seq_input* ins[DMT_MAX_TAIL_SEQS];
    float* outs[DMT_MAX_TAIL_SEQS];
    int64_t lds[DMT_MAX_TAIL_SEQS];
    for (int q = 0; q < d->n_seq; ++q) {
      DMT_REQUIRE(d->seq_cfg[q] && d->seq_in[q] && d->seq_w[q] && d->seq_user_feature[q] && d->seq_item_feature[q],
                  DMT_ERR_INVALID_ARGUMENT, "dmt_forward_bf16: sequence %d descriptors missing", q);
      DMT_REQUIRE(d->seq_cfg[q]->flags & DMT_SEQ_OUT_BF16, DMT_ERR_INVALID_ARGUMENT,
                  "dmt_forward_bf16: sequence %d must write bf16 interest vectors (DMT_SEQ_OUT_BF16)", q);
      in[q] = *d->seq_in[q];
      for (int f = 0; f < d->seq_cfg[q]->n_feats && f < DMT_MAX_SEQ_FEATS; ++f) {
        const int fu = d->seq_user_feature[q][f], fi = d->seq_item_feature[q][f];
        DMT_REQUIRE(fu >= 0 && fu < n_features && fi >= 0 && fi < n_features, DMT_ERR_INVALID_ARGUMENT,
                    "dmt_forward_bf16: sequence %d pair %d names features %d / %d of %d", q, f, fu, fi, n_features);
        in[q].ids[f] = feats[fu].ids;
        in[q].offsets[f] = feats[fu].offsets;
        in[q].item_ids[f] = feats[fi].ids;
      }
      ins[q] = &in[q];
      // (bf16 elements: the tails write with DMT_SEQ_OUT_BF16)
      outs[q] = reinterpret_cast<float*>(static_cast<uint16_t*>(d->xb) + d->interest_col + (int64_t)q * d->seq_cfg[q]->d_model);
      lds[q] = d->xb_ld;
    }
    rc = seq_encode_multi_checked(d->n_seq, d->seq_cfg, ins, d->seq_w, outs, lds, d->seq_ws, d->seq_ws_bytes, late_wait, s0);
    if (rc != DMT_OK) return rc;
    DMT_CUDA_TRY(cudaEventRecord(fs->join[0], s0), "cudaEventRecord(dmt_forward_bf16)");
  }
  // ---- dense block + pooled lookups, then the MMoE on the assembled bf16 input ----
  if (d->feature_dim > 0) {
    rc = dmt_stage_dense_features_bf16(features, features_are_bf16, B, d->feature_dim, d->xb, d->xb_ld, main);
    if (rc != DMT_OK) return rc;
  }
  if (d->n_pool > 0) {
    dmt_pool_feat pf[DMT_MAX_POOL_FEATS];
    rc = patch_pool("pooled", d->n_pool, d->pool, d->pool_feature, n_features, feats, pf);
    if (rc != DMT_OK) return rc;
    rc = dmt_pool_mean_fwd_bf16(B, d->n_pool, pf, d->xb, d->xb_ld, main);
    if (rc != DMT_OK) return rc;
  }
  if (d->n_seq > 0) DMT_CUDA_TRY(cudaStreamWaitEvent(main, fs->join[0], 0), "cudaStreamWaitEvent(dmt_forward_bf16)");
  rc = dmt_mmoe_fwd_bf16in(d->mmoe_cfg, d->mmoe_w, d->xb, d->xb_ld, scores, d->mmoe_ws, d->mmoe_ws_bytes, d->mmoe_prepared,
                           main);
  if (rc != DMT_OK) return rc;
  if (!d->is_predict) DMT_CUDA_TRY(cudaStreamWaitEvent(main, fs->join[1], 0), "cudaStreamWaitEvent(dmt_forward_bf16)");
  return DMT_OK;
}
