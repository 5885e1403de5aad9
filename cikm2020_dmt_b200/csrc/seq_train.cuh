// Activations the training forward of one behaviour sequence leaves in HBM for the backward
// (dmt_seq_encode_fwd_train -> dmt_seq_encode_bwd).  Token rows are indexed by the CSR position of the
// sequence's LAST pair (the one whose lengths are the sequence lengths, mmoe_transformer_unbias.py:141-146,183);
// rows of tokens beyond the on-chip cap are written as zeros so they are inert in every contraction.
#pragma once
#include "dmt_common.cuh"

namespace dmt {

struct SeqSaved {
  // encoder, per block (TransformerModel.py:103-121)
  float* hin[DMT_MAX_BLOCKS + 1];   // [T, d]   block input; hin[n_enc] = memory
  float* qkv[DMT_MAX_BLOCKS];       // [T, 3d]  Q | K | V projections
  float* z1[DMT_MAX_BLOCKS];        // [T, d]   attention context + residual (LayerNorm input)
  float* a[DMT_MAX_BLOCKS];         // [T, d]   LayerNorm output = FF input
  float* f1[DMT_MAX_BLOCKS];        // [T, dff] relu(a W1 + b1)
  float* z2[DMT_MAX_BLOCKS];        // [T, d]   f1 W2 + b2 + a (LayerNorm input)
  // decoder, per block (TransformerModel.py:153-168)
  float* din[DMT_MAX_BLOCKS + 1];   // [B, d]   block input; din[n_dec] = interest vector
  float* qd[DMT_MAX_BLOCKS];        // [B, d]
  float* kvd[DMT_MAX_BLOCKS];       // [T, 2d]  K | V projections of the memory
  float* pd[DMT_MAX_BLOCKS];        // [B, H, LP] attention probabilities
  float* z1d[DMT_MAX_BLOCKS];       // [B, d]
  float* ad[DMT_MAX_BLOCKS];        // [B, d]
  float* f1d[DMT_MAX_BLOCKS];       // [B, dff]
  float* z2d[DMT_MAX_BLOCKS];       // [B, d]
  float* pack;                      // DMT_PRECISION_TF32: packed / transposed weights of the GEMM in flight
};

// floats of the weight-pack scratch: the largest set alive at once is one block's Q|K|V + W1 + W2 packs
inline size_t seq_pack_floats(const dmt_seq_cfg& c) {
  const size_t d = c.d_model, dff = c.d_ff;
  return 3 * d * d + 2 * d * dff + 3 * d * d + 16 * d + 2 * dff + 1024;
}

inline int seq_lp(const dmt_seq_cfg& c) { return c.maxlen < DMT_MAX_SEQ_LEN ? c.maxlen : DMT_MAX_SEQ_LEN; }

struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* b) : base((char*)b), off(0) {}
  float* take(size_t floats) {
    float* p = base ? (float*)(base + off) : nullptr;
    off += (floats * sizeof(float) + 255) / 256 * 256;
    return p;
  }
};

// Lays the buffers out behind `base` (may be null: size query) and returns the bytes used.
inline size_t seq_saved_carve(const dmt_seq_cfg& c, int64_t T, void* base, SeqSaved* sv) {
  Carver cv(base);
  SeqSaved s{};
  const size_t d = c.d_model, dff = c.d_ff, B = c.batch, H = c.num_heads, LP = seq_lp(c);
  for (int b = 0; b <= c.n_enc_blocks; ++b) s.hin[b] = cv.take(T * d);
  for (int b = 0; b < c.n_enc_blocks; ++b) {
    s.qkv[b] = cv.take(T * 3 * d);
    s.z1[b] = cv.take(T * d);
    s.a[b] = cv.take(T * d);
    s.f1[b] = cv.take(T * dff);
    s.z2[b] = cv.take(T * d);
  }
  for (int b = 0; b <= c.n_dec_blocks; ++b) s.din[b] = cv.take(B * d);
  for (int b = 0; b < c.n_dec_blocks; ++b) {
    s.qd[b] = cv.take(B * d);
    s.kvd[b] = cv.take(T * 2 * d);
    s.pd[b] = cv.take(B * H * LP);
    s.z1d[b] = cv.take(B * d);
    s.ad[b] = cv.take(B * d);
    s.f1d[b] = cv.take(B * dff);
    s.z2d[b] = cv.take(B * d);
  }
  s.pack = cv.take(seq_pack_floats(c));
  if (sv) *sv = s;
  return cv.off + 256;
}

}  // namespace dmt
