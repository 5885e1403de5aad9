// Shared between the host side (seq_tc_host.cu) and the kernels (seq_encode_tc3.cu) of the bf16 tensor-core sequence
// path: kernel arguments and the layout of the bf16 weight images written by dmt_seq_prepare_weights.
#pragma once
#include "dmt_common.cuh"
#include <cuda_bf16.h>

namespace dmt {

struct SeqTcArgs {
  dmt_seq_cfg cfg;
  dmt_seq_input in;
  const float* pos;
  // fp32 small vectors (global): biases + LayerNorm
  const float *bq, *bk, *bv, *ln1_g, *ln1_b;          // encoder self-attention
  const float *b1, *b2, *ln2_g, *ln2_b;               // feed-forward (shared enc/dec)
  const float *dbq, *dbk, *dbv, *ln3_g, *ln3_b;       // decoder vanilla attention
  const __nv_bfloat16* prepared;                      // bf16 weight images (see seq_prepare_kernel)
  float* out;
  int64_t out_ld;
  int32_t n_tiles;
  unsigned long long* dbg;                            // diagnostics: per-phase SM cycles (thread 0), or NULL
  void* ctx;                                          // v2: decoder attention contexts, one bf16 [k/8][128 samples][8]
                                                      //     A-operand image per 128 samples (workspace)
  int32_t chunk_feat[32];                             // 16-byte chunk c of a token -> feature pair
  int32_t chunk_off[32];                              //                          -> first column inside that row
};

// element counts (bf16 units) of the prepared images
__host__ __device__ constexpr size_t prep_wqkv(int D) { return (size_t)3 * D * D; }
__host__ __device__ constexpr size_t prep_w1(int D, int DFF) { return (size_t)D * DFF; }
// decoder block: G image (H*D outputs x D) | Wv image (D x D) | g fp32 [H*D]
__host__ __device__ constexpr size_t prep_dec(int D, int H) { return (size_t)H * D * D + (size_t)D * D + 2 * (size_t)H * D; }
// v2: block-diagonal decoder Wv image (D outputs x H*D): one K = H*D GEMM computes every head's ctx_h Wv_h
__host__ __device__ constexpr size_t prep_wvbd(int D, int H) { return (size_t)H * D * D; }
__host__ __device__ constexpr size_t prep_off_dec(int D, int DFF) { return prep_wqkv(D) + 2 * prep_w1(D, DFF); }
__host__ __device__ constexpr size_t prep_off_wvbd(int D, int DFF, int H) { return prep_off_dec(D, DFF) + prep_dec(D, H); }
__host__ __device__ constexpr size_t prep_total(int D, int DFF, int H) {
  return prep_off_wvbd(D, DFF, H) + prep_wvbd(D, H);
}

}  // namespace dmt
