// Grouped fp32 SIMT GEMM used by the training path (backward of A3-A11).
//
// One launch executes a list of independent problems
//     C = epilogue( sum_p  op(A_p) [M,K_p] * op(B_p) [K_p,N] )
// where every operand is a strided fp32 matrix in either orientation, K may be split across CTAs
// (weight gradients contract over all tokens of the batch) and the epilogue covers what the backward
// needs: residual addend, scale, bias, ReLU, ReLU-mask of a saved activation, accumulate into C and the
// column sums of B (bias gradients).  Split-K partials are reduced in a fixed order by a second kernel, so
// every gradient is run-to-run deterministic (no floating-point atomics anywhere on the path).
#pragma once
#include "dmt_common.cuh"

namespace dmt {

constexpr int kGemmMaxParts = 8;
constexpr int kGemmMaxProbs = 12;

struct GemmPart {
  const float* A;
  const float* B;
  int64_t lda, ldb;
  int K;
  int _pad;
};

struct GemmProb {
  GemmPart part[kGemmMaxParts];
  int n_parts;
  int M, N;
  int transA;            // 0: A is [M,K] row-major (lda);  1: A is stored [K,M] (element (m,k) at A[k*lda+m])
  int transB;            // 0: B is [K,N] row-major (ldb);  1: B is stored [N,K] (element (k,n) at B[n*ldb+k])
  float* C;
  int64_t ldc;
  const float* addend;   // v += addend[m,n]           (ld_add)
  int64_t ld_add;
  float alpha;           // v *= alpha
  const float* bias;     // v += bias[n]
  int relu;              // v = max(v, 0)
  const float* mask;     // v = mask[m,n] > 0 ? v : 0  (ld_mask)
  int64_t ld_mask;
  int accumulate;        // C += v instead of C = v
  float* colsum;         // colsum[n] (+)= sum_k B[k,n] (single-part problems; bias gradients) or null
  int colsum_accumulate;
  int splits;            // K splits of part 0 (single-part problems only)
  float* partial;        // [splits][M + (colsum ? 1 : 0)][N] scratch when splits > 1
  // filled by the launcher
  int tiles_m, tiles_n, cta0, red0;
  int tc_bn;             // tensor-core path: tile width of this problem
};

struct GemmGroup {
  GemmProb p[kGemmMaxProbs];
  int n;
  int total_ctas;
  int total_red;
  int tc_nmax;  // filled by the tensor-core launcher: widest 16-padded tile of the group
  int tc_stages;   //                               shared-memory stages (1 when no CTA walks more than one K chunk)
  int use_tc;   // 0: fp32 SIMT; 1: bf16 operands on tcgen05 (gemm_tc.cu), fp32 accumulate; 3: bf16x3 split
                // operands (hi*hi + hi*lo + lo*hi) on tcgen05 -- fp32-grade results
};

// dmt_precision -> GemmGroup::use_tc (DMT_PRECISION_TF32: the grouped problems that have no tf32 route -- the MMoE
// GEMMs, the per-sample rows of the decoder -- run on the split-bf16 engine)
inline int gemm_engine(int precision) {
  return precision == DMT_PRECISION_BF16 ? 1
                                         : ((precision == DMT_PRECISION_BF16X3 || precision == DMT_PRECISION_TF32) ? 3 : 0);
}
inline bool gemm_tf32(int precision) { return precision == DMT_PRECISION_TF32; }

inline void gemm_prob_init(GemmProb& p) {
  p = GemmProb{};
  p.alpha = 1.0f;
  p.splits = 1;
}

// bytes of split-K scratch problem `p` needs (0 when splits == 1)
size_t gemm_partial_bytes(const GemmProb& p);
// choose a split count for a [M,N] += A^T B contraction over K rows so that the launch fills the chip
// (the two engines tile differently, so the choice depends on which one will run the problem)
int gemm_pick_splits(int M, int N, int64_t K, bool use_tc = false, bool colsum = true);
int gemm_group_launch(GemmGroup& g, cudaStream_t st);

}  // namespace dmt
