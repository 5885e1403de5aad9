// Streaming GEMM engine of the training / row-batched path (A13; run_dnn.py:181 tf.gradients, the tf.layers.dense
// calls of TransformerModel_util.py:188-190,222-231): fp32 tensors in HBM, TMA straight into 128-byte-swizzled
// shared memory, tcgen05.mma kind::tf32 (fp32 operands, 10-bit mantissa used by the tensor core, fp32 accumulate in
// TMEM).  No conversion pass and no SIMT staging: the only per-element instructions are the epilogue's.
//
//   tf32_rows   C[M, N] = epilogue( A[M, K] . Bt[N, K]^T )      M = tokens (huge), N <= 256, K <= ~320
//               persistent CTAs, Bt resident in shared memory, A streamed through a TMA ring, two TMEM accumulators
//               (the epilogue of tile i overlaps the MMAs of tile i + 1); epilogue = addend, scale, bias, ReLU,
//               ReLU-mask of a saved activation, accumulate
//   tf32_wgrad  D[MA, NB] = P[T, MA]^T . Q[T, NB]               contraction over tokens (weight gradients): both
//               operands MN-major straight from the row-major activations, split over token ranges, fixed-order
//               reduction of the partials (deterministic), scattered into the (possibly transposed) weight tensors
//   tf32_colsum out[W] (+)= column sums of X[T, W]             bias gradients, two-stage fixed-order reduction
#pragma once
#include "dmt_common.cuh"

namespace dmt {

struct Tf32Rows {
  const float* A;      // [M, K] row-major, row stride lda (floats, multiple of 4), 16-byte aligned
  int64_t lda;
  const float* Bt;     // [N, K] row-major (element (n, k) = weight of input k -> output n), row stride ldb
  int64_t ldb;
  int64_t M;
  int N, K;            // N % 16 == 0 (N > 256 runs as several column blocks); K % 4 == 0
  float* C;            // [M, N], row stride ldc (multiple of 4), 16-byte aligned
  int64_t ldc;
  const float* bias;   // [N] or null
  const float* addend; // [M, N] (ld_add) or null:  v = (acc + addend) * alpha + bias
  int64_t ld_add;
  const float* mask;   // [M, N] (ld_mask) or null: v = mask > 0 ? v : 0
  int64_t ld_mask;
  float alpha;
  int relu;
  int accumulate;      // C += v
};

struct Tf32WgradSeg {  // rows [m0, m1) of D go to one weight tensor
  float* C;            // transposed == 0: C[(m - m0) * ldc + n];  transposed == 1: C[n * ldc + (m - m0)]
  int64_t ldc;
  int m0, m1;
  float* colsum;       // non-null: colsum[m - m0] (+)= sum_t P[t, m] (the bias gradient when P is the gradient
                       // matrix) -- rides along as one extra N = 16 MMA per K step against a block of ones
};

struct Tf32Wgrad {
  const float* P;      // [T, MA] row-major (ldp): D's row operand (the wider activation)
  int64_t ldp;
  const float* Q;      // [T, NB] row-major (ldq): D's column operand, NB % 16 == 0, NB <= 256
  int64_t ldq;
  int64_t T;
  int MA, NB;
  Tf32WgradSeg seg[4];
  int n_seg;
  int transposed;
  int accumulate;      // C += D (shared weights collect several contributions per step)
  float* partial;      // workspace of tf32_wgrad_partial_bytes(...)
};

// Weight packing (tiny, once per use): dst[dst_row + r][dst_col + c] = transpose ? W[c * ldw + r] : W[r * ldw + c]
// for r < rows, c < cols (rows / cols of the DESTINATION block); vectors: dstv[dst + i] = v[i].  The TF-layout
// kernels [in, out] become the K-major [out, in] operand `Bt` of tf32_rows, and the Q|K|V (K|V) projections are
// concatenated so that one GEMM serves them.
struct Tf32PackMat {
  const float* W;
  int64_t ldw;
  int rows, cols, transpose, dst_row, dst_col;
};
struct Tf32PackVec {
  const float* v;
  int n, dst;
};
int tf32_pack(const Tf32PackMat* mats, int n_mats, float* dst, int64_t ldd, const Tf32PackVec* vecs, int n_vecs,
              float* dstv, cudaStream_t st);

// General tiled GEMM (both operands streamed by TMA), batched over up to 4 independent problems `z` (the MMoE
// experts) or -- reduce_z -- summed over them into one output (dX of a layer whose output feeds several experts):
//   C_z[M, N] (+)= mask(relu( A_z . B_z + bias_z ))
// A_z: a_mn == 0: [M, K] row-major (lda);  a_mn == 1: stored [K, M] row-major (lda) -- the activation of a weight
//      gradient, contracted over its rows
// B_z: b_mn == 0: stored [N, K] row-major (ldb) -- a TF kernel used as dX operand;  b_mn == 1: stored [K, N]
//      row-major (ldb) -- a TF kernel in the forward, or the gradient matrix of a weight gradient
struct Tf32Gemm {
  const float* A[4];
  const float* B[4];
  int64_t lda, ldb;
  int a_mn, b_mn;
  int nz, reduce_z;
  int64_t M;
  int N, K;
  float* C[4];
  int64_t ldc;
  const float* bias[4];
  const float* mask[4];
  int64_t ld_mask;
  int relu, accumulate;
};
int tf32_gemm(const Tf32Gemm& p, cudaStream_t st);

size_t tf32_wgrad_partial_bytes(int64_t T, int MA, int NB);
int tf32_rows(const Tf32Rows& p, cudaStream_t st);
int tf32_wgrad(const Tf32Wgrad& p, cudaStream_t st);
// out[w] = (accumulate ? out[w] : 0) + sum_t X[t, w];  scratch: tf32_colsum_scratch_bytes(W)
size_t tf32_colsum_scratch_bytes(int W);
int tf32_colsum(const float* X, int64_t ldx, int64_t T, int W, float* out, int accumulate, float* scratch,
                cudaStream_t st);

}  // namespace dmt
