// dmt_selftest_umma: one 128 x N x K tcgen05 GEMM through each operand layout the library uses.
// Diagnostics entry point -- the parity tests call it so that a descriptor mistake shows up as a
// failed unit test (or a trapped launch), not as a wrong model output.
#include "dmt_common.cuh"
#include "umma.cuh"

namespace dmt {

using namespace umma;

// mode 0: A K-major (no swizzle), B K-major (no swizzle), B given as [N][K]
// mode 1: A K-major (no swizzle), B MN-major (no swizzle), B given as [K][N]
// mode 2: A, B K-major SWIZZLE_128B (64-element k-blocks), B given as [N][K]
// mode 3: A in TENSOR MEMORY (row-owning threads pack bf16 pairs and tcgen05.st them), B K-major, [N][K]
// mode 4: A in tensor memory, B MN-major, B given as [K][N]
// mode 5: A K-major compact image of 16 rows (LBO = 256 B; rows >= 16 of the MMA read neighbouring bytes and
//         only produce garbage in their own accumulator rows), B MN-major [K][N]; rows 0..15 of C are checked
// mode 6: A MN-major (given as [K][128]: the image [m/8][k][8] of the memory rows, read with M contiguous), B K-major
//         compact image of N <= 16 rows given as [N][K] (the transposed decoder-context MMA of the sequence kernel)
__global__ void __launch_bounds__(128) umma_selftest_kernel(int mode, const __nv_bfloat16* __restrict__ A,
                                                            const __nv_bfloat16* __restrict__ B,
                                                            float* __restrict__ C, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;                                   // 128*K*2 bytes
  uint8_t* sB = smem + ((128 * K * 2 + 1023) & ~1023);   // N*K*2 bytes
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }

  // ---- stage A [128][K] (mode 6: A^T given as [K][128], image [m/8][k][8]) ----
  if (mode == 6) {
    for (int i = tid; i < K * 16; i += 128) {
      const int k = i % K, g = i / K;                   // 8 rows m = 8g..8g+7 of column k
      *reinterpret_cast<uint4*>(sA + g * (K * 16) + k * 16) = *reinterpret_cast<const uint4*>(A + (size_t)k * 128 + g * 8);
    }
  } else
  for (int i = tid; i < 128 * (K / 8); i += 128) {
    const int r = i % 128, c = i / 128;                 // 16-byte chunk c of row r
    const uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)r * K + c * 8);
    uint32_t off;
    if (mode == 2) off = (c / 8) * (128 * 128) + r * 128 + (((c % 8) ^ (r % 8)) * 16);
    else if (mode == 5) off = c * 256 + r * 16;
    else off = c * (128 * 16) + r * 16;
    if (mode == 5 && r >= 16) continue;
    *reinterpret_cast<uint4*>(sA + off) = v;
  }
  // ---- stage B ----
  const bool b_mn = (mode == 1 || mode == 4 || mode == 5);
  constexpr int kTmemA = 128;   // first column of the TMEM-resident A operand (modes 3, 4); C uses [0, N)
  if (b_mn) {                 // B given [K][N]; image [n/8][k][8]
    for (int i = tid; i < K * (N / 8); i += 128) {
      const int k = i % K, g = i / K;
      const uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)k * N + g * 8);
      *reinterpret_cast<uint4*>(sB + g * (K * 16) + k * 16) = v;
    }
  } else {                    // B given [N][K]
    for (int i = tid; i < N * (K / 8); i += 128) {
      const int r = i % N, c = i / N;
      const uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)r * K + c * 8);
      uint32_t off;
      if (mode == 2) off = (c / 8) * (N * 128) + r * 128 + (((c % 8) ^ (r % 8)) * 16);
      else off = c * (N * 16) + r * 16;
      *reinterpret_cast<uint4*>(sB + off) = v;
    }
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  if (mode == 3 || mode == 4) {   // A row `tid` -> lane tid, columns kTmemA + k/2
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        r[j] = (c0 + j < K / 2) ? *reinterpret_cast<const uint32_t*>(A + (size_t)tid * K + 2 * (c0 + j)) : 0u;
      tmem_st16(tmem_addr(tbase, kTmemA + c0), r);
    }
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }

  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, mode == 6, b_mn);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t da, db;
      if (mode == 2) {
        const uint32_t blk = ks / 4, sub = ks % 4;
        da = make_smem_desc(a0 + blk * (128 * 128) + sub * 32, 16, 1024, kLayoutSW128);
        db = make_smem_desc(b0 + blk * (N * 128) + sub * 32, 16, 1024, kLayoutSW128);
      } else {
        if (mode == 5) da = make_smem_desc(a0 + ks * 2 * 256, 256, 128, kLayoutNone);
        else if (mode == 6) da = make_smem_desc(a0 + ks * 2 * 128, 128, K * 16, kLayoutNone);   // MN-major A
        else da = make_smem_desc(a0 + ks * 2 * (128 * 16), 128 * 16, 128, kLayoutNone);
        if (!b_mn) db = make_smem_desc(b0 + ks * 2 * (N * 16), N * 16, 128, kLayoutNone);
        else db = make_smem_desc(b0 + ks * 2 * 128, 128, K * 16, kLayoutNone);   // MN-major: LBO = next 8 k, SBO = next 8 n
      }
      if (mode == 3 || mode == 4) mma_bf16_ts(tbase, tbase + kTmemA + ks * 8, db, idesc, ks > 0 ? 1u : 0u);
      else mma_bf16_ss(tbase, da, db, idesc, ks > 0 ? 1u : 0u);
    }
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();

  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem_addr(tbase, c0), r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < N) C[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

}  // namespace dmt

extern "C" int dmt_selftest_umma(int32_t mode, const void* A, const void* B, float* C, int32_t N, int32_t K,
                                 void* stream) {
  DMT_REQUIRE(A && B && C, DMT_ERR_INVALID_ARGUMENT, "dmt_selftest_umma: null pointer");
  DMT_REQUIRE(mode >= 0 && mode <= 6, DMT_ERR_INVALID_ARGUMENT, "dmt_selftest_umma: mode %d", mode);
  DMT_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K >= 16 && K % 16 == 0 && (mode != 2 || K % 64 == 0) &&
                  ((mode != 3 && mode != 4) || (N <= 128 && K <= 256)),
              DMT_ERR_UNSUPPORTED_SHAPE, "dmt_selftest_umma: N=%d K=%d", N, K);
  const size_t bytes = ((128 * (size_t)K * 2 + 1023) & ~(size_t)1023) + (size_t)N * K * 2 + 1024;
  DMT_REQUIRE(bytes <= 200 * 1024, DMT_ERR_UNSUPPORTED_SHAPE, "dmt_selftest_umma: tile needs %zu B", bytes);
  cudaError_t e = cudaFuncSetAttribute(dmt::umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return dmt::cuda_fail(e, "cudaFuncSetAttribute(umma_selftest_kernel)");
  dmt::umma_selftest_kernel<<<1, 128, bytes, (cudaStream_t)stream>>>(mode, (const __nv_bfloat16*)A,
                                                                     (const __nv_bfloat16*)B, C, N, K);
  DMT_CUDA_LAUNCH_CHECK("umma_selftest_kernel");
  return DMT_OK;
}
