// dmt_seq_encode_fwd, DMT_PRECISION_BF16 (v3 tile kernel + decoder tail kernel): two tiles in flight per SM,
// tensor-memory A operands, two threads per token row.
//
// Same math as seq_tc_host.cu (gather -> X Wqkv -> masked softmax attention -> LN -> FF -> LN -> decoder), but
// organised so that the tensor pipe, the shared-memory pipe and the SIMT pipes of one SM always have two
// independent tiles to work on:
//
//   * one CTA per SM = 2 GROUPS of 128 threads; each group runs the whole per-tile program on its own 128-row
//     tile (one thread = one token row = one TMEM lane), with its own mbarrier, named barrier, 256 TMEM columns
//     and 64 KB of shared memory.  The bf16 weight images (88 KB) are shared by both groups.
//   * the softmax probabilities P_h and the ReLU hidden H never touch shared memory: the row-owning thread packs
//     them to bf16 and writes them back with tcgen05.st IN PLACE over the fp32 accumulator columns it just read;
//     P_h V_h and H W2 are issued with the A operand in tensor memory.  That removes 96 KB of activation
//     buffers per tile, which is what makes the second tile fit.
//   * decoder (one query per sample): the key/query projections are folded into qt = dvec G_h + g_h (computed on
//     CUDA cores while the P V MMAs run); every token thread takes its own score from the memory row it holds in
//     registers; the softmax is a flash-style partial softmax per (warp segment, head) -- no cross-warp max; the
//     context sum_t p_t M_t is ONE more MMA: A = the partial probabilities transposed into a compact 16-row
//     image, B = the memory rows (bf16, MN-major) with a column of ones appended, so the same accumulator row
//     carries the softmax denominator.  The per-sample tail (ctx_h Wv_h + residual -> LN -> FF -> LN) is row-batched
//     over 128 SAMPLES per tile in seq_tail_kernel, on tensor cores, instead of mat-vec loops per sample.
//
// HBM traffic per (sample, sequence): ids + embedding rows in, 128 context floats out and in again (L2), one
// interest vector out.
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "dmt_common.cuh"
#include "seq_tc.cuh"
#include "umma.cuh"

namespace dmt {

using namespace umma;

namespace {

constexpr int kD = 64, kDFF = 256, kH = 2, kDK = 32, kKC = 8, kROWB = 128 * 16;
constexpr int kInvalidId = INT_MIN;

__device__ __forceinline__ void bf16x8_to_f(const uint4& v, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 f8_to_bf16(const float* f) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]);
  v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]);
  v.w = pack_bf16x2(f[6], f[7]);
  return v;
}

// streaming (read-once) 256-bit load: one 32-byte sector of an embedding row per lane, no L1 allocation
struct f8 { float4 lo, hi; };
__device__ __forceinline__ f8 ld_stream8(const float* p) {
  f8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
               : "l"(p));
  return r;
}

// Row-local LayerNorm over 64 values held in registers (TransformerModel_util.py:58-78); four independent
// partial sums keep the dependent-add chains short.
__device__ __forceinline__ void ln64(float* y, const float* __restrict__ g, const float* __restrict__ b) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < kD; i += 4) {
    s0 += y[i]; s1 += y[i + 1]; s2 += y[i + 2]; s3 += y[i + 3];
  }
  const float mean = ((s0 + s1) + (s2 + s3)) * (1.0f / kD);
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
  for (int i = 0; i < kD; i += 4) {
    const float d0 = y[i] - mean, d1 = y[i + 1] - mean, d2 = y[i + 2] - mean, d3 = y[i + 3] - mean;
    q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
  }
  const float rstd = 1.0f / sqrtf(((q0 + q1) + (q2 + q3)) * (1.0f / kD) + kLnEps);
#pragma unroll
  for (int i = 0; i < kD; i += 4) {
    const float4 gg = *reinterpret_cast<const float4*>(g + i), bb = *reinterpret_cast<const float4*>(b + i);
    y[i] = fmaf(gg.x, (y[i] - mean) * rstd, bb.x);
    y[i + 1] = fmaf(gg.y, (y[i + 1] - mean) * rstd, bb.y);
    y[i + 2] = fmaf(gg.z, (y[i + 2] - mean) * rstd, bb.z);
    y[i + 3] = fmaf(gg.w, (y[i + 3] - mean) * rstd, bb.w);
  }
}

// per-chunk gather descriptor (chunk k = 32 bytes of a token row), staged once per CTA in shared memory so the
// software-pipelined stages read plain words instead of walking the kernel-parameter tables
struct ChunkDesc {
  const int32_t* ids;       // CSR values of the chunk's feature
  const int32_t* offs;      // CSR offsets
  const int32_t* item_ids;  // target item ids
  const float* tab;         // table + first column of the chunk
  int64_t rows;
  int32_t dim;
  int32_t dup;              // same feature as chunk k-1 (same id, same offsets)
};

template <int SLOT>
struct Tc2Layout {
  static constexpr int NS = 128 / SLOT;                 // samples per tile
  static constexpr int W = SLOT < 32 ? SLOT : 32;       // rows of one decoder softmax segment (inside a warp)
  static constexpr int NPARTS = 128 / W;                // partial softmaxes per tile
  static constexpr int PPS = SLOT / W;                  // parts per sample (2 when a slot spans two warps)
  static constexpr int NR = NPARTS * kH;                // rows of the transposed-probability image
  static constexpr int CW = SLOT < 32 ? 32 : SLOT;      // score columns a warp loads (covers its rows' slots)
  // ---- shared memory (bytes) ----
  static constexpr int oWqkv = 0;
  static constexpr int oW1 = oWqkv + 3 * kD * kD * 2;
  static constexpr int oW2 = oW1 + kD * kDFF * 2;
  static constexpr int oGrp = oW2 + kDFF * kD * 2;      // 90112
  static constexpr int gXA = 0, gQ = 16384, gK = 32768, gV = 49152, szGrp = 65536;
  // aliases inside a group's Q region, valid once the S MMAs have completed
  static constexpr int gPd = gQ;                        // [k/8][16 rows][8] bf16: LBO 256, <= 6 KB touched
  static constexpr int gQt = gQ + 8192;                 // fp32 [NS][H*D] folded decoder queries
  static constexpr int gDvec = gQ + 12288;              // fp32 [NS][D] scaled target embeddings
  static constexpr int gMx = gQ + 14336;                // fp32 [NR] partial-softmax maxima | [NR] denominators
                                                        // (rows 0..7 of Q chunk 7: only warp 0 writes there)
  // the memory image (MN-major B operand of the context MMA) takes the K region, its ones chunk the first 2 KB of V
  static constexpr int oFV = oGrp + 2 * szGrp;          // 221184
  static constexpr int vBQKV = 0, vB1 = 3 * kD, vB2 = vB1 + kDFF, vLN = vB2 + kD, nFV = vLN + 4 * kD;
  static constexpr int oPos = oFV + nFV * 4;            // bf16 [maxlen][D] learned positions
  static_assert(NS * kH * kD * 4 <= 4096 && NS * kD * 4 <= 2048 && NR <= 16, "decoder scratch aliases");
  // ---- tensor memory columns (per group, relative to its 256-column half) ----
  static constexpr int tQKV = 0;                        // [0,192)   X Wqkv
  static constexpr int tS = 0;                          // head h: [h*128, h*128+128); P_h in place at [h*128, +64)
  static constexpr int tO = 64;                         // head h: [h*128+64, +32)
  static constexpr int tFF1 = 0;                        // [0,256); H in place at [0,128)
  static constexpr int tFF2 = 128;                      // [128,192)
  static constexpr int tCtx = 192;                      // [192,256): untouched by the next tile's X Wqkv
};

// =====================================================================================================================
// The per-tile program: every token row is split across TWO threads (warps w and w + 4 of a group share
// TMEM lane quarter w): 2 groups x 256 threads = 16 warps per SM at 128 registers.  One thread per row (v2, removed)
// was limited by the dependent-issue rate of 2 warps per scheduler (issue slots 30 % busy, no pipe above 40 %); here
// every thread has half the epilogue work (half hf owns head hf / feature columns [32 hf, 32 hf + 32) / 4 of the 8
// gather chunks) and the schedulers have twice the warps to pick from.  LayerNorm and the decoder scores need the whole row: the two
// halves exchange partial sums through 2 KB of (otherwise unused) shared memory inside the Q region.
// TMEM plan per group (256 columns):  X Wqkv [0,192) -> S_0 [0,128) | S_1 [128,256) -> P_h in place [128h, +64),
// O_h [128h+64, +32) -> A W1 [0,256) -> H packed by half hf IN PLACE inside ITS OWN 128 accumulator columns:
// [128hf, 128hf+64) (so the K = 256 A operand of H W2 is two 64-column pieces) -> H W2 [64,128) -> context [192,256).
// =====================================================================================================================
constexpr int kT3Threads = 512;

// =====================================================================================================================
// One launch for ALL behaviour sequences, with length-bucketed tiles (round 2).
//
// Round 1 launched this kernel once per sequence and padded every sample to the slot size of the batch's LONGEST
// sequence (64 rows for dmt.conf's 50-token histories although the mean length is 34: ~47 % of every MMA / softmax /
// LayerNorm row was padding), paying the pipeline ramp (weight images, first gather, last read-out) and the wave
// quantisation of a persistent grid three times.  Now `seq_bucket_kernel` first orders the samples of every sequence by length class
// (> 32 | 17..32 | <= 16 tokens; `perm`, `counts` in the per-sequence workspace -- no host round trip), and ONE
// persistent kernel walks the concatenated tile list of all (sequence, class) SEGMENTS: 2 / 4 / 8 samples per 128-row
// tile.  Global tile g belongs to tile group (g mod 2*gridDim.x); a group runs the per-tile program of the segment's
// slot size (the three instantiations of the generic lambda below: same memory plan, same tensor-memory plan; KW = the
// key columns of the score window the softmax of a 64-row slot visits, >= the longest sequence), drains
// its software pipeline at a segment boundary and, at a SEQUENCE boundary, the whole CTA swaps the 88 KB of weight
// images / biases / positions / gather descriptors.
// =====================================================================================================================
struct SeqMultiArgs {
  SeqTcArgs a[DMT_MAX_TAIL_SEQS];
  const int32_t* perm[DMT_MAX_TAIL_SEQS];     // [batch] sample indices ordered by length class
  const int32_t* counts[DMT_MAX_TAIL_SEQS];   // [3] samples in the 64- / 32- / 16-row classes
  int32_t n_seq;
};

struct BucketArgs {
  const int32_t* offs[DMT_MAX_TAIL_SEQS];     // CSR offsets [B + 1] that define the sequence lengths
  int32_t* perm[DMT_MAX_TAIL_SEQS];
  int32_t* counts[DMT_MAX_TAIL_SEQS];
  int32_t batch[DMT_MAX_TAIL_SEQS];
  int32_t maxlen[DMT_MAX_TAIL_SEQS];
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ int len_class(int len) { return len > 32 ? 0 : (len > 16 ? 1 : 2); }

// one CTA per sequence: stable counting sort of the samples by length class (thread t owns a contiguous run of samples)
__global__ void __launch_bounds__(1024) seq_bucket_kernel(const __grid_constant__ BucketArgs ba) {
  __shared__ int wsum[32][3];
  __shared__ int tot_s[3];
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  griddep_launch();      // the tile kernel behind this one may set itself up (TMEM, barriers, first weight images)
  const int B = ba.batch[q], maxlen = ba.maxlen[q];
  const int32_t* offs = ba.offs[q];
  const int per = (B + 1023) / 1024;
  const int lo = min(tid * per, B), hi = min(lo + per, B);
  int c[3] = {0, 0, 0};
  for (int b = lo; b < hi; ++b) {
    const int k = len_class(min(__ldg(offs + b + 1) - __ldg(offs + b), maxlen));
    c[0] += k == 0; c[1] += k == 1; c[2] += k == 2;
  }
  int ex[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int v = c[k];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    ex[k] = v - c[k];
    if (lane == 31) wsum[warp][k] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int mine = wsum[lane][k];
      int v = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      wsum[lane][k] = v - mine;
      if (lane == 31) tot_s[k] = v;
    }
  }
  __syncthreads();
  int pos[3];
  pos[0] = wsum[warp][0] + ex[0];
  pos[1] = tot_s[0] + wsum[warp][1] + ex[1];
  pos[2] = tot_s[0] + tot_s[1] + wsum[warp][2] + ex[2];
  int32_t* perm = ba.perm[q];
  for (int b = lo; b < hi; ++b) {
    const int k = len_class(min(__ldg(offs + b + 1) - __ldg(offs + b), maxlen));
    const int p = k == 0 ? pos[0]++ : (k == 1 ? pos[1]++ : pos[2]++);
    perm[p] = b;
  }
  if (tid < 3) ba.counts[q][tid] = tot_s[tid];
}

#define T3_TICK(idx)                                                    \
  if constexpr (DBG) {                                                  \
    if (dbgp && tid == 0) {                                             \
      const long long _now = clock64();                                 \
      atomicAdd(dbgp + (idx), (unsigned long long)(_now - t_last));     \
      t_last = _now;                                                    \
    }                                                                   \
    if (dbgp && gt == 0 && blockIdx.x == 0 && tl_tile < 6 && (idx) < 12) \
      dbg0[1024 + grp * 128 + tl_tile * 16 + (idx)] = (unsigned long long)(clock64() - t_entry); \
  }

template <int N>
using ic = std::integral_constant<int, N>;

// DBG: the diagnostics of dmt_debug_seq_profile (a second instantiation: their counters and clocks cost registers the
// production kernel does not have)
template <bool DBG>
__global__ void __launch_bounds__(kT3Threads, 1) seq_encode_multi_kernel(const __grid_constant__ SeqMultiArgs m) {
  using L0 = Tc2Layout<64>;                           // the byte offsets of the memory plan do not depend on the slot size
  constexpr int D = kD, DFF = kDFF, H = kH, DK = kDK, KC = kKC, ROWB = kROWB;
  constexpr int HC = KC / 2;                          // gather chunks per half
  constexpr int tFF2 = 64;                            // H W2 accumulator (v3 plan)
  constexpr int oExLN = 6144, oExSc = 4096;           // exchange scratch inside the Q region (see Tc2Layout aliases)

  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2], cbars[2], wbar;        // per group: phase MMAs | context MMA; weights landed
  __shared__ uint32_t tmem_base_s;
  __shared__ int slen_s[2][2][8];
  __shared__ ChunkDesc sd[KC];
  // partial-softmax maxima | denominators of the group's last tile.  NOT inside the Q image like the other decoder
  // scratch: the read-out warps read them in the shadow of the next tile's X Wqkv, and the first of them to finish goes
  // on to rewrite the Q image
  __shared__ float mxs_s[2][32];

  const int tid = threadIdx.x, grp = tid >> 8, gt = tid & 255, row = gt & 127, hf = gt >> 7;
  const int wq = (gt >> 5) & 3, lane = tid & 31;      // wq: warp inside the half = TMEM lane quarter
  uint8_t* gbase = smem + L0::oGrp + grp * L0::szGrp;
  uint8_t* sXA = gbase + L0::gXA;
  uint8_t* sQ = gbase + L0::gQ;
  uint8_t* sK = gbase + L0::gK;
  float* fv = reinterpret_cast<float*>(smem + L0::oFV);
  const uint4* spos = reinterpret_cast<const uint4*>(smem + L0::oPos);
  uint64_t* bar = &bars[grp];
  uint64_t* cbar = &cbars[grp];
  const uint32_t bar_id = 1 + grp;

  // diagnostics (dmt_debug_seq_profile): [q*16 + phase] cycles of CTA 0 / group 0 per sequence, [64 + cta] cycles of
  // the whole CTA, [320 + cta] / [576 + cta] %globaltimer at entry / exit
  unsigned long long* const dbg0 = DBG ? m.a[0].dbg : nullptr;
  const long long t_entry = DBG ? clock64() : 0;
  if (DBG && dbg0 && tid == 0) dbg0[320 + blockIdx.x] = globaltimer_ns();

  if (tid < 32) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&cbars[0], 1);
    mbar_init(&cbars[1], 1);
    mbar_init(&wbar, 1);
    mbar_fence_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_base_s + grp * 256;
  const uint32_t aXA = smem_u32(sXA), aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(gbase + L0::gV);
  const uint32_t dHi = desc_hi(128, kLayoutNone), dHiV = desc_hi(ROWB, kLayoutNone);
  const uint32_t dXA = desc_lo(aXA, ROWB), dQ = desc_lo(aQ, ROWB), dK = desc_lo(aK, ROWB);
  const uint32_t dWqkv = desc_lo(smem_u32(smem + L0::oWqkv), 3 * D * 16), dW1 = desc_lo(smem_u32(smem + L0::oW1), DFF * 16),
                 dW2 = desc_lo(smem_u32(smem + L0::oW2), D * 16);
  const float sqrt_d = sqrtf((float)D);
  const float sl2 = (1.0f / sqrtf((float)DK)) * 1.4426950408889634f;
  const int c0h = hf * HC;                             // first gather chunk / 8-column group of this half
  uint32_t phase = 0, cphase = 0;
  // exchange slots: [half][row] pairs of floats (consecutive lanes = consecutive 8-byte slots)
  float2* exLN = reinterpret_cast<float2*>(sQ + oExLN);
  float2* exSc = reinterpret_cast<float2*>(sQ + oExSc);
  const int G = blockIdx.x * 2 + grp, S = 2 * gridDim.x;   // this tile group | tile groups of the grid
  int tl_tile = 0;                                     // diagnostics: tiles this group has finished (timeline capture)
  int seg_g0 = 0;                                      // global index of the current segment's first tile

  for (int q = 0; q < m.n_seq; ++q) {
    const SeqTcArgs& a = m.a[q];
    const long long t_seq = DBG ? clock64() : 0;
    // ---- sequence prologue: gather descriptors, weight images, biases / LayerNorm vectors, positions.  Every group
    //      has waited for its last context MMA (ctx_readout), i.e. for every MMA it issued: nothing reads the old
    //      images any more once all threads are here ----
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid < KC) {
      const int f = a.chunk_feat[tid];
      sd[tid].ids = a.in.ids[f];
      sd[tid].offs = a.in.offsets[f];
      sd[tid].item_ids = a.in.item_ids[f];
      sd[tid].tab = a.in.table[f] + a.chunk_off[tid];
      sd[tid].rows = a.in.rows[f];
      sd[tid].dim = a.in.dim[f];
      sd[tid].dup = (tid > 0 && (tid % HC) != 0 && f == a.chunk_feat[tid - 1]) ? 1 : 0;   // dup only inside a half
    }
    if (tid == 0) {
      mbar_expect_tx(&wbar, L0::oGrp);
      bulk_g2s(smem + L0::oWqkv, a.prepared, L0::oGrp, &wbar);
    }
    {
      for (int i = tid; i < D; i += kT3Threads) {
        fv[L0::vBQKV + i] = a.bq[i];
        fv[L0::vBQKV + D + i] = a.bk[i];
        fv[L0::vBQKV + 2 * D + i] = a.bv[i];
        fv[L0::vB2 + i] = a.b2[i];
        fv[L0::vLN + 0 * D + i] = a.ln1_g[i];
        fv[L0::vLN + 1 * D + i] = a.ln1_b[i];
        fv[L0::vLN + 2 * D + i] = a.ln2_g[i];
        fv[L0::vLN + 3 * D + i] = a.ln2_b[i];
      }
      for (int i = tid; i < DFF; i += kT3Threads) fv[L0::vB1 + i] = a.b1[i];
      uint4* pdst = reinterpret_cast<uint4*>(smem + L0::oPos);
      for (int i = tid; i < a.cfg.maxlen * KC; i += kT3Threads) {
        const float4 p0 = ldg4(a.pos + i * 8), p1 = ldg4(a.pos + i * 8 + 4);
        const float f[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        // 16-byte chunk k of position t lives at chunk k ^ (t & 7) of its 128-byte row: the threads of a warp read
        // consecutive positions, the same chunk -- unswizzled that is eight lanes on the same four banks
        pdst[(i & ~7) | ((i & 7) ^ ((i >> 3) & 7))] = f8_to_bf16(f);
      }
    }
    __syncthreads();
    const __nv_bfloat16* gDec = a.prepared + prep_off_dec(D, DFF);
    const uint4* gG = reinterpret_cast<const uint4*>(gDec);
    const float* gGb = reinterpret_cast<const float*>(gDec + (size_t)H * D * D + (size_t)D * D);
    const int32_t* const len_offs = a.in.offsets[a.cfg.n_feats - 1];
    const int maxlen = a.cfg.maxlen;
    const int zp = a.cfg.zero_pad ? 1 : 0;
    const int32_t* const perm = m.perm[q];
    void* const ctxp = a.ctx;
    unsigned long long* const dbgp = (DBG && a.dbg) ? a.dbg + q * 16 : nullptr;
    if (DBG && dbgp && tid == 0) atomicAdd(dbgp + 15, (unsigned long long)(clock64() - t_seq));   // sequence prologue
    bool wpending = gt == 0;                           // the MMA issuer waits for the images before its first MMA
    const uint32_t wpar = q & 1;

    auto run_segment = [&](auto slot_c, auto kw_c, const int seg_cnt, const int seg_base) {
      constexpr int SLOT = decltype(slot_c)::value, KW = decltype(kw_c)::value;
      using L = Tc2Layout<SLOT>;
      static_assert(KW % 8 == 0 && KW <= L::CW && (SLOT == 64 || KW == L::CW), "key window");
      constexpr int NS = L::NS, CW = L::CW, W = L::W, NR = L::NR, PPS = L::PPS;
      const int n_tiles = (seg_cnt + NS - 1) / NS;
      int first = (G - seg_g0) % S;                    // this group's first tile of the segment (global striding)
      if (first < 0) first += S;
      seg_g0 += n_tiles;
      if (first >= n_tiles) return;
      const int lmax = maxlen < SLOT ? maxlen : SLOT;
      const int slot = row / SLOT, tpos = row % SLOT;
      long long t_last = DBG ? clock64() : 0;
      // ---- software-pipelined gather: this thread loads chunks c0h .. c0h+HC-1 of its token row ----
      int pf_o0[HC], pf_o1[HC], pf_id[HC];
      int pf_l0 = 0, pf_l1 = 0, pf_len = 0;
      bool pf_valid = false;
      f8 pf_e[HC];
      int pf_tid = kInvalidId;
      float4 pf_t0 = make_float4(0.f, 0.f, 0.f, 0.f), pf_t1 = pf_t0;

      auto stage_offsets = [&](int nt) {
        pf_l0 = pf_l1 = 0;
        pf_tid = kInvalidId;
#pragma unroll
        for (int k = 0; k < HC; ++k) pf_o0[k] = pf_o1[k] = 0;
        if (nt >= n_tiles) return;
        const int si = nt * NS + slot;
        if (si < seg_cnt) {
          const int b = __ldcg(perm + seg_base + si);
          pf_l0 = __ldg(len_offs + b);
          pf_l1 = __ldg(len_offs + b + 1);
#pragma unroll
          for (int k = 0; k < HC; ++k) {
            if (k > 0 && sd[c0h + k].dup) {
              pf_o0[k] = pf_o0[k - 1];
              pf_o1[k] = pf_o1[k - 1];
            } else {
              const int32_t* of = sd[c0h + k].offs;
              pf_o0[k] = __ldg(of + b);
              pf_o1[k] = __ldg(of + b + 1);
            }
          }
        }
        if (gt < NS * KC) {
          const int st = nt * NS + gt / KC;
          if (st < seg_cnt) pf_tid = __ldg(sd[gt % KC].item_ids + __ldcg(perm + seg_base + st));
        }
      };
      auto stage_ids = [&](int nt, int par) {
        pf_len = min(pf_l1 - pf_l0, lmax);
        pf_valid = tpos < pf_len;
        if (tpos == 0 && hf == 0) slen_s[grp][par][slot] = pf_len;
#pragma unroll
        for (int k = 0; k < HC; ++k) {
          pf_id[k] = kInvalidId;
          if (pf_valid) {
            if (k > 0 && sd[c0h + k].dup) pf_id[k] = pf_id[k - 1];
            else pf_id[k] = (tpos < pf_o1[k] - pf_o0[k]) ? __ldg(sd[c0h + k].ids + pf_o0[k] + tpos) : 0;
          }
        }
      };
      auto stage_rows = [&]() {
#pragma unroll
        for (int k = 0; k < HC; ++k) {
          pf_e[k].lo = make_float4(0.f, 0.f, 0.f, 0.f);
          pf_e[k].hi = pf_e[k].lo;
          const int64_t rw = (int64_t)pf_id[k] - zp;
          if (pf_id[k] != kInvalidId && rw >= 0 && rw < sd[c0h + k].rows)
            pf_e[k] = ld_stream8(sd[c0h + k].tab + rw * sd[c0h + k].dim);
        }
        pf_t0 = make_float4(0.f, 0.f, 0.f, 0.f);
        pf_t1 = pf_t0;
        if (gt < NS * KC) {
          const int c = gt % KC;
          const int64_t rw = (int64_t)pf_tid - zp;
          if (pf_tid != kInvalidId && rw >= 0 && rw < sd[c].rows) {
            const f8 t = ld_stream8(sd[c].tab + rw * sd[c].dim);
            pf_t0 = t.lo;
            pf_t1 = t.hi;
          }
        }
      };
      // decoder contexts of the tile whose first sample is rb0.  The context MMA is issued TRANSPOSED (A = the memory
      // image read MN-major, B = the 16-row probability image): accumulator row k = feature k, column (part, head), so
      // the read-out is 16 values in each of 64 lanes (two warps) instead of 64 values in each of 8-16 lanes of one
      // warp, and the two partial softmaxes of a 64-row slot sit in the same lane (no shuffles).
      auto ctx_readout = [&](int rb0) {
        if (gt >= D) return;
        mbar_wait(cbar, cphase);
        cphase ^= 1;
        fence_after_sync();
        uint32_t c[16];
        tmem_ld16(tmem_addr(tbase, L::tCtx), c);
        tmem_ld_wait();
        const float* mxs = mxs_s[grp];
        const int kf = gt;                                 // feature column = TMEM lane
        uint8_t* dst0 = reinterpret_cast<uint8_t*>(ctxp) + (size_t)(kf >> 3) * (128 * 16) + (kf & 7) * 2;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const int si = rb0 + s;
          const int b = si < seg_cnt ? __ldcg(perm + seg_base + si) : -1;
#pragma unroll
          for (int h = 0; h < H; ++h) {
            float num, den;
            if constexpr (PPS == 2) {
              const int ia = (2 * s) * H + h, ib = (2 * s + 1) * H + h;
              const float ma = mxs[ia], mb = mxs[ib];
              const float m = fmaxf(ma, mb);
              const float wa = (ma == -INFINITY) ? 0.f : ex2_approx(ma - m), wb = (mb == -INFINITY) ? 0.f : ex2_approx(mb - m);
              num = __uint_as_float(c[ia]) * wa + __uint_as_float(c[ib]) * wb;
              den = mxs[NR + ia] * wa + mxs[NR + ib] * wb;
            } else {
              num = __uint_as_float(c[s * H + h]);
              den = mxs[NR + s * H + h];
            }
            const float v = den > 0.f ? num / den : 0.f;    // empty sequence: context 0
            if (b >= 0)
              *reinterpret_cast<unsigned short*>(dst0 + (size_t)(b >> 7) * (128 * H * D * 2) + (size_t)(h * KC) * (128 * 16) +
                                                 (size_t)(b & 127) * 16) = __bfloat16_as_ushort(__float2bfloat16(v));
          }
        }
        fence_before_sync();
      };
      // concat + sqrt(d) scale + learned position -> bf16, in registers: done early (in an MMA shadow) so that P0 is
      // four stores
      uint4 px[HC];
      auto convert_rows = [&]() {
#pragma unroll
        for (int k = 0; k < HC; ++k) {
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = 0.f;
          if (pf_valid) {
            float p[8];
            bf16x8_to_f(spos[tpos * KC + ((c0h + k) ^ (tpos & 7))], p);
            x[0] = fmaf(pf_e[k].lo.x, sqrt_d, p[0]); x[1] = fmaf(pf_e[k].lo.y, sqrt_d, p[1]);
            x[2] = fmaf(pf_e[k].lo.z, sqrt_d, p[2]); x[3] = fmaf(pf_e[k].lo.w, sqrt_d, p[3]);
            x[4] = fmaf(pf_e[k].hi.x, sqrt_d, p[4]); x[5] = fmaf(pf_e[k].hi.y, sqrt_d, p[5]);
            x[6] = fmaf(pf_e[k].hi.z, sqrt_d, p[6]); x[7] = fmaf(pf_e[k].hi.w, sqrt_d, p[7]);
          }
          px[k] = f8_to_bf16(x);
        }
      };
      int n_done = 0;
      const int tile0 = first, tstride = S;
      stage_offsets(tile0);
      stage_ids(tile0, 0);
      stage_rows();
      convert_rows();
      T3_TICK(12);                                       // pipeline prime (offsets -> ids -> rows of the first tile)

      for (int it = 0;; ++it) {
        const int tile = tile0 + it * tstride;
        if (tile >= n_tiles) break;
        const int par = it & 1;
        const int b0 = tile * NS;                          // segment-local index of the tile's first sample
        const int next_tile = tile + tstride;

        // ---- P0: this half's four (already converted) chunks -> X image ----
#pragma unroll
        for (int k = 0; k < HC; ++k) *reinterpret_cast<uint4*>(sXA + (c0h + k) * ROWB + row * 16) = px[k];
        const float4 cur_t0 = pf_t0, cur_t1 = pf_t1;
        fence_proxy_async();
        fence_before_sync();
        named_sync(bar_id, 256);
        T3_TICK(0);

        // ---- P1: [Q|K|V] = X Wqkv ----
        if (gt == 0) {
          if (wpending) {                                // this sequence's weight images landed (bulk copy issued in
            mbar_wait(&wbar, wpar);                      // its prologue)
            wpending = false;
          }
          fence_after_sync();
          constexpr uint32_t idesc = make_idesc_bf16(128, 3 * D);
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks)
            mma_bf16_ss(tbase + L::tQKV, desc_join(dXA + ks * (2 * ROWB / 16), dHi),
                        desc_join(dWqkv + ks * (2 * 3 * D), dHi), idesc, ks > 0);
          commit(bar);
        }
        if (it > 0) ctx_readout(b0 - tstride * NS);
        stage_offsets(next_tile);
        mbar_wait(bar, phase);
        phase ^= 1;
        fence_after_sync();
        T3_TICK(1);

        // ---- P2: + bias, bf16 -> Q / K / V images; half hf converts columns [96 hf, 96 hf + 96) ----
#pragma unroll 1
        for (int blk = 0; blk < 3; ++blk) {
          const int n0 = hf * 96 + blk * 32;
          uint32_t r[32];
          tmem_ld32(tmem_addr(tbase, L::tQKV + n0), r);
          tmem_ld_wait();
          const int m = n0 >> 6;                           // 0: Q, 1: K, 2: V image
          uint8_t* dstm = sQ + m * 16384 + row * 16;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int n = n0 + g * 8;
            const float4 ba = *reinterpret_cast<const float4*>(fv + L::vBQKV + n);
            const float4 bb = *reinterpret_cast<const float4*>(fv + L::vBQKV + n + 4);
            float y[8];
            y[0] = __uint_as_float(r[g * 8 + 0]) + ba.x; y[1] = __uint_as_float(r[g * 8 + 1]) + ba.y;
            y[2] = __uint_as_float(r[g * 8 + 2]) + ba.z; y[3] = __uint_as_float(r[g * 8 + 3]) + ba.w;
            y[4] = __uint_as_float(r[g * 8 + 4]) + bb.x; y[5] = __uint_as_float(r[g * 8 + 5]) + bb.y;
            y[6] = __uint_as_float(r[g * 8 + 6]) + bb.z; y[7] = __uint_as_float(r[g * 8 + 7]) + bb.w;
            const int ch = ((n0 & 63) >> 3) + g;
            *reinterpret_cast<uint4*>(dstm + ch * ROWB) = f8_to_bf16(y);
          }
        }
        fence_proxy_async();
        fence_before_sync();
        named_sync(bar_id, 256);
        T3_TICK(2);

        // ---- P3: S_h = Q_h K_h^T ----
        if (gt == 0) {
          fence_after_sync();
          constexpr uint32_t idesc = make_idesc_bf16(128, 128);
#pragma unroll
          for (int h = 0; h < H; ++h)
#pragma unroll
            for (int ks = 0; ks < DK / 16; ++ks) {
              const uint32_t ch = (h * DK) / 8 + ks * 2;
              mma_bf16_ss(tbase + L::tS + h * 128, desc_join(dQ + ch * (ROWB / 16), dHi),
                          desc_join(dK + ch * (ROWB / 16), dHi), idesc, ks > 0);
            }
          commit(bar);
        }
        stage_ids(next_tile, par ^ 1);
        mbar_wait(bar, phase);
        phase ^= 1;
        fence_after_sync();
        T3_TICK(3);

        // ---- P4: masked softmax of head hf; unnormalised P_hf packed IN PLACE ----
        const int len = slen_s[grp][par][slot];
        float inv_h = 0.f;
        {
          const int col0 = (row / CW) * CW;
          const int lo = (SLOT == CW) ? 0 : slot * SLOT - col0;
          const uint32_t sbase = tmem_addr(tbase, L::tS + hf * 128);
          uint32_t r[CW];
          tmem_ld32(sbase + col0, r);
          if constexpr (KW >= 48) tmem_ld16(sbase + col0 + 32, r + 32);
          if constexpr (KW == 56) tmem_ld8(sbase + col0 + 48, r + 48);
          if constexpr (KW == 64) tmem_ld16(sbase + col0 + 48, r + 48);
          tmem_ld_wait();
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < KW; ++j) {
            const bool ok = (unsigned)(j - lo) < (unsigned)len;
            const float v = ok ? __uint_as_float(r[j]) : -INFINITY;
            r[j] = __float_as_uint(v);
            mx = fmaxf(mx, v);
          }
          const float mxs = (mx == -INFINITY) ? 0.f : mx * sl2;
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int j = 0; j < KW; j += 4) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(r[j]), sl2, -mxs));
            const float e1 = ex2_approx(fmaf(__uint_as_float(r[j + 1]), sl2, -mxs));
            const float e2 = ex2_approx(fmaf(__uint_as_float(r[j + 2]), sl2, -mxs));
            const float e3 = ex2_approx(fmaf(__uint_as_float(r[j + 3]), sl2, -mxs));
            s0 += e0; s1 += e1; s2 += e2; s3 += e3;
            r[j / 2] = pack_bf16x2(e0, e1);                 // (j/2 <= j: the packed words trail the reads)
            r[j / 2 + 1] = pack_bf16x2(e2, e3);
          }
          const float sum = (s0 + s1) + (s2 + s3);
          inv_h = sum > 0.f ? 1.0f / sum : 0.f;
#pragma unroll
          for (int j = KW / 2; j < CW; ++j) r[j] = 0u;
          if constexpr (CW == 64) {
            tmem_st32(sbase + col0 / 2, r);
            tmem_st32(sbase + (32 - col0 / 2), r + 32);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q * 16 == col0 / 2) tmem_st16(sbase + q * 16, r);
              else tmem_st16(sbase + q * 16, r + 16);
            }
          }
        }
        if (gt < NS * KC) {
          float* dv = reinterpret_cast<float*>(gbase + L::gDvec) + (gt / KC) * D + (gt % KC) * 8;
          *reinterpret_cast<float4*>(dv) = make_float4(cur_t0.x * sqrt_d, cur_t0.y * sqrt_d, cur_t0.z * sqrt_d, cur_t0.w * sqrt_d);
          *reinterpret_cast<float4*>(dv + 4) = make_float4(cur_t1.x * sqrt_d, cur_t1.y * sqrt_d, cur_t1.z * sqrt_d, cur_t1.w * sqrt_d);
        }
        tmem_st_wait();
        fence_before_sync();
        named_sync(bar_id, 256);
        T3_TICK(4);

        // ---- P5: O_h = P_h V_h ----
        if (gt == 0) {
          fence_after_sync();
          constexpr uint32_t idesc = make_idesc_bf16(128, DK, false, true);
#pragma unroll
          for (int h = 0; h < H; ++h) {
            const uint32_t dV = desc_lo(aV + ((h * DK) / 8) * ROWB, 128);
#pragma unroll
            for (int ks = 0; ks < 128 / 16; ++ks)
              mma_bf16_ts(tbase + L::tO + h * 128, tbase + L::tS + h * 128 + ks * 8,
                          desc_join(dV + ks * (256 / 16), dHiV), idesc, ks > 0);
          }
          commit(bar);
        }
        // folded decoder queries qt[s][n] = dvec[s] . G[n] + g[n]: thread (n = row, hf) computes the samples s = hf,
        // hf + 2, ...; the contraction is split between the shadows of the P.V and the H.W2 MMAs
        constexpr int NSH = NS / 2;
        float qacc[NSH];
        auto qt_part = [&](int jc0) {
          const float* dvs = reinterpret_cast<const float*>(gbase + L::gDvec);
#pragma unroll
          for (int jc = jc0; jc < jc0 + KC / 2; ++jc) {
            float w[8];
            bf16x8_to_f(__ldg(gG + jc * (H * D) + row), w);
#pragma unroll
            for (int s = 0; s < NSH; ++s) {
              const float4 d0 = *reinterpret_cast<const float4*>(dvs + (2 * s + hf) * D + jc * 8);
              const float4 d1 = *reinterpret_cast<const float4*>(dvs + (2 * s + hf) * D + jc * 8 + 4);
              qacc[s] = fmaf(d0.x, w[0], fmaf(d0.y, w[1], fmaf(d0.z, w[2], fmaf(d0.w, w[3], qacc[s]))));
              qacc[s] = fmaf(d1.x, w[4], fmaf(d1.y, w[5], fmaf(d1.z, w[6], fmaf(d1.w, w[7], qacc[s]))));
            }
          }
        };
        {
          const float gb = __ldg(gGb + row);
#pragma unroll
          for (int s = 0; s < NSH; ++s) qacc[s] = gb;
          qt_part(0);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        fence_after_sync();
        T3_TICK(5);

        // ---- P6: A = LN(O + X): half hf owns columns [32 hf, 32 hf + 32) = head hf ----
        {
          float y[DK];
          uint32_t r[32];
          tmem_ld32(tmem_addr(tbase, L::tO + hf * 128), r);
#pragma unroll
          for (int c = 0; c < HC; ++c) bf16x8_to_f(*reinterpret_cast<const uint4*>(sXA + (c0h + c) * ROWB + row * 16), y + c * 8);
          tmem_ld_wait();
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
          for (int e = 0; e < DK; e += 2) {
            y[e] = fmaf(__uint_as_float(r[e]), inv_h, y[e]);
            y[e + 1] = fmaf(__uint_as_float(r[e + 1]), inv_h, y[e + 1]);
            s0 += y[e]; s1 += y[e + 1];
            q0 = fmaf(y[e], y[e], q0); q1 = fmaf(y[e + 1], y[e + 1], q1);
          }
          exLN[hf * 128 + row] = make_float2(s0 + s1, q0 + q1);
          named_sync(bar_id, 256);
          const float2 o = exLN[(hf ^ 1) * 128 + row];
          const float mean = ((s0 + s1) + o.x) * (1.0f / D);
          const float var = fmaxf(((q0 + q1) + o.y) * (1.0f / D) - mean * mean, 0.f);
          const float rstd = 1.0f / sqrtf(var + kLnEps);
          const float* g = fv + L::vLN + 0 * D + hf * DK;
          const float* bt = fv + L::vLN + 1 * D + hf * DK;
#pragma unroll
          for (int e = 0; e < DK; e += 4) {
            const float4 gg = *reinterpret_cast<const float4*>(g + e), bb = *reinterpret_cast<const float4*>(bt + e);
            y[e] = fmaf(gg.x, (y[e] - mean) * rstd, bb.x);
            y[e + 1] = fmaf(gg.y, (y[e + 1] - mean) * rstd, bb.y);
            y[e + 2] = fmaf(gg.z, (y[e + 2] - mean) * rstd, bb.z);
            y[e + 3] = fmaf(gg.w, (y[e + 3] - mean) * rstd, bb.w);
          }
#pragma unroll
          for (int c = 0; c < HC; ++c) *reinterpret_cast<uint4*>(sXA + (c0h + c) * ROWB + row * 16) = f8_to_bf16(y + c * 8);
        }
        fence_proxy_async();
        fence_before_sync();
        named_sync(bar_id, 256);
        T3_TICK(6);

        // ---- P7: hidden = A W1 ----
        if (gt == 0) {
          fence_after_sync();
          constexpr uint32_t idesc = make_idesc_bf16(128, DFF);
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks)
            mma_bf16_ss(tbase + L::tFF1, desc_join(dXA + ks * (2 * ROWB / 16), dHi),
                        desc_join(dW1 + ks * (2 * DFF), dHi), idesc, ks > 0);
          commit(bar);
        }
        stage_rows();
        mbar_wait(bar, phase);
        phase ^= 1;
        fence_after_sync();
        T3_TICK(7);

        // ---- P8: relu(+b1): half hf packs accumulator columns [128 hf, 128 hf + 128) into [128 hf, 128 hf + 64) ----
#pragma unroll 1
        for (int blk = 0; blk < 4; ++blk) {
          uint32_t r[32];
          tmem_ld32(tmem_addr(tbase, L::tFF1 + hf * 128 + blk * 32), r);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 bb = *reinterpret_cast<const float4*>(fv + L::vB1 + hf * 128 + blk * 32 + g * 4);
            pk[g * 2] = pack_bf16x2(fmaxf(__uint_as_float(r[g * 4]) + bb.x, 0.f), fmaxf(__uint_as_float(r[g * 4 + 1]) + bb.y, 0.f));
            pk[g * 2 + 1] = pack_bf16x2(fmaxf(__uint_as_float(r[g * 4 + 2]) + bb.z, 0.f), fmaxf(__uint_as_float(r[g * 4 + 3]) + bb.w, 0.f));
          }
          tmem_st16(tmem_addr(tbase, L::tFF1 + hf * 128 + blk * 16), pk);
        }
        tmem_st_wait();
        fence_before_sync();
        named_sync(bar_id, 256);
        T3_TICK(8);

        // ---- P9: F = H W2; the K = 256 A operand is two 64-column pieces of tensor memory ----
        if (gt == 0) {
          fence_after_sync();
          constexpr uint32_t idesc = make_idesc_bf16(128, D);
#pragma unroll
          for (int ks = 0; ks < DFF / 16; ++ks)
            mma_bf16_ts(tbase + tFF2, tbase + L::tFF1 + (ks / 8) * 128 + (ks % 8) * 8, desc_join(dW2 + ks * (2 * D), dHi),
                        idesc, ks > 0);
          commit(bar);
        }
        convert_rows();                                      // next tile's rows (requested in P7) -> bf16 registers
        {
          qt_part(KC / 2);
          float* qt = reinterpret_cast<float*>(gbase + L::gQt);
#pragma unroll
          for (int s = 0; s < NSH; ++s) qt[(2 * s + hf) * (H * D) + row] = qacc[s];
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        fence_after_sync();
        T3_TICK(9);                                          // (qt is read after P10's LayerNorm-exchange barrier)

        // ---- P10: memory = LN(F + b2 + A); decoder scores; partial softmax of head hf; images for the context MMA ----
        {
          float y[DK];
          uint32_t r[32];
          tmem_ld32(tmem_addr(tbase, tFF2 + hf * DK), r);
#pragma unroll
          for (int c = 0; c < HC; ++c) bf16x8_to_f(*reinterpret_cast<const uint4*>(sXA + (c0h + c) * ROWB + row * 16), y + c * 8);
          tmem_ld_wait();
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
          for (int e = 0; e < DK; e += 2) {
            y[e] += __uint_as_float(r[e]) + fv[L::vB2 + hf * DK + e];
            y[e + 1] += __uint_as_float(r[e + 1]) + fv[L::vB2 + hf * DK + e + 1];
            s0 += y[e]; s1 += y[e + 1];
            q0 = fmaf(y[e], y[e], q0); q1 = fmaf(y[e + 1], y[e + 1], q1);
          }
          exLN[hf * 128 + row] = make_float2(s0 + s1, q0 + q1);
          named_sync(bar_id, 256);
          {
            const float2 o = exLN[(hf ^ 1) * 128 + row];
            const float mean = ((s0 + s1) + o.x) * (1.0f / D);
            const float var = fmaxf(((q0 + q1) + o.y) * (1.0f / D) - mean * mean, 0.f);
            const float rstd = 1.0f / sqrtf(var + kLnEps);
            const float* g = fv + L::vLN + 2 * D + hf * DK;
            const float* bt = fv + L::vLN + 3 * D + hf * DK;
#pragma unroll
            for (int e = 0; e < DK; e += 4) {
              const float4 gg = *reinterpret_cast<const float4*>(g + e), bb = *reinterpret_cast<const float4*>(bt + e);
              y[e] = fmaf(gg.x, (y[e] - mean) * rstd, bb.x);
              y[e + 1] = fmaf(gg.y, (y[e + 1] - mean) * rstd, bb.y);
              y[e + 2] = fmaf(gg.z, (y[e + 2] - mean) * rstd, bb.z);
              y[e + 3] = fmaf(gg.w, (y[e + 3] - mean) * rstd, bb.w);
            }
          }
          // memory image: this half's four chunks
#pragma unroll
          for (int c = 0; c < HC; ++c) *reinterpret_cast<uint4*>(sK + (c0h + c) * ROWB + row * 16) = f8_to_bf16(y + c * 8);
          // partial decoder scores over this half's 32 columns, both heads; exchanged with the other half
          const float* qt = reinterpret_cast<const float*>(gbase + L::gQt) + slot * (H * D) + hf * DK;
          float pu[H];
#pragma unroll
          for (int h = 0; h < H; ++h) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int k = 0; k < DK; k += 4) {
              const float4 q = *reinterpret_cast<const float4*>(qt + h * D + k);
              a0 = fmaf(y[k], q.x, a0); a1 = fmaf(y[k + 1], q.y, a1);
              a0 = fmaf(y[k + 2], q.z, a0); a1 = fmaf(y[k + 3], q.w, a1);
            }
            pu[h] = a0 + a1;
          }
          exSc[hf * 128 + row] = make_float2(pu[0], pu[1]);
          named_sync(bar_id, 256);
          const float2 os = exSc[(hf ^ 1) * 128 + row];
          // this half finishes head hf
          const float dot = (hf == 0 ? pu[0] + os.x : pu[1] + os.y);
          const float u = (tpos < len) ? dot * sl2 : -INFINITY;
          const int part = (SLOT >= 32) ? wq : (wq * (32 / W) + lane / W);
          float m = u;
#pragma unroll
          for (int o = W / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          float e = (tpos < len) ? ex2_approx(u - m) : 0.f;
          e = __bfloat162float(__float2bfloat16(e));
          float dsum = e;
#pragma unroll
          for (int o = W / 2; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
          if ((lane % W) == 0) {
            mxs_s[grp][part * H + hf] = m;
            mxs_s[grp][NR + part * H + hf] = dsum;
          }
          // transposed probabilities: rows (p, hf) for every part p, column = this token
          const unsigned short eb = __bfloat16_as_ushort(__float2bfloat16(e));
          uint8_t* pd = gbase + L::gPd + (row >> 3) * 256 + (row & 7) * 2 + hf * 16;
#pragma unroll
          for (int p = 0; p < NR / H; ++p)
            *reinterpret_cast<unsigned short*>(pd + p * (H * 16)) = (p == part) ? eb : (unsigned short)0;
        }
        fence_proxy_async();
        fence_before_sync();
        named_sync(bar_id, 256);
        T3_TICK(10);

        // ---- P11: context MMA, transposed: D[feature][(part, head)] = sum_t M_t[feature] e_t  (A = memory image read
        //      MN-major, B = the compact probability image); read out in the shadow of the next tile's X Wqkv ----
        if (gt == 0) {
          fence_after_sync();
          constexpr uint32_t idesc = make_idesc_bf16(128, 16, true, false);
          const uint32_t dPd = desc_lo(smem_u32(gbase + L::gPd), 256), dM = desc_lo(aK, 128);
#pragma unroll
          for (int ks = 0; ks < 128 / 16; ++ks)
            mma_bf16_ss(tbase + L::tCtx, desc_join(dM + ks * (256 / 16), dHiV), desc_join(dPd + ks * (2 * 256 / 16), dHi),
                        idesc, ks > 0);
          commit(cbar);
        }
        n_done = it + 1;
        T3_TICK(11);
        if (DBG) ++tl_tile;
      }


      if (n_done > 0) ctx_readout((tile0 + (n_done - 1) * tstride) * NS);
      T3_TICK(13);                                       // pipeline drain (last context read-out)
      if (DBG && dbgp && tid == 0) atomicAdd(dbgp + 14, (unsigned long long)n_done);
    };

    // Dependent launch: perm / counts are seq_bucket_kernel's output, written while this grid may already be running --
    // they are read with coherent loads (ld.global.cg), never through the read-only path, whose loads the compiler may
    // hoist above the wait and the hardware may serve from a stale line.
    if (q == 0) griddep_wait();
    const int c64 = __ldcg(m.counts[q]), c32 = __ldcg(m.counts[q] + 1), c16 = __ldcg(m.counts[q] + 2);
    run_segment(ic<64>{}, ic<56>{}, c64, 0);
    run_segment(ic<32>{}, ic<32>{}, c32, c64);
    run_segment(ic<16>{}, ic<32>{}, c16, c64 + c32);
  }

  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base_s, 512);
  if (DBG && dbg0 && tid == 0) {
    dbg0[64 + blockIdx.x] = (unsigned long long)(clock64() - t_entry);
    dbg0[576 + blockIdx.x] = globaltimer_ns();
  }
}

// ---- per-sample decoder tail, row-batched: 128 samples per CTA (one thread = one sample = one TMEM lane) ----
//   o  = [ctx_0 | ctx_1] WvBD + bv (bv only for non-empty sequences) ; y = o + dvec ; av = LN3(y)
//   u  = LN2(relu(av W1 + b1) W2 + b2 + av)                                   (TransformerModel.py:157-171)
struct TailLayout {
  static constexpr int oA = 0;                          // ctx image(128 samples, H*D)  32 KB  (written by the main kernel)
  static constexpr int oWv = oA + 128 * kH * kD * 2;    // image(D, H*D)   16 KB
  static constexpr int oW1 = oWv + kH * kD * kD * 2;    // image(DFF, D)   32 KB
  static constexpr int oW2 = oW1 + kD * kDFF * 2;       // image(D, DFF)   32 KB
  static constexpr int oFV = oW2 + kDFF * kD * 2;
  static constexpr int vBV = 0, vB1 = kD, vB2 = vB1 + kDFF, vLN2 = vB2 + kD, vLN3 = vLN2 + 2 * kD, nFV = vLN3 + 2 * kD;
  static constexpr int oOut = oFV + nFV * 4;            // fp32 [128][D+1] output transpose
  static constexpr int total = oOut + 128 * (kD + 1) * 4 + 64;
  static constexpr int tA = 0, tO = 64, tFF1 = 128, tFF2 = 384;
};

struct TailBatch {                                    // the tails of up to DMT_MAX_TAIL_SEQS sequences in one launch
  SeqTcArgs a[DMT_MAX_TAIL_SEQS];
  int32_t first_tile[DMT_MAX_TAIL_SEQS + 1];          // CTA range of each sequence
  int32_t n_seq;
};

__global__ void __launch_bounds__(128, 1) seq_tail_kernel(const __grid_constant__ TailBatch tb) {
  using L = TailLayout;
  constexpr int D = kD, DFF = kDFF, H = kH, KC = kKC;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, wbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int sq = 0;
  while (sq + 1 < tb.n_seq && (int)blockIdx.x >= tb.first_tile[sq + 1]) ++sq;
  const SeqTcArgs& a = tb.a[sq];
  const int tile = blockIdx.x - tb.first_tile[sq];
  const int B = a.cfg.batch;
  const int b = tile * 128 + tid;
  const bool live = b < B;
  float* fv = reinterpret_cast<float*>(smem + L::oFV);

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&wbar, 1);
    mbar_fence_init();
    // operands by bulk async copies: this tile's context image and the three weight images
    constexpr uint32_t nA = 128 * H * D * 2, nWv = H * D * D * 2, nW = 2 * D * DFF * 2;
    mbar_expect_tx(&wbar, nA + nWv + nW);
    bulk_g2s(smem + L::oA, reinterpret_cast<const uint8_t*>(a.ctx) + (size_t)tile * nA, nA, &wbar);
    bulk_g2s(smem + L::oWv, a.prepared + prep_off_wvbd(D, DFF, H), nWv, &wbar);
    bulk_g2s(smem + L::oW1, a.prepared + prep_wqkv(D), nW, &wbar);          // w1 | w2 are contiguous
  }
  // meanwhile: this sample's target item rows and length, the small vectors
  const int nf = a.cfg.n_feats;
  const int zp = a.cfg.zero_pad ? 1 : 0;
  bool has = false;
  float dvec[D];
#pragma unroll
  for (int i = 0; i < D; ++i) dvec[i] = 0.f;
  if (live) {
    const int32_t* ol = a.in.offsets[nf - 1];
    int64_t rws[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) rws[c] = (int64_t)__ldg(a.in.item_ids[a.chunk_feat[c]] + b) - zp;
    has = (__ldg(ol + b + 1) - __ldg(ol + b)) > 0;
    const float sqrt_d = sqrtf((float)D);
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      const int f = a.chunk_feat[c];
      if (rws[c] >= 0 && rws[c] < a.in.rows[f]) {
        const float* src = a.in.table[f] + rws[c] * a.in.dim[f] + a.chunk_off[c];
        const float4 t0 = ldg4(src), t1 = ldg4(src + 4);
        dvec[c * 8 + 0] = t0.x * sqrt_d; dvec[c * 8 + 1] = t0.y * sqrt_d; dvec[c * 8 + 2] = t0.z * sqrt_d;
        dvec[c * 8 + 3] = t0.w * sqrt_d; dvec[c * 8 + 4] = t1.x * sqrt_d; dvec[c * 8 + 5] = t1.y * sqrt_d;
        dvec[c * 8 + 6] = t1.z * sqrt_d; dvec[c * 8 + 7] = t1.w * sqrt_d;
      }
    }
  }
  for (int i = tid; i < D; i += 128) {
    fv[L::vBV + i] = a.dbv[i];
    fv[L::vB2 + i] = a.b2[i];
    fv[L::vLN2 + i] = a.ln2_g[i];
    fv[L::vLN2 + D + i] = a.ln2_b[i];
    fv[L::vLN3 + i] = a.ln3_g[i];
    fv[L::vLN3 + D + i] = a.ln3_b[i];
  }
  for (int i = tid; i < DFF; i += 128) fv[L::vB1 + i] = a.b1[i];
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t dHi = desc_hi(128, kLayoutNone);
  const uint32_t dA = desc_lo(smem_u32(smem + L::oA), kROWB);
  const uint32_t dWv = desc_lo(smem_u32(smem + L::oWv), D * 16), dW1 = desc_lo(smem_u32(smem + L::oW1), DFF * 16),
                 dW2 = desc_lo(smem_u32(smem + L::oW2), D * 16);
  uint32_t phase = 0;

  // ---- o = ctx WvBD ----
  if (tid == 0) {
    mbar_wait(&wbar, 0);
    fence_after_sync();
    constexpr uint32_t idesc = make_idesc_bf16(128, D);
#pragma unroll
    for (int ks = 0; ks < H * D / 16; ++ks)
      mma_bf16_ss(tbase + L::tO, desc_join(dA + ks * (2 * kROWB / 16), dHi), desc_join(dWv + ks * (2 * D), dHi), idesc,
                  ks > 0);
    commit(&bar);
  }
  mbar_wait(&bar, phase);
  phase ^= 1;
  fence_after_sync();
  float av[D];
  {
#pragma unroll
    for (int blk = 0; blk < D / 32; ++blk) {
      uint32_t r[32];
      tmem_ld32(tmem_addr(tbase, L::tO + blk * 32), r);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e)
        av[blk * 32 + e] = __uint_as_float(r[e]) + (has ? fv[L::vBV + blk * 32 + e] : 0.f) + dvec[blk * 32 + e];
    }
    if (!live) {
#pragma unroll
      for (int i = 0; i < D; ++i) av[i] = 0.f;          // rows past the batch hold whatever the image held
    }
    ln64(av, fv + L::vLN3, fv + L::vLN3 + D);
    uint32_t pk[D / 2];
#pragma unroll
    for (int i = 0; i < D / 2; ++i) pk[i] = pack_bf16x2(av[2 * i], av[2 * i + 1]);
    tmem_st32(tmem_addr(tbase, L::tA), pk);
  }
  tmem_st_wait();
  fence_before_sync();
  __syncthreads();
  // ---- hidden = relu(av W1 + b1), packed in place ----
  if (tid == 0) {
    fence_after_sync();
    constexpr uint32_t idesc = make_idesc_bf16(128, DFF);
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks)
      mma_bf16_ts(tbase + L::tFF1, tbase + L::tA + ks * 8, desc_join(dW1 + ks * (2 * DFF), dHi), idesc, ks > 0);
    commit(&bar);
  }
  mbar_wait(&bar, phase);
  phase ^= 1;
  fence_after_sync();
  {
    uint32_t rh[2][32];
    tmem_ld32(tmem_addr(tbase, L::tFF1), rh[0]);
#pragma unroll
    for (int blk = 0; blk < DFF / 32; ++blk) {
      tmem_ld_wait();
      if (blk + 1 < DFF / 32) tmem_ld32(tmem_addr(tbase, L::tFF1 + blk * 32 + 32), rh[(blk + 1) & 1]);
      const uint32_t* r = rh[blk & 1];
      uint32_t pk[16];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 bb = *reinterpret_cast<const float4*>(fv + L::vB1 + blk * 32 + g * 4);
        pk[g * 2] = pack_bf16x2(fmaxf(__uint_as_float(r[g * 4]) + bb.x, 0.f), fmaxf(__uint_as_float(r[g * 4 + 1]) + bb.y, 0.f));
        pk[g * 2 + 1] = pack_bf16x2(fmaxf(__uint_as_float(r[g * 4 + 2]) + bb.z, 0.f), fmaxf(__uint_as_float(r[g * 4 + 3]) + bb.w, 0.f));
      }
      tmem_st16(tmem_addr(tbase, L::tFF1 + blk * 16), pk);
    }
  }
  tmem_st_wait();
  fence_before_sync();
  __syncthreads();
  // ---- f = hidden W2 ; u = LN2(f + b2 + av) ----
  if (tid == 0) {
    fence_after_sync();
    constexpr uint32_t idesc = make_idesc_bf16(128, D);
#pragma unroll
    for (int ks = 0; ks < DFF / 16; ++ks)
      mma_bf16_ts(tbase + L::tFF2, tbase + L::tFF1 + ks * 8, desc_join(dW2 + ks * (2 * D), dHi), idesc, ks > 0);
    commit(&bar);
  }
  mbar_wait(&bar, phase);
  phase ^= 1;
  fence_after_sync();
  {
    float y[D];
#pragma unroll
    for (int blk = 0; blk < D / 32; ++blk) {
      uint32_t r[32];
      tmem_ld32(tmem_addr(tbase, L::tFF2 + blk * 32), r);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) y[blk * 32 + e] = __uint_as_float(r[e]) + fv[L::vB2 + blk * 32 + e] + av[blk * 32 + e];
    }
    ln64(y, fv + L::vLN2, fv + L::vLN2 + D);
    float* so = reinterpret_cast<float*>(smem + L::oOut) + tid * (D + 1);
#pragma unroll
    for (int i = 0; i < D; ++i) so[i] = y[i];
  }
  fence_before_sync();
  __syncthreads();
  // coalesced write: one warp per sample row
  for (int r = warp; r < 128; r += 4) {
    const int bb = tile * 128 + r;
    if (bb >= B) break;
    const float* so = reinterpret_cast<const float*>(smem + L::oOut) + r * (D + 1);
    if (a.cfg.flags & DMT_SEQ_OUT_BF16) {             // straight into the bf16 MMoE input
      __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(a.out) + (int64_t)bb * a.out_ld;
      ob[lane] = __float2bfloat16(so[lane]);
      ob[32 + lane] = __float2bfloat16(so[32 + lane]);
    } else {
      a.out[(int64_t)bb * a.out_ld + lane] = so[lane];
      a.out[(int64_t)bb * a.out_ld + 32 + lane] = so[32 + lane];
    }
  }
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// (A programmatic dependent launch behind the tile kernel was measured and dropped: the early tail CTAs -- 113 KB of
// shared memory and all of tensor memory each -- take the SMs that the dense / pooled kernels fill while the tile
// kernel's last CTAs retire: 337 -> 347 us per step.)
int launch_tails(const TailBatch& tb, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(seq_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TailLayout::total);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(seq_tail_kernel)");
  seq_tail_kernel<<<tb.first_tile[tb.n_seq], 128, TailLayout::total, st>>>(tb);
  DMT_CUDA_LAUNCH_CHECK("seq_tail_kernel");
  return DMT_OK;
}

}  // namespace

// shared-memory budget of the v2 kernel: the position table is the only size that depends on the configuration
bool seq_tc2_supported(const dmt_seq_cfg* cfg) {
  return Tc2Layout<64>::oPos + cfg->maxlen * kD * 2 + 64 + 1024 <= 227 * 1024 && cfg->n_feats <= kKC;   // + static smem
}

// decoder contexts: one bf16 A-operand image (128 samples x H*D) per tail tile
size_t seq_tc2_ctx_bytes(const dmt_seq_cfg* cfg) { return (size_t)((cfg->batch + 127) / 128) * (128 * kH * kD * 2) + 256; }

// dmt_seq_tail_fwd: the deferred tails of several sequences, one launch
int seq_tails_launch(int n, const SeqTcArgs* args, cudaStream_t st) {
  TailBatch tb;
  tb.n_seq = n;
  int t = 0;
  for (int i = 0; i < n; ++i) {
    tb.a[i] = args[i];
    tb.first_tile[i] = t;
    t += (args[i].cfg.batch + 127) / 128;
  }
  for (int i = n; i <= DMT_MAX_TAIL_SEQS; ++i) tb.first_tile[i] = t;
  if (t == 0) return DMT_OK;
  return launch_tails(tb, st);
}

// diagnostics (dmt_debug_seq_timer): CUDA events around every seq_encode_multi_kernel launch, on its launch stream --
// bench.py's roofline of the dominant kernel is that kernel's own live duration, not the stage's three launches
constexpr int kTimerCap = 4096;
static struct {
  bool on = false;
  int n = 0;
  cudaEvent_t e0[kTimerCap], e1[kTimerCap];
} g_seq_timer;

int seq_timer_enable(int on) {
  for (int i = 0; i < g_seq_timer.n; ++i) {
    cudaEventDestroy(g_seq_timer.e0[i]);
    cudaEventDestroy(g_seq_timer.e1[i]);
  }
  g_seq_timer.n = 0;
  g_seq_timer.on = on != 0;
  return DMT_OK;
}

int seq_timer_read(float* total_ms, int32_t* launches) {
  float sum = 0.f;
  for (int i = 0; i < g_seq_timer.n; ++i) {
    cudaError_t e = cudaEventSynchronize(g_seq_timer.e1[i]);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize(dmt_debug_seq_timer_read)");
    float ms = 0.f;
    e = cudaEventElapsedTime(&ms, g_seq_timer.e0[i], g_seq_timer.e1[i]);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventElapsedTime(dmt_debug_seq_timer_read)");
    sum += ms;
  }
  *total_ms = sum;
  *launches = g_seq_timer.n;
  return DMT_OK;
}

// workspace tail of the bucketed launch: [perm: batch int32 | counts: 4 int32]
size_t seq_tc_sched_bytes(const dmt_seq_cfg* cfg) { return ((size_t)cfg->batch * 4 + 16 + 255) / 256 * 256; }

// dmt_seq_encode_multi_fwd: every behaviour sequence of the step -- length classes, ONE tile-kernel launch, one tail
// launch
// wait_before_encode (optional): the stream waits for this event AFTER the length-class kernel -- which only reads the
// batch's offsets -- and before the tile kernel (dmt_forward_bf16: the classes of step i + 1 are formed while step i's
// MMoE still runs)
int seq_encode_multi_launch(int n, const SeqTcArgs* args, void* const* scheds, bool defer_tail, cudaEvent_t wait_before_encode,
                            cudaStream_t st) {
  SeqMultiArgs m;
  BucketArgs ba;
  memset(&m, 0, sizeof(m));
  memset(&ba, 0, sizeof(ba));
  m.n_seq = n;
  long long ub_tiles = 0;                             // upper bound of the tile count (the classes are known on the device only)
  int maxlen = 0;
  for (int i = 0; i < n; ++i) {
    m.a[i] = args[i];
    int32_t* perm = static_cast<int32_t*>(scheds[i]);
    int32_t* counts = perm + args[i].cfg.batch;
    m.perm[i] = perm;
    m.counts[i] = counts;
    ba.offs[i] = args[i].in.offsets[args[i].cfg.n_feats - 1];
    ba.perm[i] = perm;
    ba.counts[i] = counts;
    ba.batch[i] = args[i].cfg.batch;
    ba.maxlen[i] = args[i].cfg.maxlen;
    ub_tiles += (args[i].cfg.batch + 1) / 2 + 2;
    if (args[i].cfg.maxlen > maxlen) maxlen = args[i].cfg.maxlen;
  }
  const int total = Tc2Layout<64>::oPos + maxlen * kD * 2 + 64;
  const int sms = sm_count_cached();
  const long long pairs = (ub_tiles + 1) / 2;
  const int grid = pairs < sms ? (int)pairs : sms;
  const bool dbg = args[0].dbg != nullptr;
  auto kern = dbg ? seq_encode_multi_kernel<true> : seq_encode_multi_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, total);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(seq_encode_multi_kernel)");
  const bool timed = g_seq_timer.on && g_seq_timer.n < kTimerCap;
  if (timed) {      // (created before the launches: the event records then sit back to back with the kernel launch)
    cudaEventCreate(&g_seq_timer.e0[g_seq_timer.n]);
    cudaEventCreate(&g_seq_timer.e1[g_seq_timer.n]);
  }
  seq_bucket_kernel<<<n, 1024, 0, st>>>(ba);
  DMT_CUDA_LAUNCH_CHECK("seq_bucket_kernel");
  if (wait_before_encode) {
    e = cudaStreamWaitEvent(st, wait_before_encode, 0);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamWaitEvent(seq_encode_multi_launch)");
  }
  if (timed) cudaEventRecord(g_seq_timer.e0[g_seq_timer.n], st);
  if (!timed && !wait_before_encode) {
    // directly behind the length-class kernel: programmatic dependent launch, the prologue of sequence 0 runs beside it
    e = launch_pdl(kern, dim3(grid), dim3(kT3Threads), (size_t)total, st, m);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(seq_encode_multi_kernel)");
  } else {
    kern<<<grid, kT3Threads, total, st>>>(m);
  }
  if (timed) cudaEventRecord(g_seq_timer.e1[g_seq_timer.n++], st);
  DMT_CUDA_LAUNCH_CHECK("seq_encode_multi_kernel");
  if (defer_tail) return DMT_OK;
  return seq_tails_launch(n, args, st);
}

}  // namespace dmt
