// Training-mode dropout: tf.layers.dropout / tf.nn.dropout are inverted dropout -- an element is
// kept with probability 1 - rate and scaled by 1 / (1 - rate).  Sites of the hot path:
//   encoder input (TransformerModel.py:101), decoder input (:151), attention probabilities after the query
//   mask (TransformerModel_util.py:51), the two hidden layers of the bias tower (mmoe_transformer_unbias.py:
//   272,280).
// TF's RNG stream cannot be reproduced, so the mask comes from a counter-based hash of (seed, site, element
// index): the backward recomputes it instead of storing it, and the CPU oracle can be driven with the very
// same mask (tests/test_gpu_dropout.py restates this function in torch), which makes training mode
// parity-testable exactly.
#pragma once
#include <stdint.h>

namespace dmt {

enum DropoutSite : uint32_t {
  kSiteEncIn = 0,        // + element (token row * d + c)
  kSiteDecIn = 1,        // sample * d + c
  kSiteSelfProbs = 2,    // + block; ((sample * H + h) * LP + q) * LP + k
  kSiteVanillaProbs = 6, // + block; (sample * H + h) * LP + k
  kSiteBias = 10         // + layer; sample * units + n
};

__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

struct Dropout {
  uint32_t s0;       // per-site seed
  uint32_t thresh;   // drop iff rand < thresh
  float scale;       // 1 / (1 - rate)
  bool on;

  __host__ __device__ Dropout() : s0(0), thresh(0), scale(1.f), on(false) {}
  __host__ __device__ Dropout(float rate, uint32_t seed, uint32_t site) {
    on = rate > 0.f;
    s0 = fmix32(seed ^ (site * 0x9E3779B9u));
    const double t = (double)rate * 4294967296.0;
    thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    scale = on ? 1.0f / (1.0f - rate) : 1.0f;
  }
  // multiplier of element `idx`: 0 (dropped) or 1 / (1 - rate)
  __host__ __device__ __forceinline__ float mult(uint32_t idx) const {
    if (!on) return 1.0f;
    const uint32_t r = fmix32(fmix32(idx ^ s0) + s0 * 0x9E3779B1u);
    return r < thresh ? 0.f : scale;
  }
};

}  // namespace dmt
