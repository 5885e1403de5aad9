"""Host-side mirror of csrc/dropout.cuh: the counter-based keep mask of the training-mode dropout sites
(TransformerModel.py:101,151; TransformerModel_util.py:51; mmoe_transformer_unbias.py:272,280).

The device kernels recompute the mask from (seed, site, element index) in forward and backward; this module
derives the per-step / per-sequence seeds and restates the hash in torch integer arithmetic so that the CPU oracle
can be driven with exactly the same mask (tests/test_gpu_dropout.py)."""
import torch

SITE_ENC_IN, SITE_DEC_IN, SITE_SELF_PROBS, SITE_VANILLA_PROBS, SITE_BIAS = 0, 1, 2, 6, 10
_M32 = 0xFFFFFFFF


def fmix32(h: int) -> int:
    h &= _M32
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & _M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & _M32
    h ^= h >> 16
    return h


def step_seed(base_seed: int, step: int, stream: int = 0) -> int:
    """Seed of one training step (`stream` separates the behaviour sequences / the bias tower / the ranks)."""
    return fmix32(fmix32(base_seed ^ (step * 0x9E3779B9)) + stream * 0x7F4A7C15)


def _fmix32_t(h: torch.Tensor) -> torch.Tensor:
    h = h & _M32
    h = h ^ (h >> 16)
    h = (h * 0x85EBCA6B) & _M32
    h = h ^ (h >> 13)
    h = (h * 0xC2B2AE35) & _M32
    h = h ^ (h >> 16)
    return h


def multiplier(rate: float, seed: int, site: int, idx: torch.Tensor) -> torch.Tensor:
    """0 or 1/(1-rate) for every element index in `idx` (int64 tensor of uint32 values): Dropout::mult."""
    if rate <= 0:
        return torch.ones(idx.shape, dtype=torch.float64)
    s0 = fmix32(seed ^ ((site * 0x9E3779B9) & _M32))
    t = float(torch.tensor(rate, dtype=torch.float32)) * 4294967296.0     # the device sees the rate as fp32
    thresh = 4294967295 if t >= 4294967295.0 else int(t)
    r = _fmix32_t((_fmix32_t((idx.to(torch.int64) & _M32) ^ s0) + ((s0 * 0x9E3779B1) & _M32)) & _M32)
    scale = 1.0 / (1.0 - float(torch.tensor(rate, dtype=torch.float32)))
    return torch.where(r < thresh, torch.zeros((), dtype=torch.float64), torch.full((), scale, dtype=torch.float64))
