"""Host mirror of the dropout hash (cikm2020_dmt_b200/dropout.py): vectorised torch version == scalar restatement,
keep fraction, inverted-dropout scale."""
import torch

from cikm2020_dmt_b200 import dropout as DO


def _scalar(rate, seed, site, idx):
    s0 = DO.fmix32(seed ^ ((site * 0x9E3779B9) & 0xFFFFFFFF))
    import struct
    rate32 = struct.unpack("f", struct.pack("f", rate))[0]
    thresh = min(int(rate32 * 4294967296.0), 4294967295)
    r = DO.fmix32((DO.fmix32((idx & 0xFFFFFFFF) ^ s0) + ((s0 * 0x9E3779B1) & 0xFFFFFFFF)) & 0xFFFFFFFF)
    return 0.0 if r < thresh else 1.0 / (1.0 - rate32)


def test_vectorised_hash_equals_scalar():
    idx = torch.tensor([0, 1, 2, 63, 64, 12345, 2 ** 31 - 1, 2 ** 32 - 1, 987654321])
    for rate, seed, site in [(0.1, 1, 0), (0.5, 0xC0FFEE, 11), (0.9, 2 ** 32 - 1, 6)]:
        got = DO.multiplier(rate, seed, site, idx)
        want = torch.tensor([_scalar(rate, seed, site, int(i)) for i in idx], dtype=torch.float64)
        assert torch.allclose(got, want, rtol=1e-7, atol=0)


def test_keep_fraction_and_mean_preservation():
    idx = torch.arange(400000)
    for rate in (0.1, 0.5):
        m = DO.multiplier(rate, 42, 3, idx)
        keep = (m > 0).double().mean().item()
        assert abs(keep - (1 - rate)) < 5e-3
        assert abs(m.mean().item() - 1.0) < 1e-2            # inverted dropout preserves the expectation
    assert torch.equal(DO.multiplier(0.0, 1, 1, idx[:10]), torch.ones(10, dtype=torch.float64))


def test_sites_and_steps_decorrelate():
    idx = torch.arange(100000)
    a = DO.multiplier(0.5, DO.step_seed(7, 1, 0), 0, idx) > 0
    b = DO.multiplier(0.5, DO.step_seed(7, 2, 0), 0, idx) > 0
    c = DO.multiplier(0.5, DO.step_seed(7, 1, 0), 1, idx) > 0
    for x in (b, c):
        agree = (a == x).double().mean().item()
        assert abs(agree - 0.5) < 1e-2
