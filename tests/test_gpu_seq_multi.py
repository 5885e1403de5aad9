"""GPU parity of `dmt_seq_encode_multi_fwd`: every behaviour sequence in ONE persistent tile-kernel launch over
length-bucketed tiles (trans_core x 3, mmoe_transformer_unbias.py:150-216) against the CPU oracle, against the
per-sequence launches, and the device-side length-class schedule itself (bit-exact integer work).

Tolerance: the bf16 tensor-core path's (tests/test_gpu_parity.py): interest vectors atol 6e-2, mean abs error 1e-2.
"""
import ctypes as C

import pytest
import torch

from conftest import make_plan, SMALL_ROWS

pytestmark = pytest.mark.gpu

ATOL_BF16, MEAN_BF16 = 6e-2, 1e-2


def _models(conf_file, batch, seed, overrides=None, **gen):
    from cikm2020_dmt_b200.params import ParamStore
    from cikm2020_dmt_b200.data import synthetic_batch, batch_to
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    from oracle import dmt_oracle as O
    conf, plan = make_plan(conf_file, overrides=overrides)
    store = ParamStore(plan, device="cuda", seed=seed + 1).randomize_(seed + 2)
    tc = mmoe_transformer_unbias(plan, params=store, precision="bf16")
    host = synthetic_batch(plan, batch, seed=seed + 3, table_rows=SMALL_ROWS, **gen)
    P = O.params_from_store(store, torch.float32)
    return plan, tc, host, batch_to(host, "cuda"), P, O


def _multi(tc, plan, dev, B):
    x_ld = (plan.mmoe_in + 3) // 4 * 4
    x = torch.zeros(B, x_ld, device="cuda")
    tc._stream_h = None
    keep = tc.seq_encode_multi(dev, x, x_ld, B)
    torch.cuda.synchronize()
    c0 = plan.interest_col
    return x[:, c0:c0 + len(plan.sequences) * plan.d_model].clone(), keep


def _single(tc, plan, dev, B):
    out = torch.zeros(B, len(plan.sequences) * plan.d_model, device="cuda")
    for s in range(len(plan.sequences)):
        tc.seq_encode(dev, s, out.data_ptr() + 4 * s * plan.d_model, out.stride(0), B)
    torch.cuda.synchronize()
    return out


def _check(out, want):
    err = (out.double().cpu() - want.double()).abs()
    assert err.max().item() < ATOL_BF16 and err.mean().item() < MEAN_BF16, (err.max().item(), err.mean().item())


@pytest.mark.parametrize("gen", [dict(), dict(full_length=True), dict(seq_lens=[16, 30, 10])])
def test_multi_matches_oracle(gen):
    B = 37
    plan, tc, host, dev, P, O = _models("dmt_d64.conf", B, 31, **gen)
    want = O.trans_core(plan, P, O.generate_data(plan, P, host), training=False)
    out, _ = _multi(tc, plan, dev, B)
    _check(out, want)


@pytest.mark.parametrize("lens", [[1, 50, 1, 17, 50], [16, 17, 32, 33, 16, 32, 33, 17, 1, 50, 2],
                                  [33], [3], [20], [50] * 9, [5] * 19, [24] * 7])
def test_multi_class_boundaries_and_partial_tiles(lens):
    """16|17 and 32|33 are the class boundaries; single-sample classes; classes that do not fill their last tile;
    classes with no sample at all."""
    from cikm2020_dmt_b200.data import SparseIds, batch_to
    B = len(lens)
    plan, tc, host, dev, P, O = _models("dmt_d64.conf", B, 35)
    g = torch.Generator().manual_seed(12)
    off = torch.zeros(B + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(torch.tensor(lens), 0)
    seq = plan.sequences[0]
    for f, uf in enumerate(seq.user_features):
        V = plan.tables[seq.tables[f]].rows
        vals = torch.randint(1, V, (int(off[-1]),), generator=g, dtype=torch.int32)
        vals[0] = 0
        host[uf] = SparseIds(vals, off)
    dev = batch_to(host, "cuda")
    want = O.trans_core(plan, P, O.generate_data(plan, P, host), training=False)
    out, _ = _multi(tc, plan, dev, B)
    _check(out, want)


@pytest.mark.parametrize("maxlen", [40, 50, 55])
def test_multi_many_tiles_schedule_and_determinism(maxlen):
    """2400 samples: every tile group walks several tiles of several segments (pipeline drain / prime at the segment
    boundaries, weight-image swap at the sequence boundaries, mbarrier phases carried across them)."""
    from cikm2020_dmt_b200 import keys as K
    from cikm2020_dmt_b200 import abi
    B = 2400
    plan, tc, host, dev, P, O = _models("dmt_d64.conf", B, 77, overrides={(K.MODEL, "transformer_maxlen_k"): str(maxlen)})
    assert plan.maxlen_k == maxlen
    want = O.trans_core(plan, P, O.generate_data(plan, P, host), training=False)
    out, keep = _multi(tc, plan, dev, B)
    _check(out, want)
    out2, _ = _multi(tc, plan, dev, B)
    assert torch.equal(out, out2)                      # no atomics, fixed schedule: bit-identical run to run
    # per-sequence launches (64-row slots for every sample): same math up to the softmax summation order
    ref = _single(tc, plan, dev, B)
    assert (out - ref).abs().max().item() < ATOL_BF16
    # the device-side schedule: perm = stable counting sort of the samples by length class, counts = class sizes
    lib = abi.load()
    for s, seq in enumerate(plan.sequences):
        cfg = tc._seq_cfg(dev, seq, B, abi.PRECISION_BF16)
        total = lib.dmt_seq_encode_workspace_bytes(C.byref(cfg), 0)
        sched = (B * 4 + 16 + 255) // 256 * 256
        ws = tc._prepared[s][1]
        words = ws[total - sched: total - sched + 4 * (B + 3)].cpu().view(torch.int32)
        perm, counts = words[:B].long(), words[B:B + 3].tolist()
        off = host[seq.user_features[-1]].offsets.long()
        ln = torch.clamp(off[1:] - off[:-1], max=maxlen)
        cls = torch.where(ln > 32, 0, torch.where(ln > 16, 1, 2))
        assert counts == [int((cls == k).sum()) for k in range(3)]
        want_perm = torch.cat([torch.nonzero(cls == k).flatten() for k in range(3)])
        assert torch.equal(perm, want_perm)


def test_multi_inference_matches_oracle_and_single_launch_path():
    """The plugin's inference() routes the bf16 path through the multi launch; DMT_SEQ_MULTI=0 keeps the per-sequence
    launches: both within the bf16 logit tolerance of the oracle."""
    B = 300
    plan, tc, host, dev, P, O = _models("dmt_d64.conf", B, 41)
    assert tc.seq_multi
    (yr, yb) = tc.inference(dev, is_train=False)
    (wr, wb) = O.inference(plan, P, host, is_train=False)
    torch.cuda.synchronize()
    for t in range(2):
        err = (yr[t].double().cpu() - wr[t].double()).abs()
        assert bool((err <= 5e-2 + 2e-2 * wr[t].double().abs()).all()), err.max().item()
    got = [y.clone() for y in yr]
    tc.seq_multi = False
    (yr1, _) = tc.inference(dev, is_train=False)
    torch.cuda.synchronize()
    for t in range(2):
        assert (got[t] - yr1[t]).abs().max().item() < 5e-2


def test_multi_rejects_bad_arguments():
    from cikm2020_dmt_b200 import abi
    lib = abi.load()
    assert lib.dmt_seq_encode_multi_fwd(abi.MAX_TAIL_SEQS + 1, None, None, None, None, None, None, None, None) == -1
    assert lib.dmt_seq_encode_multi_fwd(0, None, None, None, None, None, None, None, None) == 0
    assert lib.dmt_seq_encode_multi_fwd(1, None, None, None, None, None, None, None, None) == -1
