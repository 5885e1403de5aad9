"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances:
  * gather rows / indices: bit-exact;
  * fp32 kernels vs the fp64 oracle: logits atol 1e-4 (+ rtol 1e-4), loss rtol 1e-5.
"""
import ctypes as C

import pytest
import torch

from conftest import make_plan

pytestmark = pytest.mark.gpu

ATOL_F32 = 1e-4
RTOL_F32 = 1e-4


def _setup(conf_file, batch, seed=0, rows=None, **gen):
    from cikm2020_dmt_b200.params import ParamStore
    from cikm2020_dmt_b200.data import synthetic_batch, batch_to
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    from oracle import dmt_oracle as O
    conf, plan = make_plan(conf_file, rows=rows)
    store = ParamStore(plan, device="cuda", seed=seed + 1).randomize_(seed + 2)
    model = mmoe_transformer_unbias(plan, params=store)
    from conftest import SMALL_ROWS
    host = synthetic_batch(plan, batch, seed=seed + 3, table_rows=SMALL_ROWS if rows is None else rows, **gen)
    P = O.params_from_store(store)
    return plan, model, host, batch_to(host, "cuda"), P, O


def _close(got, want, atol=ATOL_F32, rtol=RTOL_F32):
    got = got.detach().double().cpu()
    want = want.detach().double().cpu()
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    assert bool((err <= tol).all()), "max abs err %.3e (tol %.1e)" % (err.max().item(), atol)


# ------------------------------------------------------------------ gather: bit-exact
@pytest.mark.parametrize("dim", [32, 8, 5])
@pytest.mark.parametrize("zero_pad", [0, 1])
def test_embed_gather_identity_table_bit_exact(dim, zero_pad):
    """SURVEY 4 KAT (i): table[r, :] = r  =>  output == idx-1 (zero_pad; 0 -> 0-vector) or idx."""
    from cikm2020_dmt_b200 import abi
    lib = abi.load()
    rows, n = 5003, 40000
    table = torch.arange(rows, dtype=torch.float32, device="cuda")[:, None].repeat(1, dim).contiguous()
    g = torch.Generator().manual_seed(dim * 2 + zero_pad)
    ids = torch.randint(0, rows + (1 if zero_pad else 0), (n,), generator=g, dtype=torch.int32)
    ids[:3] = torch.tensor([0, 1, rows - 1])
    ids_d = ids.cuda()
    out = torch.empty(n, dim, device="cuda")
    abi.check(lib.dmt_embed_gather(table.data_ptr(), rows, dim, ids_d.data_ptr(), n, zero_pad, out.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    want = ids.float() - (1.0 if zero_pad else 0.0)
    want = torch.where(want < 0, torch.zeros_like(want), want)
    assert torch.equal(out.cpu(), want[:, None].expand(n, dim))


def test_embed_gather_random_table_bit_exact_vs_oracle():
    from cikm2020_dmt_b200 import abi
    from oracle import dmt_oracle as O
    lib = abi.load()
    g = torch.Generator().manual_seed(7)
    table = torch.randn(3001, 32, generator=g)
    ids = torch.randint(0, 3002, (10000,), generator=g, dtype=torch.int32)
    out = torch.empty(10000, 32, device="cuda")
    td, idd = table.cuda(), ids.cuda()
    abi.check(lib.dmt_embed_gather(td.data_ptr(), 3001, 32, idd.data_ptr(), 10000, 1, out.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    want = O.embedding(table, zero_pad=True)[ids.long()]
    assert torch.equal(out.cpu(), want)


def test_embed_gather_empty_and_errors():
    from cikm2020_dmt_b200 import abi
    lib = abi.load()
    t = torch.zeros(4, 8, device="cuda")
    assert lib.dmt_embed_gather(t.data_ptr(), 4, 8, t.data_ptr(), 0, 1, t.data_ptr(), None) == 0
    assert lib.dmt_embed_gather(None, 4, 8, t.data_ptr(), 1, 1, t.data_ptr(), None) == -1
    assert b"null" in lib.dmt_last_error()


# ------------------------------------------------------------------ per-stage parity
@pytest.mark.parametrize("conf_file", ["dmt_d64.conf", "dmt.conf"])
def test_seq_encode_matches_oracle(conf_file):
    plan, model, host, dev, P, O = _setup(conf_file, 33)
    seq_data = O.generate_data(plan, P, host)
    want = O.trans_core(plan, P, seq_data, training=False)
    B = 33
    out = torch.zeros(B, len(plan.sequences) * plan.d_model, device="cuda")
    for s in range(len(plan.sequences)):
        model.seq_encode(dev, s, out.data_ptr() + 4 * s * plan.d_model, out.stride(0), B)
    torch.cuda.synchronize()
    _close(out, want)


def test_seq_encode_edge_lengths():
    """length-1 'unknow' sequences (index 0 -> zero row), full-length 50, and ragged in one batch."""
    from cikm2020_dmt_b200.data import SparseIds, batch_to
    plan, model, host, dev, P, O = _setup("dmt_d64.conf", 6, seed=5)
    lens = [1, 50, 1, 17, 50, 2]
    g = torch.Generator().manual_seed(11)
    off = torch.zeros(7, dtype=torch.int32)
    off[1:] = torch.cumsum(torch.tensor(lens), 0)
    seq = plan.sequences[0]
    for f, uf in enumerate(seq.user_features):
        V = plan.tables[seq.tables[f]].rows
        vals = torch.randint(1, V, (int(off[-1]),), generator=g, dtype=torch.int32)
        vals[0] = 0            # sample 0: the single token is 'unknow'
        vals[int(off[2])] = 0  # sample 2 too
        host[uf] = SparseIds(vals, off)
    dev = batch_to(host, "cuda")
    want = O.trans_core(plan, P, O.generate_data(plan, P, host), training=False)
    out = torch.zeros(6, len(plan.sequences) * plan.d_model, device="cuda")
    for s in range(len(plan.sequences)):
        model.seq_encode(dev, s, out.data_ptr() + 4 * s * plan.d_model, out.stride(0), 6)
    torch.cuda.synchronize()
    _close(out, want)


def test_pool_mean_matches_oracle_with_weights():
    from cikm2020_dmt_b200.data import batch_to
    plan, model, host, dev, P, O = _setup("dmt.conf", 29, seed=3)
    g = torch.Generator().manual_seed(5)
    for p in plan.pooled[5:9]:     # attach `<feature>Wts` to a few features
        host[p.feature + "Wts"] = torch.rand(host[p.feature].values.numel(), generator=g) + 0.25
    dev = batch_to(host, "cuda")
    want = O.embedding_combiner(plan, P, host)
    x = torch.zeros(29, plan.interest_col, device="cuda")
    from cikm2020_dmt_b200 import abi
    abi.check(model.lib.dmt_copy_dense_features(dev["features"].data_ptr(), 29, plan.feature_dim, x.data_ptr(),
                                                x.stride(0), torch.cuda.current_stream().cuda_stream))
    model.pool_mean(dev, plan.pooled, False, x, 29)
    torch.cuda.synchronize()
    _close(x, want, atol=1e-6, rtol=1e-6)
    # the dense block is a copy: bit-exact
    assert torch.equal(x[:, :plan.feature_dim].cpu(), host["features"])


def test_mmoe_matches_oracle():
    plan, model, host, dev, P, O = _setup("dmt_d64.conf", 150, seed=9)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(150, plan.mmoe_in, generator=g) * 0.3
    tasks = O.expert_gate(plan, P, x.double())
    want = torch.stack([O.build_tower(plan, P, t, O.TASK_NAMES[i]).squeeze(1) for i, t in enumerate(tasks)])
    xd = x.cuda()
    logits = torch.zeros(2, 150, device="cuda")
    model.mmoe(xd, 150, logits)
    torch.cuda.synchronize()
    _close(logits, want, atol=2e-4, rtol=1e-4)


# ------------------------------------------------------------------ end to end
@pytest.mark.parametrize("conf_file,batch", [("dmt_d64.conf", 64), ("dmt.conf", 37)])
def test_inference_and_loss_match_oracle(conf_file, batch):
    from cikm2020_dmt_b200.inference import Inference
    plan, model, host, dev, P, O = _setup(conf_file, batch, seed=21)
    (yr, yb) = model.inference(dev, is_train=False)
    (wr, wb) = O.inference(plan, P, host, is_train=False)
    torch.cuda.synchronize()
    _close(yr[0], wr[0], atol=2e-4)
    _close(yr[1], wr[1], atol=2e-4)
    _close(yb, wb, atol=1e-5)
    assert yr[0].shape == (batch, 1) and yb.shape == (batch, 1)
    for unbias in ("two_head_add", "two_head_multiply"):
        for rel in ("ctr", "ctr_rel"):
            loss, probs, dlog = model.loss((yr, yb), dev["mask"], loss_unbias_method=unbias,
                                           loss_ctr_rel_method=rel, want_probs=True, want_grads=True)
            want = O.logit_loss_unbias(plan, (wr, wb), host["mask"], unbias, rel)
            torch.cuda.synchronize()
            assert abs(loss.item() - want.item()) <= 2e-4 * max(1.0, abs(want.item()))
            p_ctr, p_cvr = O.probabilities((wr, wb), unbias)
            _close(probs[0], p_ctr.squeeze(1), atol=1e-4)
            _close(probs[1], p_cvr.squeeze(1), atol=1e-4)
    # is_predict returns only the relevance logits (mmoe_transformer_unbias.py:312-316)
    yr2 = model.inference(dev, is_train=False, is_predict=True)
    assert isinstance(yr2, tuple) and len(yr2) == 2


def test_loss_gradients_match_oracle_autograd():
    plan, model, host, dev, P, O = _setup("dmt_d64.conf", 48, seed=4)
    g = torch.Generator().manual_seed(3)
    click = torch.randn(48, 1, generator=g, dtype=torch.float64, requires_grad=True)
    order = (torch.randn(48, 1, generator=g, dtype=torch.float64) - 2).requires_grad_(True)
    ybias = (torch.randn(48, 1, generator=g, dtype=torch.float64) * 0.5).requires_grad_(True)
    for unbias in ("two_head_add", "two_head_multiply"):
        for p in (click, order, ybias):
            p.grad = None
        want = O.logit_loss_unbias(plan, ((click, order), ybias), host["mask"], unbias, "ctr_rel")
        want.backward()
        logits = ((click.detach().float().cuda(), order.detach().float().cuda()), ybias.detach().float().cuda())
        loss, probs, dlog = model.loss(logits, dev["mask"], loss_unbias_method=unbias,
                                       loss_ctr_rel_method="ctr_rel", want_probs=True, want_grads=True)
        torch.cuda.synchronize()
        assert abs(loss.item() - want.item()) <= 1e-5 * max(1.0, abs(want.item()))
        _close(dlog[0], click.grad.squeeze(1), atol=1e-6, rtol=1e-4)
        _close(dlog[1], order.grad.squeeze(1), atol=1e-6, rtol=1e-4)
        _close(dlog[2], ybias.grad.squeeze(1), atol=1e-6, rtol=1e-4)


def test_sample_output_is_invariant_to_batch_composition():
    """SURVEY 0.4 / KAT (iv): padded rows are inert, so a sample's logits equal its single-sample run."""
    from cikm2020_dmt_b200.data import SparseIds, batch_to
    plan, model, host, dev, P, O = _setup("dmt_d64.conf", 16, seed=8)
    (yr, yb) = model.inference(dev, is_train=False)
    full = torch.stack([yr[0].squeeze(1), yr[1].squeeze(1), yb.squeeze(1)]).clone()
    for b in (0, 7, 15):
        one = {}
        for k, v in host.items():
            if isinstance(v, SparseIds):
                lo, hi = int(v.offsets[b]), int(v.offsets[b + 1])
                one[k] = SparseIds(v.values[lo:hi].clone(), torch.tensor([0, hi - lo], dtype=torch.int32))
            else:
                one[k] = v[b:b + 1].clone()
        (r1, b1) = model.inference(batch_to(one, "cuda"), is_train=False)
        got = torch.stack([r1[0].squeeze(1), r1[1].squeeze(1), b1.squeeze(1)])
        assert torch.equal(got[:, 0].cpu(), full[:, b].cpu())   # same arithmetic, same order: bit-identical


def test_full_size_batch_properties():
    """BASELINE config 2 size (B=4096, full tables are not needed for the property): finite outputs,
    and a strided subsample agrees with the oracle."""
    plan, model, host, dev, P, O = _setup("dmt_d64.conf", 4096, seed=13)
    (yr, yb) = model.inference(dev, is_train=False)
    torch.cuda.synchronize()
    assert torch.isfinite(yr[0]).all() and torch.isfinite(yr[1]).all() and torch.isfinite(yb).all()
    from cikm2020_dmt_b200.data import SparseIds
    idx = list(range(0, 4096, 257))
    sub = {}
    for k, v in host.items():
        if isinstance(v, SparseIds):
            rows = [v.values[int(v.offsets[b]):int(v.offsets[b + 1])].tolist() for b in idx]
            sub[k] = SparseIds.from_lists(rows)
        else:
            sub[k] = v[idx]
    (wr, wb) = O.inference(plan, P, sub, is_train=False)
    _close(yr[0][idx], wr[0], atol=2e-4)
    _close(yr[1][idx], wr[1], atol=2e-4)
    _close(yb[idx], wb, atol=1e-5)


def test_abi_rejects_bad_arguments():
    from cikm2020_dmt_b200 import abi
    lib = abi.load()
    cfg = abi.SeqCfg(4, 64, 256, 3, 1, 1, 50, 1, 5, 0)     # 64 % 3 != 0
    assert lib.dmt_seq_encode_fwd(C.byref(cfg), C.byref(abi.SeqInput()), C.byref(abi.SeqWeights()), None, 64,
                                  None, 0, None) == -1
    cfg = abi.SeqCfg(4, 64, 256, 2, 1, 1, 500, 1, 5, 0)    # maxlen > DMT_MAX_SEQ_LEN
    assert lib.dmt_seq_encode_fwd(C.byref(cfg), C.byref(abi.SeqInput()), C.byref(abi.SeqWeights()), None, 64,
                                  None, 0, None) == -2
    assert b"maxlen" in lib.dmt_last_error()


# ------------------------------------------------------------------ bf16 tensor-core path
# Tolerance (SURVEY 8c, bf16 tensor-core path): interest vectors are LayerNorm outputs of O(1)
# magnitude computed from bf16-rounded operands with fp32 accumulation: atol 6e-2, and the mean
# absolute error must stay below 1e-2; logits atol 5e-2 / rtol 2e-2.
ATOL_BF16, MEAN_BF16 = 6e-2, 1e-2


def _bf16_model(plan, store):
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    return mmoe_transformer_unbias(plan, params=store, precision="bf16")


@pytest.mark.parametrize("gen", [dict(), dict(full_length=True), dict(seq_lens=[16, 30, 10])])
def test_seq_encode_bf16_matches_oracle(gen):
    plan, model, host, dev, P, O = _setup("dmt_d64.conf", 37, seed=31, **gen)
    tc = _bf16_model(plan, model.params)
    want = O.trans_core(plan, P, O.generate_data(plan, P, host), training=False)
    B = 37
    out = torch.zeros(B, len(plan.sequences) * plan.d_model, device="cuda")
    ref = torch.zeros_like(out)
    if "seq_lens" in gen:
        dev["__max_len__"] = {i: l for i, l in enumerate(gen["seq_lens"])}
    for s in range(len(plan.sequences)):
        tc.seq_encode(dev, s, out.data_ptr() + 4 * s * plan.d_model, out.stride(0), B)
        model.seq_encode(dev, s, ref.data_ptr() + 4 * s * plan.d_model, ref.stride(0), B)
    torch.cuda.synchronize()
    err = (out.double().cpu() - want).abs()
    assert err.max().item() < ATOL_BF16, "max abs err %.3e" % err.max().item()
    assert err.mean().item() < MEAN_BF16, "mean abs err %.3e" % err.mean().item()
    # and against the fp32 CUDA path (same inputs, same weights)
    assert (out - ref).abs().max().item() < ATOL_BF16


def test_seq_encode_bf16_edge_lengths_and_batch_tail():
    """length-1 'unknow' sequences, full length, a batch that does not fill the last tile."""
    from cikm2020_dmt_b200.data import SparseIds, batch_to
    plan, model, host, dev, P, O = _setup("dmt_d64.conf", 5, seed=35)
    lens = [1, 50, 1, 17, 50]
    g = torch.Generator().manual_seed(12)
    off = torch.zeros(6, dtype=torch.int32)
    off[1:] = torch.cumsum(torch.tensor(lens), 0)
    seq = plan.sequences[0]
    for f, uf in enumerate(seq.user_features):
        V = plan.tables[seq.tables[f]].rows
        vals = torch.randint(1, V, (int(off[-1]),), generator=g, dtype=torch.int32)
        vals[0] = 0
        host[uf] = SparseIds(vals, off)
    dev = batch_to(host, "cuda")
    tc = _bf16_model(plan, model.params)
    want = O.trans_core(plan, P, O.generate_data(plan, P, host), training=False)
    out = torch.zeros(5, len(plan.sequences) * plan.d_model, device="cuda")
    for s in range(len(plan.sequences)):
        tc.seq_encode(dev, s, out.data_ptr() + 4 * s * plan.d_model, out.stride(0), 5)
    torch.cuda.synchronize()
    err = (out.double().cpu() - want).abs()
    assert err.max().item() < ATOL_BF16 and err.mean().item() < MEAN_BF16, (err.max().item(), err.mean().item())


@pytest.mark.parametrize("maxlen", [40, 50, 55])
def test_seq_encode_bf16_many_tiles_per_group(maxlen):
    """2400 samples = 1200 (64-row slots) / 300 (16-row slots) tiles: every CTA group of the persistent kernel loops
    over several tiles (software-pipelined gather, deferred decoder-context read-out, mbarrier phase tracking), and
    the three softmax key windows (maxlen 40 -> 48 keys, 50 -> 56, 55 -> 64) are compiled."""
    from cikm2020_dmt_b200.params import ParamStore
    from cikm2020_dmt_b200.data import synthetic_batch, batch_to
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    from conftest import SMALL_ROWS
    from oracle import dmt_oracle as O
    from cikm2020_dmt_b200 import keys as K
    conf, plan = make_plan("dmt_d64.conf", overrides={(K.MODEL, "transformer_maxlen_k"): str(maxlen)})
    assert plan.maxlen_k == maxlen
    store = ParamStore(plan, device="cuda", seed=5).randomize_(6)
    B = 2400
    host = synthetic_batch(plan, B, seed=77, table_rows=SMALL_ROWS)
    dev = batch_to(host, "cuda")
    tc = mmoe_transformer_unbias(plan, params=store, precision="bf16")
    P = O.params_from_store(store, torch.float32)
    want = O.trans_core(plan, P, O.generate_data(plan, P, host), training=False)
    out = torch.zeros(B, len(plan.sequences) * plan.d_model, device="cuda")
    for s in range(len(plan.sequences)):
        tc.seq_encode(dev, s, out.data_ptr() + 4 * s * plan.d_model, out.stride(0), B)
    torch.cuda.synchronize()
    err = (out.double().cpu() - want.double()).abs()
    assert err.max().item() < ATOL_BF16 and err.mean().item() < MEAN_BF16, (err.max().item(), err.mean().item())
    # run-to-run determinism of the whole path (no atomics anywhere)
    out2 = torch.zeros_like(out)
    for s in range(len(plan.sequences)):
        tc.seq_encode(dev, s, out2.data_ptr() + 4 * s * plan.d_model, out2.stride(0), B)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)


def test_inference_bf16_matches_oracle():
    plan, model, host, dev, P, O = _setup("dmt_d64.conf", 300, seed=41)
    tc = _bf16_model(plan, model.params)
    (yr, yb) = tc.inference(dev, is_train=False)
    (wr, wb) = O.inference(plan, P, host, is_train=False)
    torch.cuda.synchronize()
    _close(yr[0], wr[0], atol=5e-2, rtol=2e-2)
    _close(yr[1], wr[1], atol=5e-2, rtol=2e-2)
    _close(yb, wb, atol=1e-5)
    loss = tc.loss((yr, yb), dev["mask"])
    want = O.logit_loss_unbias(plan, (wr, wb), host["mask"])
    assert abs(loss.item() - want.item()) <= 1e-2 * max(1.0, abs(want.item()))


def test_bf16_rejects_unbuilt_shapes():
    plan, model, host, dev, P, O = _setup("dmt.conf", 4, seed=1)
    from cikm2020_dmt_b200 import abi
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    tc = mmoe_transformer_unbias(plan, params=model.params, precision="bf16")
    out = torch.zeros(4, plan.d_model, device="cuda")
    with pytest.raises(abi.DmtError) as ei:
        tc.seq_encode(dev, 0, out.data_ptr(), out.stride(0), 4)
    assert ei.value.code == -2


@pytest.mark.parametrize("batch", [150, 1024])
def test_mmoe_bf16_matches_oracle(batch):
    """TMA + tcgen05 expert GEMMs (bf16 operands, fp32 accumulate): logits atol 5e-2 / rtol 2e-2."""
    plan, model, host, dev, P, O = _setup("dmt_d64.conf", 8, seed=9)
    tc = _bf16_model(plan, model.params)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(batch, plan.mmoe_in, generator=g) * 0.3
    tasks = O.expert_gate(plan, P, x.double())
    want = torch.stack([O.build_tower(plan, P, t, O.TASK_NAMES[i]).squeeze(1) for i, t in enumerate(tasks)])
    xd = x.cuda()
    logits = torch.zeros(2, batch, device="cuda")
    tc.mmoe(xd, batch, logits)
    ref = torch.zeros(2, batch, device="cuda")
    model.mmoe(xd, batch, ref)
    torch.cuda.synchronize()
    _close(logits, want, atol=5e-2, rtol=2e-2)
    assert (logits - ref).abs().mean().item() < 1e-2


def test_prefetch_pointer_staging_matches_views():
    """`prefetch(packed, views=False)` (raw DevArray descriptors) must give bit-identical scores to the torch-view
    staging of the same packed batch, on both kernel paths."""
    from cikm2020_dmt_b200.data import PackedBatch
    plan, model, host, dev, P, O = _setup("dmt_d64.conf", 260, seed=61)
    tc = _bf16_model(plan, model.params)
    packed = PackedBatch(host)
    for m in (model, tc):
        (a_r, a_b) = m.inference(m.prefetch(packed, views=True), is_train=False)
        a = [a_r[0].clone(), a_r[1].clone(), a_b.clone()]
        (b_r, b_b) = m.inference(m.prefetch(packed, views=False), is_train=False)
        torch.cuda.synchronize()
        assert torch.equal(a[0], b_r[0]) and torch.equal(a[1], b_r[1]) and torch.equal(a[2], b_b)


def test_compact_packed_batch_widen_and_bf16_features():
    """Compact host->device format (data.py / stage_batch.cu): uint16 id arrays are restored bit-exactly by
    dmt_widen_u16, the bf16 feature block lands in the MMoE input as the RNE-rounded fp32 values, and the bf16
    path's scores from a compact batch stay within the bf16 tolerance of the oracle (and within 2e-2 of the scores
    from the wide batch: only the gate inputs see bf16-rounded instead of fp32 features)."""
    from cikm2020_dmt_b200.data import PackedBatch, SparseIds
    plan, model, host, dev, P, O = _setup("dmt_d64.conf", 333, seed=71)       # ragged: 333 % 8 != 0 tails
    tc = _bf16_model(plan, model.params)
    keys = set(plan.all_id_features()) | {"features"}
    packed = PackedBatch(host, compact=True, keys=keys)
    assert packed.narrow and packed.nbytes < PackedBatch(host).nbytes
    staged = packed.to("cuda")
    torch.cuda.synchronize()
    for k in keys:
        v = host[k]
        if isinstance(v, SparseIds):
            assert staged[k].values.dtype == torch.int32
            assert torch.equal(staged[k].values.cpu(), v.values), k
            assert torch.equal(staged[k].offsets.cpu(), v.offsets), k
    assert staged["features"].dtype == torch.bfloat16
    assert torch.equal(staged["features"].cpu(), host["features"].to(torch.bfloat16))
    (wr, wb) = O.inference(plan, P, host, is_train=False)
    for views in (True, False):
        (yr, yb) = tc.inference(tc.prefetch(packed, views=views), is_train=False)
        torch.cuda.synchronize()
        _close(yr[0], wr[0], atol=5e-2, rtol=2e-2)
        _close(yr[1], wr[1], atol=5e-2, rtol=2e-2)
        _close(yb, wb, atol=1e-5)
        x = tc._last["x"][:, :plan.feature_dim].cpu()
        assert torch.equal(x, host["features"].to(torch.bfloat16).float())
    (zr, zb) = tc.inference(PackedBatch(host), is_train=False)
    torch.cuda.synchronize()
    assert (zr[0] - yr[0]).abs().max().item() < 2e-2 and (zr[1] - yr[1]).abs().max().item() < 2e-2
    # the fp32 path refuses bf16 features instead of silently widening them
    with pytest.raises(ValueError):
        model.inference(packed, is_train=False)


@pytest.mark.parametrize("conf_file,batch", [("dmt.conf", 300), ("dmt_d64.conf", 300)])
def test_inference_tf32_pipeline_matches_oracle(conf_file, batch):
    """precision='tf32': the row-batched pipeline (dense projections / feed-forward / MMoE experts on the TMA-fed
    tcgen05 kind::tf32 engine, attention and LayerNorm fp32 on CUDA cores) -- the tensor-core path for ANY shape
    with d_model, d_ff multiples of 16, in particular the reference's own dmt.conf (d_model 80, 4 heads of 20,
    d_ff 320, MMoE input 1199).  tf32 keeps 10 operand mantissa bits: interest vectors atol 2e-2 / mean 3e-3,
    logits atol 2e-2 + rtol 1e-2, loss rtol 5e-3 against the fp64 oracle."""
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    plan, model, host, dev, P, O = _setup(conf_file, batch, seed=83)
    tf = mmoe_transformer_unbias(plan, params=model.params, precision="tf32")
    want = O.trans_core(plan, P, O.generate_data(plan, P, host), training=False)
    out = torch.zeros(batch, len(plan.sequences) * plan.d_model, device="cuda")
    for s in range(len(plan.sequences)):
        tf.seq_encode(dev, s, out.data_ptr() + 4 * s * plan.d_model, out.stride(0), batch)
    torch.cuda.synchronize()
    err = (out.double().cpu() - want.double()).abs()
    assert err.max().item() < 2e-2 and err.mean().item() < 3e-3, (err.max().item(), err.mean().item())
    (yr, yb) = tf.inference(dev, is_train=False)
    (wr, wb) = O.inference(plan, P, host, is_train=False)
    torch.cuda.synchronize()
    _close(yr[0], wr[0], atol=2e-2, rtol=1e-2)
    _close(yr[1], wr[1], atol=2e-2, rtol=1e-2)
    _close(yb, wb, atol=1e-5)
    loss = tf.loss((yr, yb), dev["mask"])
    want_loss = O.logit_loss_unbias(plan, (wr, wb), host["mask"])
    assert abs(loss.item() - want_loss.item()) <= 5e-3 * max(1.0, abs(want_loss.item()))
    # is_predict returns the logit pair only, from the same kernels
    (pc, po) = tf.inference(dev, is_train=False, is_predict=True)
    torch.cuda.synchronize()
    assert torch.equal(pc, yr[0]) and torch.equal(po, yr[1])


def test_bench_size_bf16_and_tf32_against_fp32_path_and_oracle():
    """The bench's own size and tables (BASELINE config 2: B = 4096, Sku 5,000,000 x 32 and the other full
    vocabularies): every sample's logits from the bf16 fused tcgen05 kernels and from the tf32 pipeline against
    the fp32 CUDA path on the same device batch (bf16: atol 5e-2 + rtol 2e-2; tf32: atol 2e-2 + rtol 1e-2), and
    a strided subsample of all three against the fp64 oracle."""
    from cikm2020_dmt_b200.conf import Conf
    from cikm2020_dmt_b200.data import SparseIds, batch_to, synthetic_batch
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    from cikm2020_dmt_b200.params import ParamStore
    from cikm2020_dmt_b200.plan import build_plan
    from conftest import CONF_DIR
    from oracle import dmt_oracle as O
    plan = build_plan(Conf(CONF_DIR, "dmt_d64.conf"))
    assert plan.tables["Sku"].rows == 5000000
    store = ParamStore(plan, device="cuda", seed=3).randomize_(4)
    host = synthetic_batch(plan, 4096, seed=17)
    dev = batch_to(host, "cuda")
    outs = {}
    for prec in ("f32", "bf16", "tf32"):
        m = mmoe_transformer_unbias(plan, params=store, precision=prec)
        (yr, yb) = m.inference(dev, is_train=False)
        torch.cuda.synchronize()
        outs[prec] = (yr[0].clone(), yr[1].clone(), yb.clone())
        del m
    for prec, atol, rtol in (("bf16", 5e-2, 2e-2), ("tf32", 2e-2, 1e-2)):
        for t in range(2):
            _close(outs[prec][t], outs["f32"][t], atol=atol, rtol=rtol)
        _close(outs[prec][2], outs["f32"][2], atol=1e-5)
    idx = list(range(0, 4096, 173))
    sub = {}
    for k, v in host.items():
        if isinstance(v, SparseIds):
            sub[k] = SparseIds.from_lists([v.values[int(v.offsets[b]):int(v.offsets[b + 1])].tolist() for b in idx])
        else:
            sub[k] = v[idx]
    (wr, wb) = O.inference(plan, O.params_from_store(store, torch.float32), sub, is_train=False, lean=True)
    for prec, atol, rtol in (("f32", 2e-4, 2e-4), ("bf16", 5e-2, 2e-2), ("tf32", 2e-2, 1e-2)):
        _close(outs[prec][0][idx], wr[0], atol=atol, rtol=rtol)
        _close(outs[prec][1][idx], wr[1], atol=atol, rtol=rtol)
