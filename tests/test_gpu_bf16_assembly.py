"""GPU parity of the bf16 MMoE-input assembly (round 2): the producers of the MMoE input write bf16 columns in place
(dmt_stage_dense_features_bf16, dmt_pool_mean_fwd_bf16 through the grouped pooled-lookup kernel, the sequence tails
with DMT_SEQ_OUT_BF16) and dmt_mmoe_fwd_bf16in consumes them -- base.py:93-124 and
mmoe_transformer_unbias.py:63-126,218-232 for the bf16 tensor-core path.

Bars: byte movement and type conversion bit-exact; pooled means fp32 within 1e-6 of the fp64 oracle and the bf16
output == round-to-nearest of the fp32 output; logits within the bf16 tolerance (atol 5e-2 / rtol 2e-2).
"""
import ctypes as C

import pytest
import torch

from conftest import make_plan, SMALL_ROWS

pytestmark = pytest.mark.gpu


def _setup(conf_file, batch, seed=0, precision="f32", **gen):
    from cikm2020_dmt_b200.params import ParamStore
    from cikm2020_dmt_b200.data import synthetic_batch, batch_to
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    from oracle import dmt_oracle as O
    conf, plan = make_plan(conf_file)
    store = ParamStore(plan, device="cuda", seed=seed + 1).randomize_(seed + 2)
    model = mmoe_transformer_unbias(plan, params=store, precision=precision)
    host = synthetic_batch(plan, batch, seed=seed + 3, table_rows=SMALL_ROWS, **gen)
    P = O.params_from_store(store)
    return plan, model, host, batch_to(host, "cuda"), P, O


@pytest.mark.parametrize("batch,dim", [(1, 615), (37, 615), (4096, 615), (5, 7), (3, 8), (1, 1)])
def test_stage_dense_features_bf16_bit_exact(batch, dim):
    from cikm2020_dmt_b200 import abi
    lib = abi.load()
    g = torch.Generator().manual_seed(batch * 131 + dim)
    src = torch.randn(batch, dim, generator=g).cuda()
    ld = (dim + 7) // 8 * 8 + 8
    st = torch.cuda.current_stream().cuda_stream
    for is_bf16 in (0, 1):
        inp = src.to(torch.bfloat16).contiguous() if is_bf16 else src
        out = torch.full((batch, ld), -7.0, dtype=torch.bfloat16, device="cuda")
        abi.check(lib.dmt_stage_dense_features_bf16(inp.data_ptr(), is_bf16, batch, dim, out.data_ptr(), ld, st))
        torch.cuda.synchronize()
        assert torch.equal(out[:, :dim], src.to(torch.bfloat16))
        assert bool((out[:, dim:] == -7.0).all())            # nothing outside the feature columns is touched


@pytest.mark.parametrize("conf_file", ["dmt_d64.conf", "dmt.conf"])
def test_grouped_pool_mean_fp32_and_bf16(conf_file):
    """Features of one behaviour sequence share their offsets tensor (batch_to keeps shared storage shared) and walk
    the tokens together; the same features with separately stored offsets run as single-feature groups: bit-identical
    (the summation order of a feature depends on its row width only)."""
    from cikm2020_dmt_b200.data import batch_to, SparseIds
    B = 29
    plan, model, host, dev, P, O = _setup(conf_file, B, seed=3)
    g = torch.Generator().manual_seed(5)
    for p in plan.pooled[5:9]:     # `<feature>Wts` on a few features
        host[p.feature + "Wts"] = torch.rand(host[p.feature].values.numel(), generator=g) + 0.25
    dev = batch_to(host, "cuda")
    shared = {dev[p.feature].offsets.data_ptr() for p in plan.pooled}
    assert len(shared) < len(plan.pooled)                     # the sequences' features do share offsets
    want = O.embedding_combiner(plan, P, host)[:, plan.feature_dim:]
    x = torch.zeros(B, plan.interest_col, device="cuda")
    model.pool_mean(dev, plan.pooled, False, x, B)
    torch.cuda.synchronize()
    got = x[:, plan.feature_dim:]
    err = (got.double().cpu() - want).abs()
    assert bool((err <= 1e-6 + 1e-6 * want.abs()).all()), err.max().item()
    # separately stored offsets -> one group per feature
    solo = dict(dev)
    for p in plan.pooled:
        v = dev[p.feature]
        solo[p.feature] = SparseIds(v.values, v.offsets.clone(), v.weights)
    x1 = torch.zeros_like(x)
    model.pool_mean(solo, plan.pooled, False, x1, B)
    x2 = torch.zeros_like(x)
    model.pool_mean(dev, plan.pooled, False, x2, B)
    torch.cuda.synchronize()
    assert torch.equal(x, x2) and torch.equal(x, x1)
    # bf16 output == the fp32 output rounded to nearest
    ld = (plan.mmoe_in + 7) // 8 * 8
    xb = torch.zeros(B, ld, dtype=torch.bfloat16, device="cuda")
    model.pool_mean(dev, plan.pooled, False, xb, B)
    torch.cuda.synchronize()
    assert torch.equal(xb[:, plan.feature_dim:plan.interest_col], got.to(torch.bfloat16))
    assert bool((xb[:, :plan.feature_dim] == 0).all()) and bool((xb[:, plan.interest_col:] == 0).all())


def test_grouped_pool_mean_edge_rows():
    """empty rows (absent from the SparseTensor -> 0), ids outside the table (-> zero row, weight still counted),
    more than 8 tokens (several rounds) and a ragged last round."""
    from cikm2020_dmt_b200 import abi
    from cikm2020_dmt_b200.data import SparseIds
    lib = abi.load()
    rows, dim = 50, 8
    g = torch.Generator().manual_seed(3)
    table = torch.randn(rows, dim, generator=g)
    lists = [[], [3], [1, 2, 3, 4, 5, 6, 7, 8], [9] * 9, list(range(23)), [49, 50, -1, 0], []]
    sp = SparseIds.from_lists(lists)
    B = len(lists)
    w = torch.rand(sp.values.numel(), generator=g) + 0.5
    want = torch.zeros(B, dim, dtype=torch.float64)
    off = sp.offsets.tolist()
    for b in range(B):
        num, den = torch.zeros(dim, dtype=torch.float64), 0.0
        for t in range(off[b], off[b + 1]):
            i = int(sp.values[t])
            if 0 <= i < rows:
                num += float(w[t]) * table[i].double()
            den += float(w[t])
        if off[b + 1] > off[b]:
            want[b] = num / den
    tab_d, ids_d, off_d, w_d = table.cuda(), sp.values.cuda(), sp.offsets.cuda(), w.cuda()
    feat = (abi.PoolFeat * 1)()
    feat[0].table, feat[0].rows, feat[0].dim, feat[0].out_col = tab_d.data_ptr(), rows, dim, 4
    feat[0].ids, feat[0].offsets, feat[0].weights = ids_d.data_ptr(), off_d.data_ptr(), w_d.data_ptr()
    out = torch.full((B, 16), 9.0, device="cuda")
    abi.check(lib.dmt_pool_mean_fwd(B, 1, feat, out.data_ptr(), 16, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert (out[:, 4:12].double().cpu() - want).abs().max().item() < 1e-6
    assert bool((out[:, :4] == 9.0).all()) and bool((out[:, 12:] == 9.0).all())


@pytest.mark.parametrize("batch", [150, 1024])
def test_mmoe_bf16_input_matches_fp32_input_entry(batch):
    """dmt_mmoe_fwd_bf16in on xb vs dmt_mmoe_fwd(bf16) on the same values as fp32: identical expert GEMM operands;
    the gate logits come out of the layer-0 GEMM (bf16 gate kernels, fp32 accumulate) instead of an fp32 pass."""
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    plan, model, host, dev, P, O = _setup("dmt_d64.conf", 8, seed=9)
    tc = mmoe_transformer_unbias(plan, params=model.params, precision="bf16")
    g = torch.Generator().manual_seed(2)
    ld = (plan.mmoe_in + 7) // 8 * 8
    xb = torch.zeros(batch, ld, dtype=torch.bfloat16)
    xb[:, :plan.mmoe_in] = (torch.randn(batch, plan.mmoe_in, generator=g) * 0.3).to(torch.bfloat16)
    xb[:, plan.mmoe_in:] = float("nan")            # the padding columns are never read (TMA zero-fills beyond in_dim)
    x32 = xb[:, :plan.mmoe_in].float()
    tasks = O.expert_gate(plan, P, x32.double())
    want = torch.stack([O.build_tower(plan, P, t, O.TASK_NAMES[i]).squeeze(1) for i, t in enumerate(tasks)])
    a = torch.zeros(2, batch, device="cuda")
    b = torch.zeros(2, batch, device="cuda")
    tc.mmoe(xb.cuda(), batch, a)
    tc.mmoe(x32.cuda().contiguous(), batch, b)
    torch.cuda.synchronize()
    assert (a - b).abs().max().item() < 1e-2, (a - b).abs().max().item()
    err = (a.double().cpu() - want).abs()
    assert bool((err <= 5e-2 + 2e-2 * want.abs()).all()), err.max().item()


@pytest.mark.parametrize("compact", [False, True])
def test_inference_bf16_assembly_matches_oracle_and_fp32_assembly(compact):
    from cikm2020_dmt_b200.data import PackedBatch
    B = 300
    plan, tc, host, dev, P, O = _setup("dmt_d64.conf", B, seed=41, precision="bf16")
    assert tc.x_bf16 and tc.seq_multi
    if compact:                                   # bf16 features + uint16 ids from the packed host buffer
        pk = PackedBatch(host, compact=True)
        dev = pk.to("cuda")
        dev["__max_len__"] = pk.max_len(plan)
    (yr, yb) = tc.inference(dev, is_train=False)
    torch.cuda.synchronize()
    (wr, wb) = O.inference(plan, P, host, is_train=False)
    for t in range(2):
        err = (yr[t].double().cpu() - wr[t]).abs()
        assert bool((err <= 5e-2 + 2e-2 * wr[t].abs()).all()), err.max().item()
    assert (yb.double().cpu() - wb).abs().max().item() < 1e-5
    got = [y.clone() for y in yr]
    tc.x_bf16 = False
    (yr1, _) = tc.inference(dev, is_train=False)
    torch.cuda.synchronize()
    for t in range(2):
        assert (got[t] - yr1[t]).abs().max().item() < 3e-2


@pytest.mark.parametrize("B", [1, 2, 129])
def test_bf16_inference_small_and_ragged_batches(B):
    """One sample, two samples, one sample past a 128-row tile: every kernel of the bf16 route has a partial last tile."""
    plan, tc, host, dev, P, O = _setup("dmt_d64.conf", B, seed=50 + B, precision="bf16")
    (yr, yb) = tc.inference(dev, is_train=False)
    torch.cuda.synchronize()
    (wr, wb) = O.inference(plan, P, host, is_train=False)
    for t in range(2):
        err = (yr[t].double().cpu() - wr[t]).abs()
        assert bool((err <= 5e-2 + 2e-2 * wr[t].abs()).all()), err.max().item()
    assert (yb.double().cpu() - wb).abs().max().item() < 1e-5


@pytest.mark.parametrize("is_predict", [False, True])
def test_native_forward_driver_equals_the_per_operator_path(is_predict):
    """dmt_forward_bf16 issues the same entry points as the Python path: bit-identical scores -- for a resident dict
    batch (pointer table from the tensors) and for a prefetched compact PackedBatch (base + feature_offsets)."""
    from cikm2020_dmt_b200.data import PackedBatch
    B = 300
    plan, tc, host, dev, P, O = _setup("dmt_d64.conf", B, seed=43, precision="bf16")
    assert tc.fwd_native

    def run(inputs):
        out = tc.inference(inputs, is_train=False, is_predict=is_predict)
        torch.cuda.synchronize()
        return tc.last_scores[:plan.num_tasks + (0 if is_predict else 1)].clone()

    native = run(dev)
    launches0 = tc.launches
    assert torch.equal(run(dev), native)                 # (weights prepared by the first call)
    # bias pool + tower | length classes + tile kernel + tails | dense + pooled | layer-0 GEMM + fused tail + mixture
    assert tc.launches - launches0 == (8 if is_predict else 10)
    keys = set(plan.all_id_features()) | {"features"}
    pk = PackedBatch(host, compact=True, keys=keys)
    staged_native = run(tc.prefetch(pk, views=False))
    tc.fwd_native = False
    python = run(dev)
    staged_python = run(tc.prefetch(pk, views=False))
    assert torch.equal(native, python)
    assert torch.equal(staged_native, staged_python)
    # compact batch: bf16 features / widened uint16 ids -- same ids, features rounded once either way
    assert (native - staged_native).abs().max().item() < 3e-2
    # a second batch size gets its own descriptor; a parameter change rebuilds it
    tc.fwd_native = True
    plan2, _, host2, dev2, _, _ = _setup("dmt_d64.conf", 64, seed=44, precision="bf16")
    a = run(dev2)
    tc.fwd_native = False
    b = run(dev2)
    assert torch.equal(a, b)
    tc.fwd_native = True
    tc.params.dense.mul_(1.01)
    tc.invalidate_prepared()
    c = run(dev)
    tc.fwd_native = False
    d = run(dev)
    assert torch.equal(c, d) and not torch.equal(c, native)


@pytest.mark.parametrize("nbytes", [1, 2, 3])
@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 4099, 70001])
def test_widen_ids_bit_exact(nbytes, n):
    """dmt_widen_ids: 1- / 2- / 3-byte little-endian ids -> int32, vector body + bytewise tail."""
    import numpy as np
    from cikm2020_dmt_b200 import abi
    lib = abi.load()
    rng = np.random.default_rng(n * 7 + nbytes)
    ids = rng.integers(0, 1 << (8 * nbytes), size=n, dtype=np.int64).astype(np.int32)
    if n:
        ids[0], ids[-1] = (1 << (8 * nbytes)) - 1, 0
    raw = np.ascontiguousarray(ids.astype("<u4")).view(np.uint8).reshape(-1, 4)[:, :nbytes].reshape(-1)
    src = torch.zeros(max(raw.size, 1) + 64, dtype=torch.uint8)
    src[:raw.size] = torch.from_numpy(np.ascontiguousarray(raw))
    src = src.cuda()
    dst = torch.full((n + 8,), -5, dtype=torch.int32, device="cuda")
    desc = (abi.WidenIdsDesc * 1)()
    desc[0].src, desc[0].dst, desc[0].n, desc[0].bytes = src.data_ptr(), dst.data_ptr(), n, nbytes
    abi.check(lib.dmt_widen_ids(1, desc, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(dst[:n].cpu(), torch.from_numpy(ids))
    assert bool((dst[n:] == -5).all())


def test_compact_batch_with_24_bit_ids_matches_wide_batch():
    """Sku / Brand ids beyond 16 bits travel as 3 bytes: the staged batch equals the wide one id for id."""
    from cikm2020_dmt_b200.data import PackedBatch, SparseIds, synthetic_batch
    rows = {"Sku": 5000000, "Brand": 190000, "Shopid": 230000, "Cid3": 12000, "Cid2": 500}
    conf, plan = make_plan("dmt_d64.conf", rows=rows)
    host = synthetic_batch(plan, 129, seed=5, table_rows=rows)
    keys = set(plan.all_id_features()) | {"features"}
    packed = PackedBatch(host, compact=True, keys=keys)
    widths = sorted({v[2] for v in packed.narrow.values()})
    assert widths == [1, 2, 3]
    staged = packed.to("cuda")
    torch.cuda.synchronize()
    for k in keys:
        v = host[k]
        if isinstance(v, SparseIds):
            assert torch.equal(staged[k].values.cpu(), v.values), k
            assert torch.equal(staged[k].offsets.cpu(), v.offsets), k


def test_new_entries_reject_bad_arguments():
    from cikm2020_dmt_b200 import abi
    lib = abi.load()
    assert lib.dmt_stage_dense_features_bf16(None, 0, 4, 8, None, 8, None) == -1
    assert lib.dmt_pool_mean_fwd_bf16(4, 1, None, None, 8, None) == -1
    cfg = abi.MmoeCfg()
    assert lib.dmt_mmoe_fwd_bf16in(C.byref(cfg), None, None, 8, None, None, 0, None, None) == -1
    assert lib.dmt_forward_bf16(None, 0, None, None, 0, None, None) == -1
    bad = (abi.WidenIdsDesc * 1)()
    bad[0].src, bad[0].dst, bad[0].n, bad[0].bytes = 16, 16, 4, 4
    assert lib.dmt_widen_ids(1, bad, None) == -1
    desc = abi.FwdDesc()
    assert lib.dmt_forward_bf16(C.byref(desc), 0, None, None, 0, None, None) == -1
