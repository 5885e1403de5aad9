"""K9/K10: segmented sparse-gradient scatter-add + TF-1 Adam, against the oracle's dense TFAdam.

Integer part (which row every lookup hits, zero-pad offset, skipped index 0) is exact; the fp32 update is
compared with the fp64 oracle at atol 1e-6 / rtol 1e-5 after several steps."""
import math

import pytest
import torch

from conftest import make_plan

pytestmark = pytest.mark.gpu


def _model(conf="dmt_d64.conf"):
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    from cikm2020_dmt_b200.params import ParamStore
    conf, plan = make_plan(conf)
    store = ParamStore(plan, device="cuda", seed=5).randomize_(6)
    return plan, mmoe_transformer_unbias(plan, params=store)


def test_adam_dense_matches_tf_adam_over_steps():
    from cikm2020_dmt_b200.optim import TFAdam
    from oracle import dmt_oracle as O
    plan, model = _model()
    opt = TFAdam(model, 1e-3)
    ref_p = {"w": model.params.dense.detach().double().cpu().clone()}
    ref = O.TFAdam(ref_p, lr=1e-3)
    g = torch.Generator().manual_seed(1)
    for step in range(4):
        grad = torch.randn(model.params.dense.numel(), generator=g) * (0.1 if step else 1.0)
        grad[::7] = 0.0
        opt.begin_step()
        opt.step_dense(grad.cuda(), grad_scale=0.5)
        ref.step({"w": grad.double() * 0.5})
    torch.cuda.synchronize()
    got = model.params.dense.double().cpu()
    assert torch.allclose(got, ref_p["w"], atol=1e-6, rtol=1e-5), (got - ref_p["w"]).abs().max()
    assert torch.allclose(opt.m_dense.double().cpu(), ref.m["w"], atol=1e-7, rtol=1e-5)
    assert torch.allclose(opt.v_dense.double().cpu(), ref.v["w"], atol=1e-9, rtol=1e-5)


@pytest.mark.parametrize("table_name,dim", [("Sku", 32), ("Cid2", 8)])
def test_sparse_scatter_adam_matches_dense_tf_adam(table_name, dim):
    """Every lookup kind that feeds a table in one step: sequence tokens (zero-pad, row = id-1, id 0 skipped),
    target items (one gradient row per sample), pooled mean with and without `Wts` (row = id, scaled w/sum_w)."""
    from cikm2020_dmt_b200.optim import LookupGrad, TFAdam
    from oracle import dmt_oracle as O
    plan, model = _model()
    name = plan.tables[table_name].scope
    table = model.params.tables[name]
    V = table.shape[0]
    assert table.shape[1] == dim
    opt = TFAdam(model, 1e-3)
    ref_p = {"t": table.detach().double().cpu().clone()}
    ref = O.TFAdam(ref_p, lr=1e-3)
    g = torch.Generator().manual_seed(dim)
    B = 40
    for step in range(3):
        lens = torch.randint(1, 9, (B,), generator=g)
        off = torch.zeros(B + 1, dtype=torch.int32)
        off[1:] = torch.cumsum(lens, 0)
        n = int(off[-1])
        ids_tok = torch.randint(0, V + 1, (n,), generator=g, dtype=torch.int32)     # includes 0 and V (-> row V-1)
        ids_tok[:5] = torch.tensor([0, 1, V, 7, 7])
        g_tok = torch.randn(n, 64, generator=g)                                      # token gradients, slice at col 8
        ids_item = torch.randint(0, V + 1, (B,), generator=g, dtype=torch.int32)
        g_item = torch.randn(B, 64, generator=g)
        ids_pool = torch.randint(0, V, (n,), generator=g, dtype=torch.int32)
        ids_pool[-3:] = 7
        wts = torch.rand(n, generator=g) + 0.5
        g_pool = torch.randn(B, 300, generator=g)
        col_tok, col_pool_w, col_pool = 8, 100, 200
        one = torch.arange(B + 1, dtype=torch.int32)
        srcs = [
            LookupGrad(ids_tok.cuda(), g_tok.cuda(), col_tok, -1),
            LookupGrad(ids_item.cuda(), g_item.cuda(), col_tok, -1, offsets=one.cuda()),
            LookupGrad(ids_pool.cuda(), g_pool.cuda(), col_pool_w, 0, offsets=off.cuda(), weights=wts.cuda(), mean=True),
            LookupGrad(ids_pool.cuda(), g_pool.cuda(), col_pool, 0, offsets=off.cuda(), mean=True),
        ]
        opt.begin_step()
        opt.step_table(name, srcs, grad_scale=0.25)
        # oracle: densify, then dense TF-Adam (run_dnn.py:63-72 + AdamOptimizer)
        dense = torch.zeros(V, dim, dtype=torch.float64)
        seg = torch.repeat_interleave(torch.arange(B), lens)
        ok = ids_tok > 0
        dense.index_add_(0, (ids_tok[ok] - 1).long(), g_tok[ok][:, col_tok:col_tok + dim].double())
        ok = ids_item > 0
        dense.index_add_(0, (ids_item[ok] - 1).long(), g_item[ok][:, col_tok:col_tok + dim].double())
        wsum = torch.zeros(B, dtype=torch.float64).index_add_(0, seg, wts.double())
        dense.index_add_(0, ids_pool.long(), g_pool[seg][:, col_pool_w:col_pool_w + dim].double()
                         * (wts.double() / wsum[seg])[:, None])
        dense.index_add_(0, ids_pool.long(), g_pool[seg][:, col_pool:col_pool + dim].double()
                         / lens.double()[seg][:, None])
        ref.step({"t": dense * 0.25})
    torch.cuda.synchronize()
    got = table.double().cpu()
    assert torch.allclose(got, ref_p["t"], atol=1e-6, rtol=1e-5), (got - ref_p["t"]).abs().max()
    assert torch.allclose(opt.m_tab[name].double().cpu(), ref.m["t"], atol=1e-6, rtol=1e-4)
    assert torch.allclose(opt.v_tab[name].double().cpu(), ref.v["t"], atol=1e-8, rtol=1e-4)
    assert int(opt.touched[name].sum()) == 0          # marks are cleared by the dense pass
    # rows that never received a gradient still moved after step 1 (dense semantics) ...
    # ... unless their moments are exactly zero: a never-touched row is bit-identical to its initial value
    untouched = (ref.v["t"].abs().sum(1) == 0)
    if not untouched.any():      # tiny tables: every row was hit
        return
    init = O.params_from_store(__import__("cikm2020_dmt_b200.params", fromlist=["ParamStore"]).ParamStore(
        plan, device="cpu", seed=5).randomize_(6))[name]
    assert torch.equal(got[untouched].float(), init[untouched].float())


def test_scatter_is_deterministic():
    from cikm2020_dmt_b200.optim import LookupGrad, TFAdam
    plan, model = _model()
    name = plan.tables["Brand"].scope
    V, dim = model.params.tables[name].shape
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 50, (20000,), generator=g, dtype=torch.int32).cuda()     # heavy duplication
    grad = torch.randn(20000, dim, generator=g).cuda()
    outs = []
    for rep in range(2):
        plan2, m2 = _model()
        opt = TFAdam(m2, 1e-3)
        opt.begin_step()
        opt.step_table(name, [LookupGrad(ids, grad, 0, -1)])
        torch.cuda.synchronize()
        outs.append(m2.params.tables[name].clone())
    assert torch.equal(outs[0], outs[1])
