"""The CPU oracle against the known-answer cases the reference's semantics give (SURVEY 4 test plan):
row offsets, padding inertness, shared feed-forward, loss and Adam closed forms."""
import math

import pytest
import torch

from conftest import SMALL_ROWS, make_plan


@pytest.fixture(scope="module")
def ctx():
    from cikm2020_dmt_b200.data import synthetic_batch
    from cikm2020_dmt_b200.params import ParamStore
    from oracle import dmt_oracle as O
    conf, plan = make_plan("dmt.conf")
    store = ParamStore(plan, seed=3).randomize_(4)
    batch = synthetic_batch(plan, 12, seed=5, table_rows=SMALL_ROWS)
    return plan, store, batch, O


def test_identity_table_row_offsets(ctx):
    """KAT (i): with table[r,:] = r the sequence/target path returns idx-1 (0 -> zeros), the pooled
    path returns idx (base.py:87-89 vs :115-116)."""
    plan, store, batch, O = ctx
    P = O.params_from_store(store)
    for t in plan.tables.values():
        P[t.scope] = torch.arange(t.rows, dtype=torch.float64)[:, None].repeat(1, t.dim)
    seq_data = O.generate_data(plan, P, batch)
    seq = plan.sequences[0]
    ids = O.sparse_to_dense(batch[seq.user_features[0]])
    got = seq_data[0][2][:, :, 0]
    want = torch.where(ids > 0, ids - 1, torch.zeros_like(ids)).double()
    assert torch.equal(got, want)
    item = batch[seq.item_features[0]].values.long()
    assert torch.equal(seq_data[0][3][:, 0], torch.where(item > 0, item - 1, torch.zeros_like(item)).double())
    feats = O.embedding_combiner(plan, P, batch)
    p0 = plan.pooled[0]                           # item sku, one id per sample -> mean == the row itself
    assert torch.equal(feats[:, p0.col], item.double())


def test_lean_lookup_equals_materialised_concat(ctx):
    plan, store, batch, O = ctx
    P = O.params_from_store(store)
    a = O.inference(plan, P, batch, lean=False)
    b = O.inference(plan, P, batch, lean=True)
    assert torch.equal(a[0][0], b[0][0]) and torch.equal(a[1], b[1])


def test_padded_rows_are_inert(ctx):
    """each sample's output equals its un-padded single-sample output."""
    from cikm2020_dmt_b200.data import SparseIds
    plan, store, batch, O = ctx
    P = O.params_from_store(store)
    (yr, yb) = O.inference(plan, P, batch)
    for b in (0, 5, 11):
        one = {}
        for k, v in batch.items():
            if isinstance(v, SparseIds):
                lo, hi = int(v.offsets[b]), int(v.offsets[b + 1])
                one[k] = SparseIds(v.values[lo:hi], torch.tensor([0, hi - lo], dtype=torch.int32))
            else:
                one[k] = v[b:b + 1]
        (r1, b1) = O.inference(plan, P, one)
        assert abs(r1[0].item() - yr[0][b].item()) < 1e-12
        assert abs(r1[1].item() - yr[1][b].item()) < 1e-12
        assert abs(b1.item() - yb[b].item()) < 1e-12


def test_shared_feed_forward_gets_both_gradients(ctx):
    """SURVEY 0.3 / KAT (v): encoder and decoder use the same FF variables."""
    plan, store, batch, O = ctx
    P = O.params_from_store(store)
    loss, grads, _ = O.loss_and_grads(plan, P, batch)
    seq = plan.sequences[0]
    name = seq.scope + "/num_blocks_0/positionwise_feedforward/dense/kernel"
    g_all = grads[name].clone()

    # decoder-only contribution: detach the encoder memory
    P2 = O.params_from_store(store, requires_grad=True)
    sd = O.generate_data(plan, P2, batch)
    mask, lens, seq_emb, tar = sd[0]
    mem = O.encode(plan, P2, seq.scope, seq_emb, lens, False).detach()
    dec = O.decode(plan, P2, seq.scope, tar[:, None, :], torch.ones(len(lens), dtype=torch.int64), mem, lens, False)
    dec.sum().backward()
    assert P2[name].grad.abs().sum() > 0          # the decoder alone reaches the shared kernel
    assert g_all.abs().sum() > 0
    # and there is exactly one FF kernel per block per sequence in the inventory
    assert sum(1 for k in P if k.startswith(seq.scope) and k.endswith("positionwise_feedforward/dense/kernel")) == 1


def test_gradients_at_padded_positions_and_zero_index(ctx):
    plan, store, batch, O = ctx
    P = O.params_from_store(store)
    loss, grads, _ = O.loss_and_grads(plan, P, batch)
    seq = plan.sequences[2]                                  # cart: max len 10 < maxlen_k 50
    pos = grads[seq.scope + "/positional_encoding_k_position_learn/embedding_position_learn"]
    T = int((batch[seq.user_features[0]].offsets[1:] - batch[seq.user_features[0]].offsets[:-1]).max())
    assert pos[T:].abs().sum() == 0 and pos[:T].abs().sum() > 0
    # the last table row is unreachable on the zero-pad path but reachable on the pooled path;
    # an index that never occurs gets exactly zero gradient
    sku = grads[plan.tables["Sku"].scope]
    used = set()
    for p in plan.pooled:
        if p.table == "Sku":
            ids = batch[p.feature].values.long()
            used |= set(ids.tolist()) | set((ids - 1).clamp(min=0).tolist())
    unused = [r for r in range(plan.tables["Sku"].rows) if r not in used][:50]
    assert sku[unused].abs().sum() == 0


def test_loss_hand_computed(ctx):
    """KAT (vi): five rows, one per label; xent = -log(p_y) away from the clip, weights from dmt.conf."""
    plan, store, batch, O = ctx
    click = torch.tensor([[0.3], [-1.0], [2.0], [0.5], [-0.2]], dtype=torch.float64)
    order = torch.tensor([[-2.0], [-3.0], [0.1], [1.0], [-1.5]], dtype=torch.float64)
    ybias = torch.tensor([[0.1], [0.2], [-0.3], [0.0], [0.4]], dtype=torch.float64)
    mask = torch.eye(5, dtype=torch.float64)                  # labels 0,1,2,4,5
    got = O.logit_loss_unbias(plan, ((click, order), ybias), mask, "two_head_add", "ctr_rel")
    sig = lambda z: 1 / (1 + math.exp(-z))
    w_ctr, w_cvr = [1, 15, 15, 15, 15], [1, 1, 1, 400, 400]
    y_clk, y_ord = [0, 1, 1, 1, 1], [0, 0, 0, 1, 1]
    tot = 0.0
    for b in range(5):
        c, o, yb = click[b].item(), order[b].item(), ybias[b].item()
        xe = lambda p, y: -math.log(p if y else 1 - p)
        tot += w_ctr[b] * (xe(sig(c + yb), y_clk[b]) + xe(sig(c), y_clk[b])) / 5
        tot += w_cvr[b] * (xe(sig(o + yb), y_ord[b]) + xe(sig(o), y_ord[b])) / 5
    assert abs(got.item() - tot) < 1e-12
    plain = O.logit_loss_unbias(plan, ((click, order), ybias), mask, "two_head_add", "ctr")
    assert plain.item() < got.item()
    # saturated probabilities hit the Keras clip at 1e-7 instead of producing inf
    sat = O.cal_cross_entropy(torch.tensor([[1.0], [0.0]], dtype=torch.float64), torch.tensor([0.0, 1.0]))
    assert torch.allclose(sat, torch.full((2,), -math.log(1e-7), dtype=torch.float64), rtol=1e-6)


def test_tf_adam_closed_form():
    """KAT (vii): first two steps of tf.train.AdamOptimizer; at t=1 the step is lr*g/(|g|+eps*...)."""
    from oracle import dmt_oracle as O
    p = {"w": torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)}
    opt = O.TFAdam(p, lr=1e-3)
    g1 = torch.tensor([0.1, -0.2, 0.0], dtype=torch.float64)
    opt.step({"w": g1})
    lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
    m, v = 0.1 * g1, 0.001 * g1 * g1
    want = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64) - lr_t * m / (v.sqrt() + 1e-8)
    assert torch.allclose(p["w"], want, atol=1e-15)
    assert p["w"][2].item() == 0.5                      # zero gradient, zero moments: no movement yet
    g2 = torch.tensor([0.0, 0.0, 0.0], dtype=torch.float64)
    before = p["w"].clone()
    opt.step({"w": g2})
    # dense semantics: rows with zero gradient still move while m != 0
    assert (p["w"][:2] != before[:2]).all()
    lr_t2 = 1e-3 * math.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)
    m2, v2 = 0.9 * m, 0.999 * v
    assert torch.allclose(p["w"], before - lr_t2 * m2 / (v2.sqrt() + 1e-8), atol=1e-15)
    assert O.piecewise_constant(5, [10], [1e-3, 1e-4]) == 1e-3
    assert O.piecewise_constant(10, [10], [1e-3, 1e-4]) == 1e-3
    assert O.piecewise_constant(11, [10], [1e-3, 1e-4]) == 1e-4


def test_fp32_oracle_close_to_fp64(ctx):
    plan, store, batch, O = ctx
    a = O.inference(plan, O.params_from_store(store, torch.float64), batch)
    b = O.inference(plan, O.params_from_store(store, torch.float32), batch)
    assert (a[0][0] - b[0][0].double()).abs().max() < 5e-5
