"""Real inputs end to end (BASELINE config 1's data: jd_recsys_demo TFRecords): committed demo records -> TFRecord
reader -> id lookup -> CUDA forward / training step, against the CPU oracle on the same batch; streaming AUC over
the batches; checkpoint -> exact resume."""
import os

import pytest
import torch

from conftest import CONF_DIR, SMALL_ROWS, make_plan

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _demo_batch(conf, plan):
    from cikm2020_dmt_b200 import tfrecord as T
    from cikm2020_dmt_b200.data import SparseIds
    names = {e[0] for e in list(conf.embedding_list) + list(conf.embedding_list_bias)}
    tables = T.LookupTables(conf, "", vocab_override={n: [] for n in names})       # every id takes the hash route
    payloads = list(T.read_records(os.path.join(GOLD, "demo_records.tfrecord")))
    batch = T.ExampleBatcher(conf, tables).batch(payloads)
    # fold the 5M-row indices into the small test tables
    rows = {p.feature: plan.tables[p.table].rows for p in plan.pooled}
    rows.update({p.feature: min(plan.bias_tables[p.table].rows, rows.get(p.feature, 1 << 60)) for p in plan.bias_pooled})
    for f, r in rows.items():
        sp = batch[f]
        batch[f] = SparseIds((sp.values % r).to(torch.int32), sp.offsets, sp.weights)
    return {k: v for k, v in batch.items() if k != "header"}, batch["header"]


def test_demo_records_forward_loss_and_auc_match_oracle():
    from cikm2020_dmt_b200.conf import Conf
    from cikm2020_dmt_b200.data import batch_to
    from cikm2020_dmt_b200.params import ParamStore
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    from cikm2020_dmt_b200 import metrics as M
    from oracle import dmt_oracle as O
    conf, plan = make_plan("dmt_demo.conf")
    host, headers = _demo_batch(conf, plan)
    store = ParamStore(plan, device="cuda", seed=2).randomize_(3)
    model = mmoe_transformer_unbias(plan, params=store)
    dev = batch_to(host, "cuda")
    (click, order), y_bias = model.inference(dev, is_train=False)
    loss, probs, _ = model.loss(((click, order), y_bias), dev["mask"], want_probs=True)
    P = O.params_from_store(store)
    (rc, ro), rb = O.inference(plan, P, host)
    assert torch.allclose(click.double().cpu(), rc, atol=2e-4, rtol=2e-4)
    assert torch.allclose(order.double().cpu(), ro, atol=2e-4, rtol=2e-4)
    assert torch.allclose(y_bias.double().cpu(), rb, atol=1e-4, rtol=1e-4)
    ref_loss = O.logit_loss_unbias(plan, ((rc, ro), rb), host["mask"])
    assert abs(loss.item() - ref_loss.item()) <= 2e-4 * abs(ref_loss.item())
    # streaming click AUC from the device probabilities == the same metric from the oracle's probabilities
    p_ctr, p_cvr = O.probabilities(((rc, ro), rb))
    y_clk, _ = M.click_order_labels(host["mask"])
    got, want = M.StreamingBinaryMetrics(), M.StreamingBinaryMetrics()
    got.update(y_clk, probs[0].cpu())
    want.update(y_clk, p_ctr.reshape(-1))
    assert abs(got.result()["auc"] - want.result()["auc"]) < 1e-6
    # session metrics run on the real headers
    sets, at = M.offline_metrics(conf["schema"]["header_schema"], headers, (probs[0] + probs[1]).cpu().tolist())
    assert len(at) == 7 and all(0.0 <= v <= 1.0 for v in sets[M.CLICK][0])


def test_demo_records_tensor_core_path_auc_within_1e3():
    """the reduced-precision tensor-core path must keep the click AUC of the reference's demo records
    within 1e-3 of the oracle's.  The demo conf is the reference's own shape (d_model 80, 4 heads, d_ff 320), which
    runs on the tf32 pipeline (precision='tf32')."""
    from cikm2020_dmt_b200.data import batch_to
    from cikm2020_dmt_b200.params import ParamStore
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    from cikm2020_dmt_b200 import metrics as M
    from oracle import dmt_oracle as O
    conf, plan = make_plan("dmt_demo.conf")
    assert (plan.d_model, plan.num_heads, plan.d_ff) == (80, 4, 320)
    host, headers = _demo_batch(conf, plan)
    store = ParamStore(plan, device="cuda", seed=2).randomize_(3)
    model = mmoe_transformer_unbias(plan, params=store, precision="tf32")
    dev = batch_to(host, "cuda")
    (click, order), y_bias = model.inference(dev, is_train=False)
    loss, probs, _ = model.loss(((click, order), y_bias), dev["mask"], want_probs=True)
    torch.cuda.synchronize()
    (rc, ro), rb = O.inference(plan, O.params_from_store(store), host)
    assert torch.allclose(click.double().cpu(), rc, atol=2e-2, rtol=1e-2)
    assert torch.allclose(order.double().cpu(), ro, atol=2e-2, rtol=1e-2)
    p_ctr, p_cvr = O.probabilities(((rc, ro), rb))
    y_clk, y_ord = M.click_order_labels(host["mask"])
    for y, got_p, want_p in ((y_clk, probs[0], p_ctr), (y_ord, probs[1], p_cvr)):
        if not 0 < float(y.sum()) < y.numel():      # a label that never / always fires has no AUC
            continue
        got, want = M.StreamingBinaryMetrics(), M.StreamingBinaryMetrics()
        got.update(y, got_p.cpu())
        want.update(y, want_p.reshape(-1))
        assert abs(got.result()["auc"] - want.result()["auc"]) < 1e-3


def test_demo_records_bf16_fused_path_auc_within_1e3():
    """SURVEY 8c for the bf16 route (fused tcgen05 tile kernels, one launch over all sequences, bf16 MMoE assembly,
    native forward driver): the demo records with the d_model-64 shape those kernels are built for -- logits within
    the bf16 tolerance, click / order AUC within 1e-3 of the oracle's."""
    from cikm2020_dmt_b200.data import batch_to
    from cikm2020_dmt_b200.params import ParamStore
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    from cikm2020_dmt_b200 import metrics as M
    from oracle import dmt_oracle as O
    conf, plan = make_plan("dmt_demo_d64.conf")
    assert (plan.d_model, plan.num_heads, plan.d_ff) == (64, 2, 256)
    host, headers = _demo_batch(conf, plan)
    store = ParamStore(plan, device="cuda", seed=2).randomize_(3)
    model = mmoe_transformer_unbias(plan, params=store, precision="bf16")
    assert model.seq_multi and model.x_bf16 and model.fwd_native
    dev = batch_to(host, "cuda")
    (click, order), y_bias = model.inference(dev, is_train=False)
    loss, probs, _ = model.loss(((click, order), y_bias), dev["mask"], want_probs=True)
    torch.cuda.synchronize()
    (rc, ro), rb = O.inference(plan, O.params_from_store(store), host)
    assert torch.allclose(click.double().cpu(), rc, atol=5e-2, rtol=2e-2)
    assert torch.allclose(order.double().cpu(), ro, atol=5e-2, rtol=2e-2)
    p_ctr, p_cvr = O.probabilities(((rc, ro), rb))
    y_clk, y_ord = M.click_order_labels(host["mask"])
    for y, got_p, want_p in ((y_clk, probs[0], p_ctr), (y_ord, probs[1], p_cvr)):
        if not 0 < float(y.sum()) < y.numel():      # a label that never / always fires has no AUC
            continue
        got, want = M.StreamingBinaryMetrics(), M.StreamingBinaryMetrics()
        got.update(y, got_p.cpu())
        want.update(y, want_p.reshape(-1))
        assert abs(got.result()["auc"] - want.result()["auc"]) < 1e-3


def test_checkpoint_resume_is_exact(tmp_path):
    from cikm2020_dmt_b200 import checkpoint as CK
    from cikm2020_dmt_b200.data import synthetic_batch, batch_to
    from cikm2020_dmt_b200.train import Trainer
    conf, plan = make_plan("dmt_d64.conf")         # the conf's dropout stays ON: resume must restore the seed stream
    bs = [batch_to(synthetic_batch(plan, 32, seed=40 + i, table_rows=SMALL_ROWS), "cuda") for i in range(3)]
    a = Trainer(plan, "cuda", seed=5, randomize=6)
    a.train_step(bs[0])
    a.train_step(bs[1])
    d = str(tmp_path / "ck")
    CK.save(d, a.opt.t, a.store, optimizer=a.opt, extra={"train_calls": a.model._train_calls})
    a.train_step(bs[2])
    b = Trainer(plan, "cuda", seed=99)             # different init: everything must come from the checkpoint
    extra = CK.load(d, CK.latest(d), b.store, optimizer=b.opt)
    b.model._train_calls = int(extra["train_calls"])
    b.model.dropout_base_seed = a.model.dropout_base_seed
    b.model.invalidate_prepared()
    b.train_step(bs[2])
    torch.cuda.synchronize()
    for name, v in a.store.named_parameters():
        assert torch.equal(v, b.store.views[name]), name
    assert b.opt.t == a.opt.t == 3
