"""Pin the oracle (and the parameter inventory) against the REFERENCE's own model-building code.

tests/golden/ref_graph.npz was produced by tests/golden/make_golden_graph.py, which imports the
unmodified reference modules (inference_mlp.Inference -> net.mmoe_transformer_unbias -> TransformerModel*,
base) from /root/reference and executes them under the TF-1 API shim (oracle/tf1_shim).  Everything the
reference decides in Python -- variable names and sharing, concat order, row offsets, masks, loss wiring --
is therefore the reference's; only TF primitives are restated.
"""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def gold():
    from cikm2020_dmt_b200.conf import Conf
    from cikm2020_dmt_b200.data import SparseIds
    from cikm2020_dmt_b200.plan import build_plan
    z = np.load(os.path.join(GOLD, "ref_graph.npz"))
    plan = build_plan(Conf(GOLD + "/", "ref_graph.conf"))
    variables = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("var/")}
    batch = {}
    for k in z.files:
        if not k.startswith("in/"):
            continue
        parts = k[3:].split("/")
        if len(parts) == 1:
            batch[parts[0]] = torch.from_numpy(z[k])
    for k in z.files:
        if k.startswith("in/") and k.endswith("/values"):
            name = k[3:-7]
            batch[name] = SparseIds(torch.from_numpy(z[k]), torch.from_numpy(z["in/%s/offsets" % name]))
    outs = {k[4:]: z[k] for k in z.files if k.startswith("out/")}
    return plan, variables, batch, outs, [str(n) for n in z["var_order"]]


def test_variable_inventory_matches_reference_graph(gold):
    """Names AND shapes of every variable the reference graph creates == our parameter store."""
    from cikm2020_dmt_b200.params import dense_param_specs, table_specs
    plan, variables, batch, outs, order = gold
    ours = {s.name: tuple(s.shape) for s in dense_param_specs(plan) + table_specs(plan)}
    theirs = {n: tuple(v.shape) for n, v in variables.items()}
    assert set(ours) == set(theirs), (sorted(set(ours) - set(theirs))[:5], sorted(set(theirs) - set(ours))[:5])
    for n in ours:
        assert ours[n] == theirs[n], (n, ours[n], theirs[n])
    # the feed-forward of encoder and decoder is ONE set of variables
    ff = [n for n in order if "positionwise_feedforward/dense/kernel" in n]
    assert len(ff) == len(plan.sequences)


def test_oracle_matches_reference_graph_outputs(gold):
    from oracle import dmt_oracle as O
    plan, variables, batch, outs, _ = gold
    P = {k: v.double() for k, v in variables.items()}
    (y_rel, y_bias), aux = O.inference(plan, P, batch, is_train=False, return_aux=True)
    assert np.abs(aux["interest"].numpy() - outs["interest_state"]).max() < 1e-9
    assert np.abs(y_rel[0].numpy() - outs["click_logit"]).max() < 1e-10
    assert np.abs(y_rel[1].numpy() - outs["order_logit"]).max() < 1e-10
    assert np.abs(y_bias.numpy() - outs["y_bias"]).max() < 1e-10
    for unbias in ("two_head_add", "two_head_multiply"):
        for rel in ("ctr", "ctr_rel"):
            got = O.logit_loss_unbias(plan, (y_rel, y_bias), batch["mask"], unbias, rel).item()
            want = float(outs["loss/%s/%s" % (unbias, rel)])
            assert abs(got - want) < 1e-10 * max(1.0, abs(want)), (unbias, rel, got, want)
    # is_predict drops the bias tower but not the relevance logits
    y2 = O.inference(plan, P, batch, is_train=False, is_predict=True)
    assert torch.equal(y2[0], y_rel[0])


@pytest.mark.gpu
@pytest.mark.parametrize("precision,atol", [("f32", 2e-4)])
def test_cuda_path_matches_reference_graph_outputs(gold, precision, atol):
    """The CUDA path, through the C ABI, against the vectors the reference's own graph produced."""
    from cikm2020_dmt_b200.data import batch_to
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    from cikm2020_dmt_b200.params import ParamStore
    plan, variables, batch, outs, _ = gold
    store = ParamStore(plan, device="cuda", init=False)
    store.load_state_dict(variables)
    model = mmoe_transformer_unbias(plan, params=store, precision=precision)
    (yr, yb) = model.inference(batch_to(batch, "cuda"), is_train=False)
    torch.cuda.synchronize()
    assert np.abs(yr[0].cpu().numpy() - outs["click_logit"]).max() < atol
    assert np.abs(yr[1].cpu().numpy() - outs["order_logit"]).max() < atol
    assert np.abs(yb.cpu().numpy() - outs["y_bias"]).max() < atol
    x = model._last["x"]
    got = x[:, plan.interest_col:plan.interest_col + 3 * plan.d_model].cpu().numpy()
    assert np.abs(got - outs["interest_state"]).max() < atol
    for unbias in ("two_head_add", "two_head_multiply"):
        for rel in ("ctr", "ctr_rel"):
            loss = model.loss((yr, yb), batch["mask"].cuda(), loss_unbias_method=unbias, loss_ctr_rel_method=rel)
            want = float(outs["loss/%s/%s" % (unbias, rel)])
            assert abs(loss.item() - want) < 2e-4 * max(1.0, abs(want))
