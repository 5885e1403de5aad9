"""Training-mode dropout: the CUDA training forward/backward with the conf's dropout rates against the
CPU oracle driven with the SAME keep masks.

TF's RNG stream is not reproducible, so the mask is a counter-based hash of (seed, site, element index)
(csrc/dropout.cuh, restated in cikm2020_dmt_b200/dropout.py); the oracle's dropout sites call the hook below, which
maps each oracle tensor to the element indices the kernels use.  With identical masks the comparison is exact up
to the usual fp32 (2e-4) / split-bf16 (2e-3) tolerances of tests/test_gpu_backward.py.
"""
import pytest
import torch

from conftest import SMALL_ROWS, make_plan

pytestmark = pytest.mark.gpu


def _hook(plan, host, seed):
    from cikm2020_dmt_b200 import dropout as DO
    LP = min(plan.maxlen_k, 64)
    H, d = plan.num_heads, plan.d_model
    by_scope = {seq.scope: seq for seq in plan.sequences}

    def hook(site, x, rate):
        kind = site[0]
        if kind == "bias":
            B, U = x.shape
            idx = torch.arange(B)[:, None] * U + torch.arange(U)[None, :]
            return DO.multiplier(rate, DO.step_seed(seed, 0, 0), DO.SITE_BIAS + site[1], idx)
        scope = site[1]
        seq = next(s for sc, s in by_scope.items() if scope.startswith(sc))
        sseed = DO.step_seed(seed, 0, 1 + seq.index)
        if kind == "enc_in":
            B, T, _ = x.shape
            off = host[seq.user_features[-1]].offsets[:-1].long()
            idx = ((off[:, None] + torch.arange(T)[None, :])[:, :, None] * d + torch.arange(d)[None, None, :])
            return DO.multiplier(rate, sseed, DO.SITE_ENC_IN, idx)
        if kind == "dec_in":
            B = x.shape[0]
            idx = (torch.arange(B)[:, None] * d + torch.arange(d)[None, :]).view(B, 1, d)
            return DO.multiplier(rate, sseed, DO.SITE_DEC_IN, idx)
        assert kind == "probs"
        blk = int(scope.split("/num_blocks_")[1].split("/")[0])
        HB, Tq, Tk = x.shape
        B = HB // H
        n = torch.arange(HB)
        h, b = n // B, n % B                                    # heads are stacked on the batch axis (:193-195)
        bh = (b * H + h)[:, None, None]
        if scope.endswith("self-attention"):
            idx = (bh * LP + torch.arange(Tq)[None, :, None]) * LP + torch.arange(Tk)[None, None, :]
            return DO.multiplier(rate, sseed, DO.SITE_SELF_PROBS + blk, idx)
        idx = (bh * LP + torch.arange(Tk)[None, None, :]).expand(HB, Tq, Tk)
        return DO.multiplier(rate, sseed, DO.SITE_VANILLA_PROBS + blk, idx)

    return hook


def _setup(conf_file, batch, seed, precision="f32", train_gemm=None, overrides=None):
    from cikm2020_dmt_b200.params import ParamStore
    from cikm2020_dmt_b200.data import synthetic_batch, batch_to
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    from oracle import dmt_oracle as O
    conf, plan = make_plan(conf_file, overrides=overrides)
    assert plan.dropout_rate > 0 and any(r > 0 for r in plan.dropout_rate_bias)   # the reference's training conf
    store = ParamStore(plan, device="cuda", seed=seed + 1).randomize_(seed + 2)
    model = mmoe_transformer_unbias(plan, params=store, precision=precision, train_gemm=train_gemm)
    host = synthetic_batch(plan, batch, seed=seed + 3, table_rows=SMALL_ROWS)
    host["mask"] = torch.nn.functional.one_hot(torch.arange(batch) % 5, 5).float()
    return plan, model, store, host, batch_to(host, "cuda"), O


def _compare(plan, model, store, host, dev, O, seed, rel_tol):
    P = O.params_from_store(store)
    O.DROPOUT_HOOK = _hook(plan, host, seed)
    try:
        loss_ref, grads_ref, _ = O.loss_and_grads(plan, P, host, is_train=True)
    finally:
        O.DROPOUT_HOOK = None
    loss, G = model.compute_gradients(dev, is_train=True, dropout_seed=seed)
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) <= rel_tol * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    scale = max(float(g.abs().max()) for g in grads_ref.values())
    bad = []
    for name in [s.name for s in store.specs] + list(store.tables):
        want = grads_ref.get(name)
        want = torch.zeros_like(P[name]) if want is None else want.double()
        got = G.table_dense(store, name) if name in store.tables else G[name]
        got = got.detach().double().cpu().reshape(want.shape)
        err = (got - want).abs().max().item()
        if err > rel_tol * want.abs().max().item() + 5e-3 * rel_tol * scale:
            bad.append((name[-50:], err, want.abs().max().item()))
    assert not bad, bad
    return loss.item(), loss_ref.item()


def test_training_mode_matches_oracle_with_same_masks_fp32():
    plan, model, store, host, dev, O = _setup("dmt_d64.conf", 40, seed=3)
    _compare(plan, model, store, host, dev, O, seed=0xC0FFEE, rel_tol=2e-4)


def test_training_mode_matches_oracle_dmt_conf_pipeline_bf16x3():
    plan, model, store, host, dev, O = _setup("dmt.conf", 72, seed=5, precision="bf16", train_gemm="bf16x3")
    _compare(plan, model, store, host, dev, O, seed=12345, rel_tol=3e-3)


def test_training_mode_matches_oracle_tf32_tensor_core_attention():
    """dmt_d64.conf on the tf32 engine, whose self-attention runs in the tcgen05 kernel (attn_tc.cu): the kernel must
    draw the attention-probability keep masks per (sample, head, query, key) exactly like the per-sample fp32 kernels.
    (i) training-mode forward with the same seed: interest vectors of the tf32 model within 1e-2 of the fp32 model's
    (one wrong mask index moves an interest vector by O(0.1 .. 1)); (ii) loss within 3e-3 and gradient cosine >= 0.99
    against the oracle driven with the same masks (tf32 is a bf16-class engine; dropout + 400x class weights amplify
    its operand truncation)."""
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    plan, model, store, host, dev, O = _setup("dmt_d64.conf", 200, seed=21, precision="bf16", train_gemm="tf32")
    exact = mmoe_transformer_unbias(plan, params=store, precision="f32")
    c0, c1 = plan.interest_col, plan.interest_col + len(plan.sequences) * plan.d_model
    model.inference(dev, is_train=True, dropout_seed=4242)
    got = model._last["x"][:, c0:c1].clone()
    exact.inference(dev, is_train=True, dropout_seed=4242)
    want = exact._last["x"][:, c0:c1].clone()
    torch.cuda.synchronize()
    assert (got - want).abs().max().item() < 1e-2
    P = O.params_from_store(store)
    O.DROPOUT_HOOK = _hook(plan, host, 4242)
    try:
        loss_ref, grads_ref, _ = O.loss_and_grads(plan, P, host, is_train=True)
    finally:
        O.DROPOUT_HOOK = None
    loss, G = model.compute_gradients(dev, is_train=True, dropout_seed=4242)
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) <= 3e-3 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    for name in ("DnnModel/embedding_trans/Sku/embedding", "DnnModel/mmoe_layers/expert-0/expert-layer-0/weights"):
        want = grads_ref[name].double()
        got = (G.table_dense(store, name) if name in store.tables else G[name]).detach().double().cpu().reshape(want.shape)
        cos = float((got * want).sum() / (got.norm() * want.norm() + 1e-300))
        assert cos >= 0.99, (name, cos)


def test_training_mode_two_blocks():
    ov = {("model", "transformer_num_blocks_encode"): "2", ("model", "transformer_num_blocks_decode"): "2"}
    plan, model, store, host, dev, O = _setup("dmt_d64.conf", 24, seed=9, overrides=ov)
    _compare(plan, model, store, host, dev, O, seed=77, rel_tol=2e-4)


def test_dropout_seed_semantics():
    plan, model, store, host, dev, O = _setup("dmt_d64.conf", 64, seed=11)
    l_eval, _ = model.compute_gradients(dev, is_train=False)
    l_eval = l_eval.item()
    l_a, Ga = model.compute_gradients(dev, is_train=True, dropout_seed=1)
    l_a, ga = l_a.item(), Ga.dense.clone()
    l_b, Gb = model.compute_gradients(dev, is_train=True, dropout_seed=1)
    assert l_a == l_b.item() and torch.equal(ga, Gb.dense)                 # same seed: bit-identical
    l_c, _ = model.compute_gradients(dev, is_train=True, dropout_seed=2)
    assert l_c.item() != l_a and l_a != l_eval                             # masks differ / are active
    l_auto1, _ = model.compute_gradients(dev)                              # default: training mode, fresh seed per call
    l_auto2, _ = model.compute_gradients(dev)
    assert l_auto1.item() != l_auto2.item()
    # eval mode equals the inference path's loss
    out = model.inference(dev, is_train=False)
    assert abs(model.loss(out, dev["mask"]).item() - l_eval) <= 1e-5 * abs(l_eval)


def test_inference_in_training_mode_is_the_forward_of_compute_gradients():
    """`inference(inputs, is_train=True)` (run_dnn.py:154) runs the dropout sites: with the seed of a
    `compute_gradients` call its scores give the same loss, with another seed a different one."""
    plan, model, store, host, dev, O = _setup("dmt_d64.conf", 48, seed=21)
    l_ref, _ = model.compute_gradients(dev, is_train=True, dropout_seed=4242)
    l_ref = l_ref.item()
    out = model.inference(dev, is_train=True, dropout_seed=4242)
    l_inf = model.loss(out, dev["mask"]).item()
    assert abs(l_inf - l_ref) <= 1e-5 * abs(l_ref), (l_inf, l_ref)
    out2 = model.inference(dev, is_train=True, dropout_seed=4243)
    assert abs(model.loss(out2, dev["mask"]).item() - l_ref) > 1e-6 * abs(l_ref)
    (click, order) = model.inference(dev, is_train=True, is_predict=True, dropout_seed=4242)
    torch.cuda.synchronize()
    assert click.shape == (48, 1) and order.shape == (48, 1)
