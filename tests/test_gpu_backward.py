"""A13 gradients: the CUDA training forward + backward (through the C ABI) against autograd over the CPU
oracle (fp64) on the same seeded inputs and parameters.

The reference obtains its gradients from `optimizer.compute_gradients(loss)` (run_dnn.py:181) = tf.gradients
of the graph the oracle restates, so autograd over the restatement is the reference gradient.  Dropout is
compared at rate 0 (TF RNG streams are not reproducible).

Tolerance (fp32 kernels vs fp64 oracle): per variable, |got - want| <= 2e-4 * max|want| + 1e-6 * G, where
G = the largest gradient entry of any variable (a gradient that is analytically zero -- the key bias of an
attention layer, softmax being shift invariant -- comes out as fp32 cancellation noise of that scale).
"""
import pytest
import torch

from conftest import SMALL_ROWS, make_plan

pytestmark = pytest.mark.gpu

NO_DROPOUT = {("model", "transformer_dropout_rate"): "0.0", ("model", "dropout_rate_bias"): "0.0,0.0"}


def _setup(conf_file, batch, seed=0, overrides=None, precision="f32", train_gemm=None, **gen):
    from cikm2020_dmt_b200.params import ParamStore
    from cikm2020_dmt_b200.data import synthetic_batch, batch_to
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    from oracle import dmt_oracle as O
    ov = dict(NO_DROPOUT)
    ov.update(overrides or {})
    conf, plan = make_plan(conf_file, overrides=ov)
    store = ParamStore(plan, device="cuda", seed=seed + 1).randomize_(seed + 2)
    model = mmoe_transformer_unbias(plan, params=store, precision=precision, train_gemm=train_gemm)
    host = synthetic_batch(plan, batch, seed=seed + 3, table_rows=SMALL_ROWS, **gen)
    # make the rare positive labels present so every loss branch carries gradient
    lab = torch.arange(batch) % 5
    host["mask"] = torch.nn.functional.one_hot(lab, 5).float()
    return plan, model, store, host, batch_to(host, "cuda"), O


def _check(name, got, want, rtol=2e-4, atol=1e-7):
    got = got.detach().double().cpu().reshape(want.shape)
    want = want.detach().double().cpu()
    err = (got - want).abs().max().item()
    ref = want.abs().max().item()
    assert err <= rtol * ref + atol, "%s: max abs err %.3e vs max |ref| %.3e" % (name, err, ref)


def _compare_all(plan, model, store, host, dev, O, **loss_kw):
    P = O.params_from_store(store)
    for p in P.values():
        p.requires_grad_(True)
    logits = O.inference(plan, P, host, is_train=False)
    loss_ref = O.logit_loss_unbias(plan, logits, host["mask"], **loss_kw)
    loss_ref.backward()
    loss, G = model.compute_gradients(dev, **loss_kw)
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) <= 2e-5 * abs(loss_ref.item()) + 1e-6
    checked = 0
    scale = max(float(p.grad.abs().max()) for p in P.values() if p.grad is not None)
    for spec in store.specs:
        want = P[spec.name].grad
        if want is None:
            want = torch.zeros_like(P[spec.name])
        _check(spec.name, G[spec.name], want, atol=1e-6 * scale)
        checked += 1
    for name in store.tables:
        want = P[name].grad
        if want is None:
            want = torch.zeros_like(P[name])
        _check(name, G.table_dense(store, name), want, atol=1e-6 * scale)
        checked += 1
    assert checked == len(store.specs) + len(store.tables)
    return G


def test_gradients_match_oracle_autograd_d64():
    plan, model, store, host, dev, O = _setup("dmt_d64.conf", 48)
    _compare_all(plan, model, store, host, dev, O)


def test_gradients_match_oracle_autograd_dmt_conf_d80_h4():
    plan, model, store, host, dev, O = _setup("dmt.conf", 40, seed=4)
    _compare_all(plan, model, store, host, dev, O)


def test_gradients_position_sin_cos():
    """transformer_position_encoding_method=position_sin_cos (the reference parser's default): the position table is
    a constant of the graph -- no variable, no gradient; everything else differentiates as before (fp32 and tf32)."""
    ov = {("model", "transformer_position_encoding_method"): "position_sin_cos"}
    plan, model, store, host, dev, O = _setup("dmt_d64.conf", 48, overrides=ov)
    assert not any("positional_encoding" in s.name for s in store.specs)
    _compare_all(plan, model, store, host, dev, O)
    plan, model, store, host, dev, O = _setup("dmt.conf", 40, overrides=ov, precision="bf16", train_gemm="tf32")
    loss, G = model.compute_gradients(dev)
    torch.cuda.synchronize()
    loss_ref, grads_ref, _ = O.loss_and_grads(plan, O.params_from_store(store), host)
    assert abs(loss.item() - loss_ref.item()) <= 3e-3 * abs(loss_ref.item())


def test_gradients_two_blocks_and_ctr_rel_multiply():
    ov = {("model", "transformer_num_blocks_encode"): "2", ("model", "transformer_num_blocks_decode"): "2"}
    plan, model, store, host, dev, O = _setup("dmt_d64.conf", 24, seed=7, overrides=ov)
    _compare_all(plan, model, store, host, dev, O, loss_unbias_method="two_head_multiply",
                 loss_ctr_rel_method="ctr_rel")


@pytest.mark.parametrize("conf_file,batch,engine,rel_tol,cos_tol",
                         [("dmt_d64.conf", 200, "bf16x3", 2e-3, 0.99999), ("dmt.conf", 72, "bf16x3", 2e-3, 0.99999),
                          ("dmt_d64.conf", 200, "bf16", 1e-1, 0.995),
                          ("dmt_d64.conf", 200, "tf32", 1e-1, 0.998), ("dmt.conf", 72, "tf32", 1e-1, 0.998),
                          ("dmt_d64.conf", 2100, "tf32", 6e-2, 0.999)])
def test_gradients_tensor_core_gemms(conf_file, batch, engine, rel_tol, cos_tol):
    """Training with every GEMM of the path (MMoE forward, all dgrad / wgrad contractions) on tcgen05, fp32
    accumulation, everything else fp32.
      * 'bf16x3' (the default of the bf16 model): operands split hi + lo, three MMAs per product -> per variable
        relative Frobenius error <= 2e-3 and cosine >= 0.99999 against the fp64 oracle gradient;
      * 'bf16': plain bf16 operands (2^-9 rounding per operand through a chain of up to eight GEMMs and a
        loss whose class weights reach 400): relative error <= 1e-1, cosine >= 0.995.
      * 'tf32': the per-token GEMMs of the sequence pipeline on the TMA-fed kind::tf32 engine straight from the
        fp32 activations (operand mantissa truncated to 10 bits), MMoE included: a bf16-class engine -- relative error
        <= 1e-1, cosine >= 0.998 on the small batches and <= 6e-2 / >= 0.999 on 2100 samples (2100 samples: several row tiles per persistent CTA and several token splits).
    Variables whose exact gradient is ~0 are compared absolutely."""
    plan, model, store, host, dev, O = _setup(conf_file, batch, seed=31, precision="bf16", train_gemm=engine)
    P = O.params_from_store(store)
    loss_ref, grads_ref, _ = O.loss_and_grads(plan, P, host)
    loss, G = model.compute_gradients(dev)
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) <= {"bf16": 1e-2, "tf32": 3e-3}.get(engine, 1e-4) * abs(loss_ref.item())
    scale = max(float(g.abs().max()) for g in grads_ref.values())
    bad = []
    for name in [s.name for s in store.specs] + list(store.tables):
        want = grads_ref.get(name)
        want = torch.zeros_like(P[name]) if want is None else want.double()
        got = (G[name] if name in G.views and name not in store.tables else G.table_dense(store, name))
        got = got.detach().double().cpu().reshape(want.shape)
        wn = want.norm().item()
        if wn <= 1e-5 * scale * want.numel() ** 0.5:          # analytically (near) zero gradient
            if (got - want).abs().max().item() > 2e-3 * scale:
                bad.append((name, "abs", (got - want).abs().max().item()))
            continue
        rel = (got - want).norm().item() / wn
        cos = float((got * want).sum() / (got.norm() * want.norm() + 1e-300))
        if rel > rel_tol or cos < cos_tol:
            bad.append((name[-50:], round(rel, 4), round(cos, 5)))
    assert not bad, bad


@pytest.mark.parametrize("engine", ["bf16x3", "tf32"])
def test_bf16_gemm_gradients_are_deterministic(engine):
    plan, model, store, host, dev, O = _setup("dmt_d64.conf", 300, seed=5, precision="bf16", train_gemm=engine)
    _, G = model.compute_gradients(dev)
    first = G.dense.clone()
    _, G = model.compute_gradients(dev)
    torch.cuda.synchronize()
    assert torch.equal(first, G.dense)


def test_gradients_are_deterministic_and_batch_split_additive():
    """No floating-point atomics on the path: two runs give bit-identical gradients; and since the loss is a
    batch mean, the gradient of a batch is the mean of the gradients of its halves (the data-parallel
    `average_gradients` identity, run_dnn.py:45-80)."""
    from cikm2020_dmt_b200.data import synthetic_batch, batch_to
    plan, model, store, host, dev, O = _setup("dmt_d64.conf", 64, seed=11)
    _, G = model.compute_gradients(dev)
    first = G.dense.clone()
    _, G = model.compute_gradients(dev)
    torch.cuda.synchronize()
    assert torch.equal(first, G.dense)


def test_train_steps_follow_oracle_adam():
    """Three full steps (forward, backward, TF-1 Adam over dense + sparse variables) track the oracle."""
    from cikm2020_dmt_b200.optim import TFAdam
    from cikm2020_dmt_b200.data import synthetic_batch, batch_to
    plan, model, store, host, dev, O = _setup("dmt_d64.conf", 32, seed=21)
    opt = TFAdam(model, 1e-3)
    P = O.params_from_store(store)
    ref = O.TFAdam(P, lr=1e-3)
    losses = []
    noise = set()   # variables whose exact gradient is zero (attention key biases): Adam amplifies fp32 noise
    for step in range(3):
        h = synthetic_batch(plan, 32, seed=100 + step, table_rows=SMALL_ROWS)
        h["mask"] = torch.nn.functional.one_hot((torch.arange(32) + step) % 5, 5).float()
        loss_ref, grads_ref, _ = O.loss_and_grads(plan, P, h)
        noise |= {k for k, g in grads_ref.items() if float(g.abs().max()) < 1e-12}
        ref.step(grads_ref)
        loss, G = model.compute_gradients(batch_to(h, "cuda"))
        opt.apply_gradients(G)
        losses.append((loss.item(), loss_ref.item()))
    torch.cuda.synchronize()
    for got, want in losses:
        assert abs(got - want) <= 1e-3 * abs(want) + 1e-5, losses
    # Adam's first steps move every coordinate by ~lr regardless of gradient scale, so compare absolutely
    assert all(k.endswith("attention/dense_1/bias") for k in noise), noise
    for name, v in store.named_parameters():
        if name in noise:
            continue
        got = v.detach().double().cpu()
        want = P[name].detach()
        err = (got - want).abs()
        # a coordinate whose exact gradient is ~1e-9 may take an Adam step of the wrong sign in fp32
        # (|step| ~ lr * |g| / (|g| + eps)); everything else must track closely
        assert err.max().item() <= 1e-3, "%s drifted by %.3e after 3 steps" % (name, err.max().item())
        assert err.mean().item() <= 2e-5, "%s mean drift %.3e after 3 steps" % (name, err.mean().item())


@pytest.mark.parametrize("name", ["sgd", "adagrad"])
def test_train_steps_follow_oracle_sgd_and_adagrad(name):
    """`get_optimizer('sgd' | 'adagrad', lr)` (inference_mlp.py:266-271): three full steps through the same sorted
    segmented-reduction kernels (`dmt_adam_cfg.kind`) track tf.train.GradientDescentOptimizer / AdagradOptimizer
    (initial accumulator 0.1) restated in the oracle; rows without a gradient do not move."""
    from cikm2020_dmt_b200.inference import Inference
    from cikm2020_dmt_b200.optim import TFAdagrad, TFGradientDescent
    from cikm2020_dmt_b200.data import synthetic_batch, batch_to
    plan, model, store, host, dev, O = _setup("dmt_d64.conf", 32, seed=23)
    lr = 0.05
    opt = {"sgd": TFGradientDescent, "adagrad": TFAdagrad}[name](model, lr)
    P = O.params_from_store(store)
    ref = {"sgd": O.TFGradientDescent, "adagrad": O.TFAdagrad}[name](P, lr=lr)
    sku = "DnnModel/embedding_trans/Sku/embedding"
    before = store.tables[sku].clone()
    touched = torch.zeros(before.shape[0], dtype=torch.bool)
    for step in range(3):
        h = synthetic_batch(plan, 32, seed=300 + step, table_rows=SMALL_ROWS)
        h["mask"] = torch.nn.functional.one_hot((torch.arange(32) + step) % 5, 5).float()
        loss_ref, grads_ref, _ = O.loss_and_grads(plan, P, h)
        ref.step(grads_ref)
        loss, G = model.compute_gradients(batch_to(h, "cuda"))
        opt.apply_gradients(G)
        assert abs(loss.item() - loss_ref.item()) <= 1e-3 * abs(loss_ref.item()) + 1e-5
        g = grads_ref[sku]
        touched |= (g.to_dense() if g.is_sparse else g).abs().sum(1) > 0
    torch.cuda.synchronize()
    for pname, v in store.named_parameters():
        if pname.endswith("attention/dense_1/bias"):        # exact gradient 0: fp32 noise (Adagrad normalises it up)
            continue
        err = (v.detach().double().cpu() - P[pname].detach()).abs()
        assert err.max().item() <= 2e-3 and err.mean().item() <= 2e-5, (pname, err.max().item(), err.mean().item())
    after = store.tables[sku].cpu()
    assert torch.equal(after[~touched], before.cpu()[~touched])      # no gradient, no movement
