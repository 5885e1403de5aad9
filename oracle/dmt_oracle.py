"""CPU oracle for the DMT hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A literal PyTorch-CPU restatement of the reference's TF-1.12 graph for
`model_type = mmoe_transformer_unbias`.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s CPU-baseline / `--impl reference` legs may import this module; the
product package (`cikm2020_dmt_b200/`) never does.

Where the arithmetic lives: in the un-vendored third-party `tensorflow==1.12`
(pinned only by /root/reference/README.md:14).  TF cannot be installed here, so every
function below restates the *published* semantics of the TF op the reference calls, and
cites the reference call site it follows (paths relative to /root/reference/DMT_code).

Parity pin status: the reference ships no tests or golden vectors.  The
structure of this restatement (scope sharing, concat order, row offsets, loss
wiring) is pinned by running the reference's OWN unmodified model-building code
under a TF-1 API shim (`oracle/tf1_shim`, primitives restated) -- see
`tests/golden/make_golden.py` and `tests/test_oracle_vs_reference_graph.py`.  The TF
primitives themselves (matmul, softmax, moments, Adam) remain restated from their
documentation: **parity unpinned against a TF-1.12 binary**.

Quirks reproduced on purpose:
  * zero-pad off-by-one: lookup index i reads variable row i-1 on the sequence /
    target path (base.py:87-89), row i on the pooled path (base.py:115-116);
  * encoder block i and decoder block i share the feed-forward + its LayerNorm
    (TransformerModel.py:107,121,155,168); no output projection W_O;
  * query masking AFTER the softmax with -2**32+1 (TransformerModel_util.py:48,90-97);
  * LayerNorm epsilon 1e-8 inside the sqrt, biased variance (TransformerModel_util.py:58-78);
  * Keras clipped sparse-categorical cross-entropy (inference_mlp.py:162-168);
  * TF Adam with epsilon outside the bias correction, dense over every row.
"""
import math
from typing import Dict, List, Optional

import torch

PADDING_NUM = float(-2 ** 32 + 1)      # TransformerModel_util.py:81
LN_EPS = 1e-8                          # TransformerModel_util.py:58
KERAS_EPS = 1e-7                       # keras.backend.epsilon()
TASK_NAMES = ("click", "order")        # mmoe_transformer_unbias.py:301


# --------------------------------------------------------------------------- inputs
def _lengths(sp):
    return (sp.offsets[1:] - sp.offsets[:-1]).long()


def sparse_to_dense(sp, pad=0):
    """tf.sparse.to_dense of a left-packed SparseTensor whose dense_shape is
    [B, max_len_in_batch] (tfrecord_mask.py batches VarLen features that way)."""
    lens = _lengths(sp)
    B = lens.numel()
    T = int(lens.max().item()) if B else 0
    out = torch.full((B, T), pad, dtype=torch.int64)
    m = torch.arange(T)[None, :] < lens[:, None]
    out[m] = sp.values.long()
    return out


def _wts(inputs, feature):
    """`inputs[feature + 'Wts']` when present (base.py:107-111)."""
    sp = inputs[feature]
    w = inputs.get(feature + "Wts")
    if w is not None:
        return w.values if hasattr(w, "values") and not torch.is_tensor(w) else w
    return getattr(sp, "weights", None)


# --------------------------------------------------------------------------- A1 base.embedding
def embedding(var, zero_pad=False):
    """base.py:81-91 -- with zero_pad the lookup table is concat(0[1,D], var)."""
    if zero_pad:
        return torch.cat([torch.zeros(1, var.shape[1], dtype=var.dtype), var], dim=0)
    return var


def lookup_zero_pad_lean(var, idx):
    """Same values as embedding(var, True)[idx] without materialising the concat
    (row_zp).  Used by the 'lean' CPU baseline only."""
    rows = var[(idx - 1).clamp(min=0)]
    return rows * (idx > 0).unsqueeze(-1).to(var.dtype)


# --------------------------------------------------------------------------- A2 generate_data
def generate_data(plan, P, inputs, lean=False):
    """mmoe_transformer_unbias.py:130-186.  Returns per sequence
    [mask [B,T], lens [B], seq_emb [B,T,d], tar_sku_emb [B,d]]."""
    out = []
    for seq in plan.sequences:
        seq_features, tar_features = [], []
        mask = lens = None
        for f, (user_feature, item_feature) in enumerate(zip(seq.user_features, seq.item_features)):
            # :141-146 mask / lens are recomputed per pair; the last pair wins (:183)
            dense_ids = sparse_to_dense(inputs[user_feature])
            l = _lengths(inputs[user_feature])
            mask = (torch.arange(dense_ids.shape[1])[None, :] < l[:, None]).to(torch.int32)
            lens = mask.sum(1)
            var = P[plan.tables[seq.tables[f]].scope]
            item_ids = inputs[item_feature].values.long()      # flat .values (:158)
            if lean and plan.zero_pad:
                seq_features.append(lookup_zero_pad_lean(var, dense_ids))
                tar_features.append(lookup_zero_pad_lean(var, item_ids))
            else:
                seq_features.append(embedding(var, plan.zero_pad)[dense_ids])       # :154-155
                tar_features.append(embedding(var, plan.zero_pad)[item_ids])        # :157-158
        out.append([mask, lens, torch.cat(seq_features, -1), torch.cat(tar_features, -1)])   # :181-183
    return out


# --------------------------------------------------------------------------- A4-A6 transformer ops
def ln(x, beta, gamma, epsilon=LN_EPS):
    """TransformerModel_util.py:58-78 (tf.nn.moments: biased variance)."""
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    return gamma * ((x - mean) / ((var + epsilon) ** 0.5)) + beta


def _mask(inputs, query_masks, key_masks, kind):
    """TransformerModel_util.py:80-108."""
    paddings = torch.ones_like(inputs) * PADDING_NUM
    if kind == "key":
        h = inputs.shape[0] // key_masks.shape[0]
        km = key_masks.bool().repeat(h, 1)[:, None, :].expand(-1, query_masks.shape[1], -1)
        return torch.where(km, inputs, paddings)
    h = inputs.shape[0] // query_masks.shape[0]
    qm = query_masks.bool().repeat(h, 1)[:, :, None].expand(-1, -1, key_masks.shape[1])
    return torch.where(qm, inputs, paddings)


def scaled_dot_product_attention(Q, K, V, query_masks, key_masks, dropout_rate, training, site=None):
    """TransformerModel_util.py:11-56."""
    d_k = Q.shape[-1]
    outputs = torch.matmul(Q, K.transpose(1, 2))              # :30
    outputs = outputs / d_k ** 0.5                            # :33
    outputs = _mask(outputs, query_masks, key_masks, "key")   # :36
    outputs = torch.softmax(outputs, dim=-1)                  # :43
    outputs = _mask(outputs, query_masks, key_masks, "query") # :48 (after the softmax)
    outputs = _dropout(outputs, dropout_rate, training, site)     # :51
    return torch.matmul(outputs, V)                           # :54


def multihead_attention(P, scope, queries, keys, values, queries_length, keys_length,
                        num_heads, dropout_rate, training):
    """TransformerModel_util.py:160-209; `scope` ends in self-attention / vanilla_attention."""
    query_masks = torch.arange(queries.shape[1])[None, :] < queries_length[:, None]   # :182
    key_masks = torch.arange(keys.shape[1])[None, :] < keys_length[:, None]           # :183
    Q = queries @ P[scope + "/dense/kernel"] + P[scope + "/dense/bias"]               # :188
    K = keys @ P[scope + "/dense_1/kernel"] + P[scope + "/dense_1/bias"]              # :189
    V = values @ P[scope + "/dense_2/kernel"] + P[scope + "/dense_2/bias"]            # :190
    Q_ = torch.cat(torch.chunk(Q, num_heads, dim=2), dim=0)                           # :193-195
    K_ = torch.cat(torch.chunk(K, num_heads, dim=2), dim=0)
    V_ = torch.cat(torch.chunk(V, num_heads, dim=2), dim=0)
    outputs = scaled_dot_product_attention(Q_, K_, V_, query_masks, key_masks, dropout_rate, training,
                                           site=("probs", scope))
    outputs = torch.cat(torch.chunk(outputs, num_heads, dim=0), dim=2)                # :201
    outputs = outputs + queries                                                       # :204
    return ln(outputs, P[scope + "/ln/beta"], P[scope + "/ln/gamma"])                 # :207


def ff(P, scope, inputs):
    """TransformerModel_util.py:212-235."""
    outputs = torch.relu(inputs @ P[scope + "/dense/kernel"] + P[scope + "/dense/bias"])
    outputs = outputs @ P[scope + "/dense_1/kernel"] + P[scope + "/dense_1/bias"]
    outputs = outputs + inputs
    return ln(outputs, P[scope + "/ln/beta"], P[scope + "/ln/gamma"])


# TF's RNG stream cannot be reproduced.  Tests that compare TRAINING mode install a hook
# `(site, x, rate) -> multiplier tensor (0 or 1/(1-rate))` so that the oracle drops exactly the elements the
# CUDA path drops (tests/test_gpu_dropout.py); without a hook the sites draw from torch's generator.
DROPOUT_HOOK = None


def _dropout(x, rate, training, site=None):
    """tf.layers.dropout(x, rate, training) -- inverted dropout, keep probability 1 - rate."""
    if training and rate > 0:
        if DROPOUT_HOOK is not None:
            return x * DROPOUT_HOOK(site, x, rate).to(x.dtype)
        return torch.nn.functional.dropout(x, rate, True)
    return x


def sin_cos_position_table(maxlen, E):
    """positional_encoding (TransformerModel_util.py:237-278): numpy fp64 table, cast to fp32 by
    tf.convert_to_tensor(position_enc, tf.float32)."""
    import numpy as np
    position_enc = np.array([[pos / np.power(10000, (i - i % 2) / E) for i in range(E)] for pos in range(maxlen)])
    position_enc[:, 0::2] = np.sin(position_enc[:, 0::2])
    position_enc[:, 1::2] = np.cos(position_enc[:, 1::2])
    return torch.from_numpy(position_enc.astype(np.float32))


def encode(plan, P, scope, seq_emb, seqlens, training):
    """TransformerModel.py:84-123 + positional_encoding_learn (util:281-316)."""
    enc = seq_emb * plan.d_model ** 0.5                                               # :97
    T = enc.shape[1]
    if getattr(plan, "position_encoding_method", "position_learn") == "position_sin_cos":
        pos = sin_cos_position_table(plan.maxlen_k, plan.d_model).to(enc.dtype)         # :63-65, a constant
    else:
        pos = P[scope + "/positional_encoding_k_position_learn/embedding_position_learn"]
    enc = enc + pos[torch.arange(T)][None, :, :]                                      # :67-69
    enc = _dropout(enc, plan.dropout_rate, training, ("enc_in", scope))               # :101
    for i in range(plan.num_blocks_encode):
        blk = "%s/num_blocks_%d" % (scope, i)
        enc = multihead_attention(P, blk + "/self-attention", enc, enc, enc, seqlens, seqlens,
                                  plan.num_heads, plan.dropout_rate, training)
        enc = ff(P, blk + "/positionwise_feedforward", enc)
    return enc


def decode(plan, P, scope, query_emb, query_length, memory, key_length, training):
    """TransformerModel.py:125-171."""
    dec = query_emb * plan.d_model ** 0.5                                             # :147
    dec = _dropout(dec, plan.dropout_rate, training, ("dec_in", scope))               # :151
    for i in range(plan.num_blocks_decode):
        blk = "%s/num_blocks_%d" % (scope, i)
        dec = multihead_attention(P, blk + "/vanilla_attention", dec, memory, memory,
                                  query_length, key_length, plan.num_heads, plan.dropout_rate, training)
        dec = ff(P, blk + "/positionwise_feedforward", dec)      # same variables as the encoder's FF
    return dec


def trans_core(plan, P, seq_data, training):
    """mmoe_transformer_unbias.py:189-223 (input/output MLP options are off in dmt.conf)."""
    states = []
    for seq, (mask, lens, seq_emb, tar_emb) in zip(plan.sequences, seq_data):
        seq_q = tar_emb[:, None, :]
        q_lens = torch.ones(seq_q.shape[0], dtype=torch.int64)
        memory = encode(plan, P, seq.scope, seq_emb, lens, training)
        dec = decode(plan, P, seq.scope, seq_q, q_lens, memory, lens, training)
        states.append(dec.squeeze(1))                                                 # TransformerModel.py:58
    return torch.cat(states, -1)


# --------------------------------------------------------------------------- A9 pooled embeddings
def embedding_lookup_sparse_mean(var, sp, weights):
    """tf.nn.embedding_lookup_sparse(var, ids, weights, combiner='mean'):
    sum_j w_j * var[id_j] / sum_j w_j per row (base.py:116)."""
    lens = _lengths(sp)
    B = lens.numel()
    seg = torch.repeat_interleave(torch.arange(B), lens)
    rows = var[sp.values.long()]
    w = torch.ones(rows.shape[0], dtype=var.dtype) if weights is None else weights.to(var.dtype)
    num = torch.zeros(B, var.shape[1], dtype=var.dtype).index_add_(0, seg, rows * w[:, None])
    den = torch.zeros(B, dtype=var.dtype).index_add_(0, seg, w)
    return num / den[:, None]


def embedding_combiner(plan, P, inputs):
    """base.py:93-134 (sim_embed empty in dmt.conf)."""
    cols = [inputs["features"].to(P[next(iter(P))].dtype)] if plan.is_use_feature else []
    for p in plan.pooled:
        var = P[plan.tables[p.table].scope]           # raw table: row == index
        cols.append(embedding_lookup_sparse_mean(var, inputs[p.feature], _wts(inputs, p.feature)))
    return torch.cat(cols, 1)


# --------------------------------------------------------------------------- A10 MMoE + towers
def dense_layer(P, scope, x, activation):
    """base.py:39-68 with is_bn / is_dropout off."""
    y = x @ P[scope + "/weights"] + P[scope + "/biases"]
    return activation(y)


def expert_gate(plan, P, x):
    """mmoe_transformer_unbias.py:63-105."""
    experts = []
    for e in range(plan.num_experts):
        y = x
        for l in range(len(plan.hidden_units_bottom)):
            y = dense_layer(P, "DnnModel/mmoe_layers/expert-%d/expert-layer-%d" % (e, l), y, torch.relu)
        experts.append(y)
    gates = [dense_layer(P, "DnnModel/mmoe_layers/gates-%d/gates-layer-0" % t, x,
                         lambda z: torch.softmax(z, -1)) for t in range(plan.num_tasks)]
    stacked = torch.stack(experts, -1)                              # [B, H, E]
    return [(stacked * g[:, None, :]).sum(2) for g in gates]        # :99-104


def build_tower(plan, P, task_layer, name):
    """mmoe_transformer_unbias.py:107-126."""
    y = task_layer
    for l in range(len(plan.hidden_units_task)):
        y = dense_layer(P, "DnnModel/%s/%s-fc-%d" % (name, name, l), y, torch.relu)
    return dense_layer(P, "DnnModel/%s/%s-output" % (name, name), y, lambda z: z)


# --------------------------------------------------------------------------- A11 bias net
def embedding_mlp_bias(plan, P, inputs, training):
    """mmoe_transformer_unbias.py:235-289."""
    cols = []
    for p in plan.bias_pooled:
        var = P[plan.bias_tables[p.table].scope]
        cols.append(embedding_lookup_sparse_mean(var, inputs[p.feature], _wts(inputs, p.feature)))
    y = torch.cat(cols, 1)
    n_hidden = len(plan.hidden_units_bias)
    for l in range(n_hidden):
        y = torch.relu(y @ P["DnnModel/layer_bias%d/kernel" % l] + P["DnnModel/layer_bias%d/bias" % l])
        y = _dropout(y, plan.dropout_rate_bias[l], training, ("bias", l))
    return y @ P["DnnModel/layer_bias%d/kernel" % n_hidden] + P["DnnModel/layer_bias%d/bias" % n_hidden]


# --------------------------------------------------------------------------- inference
def inference(plan, P, inputs, is_train=False, is_predict=False, lean=False, return_aux=False):
    """mmoe_transformer_unbias.py:293-316 (+ embedding_trans :226-233)."""
    seq_data = generate_data(plan, P, inputs, lean=lean)
    interest = trans_core(plan, P, seq_data, is_train)
    features = embedding_combiner(plan, P, inputs)
    x = torch.cat([features, interest], -1)
    tasks = expert_gate(plan, P, x)
    y_rel = tuple(build_tower(plan, P, t, TASK_NAMES[i]) for i, t in enumerate(tasks))
    aux = {"x": x, "interest": interest, "seq_data": seq_data}
    if is_predict:
        return (y_rel, aux) if return_aux else y_rel
    y_bias = embedding_mlp_bias(plan, P, inputs, is_train)
    out = (y_rel, y_bias)
    return (out, aux) if return_aux else out


# --------------------------------------------------------------------------- A12 loss
def cal_cross_entropy(output, labels):
    """inference_mlp.py:162-168: Keras sparse_categorical_crossentropy(from_logits=False)
    = clip to [eps, 1-eps], log, sparse softmax cross-entropy with those 'logits'."""
    p = output.reshape(-1, 1)
    p = torch.cat([1 - p, p], -1)
    logits = torch.log(p.clamp(KERAS_EPS, 1 - KERAS_EPS))
    logp = torch.log_softmax(logits, -1)
    return -logp.gather(1, labels.long().reshape(-1, 1)).squeeze(1)


def probabilities(logits, loss_unbias_method="two_head_add"):
    """run_dnn.py:90-100 / inference_mlp.py:176-185."""
    (click_logit, order_logit), y_bias = logits
    if loss_unbias_method == "two_head_multiply":
        return (torch.sigmoid(click_logit) * torch.sigmoid(y_bias),
                torch.sigmoid(order_logit) * torch.sigmoid(y_bias))
    return torch.sigmoid(click_logit + y_bias), torch.sigmoid(order_logit + y_bias)


def logit_loss_unbias(plan, logits, mask, loss_unbias_method=None, loss_ctr_rel_method=None):
    """inference_mlp.py:173-223 (fixed loss weights)."""
    loss_unbias_method = loss_unbias_method or plan.loss_unbias_method
    loss_ctr_rel_method = loss_ctr_rel_method or plan.loss_ctr_rel_method
    (click_logit, order_logit), _ = logits
    p_ctr, p_cvr = probabilities(logits, loss_unbias_method)
    mask = mask.to(click_logit.dtype)
    labels_clk = mask[:, 1:5].sum(-1)                      # :192
    labels_order = mask[:, 3] + mask[:, 4]                 # :193
    losses = []
    for p, p_rel, y, w in ((p_ctr, torch.sigmoid(click_logit), labels_clk, plan.weight_ctr),
                           (p_cvr, torch.sigmoid(order_logit), labels_order, plan.weight_ecvr)):
        xent = cal_cross_entropy(p, y)
        if loss_ctr_rel_method == "ctr_rel":
            xent = xent + cal_cross_entropy(p_rel, y)
        mask_weight = mask * torch.tensor(w, dtype=mask.dtype)          # [B,5]
        entropy_mat = mask_weight.t() * xent                            # [5,B]
        losses.append(entropy_mat.mean(1).sum())
    return plan.loss_weight[0] * losses[0] + plan.loss_weight[1] * losses[1]


# --------------------------------------------------------------------------- A13 TF-1 Adam
class TFAdam:
    """tf.train.AdamOptimizer(lr) (inference_mlp.py:272-273), dense semantics:
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; theta -= lr_t*m/(sqrt(v)+eps)."""

    def __init__(self, params: Dict[str, torch.Tensor], lr=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.params = params
        # the hyper-parameters reach TF's ApplyAdam kernel as float32 scalars: (1 - beta2) is
        # 1 - float32(0.999) = 0.00099998713, not 0.001
        f32 = lambda x: float(torch.tensor(x, dtype=torch.float32))
        self.lr, self.b1, self.b2, self.eps = f32(lr), f32(beta1), f32(beta2), f32(epsilon)
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, grads: Dict[str, torch.Tensor], lr=None):
        self.t += 1
        lr = self.lr if lr is None else float(torch.tensor(lr, dtype=torch.float32))
        lr_t = lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        with torch.no_grad():
            for k, p in self.params.items():
                g = grads.get(k)
                if g is None:
                    g = torch.zeros_like(p)
                if g.is_sparse:
                    g = g.to_dense()     # run_dnn.py:63-72 densifies IndexedSlices
                self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
                self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                p.sub_(lr_t * self.m[k] / (self.v[k].sqrt() + self.eps))


class TFGradientDescent:
    """tf.train.GradientDescentOptimizer(lr) (inference_mlp.py:266-267): theta -= lr * g."""

    def __init__(self, params: Dict[str, torch.Tensor], lr=1e-3):
        self.params, self.lr = params, float(torch.tensor(lr, dtype=torch.float32))

    def step(self, grads: Dict[str, torch.Tensor], lr=None):
        lr = self.lr if lr is None else float(torch.tensor(lr, dtype=torch.float32))
        with torch.no_grad():
            for k, p in self.params.items():
                g = grads.get(k)
                if g is not None:
                    p.sub_(lr * (g.to_dense() if g.is_sparse else g))


class TFAdagrad:
    """tf.train.AdagradOptimizer(lr) (inference_mlp.py:270-271; TF-1.12 default initial_accumulator_value = 0.1):
    acc += g^2 ; theta -= lr * g / sqrt(acc)."""

    def __init__(self, params: Dict[str, torch.Tensor], lr=1e-3, initial_accumulator_value=0.1):
        self.params, self.lr = params, float(torch.tensor(lr, dtype=torch.float32))
        self.acc = {k: torch.full_like(v, initial_accumulator_value) for k, v in params.items()}

    def step(self, grads: Dict[str, torch.Tensor], lr=None):
        lr = self.lr if lr is None else float(torch.tensor(lr, dtype=torch.float32))
        with torch.no_grad():
            for k, p in self.params.items():
                g = grads.get(k)
                if g is None:
                    continue
                g = g.to_dense() if g.is_sparse else g
                self.acc[k].addcmul_(g, g)
                p.sub_(lr * g / self.acc[k].sqrt())


def piecewise_constant(step, boundaries, values):
    """tf.train.piecewise_constant (run_dnn.py:125-126): values[0] while step <= boundaries[0]."""
    for b, v in zip(boundaries, values):
        if step <= b:
            return v
    return values[len(boundaries)]


# --------------------------------------------------------------------------- helpers for tests / baselines
def params_from_store(store, dtype=torch.float64, requires_grad=False) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in store.named_parameters():
        t = v.detach().to("cpu", dtype).clone()
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def loss_and_grads(plan, P, inputs, is_train=False):
    for p in P.values():
        p.requires_grad_(True)
        p.grad = None
    logits = inference(plan, P, inputs, is_train=is_train)
    loss = logit_loss_unbias(plan, logits, inputs["mask"])
    loss.backward()
    return loss.detach(), {k: p.grad for k, p in P.items() if p.grad is not None}, logits
