"""Golden vectors from the reference's OWN model-building code, executed under the TF-1 API shim.

Runs, unmodified, from /root/reference/DMT_code:
    model/inference_mlp.py::Inference            (plugin dispatch, loss_multi_task_unbias)
    model/net/mmoe_transformer_unbias.py         (generate_data, trans_core, MMoE, bias net)
    model/net/TransformerModel.py, TransformerModel_util.py, base.py, mmoe.py
    conf/recsys_conf.py::Conf                    (the config object the model reads)
with `tensorflow` resolved to oracle/tf1_shim/tensorflow (primitives restated on torch fp64).

The config is the reference's dmt.conf with ONLY (a) [path] localised, (b) vocabulary sizes and MLP
widths shrunk so the fixture stays small (the structure -- 23 emb entries, 3 sequences x 5 pairs,
d_model 80, 4 heads, 1+1 blocks, 4 experts, bias tower -- is untouched).

Output: tests/golden/ref_graph.npz  (variables by TF name as fp32, inputs, outputs as fp64) and
tests/golden/ref_graph.conf (the exact config text both sides parse).
"""
import contextlib
import io
import os
import re
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
REF_CODE = os.path.join(REF, "DMT_code")

SMALL = {"Sku:5000000:": "Sku:600:", "Brand:190000:": "Brand:200:", "Shopid:230000:": "Shopid:220:",
         "Cid3:12000:": "Cid3:100:", "Cid2:500:": "Cid2:30:"}
MODEL_EDITS = {"hidden_units_bottom": "48,24,16", "hidden_units_task": "8", "transformer_d_ff": "32",
               "feature_dimension": "16", "batch_size": "8"}


VARIANT = sys.argv[1] if len(sys.argv) > 1 else ""          # "" | "sincos" (transformer_position_encoding_method)
STEM = "ref_graph" + ("_" + VARIANT if VARIANT else "")


def small_reference_conf_text():
    with open(os.path.join(REF_CODE, "conf/settings/dmt.conf")) as fh:
        text = fh.read()
    if VARIANT == "sincos":
        text, n = re.subn(r"(?m)^transformer_position_encoding_method\s*=.*$",
                          "transformer_position_encoding_method=position_sin_cos", text)
        assert n == 1
    for a, b in SMALL.items():
        text = text.replace(a, b)
    for k, v in MODEL_EDITS.items():
        text, n = re.subn(r"(?m)^%s\s*=.*$" % k, "%s = %s" % (k, v), text)
        assert n == 1, k
    stat = os.path.join(REF, "jd_recsys_demo/stat/stat/part-00000")
    text = re.sub(r"(?m)^train_data_stat_path\s*=.*$", "train_data_stat_path = " + stat, text)
    text = re.sub(r"(?m)^output_path\s*=.*$", "output_path = ./out/", text)
    return text


def to_tf_sparse(tf, sp):
    lens = (sp.offsets[1:] - sp.offsets[:-1]).long()
    B, T = lens.numel(), int(lens.max())
    rows = torch.repeat_interleave(torch.arange(B), lens)
    cols = torch.cat([torch.arange(int(l)) for l in lens]) if B else torch.zeros(0, dtype=torch.long)
    return tf.SparseTensor(torch.stack([rows, cols], 1), sp.values.long(), [B, T])


def golden_graph():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "tf1_shim"))      # `import tensorflow` -> the shim
    for sub in ("model", "model/net", "util", "conf"):
        sys.path.insert(1, os.path.join(REF_CODE, sub))
    import tensorflow as tf
    assert tf.__version__.endswith("shim")
    import recsys_conf                       # reference
    import inference_mlp                     # reference
    from cikm2020_dmt_b200.conf import Conf
    from cikm2020_dmt_b200.data import SparseIds, synthetic_batch
    from cikm2020_dmt_b200.plan import build_plan

    text = small_reference_conf_text()
    conf_out = os.path.join(HERE, STEM + ".conf")
    with open(conf_out, "w") as fh:
        fh.write(re.sub(r"(?m)^train_data_stat_path\s*=.*$", "train_data_stat_path =", text))

    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "dmt.conf"), "w") as fh:
            fh.write(text)
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                ref_conf = recsys_conf.Conf(tmp + "/", "dmt.conf")
        finally:
            os.chdir(cwd)

    # inputs: our synthetic generator on the same (small) plan, then edge cases by hand
    plan = build_plan(Conf(HERE + "/", STEM + ".conf"))
    rows = {n: t.rows for n, t in plan.tables.items()}
    B = 8
    batch = synthetic_batch(plan, B, seed=424242, table_rows=rows)
    g = torch.Generator().manual_seed(7)
    for seq, lens in zip(plan.sequences, ([1, 50, 17, 50, 3, 1, 29, 50], [1, 50, 2, 50, 50, 7, 1, 33],
                                          [10, 1, 4, 10, 10, 2, 1, 9])):
        off = torch.zeros(B + 1, dtype=torch.int32)
        off[1:] = torch.cumsum(torch.tensor(lens), 0)
        for f in list(seq.user_features) + [seq.ts_feature]:
            table = [p.table for p in plan.pooled if p.feature == f][0]
            vals = torch.randint(1, plan.tables[table].rows, (int(off[-1]),), generator=g, dtype=torch.int32)
            vals[0] = 0                       # sample 0: the single token is 'unknow' (index 0)
            batch[f] = SparseIds(vals, off)
    for p in plan.pooled[:5]:                 # item features: make the highest index (V-1) appear
        batch[p.feature].values[1] = plan.tables[p.table].rows - 1
    wts_feats = [plan.pooled[6].feature, plan.pooled[13].feature]     # two `<feature>Wts` tensors
    for f in wts_feats:
        batch[f + "Wts"] = torch.rand(batch[f].values.numel(), generator=g) + 0.5

    features = {"features": tf.constant(batch["features"].double())}
    for k, v in batch.items():
        if isinstance(v, SparseIds):
            features[k] = to_tf_sparse(tf, v)
    for f in wts_feats:
        sp = features[f]
        features[f + "Wts"] = tf.SparseTensor(sp.indices, batch[f + "Wts"].double(), sp.dense_shape)
    mask = tf.constant(batch["mask"].double())
    labels = tf.constant(batch["label"].double())

    def run(reuse, is_predict=False):
        with contextlib.redirect_stdout(io.StringIO()):
            with tf.variable_scope("DnnModel", reuse=reuse):          # run_dnn.py:150
                inf = inference_mlp.Inference(ref_conf)
                out = inf.inference(features, is_train=False, is_predict=is_predict)
                if is_predict:
                    return out, None, inf
                losses = {}
                for unbias in ("two_head_add", "two_head_multiply"):
                    for rel in ("ctr", "ctr_rel"):
                        losses[unbias + "/" + rel] = inf.loss_multi_task_unbias(
                            out, labels, mask, is_train=False, loss_unbias_method=unbias, loss_ctr_rel_method=rel)
                return out, losses, inf

    tf.reset_default_graph(seed=20201019)
    run(reuse=False)                                                  # creates the variables
    gen = torch.Generator().manual_seed(99)
    for name, v in tf.global_variables():                             # perturb constant inits, round to fp32
        if v.dim() == 1:
            v.add_(torch.randn(v.shape, generator=gen, dtype=torch.float64) * 0.05)
        v.copy_(v.float().double())
    (y_rel, y_bias), losses, inf = run(reuse=True)
    y_rel_pred, _, _ = run(reuse=True, is_predict=True)
    assert torch.equal(y_rel_pred[0], y_rel[0])
    interest = inf.model.interest_state

    out = {"var/" + n: v.float().numpy() for n, v in tf.global_variables()}
    out["var_order"] = np.array([n for n, _ in tf.global_variables()])
    for k, v in batch.items():
        if isinstance(v, SparseIds):
            out["in/%s/values" % k] = v.values.numpy()
            out["in/%s/offsets" % k] = v.offsets.numpy()
        else:
            out["in/" + k] = v.numpy()
    out["out/click_logit"] = y_rel[0].numpy()
    out["out/order_logit"] = y_rel[1].numpy()
    out["out/y_bias"] = y_bias.numpy()
    out["out/interest_state"] = interest.numpy()
    for k, v in losses.items():
        out["out/loss/" + k] = np.float64(v.item())
    np.savez_compressed(os.path.join(HERE, STEM + ".npz"), **out)
    n_par = sum(v.numel() for _, v in tf.global_variables())
    print("wrote " + STEM + ".npz: %d variables, %d parameters, click_logit[:3] = %s, loss = %.6f"
          % (len(tf.global_variables()), n_par, y_rel[0][:3].flatten().tolist(), losses["two_head_add/ctr_rel"].item()))


if __name__ == "__main__":
    golden_graph()
