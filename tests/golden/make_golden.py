"""Generate the committed golden fixtures from the REFERENCE's own Python, run in this container.

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

Writes (small, committed):
  tests/golden/conf_dmt.json        -- typed config the reference's `Conf` derives from its dmt.conf
  tests/golden/ref_graph_*.pt       -- inputs / parameters / outputs of the reference's UNMODIFIED
                                       model-building code (`mmoe_transformer_unbias.inference`,
                                       `Inference.loss_multi_task_unbias`) executed under the TF-1 API
                                       shim in `oracle/tf1_shim` (see that package's docstring)

Nothing in tests/, bench.py or the package reads /root/reference at run time; only this script does.
"""
import json
import os
import re
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
REF_CODE = os.path.join(REF, "DMT_code")
sys.path.insert(0, ROOT)


def localised_reference_conf(tmpdir, ord_suffix=None):
    """The reference's dmt.conf with ONLY its [path] entries pointed at the shipped demo data
    (the originals are `~/jd_recsys/...` on the author's machine; `Conf` opens the stat file,
    recsys_conf.py:139-151,340-347)."""
    with open(os.path.join(REF_CODE, "conf/settings/dmt.conf")) as fh:
        text = fh.read()
    stat = os.path.join(REF, "jd_recsys_demo/stat/stat/part-00000")
    text = re.sub(r"(?m)^train_data_stat_path\s*=.*$", "train_data_stat_path = " + stat, text)
    text = re.sub(r"(?m)^output_path\s*=.*$", "output_path = %s/out/" % tmpdir, text)
    if ord_suffix:
        text = text.replace("_12m_50", ord_suffix)
    path = os.path.join(tmpdir, "dmt.conf")
    with open(path, "w") as fh:
        fh.write(text)
    return tmpdir + "/", "dmt.conf"


def golden_conf():
    sys.path.insert(0, os.path.join(REF_CODE, "conf"))
    sys.path.insert(0, os.path.join(REF_CODE, "util"))
    import recsys_conf   # the reference's module, unmodified
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            conf_path, conf_file = localised_reference_conf(tmp)
            c = recsys_conf.Conf(conf_path, conf_file)
        finally:
            os.chdir(cwd)
    model = {k: v for k, v in c["model"].items()}
    out = {
        "model": model,
        "parameter": c["parameter"],
        "class_weight": c["class_weight"],
        "embedding_list": c.embedding_list,
        "embedding_list_bias": c.embedding_list_bias,
        "attention_embed_pairs": c.attention_embed_pairs,
        "attention_embed_seq_ts": c.attention_embed_seq_ts,
        "attrs": {k: getattr(c, k) for k in (
            "tag", "model_type", "zero_pad", "is_unbias_model", "loss_unbias_method", "dropout_rate_bias",
            "loss_ctr_rel_method", "is_use_feature", "d_model", "d_ff", "num_heads", "num_blocks_encode",
            "num_blocks_decode", "maxlen_k", "maxlen_q", "dropout_rate", "is_trans_input_by_mlp",
            "position_encoding_method", "is_use_seq_ts", "is_trans_out_concat_item", "is_trans_out_by_mlp",
            "is_decoder_add_pos_emb", "weight_ctr", "weight_ecvr", "labels", "propensity_em",
            "propensity_em_type")},
        "label_cnt_lst": c.label_cnt_lst,
    }
    with open(os.path.join(HERE, "conf_dmt.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True, default=list)
    print("wrote conf_dmt.json")


if __name__ == "__main__":
    if not os.path.isdir(REF_CODE):
        sys.exit("needs %s" % REF_CODE)
    golden_conf()
    try:
        from make_golden_graph import golden_graph
    except ImportError:
        golden_graph = None
    if golden_graph is not None:
        golden_graph()
