#!/usr/bin/env python
"""Golden fixtures for the input pipeline and the evaluation metrics.  Run HERE (needs
/root/reference); the outputs are committed:

  tests/golden/demo_records.tfrecord   the first 6 train + 4 test records of jd_recsys_demo, verbatim TFRecord framing
  tests/golden/demo_records.json       what an INDEPENDENT parser (google.protobuf with a dynamically built
                                       tf.train.Example descriptor) reads from them, + the Cid2 / Cid3 indices the
                                       reference's own ID_TABLES give (list position, conf/idtables/*.py imported)
  tests/golden/metrics.json            inputs and outputs of the reference's own metrics/metrics.py
                                       (get_offline_metrics, get_offline_metrics_auc) on seeded synthetic sessions
"""
import importlib
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def example_class():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="ex.proto", package="g", syntax="proto3")
    def msg(name):
        m = fd.message_type.add(); m.name = name; return m
    F = descriptor_pb2.FieldDescriptorProto
    m = msg("BytesList"); f = m.field.add(name="value", number=1, type=F.TYPE_BYTES, label=F.LABEL_REPEATED)
    m = msg("FloatList"); f = m.field.add(name="value", number=1, type=F.TYPE_FLOAT, label=F.LABEL_REPEATED)
    m = msg("Int64List"); f = m.field.add(name="value", number=1, type=F.TYPE_INT64, label=F.LABEL_REPEATED)
    m = msg("Feature")
    m.oneof_decl.add(name="kind")
    m.field.add(name="bytes_list", number=1, type=F.TYPE_MESSAGE, type_name=".g.BytesList", label=F.LABEL_OPTIONAL, oneof_index=0)
    m.field.add(name="float_list", number=2, type=F.TYPE_MESSAGE, type_name=".g.FloatList", label=F.LABEL_OPTIONAL, oneof_index=0)
    m.field.add(name="int64_list", number=3, type=F.TYPE_MESSAGE, type_name=".g.Int64List", label=F.LABEL_OPTIONAL, oneof_index=0)
    m = msg("Features")
    e = m.nested_type.add(name="FeatureEntry"); e.options.map_entry = True
    e.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    e.field.add(name="value", number=2, type=F.TYPE_MESSAGE, type_name=".g.Feature", label=F.LABEL_OPTIONAL)
    m.field.add(name="feature", number=1, type=F.TYPE_MESSAGE, type_name=".g.Features.FeatureEntry", label=F.LABEL_REPEATED)
    m = msg("Example")
    m.field.add(name="features", number=1, type=F.TYPE_MESSAGE, type_name=".g.Features", label=F.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("g.Example"))


def raw_records(path, n):
    out = []
    with open(path, "rb") as fh:
        for _ in range(n):
            head = fh.read(12)
            ln = struct.unpack("<Q", head[:8])[0]
            body = fh.read(ln + 4)
            out.append(head + body)
    return out


def main():
    train = REF + "/jd_recsys_demo/train/2019-12-04_2019-12-04/train/2019-12-04_2019-12-04/data/part-r-00000"
    test = REF + "/jd_recsys_demo/2019-12-04_2019-12-04/test_ord/2019-12-19_2019-12-19/data/part-r-00000"
    recs = raw_records(train, 6) + raw_records(test, 4)
    with open(os.path.join(HERE, "demo_records.tfrecord"), "wb") as fh:
        for r in recs:
            fh.write(r)
    Example = example_class()
    sys.path.insert(0, REF + "/DMT_code/conf")
    tabs = {n: importlib.import_module("idtables." + n).ID_TABLES[n] for n in ("Cid2", "Cid3")}
    pos = {n: {s: i for i, s in reversed(list(enumerate(t)))} for n, t in tabs.items()}
    gold = {"vocab_sizes": {n: len(t) for n, t in tabs.items()}, "records": []}
    for r in recs:
        ln = struct.unpack("<Q", r[:8])[0]
        ex = Example.FromString(r[12:12 + ln])
        feat = ex.features.feature
        rec = {"label": feat["label"].float_list.value[0], "mask": list(feat["mask"].float_list.value),
               "features_sum": float(np.sum(np.asarray(feat["features"].float_list.value, dtype=np.float64))),
               "features_len": len(feat["features"].float_list.value),
               "header": feat["header"].bytes_list.value[0].decode(), "n_keys": len(feat), "ids": {}, "wts_sum": {},
               "index": {}}
        for k in sorted(feat):
            if feat[k].WhichOneof("kind") == "bytes_list" and k != "header":
                vals = [v.decode() for v in feat[k].bytes_list.value]
                rec["ids"][k] = vals
                if k + "Wts" in feat:
                    rec["wts_sum"][k] = float(sum(feat[k + "Wts"].float_list.value))
        for k, tab in (("item_c2", "Cid2"), ("item_c3", "Cid3"), ("clk_seq_c2_7d_50", "Cid2"), ("clk_seq_c3_7d_50", "Cid3"),
                       ("near_expo_seq_c2", "Cid2"), ("near_expo_seq_c3", "Cid3")):
            rec["index"][k] = [pos[tab].get(v, -1) for v in rec["ids"].get(k, [])]      # -1: out of vocabulary
        gold["records"].append(rec)
    with open(os.path.join(HERE, "demo_records.json"), "w") as fh:
        json.dump(gold, fh)

    # ---- metrics: the reference's own implementation on seeded synthetic sessions
    sys.path.insert(0, REF + "/DMT_code/metrics")
    import metrics as RM
    # the reference's cal_auc relies on roc_auc_score RAISING for a one-class group (`except: return 1`), the
    # behaviour of the scikit-learn of its era; the scikit-learn installed here only warns and returns nan
    _roc = RM.roc_auc_score

    def roc_auc_score_raising(y, s):
        if len(set(int(v) for v in y)) < 2:
            raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
        return _roc(y, s)

    RM.roc_auc_score = roc_auc_score_raising
    rng = np.random.default_rng(20201019)
    schema = "expid,pin,expo_time,sid,pos,sku,uuid,click_time,order_id,label,reqsig,page,index".split(",")
    headers, scores = [], []
    for sid in range(60):
        uuid = "u%d" % (sid % 23)
        n = int(rng.integers(1, 25))
        for j in range(n):
            label = int(rng.choice([0, 1, 2, 4, 5], p=[0.7, 0.05, 0.17, 0.04, 0.04]))
            cols = ["e", "p", "t", "s%d" % sid, str(j), "sku", uuid, "c", "o", str(label), "r", "1", str(j)]
            headers.append("\t".join(cols).encode())
            scores.append(float(np.round(rng.random() * 2, 2)))      # rounded: ties exercise the tie rules
    sets, at = RM.get_offline_metrics(schema, headers, scores)
    auc = RM.get_offline_metrics_auc(schema, headers, scores)
    gold = {"schema": schema, "headers": [h.decode() for h in headers], "scores": scores, "at_list": list(at),
            "pre": {str(a): list(map(float, sets[a][0])) for a in sets},
            "mrr": {str(a): list(map(float, sets[a][1])) for a in sets},
            "auc": {str(a): float(auc[a][0]) for a in auc}}
    with open(os.path.join(HERE, "metrics.json"), "w") as fh:
        json.dump(gold, fh)
    print("wrote fixtures:", len(recs), "records,", len(headers), "metric rows")


if __name__ == "__main__":
    main()
