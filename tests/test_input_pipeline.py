"""TFRecord -> CSR batch reader and string-id -> index lookup, against fixtures produced by an
independent parser (google.protobuf) and the reference's own ID tables (tests/golden/make_golden_data.py)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import CONF_DIR

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD, "demo_records.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def payloads():
    from cikm2020_dmt_b200 import tfrecord as T
    return list(T.read_records(os.path.join(GOLD, "demo_records.tfrecord"), verify=True))   # both CRCs checked


def test_framing_and_crc(payloads, tmp_path):
    from cikm2020_dmt_b200 import tfrecord as T
    assert len(payloads) == 10 and all(len(p) > 1000 for p in payloads)
    assert T._crc32c(b"123456789") == 0xE3069283                        # CRC-32C check value
    raw = open(os.path.join(GOLD, "demo_records.tfrecord"), "rb").read()
    bad = bytearray(raw)
    bad[40] ^= 0xFF                                                     # corrupt one payload byte
    p = tmp_path / "bad.tfrecord"
    p.write_bytes(bytes(bad))
    with pytest.raises(IOError):
        list(T.read_records(str(p), verify=True))
    p.write_bytes(raw[:-3])                                             # truncated file
    with pytest.raises(IOError):
        list(T.read_records(str(p)))


def test_example_parser_matches_protobuf(payloads, gold):
    from cikm2020_dmt_b200 import tfrecord as T
    for payload, want in zip(payloads, gold["records"]):
        ex = T.parse_example(payload)
        assert len(ex) == want["n_keys"]
        assert float(ex["label"][0]) == want["label"]
        assert [float(v) for v in ex["mask"]] == want["mask"]
        assert len(ex["features"]) == want["features_len"] == 615
        assert abs(float(np.sum(ex["features"].astype(np.float64))) - want["features_sum"]) < 1e-9
        assert ex["header"][0].decode() == want["header"]
        for k, ids in want["ids"].items():
            assert [v.decode() for v in ex[k]] == ids, k
        for k, s in want["wts_sum"].items():
            assert abs(float(np.sum(ex[k + "Wts"])) - s) < 1e-6


def test_batcher_builds_the_input_contract(payloads, gold):
    """parse_single_line (tfrecord_mask.py:23-84): CSR ids, Wts, mask, em_position / em_page from the header."""
    from cikm2020_dmt_b200.conf import Conf
    from cikm2020_dmt_b200 import tfrecord as T
    conf = Conf(CONF_DIR, "dmt_demo.conf")
    vocab = {n: [] for n in ("Sku", "Brand", "Shopid", "TimeClick", "TimeOrder", "TimeCart")}   # hashed only (no 73 MB table)
    idt = "/root/reference/DMT_code/conf/idtables"
    have_ref = os.path.isdir(idt)
    if not have_ref:
        vocab.update({"Cid2": [], "Cid3": []})
    tables = T.LookupTables(conf, idt, vocab_override=vocab)
    batch = T.ExampleBatcher(conf, tables).batch(payloads)
    B = len(payloads)
    assert batch["features"].shape == (B, 615) and batch["mask"].shape == (B, 5)
    assert torch.equal(batch["mask"].sum(1), torch.ones(B))
    for b, want in enumerate(gold["records"]):
        cols = want["header"].split("\t")
        assert int(batch["em_position"][b]) == min(int(cols[4]), 400)
        assert int(batch["em_page"][b]) == min(int(cols[11]), 100)
        assert float(batch["label"][b]) == want["label"]
    for f in ("clk_seq_sku_7d_50", "ord_seq_sku_12m_10", "cart_seq_sku_12m_10", "item_fea_sku", "near_expo_seq_c2"):
        sp = batch[f]
        lens = (sp.offsets[1:] - sp.offsets[:-1]).tolist()
        assert lens == [len(r["ids"][f]) for r in gold["records"]], f
        assert sp.values.dtype == torch.int32 and sp.offsets.dtype == torch.int32
        assert torch.equal(sp.weights, torch.ones_like(sp.weights))                 # every Wts is 1.0 in the data
        rows = next(int(e[1]) for e in list(conf.embedding_list) + list(conf.embedding_list_bias) if e[3] == f)
        assert int(sp.values.min()) >= 0 and int(sp.values.max()) < rows
    # the five features of one behaviour sequence have equal lengths (what generate_data assumes, :141-146)
    assert torch.equal(batch["clk_seq_sku_7d_50"].offsets, batch["clk_seq_shop_7d_50"].offsets)
    if have_ref:   # in-vocabulary indices are exact: list position in the reference's ID_TABLES
        for f, tab in (("item_c2", "Cid2"), ("item_c3", "Cid3"), ("clk_seq_c2_7d_50", "Cid2"),
                       ("clk_seq_c3_7d_50", "Cid3"), ("near_expo_seq_c2", "Cid2")):
            got = batch[f].values.tolist()
            want = [i for r in gold["records"] for i in r["index"][f]]
            nv = gold["vocab_sizes"][tab]
            assert len(got) == len(want)
            for g, w in zip(got, want):
                assert (g == w) if w >= 0 else (g >= nv), (f, g, w)     # OOV -> a hash bucket behind the vocabulary


def test_id_table_semantics():
    from cikm2020_dmt_b200.tfrecord import IdTable, fingerprint64
    t = IdTable("T", ["unknow", "a", "b"], 10)
    assert t.lookup([b"unknow", b"a", b"b"]).tolist() == [0, 1, 2]
    oov = t.lookup([b"zzz", b"zzz", b"another"])
    assert oov[0] == oov[1] and all(3 <= v < 10 for v in oov)
    assert oov[0] == 3 + fingerprint64(b"zzz") % 7
    assert IdTable("T", ["unknow", "a"], 2).lookup([b"nope"]).tolist() == [0]       # no buckets: default_value=0
    with pytest.raises(ValueError):
        IdTable("T", ["a", "b", "c"], 2)
    assert fingerprint64(b"") == 0x9AE16A3B2F90404F                                 # k2: the empty-string fingerprint
    # known-answer vector from the TensorFlow documentation of tf.strings.to_hash_bucket_fast (the op behind the
    # OOV buckets of index_table_from_tensor, index_tables.py:18-28):
    #   tf.strings.to_hash_bucket_fast(["Hello", "TensorFlow", "2.x"], 3) -> [0, 2, 2]
    # It pins the three sub-branches of FarmHash's 0..16-byte class (1-3, 4-7, 8-16 bytes) -- the class every id
    # string of the reference's vocabularies falls in (longest entry of conf/idtables: 10 bytes).  The 17-32 /
    # 33-64 / > 64 byte classes have no published vector available offline and stay unpinned (they only run).
    assert [fingerprint64(s) % 3 for s in (b"Hello", b"TensorFlow", b"2.x")] == [0, 2, 2]
    for n in (1, 3, 4, 7, 8, 16, 17, 32, 33, 64, 65, 200):                          # every length class runs
        assert 0 <= fingerprint64(bytes(range(n % 251)) * 1 if n < 251 else b"x" * n) < 2 ** 64


def test_batches_shard_by_rank(tmp_path):
    from cikm2020_dmt_b200.conf import Conf
    from cikm2020_dmt_b200 import tfrecord as T
    conf = Conf(CONF_DIR, "dmt_demo.conf")
    prefix = os.path.join(GOLD, "demo_records.tfrecord")
    all_b = list(T.batches(conf, None, prefix, batch_size=3))
    assert [b["features"].shape[0] for b in all_b] == [3, 3, 3, 1]
    r0 = list(T.batches(conf, None, prefix, batch_size=3, world=2, rank=0))
    r1 = list(T.batches(conf, None, prefix, batch_size=3, world=2, rank=1))
    # 10 records -> batches 3,3,3,1: only the complete group (b0, b1) is handed out -- every rank sees the same
    # number of steps (a partial step would hang the collectives of the last step), the remainder is dropped
    assert len(r0) == 1 and len(r1) == 1
    assert torch.equal(r0[0]["features"], all_b[0]["features"]) and torch.equal(r1[0]["features"], all_b[1]["features"])
    for world in (2, 3, 4):
        for bs in (1, 2, 3):
            per_rank = [list(T.batches(conf, None, prefix, batch_size=bs, world=world, rank=r)) for r in range(world)]
            assert len({len(x) for x in per_rank}) == 1, (world, bs)
            assert len(per_rank[0]) == (10 // bs) // world
            assert all(b["features"].shape[0] == bs for x in per_rank for b in x)
    two_epochs = list(T.batches(conf, None, prefix, batch_size=5, epochs=2, shuffle_size=4, drop_remainder=True))
    assert len(two_epochs) == 4


def test_packed_batch_pointer_staging_matches_views():
    """`PackedBatch.unpack_ptrs` (DevArray descriptors: address + length) must describe exactly the arrays that
    `unpack` returns as torch views -- checked on a host copy by reading the memory behind the addresses."""
    import ctypes
    import torch
    from conftest import make_plan
    from cikm2020_dmt_b200.data import DevArray, PackedBatch, SparseIds, synthetic_batch
    conf, plan = make_plan("dmt_d64.conf")
    host = synthetic_batch(plan, 37, seed=5)
    packed = PackedBatch(host, pin=False)
    buf = packed.host.clone()
    views, ptrs = packed.unpack(buf), packed.unpack_ptrs(buf)
    assert set(views) == set(k for k in ptrs if not k.startswith("__"))
    n_sparse = 0
    for k, v in views.items():
        p = ptrs[k]
        if isinstance(v, SparseIds):
            n_sparse += 1
            for a, b in ((v.values, p.values), (v.offsets, p.offsets), (v.weights, p.weights)):
                assert (a is None) == (b is None)
                if a is None:
                    continue
                assert isinstance(b, DevArray) and b.numel() == a.numel() and b.dtype == a.dtype
                assert b.data_ptr() == a.data_ptr()
                raw = (ctypes.c_char * (a.numel() * 4)).from_address(b.data_ptr())
                assert bytes(raw) == a.contiguous().view(torch.uint8).numpy().tobytes()
        else:
            assert torch.equal(v, p) and v.data_ptr() == p.data_ptr()
    assert n_sparse >= 20


def test_compact_packed_batch_host_roundtrip():
    """PackedBatch(compact=True): id arrays are stored in the narrowest byte width that holds them (1: time buckets,
    2: vocabularies < 65536, 3: < 2^24), `features` as bf16, only the requested keys are packed; the host-side unpack
    returns the original ids / offsets and the RNE-rounded features, and `max_len` is unchanged."""
    import torch
    from conftest import make_plan
    from cikm2020_dmt_b200.data import PackedBatch, SparseIds, synthetic_batch
    conf, plan = make_plan("dmt_d64.conf", rows={"Sku": 200000, "Brand": 70000, "Shopid": 900, "Cid3": 300, "Cid2": 60})
    host = synthetic_batch(plan, 37, seed=5, table_rows={"Sku": 200000, "Brand": 70000, "Shopid": 900, "Cid3": 300,
                                                         "Cid2": 60})
    keys = set(plan.all_id_features()) | {"features"}
    wide = PackedBatch(host, pin=False)
    packed = PackedBatch(host, pin=False, compact=True, keys=keys)
    assert packed.nbytes < wide.nbytes and packed.wide_bytes > 0
    out = packed.unpack(packed.host)
    assert set(out) == keys                                    # label / mask were not packed
    width = {k: packed.narrow[o][2] for k, kind, dt, shape, o in packed.layout if kind == "v" and o in packed.narrow}
    assert width["clk_seq_ts_7d_50"] == 1 and width["clk_seq_c2_7d_50"] == 1       # 23 buckets / 60 categories
    assert width["clk_seq_c3_7d_50"] == 2 and width["clk_seq_shop_7d_50"] == 2     # 300 / 900 rows
    assert width["clk_seq_sku_7d_50"] == 3 and width["clk_seq_brand_7d_50"] == 3   # 200000 / 70000 rows: > 16 bits
    for k in keys:
        v = host[k]
        if isinstance(v, SparseIds):
            assert torch.equal(out[k].values, v.values) and torch.equal(out[k].offsets, v.offsets), k
        else:
            assert torch.equal(out[k], v.to(torch.bfloat16))
    assert packed.max_len(plan) == wide.max_len(plan)
