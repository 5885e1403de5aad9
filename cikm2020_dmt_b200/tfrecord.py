"""TFRecord -> CSR batch reader and string-id -> index lookup: the input pipeline in front
of the hot path, producing exactly the feature dict of the A0 input contract.

Replaces, without TensorFlow:

* `tf.data.TFRecordDataset` + `tf.parse_single_example` with the schema of `parse_single_line`
  (data_feed/tfrecord_mask.py:23-84): FixedLen `label` f32, `mask` f32 [len(train_weight)], `features` f32
  [feature_dimension], `header` bytes; VarLen bytes ids + float `<feature>Wts` for every `[embedding] emb` /
  `emb_bias` feature; `em_position` = header field 4 capped at 400, `em_page` = header field 11 capped at 100;
* `LookupTables.transform_id2index` (data_feed/index_tables.py:5-45) = `tf.contrib.lookup.index_table_from_tensor(
  mapping=ID_TABLES[name], num_oov_buckets=V - len(mapping))`: in-vocabulary id -> list position, out-of-vocabulary
  -> `len(mapping) + Fingerprint64(id) % buckets`.  In-vocabulary indices are exact; the OOV hash is TF's FarmHash
  Fingerprint64, an un-vendored third-party routine restated here from its published algorithm -- **parity of OOV
  indices unpinned** (no TF binary to check against);
* `repeat / shuffle / batch` and the per-tower `get_next()` of `get_multi_towers_batch` (:120-158): every
  data-parallel rank takes every `world`-th batch.

Wire formats: TFRecord framing = u64 length, u32 masked crc32c(length), payload, u32 masked crc32c(payload);
`tf.train.Example` = protobuf `Example{1: Features{1: map<string, Feature>}}`, `Feature` = oneof
`{1: BytesList, 2: FloatList, 3: Int64List}` (packed or unpacked repeated fields).
"""
import glob
import os
import re
import struct
from typing import Dict, Iterator, List, Optional

import numpy as np
import torch

from .data import SparseIds, share_offsets

# ----------------------------------------------------------------------------- TFRecord framing
_CRC_TABLE = None


def _crc32c(data: bytes) -> int:
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab.append(c)
        _CRC_TABLE = tab
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked_crc(data: bytes) -> int:
    c = _crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def read_records(path: str, verify: bool = False) -> Iterator[bytes]:
    """Payloads of one TFRecord file.  `verify` checks both masked CRC32C fields (slow pure Python; tests)."""
    with open(path, "rb") as fh:
        while True:
            head = fh.read(12)
            if not head:
                return
            if len(head) < 12:
                raise IOError("%s: truncated record header" % path)
            (n,), (hcrc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            body = fh.read(n + 4)
            if len(body) < n + 4:
                raise IOError("%s: truncated record body" % path)
            payload = body[:n]
            if verify:
                if _masked_crc(head[:8]) != hcrc:
                    raise IOError("%s: length CRC mismatch" % path)
                if _masked_crc(payload) != struct.unpack("<I", body[n:])[0]:
                    raise IOError("%s: payload CRC mismatch" % path)
            yield payload


# ----------------------------------------------------------------------------- tf.train.Example wire parser
def _varint(buf: bytes, pos: int):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes):
    """(field number, wire type, value) triples of one message; length-delimited values are memoryview slices."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield num, wt, v


def _feature(buf: bytes):
    for num, wt, v in _fields(buf):
        if num == 1:      # BytesList
            return [bytes(x) for n2, _, x in _fields(v) if n2 == 1]
        if num == 2:      # FloatList: packed (one length-delimited blob) or repeated fixed32
            out = []
            for n2, wt2, x in _fields(v):
                if n2 != 1:
                    continue
                out.append(np.frombuffer(x, dtype="<f4"))
            return np.concatenate(out) if out else np.zeros(0, np.float32)
        if num == 3:      # Int64List
            out = []
            for n2, wt2, x in _fields(v):
                if n2 != 1:
                    continue
                if wt2 == 0:
                    out.append(x)
                else:
                    p = 0
                    while p < len(x):
                        val, p = _varint(x, p)
                        out.append(val)
            return np.asarray([o - (1 << 64) if o >= (1 << 63) else o for o in out], dtype=np.int64)
    return []


def parse_example(payload: bytes) -> Dict[str, object]:
    """{feature name: list of bytes | float32 array | int64 array} of one serialized tf.train.Example."""
    out = {}
    for num, _, features in _fields(payload):
        if num != 1:
            continue
        for n2, _, entry in _fields(features):
            if n2 != 1:
                continue
            key, val = None, None
            for n3, _, x in _fields(entry):
                if n3 == 1:
                    key = bytes(x).decode()
                elif n3 == 2:
                    val = _feature(x)
            if key is not None:
                out[key] = val if val is not None else []
    return out


# ----------------------------------------------------------------------------- FarmHash Fingerprint64
_K0, _K1, _K2 = 0xC3A5C85C97CB3127, 0xB492B66FBE98F273, 0x9AE16A3B2F90404F
_M64 = (1 << 64) - 1


def _rot(v, s):
    return ((v >> s) | (v << (64 - s))) & _M64 if s else v


def _f64(s, i):
    return struct.unpack_from("<Q", s, i)[0]


def _f32(s, i):
    return struct.unpack_from("<I", s, i)[0]


def _shift_mix(v):
    return v ^ (v >> 47)


def _hash_len16(u, v, mul):
    a = ((u ^ v) * mul) & _M64
    a ^= a >> 47
    b = ((v ^ a) * mul) & _M64
    b ^= b >> 47
    return (b * mul) & _M64


def _weak_hash32(w, x, y, z, a, b):
    a = (a + w) & _M64
    b = _rot((b + a + z) & _M64, 21)
    c = a
    a = (a + x + y) & _M64
    b = (b + _rot(a, 44)) & _M64
    return (a + z) & _M64, (b + c) & _M64


def fingerprint64(s: bytes) -> int:
    """farmhash::Fingerprint64 (= farmhashna::Hash64), the hash behind tf.string_to_hash_bucket_fast and the OOV
    buckets of index_table_from_tensor.  Restated from the published algorithm; unpinned against a TF binary."""
    n = len(s)
    if n <= 16:
        if n >= 8:
            mul = (_K2 + n * 2) & _M64
            a = (_f64(s, 0) + _K2) & _M64
            b = _f64(s, n - 8)
            c = (_rot(b, 37) * mul + a) & _M64
            d = ((_rot(a, 25) + b) * mul) & _M64
            return _hash_len16(c, d, mul)
        if n >= 4:
            mul = (_K2 + n * 2) & _M64
            a = _f32(s, 0)
            return _hash_len16((n + (a << 3)) & _M64, _f32(s, n - 4), mul)
        if n > 0:
            a, b, c = s[0], s[n >> 1], s[n - 1]
            y = (a + (b << 8)) & 0xFFFFFFFF
            z = (n + (c << 2)) & 0xFFFFFFFF
            return (_shift_mix((y * _K2 ^ z * _K0) & _M64) * _K2) & _M64
        return _K2
    if n <= 32:
        mul = (_K2 + n * 2) & _M64
        a = (_f64(s, 0) * _K1) & _M64
        b = _f64(s, 8)
        c = (_f64(s, n - 8) * mul) & _M64
        d = (_f64(s, n - 16) * _K2) & _M64
        return _hash_len16((_rot((a + b) & _M64, 43) + _rot(c, 30) + d) & _M64,
                           (a + _rot((b + _K2) & _M64, 18) + c) & _M64, mul)
    if n <= 64:
        mul = (_K2 + n * 2) & _M64
        a = (_f64(s, 0) * _K2) & _M64
        b = _f64(s, 8)
        c = (_f64(s, n - 8) * mul) & _M64
        d = (_f64(s, n - 16) * _K2) & _M64
        y = (_rot((a + b) & _M64, 43) + _rot(c, 30) + d) & _M64
        z = _hash_len16(y, (a + _rot((b + _K2) & _M64, 18) + c) & _M64, mul)
        e = (_f64(s, 16) * mul) & _M64
        f = _f64(s, 24)
        g = ((y + _f64(s, n - 32)) * mul) & _M64
        h = ((z + _f64(s, n - 24)) * mul) & _M64
        return _hash_len16((_rot((e + f) & _M64, 43) + _rot(g, 30) + h) & _M64,
                           (e + _rot((f + a) & _M64, 18) + g) & _M64, mul)
    seed = 81
    x = seed
    y = (seed * _K1 + 113) & _M64
    z = (_shift_mix((y * _K2 + 113) & _M64) * _K2) & _M64
    v, w = (0, 0), (0, 0)
    x = (x * _K2 + _f64(s, 0)) & _M64
    end = ((n - 1) // 64) * 64
    last64 = end + ((n - 1) & 63) - 63
    i = 0
    while True:
        x = (_rot((x + y + v[0] + _f64(s, i + 8)) & _M64, 37) * _K1) & _M64
        y = (_rot((y + v[1] + _f64(s, i + 48)) & _M64, 42) * _K1) & _M64
        x ^= w[1]
        y = (y + v[0] + _f64(s, i + 40)) & _M64
        z = (_rot((z + w[0]) & _M64, 33) * _K1) & _M64
        v = _weak_hash32(_f64(s, i), _f64(s, i + 8), _f64(s, i + 16), _f64(s, i + 24), (v[1] * _K1) & _M64,
                         (x + w[0]) & _M64)
        w = _weak_hash32(_f64(s, i + 32), _f64(s, i + 40), _f64(s, i + 48), _f64(s, i + 56), (z + w[1]) & _M64,
                         (y + _f64(s, i + 16)) & _M64)
        z, x = x, z
        i += 64
        if i == end:
            break
    mul = (_K1 + ((z & 0xFF) << 1)) & _M64
    i = last64
    w = ((w[0] + ((n - 1) & 63)) & _M64, w[1])
    v = ((v[0] + w[0]) & _M64, v[1])
    w = ((w[0] + v[0]) & _M64, w[1])
    x = (_rot((x + y + v[0] + _f64(s, i + 8)) & _M64, 37) * mul) & _M64
    y = (_rot((y + v[1] + _f64(s, i + 48)) & _M64, 42) * mul) & _M64
    x ^= (w[1] * 9) & _M64
    y = (y + v[0] * 9 + _f64(s, i + 40)) & _M64
    z = (_rot((z + w[0]) & _M64, 33) * mul) & _M64
    v = _weak_hash32(_f64(s, i), _f64(s, i + 8), _f64(s, i + 16), _f64(s, i + 24), (v[1] * mul) & _M64,
                     (x + w[0]) & _M64)
    w = _weak_hash32(_f64(s, i + 32), _f64(s, i + 40), _f64(s, i + 48), _f64(s, i + 56), (z + w[1]) & _M64,
                     (y + _f64(s, i + 16)) & _M64)
    z, x = x, z
    return _hash_len16((_hash_len16(v[0], w[0], mul) + _shift_mix(y) * _K0 + z) & _M64,
                       (_hash_len16(v[1], w[1], mul) + x) & _M64, mul)


# ----------------------------------------------------------------------------- id tables
class IdTable(object):
    """`index_table_from_tensor(mapping, num_oov_buckets=rows - len(mapping))` for one embedding table."""

    def __init__(self, name: str, vocab: List[str], rows: int):
        self.name, self.rows = name, int(rows)
        self.n_vocab = len(vocab)
        self.buckets = self.rows - self.n_vocab
        if self.buckets < 0:
            raise ValueError("table %s: vocabulary (%d) larger than the embedding rows (%d)" % (name, self.n_vocab, rows))
        # a later duplicate does not displace the first position (index_table_from_tensor fails on duplicates;
        # the shipped tables have none)
        self.index = {}
        for i, s in enumerate(vocab):
            self.index.setdefault(s.encode() if isinstance(s, str) else s, i)

    def lookup(self, values: List[bytes]) -> np.ndarray:
        out = np.empty(len(values), dtype=np.int32)
        idx, nv, nb = self.index, self.n_vocab, self.buckets
        for i, v in enumerate(values):
            j = idx.get(v)
            if j is None:
                j = nv + fingerprint64(v) % nb if nb > 0 else 0      # default_value=0 without buckets
            out[i] = j
        return out

    @staticmethod
    def read_vocab(path: str, name: str) -> List[str]:
        """The `ID_TABLES = {'<Name>': ['unknow', ...]}` literal of conf/idtables/<Name>.py, read with a regex
        instead of `import` (Sku.py is a 73 MB single-line literal)."""
        with open(path, "r") as fh:
            text = fh.read()
        start = text.index("[", text.index("'%s'" % name))
        return re.findall(r"'([^']*)'", text[start:text.rindex("]")])


class LookupTables(object):
    """data_feed/index_tables.py::LookupTables: one IdTable per `[embedding]` table name, one entry per feature."""

    def __init__(self, wnd_conf, idtables_dir: str, vocab_override: Optional[Dict[str, List[str]]] = None):
        self.tables: Dict[str, IdTable] = {}
        self.feature_table: Dict[str, IdTable] = {}
        for name, rows, dim, feature, side in list(wnd_conf.embedding_list) + list(wnd_conf.embedding_list_bias):
            if name not in self.tables:
                if vocab_override is not None and name in vocab_override:
                    vocab = vocab_override[name]
                else:
                    vocab = IdTable.read_vocab(os.path.join(idtables_dir, name + ".py"), name)
                self.tables[name] = IdTable(name, vocab, int(rows))
            self.feature_table.setdefault(feature, self.tables[name])

    def transform_id2index(self, feature: str, values: List[bytes]) -> np.ndarray:
        return self.feature_table[feature].lookup(values)


# ----------------------------------------------------------------------------- batches
class ExampleBatcher(object):
    """parse_single_line + transform_id2index + batching: serialized Examples -> the A0 feature dict."""

    def __init__(self, wnd_conf, tables: Optional[LookupTables]):
        from . import keys as K
        self.conf = wnd_conf
        self.tables = tables
        self.feature_dim = wnd_conf[K.MODEL][K.FEAT_DIM]
        self.weight_num = len(wnd_conf.train_weight) if hasattr(wnd_conf, "train_weight") else 5
        self.id_features = [e[3] for e in wnd_conf.embedding_list]
        for e in wnd_conf.embedding_list_bias:
            if e[3] not in self.id_features:
                self.id_features.append(e[3])

    def batch(self, payloads: List[bytes]) -> Dict[str, object]:
        B = len(payloads)
        labels = np.zeros(B, np.float32)
        mask = np.zeros((B, self.weight_num), np.float32)
        feats = np.zeros((B, self.feature_dim), np.float32)
        headers, pos, page = [], np.zeros(B, np.int32), np.zeros(B, np.int32)
        ids = {f: [] for f in self.id_features}
        wts = {f: [] for f in self.id_features}
        for b, payload in enumerate(payloads):
            ex = parse_example(payload)
            labels[b] = ex["label"][0]
            mask[b] = ex["mask"]
            feats[b] = ex["features"]
            header = ex["header"][0]
            headers.append(header)
            cols = header.split(b"\t")
            pos[b] = min(int(cols[4]), 400)          # tfrecord_mask.py:64-65
            page[b] = min(int(cols[11]), 100)        # tfrecord_mask.py:66-67
            for f in self.id_features:
                ids[f].append(ex.get(f, []))
                wts[f].append(np.asarray(ex.get(f + "Wts", np.zeros(0, np.float32)), dtype=np.float32))
        out: Dict[str, object] = {
            "features": torch.from_numpy(feats), "mask": torch.from_numpy(mask), "label": torch.from_numpy(labels),
            "em_position": torch.from_numpy(pos), "em_page": torch.from_numpy(page), "header": headers,
        }
        for f in self.id_features:
            lens = np.fromiter((len(r) for r in ids[f]), dtype=np.int64, count=B)
            off = np.zeros(B + 1, np.int32)
            off[1:] = np.cumsum(lens)
            flat = [v for r in ids[f] for v in r]
            if self.tables is not None:
                values = self.tables.transform_id2index(f, flat)
            else:
                values = np.zeros(len(flat), np.int32)
            w = np.concatenate(wts[f]) if len(flat) else np.zeros(0, np.float32)
            if w.size != len(flat):                 # a feature without Wts: unit weights (base.py:107-111)
                w = np.ones(len(flat), np.float32)
            out[f] = SparseIds(torch.from_numpy(values.astype(np.int32)), torch.from_numpy(off), torch.from_numpy(w))
        return share_offsets(out)       # the parallel id lists of one behaviour sequence: one offsets tensor


def list_files(prefix: str) -> List[str]:
    """`tf.data.Dataset.list_files(path + '*')` on a local path, sorted (the reference shuffles with seed 131)."""
    return sorted(p for p in glob.glob(prefix + "*") if os.path.isfile(p))


def batches(wnd_conf, tables, file_prefix: str, batch_size: int, epochs: int = 1, shuffle_size: int = 0,
            seed: int = 131, world: int = 1, rank: int = 0, drop_remainder: bool = False) -> Iterator[Dict]:
    """get_multi_towers_batch (tfrecord_mask.py:120-158): repeat -> shuffle buffer -> batch; data-parallel rank
    `rank` of `world` takes every world-th batch (each tower calls iterator.get_next() in turn, :152-157).
    With world > 1 all ranks yield the same number of (full) batches."""
    batcher = ExampleBatcher(wnd_conf, tables)
    rng = np.random.Generator(np.random.PCG64(seed))

    def stream():
        for _ in range(epochs):
            for path in list_files(file_prefix):
                for rec in read_records(path):
                    yield rec

    def shuffled():
        if shuffle_size <= 1:
            yield from stream()
            return
        buf = []
        for rec in stream():
            if len(buf) < shuffle_size:
                buf.append(rec)
                continue
            j = int(rng.integers(len(buf)))
            yield buf[j]
            buf[j] = rec
        rng.shuffle(buf)
        yield from buf

    # world > 1: every step of synchronous data-parallel training runs collectives on all ranks, so all ranks must
    # see the SAME number of steps: only complete groups of `world` FULL batches are handed out; the trailing
    # partial group and the remainder batch are dropped (the reference pulls all towers' batches of a step from
    # one iterator and abandons the whole step on OutOfRangeError, run_dnn.py:148-156,330-341).
    cur, n, mine = [], 0, None
    for rec in shuffled():
        cur.append(rec)
        if len(cur) == batch_size:
            if world == 1:
                yield batcher.batch(cur)
            else:
                if n % world == rank:
                    mine = cur
                if n % world == world - 1:          # the group is complete: every rank has its batch
                    yield batcher.batch(mine)
                    mine = None
            n += 1
            cur = []
    if cur and not drop_remainder and world == 1:
        yield batcher.batch(cur)
