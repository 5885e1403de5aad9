"""Batch container + seeded synthetic TFRecord-shaped batches.

The reference hands the model a dict of `tf.SparseTensor`s, one per id feature,
left-packed `[B, maxlen_in_batch]` (tfrecord_mask.py:23-84 + index_tables.py:37-45),
plus `features` fp32 `[B, 615]`, `mask` fp32 `[B, 5]`, `label` fp32 `[B]`.  A
left-packed SparseTensor is exactly a CSR row-partition, so the drop-in keeps the
same keys and carries each feature as `SparseIds(values int32 [nnz], offsets int32
[B+1])` (+ optional `weights`, the `<feature>Wts` tensor).

`synthetic_batch` follows SURVEY 8(d): seed 20201019, length/id/label
distributions measured on `jd_recsys_demo`.
"""
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

SEED = 20201019
LABEL_VALUES = (0, 1, 2, 4, 5)
LABEL_PRIOR = (0.9322, 0.0088, 0.0569, 0.0015, 0.00065)   # jd_recsys_demo/stat/stat/part-00000


class DevArray:
    """A 1-D device array known only by address: what the C ABI needs of an id / offset / weight array
    (`data_ptr()`, `numel()`, `dtype`, `device`).  `PackedBatch.to(..., views=False)` hands these out instead of
    torch views: staging a batch for inference then costs two small objects per feature instead of ~100 tensor
    views.  Not a tensor -- the training path (which sorts and indexes ids with torch) asks for real views."""
    __slots__ = ("_p", "_n", "dtype", "device")

    def __init__(self, ptr, numel, dtype, device):
        self._p, self._n, self.dtype, self.device = ptr, numel, dtype, device

    def data_ptr(self):
        return self._p

    def numel(self):
        return self._n


@dataclass
class SparseIds:
    values: torch.Tensor             # int32 [nnz], post-lookup index in [0, V)
    offsets: torch.Tensor            # int32 [B+1]
    weights: Optional[torch.Tensor] = None   # fp32 [nnz] (`<feature>Wts`), None == all ones

    @property
    def batch_size(self):
        return self.offsets.numel() - 1

    def to(self, device, non_blocking=False):
        return SparseIds(self.values.to(device, non_blocking=non_blocking),
                         self.offsets.to(device, non_blocking=non_blocking),
                         None if self.weights is None else self.weights.to(device, non_blocking=non_blocking))

    def lengths(self):
        return self.offsets[1:] - self.offsets[:-1]

    def to_dense(self, pad=0):
        """tf.sparse.to_dense of the left-packed tensor: [B, max_len]."""
        lens = self.lengths().long()
        B, T = lens.numel(), int(lens.max().item()) if lens.numel() else 0
        out = torch.full((B, T), pad, dtype=torch.int64, device=self.values.device)
        pos = torch.arange(T, device=self.values.device)[None, :]
        m = pos < lens[:, None]
        out[m] = self.values.long()
        return out

    @staticmethod
    def from_lists(rows, weights=None):
        lens = [len(r) for r in rows]
        off = np.zeros(len(rows) + 1, dtype=np.int32)
        off[1:] = np.cumsum(lens)
        vals = np.asarray([v for r in rows for v in r], dtype=np.int32)
        w = None
        if weights is not None:
            w = torch.from_numpy(np.asarray([v for r in weights for v in r], dtype=np.float32))
        return SparseIds(torch.from_numpy(vals), torch.from_numpy(off), w)


def batch_to(batch: Dict, device, non_blocking=False) -> Dict:
    """Move a batch; the longest row of every CSR feature whose offsets are still on the host is recorded under
    `__max_len__` (feature name -> length) so the model never needs a device->host read to size its row slots."""
    out = {}
    max_len = dict(batch.get("__max_len__") or {})
    moved = {}      # offsets tensors that share storage in the source (the id lists of one behaviour sequence) stay
                    # ONE tensor on the device: the pooled-lookup kernel walks such features together

    def move_offsets(t):
        key = (t.data_ptr(), t.dtype, tuple(t.shape))
        if key not in moved:
            moved[key] = t.to(device, non_blocking=non_blocking)
        return moved[key]

    for k, v in batch.items():
        if k == "__max_len__":
            continue
        if isinstance(v, SparseIds) and v.offsets.device.type == "cpu" and k not in max_len:
            off = v.offsets
            max_len[k] = int((off[1:] - off[:-1]).max()) if off.numel() > 1 else 0
        if isinstance(v, SparseIds):
            out[k] = SparseIds(v.values.to(device, non_blocking=non_blocking), move_offsets(v.offsets),
                               None if v.weights is None else v.weights.to(device, non_blocking=non_blocking))
        else:
            out[k] = v.to(device, non_blocking=non_blocking) if hasattr(v, "to") else v
    out["__max_len__"] = max_len
    return out


def share_offsets(batch: Dict) -> Dict:
    """CSR features of a HOST batch whose offsets arrays are equal (the parallel id lists of one behaviour sequence:
    sku / time bucket / category / brand / shop of the same events) are made to share ONE offsets tensor, in place.
    `PackedBatch` then stores it once and `batch_to` moves it once, and `dmt_pool_mean_fwd` walks those features
    together (it groups by offsets pointer, which proves equal lengths without reading device memory)."""
    seen = {}
    for k, v in batch.items():
        if isinstance(v, SparseIds) and v.offsets.device.type == "cpu":
            key = (v.offsets.numel(), v.offsets.dtype, v.offsets.contiguous().numpy().tobytes())
            if key in seen:
                v.offsets = seen[key]
            else:
                seen[key] = v.offsets
    return batch


def _seq_lengths(rng, B, max_len, kind):
    u = rng.random(B)
    if kind == "clk":
        lens = np.where(u < 0.45, max_len, rng.integers(1, max(max_len, 2), size=B))
    elif kind == "near":
        lens = np.where(u < 0.66, max_len, rng.integers(1, max(max_len, 2), size=B))
    else:
        mid = rng.integers(2, max(max_len, 3), size=B) if max_len > 2 else np.ones(B, dtype=np.int64)
        lens = np.where(u < 0.6, max_len, np.where(u < 0.75, 1, mid))
    return np.clip(lens, 1, max_len).astype(np.int64)


def _draw_ids(rng, n, V, mode, oov_rows=0):
    if V <= 2:
        return np.zeros(n, dtype=np.int32)
    if mode == "uniform":
        return rng.integers(1, V, size=n).astype(np.int32)
    # zipf over a fixed pseudo-random permutation of the in-vocab range, plus OOV buckets + index 0
    hi = max(V - oov_rows, 2)
    ranks = rng.zipf(1.05, size=n).astype(np.int64)
    ranks = np.minimum(ranks, hi - 1) - 1
    ids = 1 + (ranks * 2654435761 % (hi - 1))
    if oov_rows > 0:
        oov = rng.random(n) < 0.34
        ids = np.where(oov, rng.integers(V - oov_rows, V, size=n), ids)
    ids = np.where(rng.random(n) < 0.01, 0, ids)
    return ids.astype(np.int32)


def synthetic_batch(plan, batch_size, seed=SEED, id_mode="uniform", table_rows=None,
                    seq_lens=None, full_length=False, near_len=6) -> Dict:
    """One synthetic batch with the reference's feature names.

    table_rows: optional {table name: V} override (vocab sweep, small test tables).
    seq_lens:   optional per-sequence max length (defaults: the `_<N>` suffix of the
                feature name, capped at maxlen_k).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    B = batch_size
    rows = {n: t.rows for n, t in plan.tables.items()}
    rows.update({"bias:" + n: t.rows for n, t in plan.bias_tables.items()})
    for n, v in (table_rows or {}).items():
        if n in rows:
            rows[n] = v
        if "bias:" + n in rows:
            rows["bias:" + n] = min(rows["bias:" + n], v)
    batch: Dict = {}
    dense = rng.normal(0.0, 0.2, size=(B, plan.feature_dim)).clip(-0.99, 0.99)
    dense[rng.random(dense.shape) < 0.02] = 0.0
    batch["features"] = torch.from_numpy(dense.astype(np.float32))
    lab = rng.choice(len(LABEL_VALUES), size=B, p=np.asarray(LABEL_PRIOR) / sum(LABEL_PRIOR))
    mask = np.zeros((B, len(LABEL_VALUES)), dtype=np.float32)
    mask[np.arange(B), lab] = 1.0
    batch["mask"] = torch.from_numpy(mask)
    batch["label"] = torch.from_numpy(np.asarray(LABEL_VALUES, dtype=np.float32)[lab])

    feat_table = {p.feature: p.table for p in plan.pooled}
    done = set()
    kinds = ("clk", "ord", "cart")
    for seq in plan.sequences:
        tail = seq.user_features[0].rsplit("_", 1)[-1]
        L = int(tail) if tail.isdigit() else seq.maxlen
        if seq_lens is not None:
            L = seq_lens[seq.index]
        L = min(L, seq.maxlen)
        lens = np.full(B, L, dtype=np.int64) if full_length else \
            _seq_lengths(rng, B, L, kinds[seq.index] if seq.index < 3 else "ord")
        off = np.zeros(B + 1, dtype=np.int32)
        off[1:] = np.cumsum(lens)
        off_t = torch.from_numpy(off)
        n = int(off[-1])
        feats = list(seq.user_features) + ([seq.ts_feature] if seq.ts_feature else [])
        for f in feats:
            if f in done or f not in feat_table:
                continue
            V = rows[feat_table[f]]
            oov = 302 if feat_table[f] == "Sku" and V > 100000 else 0
            batch[f] = SparseIds(torch.from_numpy(_draw_ids(rng, n, V, id_mode, oov)), off_t)
            done.add(f)
    one = torch.arange(B + 1, dtype=torch.int32)
    for p in plan.pooled:
        if p.feature in done:
            continue
        if p.side == "i":
            batch[p.feature] = SparseIds(torch.from_numpy(_draw_ids(rng, B, rows[p.table], id_mode)), one)
        else:   # a pooled-only user feature that belongs to no sequence
            lens = _seq_lengths(rng, B, 10, "ord")
            off = np.zeros(B + 1, dtype=np.int32)
            off[1:] = np.cumsum(lens)
            batch[p.feature] = SparseIds(torch.from_numpy(_draw_ids(rng, int(off[-1]), rows[p.table], id_mode)),
                                         torch.from_numpy(off))
        done.add(p.feature)
    for p in plan.bias_pooled:
        if p.feature in done:
            continue
        lens = _seq_lengths(rng, B, near_len, "near")
        off = np.zeros(B + 1, dtype=np.int32)
        off[1:] = np.cumsum(lens)
        V = min(rows["bias:" + p.table], rows.get(p.table, 1 << 60))
        batch[p.feature] = SparseIds(torch.from_numpy(_draw_ids(rng, int(off[-1]), V, id_mode)),
                                     torch.from_numpy(off))
        done.add(p.feature)
    return batch


def batch_tokens(plan, batch) -> int:
    """Valid tokens across the behaviour sequences (the unit of the gather roofline)."""
    n = 0
    for seq in plan.sequences:
        n += int(batch[seq.user_features[-1]].offsets[-1])
    return n


class PackedBatch:
    """A batch serialised into ONE pinned host buffer (what a data-loader worker hands over), so
    the host->device step is a single async copy.  Arrays are 256-byte aligned; tensors that
    share storage in the source batch (the offsets of the features of one sequence) are stored
    once."""

    ALIGN = 256

    def __init__(self, batch: Dict, pin=True, compact=False, keys=None):
        """compact: ship id arrays in the narrowest byte width that holds their ids -- 1 byte (time buckets), 2 bytes
        (category vocabularies < 65536), 3 bytes (Sku / Brand / Shopid: < 2^24) -- and
        fp32 `features` as bf16 -- for the bf16 tensor-core path, which rounds the features to bf16 before its
        first GEMM anyway.  `to()` widens the ids back to int32 on the device (dmt_widen_ids, one launch per
        batch) and hands `features` over as a bf16 tensor.  keys: only these entries are packed (an inference
        batch needs neither `label` nor the propensity arrays)."""
        self.layout = []          # (key, kind, dtype, shape, byte offset); kind in {t, v, o, w}
        self.narrow = {}          # byte offset of a narrow id array -> (byte offset in the wide buffer, n, bytes per id)
        self.wide_bytes = 0
        self.compact = bool(compact)
        blobs, seen, off = [], {}, 0

        def add(key, kind, t):
            nonlocal off
            t = t.contiguous()
            ident = (t.data_ptr(), t.dtype, tuple(t.shape))
            if ident in seen and t.numel() > 0:
                self.layout.append((key, kind, seen[ident][1], tuple(t.shape), seen[ident][0]))
                return
            stored, logical = t, t.dtype
            if compact and kind == "v" and t.dtype == torch.int32 and t.numel() > 0 and \
                    int(t.min()) >= 0 and int(t.max()) < (1 << 24):
                # the narrowest little-endian width that holds every id: 1 byte (time buckets), 2 (categories),
                # 3 (Sku / Brand / Shopid: vocabularies < 2^24)
                top = int(t.max())
                nb = 1 if top < 256 else (2 if top < 65536 else 3)
                raw = np.ascontiguousarray(t.numpy().astype("<u4")).view(np.uint8).reshape(-1, 4)
                stored = torch.from_numpy(np.ascontiguousarray(raw[:, :nb]).reshape(-1))
                self.narrow[off] = (self.wide_bytes, t.numel(), nb)
                self.wide_bytes += (t.numel() * 4 + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            elif compact and kind == "t" and key == "features" and t.dtype == torch.float32:
                stored = t.to(torch.bfloat16)
                logical = torch.bfloat16
            seen[ident] = (off, logical)
            self.layout.append((key, kind, logical, tuple(t.shape), off))
            blobs.append((off, stored))
            off += (stored.numel() * stored.element_size() + self.ALIGN - 1) // self.ALIGN * self.ALIGN

        for k, v in batch.items():
            if keys is not None and k not in keys:
                continue
            if isinstance(v, SparseIds):
                add(k, "v", v.values)
                add(k, "o", v.offsets)
                if v.weights is not None:
                    add(k, "w", v.weights)
            elif torch.is_tensor(v):
                add(k, "t", v)
        self.nbytes = max(off, self.ALIGN)
        self.host = torch.empty(self.nbytes, dtype=torch.uint8, pin_memory=pin and torch.cuda.is_available())
        for o, t in blobs:
            n = t.numel() * t.element_size()
            if n:
                self.host[o:o + n].copy_(t.view(-1).view(torch.uint8))
        self._widen_cache = {}

    def max_len(self, plan) -> Dict:
        """{sequence index: longest sequence in this batch}, computed once from the host copy."""
        if not hasattr(self, "_max_len"):
            host = self.unpack(self.host)
            self._max_len = {}
            for seq in plan.sequences:
                off = host[seq.user_features[-1]].offsets
                self._max_len[seq.index] = int((off[1:] - off[:-1]).max()) if off.numel() > 1 else 0
        return self._max_len

    def _fast_plan(self):
        """All entries are 4-byte types: view the buffer once as int32 and once as float32 and cut both with ONE
        split_with_sizes each (entry, padding, entry, padding, ...): two tensor ops instead of three per entry."""
        if not hasattr(self, "_fast"):
            ok = all(dt in (torch.int32, torch.float32) for _, _, dt, _, _ in self.layout) and not self.narrow
            uniq = sorted({o for _, _, _, _, o in self.layout})
            sizes, index, pos = [], {}, 0
            if ok:
                ends = {}
                for _, _, _, shape, o in self.layout:
                    n = 1
                    for d in shape:
                        n *= d
                    ends[o] = max(ends.get(o, 0), n)
                for o in uniq:
                    if o % 4 or o < pos * 4:
                        ok = False
                        break
                    if o // 4 > pos:
                        sizes.append(o // 4 - pos)      # padding
                    index[o] = len(sizes)
                    sizes.append(ends[o])
                    pos = o // 4 + ends[o]
                total = self.nbytes // 4
                if ok and total >= pos:
                    if total > pos:
                        sizes.append(total - pos)
                else:
                    ok = False
            self._fast = (sizes, index) if ok else None
        return self._fast

    def unpack(self, buf: torch.Tensor, wide: Optional[torch.Tensor] = None) -> Dict:
        """Views over `buf` (a uint8 tensor holding a copy of `self.host`, on any device).  narrow id arrays
        of a compact batch are views of `wide` (the buffer `to()` widened them into) or, without it (host side),
        converted copies."""
        out, parts = {}, {}
        fast = self._fast_plan()
        if fast is not None and buf.numel() >= self.nbytes:
            sizes, index = fast
            b8 = buf[:self.nbytes]
            seg_i = b8.view(torch.int32).split_with_sizes(sizes)
            seg_f = b8.view(torch.float32).split_with_sizes(sizes)
            for key, kind, dtype, shape, o in self.layout:
                t = (seg_i if dtype == torch.int32 else seg_f)[index[o]]
                n = 1
                for d in shape:
                    n *= d
                if t.numel() != n:
                    t = t[:n]
                if len(shape) != 1:
                    t = t.view(shape)
                if kind == "t":
                    out[key] = t
                else:
                    parts.setdefault(key, {})[kind] = t
        else:
            for key, kind, dtype, shape, o in self.layout:
                n = 1
                for d in shape:
                    n *= d
                if o in self.narrow and kind == "v":
                    wo, _, nb = self.narrow[o]
                    if wide is not None:
                        t = wide[wo:wo + 4 * n].view(torch.int32).view(shape)
                    else:       # host side: bytes -> int32
                        b = buf[o:o + nb * n].cpu().numpy().reshape(-1, nb).astype(np.uint32)
                        v = sum(b[:, k] << (8 * k) for k in range(nb)) if n else np.zeros(0, np.uint32)
                        t = torch.from_numpy(v.astype(np.int32)).view(shape)
                else:
                    nb = n * torch.empty((), dtype=dtype).element_size()
                    t = buf[o:o + nb].view(dtype).view(shape)
                if kind == "t":
                    out[key] = t
                else:
                    parts.setdefault(key, {})[kind] = t
        for key, p in parts.items():
            out[key] = SparseIds(p["v"], p["o"], p.get("w"))
        return out

    def feature_offsets(self, names):
        """[len(names), 3] int64 byte offsets of (ids, offsets, weights) of the CSR features `names` inside the packed
        buffer -- ids of a narrow-stored array: inside the WIDE buffer, flagged by bit 62; no weights: -1 -- plus the
        element counts [len(names), 2] (ids, offsets).  Computed once per batch when the data loader builds it; the
        native forward driver (dmt_forward_bf16) turns it into its pointer table with one vector add per call."""
        key = tuple(names)
        cache = self.__dict__.setdefault("_feat_offsets", {})
        hit = cache.get(key)
        if hit is None:
            if not hasattr(self, "_ptr_plan"):
                self.unpack_ptrs(self.host)      # (builds the plan; the descriptors it returns are dropped)
            sparse = {k: (v, o, w) for k, v, o, w in self._ptr_plan[1]}
            tab = np.full((len(names), 3), -1, dtype=np.int64)
            cnt = np.zeros((len(names), 2), dtype=np.int64)
            for i, n in enumerate(names):
                if n not in sparse:
                    raise KeyError("feature %r is not in this packed batch" % n)
                v, o, w = sparse[n]
                if v[2] != torch.int32 or o[2] != torch.int32:
                    raise TypeError("feature %r: ids/offsets must be int32" % n)
                tab[i, 0] = v[0] | ((1 << 62) if v[3] else 0)
                tab[i, 1] = o[0]
                if w is not None:
                    tab[i, 2] = w[0]
                cnt[i] = (v[1], o[1])
            hit = cache[key] = (tab, cnt)
        return hit

    def unpack_ptrs(self, buf: torch.Tensor, wide: Optional[torch.Tensor] = None) -> Dict:
        """Like `unpack`, but the id / offset / weight arrays come back as `DevArray`s (address + length) and only
        the dense tensors ('features', 'mask', ...) as torch views."""
        if not hasattr(self, "_ptr_plan"):
            dense, sparse = [], {}
            for key, kind, dtype, shape, o in self.layout:
                n = 1
                for d in shape:
                    n *= d
                if kind == "t":
                    dense.append((key, dtype, shape, o, n * (4 if dtype in (torch.int32, torch.float32) else
                                                            torch.empty((), dtype=dtype).element_size())))
                elif kind == "v" and o in self.narrow:
                    sparse.setdefault(key, {})[kind] = (self.narrow[o][0], n, dtype, True)
                else:
                    sparse.setdefault(key, {})[kind] = (o, n, dtype, False)
            self._ptr_plan = (dense, [(k, p["v"], p["o"], p.get("w")) for k, p in sparse.items()])
        dense, sparse = self._ptr_plan
        base, dev = buf.data_ptr(), buf.device
        wbase = wide.data_ptr() if wide is not None else 0
        out = {}
        for key, dtype, shape, o, nb in dense:
            out[key] = buf[o:o + nb].view(dtype).view(shape)
        for key, v, off, w in sparse:
            out[key] = SparseIds(DevArray((wbase if v[3] else base) + v[0], v[1], v[2], dev),
                                 DevArray(base + off[0], off[1], off[2], dev),
                                 None if w is None else DevArray(base + w[0], w[1], w[2], dev))
        out["__buffer__"] = (buf, wide)  # keeps the storage alive as long as the descriptors
        out["__packed__"] = self
        return out

    def to(self, device, out: Optional[torch.Tensor] = None, views: bool = True,
           wide: Optional[torch.Tensor] = None, stream: Optional[int] = None) -> Dict:
        """One async copy of the pinned buffer into `out` (+, for a compact batch, ONE dmt_widen_ids launch that
        restores the int32 id arrays into `wide`); both are enqueued on the current stream (`stream` = its raw
        handle, looked up when omitted)."""
        if out is None:
            out = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        out[:self.nbytes].copy_(self.host, non_blocking=True)
        if self.narrow:
            if wide is None:
                wide = torch.empty(self.wide_bytes, dtype=torch.uint8, device=device)
            self.widen(out, wide, stream)
        return self.unpack(out, wide) if views else self.unpack_ptrs(out, wide)

    def widen(self, buf: torch.Tensor, wide: torch.Tensor, stream: Optional[int] = None):
        """1- / 2- / 3-byte id arrays of `buf` -> int32 arrays in `wide` (C-ABI kernel; there is no host fallback)."""
        from . import abi
        if wide.numel() < self.wide_bytes:
            raise ValueError("wide buffer holds %d bytes, the batch needs %d" % (wide.numel(), self.wide_bytes))
        key = (buf.data_ptr(), wide.data_ptr())
        descs = self._widen_cache.get(key)
        if descs is None:
            items = sorted(self.narrow.items())
            descs = (abi.WidenIdsDesc * len(items))()
            for i, (o, (wo, n, nb)) in enumerate(items):
                descs[i].src, descs[i].dst, descs[i].n, descs[i].bytes = key[0] + o, key[1] + wo, n, nb
            if len(self._widen_cache) > 8:
                self._widen_cache.clear()
            self._widen_cache[key] = descs
        if stream is None:
            stream = torch.cuda.current_stream(buf.device).cuda_stream
        abi.check(abi.load().dmt_widen_ids(len(descs), descs, stream))
