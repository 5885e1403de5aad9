"""Host side of the DMT model plugin: same protocol as the reference's
`model/net/mmoe_transformer_unbias.py` (class of the same name, `__init__(wnd_conf)`,
`inference(inputs, is_train, is_predict)` :293-316), every operator executed by the
hand-written CUDA library behind `include/dmt_b200.h`.

This module only marshals pointers: it owns the device buffers (PyTorch tensors are the
allocator), fills the POD descriptors and enqueues the C-ABI calls on the current CUDA
stream.  There is no PyTorch math on the path and no CPU fallback.
"""
import ctypes as C
import os

import numpy as np
import torch

from .. import abi
from ..data import PackedBatch, SparseIds
from ..params import ParamStore, TASK_NAMES
from ..plan import build_plan


def _is_host(t):
    return t.device.type == "cpu"


class mmoe_transformer_unbias(object):
    def __init__(self, wnd_conf, device=None, params=None, precision="f32", seed=20201019, train_gemm=None):
        """precision: 'f32' (CUDA cores, exact-parity path) | 'bf16' (fused tcgen05 tile kernels; d_model 64 / 2 heads)
        | 'tf32' (any d_model / d_ff that are multiples of 16, e.g. dmt.conf's 80 / 320 with 4 heads: the row-batched
        pipeline -- every dense projection, the feed-forward and the MMoE experts on the TMA-fed tcgen05 kind::tf32
        engine straight from fp32 activations, attention / LayerNorm on CUDA cores in fp32).
        train_gemm: engine of the GEMMs of the TRAINING path -- 'f32' (SIMT), 'bf16' (tcgen05, bf16 operands),
        'bf16x3' (tcgen05, split hi+lo operands: fp32-grade) or 'tf32' (the per-token GEMMs of the sequence pipeline on
        the TMA-fed tcgen05 kind::tf32 engine, no operand conversion pass; MMoE on bf16x3).
        Default: 'f32' with precision 'f32', else 'bf16x3'."""
        self.wnd_conf = wnd_conf
        self.plan = wnd_conf if hasattr(wnd_conf, "mmoe_in") else build_plan(wnd_conf)
        if not torch.cuda.is_available():
            raise RuntimeError("dmt_b200 needs a CUDA device: the product path has no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = abi.load()
        self.precision = {"f32": abi.PRECISION_F32, "bf16": abi.PRECISION_BF16, "tf32": abi.PRECISION_TF32}[precision]
        if train_gemm is None:
            train_gemm = {"f32": "f32", "tf32": "tf32"}.get(precision, "bf16x3")
        self.train_precision = {"f32": abi.PRECISION_F32, "bf16": abi.PRECISION_BF16,
                                "bf16x3": abi.PRECISION_BF16X3, "tf32": abi.PRECISION_TF32}[train_gemm]
        self.params = params if params is not None else ParamStore(self.plan, device=self.device, seed=seed)
        pdev = self.params.dense.device
        if pdev.type != "cuda" or pdev.index != self.device.index:
            raise ValueError("ParamStore lives on %s, model on %s" % (self.params.device, self.device))
        self._buffers = {}
        self._prepared = {}          # seq index -> (params version, bf16 weight images)
        self.params_version = 0      # bump (invalidate_prepared) whenever the parameters change
        self._events = None          # bench hook: {stage: [(start, stop), ...]} CUDA events
        self.launches = 0            # kernels of this library enqueued so far
        self._stream_h = None        # set for the duration of inference() / compute_gradients()
        self.seq_streams = os.environ.get("DMT_SEQ_STREAMS", "1") != "0"   # one stream per behaviour sequence
        self.seq_multi = os.environ.get("DMT_SEQ_MULTI", "1") != "0"       # bf16: one launch over all sequences
        self.x_bf16 = os.environ.get("DMT_X_BF16", "1") != "0"             # bf16: MMoE input assembled in bf16
        self.fwd_native = os.environ.get("DMT_FWD_NATIVE", "1") != "0"     # bf16: one C call per forward
        self._fwd_states = {}
        self._pool_static = {}       # (bias, n specs) -> per-feature static descriptor parts
        self._v2_ok = True           # bf16 path: the decoder tails of all sequences run as one deferred launch
        self._bind_weights()

    # ------------------------------------------------------------------ per-stage device timing
    def enable_stage_timing(self, on=True):
        self._events = {} if on else None

    class _Stage(object):
        def __init__(self, model, name, launches):
            self.m, self.name, self.n = model, name, launches

        def __enter__(self):
            self.m.launches += self.n
            if self.m._events is not None:
                self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                self.ev[0].record(torch.cuda.current_stream(self.m.device))

        def __exit__(self, *exc):
            if self.m._events is not None:
                self.ev[1].record(torch.cuda.current_stream(self.m.device))
                self.m._events.setdefault(self.name, []).append(self.ev)
            return False

    def stage_times_ms(self):
        """{stage: (total ms, launches)} -- call after a synchronize."""
        out = {}
        for name, evs in (self._events or {}).items():
            out[name] = (sum(a.elapsed_time(b) for a, b in evs), len(evs))
        return out

    # ------------------------------------------------------------------ weight descriptors
    def _bind(self, P):
        """Fill the POD descriptors from `P[name] -> tensor` (the parameters, or a gradient buffer with the
        same layout)."""
        plan = self.plan
        seq_w = []
        for seq in plan.sequences:
            w = abi.SeqWeights()
            S = seq.scope
            if plan.position_encoding_method == "position_sin_cos":
                # a constant of the graph (TransformerModel_util.py:237-278): no variable, its "gradient" lands in a
                # scratch table that nothing reads
                w.pos = abi.ptr(self.params.position_table(seq) if P is self.params else
                                self._buf("pos_grad_sink_%d" % seq.index, (plan.maxlen_k, plan.d_model)))
            else:
                w.pos = abi.ptr(P[S + "/positional_encoding_k_position_learn/embedding_position_learn"])

            def attn(dst, base):
                dst.q = abi.dense(P[base + "/dense/kernel"], P[base + "/dense/bias"])
                dst.k = abi.dense(P[base + "/dense_1/kernel"], P[base + "/dense_1/bias"])
                dst.v = abi.dense(P[base + "/dense_2/kernel"], P[base + "/dense_2/bias"])
                dst.ln = abi.LayerNorm(abi.ptr(P[base + "/ln/gamma"]), abi.ptr(P[base + "/ln/beta"]))

            for b in range(plan.num_blocks_encode):
                attn(w.enc_attn[b], "%s/num_blocks_%d/self-attention" % (S, b))
            for b in range(plan.num_blocks_decode):
                attn(w.dec_attn[b], "%s/num_blocks_%d/vanilla_attention" % (S, b))
            for b in range(max(plan.num_blocks_encode, plan.num_blocks_decode)):
                base = "%s/num_blocks_%d/positionwise_feedforward" % (S, b)
                w.ff[b].w1 = abi.dense(P[base + "/dense/kernel"], P[base + "/dense/bias"])
                w.ff[b].w2 = abi.dense(P[base + "/dense_1/kernel"], P[base + "/dense_1/bias"])
                w.ff[b].ln = abi.LayerNorm(abi.ptr(P[base + "/ln/gamma"]), abi.ptr(P[base + "/ln/beta"]))
            seq_w.append(w)

        mw = abi.MmoeWeights()
        for e in range(plan.num_experts):
            for l in range(len(plan.hidden_units_bottom)):
                base = "DnnModel/mmoe_layers/expert-%d/expert-layer-%d" % (e, l)
                mw.expert[e][l] = abi.dense(P[base + "/weights"], P[base + "/biases"])
        for t in range(plan.num_tasks):
            base = "DnnModel/mmoe_layers/gates-%d/gates-layer-0" % t
            mw.gate[t] = abi.dense(P[base + "/weights"], P[base + "/biases"])
            name = TASK_NAMES[t]
            for l in range(len(plan.hidden_units_task)):
                base = "DnnModel/%s/%s-fc-%d" % (name, name, l)
                mw.tower[t][l] = abi.dense(P[base + "/weights"], P[base + "/biases"])
            base = "DnnModel/%s/%s-output" % (name, name)
            mw.tower_out[t] = abi.dense(P[base + "/weights"], P[base + "/biases"])

        bw = abi.BiasWeights()
        for l in range(len(plan.hidden_units_bias) + 1):
            bw.layer[l] = abi.dense(P["DnnModel/layer_bias%d/kernel" % l], P["DnnModel/layer_bias%d/bias" % l])
        return seq_w, mw, bw

    def _bind_weights(self):
        self._seq_w, self._mmoe_w, self._bias_w = self._bind(self.params)

    # ------------------------------------------------------------------ buffers
    def _buf(self, name, shape, dtype=torch.float32):
        key = (name, tuple(shape), dtype)
        t = self._buffers.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._buffers[key] = t
        return t

    def _scratch(self, name, nbytes):
        """Grow-only byte buffer (sizes that depend on the number of tokens change every batch)."""
        t = self._buffers.get(("scratch", name))
        if t is None or t.numel() < nbytes:
            t = torch.empty((int(nbytes * 1.25) + 4095) // 4096 * 4096, dtype=torch.uint8, device=self.device)
            self._buffers[("scratch", name)] = t
        return t

    def _seq_side_streams(self, n):
        st = getattr(self, "_side_streams", None)
        if st is None or len(st) < n:
            # stream 0 carries the (whole-SM, persistent) sequence kernels: highest priority, see csrc/forward.cu
            st = self._side_streams = [torch.cuda.Stream(self.device, priority=-1 if i == 0 else 0) for i in range(n)]
        return st

    def _ev_pool(self, name):
        """Reusable (timing-free) events for the fork / join of the side streams."""
        pool = getattr(self, "_events_pool", None)
        if pool is None:
            pool = self._events_pool = {}
        ev = pool.get(name)
        if ev is None:
            ev = pool[name] = torch.cuda.Event()
        return ev

    def _stream(self):
        """Raw handle of the current CUDA stream; looked up once per public entry point (the torch call costs
        ~5 us and every stage needs it)."""
        h = self._stream_h
        if h is None:
            h = torch.cuda.current_stream(self.device).cuda_stream
        return h

    def _dev(self, t):
        if t is None:
            return None
        return t.to(self.device, non_blocking=True) if _is_host(t) else t

    def stage_inputs(self, inputs):
        """Bring a batch to the device exactly once.  A `PackedBatch` is one pinned buffer -> one
        async copy; a plain dict is copied tensor by tensor (tensors shared between features, e.g.
        the offsets of one sequence, are copied once)."""
        if isinstance(inputs, PackedBatch):
            buf = self._scratch("packed_in", inputs.nbytes)
            wide = self._scratch("packed_wide", inputs.wide_bytes) if inputs.narrow else None
            out = inputs.to(self.device, out=buf, wide=wide)
            self.launches += 1 if inputs.narrow else 0          # dmt_widen_u16
            out["__max_len__"] = inputs.max_len(self.plan)
            return out
        if inputs.get("__staged__") is self:
            ready = inputs.get("__ready__")
            if ready is not None:     # produced by prefetch(): order the consumer after the copy
                torch.cuda.current_stream(self.device).wait_event(ready)
                self._pf_busy[inputs["__slot__"]] = False      # its consumer is enqueued: the slot may rotate
            return inputs
        cache, out = {}, {"__staged__": self}

        def dev(t):
            if t is None or not _is_host(t):
                return t
            key = (t.data_ptr(), t.dtype, tuple(t.shape))
            if key not in cache:
                cache[key] = t.to(self.device, non_blocking=True)
            return cache[key]

        for k, v in inputs.items():
            if isinstance(v, SparseIds):
                out[k] = SparseIds(dev(v.values), dev(v.offsets), dev(v.weights))
            elif torch.is_tensor(v):
                out[k] = dev(v)
            else:
                out[k] = v
        return out

    def prefetch(self, packed, views=True):
        """Start the host->device copy of a `PackedBatch` on a side stream and return the staged batch; pass
        it to `inference` / `compute_gradients` later.  views=False (inference only): the id arrays are handed
        over as raw `DevArray` descriptors instead of torch views -- the cheapest way to stage a batch.  The copy overlaps whatever the compute stream is doing
        (what a data-loader prefetch thread does for the reference's tf.data pipeline,
        tfrecord_mask.py:140-157).  Three rotating device buffers: a buffer is rewritten only after the
        compute enqueued before this call -- which includes its previous consumer -- has finished; prefetching
        a third batch while two are still unconsumed would overwrite one of them and raises instead."""
        if not isinstance(packed, PackedBatch):
            raise TypeError("prefetch takes a PackedBatch (one pinned host buffer)")
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._pf_slot = 0
            self._pf_busy = [False, False, False]
        cur = torch.cuda.current_stream(self.device)
        nxt = (self._pf_slot + 1) % 3
        if self._pf_busy[nxt]:
            raise RuntimeError("prefetch: the batch staged 3 calls ago has not been consumed yet (pass it to "
                               "inference / compute_gradients first); only 3 device buffers rotate")
        self._pf_slot = nxt
        self._pf_busy[nxt] = True
        buf = self._scratch("prefetch_%d" % self._pf_slot, packed.nbytes)
        buf.record_stream(self._copy_stream)
        wide = None
        if packed.narrow:       # compact batch: the uint16 id arrays are widened on the copy stream as well
            wide = self._scratch("prefetch_wide_%d" % self._pf_slot, packed.wide_bytes)
            wide.record_stream(self._copy_stream)
            self.launches += 1
        done = torch.cuda.Event()
        done.record(cur)
        self._copy_stream.wait_event(done)
        with torch.cuda.stream(self._copy_stream):
            out = packed.to(self.device, out=buf, views=views, wide=wide)
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        out["__max_len__"] = packed.max_len(self.plan)
        out["__staged__"] = self
        out["__ready__"] = ready
        out["__slot__"] = self._pf_slot
        return out

    def _sparse(self, inputs, name, role="pool"):
        """`role` distinguishes the lookups of one feature: 'pool' (embedding_combiner, row = id) and 'seq<i>'
        (generate_data of sequence i, row = id-1).  A row-sharded table re-maps them to different compact rows
        (`inputs['__remap__'][(role, name)]`, see shard.py / train.py); otherwise all read `inputs[name]`."""
        remap = inputs.get("__remap__")
        cache = inputs.get("__sparse__")
        if cache is None:
            cache = inputs["__sparse__"] = {}
        ckey = (role, name) if remap else name
        sp = remap.get((role, name)) if remap else None
        if sp is None:
            sp = inputs[name]
        hit = cache.get(ckey)
        if hit is not None and hit[0] is sp and hit[1] is inputs.get(name + "Wts"):
            return hit[2]
        if not isinstance(sp, SparseIds):
            raise TypeError("feature %r must be a SparseIds (CSR) value" % name)
        w = sp.weights
        wts = inputs.get(name + "Wts")
        if wts is not None:
            w = wts.values if isinstance(wts, SparseIds) else wts
            if w.dtype != torch.float32:
                w = w.float()
        out = SparseIds(self._dev(sp.values), self._dev(sp.offsets), self._dev(w))
        if out.values.dtype != torch.int32 or out.offsets.dtype != torch.int32:
            raise TypeError("feature %r: ids/offsets must be int32" % name)
        cache[ckey] = (sp, wts, out)
        return out

    # ------------------------------------------------------------------ forward pieces
    def invalidate_prepared(self):
        """Call after the parameters changed (optimizer step, checkpoint load)."""
        self.params_version += 1

    def _copy_dense(self, feats, batch, x, x_ld, keep, precision):
        """base.py:95-96: the dense `features` block into the first columns of the MMoE input.  fp32 [B, F], or --
        bf16 tensor-core path only -- the bf16 block of a compact PackedBatch (that path rounds the features to
        bf16 before its first GEMM anyway)."""
        plan, lib = self.plan, self.lib
        if tuple(feats.shape) != (batch, plan.feature_dim) or feats.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("'features' must be fp32 (or bf16) [%d, %d]" % (batch, plan.feature_dim))
        if feats.dtype == torch.bfloat16 and precision != abi.PRECISION_BF16:
            raise ValueError("bf16 'features' (compact batch) are only accepted by the bf16 inference path")
        feats = feats.contiguous()
        fn = lib.dmt_copy_dense_features if feats.dtype == torch.float32 else lib.dmt_copy_dense_features_bf16
        with self._Stage(self, "copy_dense", 1):
            abi.check(fn(feats.data_ptr(), batch, plan.feature_dim, x.data_ptr(), x_ld, self._stream()))
        keep.append(feats)

    def _stage_dense_bf16(self, feats, batch, xb, xb_ld, keep):
        """base.py:95-96 for the bf16 MMoE input: fp32 or bf16 `features` -> bf16 columns [0, feature_dim)."""
        plan = self.plan
        if tuple(feats.shape) != (batch, plan.feature_dim) or feats.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("'features' must be fp32 (or bf16) [%d, %d]" % (batch, plan.feature_dim))
        feats = feats.contiguous()
        with self._Stage(self, "copy_dense", 1):
            abi.check(self.lib.dmt_stage_dense_features_bf16(feats.data_ptr(), 1 if feats.dtype == torch.bfloat16 else 0,
                                                             batch, plan.feature_dim, xb.data_ptr(), xb_ld,
                                                             self._stream()))
        keep.append(feats)

    def _seq_len_hint(self, inputs, seq):
        """Upper bound on this sequence's lengths (selects the row-slot size of the bf16 tile kernels, which clamp
        every sequence to it).  Exact when known: `inputs['__max_len__']` ({sequence index | feature name: longest
        sequence}, filled by `PackedBatch` / `batch_to` from the host copy), or computed here when the offsets
        are still host tensors.  Otherwise transformer_maxlen_k -- the feature-name suffix (`..._12m_10`) is NOT
        a bound: the reference sizes the sequence by the batch's own longest row
        (mmoe_transformer_unbias.py:141-146)."""
        return self._seq_len_bound(inputs, seq)[0]

    def _seq_len_bound(self, inputs, seq):
        """(bound, exact): `exact` = the bound is the batch's true longest sequence and it does not exceed
        transformer_maxlen_k (DMT_SEQ_LEN_EXACT)."""
        name = seq.user_features[-1]
        hints = inputs.get("__max_len__")
        hint = None
        if isinstance(hints, dict):
            hint = hints.get(seq.index, hints.get(name))
        if hint is None:
            sp = inputs.get(name)
            off = getattr(sp, "offsets", None)
            if torch.is_tensor(off) and _is_host(off) and off.numel() > 1:
                hint = int((off[1:] - off[:-1]).max())
        exact = hint is not None and int(hint) <= seq.maxlen
        if hint is None:
            hint = seq.maxlen
        return max(1, min(int(hint), seq.maxlen)), exact

    def _prepared_for(self, seq_index, cfg):
        ver, buf = self._prepared.get(seq_index, (-1, None))
        nbytes = self.lib.dmt_seq_encode_workspace_bytes(C.byref(cfg), 0)
        if buf is None or buf.numel() < nbytes:     # the workspace also holds batch-sized scratch: grow-only
            buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            ver = -1
        nbytes = buf.numel()
        if ver != self.params_version:
            stream = self._stream()
            with self._Stage(self, "prepare_weights", 1):
                abi.check(self.lib.dmt_seq_prepare_weights(C.byref(cfg), C.byref(self._seq_w[seq_index]),
                                                           buf.data_ptr(), nbytes, stream))
            self._prepared[seq_index] = (self.params_version, buf)
        return buf, nbytes

    def _seq_cfg(self, inputs, seq, batch, precision, dropout_rate=0.0, dropout_seed=0):
        plan = self.plan
        bound, exact = self._seq_len_bound(inputs, seq)
        return abi.SeqCfg(batch, plan.d_model, plan.d_ff, plan.num_heads, plan.num_blocks_encode,
                          plan.num_blocks_decode, plan.maxlen_k, 1 if plan.zero_pad else 0,
                          len(seq.user_features), precision, bound, abi.SEQ_LEN_EXACT if exact else 0,
                          float(dropout_rate), int(dropout_seed) & 0xFFFFFFFF)

    def _seq_input(self, inputs, seq, batch):
        si = abi.SeqInput()
        keep = []
        for f, (uf, itf) in enumerate(zip(seq.user_features, seq.item_features)):
            table = self.params.table(seq.tables[f])
            u = self._sparse(inputs, uf, "seq%d" % seq.index)
            it = self._sparse(inputs, itf, "seq%d" % seq.index)
            if u.offsets.numel() != batch + 1:
                raise ValueError("feature %r: offsets has %d entries, batch is %d" % (uf, u.offsets.numel(), batch))
            if it.values.numel() != batch:
                # the reference uses the flat .values of the item feature (:158): one id per sample
                raise ValueError("item feature %r must hold exactly one id per sample" % itf)
            keep += [u, it]
            si.table[f] = abi.ptr(table)
            si.rows[f] = table.shape[0]
            si.dim[f] = table.shape[1]
            si.ids[f] = abi.ptr(u.values)
            si.offsets[f] = abi.ptr(u.offsets)
            si.item_ids[f] = abi.ptr(it.values)
        return si, keep

    def seq_encode(self, inputs, seq_index, out, out_ld, batch, deferred=None):
        """A2-A8 for one behaviour sequence; writes [B, d_model] at `out` (row stride out_ld).
        deferred: a list -> bf16 path only: the decoder tail of this sequence is left to `seq_tails(deferred)`,
        which runs the tails of all sequences as one launch."""
        plan = self.plan
        seq = plan.sequences[seq_index]
        cfg = self._seq_cfg(inputs, seq, batch, self.precision)
        if self.precision == abi.PRECISION_TF32:      # row-batched pipeline (its intermediates live in `saved`)
            si, keep = self._seq_input(inputs, seq, batch)
            n_tok = self._sparse(inputs, seq.user_features[-1], "seq%d" % seq.index).values.numel()
            saved = self._scratch("seq_saved_%d" % seq_index, self.lib.dmt_seq_saved_bytes(C.byref(cfg), n_tok))
            with self._Stage(self, "seq_encode", 1):
                abi.check(self.lib.dmt_seq_encode_fwd_train(C.byref(cfg), C.byref(si), C.byref(self._seq_w[seq_index]),
                                                            out, out_ld, n_tok, saved.data_ptr(), saved.numel(),
                                                            self._stream()))
            return keep
        ws_ptr, ws_bytes = None, 0
        if self.precision == abi.PRECISION_BF16:
            ws, ws_bytes = self._prepared_for(seq_index, cfg)
            ws_ptr = ws.data_ptr()
        defer = deferred is not None and self.precision == abi.PRECISION_BF16 and self._v2_ok
        if defer:
            cfg.flags |= abi.SEQ_DEFER_TAIL
        si, keep = self._seq_input(inputs, seq, batch)
        stream = self._stream()
        # bf16: the fused tile kernel (+ the row-batched decoder tail kernel unless deferred)
        with self._Stage(self, "seq_encode", 2 if self.precision == abi.PRECISION_BF16 and not defer else 1):
            abi.check(self.lib.dmt_seq_encode_fwd(C.byref(cfg), C.byref(si), C.byref(self._seq_w[seq_index]),
                                                  out, out_ld, ws_ptr, ws_bytes, stream))
        if defer:
            deferred.append((cfg, si, self._seq_w[seq_index], out, out_ld, ws_ptr))
        return keep

    def seq_encode_multi(self, inputs, x, x_ld, batch):
        """A2-A8 for ALL behaviour sequences (bf16 path): `dmt_seq_encode_multi_fwd` -- length classes on the
        device, one persistent tile-kernel launch over every (sequence, class) segment, one tail launch; writes
        the interest vectors into their columns of the MMoE input `x` (fp32, or bf16: DMT_SEQ_OUT_BF16)."""
        plan = self.plan
        n = len(plan.sequences)
        items, keep = [], []
        esz = x.element_size()
        for s, seq in enumerate(plan.sequences):
            cfg = self._seq_cfg(inputs, seq, batch, self.precision)
            if x.dtype == torch.bfloat16:
                cfg.flags |= abi.SEQ_OUT_BF16
            ws, ws_bytes = self._prepared_for(s, cfg)
            si, kp = self._seq_input(inputs, seq, batch)
            keep += kp
            col = plan.interest_col + s * plan.d_model
            items.append((cfg, si, self._seq_w[s], x.data_ptr() + esz * col, x_ld, ws.data_ptr(), ws_bytes))
        keep.append(items)
        cfgs = (C.c_void_p * n)(*[C.addressof(d[0]) for d in items])
        ins = (C.c_void_p * n)(*[C.addressof(d[1]) for d in items])
        wts = (C.c_void_p * n)(*[C.addressof(d[2]) for d in items])
        outs = (C.c_void_p * n)(*[d[3] for d in items])
        lds = (C.c_int64 * n)(*[d[4] for d in items])
        wss = (C.c_void_p * n)(*[d[5] for d in items])
        wsb = (C.c_size_t * n)(*[d[6] for d in items])
        with self._Stage(self, "seq_encode", 3):
            abi.check(self.lib.dmt_seq_encode_multi_fwd(n, cfgs, ins, wts, outs, lds, wss, wsb, self._stream()))
        return keep

    def seq_tails(self, deferred):
        """dmt_seq_tail_fwd over the sequences collected by `seq_encode(..., deferred=list)`."""
        n = len(deferred)
        if n == 0:
            return
        cfgs = (C.c_void_p * n)(*[C.addressof(d[0]) for d in deferred])
        ins = (C.c_void_p * n)(*[C.addressof(d[1]) for d in deferred])
        wts = (C.c_void_p * n)(*[C.addressof(d[2]) for d in deferred])
        outs = (C.c_void_p * n)(*[d[3] for d in deferred])
        lds = (C.c_int64 * n)(*[d[4] for d in deferred])
        wss = (C.c_void_p * n)(*[d[5] for d in deferred])
        with self._Stage(self, "seq_encode", 1):
            abi.check(self.lib.dmt_seq_tail_fwd(n, cfgs, ins, wts, outs, lds, wss, self._stream()))

    def pool_mean(self, inputs, specs, tables_bias, out, batch):
        plan = self.plan
        key = (id(specs), tables_bias, len(specs))
        arr = self._pool_static.get(key)
        if arr is None:                      # descriptor array reused across calls (every field is rewritten)
            arr = self._pool_static[key] = (abi.PoolFeat * len(specs))()
        keep = []
        for i, p in enumerate(specs):
            table = self.params.table(p.table, bias=tables_bias)
            sp = self._sparse(inputs, p.feature)
            if sp.offsets.numel() != batch + 1:
                raise ValueError("feature %r: offsets has %d entries, batch is %d" % (p.feature, sp.offsets.numel(), batch))
            keep.append(sp)
            e = arr[i]
            e.table, e.rows, e.dim, e.out_col = table.data_ptr(), table.shape[0], table.shape[1], p.col
            e.ids, e.offsets = sp.values.data_ptr(), sp.offsets.data_ptr()
            e.weights = None if sp.weights is None else sp.weights.data_ptr()
        stream = self._stream()
        fn = self.lib.dmt_pool_mean_fwd_bf16 if out.dtype == torch.bfloat16 else self.lib.dmt_pool_mean_fwd
        for s in range(0, len(specs), abi.MAX_POOL_FEATS):
            n = min(abi.MAX_POOL_FEATS, len(specs) - s)
            sub = C.cast(C.byref(arr, s * C.sizeof(abi.PoolFeat)), C.POINTER(abi.PoolFeat))
            with self._Stage(self, "pool_mean", 1):
                abi.check(fn(batch, n, sub, out.data_ptr(), out.stride(0), stream))
        return keep

    def _mmoe_cfg(self, batch, precision):
        plan = self.plan
        cfg = abi.MmoeCfg()
        cfg.batch, cfg.in_dim, cfg.n_experts = batch, plan.mmoe_in, plan.num_experts
        cfg.n_layers = len(plan.hidden_units_bottom)
        for i, u in enumerate(plan.hidden_units_bottom):
            cfg.units[i] = u
        cfg.n_tasks = plan.num_tasks
        cfg.n_tower_layers = len(plan.hidden_units_task)
        for i, u in enumerate(plan.hidden_units_task):
            cfg.tower_units[i] = u
        cfg.precision = precision
        return cfg

    def mmoe(self, x, batch, logits):
        cfg = self._mmoe_cfg(batch, self.precision)
        nbytes = self.lib.dmt_mmoe_workspace_bytes(C.byref(cfg))
        ws = self._buf("mmoe_ws", ((nbytes + 255) // 256 * 256,), torch.uint8)
        stream = self._stream()
        prep_ptr = None
        launches = cfg.n_layers + 1
        if self.precision == abi.PRECISION_BF16:
            ver, prep = self._prepared.get("mmoe", (-1, None))
            pbytes = self.lib.dmt_mmoe_prepared_bytes(C.byref(cfg))
            if prep is None:
                prep = torch.empty(pbytes, dtype=torch.uint8, device=self.device)
            if ver != self.params_version:
                with self._Stage(self, "prepare_weights", 2 * cfg.n_layers):
                    abi.check(self.lib.dmt_mmoe_prepare_weights(C.byref(cfg), C.byref(self._mmoe_w),
                                                                prep.data_ptr(), pbytes, stream))
                self._prepared["mmoe"] = (self.params_version, prep)
            prep_ptr = prep.data_ptr()
            # fp32 input: conversion + gates, one GEMM per layer, head; bf16 input: the gates ride in the layer-0 GEMM;
            # hidden_units_bottom = (x, 256, 128) with one tower layer: layers 2 + 3 + tower are one kernel + mixture
            fused = (cfg.n_layers == 3 and cfg.units[1] == 256 and cfg.units[2] == 128 and cfg.n_tower_layers == 1
                     and cfg.n_tasks * cfg.tower_units[0] <= 64)
            launches = (0 if x.dtype == torch.bfloat16 else 1) + (3 if fused else cfg.n_layers + 1)
        fwd = self.lib.dmt_mmoe_fwd_bf16in if x.dtype == torch.bfloat16 else self.lib.dmt_mmoe_fwd
        with self._Stage(self, "mmoe", launches):
            abi.check(fwd(C.byref(cfg), C.byref(self._mmoe_w), x.data_ptr(), x.stride(0),
                          logits.data_ptr(), ws.data_ptr(), nbytes, prep_ptr, stream))

    def _bias_cfg(self, batch, passthrough=False, loss_unbias_method=None, loss_ctr_rel_method=None,
                  dropout_rates=None, dropout_seed=0):
        plan = self.plan
        cfg = abi.BiasLossCfg()
        cfg.batch = batch
        cfg.in_dim = 1 if passthrough else plan.bias_width
        cfg.n_hidden = -1 if passthrough else len(plan.hidden_units_bias)
        if not passthrough:
            for i, u in enumerate(plan.hidden_units_bias):
                cfg.units[i] = u
        cfg.two_head_multiply = 1 if (loss_unbias_method or plan.loss_unbias_method) == "two_head_multiply" else 0
        cfg.ctr_rel = 1 if (loss_ctr_rel_method or plan.loss_ctr_rel_method) == "ctr_rel" else 0
        for i in range(5):
            cfg.weight_ctr[i] = plan.weight_ctr[i]
            cfg.weight_ecvr[i] = plan.weight_ecvr[i]
        cfg.loss_weight[0], cfg.loss_weight[1] = plan.loss_weight[0], plan.loss_weight[1]
        for i, r in enumerate(dropout_rates or []):
            if i < abi.MAX_LAYERS:
                cfg.dropout_rate[i] = float(r)
        cfg.dropout_seed = int(dropout_seed) & 0xFFFFFFFF
        return cfg

    # ------------------------------------------------------------------ plugin protocol
    def inference(self, inputs, is_train=True, is_predict=False, dropout_seed=None):
        """mmoe_transformer_unbias.py:293-316.  Returns ((click_logit [B,1], order_logit [B,1]),
        y_bias [B,1]) or, with is_predict, just the logit pair.  With is_train and non-zero dropout rates the
        forward of the TRAINING graph runs (the dropout sites of TransformerModel.py:101,151 /
        TransformerModel_util.py:51 / mmoe_transformer_unbias.py:272,280 active, masks = the counter-based hash of
        `dropout_seed`, a fresh seed per call unless given) -- the same forward `compute_gradients` differentiates."""
        self._stream_h = torch.cuda.current_stream(self.device).cuda_stream
        try:
            plan = self.plan
            if is_train and (plan.dropout_rate > 0 or any(r > 0 for r in plan.dropout_rate_bias)):
                return self._inference_train(inputs, is_predict, dropout_seed)
            if self.precision == abi.PRECISION_TF32:      # the row-batched pipeline, dropout off
                return self._inference_train(inputs, is_predict, 0, dropout=False, engine=abi.PRECISION_TF32)
            return self._inference(inputs, is_train, is_predict)
        finally:
            self._stream_h = None

    def _inference_train(self, inputs, is_predict, dropout_seed, dropout=True, engine=None):
        """Forward of the training graph through the row-batched pipeline (dropout active unless dropout=False); the
        activations it leaves in the `saved` scratch are not kept for a backward."""
        from .. import dropout as DO
        plan, lib = self.plan, self.lib
        rate = float(plan.dropout_rate) if dropout else 0.0
        rates_bias = [float(r) for r in plan.dropout_rate_bias] if dropout else []
        if dropout_seed is None:
            self._train_calls = getattr(self, "_train_calls", 0) + 1
            dropout_seed = DO.step_seed(getattr(self, "dropout_base_seed", 20201019), self._train_calls)
        self.last_dropout_seed = dropout_seed
        inputs = self.stage_inputs(inputs)
        feats = inputs["features"] if plan.is_use_feature else None
        batch = inputs[plan.pooled[0].feature].offsets.numel() - 1
        stream = self._stream()
        F32 = self.train_precision if engine is None else engine
        x_ld = (plan.mmoe_in + 3) // 4 * 4
        x = self._buf("x", (batch, x_ld))
        keep = []
        side = self._seq_side_streams(len(plan.sequences)) if self._events is None and self.seq_streams else None
        main = torch.cuda.current_stream(self.device)
        if side is not None:     # one stream per behaviour sequence; the dense copy / pooled means overlap on `main`
            fork = self._ev_pool("fork")
            fork.record(main)
        if feats is not None:
            if feats.dtype != torch.float32 or feats.shape != (batch, plan.feature_dim):
                raise ValueError("'features' must be fp32 [%d, %d]" % (batch, plan.feature_dim))
            feats = feats.contiguous()
            with self._Stage(self, "copy_dense", 1):
                abi.check(lib.dmt_copy_dense_features(feats.data_ptr(), batch, plan.feature_dim,
                                                      x.data_ptr(), x_ld, stream))
            keep.append(feats)
        keep += self.pool_mean(inputs, plan.pooled, False, x, batch)
        for s, seq in enumerate(plan.sequences):
            cfg = self._seq_cfg(inputs, seq, batch, F32, rate, DO.step_seed(dropout_seed, 0, 1 + seq.index))
            si, kp = self._seq_input(inputs, seq, batch)
            keep += kp
            n_tok = self._sparse(inputs, seq.user_features[-1], "seq%d" % seq.index).values.numel()
            nbytes = lib.dmt_seq_saved_bytes(C.byref(cfg), n_tok)
            saved = self._scratch("seq_saved_%d" % s, nbytes)
            col = plan.interest_col + s * plan.d_model
            if side is not None:
                side[s].wait_event(fork)
            with self._Stage(self, "seq_encode_train", 1):
                abi.check(lib.dmt_seq_encode_fwd_train(C.byref(cfg), C.byref(si), C.byref(self._seq_w[s]),
                                                       x.data_ptr() + 4 * col, x_ld, n_tok, saved.data_ptr(),
                                                       saved.numel(), side[s].cuda_stream if side is not None else stream))
            if side is not None:
                ev = self._ev_pool("join%d" % s)
                ev.record(side[s])
                main.wait_event(ev)
        mcfg = self._mmoe_cfg(batch, F32)
        mws_bytes = lib.dmt_mmoe_train_workspace_bytes(C.byref(mcfg))
        mws = self._buf("mmoe_ws_f32", ((mws_bytes + 255) // 256 * 256,), torch.uint8)
        self._score_flip = getattr(self, "_score_flip", 0) ^ 1
        scores = self._buf("scores%d" % self._score_flip, (plan.num_tasks + 1, batch))
        self.last_scores = scores
        logits = scores[:plan.num_tasks]
        with self._Stage(self, "mmoe", mcfg.n_layers + 1):
            abi.check(lib.dmt_mmoe_fwd_train(C.byref(mcfg), C.byref(self._mmoe_w), x.data_ptr(), x_ld,
                                             logits.data_ptr(), mws.data_ptr(), mws_bytes, stream))
        self._last = {"x": x, "batch": batch, "keep": keep}
        y_rel = tuple(logits[t].view(batch, 1) for t in range(plan.num_tasks))
        if is_predict:
            return y_rel
        bias_in = self._buf("bias_in", (batch, plan.bias_width))
        keep += self.pool_mean(inputs, plan.bias_pooled, True, bias_in, batch)
        y_bias = scores[plan.num_tasks]
        bcfg = self._bias_cfg(batch, dropout_rates=rates_bias, dropout_seed=DO.step_seed(dropout_seed, 0, 0))
        with self._Stage(self, "bias_tower", 1):
            abi.check(lib.dmt_bias_loss_fwd(C.byref(bcfg), C.byref(self._bias_w), bias_in.data_ptr(),
                                            bias_in.stride(0), logits.data_ptr(), None, y_bias.data_ptr(),
                                            None, None, None, None, stream))
        return (y_rel, y_bias.view(batch, 1))

    def _inference(self, inputs, is_train, is_predict):
        plan = self.plan
        inputs = self.stage_inputs(inputs)
        feats = inputs["features"] if plan.is_use_feature else None
        first = inputs[plan.pooled[0].feature]
        batch = first.offsets.numel() - 1
        stream = self._stream()
        x_ld = (plan.mmoe_in + 3) // 4 * 4
        x = self._buf("x", (batch, x_ld))
        keep = []
        deferred = [] if len(plan.sequences) <= abi.MAX_TAIL_SEQS else None
        if (self.precision == abi.PRECISION_BF16 and self._v2_ok and self.seq_multi and batch > 0
                and 0 < len(plan.sequences) <= abi.MAX_TAIL_SEQS):
            # bf16: all sequences in ONE persistent tile-kernel launch over length-bucketed tiles; the MMoE input is
            # assembled in bf16 by its producers (no fp32 copy of x, no conversion pass)
            if self.x_bf16 and self.fwd_native and self._events is None and self.seq_streams:
                out = self._inference_native(inputs, feats, batch, is_predict)     # the same calls, issued natively
                if out is not None:
                    return out
            if self.x_bf16:
                x_ld = (plan.mmoe_in + 7) // 8 * 8
                x = self._buf("xb", (batch, x_ld), torch.bfloat16)
            # the sequence launches run on a side stream: the dense copy / pooled lookups fill the SMs while the
            # 3-CTA length-class kernel runs and as the persistent kernel's CTAs retire (they write disjoint columns)
            side = self._seq_side_streams(2) if self._events is None and self.seq_streams else None
            scores, y_bias = self._next_scores(batch), None
            if side is not None:
                main = torch.cuda.current_stream(self.device)
                fork = self._ev_pool("fork")
                fork.record(main)
                if not is_predict:                   # the bias branch depends on nothing else: its own stream
                    side[1].wait_event(fork)
                    self._stream_h = side[1].cuda_stream
                    y_bias = self._bias_branch(inputs, batch, scores, keep)
                    join_b = self._ev_pool("join1")
                    join_b.record(side[1])
                side[0].wait_event(fork)
                self._stream_h = side[0].cuda_stream
            keep += self.seq_encode_multi(inputs, x, x_ld, batch)
            if side is not None:
                join = self._ev_pool("join0")
                join.record(side[0])
                self._stream_h = main.cuda_stream
            if feats is not None:
                if self.x_bf16:
                    self._stage_dense_bf16(feats, batch, x, x_ld, keep)
                else:
                    self._copy_dense(feats, batch, x, x_ld, keep, self.precision)
            keep += self.pool_mean(inputs, plan.pooled, False, x, batch)
            if side is not None:
                main.wait_event(join)
            out = self._inference_head(inputs, x, batch, keep, is_predict, scores, y_bias)
            if side is not None and not is_predict:
                main.wait_event(join_b)
            return out
        # The behaviour sequences are independent of each other and of the dense / pooled columns until the MMoE
        # input: each runs on its own stream, so the CTAs of the next sequence's (persistent, one-CTA-per-SM) kernel
        # start on an SM the moment the previous kernel's CTA there retires -- no tail bubble, prologues (weight
        # images, TMEM allocation) hidden.  Per-stage timing (bench hook) serialises them so that each kernel's
        # duration is its own.
        side = self._seq_side_streams(len(plan.sequences)) if self._events is None and self.seq_streams else None
        if side is not None:
            main = torch.cuda.current_stream(self.device)
            fork = self._ev_pool("fork")
            fork.record(main)
            joins = []
            for s in range(len(plan.sequences)):
                side[s].wait_event(fork)
                self._stream_h = side[s].cuda_stream
                col = plan.interest_col + s * plan.d_model
                keep += self.seq_encode(inputs, s, x.data_ptr() + 4 * col, x_ld, batch, deferred)
                ev = self._ev_pool("join%d" % s)
                ev.record(side[s])
                joins.append(ev)
            self._stream_h = main.cuda_stream
        if feats is not None:
            self._copy_dense(feats, batch, x, x_ld, keep, self.precision)
        keep += self.pool_mean(inputs, plan.pooled, False, x, batch)
        if side is not None:
            for ev in joins:
                main.wait_event(ev)
        else:
            for s in range(len(plan.sequences)):
                col = plan.interest_col + s * plan.d_model
                keep += self.seq_encode(inputs, s, x.data_ptr() + 4 * col, x_ld, batch, deferred)
        if deferred:
            self.seq_tails(deferred)       # one launch for the decoder tails of every sequence
        return self._inference_head(inputs, x, batch, keep, is_predict)

    # ------------------------------------------------------------------ native forward driver (dmt_forward_bf16)
    def _fwd_feature_names(self):
        names = list(self.plan.all_id_features())
        for seq in self.plan.sequences:
            for n in list(seq.user_features) + list(seq.item_features):
                if n not in names:
                    names.append(n)
        return names

    def _fwd_state(self, batch, is_predict):
        """Static half of the forward: `dmt_fwd_desc` + everything it points to (descriptor templates, feature
        bindings, buffers), built once per (batch size, is_predict, parameter version)."""
        key = (batch, bool(is_predict))
        st = self._fwd_states.get(key)
        if st is not None and st["version"] == self.params_version:
            return st
        plan, lib = self.plan, self.lib
        names = self._fwd_feature_names()
        index = {n: i for i, n in enumerate(names)}
        st = {"version": self.params_version, "names": names, "keep": []}
        d = abi.FwdDesc()
        d.batch, d.n_seq, d.is_predict = batch, len(plan.sequences), 1 if is_predict else 0
        d.feature_dim = plan.feature_dim if plan.is_use_feature else 0
        d.interest_col = plan.interest_col

        def pool_templates(specs, bias):
            arr = (abi.PoolFeat * max(len(specs), 1))()
            feat = (C.c_int32 * max(len(specs), 1))()
            for i, p in enumerate(specs):
                table = self.params.table(p.table, bias=bias)
                arr[i].table, arr[i].rows, arr[i].dim, arr[i].out_col = table.data_ptr(), table.shape[0], table.shape[1], p.col
                feat[i] = index[p.feature]
            st["keep"] += [arr, feat]
            return arr, feat

        if len(plan.pooled) > abi.MAX_POOL_FEATS or len(plan.bias_pooled) > abi.MAX_POOL_FEATS:
            raise ValueError("more than %d pooled lookups" % abi.MAX_POOL_FEATS)
        arr, feat = pool_templates(plan.pooled, False)
        d.n_pool, d.pool, d.pool_feature = len(plan.pooled), C.addressof(arr), C.addressof(feat)
        arr, feat = pool_templates(plan.bias_pooled, True)
        d.n_bias_pool, d.bias_pool, d.bias_pool_feature = len(plan.bias_pooled), C.addressof(arr), C.addressof(feat)
        for s, seq in enumerate(plan.sequences):
            cfg = abi.SeqCfg(batch, plan.d_model, plan.d_ff, plan.num_heads, plan.num_blocks_encode,
                             plan.num_blocks_decode, plan.maxlen_k, 1 if plan.zero_pad else 0, len(seq.user_features),
                             abi.PRECISION_BF16, seq.maxlen, abi.SEQ_OUT_BF16, 0.0, 0)
            ws, ws_bytes = self._prepared_for(s, cfg)
            si = abi.SeqInput()
            uf = (C.c_int32 * abi.MAX_SEQ_FEATS)()
            itf = (C.c_int32 * abi.MAX_SEQ_FEATS)()
            for f, (u, it) in enumerate(zip(seq.user_features, seq.item_features)):
                table = self.params.table(seq.tables[f])
                si.table[f], si.rows[f], si.dim[f] = abi.ptr(table), table.shape[0], table.shape[1]
                uf[f], itf[f] = index[u], index[it]
            st["keep"] += [cfg, ws, si, uf, itf]
            d.seq_cfg[s], d.seq_in[s], d.seq_w[s] = C.addressof(cfg), C.addressof(si), C.addressof(self._seq_w[s])
            d.seq_user_feature[s], d.seq_item_feature[s] = C.addressof(uf), C.addressof(itf)
            d.seq_ws[s], d.seq_ws_bytes[s] = ws.data_ptr(), ws_bytes
        mcfg = self._mmoe_cfg(batch, abi.PRECISION_BF16)
        nbytes = lib.dmt_mmoe_workspace_bytes(C.byref(mcfg))
        mws = self._buf("mmoe_ws", ((nbytes + 255) // 256 * 256,), torch.uint8)
        ver, prep = self._prepared.get("mmoe", (-1, None))
        pbytes = lib.dmt_mmoe_prepared_bytes(C.byref(mcfg))
        if prep is None:
            prep = torch.empty(pbytes, dtype=torch.uint8, device=self.device)
        if ver != self.params_version:
            abi.check(lib.dmt_mmoe_prepare_weights(C.byref(mcfg), C.byref(self._mmoe_w), prep.data_ptr(), pbytes,
                                                   self._stream()))
            self._prepared["mmoe"] = (self.params_version, prep)
        d.mmoe_cfg, d.mmoe_w = C.addressof(mcfg), C.addressof(self._mmoe_w)
        d.mmoe_ws, d.mmoe_ws_bytes, d.mmoe_prepared = mws.data_ptr(), nbytes, prep.data_ptr()
        bcfg = self._bias_cfg(batch)
        bias_in = self._buf("bias_in", (batch, plan.bias_width))
        d.bias_cfg, d.bias_w = C.addressof(bcfg), C.addressof(self._bias_w)
        d.bias_in, d.bias_ld = bias_in.data_ptr(), bias_in.stride(0)
        x_ld = (plan.mmoe_in + 7) // 8 * 8
        xb = self._buf("xb", (batch, x_ld), torch.bfloat16)
        d.xb, d.xb_ld = xb.data_ptr(), x_ld
        st["keep"] += [mcfg, mws, prep, bcfg, bias_in, xb]
        st["desc"], st["xb"] = d, xb
        # kernel launches per call: bias pool + tower | length classes + tile kernel + tails | dense + pooled |
        # MMoE layer 0 + fused tail + mixture
        st["launches"] = (0 if is_predict else 2) + 3 + (1 if d.feature_dim else 0) + 1 + 3
        self._fwd_states[key] = st
        return st

    def _fwd_table(self, inputs, st, batch):
        """Per-call half: the [n_features, 3] pointer table (ids, offsets, weights)."""
        names = st["names"]
        packed = inputs.get("__packed__")
        if packed is not None and "__buffer__" in inputs and not inputs.get("__remap__"):
            tab, cnt = packed.feature_offsets(names)
            ok = packed.__dict__.setdefault("_fwd_ok", set())        # validated once per (packed batch, model)
            if (id(self), batch) not in ok:
                if any(n + "Wts" in inputs for n in names):
                    return None
                item = {n for seq in self.plan.sequences for n in seq.item_features}
                for i, n in enumerate(names):
                    if cnt[i, 1] != batch + 1:
                        raise ValueError("feature %r: offsets has %d entries, batch is %d" % (n, cnt[i, 1], batch))
                    if n in item and cnt[i, 0] != batch:
                        raise ValueError("item feature %r must hold exactly one id per sample" % n)
                ok.add((id(self), batch))
            buf, wide = inputs["__buffer__"]
            base, wbase = buf.data_ptr(), (wide.data_ptr() if wide is not None else 0)
            out = tab.copy()
            in_wide = (out[:, 0] >> 62) & 1
            out[:, 0] = (out[:, 0] & ((1 << 62) - 1)) + np.where(in_wide == 1, wbase, base)
            out[:, 1] += base
            out[:, 2] = np.where(tab[:, 2] >= 0, tab[:, 2] + base, 0)
            return out
        # a resident batch keeps its pointer table; it is rebuilt when any feature entry has been replaced since
        stamp = tuple(id(inputs.get(n)) for n in names) + tuple(id(inputs.get(n + "Wts")) for n in names)
        cached = inputs.get("__fwd_table__")
        if cached is not None and cached[0] == stamp:
            return cached[1]
        item = {n for seq in self.plan.sequences for n in seq.item_features}
        out = np.zeros((len(names), 3), dtype=np.int64)
        for i, n in enumerate(names):
            sp = self._sparse(inputs, n)
            if sp.offsets.numel() != batch + 1:
                raise ValueError("feature %r: offsets has %d entries, batch is %d" % (n, sp.offsets.numel(), batch))
            if n in item and sp.values.numel() != batch:
                raise ValueError("item feature %r must hold exactly one id per sample" % n)
            out[i, 0], out[i, 1] = sp.values.data_ptr(), sp.offsets.data_ptr()
            out[i, 2] = 0 if sp.weights is None else sp.weights.data_ptr()
        inputs["__fwd_table__"] = (stamp, out)
        return out

    def _inference_native(self, inputs, feats, batch, is_predict):
        """One `dmt_forward_bf16` call: descriptor templates patched and the three branches issued natively."""
        plan = self.plan
        if inputs.get("__remap__"):
            return None
        st = self._fwd_state(batch, is_predict)
        table = self._fwd_table(inputs, st, batch)
        if table is None:
            return None
        fptr, fbf16 = None, 0
        if feats is not None:
            if tuple(feats.shape) != (batch, plan.feature_dim) or feats.dtype not in (torch.float32, torch.bfloat16):
                raise ValueError("'features' must be fp32 (or bf16) [%d, %d]" % (batch, plan.feature_dim))
            feats = feats.contiguous()
            fptr, fbf16 = feats.data_ptr(), 1 if feats.dtype == torch.bfloat16 else 0
        scores = self._next_scores(batch)
        ready = inputs.get("__ready__")       # prefetched batch: the copy's event (the length classes may start early)
        st["desc"].inputs_ready = ready.cuda_event if ready is not None else None
        abi.check(self.lib.dmt_forward_bf16(C.byref(st["desc"]), table.shape[0], table.ctypes.data, fptr, fbf16,
                                            scores.data_ptr(), self._stream()))
        self.launches += st["launches"]
        self._last = {"x": st["xb"], "batch": batch, "keep": [table, feats, inputs]}
        logits = scores[:plan.num_tasks]
        y_rel = tuple(logits[t].view(batch, 1) for t in range(plan.num_tasks))
        if is_predict:
            return y_rel
        return (y_rel, scores[plan.num_tasks].view(batch, 1))

    def _next_scores(self, batch):
        """Scores of one call live in ONE [num_tasks + 1, B] buffer (task logits, then y_bias) so that a caller can
        read them back with a single copy; two buffers alternate, so the result of call i stays valid while call
        i + 1 runs."""
        self._score_flip = getattr(self, "_score_flip", 0) ^ 1
        scores = self._buf("scores%d" % self._score_flip, (self.plan.num_tasks + 1, batch))
        self.last_scores = scores
        return scores

    def _bias_branch(self, inputs, batch, scores, keep):
        """embedding_combiner_bias + embedding_mlp_bias (mmoe_transformer_unbias.py:235-289, eval mode) -> y_bias =
        scores[num_tasks].  Reads nothing the rest of the forward writes: it may run on its own stream."""
        plan = self.plan
        bias_in = self._buf("bias_in", (batch, plan.bias_width))
        keep += self.pool_mean(inputs, plan.bias_pooled, True, bias_in, batch)
        y_bias = scores[plan.num_tasks]
        cfg = self._bias_cfg(batch)
        with self._Stage(self, "bias_tower", 1):
            abi.check(self.lib.dmt_bias_loss_fwd(C.byref(cfg), C.byref(self._bias_w), bias_in.data_ptr(),
                                                 bias_in.stride(0), scores.data_ptr(), None, y_bias.data_ptr(),
                                                 None, None, None, None, self._stream()))
        return y_bias

    def _inference_head(self, inputs, x, batch, keep, is_predict, scores=None, y_bias=None):
        """MMoE + towers (+ the bias tower unless it already ran) on the assembled MMoE input
        (mmoe_transformer_unbias.py:218-316)."""
        plan = self.plan
        if scores is None:
            scores = self._next_scores(batch)
        logits = scores[:plan.num_tasks]
        self.mmoe(x, batch, logits)
        self._last = {"x": x, "batch": batch, "keep": keep}
        y_rel = tuple(logits[t].view(batch, 1) for t in range(plan.num_tasks))
        if is_predict:
            return y_rel
        if y_bias is None:
            y_bias = self._bias_branch(inputs, batch, scores, keep)
        return (y_rel, y_bias.view(batch, 1))

    def loss(self, logits, mask, loss_unbias_method=None, loss_ctr_rel_method=None, want_probs=False,
             want_grads=False):
        """A12 on the outputs of `inference` (inference_mlp.py:173-223)."""
        (click, order), y_bias = logits
        batch = click.shape[0]
        last = getattr(self, "last_scores", None)
        lg = last[:self.plan.num_tasks] if last is not None and last.shape[1] == batch else \
            self._buf("logits", (self.plan.num_tasks, batch))
        if click.data_ptr() != lg[0].data_ptr() or order.data_ptr() != lg[1].data_ptr():
            lg = self._buf("logits_in", (2, batch))
            lg[0].copy_(click.reshape(-1))
            lg[1].copy_(order.reshape(-1))
        yb_in = y_bias.reshape(-1).contiguous()
        mask = self._dev(mask)
        if mask.dtype != torch.float32 or tuple(mask.shape) != (batch, 5):
            raise ValueError("mask must be fp32 [%d, 5]" % batch)
        mask = mask.contiguous()
        cfg = self._bias_cfg(batch, passthrough=True, loss_unbias_method=loss_unbias_method,
                             loss_ctr_rel_method=loss_ctr_rel_method)
        yb_out = self._buf("y_bias_pass", (batch,))
        loss = self._buf("loss", (1,))
        probs = self._buf("probs", (2, batch)) if want_probs else None
        dlog = self._buf("dlogits", (3, batch)) if want_grads else None
        scratch = self._buf("loss_scratch", (self.lib.dmt_loss_scratch_bytes(batch),), torch.uint8)
        stream = self._stream()
        with self._Stage(self, "loss", 2):
            abi.check(self.lib.dmt_bias_loss_fwd(C.byref(cfg), C.byref(self._bias_w), yb_in.data_ptr(), 1,
                                                 lg.data_ptr(), mask.data_ptr(), yb_out.data_ptr(),
                                                 abi.ptr(probs), loss.data_ptr(), abi.ptr(dlog),
                                                 scratch.data_ptr(), stream))
        self._last_loss = {"mask": mask, "yb": yb_in}
        out = loss[0].clone()       # the buffer is reused by the next call
        if want_probs or want_grads:
            return out, probs, dlog
        return out

    # ------------------------------------------------------------------ A13: gradients
    def bind_grad_buffer(self, flat):
        """Use `flat` (fp32, laid out like `params.dense`, e.g. the front of a data-parallel allreduce bucket)
        as the dense gradient buffer."""
        if flat.numel() != self.params.dense.numel() or flat.dtype != torch.float32 or not flat.is_contiguous():
            raise ValueError("gradient buffer must be a contiguous fp32 tensor of %d elements" % self.params.dense.numel())
        self._grad_dense = flat
        gviews = {sp.name: flat[sp.offset:sp.offset + sp.numel].view(sp.shape) for sp in self.params.specs}
        self._seq_g, self._mmoe_g, self._bias_g = self._bind(gviews)
        self._grad_views = gviews

    def compute_gradients(self, inputs, mask=None, loss_unbias_method=None, loss_ctr_rel_method=None,
                          is_train=True, dropout_seed=None):
        """`optimizer.compute_gradients(loss)` of the reference's training graph (run_dnn.py:154-181, built with
        is_train=True): one forward that saves activations + the backward.  Returns `(loss, Gradients)`; the
        loss is the batch mean, so averaging `Gradients` over data-parallel ranks equals `average_gradients`
        (run_dnn.py:45-80).

        is_train activates the dropout sites of the graph (transformer_dropout_rate at the encoder / decoder
        inputs and the attention probabilities, dropout_rate_bias in the bias tower); the keep masks are a
        counter-based hash of (dropout_seed, site, element), a fresh seed per call unless one is given."""
        from ..optim import Gradients, LookupGrad
        from .. import dropout as DO
        plan, lib = self.plan, self.lib
        rate = float(plan.dropout_rate) if is_train else 0.0
        rates_bias = [float(r) for r in plan.dropout_rate_bias] if is_train else []
        if dropout_seed is None:
            self._train_calls = getattr(self, "_train_calls", 0) + 1
            dropout_seed = DO.step_seed(getattr(self, "dropout_base_seed", 20201019), self._train_calls)
        self.last_dropout_seed = dropout_seed
        inputs = self.stage_inputs(inputs)
        if "__buffer__" in inputs:
            raise TypeError("compute_gradients sorts and indexes the id arrays with torch: stage the batch with "
                            "tensor views (prefetch(packed) / PackedBatch.to(device)), not views=False")
        mask = self._dev(inputs["mask"] if mask is None else mask)
        feats = inputs["features"] if plan.is_use_feature else None
        batch = inputs[plan.pooled[0].feature].offsets.numel() - 1
        if mask.dtype != torch.float32 or tuple(mask.shape) != (batch, 5):
            raise ValueError("mask must be fp32 [%d, 5]" % batch)
        mask = mask.contiguous()
        stream = self._stream()
        F32 = self.train_precision   # engine of the training-path GEMMs (activations / gradients stay fp32)

        if getattr(self, "_grad_dense", None) is None:
            self.bind_grad_buffer(torch.zeros_like(self.params.dense))
        g_dense = self._grad_dense
        g_dense.zero_()

        # ---- forward (activations saved)
        x_ld = (plan.mmoe_in + 3) // 4 * 4
        x = self._buf("x", (batch, x_ld))
        keep = []
        side = self._seq_side_streams(len(plan.sequences)) if self._events is None and self.seq_streams else None
        main = torch.cuda.current_stream(self.device)
        if side is not None:     # the sequence pipelines fork here: the dense copy / pooled means overlap them on `main`
            fork = self._ev_pool("fork")
            fork.record(main)
        if feats is not None:
            if feats.dtype != torch.float32 or feats.shape != (batch, plan.feature_dim):
                raise ValueError("'features' must be fp32 [%d, %d]" % (batch, plan.feature_dim))
            feats = feats.contiguous()
            with self._Stage(self, "copy_dense", 1):
                abi.check(lib.dmt_copy_dense_features(feats.data_ptr(), batch, plan.feature_dim,
                                                      x.data_ptr(), x_ld, stream))
            keep.append(feats)
        keep += self.pool_mean(inputs, plan.pooled, False, x, batch)
        seq_state = []
        # one stream per behaviour sequence (independent pipelines of many short kernels: their tails and the
        # low-occupancy ones overlap); serial under per-stage timing
        for s, seq in enumerate(plan.sequences):
            cfg = self._seq_cfg(inputs, seq, batch, F32, rate, DO.step_seed(dropout_seed, 0, 1 + seq.index))
            si, kp = self._seq_input(inputs, seq, batch)
            keep += kp
            users = [self._sparse(inputs, uf, "seq%d" % seq.index) for uf in seq.user_features]
            n_tok = users[-1].values.numel()
            if any(u.values.numel() != n_tok for u in users):
                raise ValueError("sequence %d: the id features of one behaviour sequence must have equal lengths" % s)
            nbytes = lib.dmt_seq_saved_bytes(C.byref(cfg), n_tok)
            saved = self._scratch("seq_saved_%d" % s, nbytes)
            col = plan.interest_col + s * plan.d_model
            if side is not None:
                side[s].wait_event(fork)
            with self._Stage(self, "seq_encode_train", 1):
                abi.check(lib.dmt_seq_encode_fwd_train(C.byref(cfg), C.byref(si), C.byref(self._seq_w[s]),
                                                       x.data_ptr() + 4 * col, x_ld, n_tok, saved.data_ptr(),
                                                       saved.numel(), side[s].cuda_stream if side is not None else stream))
            if side is not None:
                ev = self._ev_pool("join%d" % s)
                ev.record(side[s])
                main.wait_event(ev)
            seq_state.append((cfg, si, users, n_tok, saved, col))
        mcfg = self._mmoe_cfg(batch, F32)
        mws_bytes = lib.dmt_mmoe_train_workspace_bytes(C.byref(mcfg))
        mws = self._buf("mmoe_ws_f32", ((mws_bytes + 255) // 256 * 256,), torch.uint8)
        logits = self._buf("logits", (plan.num_tasks, batch))
        with self._Stage(self, "mmoe", mcfg.n_layers + 1):
            abi.check(lib.dmt_mmoe_fwd_train(C.byref(mcfg), C.byref(self._mmoe_w), x.data_ptr(), x_ld,
                                             logits.data_ptr(), mws.data_ptr(), mws_bytes, stream))
        bias_in = self._buf("bias_in", (batch, plan.bias_width))
        keep += self.pool_mean(inputs, plan.bias_pooled, True, bias_in, batch)
        y_bias = self._buf("y_bias", (batch,))
        bcfg = self._bias_cfg(batch, loss_unbias_method=loss_unbias_method, loss_ctr_rel_method=loss_ctr_rel_method,
                              dropout_rates=rates_bias, dropout_seed=DO.step_seed(dropout_seed, 0, 0))
        loss = self._buf("loss", (1,))
        dlog = self._buf("dlogits", (3, batch))
        scratch = self._buf("loss_scratch", (lib.dmt_loss_scratch_bytes(batch),), torch.uint8)
        with self._Stage(self, "bias_loss", 2):
            abi.check(lib.dmt_bias_loss_fwd(C.byref(bcfg), C.byref(self._bias_w), bias_in.data_ptr(),
                                            bias_in.stride(0), logits.data_ptr(), mask.data_ptr(), y_bias.data_ptr(),
                                            None, loss.data_ptr(), dlog.data_ptr(), scratch.data_ptr(), stream))

        # ---- backward
        d_bias_in = self._buf("d_bias_in", (batch, plan.bias_width))
        nb = lib.dmt_bias_bwd_workspace_bytes(C.byref(bcfg))
        bws = self._buf("bias_bwd_ws", ((nb + 255) // 256 * 256,), torch.uint8)
        with self._Stage(self, "bias_bwd", 3):
            abi.check(lib.dmt_bias_bwd(C.byref(bcfg), C.byref(self._bias_w), bias_in.data_ptr(), bias_in.stride(0),
                                       dlog[2].data_ptr(), C.byref(self._bias_g), d_bias_in.data_ptr(),
                                       d_bias_in.stride(0), bws.data_ptr(), bws.numel(), stream))
        dx = self._buf("dx", (batch, x_ld))
        nb = lib.dmt_mmoe_bwd_workspace_bytes(C.byref(mcfg))
        mbws = self._buf("mmoe_bwd_ws", ((nb + 255) // 256 * 256,), torch.uint8)
        col0 = plan.feature_dim if plan.is_use_feature else 0
        with self._Stage(self, "mmoe_bwd", 4 + 4 * mcfg.n_layers):
            abi.check(lib.dmt_mmoe_bwd(C.byref(mcfg), C.byref(self._mmoe_w), x.data_ptr(), x_ld, mws.data_ptr(),
                                       dlog.data_ptr(), C.byref(self._mmoe_g), dx.data_ptr(), x_ld, col0,
                                       mbws.data_ptr(), mbws.numel(), stream))
        lookups = {}

        def add(scope, lg):
            lookups.setdefault(scope, []).append(lg)

        for p in plan.pooled:
            sp = self._sparse(inputs, p.feature)
            add(plan.tables[p.table].scope, LookupGrad(sp.values, dx, p.col, 0, sp.offsets, sp.weights, True))
        for p in plan.bias_pooled:
            sp = self._sparse(inputs, p.feature)
            add(plan.bias_tables[p.table].scope,
                LookupGrad(sp.values, d_bias_in, p.col, 0, sp.offsets, sp.weights, True))
        id_off = -1 if plan.zero_pad else 0
        if side is not None:
            fork = self._ev_pool("fork_bwd")
            fork.record(main)
        for s, seq in enumerate(plan.sequences):
            cfg, si, users, n_tok, saved, col = seq_state[s]
            d_tok = self._scratch("d_tokens_%d" % s, max(n_tok, 1) * plan.d_model * 4)
            d_tok = d_tok[:max(n_tok, 1) * plan.d_model * 4].view(torch.float32).view(max(n_tok, 1), plan.d_model)
            d_tar = self._buf("d_target_%d" % s, (batch, plan.d_model))
            nb = lib.dmt_seq_bwd_workspace_bytes(C.byref(cfg), n_tok)
            sws = self._scratch("seq_bwd_ws_%d" % s if side is not None else "seq_bwd_ws", nb)
            if side is not None:
                side[s].wait_event(fork)
            with self._Stage(self, "seq_bwd", 30):
                abi.check(lib.dmt_seq_encode_bwd(C.byref(cfg), C.byref(si), C.byref(self._seq_w[s]), n_tok,
                                                 saved.data_ptr(), saved.numel(), dx.data_ptr() + 4 * col, x_ld,
                                                 C.byref(self._seq_g[s]), d_tok.data_ptr(), d_tar.data_ptr(),
                                                 sws.data_ptr(), sws.numel(),
                                                 side[s].cuda_stream if side is not None else stream))
            if side is not None:
                ev = self._ev_pool("join_bwd%d" % s)
                ev.record(side[s])
                main.wait_event(ev)
            for f, u in enumerate(users):
                scope = plan.tables[seq.tables[f]].scope
                if n_tok:
                    add(scope, LookupGrad(u.values, d_tok, seq.col_offsets[f], id_off))
                it = self._sparse(inputs, seq.item_features[f], "seq%d" % seq.index)
                add(scope, LookupGrad(it.values, d_tar, seq.col_offsets[f], id_off))
        self._keep_train = (keep, seq_state, inputs)
        return loss[0].clone(), Gradients(g_dense, lookups, self._grad_views)   # (the loss buffer is reused)

    def l2_norm(self, inputs):
        raise NotImplementedError("l2_norm is gated off by wnd_wd = 0.0 in dmt.conf (run_dnn.py:174)")

    def embedding_update(self, sess=None):
        raise NotImplementedError("update_emb is empty in dmt.conf (base.py:178-196 is unused)")
