"""Synchronous data-parallel training of the DMT hot path: the replacement for the tower loop of
`run_dnn.py:148-207` (`compute_gradients` per tower -> `average_gradients` -> `apply_gradients`).

One process per GPU.  Per step and rank:

  1. row-sharded tables (Sku): all-to-all the lookups of the local batch to the owning ranks and bring the rows
     back as a compact table (shard.py);
  2. `compute_gradients` on the local batch (training forward + backward, C-ABI kernels);
  3. ONE NCCL allreduce of a flat bucket = [all dense variables | densified gradients of the replicated small
     tables] -- `average_gradients` of the reference (each tower loss is a batch mean, so mean over ranks ==
     mean over the global batch);
  4. sharded tables: one gradient row per compact row travels back to its owner (all-to-all), which runs the
     sorted segmented reduction + TF-1 Adam on its shard; every rank runs the dense Adam on its replicas.

With world == 1 steps 1, 3 and 4's exchange vanish and the tables take the sparse path of optim.TFAdam.
Host code only; the collectives are torch.distributed (NCCL over NVLink), the arithmetic is the C-ABI library.
"""
import contextlib
import ctypes as C
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import abi
from .data import SparseIds
from .net.mmoe_transformer_unbias import mmoe_transformer_unbias
from .optim import LookupGrad, TFAdam
from .params import ParamStore
from .shard import RowExchange, RowShard, remap_ids


_nullctx = contextlib.nullcontext


class Trainer(object):
    def __init__(self, plan, device, learning_rate=None, world=1, rank=0, seed=20201019, sharded_tables=("Sku",),
                 group=None, randomize=None, force_dp_path=False, precision="f32", train_gemm=None,
                 step_boundary=None, global_step=0):
        """learning_rate / step_boundary: default to the conf's `[model] learning_rate` / `step_boundary` lists
        (piecewise-constant schedule over `global_step`, run_dnn.py:119-126); a float = constant rate.
        force_dp_path: run the data-parallel code path (compact tables, densified replicas, bucket) even
        with world == 1 -- the single-GPU test of that path."""
        if learning_rate is None:
            learning_rate = list(getattr(plan, "learning_rate", [1e-3]))
            if step_boundary is None:
                step_boundary = list(getattr(plan, "step_boundary", []))
        self.plan, self.device = plan, torch.device(device)
        self.world, self.rank, self.group = int(world), int(rank), group
        self.dp = self.world > 1 or force_dp_path
        self.sharded: Dict[str, RowShard] = {}
        row_shards = {}
        if self.dp:
            for name in sharded_tables:
                if name in plan.tables:
                    sh = RowShard(plan.tables[name].rows, self.world, self.rank)
                    self.sharded[plan.tables[name].scope] = sh
                    row_shards[plan.tables[name].scope] = (sh.lo, sh.hi)
        store = ParamStore(plan, device="cpu", seed=seed, row_shards=None)
        if randomize is not None:
            store.randomize_(randomize)
        if row_shards:
            for scope, (lo, hi) in row_shards.items():
                store.tables[scope] = store.tables[scope][lo:hi].clone()
        store = _store_to(store, self.device)
        self.model = mmoe_transformer_unbias(plan, device=self.device, params=store, precision=precision,
                                             train_gemm=train_gemm)
        self.store = store
        self.model.dropout_base_seed = (seed * 0x9E3779B1 + 7919 * self.rank) & 0xFFFFFFFF   # ranks draw different masks
        self.opt = TFAdam(self.model, learning_rate, step_boundary=step_boundary, global_step=global_step)
        self.learning_rate = learning_rate
        self.lib = self.model.lib
        self.replicated = [k for k in store.tables if k not in self.sharded]
        if self.dp:
            n_dense = store.dense.numel()
            sizes = [store.tables[k].numel() for k in self.replicated]
            pad4 = lambda n: (n + 3) // 4 * 4      # keep every region 16-byte aligned for the float4 Adam
            self.bucket = torch.zeros(n_dense + sum(pad4(n) for n in sizes), dtype=torch.float32, device=self.device)
            self.model.bind_grad_buffer(self.bucket[:n_dense])
            self.table_grad = {}
            off = n_dense
            for k, n in zip(self.replicated, sizes):
                self.table_grad[k] = self.bucket[off:off + n].view(store.tables[k].shape)
                off += pad4(n)
        self.last_loss = None
        self.last_a2a_bytes = 0           # all-to-all payload of the last step (this rank's sends)
        self.overlap_push = os.environ.get("DMT_DP_OVERLAP", "1") != "0"   # sharded-table push beside the allreduce

    @property
    def allreduce_bytes(self):
        """Payload of the one gradient allreduce per step (dense variables + replicated small tables)."""
        return int(self.bucket.numel() * 4) if self.dp else 0

    # ------------------------------------------------------------------ stage timing (bench hook)
    def enable_stage_timing(self, on=True):
        self.model.enable_stage_timing(on)

    def stage_times_ms(self):
        return self.model.stage_times_ms()

    def _stage(self, name, launches=0):
        return self.model._Stage(self.model, name, launches)

    # ------------------------------------------------------------------ row-sharded forward exchange
    def _exchange(self, staged):
        """Fetch the rows of every sharded table this batch touches; returns {scope: (RowExchange, compact)} and
        installs the compact tables + re-mapped ids."""
        plan, model = self.plan, self.model
        remap = {}
        state = {}
        stream = torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None
        for scope, shard in self.sharded.items():
            tname = next(n for n, t in plan.tables.items() if t.scope == scope)
            lookups = []   # (role, feature, SparseIds, row offset)
            for p in plan.pooled:
                if p.table == tname:
                    lookups.append(("pool", p.feature, model._sparse(staged, p.feature), 0))
            zp = -1 if plan.zero_pad else 0
            for seq in plan.sequences:
                # one compact row per (sequence, lookup): the target item is looked up by every sequence and each
                # of those lookups gets its own gradient row, so the compact rows stay unique per gradient source
                role = "seq%d" % seq.index
                for f, t in enumerate(seq.tables):
                    if t != tname:
                        continue
                    for feat in (seq.user_features[f], seq.item_features[f]):
                        lookups.append((role, feat, model._sparse(staged, feat), zp))
            need = torch.cat([sp.values.to(torch.int64) + off for _, _, sp, off in lookups])
            with self._stage("sku_route"):
                ex = RowExchange(shard, need, group=self.group)
            # bytes this rank puts on the wire for the table this step: row ids out (int64) + rows back + gradient
            # rows out again (fp32 [n, dim] each way); rows it owns itself never leave the GPU
            remote = ex.n_valid - ex.send_counts[shard.rank]
            self.last_a2a_bytes += remote * 8 + 2 * remote * self.store.tables[scope].shape[1] * 4
            full = self.store.tables[scope]
            dim = full.shape[1]

            def gather_local(rows_i32, full=full, dim=dim):
                out = torch.empty(rows_i32.numel(), dim, dtype=torch.float32, device=self.device)
                if rows_i32.numel():
                    abi.check(self.lib.dmt_embed_gather(full.data_ptr(), full.shape[0], dim, rows_i32.data_ptr(),
                                                        rows_i32.numel(), 0, out.data_ptr(), stream))
                    model.launches += 1
                return out

            with self._stage("sku_fetch"):
                compact = ex.fetch(gather_local)
            if compact.shape[0] == 0:   # keep a valid pointer for the kernels
                compact = torch.zeros(1, dim, dtype=torch.float32, device=self.device)
            pos = 0
            for role, feat, sp, off in lookups:
                n = sp.values.numel()
                ids = remap_ids(ex.compact_row[pos:pos + n], role != "pool" and plan.zero_pad, ex.n_valid)
                remap[(role, feat)] = SparseIds(ids, sp.offsets, sp.weights)
                pos += n
            state[scope] = (ex, compact, full)
        return remap, state

    # ------------------------------------------------------------------ one synchronous step
    def train_step(self, inputs, lr=None):
        model, opt = self.model, self.opt
        if not self.dp:
            loss, grads = model.compute_gradients(inputs)
            with self._stage("adam"):
                opt.apply_gradients(grads, lr=lr)
            self.last_loss = loss
            return loss
        staged = dict(model.stage_inputs(inputs))
        self.last_a2a_bytes = 0
        remap, state = self._exchange(staged)
        staged["__remap__"] = remap
        try:
            for scope, (ex, compact, full) in state.items():
                self.store.tables[scope] = compact
            loss, grads = model.compute_gradients(staged)
        finally:
            for scope, (ex, compact, full) in state.items():
                self.store.tables[scope] = full
        stream = torch.cuda.current_stream(self.device).cuda_stream
        inv_world = 1.0 / self.world
        keep = []
        opt.begin_step()
        # sharded tables: per-row gradients back to the owners, then sparse Adam on the shard -- on a side stream:
        # independent of the replicated tables' densify -> allreduce -> Adam below, which runs beside it (all ranks
        # issue the all-to-all before the allreduce, so the collectives keep one order on NCCL's stream).  The
        # per-stage timing pass keeps everything on one stream.
        main = torch.cuda.current_stream(self.device) if self.device.type == "cuda" else None
        side = None
        if main is not None and state and self.model._events is None and self.overlap_push:
            if getattr(self, "_push_stream", None) is None:
                self._push_stream = torch.cuda.Stream(self.device)
            side = self._push_stream
            side.wait_stream(main)
        with (torch.cuda.stream(side) if side is not None else _nullctx()):
            pstream = side.cuda_stream if side is not None else stream
            for scope, (ex, compact, full) in state.items():
                srcs = grads.lookups.get(scope, [])
                dim = full.shape[1]
                with self._stage("sku_push", 2):
                    gc = torch.zeros(max(ex.n_valid, 1), dim, dtype=torch.float32, device=self.device)
                    if srcs and ex.n_valid:
                        keys, refs, scale, arr = self._expand(srcs, ex.n_valid, pstream)
                        abi.check(self.lib.dmt_embed_grad_scatter_rows(len(srcs), arr, keys.data_ptr(), refs.data_ptr(),
                                                                       scale.data_ptr(), keys.numel(), dim, gc.data_ptr(),
                                                                       pstream))
                        keep.append((keys, refs, scale, arr, srcs))
                    recv = ex.push_grads(gc[:ex.n_valid])
                with self._stage("adam_shard"):
                    src = [LookupGrad(ex.recv_rows.to(torch.int32), recv, 0, 0)] if recv.shape[0] else []
                    opt.step_table(scope, src, lr=lr, grad_scale=inv_world, ws_name="sorted_ws_shard")
                keep.append((gc, recv, src))
        # replicated tables: densify into the allreduce bucket (the bucket's dense part was zeroed and filled by
        # compute_gradients; zero the table part here)
        with self._stage("densify", 3):
            n_dense = self.store.dense.numel()
            self.bucket[n_dense:].zero_()
            scopes = [k for k in self.replicated if grads.lookups.get(k)]
            n_src = sum(len(grads.lookups[k]) for k in scopes)
            if scopes and (len(scopes) > abi.MAX_ADAM_TABLES or n_src > abi.MAX_MULTI_GRAD_SOURCES or
                           any(self.store.tables[k].shape[0] > (1 << 24) or self.store.tables[k].shape[1] > 128
                               for k in scopes)):
                for scope in scopes:       # beyond the limits of the multi-table entry points: one table at a time
                    srcs = grads.lookups[scope]
                    table = self.store.tables[scope]
                    keys, refs, scale, arr = self._expand(srcs, table.shape[0], stream)
                    skeys, perm = torch.sort(keys, stable=True)
                    ws = model._scratch("sorted_ws", self.lib.dmt_embed_sorted_workspace_bytes(keys.numel(), table.shape[1]))
                    abi.check(self.lib.dmt_embed_grad_densify_sorted(table.shape[0], table.shape[1], len(srcs), arr,
                                                                     skeys.data_ptr(), perm.data_ptr(), refs.data_ptr(),
                                                                     scale.data_ptr(), keys.numel(), 1.0,
                                                                     self.table_grad[scope].data_ptr(), ws.data_ptr(),
                                                                     ws.numel(), stream))
                    keep.append((keys, refs, scale, arr, skeys, perm, srcs))
            elif scopes:
                # all replicated tables in ONE expand / sort / segmented reduction, written straight into the bucket
                tabs = (abi.AdamTable * len(scopes))()
                sources, owner = [], []
                for i, scope in enumerate(scopes):
                    t = self.store.tables[scope]
                    tabs[i].rows, tabs[i].dim = t.shape[0], t.shape[1]
                    tabs[i].dense_out = self.table_grad[scope].data_ptr()
                    for lg in grads.lookups[scope]:
                        sources.append(lg)
                        owner.append(i)
                arr = (abi.GradSource * len(sources))(*[s_.to_c() for s_ in sources])
                own = (C.c_int32 * len(owner))(*owner)
                total = sum(s_.ids.numel() for s_ in sources)
                buf = model._scratch("densify_multi", total * 16 + 1024)
                keys = buf[:total * 4].view(torch.int32)
                scale = buf[total * 4:total * 8].view(torch.float32)
                refs = buf[total * 8:total * 16].view(torch.int64)
                abi.check(self.lib.dmt_embed_grad_expand_multi(len(scopes), tabs, len(sources), arr, own, keys.data_ptr(),
                                                               refs.data_ptr(), scale.data_ptr(), stream))
                skeys, perm = torch.sort(keys, stable=True)
                ws = model._scratch("sorted_ws", self.lib.dmt_embed_sorted_multi_workspace_bytes(total))
                abi.check(self.lib.dmt_embed_adam_sorted_multi(None, len(scopes), tabs, len(sources), arr,
                                                               skeys.data_ptr(), perm.data_ptr(), refs.data_ptr(),
                                                               scale.data_ptr(), total, 1.0, ws.data_ptr(), ws.numel(),
                                                               stream))
                keep.append((tabs, arr, own, keys, refs, scale, skeys, perm, sources))
        with self._stage("allreduce"):
            if self.world > 1:
                dist.all_reduce(self.bucket, group=self.group)
        with self._stage("adam"):
            opt.step_dense(grads.dense, lr=lr, grad_scale=inv_world)
            cfg = opt._cfg(lr)
            for scope in self.replicated:
                t = self.store.tables[scope]
                abi.check(self.lib.dmt_adam_dense(C.byref(cfg), t.data_ptr(), opt.m_tab[scope].data_ptr(),
                                                  opt.v_tab[scope].data_ptr(), self.table_grad[scope].data_ptr(),
                                                  t.numel(), inv_world, stream))
                model.launches += 1
        if side is not None:
            main.wait_stream(side)
        model.invalidate_prepared()
        self.last_loss = loss
        self._keep = keep
        return loss

    def _expand(self, srcs: List[LookupGrad], rows, stream):
        arr = (abi.GradSource * len(srcs))(*[s.to_c() for s in srcs])
        total = sum(s.ids.numel() for s in srcs)
        keys = torch.empty(total, dtype=torch.int32, device=self.device)
        refs = torch.empty(total, dtype=torch.int64, device=self.device)
        scale = torch.empty(total, dtype=torch.float32, device=self.device)
        abi.check(self.lib.dmt_embed_grad_expand(len(srcs), arr, rows, keys.data_ptr(), refs.data_ptr(),
                                                 scale.data_ptr(), stream))
        self.model.launches += 1
        return keys, refs, scale, arr

    def global_loss(self):
        """Mean of the per-rank batch-mean losses (run_dnn.py:195-201 logs the tower average)."""
        loss = self.last_loss.detach().clone().reshape(1)
        if self.world > 1:
            dist.all_reduce(loss, group=self.group)
            loss /= self.world
        return loss[0]


def _store_to(store: ParamStore, device):
    """Move a CPU-initialised ParamStore to `device` (views rebuilt over the moved buffers)."""
    store.device = torch.device(device)
    store.dense = store.dense.to(device)
    store.tables = {k: v.to(device) for k, v in store.tables.items()}
    store.views = {}
    for s in store.specs:
        store.views[s.name] = store.dense[s.offset:s.offset + s.numel].view(s.shape)
    store.views.update(store.tables)
    return store
