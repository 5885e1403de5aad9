"""TF-1 Adam over the DMT parameter store (`tf.train.AdamOptimizer`, inference_mlp.py:272-273;
applied to tower-averaged gradients, run_dnn.py:203-207).

Host side only: owns the m / v buffers (PyTorch tensors are the allocator), builds the POD descriptors
and enqueues the C-ABI kernels.  The one library call on the path is the key sort between
`dmt_embed_grad_expand` and `dmt_embed_adam_sorted` (`torch.sort`, i.e. CUB) -- plumbing, like the
allocator; the arithmetic (segmented gradient reduction + the Adam update) is ours.
"""
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import abi


@dataclass
class LookupGrad:
    """One group of lookups into a table + where their gradient rows live (dmt_grad_source)."""
    ids: torch.Tensor                      # int32 [n]
    grad: torch.Tensor                     # fp32 [rows, ld]
    grad_col: int = 0
    id_offset: int = 0                     # -1: zero-pad path (base.py:87-89)
    offsets: Optional[torch.Tensor] = None  # int32 [B+1]: gradient rows are per sample
    weights: Optional[torch.Tensor] = None
    mean: bool = False

    def to_c(self):
        g = abi.GradSource()
        g.ids = abi.ptr(self.ids)
        g.offsets = abi.ptr(self.offsets)
        g.weights = abi.ptr(self.weights)
        g.grad = abi.ptr(self.grad)
        g.n = self.ids.numel()
        g.grad_ld = self.grad.stride(0)
        g.grad_col = self.grad_col
        g.id_offset = self.id_offset
        g.batch = 0 if self.offsets is None else self.offsets.numel() - 1
        g.mean = 1 if self.mean else 0
        return g


class Gradients(object):
    """What `compute_gradients` hands to the optimizer: `dense` is a flat buffer laid out like
    `ParamStore.dense` (one allreduce), `lookups` maps a table's TF variable name to the lookups that feed it
    (the IndexedSlices of the reference, run_dnn.py:63-72, kept sparse), `views` names the dense slices."""

    def __init__(self, dense, lookups, views):
        self.dense, self.lookups, self.views = dense, lookups, views

    def __getitem__(self, name):
        return self.views[name]

    def table_dense(self, store, name, grad_scale=1.0):
        """Densified gradient of one table (tests / small tables only)."""
        table = store.tables[name]
        out = torch.zeros_like(table, dtype=torch.float64)
        for lg in self.lookups.get(name, []):
            ids = lg.ids.long()
            g = lg.grad[:, lg.grad_col:lg.grad_col + table.shape[1]].double()
            if lg.offsets is not None:
                lens = (lg.offsets[1:] - lg.offsets[:-1]).long()
                seg = torch.repeat_interleave(torch.arange(lens.numel(), device=ids.device), lens)
                w = torch.ones_like(ids, dtype=torch.float64) if lg.weights is None else lg.weights.double()
                if lg.mean:
                    wsum = torch.zeros(lens.numel(), dtype=torch.float64, device=ids.device).index_add_(0, seg, w)
                    w = w / wsum[seg]
                g = g[seg] * w[:, None]
            rows = ids + lg.id_offset
            ok = (rows >= 0) & (rows < table.shape[0])
            out.index_add_(0, rows[ok], g[ok])
        return out * grad_scale


def piecewise_constant(step, boundaries, values):
    """`tf.train.piecewise_constant(global_step, step_boundary, learning_rate)` (run_dnn.py:125-126):
    values[0] while step <= boundaries[0], values[i] while boundaries[i-1] < step <= boundaries[i],
    values[-1] beyond the last boundary.  TF requires len(values) == len(boundaries) + 1."""
    boundaries, values = list(boundaries), list(values)
    if len(values) != len(boundaries) + 1:
        raise ValueError("piecewise_constant: %d learning rates need %d step boundaries, got %d"
                         % (len(values), len(values) - 1, len(boundaries)))
    for b, v in zip(boundaries, values):
        if step <= b:
            return float(v)
    return float(values[-1])


class TFAdam(object):
    KIND = abi.OPT_ADAM
    SLOT_NAMES = ("Adam", "Adam_1")        # TF slot names of m / v

    def __init__(self, model, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8, step_boundary=None,
                 global_step=0):
        """learning_rate: a float, or the conf's comma list (`[model] learning_rate = 0.001,0.0001`) together
        with `step_boundary` (`[model] step_boundary`, default: the model's conf): the rate of a step is
        `piecewise_constant(global_step, step_boundary, learning_rate)` evaluated BEFORE the step increments
        `global_step` (run_dnn.py:119-126; `global_step` starts at the resumed checkpoint's step while the Adam
        beta powers `t` restart, run_dnn.py:296-306)."""
        self.model = model
        self.store = model.params
        self.lib = model.lib
        self.learning_rate = learning_rate
        if step_boundary is None and isinstance(learning_rate, (list, tuple)) and len(learning_rate) > 1:
            plan = getattr(model, "plan", None)
            step_boundary = getattr(plan, "step_boundary", None)
        if isinstance(learning_rate, (list, tuple)):
            if len(learning_rate) == 1 and not step_boundary:
                step_boundary = []
            if step_boundary is None:
                raise ValueError("a list of learning rates needs step_boundary (run_dnn.py:125-126)")
            piecewise_constant(0, step_boundary, learning_rate)      # validates the lengths
        self.step_boundary = list(step_boundary) if step_boundary is not None else None
        self.global_step = int(global_step)
        self._lr_now = None
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self.t = 0
        dev = self.store.dense.device
        self.m_dense = torch.zeros_like(self.store.dense)
        self.v_dense = torch.zeros_like(self.store.dense)
        self.m_tab = {k: torch.zeros_like(t) for k, t in self.store.tables.items()}
        self.v_tab = {k: torch.zeros_like(t) for k, t in self.store.tables.items()}
        # the "row got a gradient this step" marks of all tables live in ONE buffer (cleared with one memset)
        sizes = {k: (t.shape[0] + 255) // 256 * 256 for k, t in self.store.tables.items()}
        self._touched_all = torch.zeros(sum(sizes.values()), dtype=torch.uint8, device=dev)
        self.touched, off = {}, 0
        for k, t in self.store.tables.items():
            self.touched[k] = self._touched_all[off:off + t.shape[0]]
            off += sizes[k]
        self._multi = None            # cached descriptors of the multi-table path

    def current_lr(self, step=None):
        """Learning rate of the step that starts at `global_step` (default: the next one)."""
        lr = self.learning_rate
        if isinstance(lr, (list, tuple)):
            return piecewise_constant(self.global_step if step is None else step, self.step_boundary, lr)
        return float(lr)

    def _cfg(self, lr):
        if lr is None:
            lr = self._lr_now if self._lr_now is not None else self.current_lr()
        elif isinstance(lr, (list, tuple)):
            raise TypeError("pass a scalar learning rate per step; schedules belong to the constructor")
        return abi.AdamCfg(float(lr), self.beta1, self.beta2, self.epsilon, int(self.t), self.KIND)

    def begin_step(self):
        self._lr_now = self.current_lr()       # evaluated at the pre-increment global_step, like the TF graph
        self.global_step += 1
        self.t += 1

    def apply_gradients(self, grads, lr=None, grad_scale=1.0):
        """`optimizer.apply_gradients` (run_dnn.py:203-207): one TF-Adam step over every variable; tables
        without a lookup this step still decay (dense semantics).  All embedding tables go through ONE expand /
        sort / segmented-Adam / untouched-rows sequence (dmt_*_multi); a table too large for the packed keys, or
        more tables / lookups groups than the multi entry points take, falls back to the per-table calls."""
        self.begin_step()
        self.step_dense(grads.dense, lr=lr, grad_scale=grad_scale)
        names = list(self.store.tables)
        n_src = sum(len(grads.lookups.get(k, [])) for k in names)
        if (len(names) <= abi.MAX_ADAM_TABLES and 0 < n_src <= abi.MAX_MULTI_GRAD_SOURCES
                and all(self.store.tables[k].shape[0] <= (1 << 24) and self.store.tables[k].shape[1] <= 128 for k in names)):
            self.step_tables_multi(names, grads.lookups, lr=lr, grad_scale=grad_scale)
            return
        for name in names:
            self.step_table(name, grads.lookups.get(name, []), lr=lr, grad_scale=grad_scale)

    def step_tables_multi(self, names, lookups, lr=None, grad_scale=1.0):
        """Every table of `names` in one pass: keys = (table index << 24) | row."""
        cfg = self._cfg(lr)
        dev = self.store.dense.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        if self._multi is None or self._multi[0] != tuple(names):
            tabs = (abi.AdamTable * len(names))()
            for i, k in enumerate(names):
                t = self.store.tables[k]
                tabs[i].table, tabs[i].m, tabs[i].v = t.data_ptr(), self.m_tab[k].data_ptr(), self.v_tab[k].data_ptr()
                tabs[i].touched, tabs[i].rows, tabs[i].dim = self.touched[k].data_ptr(), t.shape[0], t.shape[1]
            self._multi = (tuple(names), tabs)
        tabs = self._multi[1]
        for i, k in enumerate(names):          # the tables may have been re-bound (row-sharded exchange)
            tabs[i].table = self.store.tables[k].data_ptr()
        sources, owner = [], []
        for i, k in enumerate(names):
            for lg in lookups.get(k, []):
                sources.append(lg)
                owner.append(i)
        arr = (abi.GradSource * len(sources))(*[s.to_c() for s in sources])
        own = (C.c_int32 * len(owner))(*owner)
        total = sum(s.ids.numel() for s in sources)
        buf = self.model._scratch("adam_multi", total * 16 + 1024)
        keys = buf[:total * 4].view(torch.int32)
        scale = buf[total * 4:total * 8].view(torch.float32)
        refs = buf[total * 8:total * 16].view(torch.int64)
        abi.check(self.lib.dmt_embed_grad_expand_multi(len(names), tabs, len(sources), arr, own, keys.data_ptr(),
                                                       refs.data_ptr(), scale.data_ptr(), stream))
        skeys, perm = torch.sort(keys, stable=True)
        ws = self.model._scratch("sorted_ws", self.lib.dmt_embed_sorted_multi_workspace_bytes(total))
        abi.check(self.lib.dmt_embed_adam_sorted_multi(C.byref(cfg), len(names), tabs, len(sources), arr, skeys.data_ptr(),
                                                       perm.data_ptr(), refs.data_ptr(), scale.data_ptr(), total,
                                                       float(grad_scale), ws.data_ptr(), ws.numel(), stream))
        abi.check(self.lib.dmt_adam_rows_untouched_multi(C.byref(cfg), len(names), tabs, stream))
        self._touched_all.zero_()
        self.model.launches += 5
        self._keep = (arr, own, keys, refs, scale, skeys, perm, sources)
        self.model.invalidate_prepared()

    def step_dense(self, grad_flat, lr=None, grad_scale=1.0):
        cfg = self._cfg(lr)
        stream = torch.cuda.current_stream(grad_flat.device).cuda_stream
        self.model.launches += 1
        abi.check(self.lib.dmt_adam_dense(C.byref(cfg), self.store.dense.data_ptr(), self.m_dense.data_ptr(),
                                          self.v_dense.data_ptr(), grad_flat.data_ptr(), grad_flat.numel(),
                                          float(grad_scale), stream))

    def step_table(self, name, sources: List[LookupGrad], lr=None, grad_scale=1.0, ws_name="sorted_ws"):
        """`name` is the TF variable name of the table; `sources` every lookup of this step into it.  ws_name: the
        scratch buffer of the segmented reduction (a caller that runs this on a side stream names its own)."""
        table = self.store.tables[name]
        cfg = self._cfg(lr)
        stream = torch.cuda.current_stream(table.device).cuda_stream
        rows, dim = table.shape
        if sources:
            arr = (abi.GradSource * len(sources))(*[s.to_c() for s in sources])
            total = sum(s.ids.numel() for s in sources)
            keys = torch.empty(total, dtype=torch.int32, device=table.device)
            refs = torch.empty(total, dtype=torch.int64, device=table.device)
            scale = torch.empty(total, dtype=torch.float32, device=table.device)
            abi.check(self.lib.dmt_embed_grad_expand(len(sources), arr, rows, keys.data_ptr(), refs.data_ptr(),
                                                     scale.data_ptr(), stream))
            skeys, perm = torch.sort(keys, stable=True)
            ws = self.model._scratch(ws_name, self.lib.dmt_embed_sorted_workspace_bytes(total, dim))
            abi.check(self.lib.dmt_embed_adam_sorted(C.byref(cfg), table.data_ptr(), self.m_tab[name].data_ptr(),
                                                     self.v_tab[name].data_ptr(), rows, dim, len(sources), arr,
                                                     skeys.data_ptr(), perm.data_ptr(), refs.data_ptr(),
                                                     scale.data_ptr(), total, float(grad_scale),
                                                     self.touched[name].data_ptr(), ws.data_ptr(), ws.numel(),
                                                     stream))
            self.model.launches += 3
            self._keep = (arr, keys, refs, scale, skeys, perm, sources)
        abi.check(self.lib.dmt_adam_rows_untouched(C.byref(cfg), table.data_ptr(), self.m_tab[name].data_ptr(),
                                                   self.v_tab[name].data_ptr(), rows, dim,
                                                   self.touched[name].data_ptr(), stream))
        self.model.launches += 1
        self.model.invalidate_prepared()


class TFGradientDescent(TFAdam):
    """`tf.train.GradientDescentOptimizer(learning_rate)` (inference_mlp.py:266-267): theta -= lr * g.  Same host
    path and kernels as TFAdam (`dmt_adam_cfg.kind = DMT_OPT_SGD`); rows without a gradient do not move, so the
    untouched-rows pass is a no-op and the moment buffers stay unused."""
    KIND = abi.OPT_SGD
    SLOT_NAMES = (None, None)


class TFAdagrad(TFAdam):
    """`tf.train.AdagradOptimizer(learning_rate)` (inference_mlp.py:270-271): acc += g^2 (initial accumulator 0.1, TF's
    default), theta -= lr * g / sqrt(acc).  The accumulator lives in the `m` buffers."""
    KIND = abi.OPT_ADAGRAD
    SLOT_NAMES = ("Adagrad", None)

    def __init__(self, model, learning_rate, initial_accumulator_value=0.1, **kw):
        super().__init__(model, learning_rate, **kw)
        self.m_dense.fill_(initial_accumulator_value)
        for t in self.m_tab.values():
            t.fill_(initial_accumulator_value)
