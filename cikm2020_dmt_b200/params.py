"""Parameter inventory of the DMT hot path: names, shapes, initialisers, storage.

Names are the TF-1 variable names the reference's scope nesting produces
(SURVEY 8a "Parameter names/shapes"), so a checkpoint key surface exists:

* tables                  base.py:81-91              Xavier-uniform over [V, D]
* position table          TransformerModel_util.py:302  Xavier-uniform [maxlen_k, d]
* attention dense x3      TransformerModel_util.py:188-190  glorot-uniform kernel, zero bias
* LayerNorm               TransformerModel_util.py:73-74    beta 0, gamma 1
* feed-forward            TransformerModel_util.py:222-226  glorot-uniform, zero bias;
                          **shared by encoder block i and decoder block i** (same scope,
                          TransformerModel.py:107,121,155,168)
* MMoE experts/gates/towers  base.py:28-37  truncated-normal(0.1) weights, bias 0.1
* bias net                mmoe_transformer_unbias.py:263-287  glorot-uniform, zero bias

Storage: every non-table parameter lives in ONE flat fp32 buffer (`dense`) so the
data-parallel allreduce and the Adam pass are single launches; tables are separate
`[V, D]` tensors (the Sku table is the only large one).
"""
import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch

ALIGN = 64  # floats; every parameter starts on a 256-byte boundary


@dataclass
class ParamSpec:
    name: str
    shape: Tuple[int, ...]
    init: str   # xavier | zeros | ones | const:<v> | trunc_normal:<std>
    offset: int = -1   # float offset inside the flat dense buffer (tables: -1)

    @property
    def numel(self):
        n = 1
        for s in self.shape:
            n *= s
        return n


POS_LEARN = "/positional_encoding_k_position_learn/embedding_position_learn"
POS_SIN_COS = "/positional_encoding_k_position_sin_cos/position_enc"     # a constant, not a TF variable


def sin_cos_table(maxlen, d):
    """positional_encoding (TransformerModel_util.py:237-278): pos / 10000^((i - i % 2) / E), sin on the even
    columns, cos on the odd ones; built in fp64 like the reference's numpy code, then cast to fp32."""
    import numpy as np
    i = np.arange(d)
    enc = np.arange(maxlen)[:, None] / np.power(10000.0, (i - i % 2) / float(d))[None, :]
    enc[:, 0::2] = np.sin(enc[:, 0::2])
    enc[:, 1::2] = np.cos(enc[:, 1::2])
    return torch.from_numpy(enc.astype(np.float32))


def seq_param_specs(plan, seq) -> List[ParamSpec]:
    d, dff, S = plan.d_model, plan.d_ff, seq.scope
    out = []
    if plan.position_encoding_method == "position_learn":        # position_sin_cos has no variable (:61-65)
        out.append(ParamSpec(S + POS_LEARN, (plan.maxlen_k, d), "xavier"))

    def attn(block, kind):
        base = "%s/num_blocks_%d/%s" % (S, block, kind)
        for sfx in ("dense", "dense_1", "dense_2"):          # Q, K, V in creation order
            out.append(ParamSpec("%s/%s/kernel" % (base, sfx), (d, d), "xavier"))
            out.append(ParamSpec("%s/%s/bias" % (base, sfx), (d,), "zeros"))
        out.append(ParamSpec(base + "/ln/beta", (d,), "zeros"))
        out.append(ParamSpec(base + "/ln/gamma", (d,), "ones"))

    for b in range(plan.num_blocks_encode):
        attn(b, "self-attention")
    for b in range(plan.num_blocks_decode):
        attn(b, "vanilla_attention")
    for b in range(max(plan.num_blocks_encode, plan.num_blocks_decode)):
        base = "%s/num_blocks_%d/positionwise_feedforward" % (S, b)
        out.append(ParamSpec(base + "/dense/kernel", (d, dff), "xavier"))
        out.append(ParamSpec(base + "/dense/bias", (dff,), "zeros"))
        out.append(ParamSpec(base + "/dense_1/kernel", (dff, d), "xavier"))
        out.append(ParamSpec(base + "/dense_1/bias", (d,), "zeros"))
        out.append(ParamSpec(base + "/ln/beta", (d,), "zeros"))
        out.append(ParamSpec(base + "/ln/gamma", (d,), "ones"))
    return out


TASK_NAMES = ("click", "order")   # mmoe_transformer_unbias.py:301


def dense_param_specs(plan) -> List[ParamSpec]:
    specs: List[ParamSpec] = []
    for seq in plan.sequences:
        specs += seq_param_specs(plan, seq)
    for e in range(plan.num_experts):
        fan_in = plan.mmoe_in
        for l, units in enumerate(plan.hidden_units_bottom):
            base = "DnnModel/mmoe_layers/expert-%d/expert-layer-%d" % (e, l)
            specs.append(ParamSpec(base + "/weights", (fan_in, units), "trunc_normal:0.1"))
            specs.append(ParamSpec(base + "/biases", (units,), "const:0.1"))
            fan_in = units
    for t in range(plan.num_tasks):
        base = "DnnModel/mmoe_layers/gates-%d/gates-layer-0" % t
        specs.append(ParamSpec(base + "/weights", (plan.mmoe_in, plan.num_experts), "trunc_normal:0.1"))
        specs.append(ParamSpec(base + "/biases", (plan.num_experts,), "const:0.1"))
    for t in range(plan.num_tasks):
        name = TASK_NAMES[t]
        fan_in = plan.hidden_units_bottom[-1]
        for l, units in enumerate(plan.hidden_units_task):
            base = "DnnModel/%s/%s-fc-%d" % (name, name, l)
            specs.append(ParamSpec(base + "/weights", (fan_in, units), "trunc_normal:0.1"))
            specs.append(ParamSpec(base + "/biases", (units,), "const:0.1"))
            fan_in = units
        base = "DnnModel/%s/%s-output" % (name, name)
        specs.append(ParamSpec(base + "/weights", (fan_in, 1), "trunc_normal:0.1"))
        specs.append(ParamSpec(base + "/biases", (1,), "const:0.1"))
    fan_in = plan.bias_width
    for l, units in enumerate(list(plan.hidden_units_bias) + [plan.output_units]):
        specs.append(ParamSpec("DnnModel/layer_bias%d/kernel" % l, (fan_in, units), "xavier"))
        specs.append(ParamSpec("DnnModel/layer_bias%d/bias" % l, (units,), "zeros"))
        fan_in = units
    off = 0
    for s in specs:
        s.offset = off
        off += (s.numel + ALIGN - 1) // ALIGN * ALIGN
    return specs


def dense_numel(specs) -> int:
    last = specs[-1]
    return last.offset + (last.numel + ALIGN - 1) // ALIGN * ALIGN


def table_specs(plan) -> List[ParamSpec]:
    out = [ParamSpec(t.scope, (t.rows, t.dim), "xavier") for t in plan.tables.values()]
    out += [ParamSpec(t.scope, (t.rows, t.dim), "xavier") for t in plan.bias_tables.values()]
    return out


def _fill(t: torch.Tensor, spec: ParamSpec, gen: torch.Generator):
    kind = spec.init
    if kind == "zeros":
        t.zero_()
    elif kind == "ones":
        t.fill_(1.0)
    elif kind.startswith("const:"):
        t.fill_(float(kind.split(":")[1]))
    elif kind == "xavier":
        # glorot/xavier uniform: limit sqrt(6/(fan_in+fan_out)), fan = (shape[0], shape[1])
        limit = math.sqrt(6.0 / (spec.shape[0] + spec.shape[1]))
        t.uniform_(-limit, limit, generator=gen)
    elif kind.startswith("trunc_normal:"):
        std = float(kind.split(":")[1])
        # TF truncated_normal: resample outside +-2 sigma (base.py:31)
        torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
    else:
        raise ValueError(kind)


class ParamStore:
    """All trainable state of the path on one device.

    `dense`   flat fp32 buffer holding every non-table parameter
    `tables`  {TF variable name: [V, D] fp32 tensor}
    `views`   {TF variable name: view} over both
    """

    def __init__(self, plan, device="cpu", seed=20201019, init=True, table_rows_override=None, row_shards=None):
        """row_shards: {TF variable name: (lo, hi)} -- keep only rows [lo, hi) of that table on this rank (the
        values are those of the full table's initialisation, so N ranks together hold exactly the 1-GPU table)."""
        self.plan = plan
        self.device = torch.device(device)
        self.specs = dense_param_specs(plan)
        self.table_specs = table_specs(plan)
        if table_rows_override:
            for s in self.table_specs:
                if s.name in table_rows_override:
                    s.shape = (table_rows_override[s.name], s.shape[1])
        gen = torch.Generator(device="cpu")
        gen.manual_seed(seed)
        dense = torch.zeros(dense_numel(self.specs), dtype=torch.float32)
        self.views: Dict[str, torch.Tensor] = {}
        tables: Dict[str, torch.Tensor] = {}
        for s in self.table_specs:
            t = torch.empty(s.shape, dtype=torch.float32)
            if init:
                _fill(t, s, gen)
            if row_shards and s.name in row_shards:
                lo, hi = row_shards[s.name]
                t = t[lo:hi].clone()
            tables[s.name] = t
        for s in self.specs:
            v = dense[s.offset:s.offset + s.numel].view(s.shape)
            if init:
                _fill(v, s, gen)
        self.dense = dense.to(self.device)
        self.tables = {k: v.to(self.device) for k, v in tables.items()}
        for s in self.specs:
            self.views[s.name] = self.dense[s.offset:s.offset + s.numel].view(s.shape)
        self.views.update(self.tables)
        # constants of the graph that are not variables (never saved, never updated): the sinusoid position table
        self.consts: Dict[str, torch.Tensor] = {}
        if plan.position_encoding_method == "position_sin_cos":
            tab = sin_cos_table(plan.maxlen_k, plan.d_model).to(self.device)
            for seq in plan.sequences:
                self.consts[seq.scope + POS_SIN_COS] = tab

    def position_table(self, seq):
        """[maxlen_k, d_model] added to the scaled token embeddings of `seq` (TransformerModel.py:61-69)."""
        if self.plan.position_encoding_method == "position_sin_cos":
            return self.consts[seq.scope + POS_SIN_COS]
        return self.views[seq.scope + POS_LEARN]

    def named_parameters(self):
        return self.views.items()

    def __getitem__(self, name):
        return self.views[name]

    def table(self, table_name, bias=False):
        spec = (self.plan.bias_tables if bias else self.plan.tables)[table_name]
        return self.tables[spec.scope]

    def randomize_(self, seed=1, scale=0.05):
        """Test helper: perturb biases / LN params away from their constant inits so
        parity tests exercise every term."""
        gen = torch.Generator(device="cpu")
        gen.manual_seed(seed)
        for s in self.specs:
            if s.init in ("zeros", "ones") or s.init.startswith("const:"):
                noise = torch.empty(s.shape, dtype=torch.float32).normal_(0, scale, generator=gen)
                self.views[s.name].add_(noise.to(self.device))
        return self

    def state_dict(self):
        return {k: v.detach().cpu().clone() for k, v in self.views.items()}

    def load_state_dict(self, sd):
        for k, v in sd.items():
            self.views[k].copy_(v.to(self.views[k].device).reshape(self.views[k].shape))
