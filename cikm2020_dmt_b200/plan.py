"""Static model plan derived from a `Conf`: tables, behaviour sequences, widths.

Everything here is host-side integer bookkeeping for the hot path of
`DMT_code/model/net/mmoe_transformer_unbias.py`:

* `tables`         one entry per distinct `[embedding] emb` name (base.py:81-91)
* `sequences`      one entry per `attention_embed` group (generate_data, :130-186):
                   the (user_feature, item_feature) pairs in config order, each
                   resolved to its table; concat order == pair order (:153-158,181-182)
* `pooled`         every `emb` entry in config order (embedding_combiner, base.py:93-124)
* `bias_pooled`    every `emb_bias` entry (embedding_combiner_bias, :235-257); these
                   tables are *separate variables* (scope `DnnModel/<Name>/embedding`,
                   not under `embedding_trans`)
* column layout of the MMoE input
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

from . import keys as K


@dataclass
class TableSpec:
    name: str
    rows: int
    dim: int
    scope: str            # TF variable name


@dataclass
class PooledSpec:
    feature: str
    table: str            # key into plan.tables / plan.bias_tables
    dim: int
    col: int              # first output column
    side: str             # 'i' | 'u'


@dataclass
class SeqSpec:
    index: int
    user_features: List[str]
    item_features: List[str]
    tables: List[str]
    dims: List[int]
    col_offsets: List[int]        # column of each feature inside the d_model-wide token
    ts_feature: Optional[str]
    maxlen: int                   # transformer_maxlen_k (size of the position table)
    scope: str                    # TF scope prefix of this sequence's transformer


@dataclass
class ModelPlan:
    d_model: int
    d_ff: int
    num_heads: int
    num_blocks_encode: int
    num_blocks_decode: int
    maxlen_k: int
    maxlen_q: int
    dropout_rate: float
    zero_pad: bool
    position_encoding_method: str
    feature_dim: int
    is_use_feature: bool
    tables: Dict[str, TableSpec] = field(default_factory=dict)
    bias_tables: Dict[str, TableSpec] = field(default_factory=dict)
    sequences: List[SeqSpec] = field(default_factory=list)
    pooled: List[PooledSpec] = field(default_factory=list)
    bias_pooled: List[PooledSpec] = field(default_factory=list)
    pooled_width: int = 0
    bias_width: int = 0
    mmoe_in: int = 0              # feature_dim + pooled_width + n_seq*d_model
    interest_col: int = 0         # first column of the interest vectors inside x
    hidden_units_bottom: List[int] = field(default_factory=list)
    hidden_units_task: List[int] = field(default_factory=list)
    hidden_units_bias: List[int] = field(default_factory=list)
    dropout_rate_bias: List[float] = field(default_factory=list)
    num_experts: int = 4
    num_tasks: int = 2
    output_units: int = 1
    weight_ctr: List[float] = field(default_factory=list)
    weight_ecvr: List[float] = field(default_factory=list)
    loss_weight: List[float] = field(default_factory=list)
    loss_unbias_method: str = "two_head_add"
    loss_ctr_rel_method: str = "ctr"
    learning_rate: List[float] = field(default_factory=lambda: [1e-3])   # [model] learning_rate (comma list)
    step_boundary: List[int] = field(default_factory=list)               # [model] step_boundary

    @property
    def d_k(self):
        return self.d_model // self.num_heads

    def all_id_features(self) -> List[str]:
        names = [p.feature for p in self.pooled]
        for p in self.bias_pooled:
            if p.feature not in names:
                names.append(p.feature)
        return names


class PlanError(ValueError):
    pass


def build_plan(conf) -> ModelPlan:
    model = conf[K.MODEL]
    if conf.model_type != "mmoe_transformer_unbias":
        # Only this model_type is runnable through Inference as shipped.
        raise PlanError("model_type %r is not on the DMT hot path" % conf.model_type)
    if conf.is_trans_input_by_mlp or conf.is_trans_out_concat_item or conf.is_trans_out_by_mlp:
        raise PlanError("transformer_is_trans_input_by_mlp / _out_concat_item / _out_by_mlp "
                        "are off in dmt.conf and not built (mmoe_transformer_unbias.py:197-216)")
    if conf.position_encoding_method not in ("position_learn", "position_sin_cos"):
        raise PlanError("transformer_position_encoding_method=%r: position_learn and position_sin_cos are built; "
                        "time_add / time_concat (TransformerModel.py:71-79) are not" % conf.position_encoding_method)
    if conf.is_decoder_add_pos_emb:
        raise PlanError("transformer_is_decoder_add_pos_emb=true is not built (TransformerModel.py:148-149)")
    if model.get(K.IS_BN) or conf.sim_embed:
        raise PlanError("is_bn / sim_embed are disabled in dmt.conf and not built (base.py:44-63,126-132)")

    plan = ModelPlan(
        d_model=conf.d_model, d_ff=conf.d_ff, num_heads=conf.num_heads,
        num_blocks_encode=conf.num_blocks_encode, num_blocks_decode=conf.num_blocks_decode,
        maxlen_k=conf.maxlen_k, maxlen_q=conf.maxlen_q, dropout_rate=conf.dropout_rate,
        zero_pad=bool(conf.zero_pad), position_encoding_method=conf.position_encoding_method,
        feature_dim=model[K.FEAT_DIM], is_use_feature=bool(conf.is_use_feature),
        hidden_units_bottom=list(model[K.hidden_units_bottom]),
        hidden_units_task=list(model[K.hidden_units_task]),
        hidden_units_bias=list(model[K.HIDDEN_UNITS_BIAS]),
        dropout_rate_bias=list(conf.dropout_rate_bias or []),
        num_experts=model[K.num_experts], output_units=model[K.OUTPUT_UNITS],
        weight_ctr=list(conf.weight_ctr), weight_ecvr=list(conf.weight_ecvr),
        loss_weight=list(conf[K.PARAMETER][K.LOSS_WEIGHT]),
        loss_unbias_method=conf.loss_unbias_method,
        loss_ctr_rel_method=conf.loss_ctr_rel_method,
        learning_rate=list(model.get(K.LEARNING_RATE) or [1e-3]),
        step_boundary=list(model.get(K.STEP_BOUNDARY) or []),
    )
    if plan.d_model % plan.num_heads:
        raise PlanError("transformer_d_model must be divisible by transformer_num_heads")

    feat_to_emb = {}
    col = plan.feature_dim if plan.is_use_feature else 0
    for name, rows, dim, feature, side in conf.embedding_list:
        spec = plan.tables.get(name)
        if spec is None:
            plan.tables[name] = TableSpec(name, rows, dim, "DnnModel/embedding_trans/%s/embedding" % name)
        elif (spec.rows, spec.dim) != (rows, dim):
            raise PlanError("table %s declared with two shapes" % name)
        feat_to_emb.setdefault(feature, (name, dim))
        plan.pooled.append(PooledSpec(feature, name, dim, col, side))
        col += dim
    plan.pooled_width = col - (plan.feature_dim if plan.is_use_feature else 0)
    plan.interest_col = col

    bcol = 0
    for name, rows, dim, feature, side in conf.embedding_list_bias:
        spec = plan.bias_tables.get(name)
        if spec is None:
            plan.bias_tables[name] = TableSpec(name, rows, dim, "DnnModel/%s/embedding" % name)
        elif (spec.rows, spec.dim) != (rows, dim):
            raise PlanError("bias table %s declared with two shapes" % name)
        plan.bias_pooled.append(PooledSpec(feature, name, dim, bcol, side))
        bcol += dim
    plan.bias_width = bcol

    ts_list = conf.attention_embed_seq_ts if conf.is_use_seq_ts else []
    for i, pairs in enumerate(conf.attention_embed_pairs):
        users, items, tables, dims, offs = [], [], [], [], []
        c = 0
        for user_feature, item_feature in pairs:
            if user_feature not in feat_to_emb or item_feature not in feat_to_emb:
                raise PlanError("attention_embed pair %s:%s has no emb entry" % (user_feature, item_feature))
            tname, dim = feat_to_emb[user_feature]
            if feat_to_emb[item_feature] != (tname, dim):
                raise PlanError("pair %s:%s maps to different tables" % (user_feature, item_feature))
            users.append(user_feature)
            items.append(item_feature)
            tables.append(tname)
            dims.append(dim)
            offs.append(c)
            c += dim
        if c != plan.d_model:
            raise PlanError("sequence %d token width %d != transformer_d_model %d" % (i, c, plan.d_model))
        stag = "sequence_%d" % i
        scope = ("DnnModel/embedding_trans/trans_%s/encode_decode_%s/encode_decode_%s" % (stag, stag, stag))
        plan.sequences.append(SeqSpec(i, users, items, tables, dims, offs,
                                      ts_list[i] if i < len(ts_list) else None,
                                      plan.maxlen_k, scope))
    plan.mmoe_in = plan.interest_col + len(plan.sequences) * plan.d_model
    return plan
