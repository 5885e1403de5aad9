"""Config-key vocabulary of the reference's INI surface.

The reference spells every section / option name as a module constant in
`DMT_code/util/util.py:5-113` and indexes `Conf` with them
(`wnd_conf[MODEL][BATCH_SIZE]`).  A drop-in has to accept the same spellings,
so the names below are the reference's; the table is generated from one
mapping instead of being written out line by line.
"""

_SECTIONS = {
    "INFO": "info",
    "PARAMETER": "parameter",
    "EXPORT_MODEL": "export_model",
    "MODEL": "model",
    "PATH": "path",
    "CLASS_WEIGHT": "class_weight",
    "SCHEMA": "schema",
    "EMBEDDING": "embedding",
    "ONLINE": "online_learning",
}

# constant-name -> option string (util.py:5-113).  Lower-case constant names are
# lower-case in the reference too (util.py:22-25,30-32,56-72,103).
_OPTIONS = {
    "TYPE": "type",
    "LABEL_WEIGHT": "label_weight",
    "LOSS_WEIGHT": "loss_weight",
    "LOSS_WEIGHT_METHOD": "loss_weight_method",
    "EXPORT_WEIGHT": "export_weight",
    "FEAT_DIM": "feature_dimension",
    "OUTPUT_UNITS": "output_units",
    "HIDDEN_UNITS": "hidden_units",
    "IS_USE_FEATURE": "is_use_feature",
    "HIDDEN_UNITS_BIAS": "hidden_units_bias",
    "loss_unbias_method": "loss_unbias_method",
    "LOSS_CTR_REL_METHOD": "loss_ctr_rel_method",
    "propensity_em": "propensity_em",
    "propensity_em_type": "propensity_em_type",
    "MODEL_TYPE": "model_type",
    "ENABLE_SSP": "enable_ssp",
    "LEARNING_RATE": "learning_rate",
    "STEP_BOUNDARY": "step_boundary",
    "OPTIMIZER": "optimizer",
    "hidden_units_bottom": "hidden_units_bottom",
    "hidden_units_task": "hidden_units_task",
    "num_experts": "num_experts",
    "DROPOUT": "dropout",
    "dropout_rate_bias": "dropout_rate_bias",
    "DROPOUT_BOTTOM": "dropout_bottom",
    "DROPOUT_TASK": "dropout_task",
    "EPOCH_NUM": "epoch_num",
    "BATCH_SIZE": "batch_size",
    "SHUFFLE_SIZE": "shuffle_size",
    "TEST_BATCH_SIZE": "test_batch_size",
    "VALIDATION_BATCH_SIZE": "validation_batch_size",
    "DEVICE": "device",
    "GPU_VISIBLE": "gpu_visible",
    "VALIDATE_STEP": "validate_step",
    "IS_BN": "is_bn",
    "BN_DECAY": "bn_decay",
    "IS_DROPOUT": "is_dropout",
    "FILTER_SHAPE": "filter_shape",
    "MAX_ITER_STEP": "max_iter_step",
    "TOTAL_EXAMPLE_NUM": "total_example_num",
    "SAVE_CKPT_NUMS": "save_ckpt_nums",
    "WND_WD": "wnd_wd",
    "L2_EMB_LAMBDA": "l2_emb_lambda",
    "zero_pad": "zero_pad",
    "TRAIN_DATA_PATH": "train_data_path",
    "TEST_DATA_PATH": "test_data_path",
    "TEST_DATA_PATH_ORD": "test_data_path_ord",
    "TRAIN_DATA_MEAN_PATH": "train_data_mean_path",
    "TRAIN_DATA_STD_PATH": "train_data_std_path",
    "TRAIN_DATA_STAT_PATH": "train_data_stat_path",
    "TRAIN_RESULT": "train_result",
    "TEST_RESULT": "test_result",
    "VALIDATION_DATA_PATH": "validation_data_path",
    "VALIDATION_RESULT": "validation_result",
    "OUTPUT_PATH": "output_path",
    "SUMMARY_PATH": "summary_path",
    "MODEL_PATH": "model_path",
    "MODEL_FROZEN_PATH": "model_frozen_path",
    "MODEL_IMP_PATH": "model_imp_path",
    "TRAIN_WEIGHT": "train_weight",
    "VALID_WEIGHT": "valid_weight",
    "WEIGHT_CTR": "weight_ctr",
    "WEIGHT_ECVR": "weight_ecvr",
    "HEADER_SCHEMA": "header_schema",
    "EMB": "emb",
    "EMB_BIAS": "emb_bias",
    "ATTENTION_EMBED": "attention_embed",
    "attention_embed_seq_ts": "attention_embed_seq_ts",
    "SIM_EMBED": "sim_embed",
    "UPDATE_EMB": "update_emb",
    "MIN_TRAIN_EXA_NUMS": "min_train_exa_num",
    "MAX_RATIO": "data_max_ratio",
    "MIN_RATIO": "data_min_ratio",
    "TO_ADDRS": "email_to_addrs",
    "SUBJECT": "email_subject",
}

# transformer_* options are spelled identically as constant and as option
# (util.py:59-71).
for _k in ("d_model", "d_ff", "num_heads", "num_blocks_encode", "num_blocks_decode",
           "maxlen_k", "maxlen_q", "dropout_rate", "is_trans_input_by_mlp",
           "position_encoding_method", "is_trans_out_concat_item",
           "is_trans_out_by_mlp", "is_decoder_add_pos_emb"):
    _OPTIONS["transformer_" + _k] = "transformer_" + _k

globals().update(_SECTIONS)
globals().update(_OPTIONS)

__all__ = sorted(list(_SECTIONS) + list(_OPTIONS)) + [
    "str_to_bool", "csv_to_int_list", "csv_to_float_list", "parse_weight"]


def str_to_bool(s):
    """util.py:116-117: only these five spellings are true."""
    return s in ("True", "true", "yes", "TRUE", "1")


def csv_to_int_list(s):
    """util.py:124-125."""
    return [int(tok) for tok in s.strip().split(",")]


def csv_to_float_list(s):
    """util.py:128-129."""
    return [float(tok) for tok in s.strip().split(",")]


def parse_weight(s):
    """`label:weight,...` -> weights ordered by ascending label (util.py:132-144)."""
    table = {}
    for item in s.split(","):
        label, weight = item.split(":")[:2]
        table[int(label)] = float(weight)
    return [table[label] for label in sorted(table)]
