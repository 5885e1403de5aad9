"""ctypes binding of `libdmt_b200.so` (include/dmt_b200.h).

The structures mirror the header field for field.  Loading fails loudly: there is no CPU or
PyTorch fallback behind this module -- if the shared library is missing the product path
raises.
"""
import ctypes as C
import os

from . import build as _build

MAX_SEQ_FEATS, MAX_BLOCKS, MAX_POOL_FEATS = 8, 4, 64
MAX_EXPERTS, MAX_TASKS, MAX_LAYERS, MAX_SEQ_LEN = 8, 4, 4, 64
PRECISION_F32, PRECISION_BF16, PRECISION_BF16X3, PRECISION_TF32 = 0, 1, 2, 3
ABI_VERSION = 6

_fp = C.c_void_p   # device pointers travel as integers


class Dense(C.Structure):
    _fields_ = [("w", _fp), ("b", _fp)]


class LayerNorm(C.Structure):
    _fields_ = [("gamma", _fp), ("beta", _fp)]


class AttnWeights(C.Structure):
    _fields_ = [("q", Dense), ("k", Dense), ("v", Dense), ("ln", LayerNorm)]


class FFWeights(C.Structure):
    _fields_ = [("w1", Dense), ("w2", Dense), ("ln", LayerNorm)]


SEQ_DEFER_TAIL = 1
SEQ_LEN_EXACT = 2
SEQ_OUT_BF16 = 4
MAX_TAIL_SEQS = 4


class SeqCfg(C.Structure):
    _fields_ = [("batch", C.c_int32), ("d_model", C.c_int32), ("d_ff", C.c_int32), ("num_heads", C.c_int32),
                ("n_enc_blocks", C.c_int32), ("n_dec_blocks", C.c_int32), ("maxlen", C.c_int32),
                ("zero_pad", C.c_int32), ("n_feats", C.c_int32), ("precision", C.c_int32),
                ("slot_len", C.c_int32), ("flags", C.c_int32), ("dropout_rate", C.c_float),
                ("dropout_seed", C.c_uint32)]


class SeqInput(C.Structure):
    _fields_ = [("table", _fp * MAX_SEQ_FEATS), ("rows", C.c_int64 * MAX_SEQ_FEATS),
                ("dim", C.c_int32 * MAX_SEQ_FEATS), ("_pad", C.c_int32 * MAX_SEQ_FEATS),
                ("ids", _fp * MAX_SEQ_FEATS), ("offsets", _fp * MAX_SEQ_FEATS),
                ("item_ids", _fp * MAX_SEQ_FEATS)]


class SeqWeights(C.Structure):
    _fields_ = [("pos", _fp), ("enc_attn", AttnWeights * MAX_BLOCKS), ("dec_attn", AttnWeights * MAX_BLOCKS),
                ("ff", FFWeights * MAX_BLOCKS)]


class PoolFeat(C.Structure):
    _fields_ = [("table", _fp), ("rows", C.c_int64), ("ids", _fp), ("offsets", _fp), ("weights", _fp),
                ("dim", C.c_int32), ("out_col", C.c_int32)]


class MmoeCfg(C.Structure):
    _fields_ = [("batch", C.c_int32), ("in_dim", C.c_int32), ("n_experts", C.c_int32), ("n_layers", C.c_int32),
                ("units", C.c_int32 * MAX_LAYERS), ("n_tasks", C.c_int32), ("n_tower_layers", C.c_int32),
                ("tower_units", C.c_int32 * MAX_LAYERS), ("precision", C.c_int32)]


class MmoeWeights(C.Structure):
    _fields_ = [("expert", (Dense * MAX_LAYERS) * MAX_EXPERTS), ("gate", Dense * MAX_TASKS),
                ("tower", (Dense * MAX_LAYERS) * MAX_TASKS), ("tower_out", Dense * MAX_TASKS)]


class BiasLossCfg(C.Structure):
    _fields_ = [("batch", C.c_int32), ("in_dim", C.c_int32), ("n_hidden", C.c_int32),
                ("units", C.c_int32 * MAX_LAYERS), ("two_head_multiply", C.c_int32), ("ctr_rel", C.c_int32),
                ("weight_ctr", C.c_float * 5), ("weight_ecvr", C.c_float * 5), ("loss_weight", C.c_float * 2),
                ("dropout_rate", C.c_float * MAX_LAYERS), ("dropout_seed", C.c_uint32)]


class BiasWeights(C.Structure):
    _fields_ = [("layer", Dense * (MAX_LAYERS + 1))]


class FwdFeature(C.Structure):
    _fields_ = [("ids", _fp), ("offsets", _fp), ("weights", _fp)]


class FwdDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("n_seq", C.c_int32), ("n_pool", C.c_int32), ("n_bias_pool", C.c_int32),
                ("feature_dim", C.c_int32), ("is_predict", C.c_int32), ("interest_col", C.c_int32), ("_pad", C.c_int32),
                ("pool", _fp), ("pool_feature", _fp), ("bias_pool", _fp), ("bias_pool_feature", _fp),
                ("seq_cfg", _fp * MAX_TAIL_SEQS), ("seq_in", _fp * MAX_TAIL_SEQS),
                ("seq_user_feature", _fp * MAX_TAIL_SEQS), ("seq_item_feature", _fp * MAX_TAIL_SEQS),
                ("seq_w", _fp * MAX_TAIL_SEQS), ("seq_ws", _fp * MAX_TAIL_SEQS),
                ("seq_ws_bytes", C.c_size_t * MAX_TAIL_SEQS),
                ("mmoe_cfg", _fp), ("mmoe_w", _fp), ("mmoe_ws", _fp), ("mmoe_ws_bytes", C.c_size_t),
                ("mmoe_prepared", _fp), ("bias_cfg", _fp), ("bias_w", _fp), ("bias_in", _fp), ("bias_ld", C.c_int64),
                ("xb", _fp), ("xb_ld", C.c_int64), ("inputs_ready", _fp)]


MAX_WIDEN = 64


class WidenDesc(C.Structure):
    _fields_ = [("src", _fp), ("dst", _fp), ("n", C.c_int64)]


class WidenIdsDesc(C.Structure):
    _fields_ = [("src", _fp), ("dst", _fp), ("n", C.c_int64), ("bytes", C.c_int32), ("_pad", C.c_int32)]


class AdamCfg(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float),
                ("step", C.c_int32), ("kind", C.c_int32)]


OPT_ADAM, OPT_SGD, OPT_ADAGRAD = 0, 1, 2


class GradSource(C.Structure):
    _fields_ = [("ids", _fp), ("offsets", _fp), ("weights", _fp), ("grad", _fp), ("n", C.c_int64),
                ("grad_ld", C.c_int64), ("grad_col", C.c_int32), ("id_offset", C.c_int32), ("batch", C.c_int32),
                ("mean", C.c_int32)]

MAX_GRAD_SOURCES = 16
MAX_ADAM_TABLES, MAX_MULTI_GRAD_SOURCES = 16, 64


class AdamTable(C.Structure):
    _fields_ = [("table", _fp), ("m", _fp), ("v", _fp), ("touched", _fp), ("rows", C.c_int64), ("dim", C.c_int32),
                ("_pad", C.c_int32), ("dense_out", _fp)]



# name -> (restype, argtypes); must list every DMT_API symbol of include/dmt_b200.h
PROTOTYPES = {
    "dmt_abi_version": (C.c_int, []),
    "dmt_last_error": (C.c_char_p, []),
    "dmt_device_sm_count": (C.c_int, []),
    "dmt_embed_gather": (C.c_int, [_fp, C.c_int64, C.c_int32, _fp, C.c_int64, C.c_int32, _fp, _fp]),
    "dmt_seq_encode_workspace_bytes": (C.c_size_t, [C.POINTER(SeqCfg), C.c_int64]),
    "dmt_seq_prepare_weights": (C.c_int, [C.POINTER(SeqCfg), C.POINTER(SeqWeights), _fp, C.c_size_t, _fp]),
    "dmt_seq_encode_fwd": (C.c_int, [C.POINTER(SeqCfg), C.POINTER(SeqInput), C.POINTER(SeqWeights), _fp,
                                     C.c_int64, _fp, C.c_size_t, _fp]),
    "dmt_seq_tail_fwd": (C.c_int, [C.c_int32, _fp, _fp, _fp, _fp, _fp, _fp, _fp]),
    "dmt_seq_encode_multi_fwd": (C.c_int, [C.c_int32, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp]),
    "dmt_pool_mean_fwd": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(PoolFeat), _fp, C.c_int64, _fp]),
    "dmt_pool_mean_fwd_bf16": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(PoolFeat), _fp, C.c_int64, _fp]),
    "dmt_copy_dense_features": (C.c_int, [_fp, C.c_int32, C.c_int32, _fp, C.c_int64, _fp]),
    "dmt_stage_dense_features_bf16": (C.c_int, [_fp, C.c_int32, C.c_int32, C.c_int32, _fp, C.c_int64, _fp]),
    "dmt_mmoe_workspace_bytes": (C.c_size_t, [C.POINTER(MmoeCfg)]),
    "dmt_mmoe_prepared_bytes": (C.c_size_t, [C.POINTER(MmoeCfg)]),
    "dmt_mmoe_prepare_weights": (C.c_int, [C.POINTER(MmoeCfg), C.POINTER(MmoeWeights), _fp, C.c_size_t, _fp]),
    "dmt_mmoe_fwd": (C.c_int, [C.POINTER(MmoeCfg), C.POINTER(MmoeWeights), _fp, C.c_int64, _fp, _fp, C.c_size_t,
                               _fp, _fp]),
    "dmt_mmoe_fwd_bf16in": (C.c_int, [C.POINTER(MmoeCfg), C.POINTER(MmoeWeights), _fp, C.c_int64, _fp, _fp,
                                      C.c_size_t, _fp, _fp]),
    "dmt_forward_bf16": (C.c_int, [C.POINTER(FwdDesc), C.c_int32, _fp, _fp, C.c_int32, _fp, _fp]),
    "dmt_loss_scratch_bytes": (C.c_size_t, [C.c_int32]),
    "dmt_bias_loss_fwd": (C.c_int, [C.POINTER(BiasLossCfg), C.POINTER(BiasWeights), _fp, C.c_int64, _fp, _fp, _fp,
                                    _fp, _fp, _fp, _fp, _fp]),
    "dmt_seq_saved_bytes": (C.c_size_t, [C.POINTER(SeqCfg), C.c_int64]),
    "dmt_seq_encode_fwd_train": (C.c_int, [C.POINTER(SeqCfg), C.POINTER(SeqInput), C.POINTER(SeqWeights), _fp,
                                           C.c_int64, C.c_int64, _fp, C.c_size_t, _fp]),
    "dmt_seq_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(SeqCfg), C.c_int64]),
    "dmt_seq_encode_bwd": (C.c_int, [C.POINTER(SeqCfg), C.POINTER(SeqInput), C.POINTER(SeqWeights), C.c_int64, _fp,
                                     C.c_size_t, _fp, C.c_int64, C.POINTER(SeqWeights), _fp, _fp, _fp, C.c_size_t,
                                     _fp]),
    "dmt_mmoe_train_workspace_bytes": (C.c_size_t, [C.POINTER(MmoeCfg)]),
    "dmt_mmoe_fwd_train": (C.c_int, [C.POINTER(MmoeCfg), C.POINTER(MmoeWeights), _fp, C.c_int64, _fp, _fp, C.c_size_t,
                                     _fp]),
    "dmt_mmoe_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(MmoeCfg)]),
    "dmt_mmoe_bwd": (C.c_int, [C.POINTER(MmoeCfg), C.POINTER(MmoeWeights), _fp, C.c_int64, _fp, _fp,
                               C.POINTER(MmoeWeights), _fp, C.c_int64, C.c_int32, _fp, C.c_size_t, _fp]),
    "dmt_bias_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(BiasLossCfg)]),
    "dmt_bias_bwd": (C.c_int, [C.POINTER(BiasLossCfg), C.POINTER(BiasWeights), _fp, C.c_int64, _fp,
                               C.POINTER(BiasWeights), _fp, C.c_int64, _fp, C.c_size_t, _fp]),
    "dmt_adam_dense": (C.c_int, [C.POINTER(AdamCfg), _fp, _fp, _fp, _fp, C.c_int64, C.c_float, _fp]),
    "dmt_embed_grad_expand": (C.c_int, [C.c_int32, C.POINTER(GradSource), C.c_int64, _fp, _fp, _fp, _fp]),
    "dmt_embed_sorted_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "dmt_embed_adam_sorted": (C.c_int, [C.POINTER(AdamCfg), _fp, _fp, _fp, C.c_int64, C.c_int32, C.c_int32,
                                        C.POINTER(GradSource), _fp, _fp, _fp, _fp, C.c_int64, C.c_float, _fp, _fp,
                                        C.c_size_t, _fp]),
    "dmt_embed_grad_densify_sorted": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.POINTER(GradSource), _fp, _fp, _fp,
                                                _fp, C.c_int64, C.c_float, _fp, _fp, C.c_size_t, _fp]),
    "dmt_embed_grad_scatter_rows": (C.c_int, [C.c_int32, C.POINTER(GradSource), _fp, _fp, _fp, C.c_int64, C.c_int32,
                                              _fp, _fp]),
    "dmt_adam_rows_untouched": (C.c_int, [C.POINTER(AdamCfg), _fp, _fp, _fp, C.c_int64, C.c_int32, _fp, _fp]),
    "dmt_embed_grad_expand_multi": (C.c_int, [C.c_int32, C.POINTER(AdamTable), C.c_int32, C.POINTER(GradSource),
                                              C.POINTER(C.c_int32), _fp, _fp, _fp, _fp]),
    "dmt_embed_sorted_multi_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "dmt_embed_adam_sorted_multi": (C.c_int, [C.POINTER(AdamCfg), C.c_int32, C.POINTER(AdamTable), C.c_int32,
                                              C.POINTER(GradSource), _fp, _fp, _fp, _fp, C.c_int64, C.c_float, _fp,
                                              C.c_size_t, _fp]),
    "dmt_adam_rows_untouched_multi": (C.c_int, [C.POINTER(AdamCfg), C.c_int32, C.POINTER(AdamTable), _fp]),
    "dmt_build_digest": (C.c_char_p, []),
    "dmt_widen_u16": (C.c_int, [C.c_int32, C.POINTER(WidenDesc), _fp]),
    "dmt_widen_ids": (C.c_int, [C.c_int32, C.POINTER(WidenIdsDesc), _fp]),
    "dmt_copy_dense_features_bf16": (C.c_int, [_fp, C.c_int32, C.c_int32, _fp, C.c_int64, _fp]),
    "dmt_debug_seq_profile": (C.c_int, [_fp]),
    "dmt_debug_seq_timer": (C.c_int, [C.c_int32]),
    "dmt_debug_seq_timer_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "dmt_selftest_umma": (C.c_int, [C.c_int32, _fp, _fp, _fp, C.c_int32, C.c_int32, _fp]),
    "dmt_selftest_tf32_rows": (C.c_int, [_fp, C.c_int64, _fp, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _fp,
                                         C.c_int64, _fp, _fp, C.c_int64, _fp, C.c_int64, C.c_float, C.c_int32,
                                         C.c_int32, _fp]),
    "dmt_selftest_tf32_gemm": (C.c_int, [_fp, C.c_int64, C.c_int32, _fp, C.c_int64, C.c_int32, C.c_int64, C.c_int32,
                                         C.c_int32, _fp, C.c_int64, _fp, _fp, C.c_int64, C.c_int32, C.c_int32, _fp]),
    "dmt_selftest_tf32_wgrad_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "dmt_selftest_tf32_wgrad": (C.c_int, [_fp, C.c_int64, _fp, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _fp,
                                          C.c_int64, C.c_int32, C.c_int32, _fp, _fp, _fp]),
    "dmt_selftest_tf32_colsum": (C.c_int, [_fp, C.c_int64, C.c_int64, C.c_int32, _fp, C.c_int32, _fp, _fp]),
}


class DmtError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("dmt_b200 error %d: %s" % (code, message))
        self.code = code


_LIB = None


def lib_path():
    return _build.LIB


def load(rebuild=True):
    """Load the shared library, (re)building it in-tree when nvcc is available and sources
    changed.  Raises if it cannot be produced -- no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.build() if rebuild else _build.LIB
    if not os.path.exists(path):
        raise RuntimeError("libdmt_b200.so is missing (%s); run `python -m cikm2020_dmt_b200.build`" % path)
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    got = lib.dmt_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError("libdmt_b200.so ABI %d != binding ABI %d" % (got, ABI_VERSION))
    built = lib.dmt_build_digest()
    built = built.decode() if built else ""
    want = _build.source_digest()
    if built != want and os.environ.get("DMT_ALLOW_STALE_LIB") != "1":
        # the ctypes structures above mirror the header the library was compiled from: a stale library means
        # mismatched layouts, i.e. memory corruption instead of an error
        raise RuntimeError("libdmt_b200.so was built from other sources (digest %s..., tree %s...) and nvcc is not "
                           "available to rebuild it; run `python -m cikm2020_dmt_b200.build` where nvcc exists"
                           % (built[:12], want[:12]))
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().dmt_last_error()
        raise DmtError(rc, msg.decode() if msg else "")


def ptr(t):
    """Device (or host) address of a torch tensor, None -> NULL."""
    return None if t is None else t.data_ptr()


def dense(w, b):
    return Dense(ptr(w), ptr(b))
