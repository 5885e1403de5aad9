"""The C-ABI shared library: builds for sm_100a, loads without a GPU, exports every symbol the header
declares, and its POD structs have the layout the ctypes binding assumes.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "dmt_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"DMT_API\s+[\w\s\*]+?\b(dmt_\w+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from cikm2020_dmt_b200 import abi
    lib = abi.load()
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(abi.PROTOTYPES) == names, "binding table and header disagree"
    assert lib.dmt_abi_version() == abi.ABI_VERSION
    out = subprocess.run(["nm", "-D", "--defined-only", abi.lib_path()], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported
    assert not any(s.startswith("_ZN3dmt") for s in exported), "internal symbols leak from the ABI"


def test_library_contains_sm100a_code():
    from cikm2020_dmt_b200 import abi
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", abi.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layouts_match_header(tmp_path):
    from cikm2020_dmt_b200 import abi
    structs = {"dmt_adam_cfg": abi.AdamCfg, "dmt_grad_source": abi.GradSource, "dmt_seq_cfg": abi.SeqCfg, "dmt_seq_input": abi.SeqInput, "dmt_seq_weights": abi.SeqWeights,
               "dmt_pool_feat": abi.PoolFeat, "dmt_mmoe_cfg": abi.MmoeCfg, "dmt_mmoe_weights": abi.MmoeWeights,
               "dmt_bias_loss_cfg": abi.BiasLossCfg, "dmt_bias_weights": abi.BiasWeights,
               "dmt_dense": abi.Dense, "dmt_attn_weights": abi.AttnWeights, "dmt_ff_weights": abi.FFWeights,
               "dmt_fwd_feature": abi.FwdFeature, "dmt_fwd_desc": abi.FwdDesc, "dmt_widen_ids_desc": abi.WidenIdsDesc}
    probes = {"dmt_seq_cfg": ["precision", "n_feats", "flags", "dropout_seed"], "dmt_seq_input": ["ids", "item_ids", "dim"],
              "dmt_seq_weights": ["dec_attn", "ff"], "dmt_pool_feat": ["weights", "out_col"],
              "dmt_mmoe_cfg": ["n_tasks", "tower_units", "precision"], "dmt_mmoe_weights": ["gate", "tower_out"],
              "dmt_bias_loss_cfg": ["ctr_rel", "weight_ecvr", "loss_weight"],
              "dmt_fwd_feature": ["weights"], "dmt_widen_ids_desc": ["n", "bytes"],
              "dmt_fwd_desc": ["pool", "seq_in", "seq_ws_bytes", "mmoe_ws_bytes", "bias_ld", "xb_ld", "inputs_ready"]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "%s"' % HEADER, "int main(void){"]
    for name in structs:
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f in probes.get(name, []):
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    lines += ["return 0;}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    for line in out.splitlines():
        key, val = line.split()
        if "." in key:
            s, f = key.split(".")
            assert getattr(structs[s], f).offset == int(val), key
        else:
            assert C.sizeof(structs[key]) == int(val), key


def test_header_is_plain_c_and_cites_the_reference():
    text = open(HEADER).read()
    assert 'extern "C"' in text and "torch" not in text.replace("no torch types", "")
    for cite in ("base.py:81-91", "mmoe_transformer_unbias.py:130-186", "TransformerModel.py:51-171",
                 "inference_mlp.py:162-223", "mmoe_transformer_unbias.py:63-126"):
        assert cite in text, cite


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cikm2020_dmt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_build_digest_does_not_depend_on_the_checkout_path(tmp_path, monkeypatch):
    """The source digest decides whether a box rebuilds the library: it must be the same for a copy of the tree
    under another path (gpurun boxes, CI checkouts), or every rank of a torchrun launch would recompile."""
    import shutil
    from cikm2020_dmt_b200 import build as B
    want = B._digest()
    pkg = tmp_path / "elsewhere" / "cikm2020_dmt_b200"
    (pkg / "csrc").mkdir(parents=True)
    (tmp_path / "elsewhere" / "include").mkdir()
    for f in os.listdir(B.CSRC):
        if f.endswith((".cu", ".cuh", ".h")):
            shutil.copy(os.path.join(B.CSRC, f), pkg / "csrc" / f)
    shutil.copy(HEADER, tmp_path / "elsewhere" / "include" / "dmt_b200.h")
    monkeypatch.setattr(B, "HERE", str(pkg))
    monkeypatch.setattr(B, "CSRC", str(pkg / "csrc"))
    assert B._digest() == want


def test_argument_errors_need_no_gpu():
    """Argument validation happens before any CUDA call: error codes and messages without a device."""
    import ctypes as C
    from cikm2020_dmt_b200 import abi
    lib = abi.load()
    rc = lib.dmt_seq_tail_fwd(abi.MAX_TAIL_SEQS + 1, None, None, None, None, None, None, None)
    assert rc < 0 and b"n_seq" in lib.dmt_last_error()
    assert lib.dmt_seq_tail_fwd(0, None, None, None, None, None, None, None) == 0
    rc = lib.dmt_seq_tail_fwd(1, None, None, None, None, None, None, None)
    assert rc < 0 and b"null" in lib.dmt_last_error()
    cfg = abi.SeqCfg(4, 64, 256, 2, 1, 1, 50, 1, 5, abi.PRECISION_BF16, 0, 0, 0.0, 0)
    assert lib.dmt_seq_encode_workspace_bytes(C.byref(cfg), 0) > 128 * 1024       # weight images + context image
    big = abi.SeqCfg(4096, 64, 256, 2, 1, 1, 50, 1, 5, abi.PRECISION_BF16, 0, 0, 0.0, 0)
    grow = lib.dmt_seq_encode_workspace_bytes(C.byref(big), 0) - lib.dmt_seq_encode_workspace_bytes(C.byref(cfg), 0)
    sched = lambda b: (b * 4 + 16 + 255) // 256 * 256                            # length-class schedule: perm + counts
    assert grow == (4096 // 128 - 1) * 128 * 128 * 2 + sched(4096) - sched(4)    # one 32 KB image per 128 samples
    assert lib.dmt_forward_bf16(None, 0, None, None, 0, None, None) == -1 and b"null" in lib.dmt_last_error()
    assert lib.dmt_stage_dense_features_bf16(None, 0, 4, 8, None, 8, None) == -1
    assert lib.dmt_pool_mean_fwd_bf16(4, 1, None, None, 8, None) == -1
    rc = lib.dmt_seq_encode_multi_fwd(abi.MAX_TAIL_SEQS + 1, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"n_seq" in lib.dmt_last_error()
    assert lib.dmt_seq_encode_multi_fwd(0, None, None, None, None, None, None, None, None) == 0
    rc = lib.dmt_seq_encode_multi_fwd(1, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"null" in lib.dmt_last_error()
