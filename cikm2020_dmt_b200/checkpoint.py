"""Checkpoint save / resume and the serving score.

* `save` / `latest` / `load`: the reference saves `tf.trainable_variables()` every `validate_step` with
  `tf.train.Saver(max_to_keep=0)` as `model.ckpt-<step>` and drops a `step-<step>.model.DONE` marker next to it
  (run_dnn.py:258-261,362-388); resuming restores those variables only (:296-306 -- the Adam slots restart from
  zero).  Here a checkpoint is `model.ckpt-<step>.npz` whose keys are the TF variable names of the reference graph
  (params.py), so a tensor exported from a TF-1 checkpoint by name drops in.  TF's own bundle format cannot be
  written without TensorFlow -- the container is the one deliberate difference.  `optimizer=` additionally stores
  the Adam slots under TF's slot names (`<var>/Adam`, `<var>/Adam_1`) and the step for an exact resume.
  A row-sharded table (train.py) is written per rank as `<name>@rows<lo>-<hi>`.
* `serving_features` / `serving_scores`: saved_model/export_model.py:88-115 + preprocess.py:17-43 -- the dense
  features are z-score-style normalised with the training mean/std and clipped to +-0.99, the model runs with
  `is_predict`, and `Scores = (w0 sigmoid(click) + w1 sigmoid(order)) / (w0 + w1)` with `[export_model]
  export_weight`.
"""
import glob
import os
import re
from typing import Dict, Optional

import numpy as np
import torch


def _ckpt_path(directory, step):
    return os.path.join(directory, "model.ckpt-%d.npz" % step)


def save(directory: str, step: int, store, optimizer=None, shards: Optional[Dict[str, tuple]] = None,
         rank: int = 0, extra: Optional[Dict[str, float]] = None) -> str:
    """Write the trainable variables (TF names) of `store`; returns the file path.  With `shards`
    ({variable name: (lo, hi)}) the named tables are row shards of this rank and the file gets a `.rank<r>` suffix
    for ranks > 0 (rank 0 also holds the replicated variables)."""
    os.makedirs(directory, exist_ok=True)
    arrays = {}
    shards = shards or {}
    for name, t in store.named_parameters():
        if name in shards:
            lo, hi = shards[name]
            arrays["%s@rows%d-%d" % (name, lo, hi)] = t.detach().cpu().numpy()
        elif rank == 0:
            arrays[name] = t.detach().cpu().numpy()
    slot_m, slot_v = getattr(optimizer, "SLOT_NAMES", ("Adam", "Adam_1"))   # TF slot names of the two state buffers
    if optimizer is not None:
        arrays["global_step"] = np.asarray(optimizer.t, dtype=np.int64)
        if hasattr(optimizer, "global_step"):      # the learning-rate schedule's clock (run_dnn.py:122-126)
            arrays["lr_global_step"] = np.asarray(optimizer.global_step, dtype=np.int64)
        if rank == 0:
            for spec in store.specs:
                sl = slice(spec.offset, spec.offset + spec.numel)
                if slot_m:
                    arrays[spec.name + "/" + slot_m] = optimizer.m_dense[sl].detach().cpu().numpy().reshape(spec.shape)
                if slot_v:
                    arrays[spec.name + "/" + slot_v] = optimizer.v_dense[sl].detach().cpu().numpy().reshape(spec.shape)
        for name in store.tables:
            if name in shards or rank == 0:
                sfx = "@rows%d-%d" % shards[name] if name in shards else ""
                if slot_m:
                    arrays[name + "/" + slot_m + sfx] = optimizer.m_tab[name].detach().cpu().numpy()
                if slot_v:
                    arrays[name + "/" + slot_v + sfx] = optimizer.v_tab[name].detach().cpu().numpy()
    for k, v in (extra or {}).items():
        arrays["extra/" + k] = np.asarray(v)
    path = _ckpt_path(directory, step) if rank == 0 else _ckpt_path(directory, step)[:-4] + ".rank%d.npz" % rank
    tmp = path + ".tmp.npz"
    np.savez(tmp, **arrays)
    os.replace(tmp, path)
    if rank == 0:
        with open(os.path.join(directory, "step-%d.model.DONE" % step), "w"):
            pass                                   # create_file(MODEL_PATH, 'step-%d.model.DONE'), run_dnn.py:385
    return path


def latest(directory: str) -> Optional[int]:
    """Highest step whose checkpoint is complete (has its .DONE marker)."""
    steps = []
    for p in glob.glob(os.path.join(directory, "step-*.model.DONE")):
        m = re.search(r"step-(\d+)\.model\.DONE$", p)
        if m and os.path.exists(_ckpt_path(directory, int(m.group(1)))):
            steps.append(int(m.group(1)))
    return max(steps) if steps else None


def load(directory: str, step: int, store, optimizer=None, shards: Optional[Dict[str, tuple]] = None,
         rank: int = 0, strict: bool = True, model=None) -> Dict[str, np.ndarray]:
    """Restore `store` (and the Adam slots when `optimizer` is given and they were saved).  A sharded table is
    cut out of a full table or assembled from the per-rank shard files, whichever the checkpoint holds; its Adam
    slots (`<var>/Adam`, `<var>/Adam_1`) are sharded by the SAME row range as the variable.
    `model` (or `optimizer.model`): its cached bf16 weight images are invalidated, so the next forward re-derives
    them from the restored parameters."""
    files = [_ckpt_path(directory, step)] + sorted(glob.glob(_ckpt_path(directory, step)[:-4] + ".rank*.npz"))
    data = {}
    for f in files:
        with np.load(f) as z:
            for k in z.files:
                data[k] = z[k]
    shards = shards or {}

    def fetch(name, shard_of=None):
        """`name`: array to restore; `shard_of`: the variable whose row range cuts it (the variable itself, or
        the variable an Adam slot belongs to)."""
        key = name if shard_of is None else shard_of
        pieces = [k for k in data if k.startswith(name + "@rows")]
        if key in shards or pieces:
            lo, hi = shards.get(key, (0, None))
            if name in data:
                return data[name][lo:hi]
            if not pieces:
                return None
            parts = sorted((int(re.search(r"@rows(\d+)-", k).group(1)), data[k]) for k in pieces)
            full = np.concatenate([p for _, p in parts], 0)
            return full[lo:hi]
        return data.get(name)

    missing = []
    for name, t in store.named_parameters():
        arr = fetch(name)
        if arr is None:
            missing.append(name)
            continue
        t.copy_(torch.from_numpy(np.ascontiguousarray(arr)).to(t.device).reshape(t.shape))
    if strict and missing:
        raise KeyError("checkpoint %s lacks %d variables, e.g. %s" % (files[0], len(missing), missing[0]))
    slot_m, slot_v = getattr(optimizer, "SLOT_NAMES", ("Adam", "Adam_1"))
    if optimizer is not None and "global_step" in data:
        optimizer.t = int(data["global_step"])
        if hasattr(optimizer, "global_step"):
            optimizer.global_step = int(data.get("lr_global_step", data["global_step"]))
        for spec in store.specs:
            sl = slice(spec.offset, spec.offset + spec.numel)
            if slot_m and spec.name + "/" + slot_m in data:
                optimizer.m_dense[sl].copy_(torch.from_numpy(data[spec.name + "/" + slot_m]).reshape(-1))
            if slot_v and spec.name + "/" + slot_v in data:
                optimizer.v_dense[sl].copy_(torch.from_numpy(data[spec.name + "/" + slot_v]).reshape(-1))
        for name in store.tables:
            m = fetch(name + "/" + slot_m, shard_of=name) if slot_m else None
            v = fetch(name + "/" + slot_v, shard_of=name) if slot_v else None
            if m is not None:
                optimizer.m_tab[name].copy_(torch.from_numpy(np.ascontiguousarray(m)))
            if v is not None:
                optimizer.v_tab[name].copy_(torch.from_numpy(np.ascontiguousarray(v)))
    if model is None and optimizer is not None:
        model = getattr(optimizer, "model", None)
    if model is not None and hasattr(model, "invalidate_prepared"):
        model.invalidate_prepared()
    return {k[6:]: v for k, v in data.items() if k.startswith("extra/")}


# ----------------------------------------------------------------------------- serving (export_model.py)
def read_const_vector(path: str) -> np.ndarray:
    """util.get_const_data: one tab-separated line of floats (jd_recsys_demo/stat/{mean,std}/part-00000)."""
    with open(path) as fh:
        return np.asarray([float(x) for x in fh.read().split()], dtype=np.float64)


def serving_features(features: torch.Tensor, mean: np.ndarray, std: np.ndarray) -> torch.Tensor:
    """export_model.py:88-96 with preprocess.vec_constant (:17-43):
    c = mean*std/(3 (std+eps)^2) + mean*std/(std+eps) - mean;  x' = clip(clip(x, 0, max) * std / (3 (std+eps)^2) - c, +-0.99)."""
    eps = 1e-7
    mean64, std64 = torch.from_numpy(mean), torch.from_numpy(std)
    c = (mean64 * std64 / ((std64 + eps) ** 2 * 3) + mean64 * std64 / (std64 + eps) - mean64).to(torch.float32)
    std32 = std64.to(torch.float32)
    x = features.to(torch.float32).clamp(min=0.0)
    out = x * std32.to(x.device) / (((std32 + torch.tensor(eps, dtype=torch.float32)) ** 2) * 3.0).to(x.device) - c.to(x.device)
    return out.clamp(-0.99, 0.99)


def serving_scores(model, inputs, export_weight, mean=None, std=None):
    """The exported graph's `Scores` (export_model.py:98-115): is_predict path, weighted mean of the sigmoids."""
    if mean is not None:
        inputs = dict(inputs)
        inputs["features"] = serving_features(inputs["features"], mean, std)
    click, order = model.inference(inputs, is_train=False, is_predict=True)
    w0, w1 = float(export_weight[0]), float(export_weight[1])
    return (w0 * torch.sigmoid(click.reshape(-1)) + w1 * torch.sigmoid(order.reshape(-1))) / (w0 + w1)
