#!/usr/bin/env python
"""bench.py -- the DMT hot path on N B200s of one node, one JSON line on stdout (rank 0).

Metric (BASELINE.json): samples/s of the forward ranking path on BASELINE config 2 -- "DMT full
3-seq fwd-only, d_model=64, 2 heads, 1 block, batch=4096, synthetic ids" -- per GPU, weak scaling
(every rank runs its own 4096-sample batches; the path is data-parallel with no data-path
collective in forward-only mode).

  value      whole-job samples/s with the batch already resident in HBM, CUDA-event timed
  e2e        the same metric through the plugin call (`Inference.inference`) from PINNED HOST
             buffers: one packed H2D copy per step + the D2H read of the logits, inside the timing
  roofline   the dominant kernel's algorithmic bytes (or FLOPs) / its live CUDA-event time, against
             MEASURED_PEAKS.json
  cpu_baseline   the CPU oracle (`oracle/dmt_oracle.py`, PyTorch CPU fp32, all host threads) on a
             bounded sample of the same workload -- a reported baseline, not the target

`--impl reference` times the reference's CPU path instead (the oracle port; TF-1.12 cannot be
installed here) and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0        # /opt/skills/guides/B200_PROFILING.md fallback
FALLBACK_BF16_TFLOPS = 1590.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="per-GPU batch (BASELINE config 2: 4096)")
    ap.add_argument("--conf", default="dmt_d64.conf")
    ap.add_argument("--id-mode", default="uniform", choices=["uniform", "zipf"])
    ap.add_argument("--precision", default="bf16", choices=["f32", "bf16", "tf32"],
                    help="headline path: bf16 = fused tcgen05 tile kernels (d_model 64 / 2 heads); tf32 = row-batched "
                         "tcgen05 pipeline (any shape, e.g. --conf dmt.conf); f32 = CUDA cores")
    ap.add_argument("--n-batches", type=int, default=4, help="distinct batches rotated through the timed loop")
    ap.add_argument("--cpu-batch", type=int, default=256, help="samples per CPU-baseline step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline time budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--small-tables", action="store_true", help="debug: tiny vocabularies")
    ap.add_argument("--train-batch", type=int, default=8192, help="per-GPU training batch (BASELINE configs 3/4: 8192)")
    ap.add_argument("--train-steps", type=int, default=10, help="timed training steps of the `train` block")
    ap.add_argument("--train-gemm", default="tf32", choices=["f32", "bf16", "bf16x3", "tf32"],
                    help="GEMM engine of the training step (BASELINE config 3 names bf16; tf32 keeps more operand bits)")
    ap.add_argument("--no-train", action="store_true", help="skip the `train` block (configs 3/4)")
    ap.add_argument("--no-f32", action="store_true", help="skip the `f32` forward block")
    ap.add_argument("--wide-batch", action="store_true",
                    help="e2e: ship int32 ids + fp32 features (the round-1 packed format) instead of the compact one")
    ap.add_argument("--f32-steps", type=int, default=10)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured"}
    return {"hbm_gbs": FALLBACK_HBM_GBS, "bf16_tflops": FALLBACK_BF16_TFLOPS, "source": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(args, device, rank):
    import torch
    from cikm2020_dmt_b200.conf import Conf
    from cikm2020_dmt_b200.data import synthetic_batch, batch_tokens, SEED
    from cikm2020_dmt_b200.plan import build_plan
    conf = Conf(os.path.join(ROOT, "conf", "settings") + "/", args.conf)
    plan = build_plan(conf)
    rows = None
    if args.small_tables:
        rows = {"Sku": 20000, "Brand": 2000, "Shopid": 2000, "Cid3": 1000, "Cid2": 100}
        for t in list(plan.tables.values()) + list(plan.bias_tables.values()):
            if t.name in rows:
                t.rows = rows[t.name]
    batches = [synthetic_batch(plan, args.batch, seed=SEED + 1000 * rank + i, id_mode=args.id_mode, table_rows=rows)
               for i in range(args.n_batches)]
    return conf, plan, batches, rows


def algorithmic_bytes_seq(plan, batch):
    """SURVEY 8(d), fused K1-K4 boundary: per (sample, sequence) L*(sum D_f*e + 20) + (sum D_f*e + 20)
    + 4 + d*a_out, fp32 tables (e=4) and fp32 interest vector (a_out=4)."""
    total = 0
    B = batch["features"].shape[0]
    for seq in plan.sequences:
        n_tok = int(batch[seq.user_features[-1]].offsets[-1])
        row_bytes = sum(seq.dims) * 4 + 4 * len(seq.dims)
        total += n_tok * row_bytes + B * (row_bytes + 4 + plan.d_model * 4)
    return total


def flops_seq(plan, batch, tail=True):
    """Algorithmic FLOPs (valid tokens only) of A2-A8; tail=False leaves out the per-sample decoder tail
    (ctx Wv + FF on B rows: seq_tail_kernel's share)."""
    d, dff = plan.d_model, plan.d_ff
    total = 0
    B = batch["features"].shape[0]
    for seq in plan.sequences:
        lens = (batch[seq.user_features[-1]].offsets[1:] - batch[seq.user_features[-1]].offsets[:-1]).double()
        n_tok = float(lens.sum())
        total += n_tok * (2 * d * 3 * d + 4 * d * dff) * plan.num_blocks_encode          # QKV + FF per token
        total += float((lens * lens).sum()) * 4 * d * plan.num_blocks_encode              # QK^T + PV
        total += (n_tok * (2 * d * 2 * d + 4 * d) + (B * (2 * d * d + 4 * d * dff) if tail else 0)) * plan.num_blocks_decode
    return total


def flops_mmoe(plan, B):
    f, k = 0, plan.mmoe_in
    for u in plan.hidden_units_bottom:
        f += 2 * k * u
        k = u
    f *= plan.num_experts
    f += 2 * plan.mmoe_in * plan.num_experts * plan.num_tasks
    return f * B


def cpu_oracle_throughput(plan, params_cpu, batch, seconds, lean=True):
    """samples/s of the CPU oracle (PyTorch CPU fp32, all host threads) on `batch`."""
    import torch
    from oracle import dmt_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    B = batch["features"].shape[0]
    with torch.no_grad():
        O.inference(plan, params_cpu, batch, is_train=False, lean=lean)     # warm-up
        times = []
        t_end = time.perf_counter() + seconds
        while len(times) < 3 or (time.perf_counter() < t_end and len(times) < 200):
            t0 = time.perf_counter()
            O.inference(plan, params_cpu, batch, is_train=False, lean=lean)
            times.append(time.perf_counter() - t0)
    return B / statistics.median(times), len(times)


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path = the oracle port."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cikm2020_dmt_b200.data import synthetic_batch, SEED
    from cikm2020_dmt_b200.params import ParamStore
    from oracle import dmt_oracle as O
    conf, plan, _, rows = build_workload(argparse.Namespace(**{**vars(args), "n_batches": 0}), "cpu", 0)
    torch.set_num_threads(os.cpu_count() or 1)
    store = ParamStore(plan, device="cpu")
    P = O.params_from_store(store, torch.float32)
    sample = synthetic_batch(plan, args.cpu_batch, seed=SEED, id_mode=args.id_mode, table_rows=rows)
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            O.inference(plan, P, sample, is_train=False, lean=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.inference(plan, P, sample, is_train=False, lean=True)
        dt = time.perf_counter() - t0
    value = args.cpu_batch * args.steps / dt
    desc = ("oracle port (PyTorch CPU fp32, lean lookups) fwd-only on %d-sample slices of the %d-sample batch"
            % (args.cpu_batch, args.batch))
    line = {
        "impl": "reference", "metric": "samples/sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args, plan), precision="f32", cpu_batch=args.cpu_batch,
                       sample="each step = one forward pass over a %d-sample slice of the %d-sample batch"
                              % (args.cpu_batch, args.batch),
                       l2="n/a (CPU arm)"),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": desc},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, plan):
    return {"workload": "BASELINE config 2: DMT full 3-seq (clk50/ord50/cart10) fwd-only, d_model=%d, %d heads, "
                        "%d+%d blocks, 615 dense + 23 pooled id features, MMoE %d experts / 2 tasks + bias tower, "
                        "per-GPU batch %d, synthetic %s ids, Sku vocabulary %d"
                        % (plan.d_model, plan.num_heads, plan.num_blocks_encode, plan.num_blocks_decode,
                           plan.num_experts, args.batch, args.id_mode, plan.tables["Sku"].rows),
            "conf": args.conf, "per_gpu_batch": args.batch, "precision": args.precision,
            "l2": "inputs larger than L2: random rows of a %.0f MB Sku table + %d rotating batches"
                  % (plan.tables["Sku"].rows * plan.tables["Sku"].dim * 4 / 1e6, args.n_batches)}


NO_DROPOUT = {("model", "transformer_dropout_rate"): "0.0", ("model", "dropout_rate_bias"): "0.0,0.0"}
SMALL_ROWS = {"Sku": 20000, "Brand": 2000, "Shopid": 2000, "Cid3": 1000, "Cid2": 100}


def make_timer(world, device):
    """timed(step_fn, steps) -> ms: CUDA events on the current stream between barriers, MAX over ranks."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step_fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    return timed


def slice_batch(batch, lo, hi):
    """Samples [lo, hi) of a host batch (CSR features re-based)."""
    from cikm2020_dmt_b200.data import SparseIds
    sub = {}
    for k, v in batch.items():
        if isinstance(v, SparseIds):
            a, b = int(v.offsets[lo]), int(v.offsets[hi])
            sub[k] = SparseIds(v.values[a:b].clone(), (v.offsets[lo:hi + 1] - a).clone(),
                               None if v.weights is None else v.weights[a:b].clone())
        elif hasattr(v, "shape"):
            sub[k] = v[lo:hi].clone()
    return sub


def run_f32_forward(args, plan, store, dev_batches, timed, world):
    """The `f32` block: the same forward workload in the reference's own arithmetic (fp32 end to end, CUDA-core
    kernels; logits within 2e-4 of the fp64 oracle -- tests/test_gpu_parity.py).  The headline `value` is the bf16
    tensor-core path (logits atol 5e-2); this is the number to hold against the fp32 reference arm."""
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    model = mmoe_transformer_unbias(plan, params=store, precision="f32")
    step = lambda i: model.inference(dev_batches[i % len(dev_batches)], is_train=False)
    for i in range(3):
        step(i)
    l0 = model.launches
    ms = timed(step, args.f32_steps)
    B = args.batch
    return {"value": world * B * args.f32_steps / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms / args.f32_steps,
            "steps": args.f32_steps, "warmup": 3, "dtype": "f32", "gpu_launches": int(model.launches - l0),
            "kernels": "seq_encode_f32_kernel + fp32 SIMT MMoE (CUDA cores)",
            "tolerance": "logits atol 2e-4 vs the fp64 oracle (tests/test_gpu_parity.py)",
            "note": "same workload and batches as `value`, inputs resident in HBM"}


def run_tf32_forward(args, conf_file, rank, device, timed, world):
    """A `tf32` block: the forward workload through precision='tf32' -- the row-batched pipeline whose dense
    projections, feed-forward and MMoE experts run on the TMA-fed tcgen05 kind::tf32 engine straight from fp32
    activations (no bf16 rounding of activations or weights; attention / LayerNorm fp32 on CUDA cores).  Runs any
    d_model / d_ff that are multiples of 16: `conf_file` = the bench conf (d_model 64) or the reference's own
    dmt.conf (d_model 80, 4 heads, d_ff 320).  Logits within atol 2e-2 + rtol 1e-2 of the fp64 oracle (tests)."""
    import torch
    from cikm2020_dmt_b200.conf import Conf
    from cikm2020_dmt_b200.data import synthetic_batch, batch_to, SEED
    from cikm2020_dmt_b200.params import ParamStore
    from cikm2020_dmt_b200.plan import build_plan
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    conf = Conf(os.path.join(ROOT, "conf", "settings") + "/", conf_file)
    plan = build_plan(conf)
    rows = None
    if args.small_tables:
        rows = SMALL_ROWS
        for t in list(plan.tables.values()) + list(plan.bias_tables.values()):
            if t.name in rows:
                t.rows = rows[t.name]
    store = ParamStore(plan, device=device)
    model = mmoe_transformer_unbias(plan, params=store, precision="tf32")
    batches = [batch_to(synthetic_batch(plan, args.batch, seed=SEED + 70000 + 1000 * rank + i, id_mode=args.id_mode,
                                        table_rows=rows), device) for i in range(3)]
    step = lambda i: model.inference(batches[i % len(batches)], is_train=False)
    for i in range(3):
        step(i)
    l0 = model.launches
    ms = timed(step, args.f32_steps)
    out = {"conf": conf_file, "d_model": plan.d_model, "num_heads": plan.num_heads, "d_ff": plan.d_ff,
           "mmoe_in": plan.mmoe_in, "value": world * args.batch * args.f32_steps / (ms / 1e3), "unit": "samples/s",
           "ms_per_step": ms / args.f32_steps, "steps": args.f32_steps, "warmup": 3,
           "dtype": "tf32 operands (fp32 storage, 10-bit operand mantissa in the tensor core) / fp32 accumulate",
           "gpu_launches": int(model.launches - l0),
           "kernels": "tf32_rows_kernel / tf32_gemm_kernel (TMA + tcgen05 kind::tf32) + fp32 SIMT attention, LayerNorm",
           "tolerance": "logits atol 2e-2 + rtol 1e-2 vs the fp64 oracle (tests/test_gpu_parity.py)"}
    del model, store, batches
    torch.cuda.empty_cache()
    return out


def run_dp_parity(args, world, rank, device):
    """N ranks on 1/N of a global batch each (row-sharded Sku, one allreduce bucket) vs ONE rank on the whole batch,
    2 optimizer steps, dropout off, small vocabularies: max / mean |delta parameter| over every variable (the
    attention key biases excluded: their exact gradient is 0 and Adam amplifies fp32 noise -- same rule as
    tests/test_gpu_dist_train.py).  With N == 1 the data-parallel code path (compact tables, densified replicas,
    bucket) runs with world 1 against the plain step."""
    import torch
    import torch.distributed as dist
    from cikm2020_dmt_b200.conf import Conf
    from cikm2020_dmt_b200.data import synthetic_batch, batch_to
    from cikm2020_dmt_b200.plan import build_plan
    from cikm2020_dmt_b200.train import Trainer
    conf = Conf(os.path.join(ROOT, "conf", "settings") + "/", args.conf, overrides=NO_DROPOUT)
    plan = build_plan(conf)
    for t in list(plan.tables.values()) + list(plan.bias_tables.values()):
        if t.name in SMALL_ROWS:
            t.rows = SMALL_ROWS[t.name]
    per = 32
    G = per * world
    engine = dict(precision="f32", train_gemm="f32")          # exact engines: the comparison is about the routing
    dp = Trainer(plan, device, seed=3, randomize=4, world=world, rank=rank, force_dp_path=(world == 1),
                 learning_rate=1e-3, **engine)
    ref = Trainer(plan, device, seed=3, randomize=4, learning_rate=1e-3, **engine)
    loss_err = 0.0
    for s in range(2):
        h = synthetic_batch(plan, G, seed=500 + s, table_rows=SMALL_ROWS)
        h["mask"] = torch.nn.functional.one_hot((torch.arange(G) + s) % 5, 5).float()
        dp.train_step(batch_to(slice_batch(h, rank * per, (rank + 1) * per), device))
        lr_ = ref.train_step(batch_to(h, device)).item()
        loss_err = max(loss_err, abs(dp.global_loss().item() - lr_) / max(abs(lr_), 1e-12))
    torch.cuda.synchronize()
    mx, mean_mx = 0.0, 0.0
    for name, v in ref.store.named_parameters():
        if name.endswith("attention/dense_1/bias"):
            continue
        got = dp.store.views[name]
        if name in dp.sharded:
            sh = dp.sharded[name]
            v = v[sh.lo:sh.hi]
        err = (v - got).abs()
        if err.numel():
            mx = max(mx, float(err.max()))
            mean_mx = max(mean_mx, float(err.mean()))
    stats = torch.tensor([mx, mean_mx, loss_err], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    mx, mean_mx, loss_err = [float(x) for x in stats]
    del dp, ref
    torch.cuda.empty_cache()
    return {"max_abs_dparam": mx, "max_mean_abs_dparam": mean_mx, "loss_rel_err": loss_err, "steps": 2,
            "ranks": world, "global_batch": G, "tol": {"max": 2e-4, "mean": 2e-6, "loss_rel": 5e-5},
            "ok": bool(mx <= 2e-4 and mean_mx <= 2e-6 and loss_err <= 5e-5),
            "what": "%d rank(s) on split batches (Sku row-sharded, allreduce bucket) vs 1 rank on the whole batch"
                    % world}


def run_train(args, world, rank, local_rank, device, timed):
    """The `train` block: BASELINE config 3 (N = 1) / config 4 (N > 1: global batch N x 8192, dense + small-table
    gradients in ONE NCCL allreduce, Sku row-sharded with all-to-all row / gradient exchange): forward with saved
    activations + backward + TF-1 Adam (dense semantics), the conf's training-mode dropout, batches resident."""
    import torch
    from cikm2020_dmt_b200.conf import Conf
    from cikm2020_dmt_b200.data import synthetic_batch, batch_to, batch_tokens, SEED
    from cikm2020_dmt_b200.plan import build_plan
    from cikm2020_dmt_b200.train import Trainer
    conf = Conf(os.path.join(ROOT, "conf", "settings") + "/", args.conf)
    plan = build_plan(conf)
    rows = None
    if args.small_tables:
        rows = SMALL_ROWS
        for t in list(plan.tables.values()) + list(plan.bias_tables.values()):
            if t.name in rows:
                t.rows = rows[t.name]
    B = args.train_batch
    batches = [synthetic_batch(plan, B, seed=SEED + 50000 + 1000 * rank + i, id_mode=args.id_mode, table_rows=rows)
               for i in range(3)]
    trainer = Trainer(plan, device, world=world, rank=rank, seed=SEED,
                      precision="f32" if args.train_gemm == "f32" else "bf16", train_gemm=args.train_gemm)
    dev_batches = [batch_to(b, device) for b in batches]
    step = lambda i: trainer.train_step(dev_batches[i % len(dev_batches)])
    for i in range(3):
        step(i)
    trainer.enable_stage_timing(True)
    l0 = trainer.model.launches
    ms_staged = timed(step, args.train_steps)
    launches = trainer.model.launches - l0
    torch.cuda.synchronize()
    stage = trainer.stage_times_ms()
    trainer.enable_stage_timing(False)
    ms = min(ms_staged, timed(step, args.train_steps))
    a2a = trainer.last_a2a_bytes
    out = {
        "workload": "BASELINE config %d: DMT training step (fwd + bwd + TF-1 Adam, dense over every row), 615 dense + "
                    "all id sequences, MMoE 2 tasks, per-GPU batch %d (global %d), d_model=%d, %d heads, Sku vocabulary "
                    "%d%s, training-mode dropout (transformer %.2g, bias tower %s)"
                    % (4 if world > 1 else 3, B, B * world, plan.d_model, plan.num_heads, plan.tables["Sku"].rows,
                       " row-sharded over %d ranks" % world if world > 1 else "", plan.dropout_rate,
                       list(plan.dropout_rate_bias)),
        "value": world * B * args.train_steps / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms / args.train_steps,
        "steps": args.train_steps, "warmup": 3, "scaling": "weak",
        "dtype": {"f32": "f32", "bf16": "bf16 GEMM operands on tcgen05 / fp32 accumulate + storage",
                  "bf16x3": "split-bf16 (hi+lo) GEMM operands on tcgen05 / fp32 accumulate + storage",
                  "tf32": "tf32 GEMM operands (TMA-fed tcgen05 kind::tf32 straight from fp32 activations; MMoE split-bf16) / "
                          "fp32 accumulate + storage"}[args.train_gemm],
        "stage_ms": {k: round(t / args.train_steps, 4) for k, (t, _) in sorted(stage.items())},
        "gpu_launches": int(launches),
        "allreduce_bytes": trainer.allreduce_bytes if world > 1 else 0,
        "a2a_bytes": int(a2a),
        "collectives": ("per step and rank: 1 NCCL allreduce of the flat [dense | replicated small tables] gradient "
                        "bucket; Sku: all_gather of split points + 3 all_to_all_single (row ids, rows, gradient rows)")
                       if world > 1 else "none (1 GPU)",
        "tokens_per_step": sum(batch_tokens(plan, b) for b in batches) / len(batches),
        "loss": float(trainer.global_loss()),
    }
    del trainer, dev_batches
    torch.cuda.empty_cache()
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    from cikm2020_dmt_b200.data import PackedBatch, batch_to, batch_tokens
    from cikm2020_dmt_b200.inference import Inference
    from cikm2020_dmt_b200.params import ParamStore

    conf, plan, batches, rows = build_workload(args, device, rank)
    store = ParamStore(plan, device=device)
    inf = Inference(conf, params=store, precision=args.precision) if not args.small_tables else None
    if inf is None:
        from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
        model = mmoe_transformer_unbias(plan, params=store, precision=args.precision)
        infer = model.inference
    else:
        model = inf.model
        infer = inf.inference
    dev_batches = [batch_to(b, device) for b in batches]
    # the data loader's product: ONE pinned buffer per batch, allocated on the NUMA node of this rank's GPU.  For the
    # bf16 path it is the compact format (uint16 ids of the small vocabularies, bf16 features, inference keys only)
    from cikm2020_dmt_b200.numa import numa_local
    compact = args.precision == "bf16" and not args.wide_batch
    infer_keys = set(plan.all_id_features()) | {"features"}
    with numa_local(local_rank) as numa_node:
        packed = [PackedBatch(b, compact=compact, keys=infer_keys if compact else None) for b in batches]
    B = args.batch
    out_host = torch.empty(2, 3, B, dtype=torch.float32).pin_memory()   # double-buffered D2H landing zone
    out_done = [None, None]
    d2h_stream = torch.cuda.Stream(device)

    timed = make_timer(world, device)

    def step_resident(i):
        infer(dev_batches[i % len(dev_batches)], is_train=False)

    staged = []         # batches whose copy is in flight: [step i, step i + 1]

    def step_e2e(i):
        # every step copies ONE batch host->device (the batch of step i + 2, on the copy stream, overlapping the
        # kernels of steps i and i + 1 -- a data-loader prefetch two batches deep, the model rotates three device
        # buffers) and reads this step's scores back
        while len(staged) < 2:
            staged.append(model.prefetch(packed[(i + len(staged)) % len(packed)], views=False))
        # (before this step's launches: the copy into the third buffer then only waits for step i - 1, the buffer's
        # previous consumer)
        staged.append(model.prefetch(packed[(i + 2) % len(packed)], views=False))
        cur = staged.pop(0)
        (yr, yb) = infer(cur, is_train=False)
        # one device->host read of the step's scores ([click logit; order logit; y_bias] x B), on a side stream so
        # the next step's kernels do not queue behind the copy (the model alternates two score buffers)
        done = torch.cuda.Event()
        done.record()
        d2h_stream.wait_event(done)
        with torch.cuda.stream(d2h_stream):
            out_host[i & 1].copy_(model.last_scores, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(d2h_stream)
        out_done[i & 1] = ev
        # the caller consumes every step's scores, one step behind the launches (keeps the launch queue fed)
        prev = out_done[(i + 1) & 1]
        if prev is not None:
            prev.synchronize()

    # ---- warm-up, then the timed device-resident region (with per-stage events + clock sampling).  The W warm-up
    #      steps are followed by an untimed pre-heat of at least 0.3 s of back-to-back steps: 20 steps of this workload
    #      are 9 ms, too short for the SM clock to leave its idle state on a cold box
    for i in range(max(args.warmup, 3)):
        step_resident(i)
    torch.cuda.synchronize()
    t_heat = time.perf_counter()
    while time.perf_counter() - t_heat < 0.3:
        for i in range(10):
            step_resident(i)
        torch.cuda.synchronize()
    model.enable_stage_timing(True)
    for i in range(3):                       # the per-stage pass runs the sequences serially: warm that path as well
        step_resident(i)
    torch.cuda.synchronize()
    model.enable_stage_timing(True)          # (drops the warm-up's events)
    launches0 = model.launches
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.15)
    # per-stage pass (CUDA events around every stage, sequences serial): feeds `roofline` and `stage_share` only; it is
    # host-bound (two events per stage), so it runs at least 60 steps to average the idle gaps out
    stage_steps = max(args.steps, 60)
    # bf16: the dominant kernel's own launches are bracketed by CUDA events on their launch stream inside the library
    # (dmt_debug_seq_timer) -- in this pass, where nothing runs beside it (in the headline region the dense copy and
    # the pooled lookups share the SMs with it)
    import ctypes as _C
    seq_timer = None
    use_seq_timer = args.precision == "bf16" and getattr(model, "seq_multi", False)
    if use_seq_timer:
        model.lib.dmt_debug_seq_timer(1)
    timed(step_resident, stage_steps)
    if use_seq_timer:
        t_k, n_k = _C.c_float(0), _C.c_int32(0)
        if model.lib.dmt_debug_seq_timer_read(_C.byref(t_k), _C.byref(n_k)) == 0 and n_k.value == stage_steps:
            seq_timer = (float(t_k.value), int(n_k.value))
        model.lib.dmt_debug_seq_timer(0)
    gpu_launches = (model.launches - launches0) * args.steps // stage_steps
    torch.cuda.synchronize()
    stage = model.stage_times_ms()
    model.enable_stage_timing(False)
    # the timed region of the headline value: exactly K steps, no per-stage events
    ms = timed(step_resident, args.steps)
    for i in range(3):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None

    # ---- every rank: the fp32 forward block and the training / data-parallel blocks (collectives inside)
    def guarded(fn, *a):
        try:
            return fn(*a)
        except Exception as exc:      # a failing side block must not lose the headline line; it is reported
            import traceback
            return {"error": "%s: %s" % (type(exc).__name__, exc), "trace": traceback.format_exc()[-800:]}

    f32_block = None if args.no_f32 else guarded(run_f32_forward, args, plan, store, dev_batches, timed, world)
    tf32_block = tf32_dmt_block = None
    if not args.no_f32:
        tf32_block = guarded(run_tf32_forward, args, args.conf, rank, device, timed, world)
        tf32_dmt_block = guarded(run_tf32_forward, args, "dmt.conf", rank, device, timed, world)
    train_block = None
    if not args.no_train:
        dev_batches = None
        torch.cuda.empty_cache()
        train_block = guarded(run_train, args, world, rank, local_rank, device, timed)
        if isinstance(train_block, dict) and "error" not in train_block:
            train_block["dp_parity"] = guarded(run_dp_parity, args, world, rank, device)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * B * args.steps / (ms / 1e3)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    pk = peaks()

    # ---- roofline of the dominant stage (live CUDA-event time inside the timed region)
    stage_total = sum(t for t, _ in stage.values()) or 1.0
    shares = {k: round(t / stage_total, 4) for k, (t, _) in stage.items()}
    dom = max(stage, key=lambda k: stage[k][0])
    steps_used = stage_steps
    mean_b = lambda fn: sum(fn(plan, batches[i % len(batches)]) for i in range(steps_used))
    if dom in ("seq_encode", "seq_encode_train"):
        t_ms, _ = stage[dom]
        multi = args.precision == "bf16" and getattr(model, "seq_multi", False)
        # bf16: ONE persistent tile-kernel launch per step over all sequences (+ length-class and tail launches, whose
        # time is in t_ms); other paths: one launch chain per sequence
        n = steps_used * (1 if multi else len(plan.sequences))
        stage_ms_per_step = t_ms / steps_used
        if multi and seq_timer is not None:
            # the kernel's own events over the K timed steps (the stage time above also holds the length-class and
            # tail launches and, in the serial per-stage pass, launch gaps)
            t_ms, n = seq_timer
        alg = mean_b(algorithmic_bytes_seq)                      # bytes over all launches of the region
        own = multi and seq_timer is not None                    # t_ms is the tile kernel's own time
        fl = mean_b((lambda p, b: flops_seq(p, b, tail=False)) if own else flops_seq)   # algorithmic FLOPs (valid tokens)
        hbm = alg / (t_ms / 1e3) / 1e9
        tfl = fl / (t_ms / 1e3) / 1e12
        name = {"bf16": ("seq_encode_multi_kernel (bf16 tcgen05, two tiles in flight per SM, fused gather->encoder->decoder "
                         "scores / contexts, all sequences in one persistent launch over length-bucketed tiles)") if own else
                        ("seq_bucket_kernel + seq_encode_multi_kernel + seq_tail_kernel (bf16 tcgen05, two tiles in flight "
                         "per SM, fused gather->encoder->decoder, all sequences in one persistent launch over "
                         "length-bucketed tiles)") if multi else
                        ("seq_encode_tc3_kernel + seq_tail_kernel (bf16 tcgen05, two tiles in flight per SM, fused "
                         "gather->encoder->decoder, per sequence)"),
                "tf32": "row-batched pipeline: seq_gather + tf32_rows_kernel x5 + attention + LayerNorm kernels (tcgen05 "
                        "kind::tf32 GEMMs, per sequence)",
                "f32": "seq_encode_f32_kernel (fp32 CUDA cores, fused gather->encoder->decoder, per sequence)"}[args.precision]
        if args.precision == "bf16":
            # SURVEY 8d crossover: the fully fused kernel has 432 FLOP/B against a 209 FLOP/B ridge -> the tensor
            # roofline is the binding one; the HBM reading is kept as a secondary field
            roofline = {"kernel": name, "bound": "tensor", "achieved": tfl, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                        "frac": tfl / pk["bf16_tflops"], "traffic": None, "peak_source": pk["source"] + " (sustained)",
                        "launches": n, "avg_launch_ms": t_ms / n, "flops_per_launch": fl / n,
                        "algorithmic_bytes_per_launch": alg / n,
                        "timing": ("CUDA events around every seq_encode_multi_kernel launch on its launch stream, in the "
                                   "per-stage pass (dmt_debug_seq_timer)") if (multi and seq_timer is not None) else
                                  "CUDA events around the stage's launches in the per-stage pass",
                        "stage_ms_per_step": stage_ms_per_step,
                        "hbm": {"achieved": hbm, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm / pk["hbm_gbs"]},
                        "note": "fully fused kernel: 432 FLOP/B vs a 209 FLOP/B ridge -> tensor-bound by the roofline; "
                                "at L<=50 it is in practice limited by the SIMT epilogues between its six dependent MMA "
                                "round trips (issue slots 30 % busy, DESIGN.md 4.1); the HBM-bound gather is reported "
                                "under embed_gather"}
        else:
            roofline = {"kernel": name, "bound": "hbm", "achieved": hbm, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": hbm / pk["hbm_gbs"], "traffic": None, "peak_source": pk["source"],
                        "launches": n, "avg_launch_ms": t_ms / n, "algorithmic_bytes_per_launch": alg / n,
                        "flops_per_launch": fl / n, "achieved_tflops": tfl}
    else:
        t_ms, n = stage[dom]
        fl = flops_mmoe(plan, B) * steps_used
        achieved = fl / (t_ms / 1e3) / 1e12
        roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"],
                    "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops"], "traffic": None,
                    "peak_source": pk["source"], "launches": n, "avg_launch_ms": t_ms / n}
    roofline["stage_share"] = shares
    # DRAM bytes per launch of the same kernel from the committed ncu capture (when the workload matches)
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            tr = json.load(fh).get("seq_encode_multi_kernel" if (args.precision == "bf16" and getattr(model, "seq_multi", False))
                                   else "seq_encode_tc_kernel")
        wl = tr["workload"]
        if (dom in ("seq_encode", "seq_encode_train") and wl["conf"] == args.conf and wl["per_gpu_batch"] == args.batch
                and wl["precision"] == args.precision and wl["id_mode"] == args.id_mode and not args.small_tables):
            roofline["traffic"] = tr["mean_dram_bytes_per_launch"]
            roofline["traffic_source"] = tr["source"]
    except Exception:
        pass

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        from cikm2020_dmt_b200.data import SparseIds
        from oracle import dmt_oracle as O
        P = O.params_from_store(store, torch.float32)
        nb = min(args.cpu_batch, B)
        sub = {}
        for k, v in batches[0].items():
            if isinstance(v, SparseIds):
                hi = int(v.offsets[nb])
                sub[k] = SparseIds(v.values[:hi], v.offsets[:nb + 1], None if v.weights is None else v.weights[:hi])
            else:
                sub[k] = v[:nb]
        sps, reps = cpu_oracle_throughput(plan, P, sub, args.cpu_seconds)
        cpu_baseline = {"value": sps, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": "oracle port (PyTorch CPU fp32, lean lookups), first %d samples of batch 0, "
                                  "median of %d forward passes" % (nb, reps)}
        del P

    # ---- BASELINE metric part 2: stand-alone embedding gather (Sku rows of this workload) vs HBM peak
    embed_gather = None
    try:
        from cikm2020_dmt_b200 import abi as _abi
        sku = store.table("Sku")
        n_ids = 4096 * 50
        gsets = [torch.randint(1, sku.shape[0] + 1, (n_ids,), device=device, dtype=torch.int32) for _ in range(4)]
        gout = torch.empty(n_ids, sku.shape[1], device=device)
        st = torch.cuda.current_stream().cuda_stream
        for i in range(3):
            _abi.check(model.lib.dmt_embed_gather(sku.data_ptr(), sku.shape[0], sku.shape[1], gsets[i].data_ptr(),
                                                  n_ids, 1, gout.data_ptr(), st))
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(20):
            _abi.check(model.lib.dmt_embed_gather(sku.data_ptr(), sku.shape[0], sku.shape[1],
                                                  gsets[i % 4].data_ptr(), n_ids, 1, gout.data_ptr(), st))
        g1.record()
        torch.cuda.synchronize()
        gms = g0.elapsed_time(g1) / 20
        gbytes = n_ids * (sku.shape[1] * 4 + 4) + n_ids * sku.shape[1] * 4
        embed_gather = {"kernel": "embed_gather_kernel (B=4096, L=50, Sku %dx%d fp32, uniform ids)" % tuple(sku.shape),
                        "bound": "hbm", "achieved": gbytes / (gms / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": gbytes / (gms / 1e3) / 1e9 / pk["hbm_gbs"], "avg_launch_ms": gms,
                        "algorithmic_bytes_per_launch": gbytes}
    except Exception as exc:   # diagnostics only; never fail the bench line
        embed_gather = {"error": str(exc)}

    h2d = sum(p.nbytes for p in packed) / len(packed)
    line = {
        "metric": "samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": workload_config(args, plan),
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(out_host[0].numel() * 4), "ms_per_step": ms_e2e / args.steps,
                "format": ("compact pinned batch: ids in the narrowest byte width (1: time buckets, 2: vocabularies "
                           "< 65536, 3: Sku / Brand / Shopid < 2^24), bf16 features, inference keys only; widened on the "
                           "device by dmt_widen_ids (1 launch, copy stream)") if compact else
                          "int32 ids + fp32 features, one pinned buffer",
                "pinned_numa_node": numa_node,
                "pipeline": "prefetch two batches deep (three rotating device buffers): step i+2's packed batch is "
                            "copied on a side stream while steps i and i+1 compute; one H2D copy and one D2H read "
                            "(side stream) per step inside the timed region; the host waits for step i-1's scores "
                            "after launching step i"},
        "embed_gather": embed_gather,
        "f32": f32_block, "tf32": tf32_block, "tf32_dmt_conf": tf32_dmt_block, "train": train_block,
        "gpu_launches": int(gpu_launches), "clocks": clocks,
        "tokens_per_step": sum(batch_tokens(plan, b) for b in batches) / len(batches),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
