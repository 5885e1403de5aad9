#!/usr/bin/env python
"""bench_train.py -- BASELINE configs 3 / 4: one DMT training step (forward with saved activations, backward,
TF-1 Adam over dense + embedding variables) per GPU batch, data-parallel over N GPUs.

    python bench_train.py --batch 8192 --steps 20 --warmup 5                 # config 3, 1 GPU
    torchrun --nproc-per-node 8 ... bench_train.py --gpus 8 --batch 8192     # config 4 (global 65536)

Prints one JSON line (rank 0): samples/s (whole job), ms/step (max over ranks, CUDA events between
barriers), the per-stage share of the step and, with --cpu-seconds > 0, the CPU oracle's training step
(autograd + dense TF-Adam, what the reference's graph does) timed on a bounded sample.
This is NOT the headline bench (bench.py = config 2, forward only); it is the measurement of SURVEY 8(d)
configs 3 and 4.
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=8192, help="per-GPU batch (BASELINE config 3: 8192)")
    ap.add_argument("--conf", default="dmt_d64.conf")
    ap.add_argument("--id-mode", default="uniform", choices=["uniform", "zipf"])
    ap.add_argument("--n-batches", type=int, default=3)
    ap.add_argument("--cpu-batch", type=int, default=256)
    ap.add_argument("--cpu-seconds", type=float, default=0.0)
    ap.add_argument("--small-tables", action="store_true")
    ap.add_argument("--no-optimizer", action="store_true", help="debug: gradients only")
    ap.add_argument("--no-dropout", action="store_true", help="set the conf's dropout rates to 0")
    ap.add_argument("--train-gemm", default="bf16x3", choices=["f32", "bf16", "bf16x3", "tf32"],
                    help="engine of the training-path GEMMs: fp32 SIMT | tcgen05 bf16 | tcgen05 split-bf16 (fp32-grade)")
    return ap.parse_args()


NO_DROPOUT = {("model", "transformer_dropout_rate"): "0.0", ("model", "dropout_rate_bias"): "0.0,0.0"}


def main():
    args = parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench_train.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    from cikm2020_dmt_b200.conf import Conf
    from cikm2020_dmt_b200.data import synthetic_batch, batch_to, batch_tokens, SEED
    from cikm2020_dmt_b200.plan import build_plan
    from cikm2020_dmt_b200.params import ParamStore
    from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
    from cikm2020_dmt_b200.train import Trainer

    conf = Conf(os.path.join(ROOT, "conf", "settings") + "/", args.conf,
                overrides=NO_DROPOUT if args.no_dropout else None)
    plan = build_plan(conf)
    rows = None
    if args.small_tables:
        rows = {"Sku": 20000, "Brand": 2000, "Shopid": 2000, "Cid3": 1000, "Cid2": 100}
        for t in list(plan.tables.values()) + list(plan.bias_tables.values()):
            if t.name in rows:
                t.rows = rows[t.name]
    batches = [synthetic_batch(plan, args.batch, seed=SEED + 1000 * rank + i, id_mode=args.id_mode, table_rows=rows)
               for i in range(args.n_batches)]
    trainer = Trainer(plan, device=device, learning_rate=conf.learning_rate if hasattr(conf, "learning_rate") else 1e-3,
                      world=world, rank=rank, seed=SEED, precision="f32" if args.train_gemm == "f32" else "bf16",
                      train_gemm=args.train_gemm)
    model = trainer.model
    dev_batches = [batch_to(b, device) for b in batches]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        if args.no_optimizer:
            model.compute_gradients(dev_batches[i % len(dev_batches)])
        else:
            trainer.train_step(dev_batches[i % len(dev_batches)])

    def timed(n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            step(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(max(args.warmup, 3)):
        step(i)
    trainer.enable_stage_timing(True)
    launches0 = model.launches
    from bench import ClockSampler
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.15)
    ms_staged = timed(args.steps)
    launches = model.launches - launches0
    torch.cuda.synchronize()
    stage = trainer.stage_times_ms()
    trainer.enable_stage_timing(False)
    ms = min(ms_staged, timed(args.steps))
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total = sum(t for t, _ in stage.values()) or 1.0
    line = {
        "metric": "training samples/sec", "value": world * args.batch * args.steps / (ms / 1e3), "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak",
        "dtype": {"f32": "f32", "bf16": "bf16 GEMM operands / fp32 accumulate+storage",
                  "bf16x3": "split-bf16 (hi+lo) GEMM operands on tcgen05 / fp32 accumulate+storage",
                  "tf32": "tf32 GEMM operands (TMA-fed tcgen05 kind::tf32) / fp32 accumulate+storage"}[args.train_gemm],
        "data": "synthetic",
        "config": {"workload": "BASELINE config %d: DMT training step (fwd + bwd + TF-1 Adam, dense over every row), "
                               "615 dense + all id sequences, MMoE 2 tasks, per-GPU batch %d, d_model=%d, %d heads, "
                               "Sku vocabulary %d %s" % (4 if world > 1 else 3, args.batch, plan.d_model, plan.num_heads,
                                                         plan.tables["Sku"].rows,
                                                         "row-sharded over %d ranks" % world if world > 1 else ""),
                   "conf": args.conf,
                   "dropout": "off" if args.no_dropout else "training mode: transformer %.2g, bias tower %s"
                              % (plan.dropout_rate, list(plan.dropout_rate_bias))},
        "stage_ms_per_step": {k: round(t / args.steps, 4) for k, (t, _) in sorted(stage.items())},
        "stage_share": {k: round(t / total, 4) for k, (t, _) in sorted(stage.items())},
        "gpu_launches": int(launches), "clocks": clocks,
        "tokens_per_step": sum(batch_tokens(plan, b) for b in batches) / len(batches),
    }
    if args.cpu_seconds > 0:
        from cikm2020_dmt_b200.data import SparseIds
        from oracle import dmt_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        P = O.params_from_store(model.params, torch.float32)
        nb = min(args.cpu_batch, args.batch)
        sub = {}
        for k, v in batches[0].items():
            if isinstance(v, SparseIds):
                hi = int(v.offsets[nb])
                sub[k] = SparseIds(v.values[:hi], v.offsets[:nb + 1], None if v.weights is None else v.weights[:hi])
            else:
                sub[k] = v[:nb]
        opt = O.TFAdam(P, lr=1e-3)
        times = []
        t_end = time.perf_counter() + args.cpu_seconds
        while len(times) < 2 or (time.perf_counter() < t_end and len(times) < 50):
            t0 = time.perf_counter()
            _, grads, _ = O.loss_and_grads(plan, P, sub)
            opt.step(grads)
            times.append(time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": nb / statistics.median(times), "unit": "samples/s",
                                "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "oracle port training step (autograd + dense TF-Adam over every table row, "
                                          "what the reference graph does), %d samples, median of %d steps"
                                          % (nb, len(times))}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
