#!/usr/bin/env python
"""BASELINE config 5: Sku row-gather sweep -- achieved HBM GB/s of `dmt_embed_gather` vs the measured peak.

    python bench_gather.py [--quick]

V in {1M, 10M, 100M} x L in {50, 200} x B in {4096 .. 32768}, D = 32 fp32 rows (128 B), uniform ids,
zero-pad addressing.  Algorithmic bytes: B*L*(D*4 + 4) read + B*L*D*4 written.
Prints one JSON line per point and a summary line.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    import torch
    from cikm2020_dmt_b200 import abi
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    lib = abi.load()
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    src = "fallback"
    if os.path.exists(p):
        peak, src = float(json.load(open(p))["hbm_gbs"]), "measured"
    D = 32
    vocabs = [1_000_000, 10_000_000] if args.quick else [1_000_000, 10_000_000, 100_000_000]
    points = []
    stream = torch.cuda.current_stream().cuda_stream
    for V in vocabs:
        table = torch.empty(V, D, device="cuda").uniform_(-1e-3, 1e-3)
        for L in (50, 200):
            for B in ((4096, 32768) if args.quick else (4096, 8192, 16384, 32768)):
                n = B * L
                g = torch.Generator(device="cuda").manual_seed(V % 1000 + L + B)
                ids = [torch.randint(1, V + 1, (n,), device="cuda", dtype=torch.int32, generator=g) for _ in range(4)]
                out = torch.empty(n, D, device="cuda")
                for i in range(3):
                    abi.check(lib.dmt_embed_gather(table.data_ptr(), V, D, ids[i % 4].data_ptr(), n, 1, out.data_ptr(), stream))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for i in range(args.iters):
                    abi.check(lib.dmt_embed_gather(table.data_ptr(), V, D, ids[i % 4].data_ptr(), n, 1, out.data_ptr(), stream))
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.iters
                byts = n * (D * 4 + 4) + n * D * 4
                gbs = byts / (ms / 1e3) / 1e9
                pt = {"V": V, "L": L, "B": B, "ms": ms, "algorithmic_bytes": byts, "achieved_gbs": gbs,
                      "frac_of_%s_peak" % src: gbs / peak, "table_mb": V * D * 4 / 1e6}
                points.append(pt)
                print(json.dumps(pt), flush=True)
                del ids, out
        del table
        torch.cuda.empty_cache()
    best = max(points, key=lambda q: q["achieved_gbs"])
    worst = min(points, key=lambda q: q["achieved_gbs"])
    print(json.dumps({"summary": "dmt_embed_gather sweep", "peak_gbs": peak, "peak_source": src,
                      "best": best, "worst": worst}))


if __name__ == "__main__":
    main()
