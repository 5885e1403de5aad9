import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from cikm2020_dmt_b200.conf import Conf
from cikm2020_dmt_b200.data import synthetic_batch, batch_to, SEED
from cikm2020_dmt_b200.plan import build_plan
from cikm2020_dmt_b200.train import Trainer
conf = Conf(os.path.join(ROOT, "conf", "settings") + "/", "dmt_d64.conf")
plan = build_plan(conf)
batches = [batch_to(synthetic_batch(plan, 8192, seed=SEED + i), "cuda") for i in range(2)]
tr = Trainer(plan, "cuda", precision="bf16", train_gemm="bf16x3")
for i in range(4):
    tr.train_step(batches[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(4):
        tr.train_step(batches[i % 2])
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print("span %.2f ms, kernel busy %.2f ms, kernels %d (4 steps)" % ((t1 - t0) / 1e3, busy / 1e3, len(evs)))
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    agg[e.name[:60]][0] += 1
    agg[e.name[:60]][1] += (e.time_range.end - e.time_range.start) / 1e3
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:22]:
    print("%-62s n=%4d total=%8.2f ms" % (k, n, t))
# gaps
gaps = []
for a, b in zip(evs[:-1], evs[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 20:
        gaps.append((g, a.name[:40], b.name[:40]))
gaps.sort(reverse=True)
print("gaps > 20us: %d, total %.2f ms" % (len(gaps), sum(g for g, _, _ in gaps) / 1e3))
for g in gaps[:12]:
    print("   %.0f us between %s -> %s" % g)
