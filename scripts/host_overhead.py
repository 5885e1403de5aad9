"""Host-side launch cost of one inference() call vs the GPU time of the step (is the step launch-bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cikm2020_dmt_b200.conf import Conf
from cikm2020_dmt_b200.plan import build_plan
from cikm2020_dmt_b200.params import ParamStore
from cikm2020_dmt_b200.data import synthetic_batch, batch_to, PackedBatch
from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
conf = Conf(os.path.join(ROOT, "conf", "settings") + "/", "dmt_d64.conf")
plan = build_plan(conf)
store = ParamStore(plan, device="cuda")
model = mmoe_transformer_unbias(plan, params=store, precision="bf16")
B = int(os.environ.get("B", 4096))
host = [synthetic_batch(plan, B, seed=i) for i in range(4)]
dev = [batch_to(b, "cuda") for b in host]
for i in range(20):
    model.inference(dev[i % 4], is_train=False)
torch.cuda.synchronize()
N = 200
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for i in range(N):
    model.inference(dev[i % 4], is_train=False)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("resident: host launch %.1f us/step, GPU %.1f us/step" % ((t1 - t0) / N * 1e6, e0.elapsed_time(e1) / N * 1e3))
keys = set(plan.all_id_features()) | {"features"}
packed = [PackedBatch(b, compact=True, keys=keys) for b in host]
staged = [None]
def step(i):
    cur = staged[0] if staged[0] is not None else model.prefetch(packed[i % 4], views=False)
    staged[0] = model.prefetch(packed[(i + 1) % 4], views=False)
    model.inference(cur, is_train=False)
for i in range(20):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
e0.record()
for i in range(N):
    step(i)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("prefetch + inference: host launch %.1f us/step, GPU %.1f us/step" % ((t1 - t0) / N * 1e6, e0.elapsed_time(e1) / N * 1e3))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for i in range(100):
    step(i)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
