"""Host-side cost of one forward step through the plugin surface (prefetch + inference + D2H), B = 4096."""
import cProfile, pstats, sys, os, time, io
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from cikm2020_dmt_b200.conf import Conf
from cikm2020_dmt_b200.data import synthetic_batch, PackedBatch, SEED
from cikm2020_dmt_b200.plan import build_plan
from cikm2020_dmt_b200.params import ParamStore
from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
conf = Conf(os.path.join(ROOT, "conf", "settings") + "/", "dmt_d64.conf")
plan = build_plan(conf)
B = 4096
store = ParamStore(plan, device="cuda")
model = mmoe_transformer_unbias(plan, params=store, precision="bf16")
packed = [PackedBatch(synthetic_batch(plan, B, seed=SEED + i)) for i in range(4)]
out_host = torch.empty(3, B).pin_memory()

def step(i):
    cur = model.prefetch(packed[i % 4], views=False)
    (yr, yb) = model.inference(cur, is_train=False)
    out_host[0].copy_(yr[0].view(-1), non_blocking=True)
    out_host[1].copy_(yr[1].view(-1), non_blocking=True)
    out_host[2].copy_(yb.view(-1), non_blocking=True)

for i in range(5):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50):
    step(i)
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("host issue time per step %.3f ms, wall per step %.3f ms" % (t_host * 20, t_all * 20))
# H2D alone
buf = torch.empty(packed[0].nbytes, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20):
    packed[i % 4].to("cuda", out=buf)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print("H2D of one packed batch: %.3f ms = %.1f GB/s (%d bytes)" % (dt * 1e3, packed[0].nbytes / dt / 1e9, packed[0].nbytes))
pr = cProfile.Profile()
pr.enable()
for i in range(50):
    step(i)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
print(s.getvalue()[:7000])
