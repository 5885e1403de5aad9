import cProfile, pstats, sys, os, time, io
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from cikm2020_dmt_b200.conf import Conf
from cikm2020_dmt_b200.data import synthetic_batch, batch_to, SEED
from cikm2020_dmt_b200.plan import build_plan
from cikm2020_dmt_b200.train import Trainer
conf = Conf(os.path.join(ROOT, "conf", "settings") + "/", "dmt_d64.conf")
plan = build_plan(conf)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
batches = [batch_to(synthetic_batch(plan, B, seed=SEED + i), "cuda") for i in range(2)]
tr = Trainer(plan, "cuda", precision="bf16", train_gemm="bf16x3")
for i in range(3):
    tr.train_step(batches[i % 2])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    tr.train_step(batches[i % 2])
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("host issue time per step %.2f ms, wall per step %.2f ms" % (t_host * 100, t_all * 100))
pr = cProfile.Profile()
pr.enable()
for i in range(10):
    tr.train_step(batches[i % 2])
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])
