"""Time dmt_pool_mean_fwd on subsets of the pooled features (which lookups cost what)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cikm2020_dmt_b200 import abi
from cikm2020_dmt_b200.conf import Conf
from cikm2020_dmt_b200.plan import build_plan
from cikm2020_dmt_b200.params import ParamStore
from cikm2020_dmt_b200.data import synthetic_batch, batch_to
from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
conf = Conf(os.path.join(ROOT, "conf", "settings") + "/", "dmt_d64.conf")
plan = build_plan(conf)
store = ParamStore(plan, device="cuda")
model = mmoe_transformer_unbias(plan, params=store, precision="bf16")
B = 4096
devs = [batch_to(synthetic_batch(plan, B, seed=100 + i), "cuda") for i in range(4)]
x = torch.zeros(B, 1088, device="cuda")

def run(name, specs):
    for i in range(3):
        model.pool_mean(devs[i % 4], specs, False, x, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        model.pool_mean(devs[i % 4], specs, False, x, B)
    e1.record()
    torch.cuda.synchronize()
    nnz = sum(int(devs[0][p.feature].values.numel()) for p in specs)
    byts = sum(int(devs[0][p.feature].values.numel()) * plan.tables[p.table].dim * 4 for p in specs)
    ms = e0.elapsed_time(e1) / 20
    print("%-28s feats %2d lookups %8d row-bytes %6.1f MB  %7.1f us  %6.0f GB/s" % (name, len(specs), nnz, byts / 1e6, ms * 1e3, byts / ms / 1e6))

P = plan.pooled
run("all", P)
run("sku only", [p for p in P if p.table == "Sku"])
run("non-sku", [p for p in P if p.table != "Sku"])
run("item feats", [p for p in P if not p.feature.startswith(("clk", "ord", "cart"))])
run("clk feats", [p for p in P if p.feature.startswith("clk")])
run("clk sku", [p for p in P if p.feature.startswith("clk") and p.table == "Sku"])
run("clk time", [p for p in P if p.feature.startswith("clk") and p.table.startswith("Time")])
for p in P:
    if p.feature.startswith("clk"):
        run(p.feature, [p])
