"""Where does the pooled-lookup kernel spend its time?  Times dmt_pool_mean_fwd(_bf16) over subsets of the features."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cikm2020_dmt_b200.conf import Conf
from cikm2020_dmt_b200.plan import build_plan
from cikm2020_dmt_b200.params import ParamStore
from cikm2020_dmt_b200.data import synthetic_batch, batch_to
from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
plan = build_plan(Conf(os.path.join(ROOT, "conf", "settings") + "/", "dmt_d64.conf"))
store = ParamStore(plan, device="cuda")
model = mmoe_transformer_unbias(plan, params=store, precision="bf16")
B = 4096
devs = [batch_to(synthetic_batch(plan, B, seed=i), "cuda") for i in range(4)]
ld = (plan.mmoe_in + 7) // 8 * 8
for dt in (torch.bfloat16, torch.float32):
    x = torch.zeros(B, ld, dtype=dt, device="cuda")
    subsets = {
        "all": list(plan.pooled),
        "sku only": [p for p in plan.pooled if p.table == "Sku"],
        "small only": [p for p in plan.pooled if p.table != "Sku"],
        "seq sku only": [p for p in plan.pooled if p.table == "Sku" and "seq" in p.feature],
        "item only": [p for p in plan.pooled if "seq" not in p.feature],
        "clk seq": [p for p in plan.pooled if p.feature.startswith("clk_seq")],
    }
    for name, specs in subsets.items():
        for i in range(3):
            model.pool_mean(devs[i % 4], specs, False, x, B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            model.pool_mean(devs[i % 4], specs, False, x, B)
        e1.record()
        torch.cuda.synchronize()
        print("%s %-14s %2d features: %.1f us" % (str(dt)[6:], name, len(specs), e0.elapsed_time(e1) / 20 * 1e3))
