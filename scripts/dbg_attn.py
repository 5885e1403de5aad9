import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import make_plan, SMALL_ROWS
from cikm2020_dmt_b200.params import ParamStore
from cikm2020_dmt_b200.data import synthetic_batch, batch_to
from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias
conf, plan = make_plan("dmt_d64.conf")
store = ParamStore(plan, device="cuda", seed=84).randomize_(85)
B = 300
host = synthetic_batch(plan, B, seed=86, table_rows=SMALL_ROWS)
dev = batch_to(host, "cuda")
tf = mmoe_transformer_unbias(plan, params=store, precision="tf32")
out = torch.zeros(B, 3 * plan.d_model, device="cuda")
for s in range(3):
    tf.seq_encode(dev, s, out.data_ptr() + 4 * s * plan.d_model, out.stride(0), B)
torch.cuda.synchronize()
torch.save(out.cpu(), sys.argv[1])
if len(sys.argv) > 2:
    ref = torch.load(sys.argv[2])
    err = (out.cpu() - ref).abs()
    print("max err", err.max().item(), "mean", err.mean().item())
    for s in range(3):
        e = err[:, s * 64:(s + 1) * 64].max(1).values
        bad = (e > 1e-2).nonzero().flatten().tolist()
        off = host[plan.sequences[s].user_features[-1]].offsets
        print("seq", s, "bad samples", len(bad), bad[:20], "starts", [int(off[b]) for b in bad[:20]], "lens", [int(off[b+1]-off[b]) for b in bad[:20]])
