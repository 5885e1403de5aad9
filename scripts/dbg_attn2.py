import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_dropout as TD
plan, model, store, host, dev, O = TD._setup("dmt_d64.conf", 200, seed=21, precision="bf16", train_gemm="tf32")
(yr, yb) = model.inference(dev, is_train=True, dropout_seed=4242)
torch.cuda.synchronize()
x = model._last["x"].clone().cpu()
torch.save(x, sys.argv[1])
if len(sys.argv) > 2:
    ref = torch.load(sys.argv[2])
    c0 = plan.interest_col
    for s in range(3):
        e = (x[:, c0 + 64 * s:c0 + 64 * (s + 1)] - ref[:, c0 + 64 * s:c0 + 64 * (s + 1)]).abs().max(1).values
        bad = (e > 1e-2).nonzero().flatten().tolist()
        off = host[plan.sequences[s].user_features[-1]].offsets
        print("seq", s, "max", e.max().item(), "bad", len(bad), bad[:16], "lens", [int(off[b + 1] - off[b]) for b in bad[:16]])
