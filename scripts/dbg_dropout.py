import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_dropout as TD
engine = sys.argv[1]
plan, model, store, host, dev, O = TD._setup("dmt_d64.conf", 200, seed=21, precision="bf16", train_gemm=engine)
P = O.params_from_store(store)
O.DROPOUT_HOOK = TD._hook(plan, host, 4242)
try:
    loss_ref, grads_ref, _ = O.loss_and_grads(plan, P, host, is_train=True)
finally:
    O.DROPOUT_HOOK = None
loss, G = model.compute_gradients(dev, is_train=True, dropout_seed=4242)
torch.cuda.synchronize()
print("loss", loss.item(), loss_ref.item())
rows = []
for name in [s.name for s in store.specs] + list(store.tables):
    want = grads_ref.get(name)
    want = torch.zeros_like(P[name]) if want is None else want.double()
    got = (G.table_dense(store, name) if name in store.tables else G[name]).detach().double().cpu().reshape(want.shape)
    cos = float((got * want).sum() / (got.norm() * want.norm() + 1e-300))
    rel = float((got - want).norm() / (want.norm() + 1e-300))
    rows.append((cos, rel, name[-60:]))
for r in sorted(rows)[:12]:
    print("%.5f %.4f %s" % r)
