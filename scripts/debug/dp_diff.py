import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from conftest import SMALL_ROWS, make_plan
from test_gpu_dist_train import _batches, NO_DROPOUT
from cikm2020_dmt_b200.data import batch_to
from cikm2020_dmt_b200.train import Trainer
conf, plan = make_plan("dmt_d64.conf", overrides=NO_DROPOUT)
a = Trainer(plan, "cuda", seed=3, randomize=4)
b = Trainer(plan, "cuda", seed=3, randomize=4, force_dp_path=True)
for step, h in enumerate(_batches(plan, 48, 3)):
    la = a.train_step(batch_to(h, "cuda")); lb = b.train_step(batch_to(h, "cuda"))
    torch.cuda.synchronize()
    print("step", step, la.item(), lb.item())
    worst = []
    for name, va in a.store.named_parameters():
        vb = b.store.views[name]
        e = (va - vb).abs()
        worst.append((e.max().item(), e.mean().item(), name[-60:]))
    worst.sort(reverse=True)
    for w in worst[:6]:
        print("   %.3e %.3e %s" % w)
