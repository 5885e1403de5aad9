"""Per-phase SM-cycle breakdown of the bf16 sequence kernel (dmt_debug_seq_profile)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cikm2020_dmt_b200 import abi
from cikm2020_dmt_b200.conf import Conf
from cikm2020_dmt_b200.plan import build_plan
from cikm2020_dmt_b200.params import ParamStore
from cikm2020_dmt_b200.data import synthetic_batch, batch_to
from cikm2020_dmt_b200.net.mmoe_transformer_unbias import mmoe_transformer_unbias

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
conf = Conf(os.path.join(ROOT, "conf", "settings") + "/", "dmt_d64.conf")
plan = build_plan(conf)
store = ParamStore(plan, device="cuda")
model = mmoe_transformer_unbias(plan, params=store, precision="bf16")
B = 4096
dev = batch_to(synthetic_batch(plan, B), "cuda")
out = torch.zeros(B, 3 * plan.d_model, device="cuda")
V1 = False      # the first-generation kernel (one tile in flight per SM) was removed in round 2
names2 = ["P0 convert+sync", "P1 QKV mma wait", "P2 QKV epilogue+sync", "P3 S mma wait", "P4 softmax->TMEM+sync",
          "P5 PV mma wait (+qt)", "P6 LN1+sync", "P7 FF1 mma wait", "P8 relu->TMEM+sync", "P9 FF2 mma wait",
          "P10 LN2+scores+images+sync", "P11 ctx mma + readout"]
names = ["top sync", "P0 convert+sync", "P1 QKV mma wait", "P2 QKV epilogue+sync", "P3 S mma wait", "P4 softmax+sync",
         "P5 PV mma wait", "P6 LN1+sync", "P7 FF1 mma wait", "P8 relu epilogue+sync", "P9 FF2 mma wait", "P10 LN2+sync",
         "P11 decoder"]
lib = abi.load()
for s in range(3):
    for it in range(2):
        cnt = torch.zeros(16, dtype=torch.int64, device="cuda")
        lib.dmt_debug_seq_profile(cnt.data_ptr())
        model.seq_encode(dev, s, out.data_ptr() + 4 * s * plan.d_model, out.stride(0), B)
        torch.cuda.synchronize()
        lib.dmt_debug_seq_profile(None)
    c = cnt.cpu().tolist()
    tot = sum(c[:13])
    ntiles = (B + 1) // 2 if s < 2 else (B + 7) // 8
    if not V1:
        ntiles = (ntiles + 1) // 2      # v2: thread 0 only sees the tiles of group 0
        print("sequence %d: %d tiles of group 0, %.0f cycles/tile in the group = %.0f cycles/tile per SM"
              % (s, ntiles, tot / ntiles, tot / ntiles / 2))
    else:
        print("sequence %d: %d tiles, %.0f cycles/tile (thread 0 view)" % (s, ntiles, tot / ntiles))
    for n, v in zip(names if V1 else names2, c):
        print("   %-24s %7.0f cyc/tile  %5.1f%%" % (n, v / ntiles, 100.0 * v / tot))
