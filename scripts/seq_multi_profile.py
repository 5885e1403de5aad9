"""Per-phase SM-cycle breakdown and per-CTA durations of the multi-sequence bf16 tile kernel
(dmt_debug_seq_profile + dmt_seq_encode_multi_fwd)."""
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
B = int(os.environ.get("B", 4096))
dev = batch_to(synthetic_batch(plan, B), "cuda")
x_ld = (plan.mmoe_in + 3) // 4 * 4
x = torch.zeros(B, x_ld, device="cuda")
names = ["P0 convert+sync", "P1 QKV mma wait", "P2 QKV epilogue+sync", "P3 S mma wait", "P4 softmax->TMEM+sync",
         "P5 PV mma wait (+qt)", "P6 LN1+sync", "P7 FF1 mma wait", "P8 relu->TMEM+sync", "P9 FF2 mma wait",
         "P10 LN2+scores+images+sync", "P11 ctx mma issue", "prime (per segment)", "drain (per segment)"]
lib = abi.load()
for it in range(3):
    cnt = torch.zeros(2048, dtype=torch.int64, device="cuda")
    lib.dmt_debug_seq_profile(cnt.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.seq_encode_multi(dev, x, x_ld, B)
    e1.record()
    torch.cuda.synchronize()
    lib.dmt_debug_seq_profile(None)
c = cnt.cpu().tolist()
print("event time of bucket + multi + tail launches: %.1f us" % (e0.elapsed_time(e1) * 1e3))
for q in range(len(plan.sequences)):
    v = c[q * 16:q * 16 + 16]
    nt = max(v[14], 1)
    tot = sum(v[:14])
    print("sequence %d: %d tiles of CTA 0 / group 0, %.0f cycles/tile in the group (prologue %d cycles)" % (q, v[14], tot / nt, v[15]))
    for n, val in zip(names, v[:14]):
        print("   %-28s %8.0f cyc/tile  %5.1f%%   (total %d)" % (n, val / nt, 100.0 * val / max(tot, 1), val))
ncta = sum(1 for i in range(256) if c[64 + i] > 0)
cyc = c[64:64 + ncta]
t0 = c[320:320 + ncta]
t1 = c[576:576 + ncta]
print("CTAs %d: cycles min %d mean %.0f max %d" % (ncta, min(cyc), sum(cyc) / ncta, max(cyc)))
print("globaltimer: first entry -> last exit %.1f us; entry skew %.1f us; exit skew %.1f us; mean CTA life %.1f us"
      % ((max(t1) - min(t0)) / 1e3, (max(t0) - min(t0)) / 1e3, (max(t1) - min(t1)) / 1e3,
         sum(b - a for a, b in zip(t0, t1)) / ncta / 1e3))

print("timeline of CTA 0 (cycles since kernel entry at the END of each phase), group 0 | group 1:")
for t in range(6):
    for i in range(12):
        a, b = c[1024 + t * 16 + i], c[1024 + 128 + t * 16 + i]
        print("  tile %d %-28s %8d | %8d" % (t, names[i], a, b))
