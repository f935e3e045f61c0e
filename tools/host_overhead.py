"""Host-side cost of enqueueing one fused CLR step (CLRPlan.run: ONE C call, 6 kernel launches) on a tiny problem, so the
GPU is never the bottleneck.  Measured on the B200 box: ~25 us of host time per step against 178 us of device time at
the bench size -- the step is device-bound with a wide margin."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uda_clr_b200 as clr
from uda_clr_b200 import synth
dev = torch.device("cuda", 0)
b = synth.make_batch(B=1, C=32, H=32, W=32, K=2, T=8, up=4, seed=1)     # tiny: the GPU is never the bottleneck
t = {k: getattr(b, k).to(dev) for k in ("xs", "ys", "xt", "oT_before", "preds", "oT", "oT_aug")}
step = clr.CLRStep(K=2, retrify=True, use_disc=True, use_cons=True)
plan = step.plan(t["xs"], t["ys"], t["xt"], oT_before=t["oT_before"], preds=t["preds"], T=8, oT=t["oT"], oT_aug=t["oT_aug"])
for _ in range(50): plan.run()
torch.cuda.synchronize()
n = 2000
t0 = time.perf_counter()
for _ in range(n): plan.run()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue per step: %.1f us (6 launches); drained after %.1f us more per step" % ((t1 - t0) / n * 1e6, (t2 - t1) / n * 1e6))
