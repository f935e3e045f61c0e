#!/usr/bin/env python3
"""Where does the host time of the zero-line drop-in go?  cProfile of the trainer protocol (bench.trainer_protocol_step) with
this package's ops: top functions by cumulative and by own time.  GPU box only."""
import cProfile
import io
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import uda_clr_b200 as clr  # noqa: E402
from bench import trainer_protocol_step  # noqa: E402
from uda_clr_b200 import synth  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    hb = [synth.make_batch(B=8, C=256, H=128, W=128, K=2, T=8, up=4, seed=1234 + s) for s in range(2)]
    devb = [{k: getattr(h, k).to(dev) for k in ("xs", "ys", "xt", "oT_before", "preds")} for h in hb]
    st = {}
    for i in range(20):
        trainer_protocol_step(clr, st, devb[i % 2], 8)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for i in range(200):
        trainer_protocol_step(clr, st, devb[i % 2], 8)
    torch.cuda.synchronize()
    pr.disable()
    for key in ("cumulative", "tottime"):
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats(key).print_stats(28)
        print(s.getvalue()[:6000])


if __name__ == "__main__":
    main()
