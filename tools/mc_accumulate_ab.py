#!/usr/bin/env python3
"""SURVEY 8(f) rank 1, measured: MC statistics from the staged ``[T*B,K,Hi,Wi]`` logits (one read, ``clr_mc_stats``) against
running statistics over the MC forwards without a staging buffer (``clr_mc_accumulate`` x T/2 + ``clr_mc_finalize``), at the
trainer's shape (B=8, K=2, 512x512, T=8, two passes per forward as in Trainer_prototype_full.py:359-368).  Also times what
the trainer's own staging costs: the slice assignment ``preds_trg[...] = logits`` per forward.  CUDA events, 200 reps."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uda_clr_b200 as clr  # noqa: E402


def timed(fn, reps=200, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    dev = torch.device("cuda:0")
    B, K, Hi, T = 8, 2, 512, 8
    g = torch.Generator(device=dev).manual_seed(0)
    fwd = [2.0 * torch.randn(2 * B, K, Hi, Hi, generator=g, device=dev) for _ in range(T // 2)]      # 4 MC forwards, 2 passes each
    spare = [torch.randn(64 * 1024 * 1024, device=dev) for _ in range(2)]                            # L2 flush between variants
    staged = torch.zeros(T * B, K, Hi, Hi, device=dev)
    acc = clr.MCAccumulator()

    def stage():
        for i, x in enumerate(fwd):
            staged[2 * B * i:2 * B * (i + 1)] = x

    def stats_staged():
        clr.mc_statistics(staged, T, B)

    def accumulate():
        for x in fwd:
            acc.add(x, passes=2)
        acc.finalize()

    stage()
    s_ref, m_ref = clr.mc_statistics(staged, T, B)
    for x in fwd:
        acc.add(x, passes=2)
    s_acc, m_acc = acc.finalize()
    Li = 4 * B * K * Hi * Hi
    out = {"shape": dict(B=B, K=K, Hi=Hi, T=T), "Li_MB": Li / 1e6,
           "max_abs_diff_std": float((s_ref - s_acc).abs().max()), "max_abs_diff_mean": float((m_ref - m_acc).abs().max()),
           "us": {"trainer_staging_copies (4 slice assignments, 2T*Li bytes)": timed(stage),
                  "clr_mc_stats on the staged logits (T*Li + 2Li bytes)": timed(stats_staged),
                  "clr_mc_accumulate x4 + clr_mc_finalize (no staging; ~T*Li + 30Li bytes)": timed(accumulate)},
           "memory_MB": {"staging buffer preds_trg": T * Li / 1e6, "accumulator state (4 maps)": 4 * Li / 1e6,
                         "dead features_trg of the reference (never needed)": 64 * 305 * 128 * 128 * 4 / 1e6}}
    out["verdict"] = ("staged path = staging copies + mc_stats; accumulate path replaces both: compare "
                      "us[staging] + us[mc_stats] with us[accumulate]")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
