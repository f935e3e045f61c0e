#!/usr/bin/env python3
"""CLR op sweep (BASELINE.json configs[4]): C = 256/305/512, H = W = 64..256, K = 2..8, B = 8..64 on one GPU.

    python tests/perf/sweep.py [--quick] [--json out.json]

(Lives under tests/ because it uses the oracle's eager port as the checker; nothing outside tests/, smoke() and
bench.py's CPU legs may import oracle/.)

Per point: the fused step (clr3 workload) is timed with CUDA events over back-to-back steps (two rotating input sets),
checked for finite outputs and -- where the eager-PyTorch port of the reference fits in memory -- compared with that
port run on the same GPU (losses to 1e-4, prototypes to 1e-5).  Prints one line per point and a JSON summary.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import uda_clr_b200 as clr  # noqa: E402
from oracle import clr_torch_port as TP  # noqa: E402  (checker only)
from uda_clr_b200 import synth  # noqa: E402


def bytes_model(B, C, HW, K, T, up):
    """SURVEY.md 8(d) ALGORITHMIC bytes of the clr3 step (the same model as bench.py:algorithmic_bytes): pooling fwd + bwd of
    both maps 4(F+Lb), retrify extras T*Li + Li + Lb, discriminative re-read F + Lb, consistency 2Li + Lb."""
    F, Lb = 4 * B * C * HW, 4 * B * K * HW
    Li = Lb * up * up
    return 4 * (F + Lb) + (T * Li + Li + Lb) + (F + Lb) + (2 * Li + Lb)


def run_point(B, C, H, K, T, up, steps, check):
    dev = torch.device("cuda", 0)
    sets = []
    for s in range(2):
        b = synth.make_batch(B=B, C=C, H=H, W=H, K=K, T=T, up=up, seed=1234 + s, image_res=True)
        sets.append({k: getattr(b, k).to(dev) for k in ("xs", "ys", "xt", "oT_before", "preds", "oT", "oT_aug")})
        del b
    step = clr.CLRStep(K=K, retrify=True, use_disc=True, use_cons=True, backprop_aug=False, global_batch=B)
    plans = [step.plan(d["xs"], d["ys"], d["xt"], oT_before=d["oT_before"], preds=d["preds"], T=T, oT=d["oT"],
                       oT_aug=d["oT_aug"], epoch=0.0) for d in sets]
    for i in range(4):
        plans[i % 2].run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        plans[i % 2].run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    p = plans[(steps - 1) % 2]
    finite = bool(torch.isfinite(p.losses[:5]).all() and torch.isfinite(p.gxs).all() and torch.isfinite(p.gxt).all())
    res = dict(B=B, C=C, H=H, K=K, T=T, ms_per_step=ms, mpixel_s=2 * B * H * H / ms / 1e3,
               gbs=bytes_model(B, C, H * H, K, T, up) / ms / 1e6, finite=finite, parity=None)
    if check:
        # same two-step protocol on the eager port (reference op sequence) on this GPU, fresh state on both sides
        d = sets[0]
        st2 = clr.CLRStep(K=K, retrify=True, use_disc=True, use_cons=True, backprop_aug=False, global_batch=B)
        pl = st2.plan(d["xs"], d["ys"], d["xt"], oT_before=d["oT_before"], preds=d["preds"], T=T, oT=d["oT"],
                      oT_aug=d["oT_aug"], epoch=0.0)
        port = TP.ClrStepPort(retrify=True, use_disc=True, use_cons=True, backprop_aug=False)
        for _ in range(2):
            pl.run()
            xs = d["xs"].clone().requires_grad_(True)
            xt = d["xt"].clone().requires_grad_(True)
            r = port.step(xs, d["ys"], xt, d["oT_before"], preds=d["preds"], features=None, T=T, oT=d["oT"],
                          oT_aug=d["oT_aug"], epoch=0.0)
        torch.cuda.synchronize()
        o = pl.outputs()
        # gxs carries A9's direct term: pixels whose hinge argument sits within float noise of the kink may legitimately
        # differ (tests/test_gpu_step.py proves they sit on the kink): their fraction must be vanishing
        bad = (pl.gxs - xs.grad).abs() > 1e-4 * xs.grad.abs().max()
        errs = dict(total=abs(float(o.total) - float(r["total"])) / abs(float(r["total"])),
                    disc=abs(float(o.disc) - float(r["disc"])) / max(abs(float(r["disc"])), 1e-30),
                    aug=abs(float(o.aug) - float(r["aug"])) / max(abs(float(r["aug"])), 1e-30),
                    Ps=float((torch.cat(o.source_prototypes) - torch.cat(r["Ps"])).abs().max() / torch.cat(r["Ps"]).abs().max()),
                    Pt=float((torch.cat(o.target_prototypes) - torch.cat(r["Pt"])).abs().max() / torch.cat(r["Pt"]).abs().max()),
                    gxt=float((pl.gxt - xt.grad).abs().max() / xt.grad.abs().max()),
                    gxs=float((pl.gxs - xs.grad).abs().max() / xs.grad.abs().max()),
                    gxs_px_above_tol=float(bad.any(dim=1).float().mean()))
        mask_mismatch = int((torch.cat(o.masks, 1) != torch.cat(r["masks"], 1)).sum())
        res["parity"] = {k: float("%.2e" % v) for k, v in errs.items()}
        res["parity"]["mask_mismatch_px"] = mask_mismatch
        res["parity"]["timeout_flag"] = float(pl.losses[7])
        res["parity_ok"] = bool(errs["total"] < 1e-4 and errs["disc"] < 1e-4 and errs["aug"] < 1e-4 and errs["Ps"] < 1e-5 and
                                errs["Pt"] < 1e-5 and errs["gxt"] < 1e-4 and (errs["gxs"] < 1e-4 or errs["gxs_px_above_tol"] < 1e-4) and
                                mask_mismatch == 0 and float(pl.losses[7]) == 0.0)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    pts = [(8, 256, 128, 2), (8, 305, 128, 2), (8, 512, 128, 2), (8, 256, 64, 2), (16, 256, 128, 2), (32, 256, 128, 2),
           (8, 256, 128, 4), (8, 256, 128, 8), (8, 256, 256, 2), (4, 305, 256, 2), (16, 512, 64, 4), (64, 256, 64, 2),
           (64, 256, 128, 2), (16, 512, 256, 2), (64, 512, 256, 2)]
    if a.quick:
        pts = pts[:5]
    out = []
    for (B, C, H, K) in pts:
        T, up = 8, 4
        feat_gb = 4 * B * C * H * H / 1e9
        check = feat_gb <= 0.6          # the eager port materialises ~8 feature-map-sized temporaries
        steps = 200 if feat_gb < 0.3 else (60 if feat_gb < 1.5 else (20 if feat_gb < 4 else 4))
        try:
            r = run_point(B, C, H, K, T, up, steps, check)
        except Exception as e:          # keep sweeping; report the failure
            r = dict(B=B, C=C, H=H, K=K, error=str(e)[:200])
        out.append(r)
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
    if a.json:
        with open(a.json, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
