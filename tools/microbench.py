#!/usr/bin/env python3
"""Per-kernel timing of the CLR kernels with CUDA events (GPU box only).

    python tools/microbench.py [--B 8 --C 256 --H 128 --K 2 --iters 20]

Inputs rotate through several buffers whose total size exceeds L2, so no iteration finds its input in cache.
"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uda_clr_b200 as clr  # noqa: E402
from uda_clr_b200 import _lib, synth  # noqa: E402
from uda_clr_b200._lib import check, ptr  # noqa: E402


def time_it(fn, iters, warmup=3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(evs):
        a.record()
        fn(i)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)  # us
    return ts[len(ts) // 2], ts[0]


def time_batch(fn, iters, warmup=3):
    """Back-to-back launches inside one event pair (amortises the event / launch gap)."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--C", type=int, default=256)
    ap.add_argument("--H", type=int, default=128)
    ap.add_argument("--K", type=int, default=2)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--nbuf", type=int, default=3)
    a = ap.parse_args()
    B, C, H, K = a.B, a.C, a.H, a.K
    HW = H * H
    dev = torch.device("cuda:0")
    lib = _lib.load()
    g = torch.Generator().manual_seed(1)
    y = synth.nested_ellipse_labels(B, K, H, H, g).to(dev)
    feats = [torch.randn(B, C, H, H, device=dev) for _ in range(a.nbuf)]
    grads = [torch.empty(B, C, H, H, device=dev) for _ in range(a.nbuf)]
    F = 4 * B * C * HW
    Lb = 4 * B * K * HW
    stream = torch.cuda.current_stream().cuda_stream
    ws_bytes = lib.clr_pool_ws_bytes(B, C, HW, K) * 2
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    sums = torch.empty(2 * K, C + 1, device=dev)
    sums2 = torch.empty(2 * K, C + 1, device=dev)
    gmat = torch.randn(2 * K, C, device=dev)
    res = {}

    def rec(name, fn, nbytes):
        med, best = time_it(fn, a.iters)
        bat = time_batch(fn, a.iters)
        res[name] = dict(us=round(med, 2), best_us=round(best, 2), batch_us=round(bat, 2),
                         GBs=round(nbytes / med / 1e3, 1), GBs_batch=round(nbytes / bat / 1e3, 1))

    def fwd(i):
        check(lib.clr_pool_fwd(ptr(feats[i % a.nbuf]), ptr(y), 0, B, C, HW, K, ptr(ws), ws_bytes, ptr(sums), stream), "fwd")

    def fwd2(i):
        check(lib.clr_pool_fwd2(ptr(feats[i % a.nbuf]), ptr(y), 0, B, ptr(feats[(i + 1) % a.nbuf]), ptr(y), 0, B,
                                C, HW, K, ptr(ws), ws_bytes, ptr(sums), ptr(sums2), stream), "fwd2")

    lib.clr_set_tunable(b"pool_impl", 1)
    rec("pool_fwd_ldg", fwd, F + Lb)
    rec("pool_fwd2_ldg", fwd2, 2 * (F + Lb))
    lib.clr_set_tunable(b"pool_impl", 2)
    for stages in (2, 3, 4):
        lib.clr_set_tunable(b"pool_stages", stages)
        rec("pool_fwd_tma_s%d" % stages, fwd, F + Lb)
        rec("pool_fwd2_tma_s%d" % stages, fwd2, 2 * (F + Lb))
    lib.clr_set_tunable(b"pool_stages", 0)
    lib.clr_set_tunable(b"pool_impl", 0)

    def bwd(i):
        check(lib.clr_pool_bwd(ptr(y), 0, B, C, HW, K, ptr(gmat), ptr(sums), 1.0, None, None, 0,
                               ptr(grads[i % a.nbuf]), stream), "bwd")
    rec("pool_bwd", bwd, F + Lb)

    doms = (_lib.BwdDom * 2)()
    for d in range(2):
        doms[d].w, doms[d].g, doms[d].sums = ptr(y), ptr(gmat), ptr(sums)
        doms[d].scale, doms[d].fmt, doms[d].B, doms[d].Kx = 1.0, 0, B, 0

    def bwd2(i):
        doms[0].grad = ptr(grads[i % a.nbuf])
        doms[1].grad = ptr(grads[(i + 1) % a.nbuf])
        check(lib.clr_pool_bwd_multi(doms, 2, C, HW, K, stream), "bwd2")
    rec("pool_bwd2", bwd2, 2 * (F + Lb))

    V = torch.randn(K, C, device=dev)
    dots = torch.empty(B, K, H, H, device=dev)

    def dts(i):
        check(lib.clr_pixel_dots(ptr(feats[i % a.nbuf]), B, C, HW, ptr(V), K, ptr(dots), None, stream), "dots")
    rec("pixel_dots", dts, F + Lb)

    # references: a device copy (read+write), a read-only reduction, a write-only fill
    rec("torch_copy", lambda i: grads[i % a.nbuf].copy_(feats[i % a.nbuf]), 2 * F)
    rec("torch_sum", lambda i: feats[i % a.nbuf].sum(), F)
    rec("torch_fill", lambda i: grads[i % a.nbuf].fill_(1.0), F)

    # fused step variants
    bt = synth.make_batch(B=B, C=C, H=H, W=H, K=K, T=8, up=4, seed=1234)
    t = {k: getattr(bt, k).to(dev) for k in ("ys", "oT_before", "preds", "oT", "oT_aug")}
    xs_l = [f.requires_grad_(True) for f in feats]
    for name, kw in (("step_align", dict(retrify=False, use_disc=False, use_cons=False)),
                     ("step_clr3_noretrify", dict(retrify=False, use_disc=True, use_cons=True)),
                     ("step_clr3", dict(retrify=True, use_disc=True, use_cons=True))):
        step = clr.CLRStep(K=K, **kw)
        wt = torch.sigmoid(t["oT_before"])
        masks = 2.0 * torch.ones(B, K, H, H, device=dev)

        def run(i):
            xs, xt = xs_l[i % a.nbuf], xs_l[(i + 1) % a.nbuf]
            xs.grad = None
            xt.grad = None
            if kw["retrify"]:
                out = step(xs, t["ys"], xt, oT_before=t["oT_before"], preds=t["preds"], T=8, oT=t["oT"], oT_aug=t["oT_aug"])
            elif kw["use_cons"]:
                out = step(xs, t["ys"], xt, wt=wt, oT=t["oT"], oT_aug=t["oT_aug"], masks=masks)
            else:
                out = step(xs, t["ys"], xt, wt=wt)
            out.total.backward()
        med, best = time_it(run, a.iters)
        bat = time_batch(run, a.iters)
        res[name] = dict(us=round(med, 2), best_us=round(best, 2), batch_us=round(bat, 2))

    # (the eager-ATen baselines of the same workloads on this GPU live in tests/perf/eager_gpu_baseline.py: they use the
    #  oracle's port, which only tests/ may import)
    print(json.dumps(dict(shape=[B, C, H, H, K], F_MB=F / 1e6, results=res), indent=1))


if __name__ == "__main__":
    main()
