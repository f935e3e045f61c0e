#!/usr/bin/env python3
"""Per-kernel timing of the CLR kernels with CUDA events (GPU box only).

    python tools/microbench.py [--B 8 --C 256 --H 128 --K 2 --iters 20]

Inputs rotate through several buffers whose total size exceeds L2, so no iteration finds its input in cache.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uda_clr_b200 as clr  # noqa: E402
from uda_clr_b200 import _lib, ops, synth  # noqa: E402
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

    def fwd(i):
        check(lib.clr_pool_fwd(ptr(feats[i % a.nbuf]), ptr(y), 0, B, C, HW, K, ptr(ws), ws_bytes, ptr(sums), stream), "fwd")
    med, best = time_it(fwd, a.iters)
    res["pool_fwd"] = dict(us=med, best_us=best, GBs=(F + Lb) / med / 1e3)

    def fwd2(i):
        check(lib.clr_pool_fwd2(ptr(feats[i % a.nbuf]), ptr(y), 0, B, ptr(feats[(i + 1) % a.nbuf]), ptr(y), 0, B,
                                C, HW, K, ptr(ws), ws_bytes, ptr(sums), ptr(sums2), stream), "fwd2")
    med, best = time_it(fwd2, a.iters)
    res["pool_fwd2"] = dict(us=med, best_us=best, GBs=2 * (F + Lb) / med / 1e3)

    def bwd(i):
        check(lib.clr_pool_bwd(ptr(y), 0, B, C, HW, K, ptr(gmat), ptr(sums), 1.0, None, None, 0,
                               ptr(grads[i % a.nbuf]), stream), "bwd")
    med, best = time_it(bwd, a.iters)
    res["pool_bwd"] = dict(us=med, best_us=best, GBs=(F + Lb) / med / 1e3)

    V = torch.randn(K, C, device=dev)
    dots = torch.empty(B, K, H, H, device=dev)

    def dts(i):
        check(lib.clr_pixel_dots(ptr(feats[i % a.nbuf]), B, C, HW, ptr(V), K, ptr(dots), None, stream), "dots")
    med, best = time_it(dts, a.iters)
    res["pixel_dots"] = dict(us=med, best_us=best, GBs=(F + Lb) / med / 1e3)

    # references: a device copy (read+write) and a read-only reduction
    def cp(i):
        grads[i % a.nbuf].copy_(feats[i % a.nbuf])
    med, best = time_it(cp, a.iters)
    res["torch_copy"] = dict(us=med, best_us=best, GBs=2 * F / med / 1e3)

    def rd(i):
        feats[i % a.nbuf].sum()
    med, best = time_it(rd, a.iters)
    res["torch_sum"] = dict(us=med, best_us=best, GBs=F / med / 1e3)

    def fill(i):
        grads[i % a.nbuf].fill_(1.0)
    med, best = time_it(fill, a.iters)
    res["torch_fill"] = dict(us=med, best_us=best, GBs=F / med / 1e3)

    print(json.dumps(dict(shape=[B, C, H, H, K], F_MB=F / 1e6, results=res), indent=1))


if __name__ == "__main__":
    main()
