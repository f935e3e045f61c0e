#!/usr/bin/env python3
"""The tougher baseline of SURVEY.md 8(d): the reference's EAGER op sequence on the same B200 (the oracle's op-for-op
torch port, run on CUDA tensors) against this library, fwd + bwd, for

  * the whole CLR step (clr3 workload),
  * the drop-in pair gen_prototype + gen_prototype_retrify under autograd,
  * the 8(f) glue ops (seg loss, uncertainty map) -- through autograd and through the raw C ABI,
  * TransNorm (8(f) rank 4) on decoder / ASPP / backbone-like activation shapes (``--only tn`` runs just this part).

    python tests/perf/eager_gpu_baseline.py [--B 8 --C 256 --H 128 --K 2 --iters 20]

Lives under tests/ because it imports the oracle (only tests/, smoke() and bench.py's CPU legs may).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import uda_clr_b200 as clr  # noqa: E402
from microbench import time_batch, time_it  # noqa: E402
from oracle import clr_torch_port as TP  # noqa: E402
from uda_clr_b200 import _lib, synth  # noqa: E402
from uda_clr_b200._lib import check, ptr  # noqa: E402


def transnorm_section(a):
    """TransNorm forward + backward (module under autograd, and the raw C ABI) vs the reference's eager ATen sequence;
    algorithmic bytes = 2 reads + 1 write forward, 4 reads + 1 write backward = 8 x 4BCHW."""
    from uda_clr_b200 import transnorm as TN
    dev = torch.device("cuda:0")
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    res = {}
    for (B, C, H) in [(16, 305, 128), (16, 256, 128), (16, 1280, 32), (16, 24, 128), (16, 32, 256)]:
        xs = [torch.randn(B, C, H, H, device=dev).requires_grad_(True) for _ in range(a.nbuf)]
        gy = torch.randn(B, C, H, H, device=dev)
        m = TN.TransNorm2d(C).to(dev)
        bufs = [torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.zeros(C, device=dev), torch.ones(C, device=dev)]
        w = torch.ones(C, device=dev, requires_grad=True)
        b_ = torch.zeros(C, device=dev, requires_grad=True)

        def ours(i):
            x = xs[i % a.nbuf]
            x.grad = None
            m(x).backward(gy)

        def eager(i):
            x = xs[i % a.nbuf]
            x.grad = None
            TP.trans_norm(x, w, b_, *bufs, True, 0.1, 1e-5).backward(gy)
        y = torch.empty(B, C, H, H, device=dev)
        gx = torch.empty(B, C, H, H, device=dev)
        save = torch.empty(5, C, device=dev)
        gw_, gb_ = torch.empty(C, device=dev), torch.empty(C, device=dev)
        wsb = lib.clr_tn_ws_bytes(C)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)

        def raw(i):
            x = xs[i % a.nbuf]
            check(lib.clr_tn_fwd(ptr(x), B, C, H * H, ptr(w), ptr(b_), ptr(bufs[0]), ptr(bufs[1]), ptr(bufs[2]), ptr(bufs[3]),
                                 0.1, 1e-5, ptr(ws), wsb, ptr(y), ptr(save), stream), "tn fwd")
            check(lib.clr_tn_bwd(ptr(x), ptr(gy), B, C, H * H, ptr(w), ptr(save), 0, ptr(ws), wsb, ptr(gx), ptr(gw_), ptr(gb_),
                                 stream), "tn bwd")
        tag = "transnorm_B%d_C%d_%dx%d" % (B, C, H, H)
        alg = 8 * 4.0 * B * C * H * H
        us_raw = time_batch(raw, a.iters)
        res[tag] = dict(ours_autograd_us=round(time_batch(ours, a.iters), 2), c_abi_us=round(us_raw, 2),
                        eager_aten_us=round(time_batch(eager, max(5, a.iters // 2)), 2),
                        algorithmic_mb=round(alg / 1e6, 1), c_abi_gbs=round(alg / us_raw / 1e3, 1))
        del xs, gy, y, gx
        torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--C", type=int, default=256)
    ap.add_argument("--H", type=int, default=128)
    ap.add_argument("--K", type=int, default=2)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--nbuf", type=int, default=3)
    ap.add_argument("--only", default="", choices=["", "tn"])
    a = ap.parse_args()
    if a.only == "tn":
        print(json.dumps(dict(results=transnorm_section(a)), indent=1))
        return
    B, C, H, K = a.B, a.C, a.H, a.K
    dev = torch.device("cuda:0")
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    res = {}
    bt = synth.make_batch(B=B, C=C, H=H, W=H, K=K, T=8, up=4, seed=1234)
    t = {k: getattr(bt, k).to(dev) for k in ("ys", "oT_before", "preds", "oT", "oT_aug")}
    xs_l = [torch.randn(B, C, H, H, device=dev).requires_grad_(True) for _ in range(a.nbuf)]

    # ---- whole step ----------------------------------------------------------------------------------------------
    port = TP.ClrStepPort(retrify=True, use_disc=True, use_cons=True, backprop_aug=False)

    def eager_step(i):
        xs, xt = xs_l[i % a.nbuf], xs_l[(i + 1) % a.nbuf]
        xs.grad = None
        xt.grad = None
        port.step(xs, t["ys"], xt, t["oT_before"], preds=t["preds"], features=None, T=8, oT=t["oT"], oT_aug=t["oT_aug"], epoch=0.0)
    med, best = time_it(eager_step, max(5, a.iters // 2))
    res["step_clr3_eager_port_on_gpu"] = dict(us=round(med, 2), best_us=round(best, 2))
    step_p = clr.CLRStep(K=K, retrify=True, use_disc=True, use_cons=True)
    plan_p = step_p.plan(xs_l[0].detach(), t["ys"], xs_l[1].detach(), oT_before=t["oT_before"], preds=t["preds"], T=8, oT=t["oT"],
                         oT_aug=t["oT_aug"])
    res["step_clr3_plan_run"] = dict(batch_us=round(time_batch(lambda i: plan_p.run(), 50), 2))

    # ---- drop-in ops ------------------------------------------------------------------------------------------------
    seeds4 = [torch.randn(1, C, 1, 1, device=dev) for _ in range(2 * K)]

    def dropin(fn_proto, fn_retr):
        def run(i):
            xs, xt = xs_l[i % a.nbuf], xs_l[(i + 1) % a.nbuf]
            xs.grad = None
            xt.grad = None
            ps = fn_proto(t["ys"], xs)
            pt = fn_retr(t["oT_before"], xt, t["preds"], None, 8, B)[:2 * K]
            sum((p * s).sum() for p, s in zip(list(ps) + list(pt), seeds4 + seeds4)).backward()
        return run
    med, best = time_it(dropin(clr.gen_prototype, clr.gen_prototype_retrify), a.iters)
    res["dropin_gen_prototype+retrify_ours"] = dict(us=round(med, 2), best_us=round(best, 2))
    med, best = time_it(dropin(TP.gen_prototype, TP.gen_prototype_retrify), max(5, a.iters // 2))
    res["dropin_gen_prototype+retrify_eager_gpu"] = dict(us=round(med, 2), best_us=round(best, 2))

    # ---- 8(f) glue at image resolution --------------------------------------------------------------------------------
    Hi = 4 * H
    oS = [torch.randn(B, K, Hi, Hi, device=dev).requires_grad_(True) for _ in range(a.nbuf)]
    bS = [torch.randn(B, 1, Hi, Hi, device=dev).requires_grad_(True) for _ in range(a.nbuf)]
    tmap = (torch.rand(B, K, Hi, Hi, device=dev) > 0.5).float()
    tbd = torch.rand(B, 1, Hi, Hi, device=dev)

    def rec(name, fn):
        med, best = time_it(fn, a.iters)
        res[name] = dict(us=round(med, 2), best_us=round(best, 2), batch_us=round(time_batch(fn, a.iters), 2))

    def seg(fn):
        def run(i):
            o, b_ = oS[i % a.nbuf], bS[i % a.nbuf]
            o.grad = None
            b_.grad = None
            fn(o, b_, tmap, tbd).backward()
        return run
    rec("seg_loss_fwd_bwd_ours_autograd", seg(clr.seg_loss))
    rec("seg_loss_fwd_bwd_eager_aten", seg(TP.seg_loss))
    segws = torch.empty(lib.clr_seg_loss_ws_bytes(), dtype=torch.uint8, device=dev)
    segout = torch.empty(4, device=dev)
    g1 = torch.empty(B, K, Hi, Hi, device=dev)
    g2 = torch.empty(B, 1, Hi, Hi, device=dev)

    def seg_raw(i):
        o, b_ = oS[i % a.nbuf], bS[i % a.nbuf]
        check(lib.clr_seg_loss_fwd(ptr(o), ptr(tmap), o.numel(), ptr(b_), ptr(tbd), b_.numel(), ptr(segws), segws.numel(),
                                   ptr(segout), stream), "seg fwd")
        check(lib.clr_seg_loss_bwd(ptr(o), ptr(tmap), o.numel(), ptr(b_), ptr(tbd), b_.numel(), None, 1.0, ptr(g1), ptr(g2),
                                   stream), "seg bwd")
    rec("seg_loss_fwd_bwd_c_abi", seg_raw)
    gw = torch.randn(B, K, Hi, Hi, device=dev)

    def ent(fn):
        def run(i):
            o = oS[i % a.nbuf]
            o.grad = None
            fn(o).backward(gw)
        return run
    rec("uncertainty_map_fwd_bwd_ours_autograd", ent(clr.uncertainty_map))
    rec("uncertainty_map_fwd_bwd_eager_aten", ent(TP.uncertainty_map))

    def ent_raw(i):
        o = oS[i % a.nbuf]
        check(lib.clr_entropy_fwd(ptr(o), o.numel(), 1e-7, ptr(g1), stream), "ent fwd")
        check(lib.clr_entropy_bwd(ptr(o), ptr(gw), o.numel(), 1e-7, ptr(g1), stream), "ent bwd")
    rec("uncertainty_map_fwd_bwd_c_abi", ent_raw)
    res.update(transnorm_section(a))
    print(json.dumps(dict(shape=[B, C, H, H, K], results=res), indent=1))


if __name__ == "__main__":
    main()
