#!/usr/bin/env python3
"""Benchmark of the CLR hot path: ``CLR loss fwd+bwd Mpixels/s`` (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload clr3|align]

A *step* is one pass of the hot path over one synthetic batch per GPU: source pooling (A1, hard labels) +
target retrify (A2: MC statistics + confidence-weighted pooling) + EMA (A4) + alignment / separation losses
(A5) + discriminative hinge (A9) + augmented-consistency BCE (A10), forward AND backward (gradients of both
feature maps written).  One *pixel* is one (b,h,w) location of one domain's feature map: a step processes
``2*B*H*W`` pixels per GPU.

Arms
  ours       the sm_100a kernels through the public API (``CLRStep.plan(...).run()``); inputs resident in HBM.
             ``e2e`` repeats the measurement with pinned HOST inputs copied to the device inside every step and
             the losses read back to the host.
  reference  the reference's own eager-PyTorch CPU implementation of the same step, timed on this box's host
             cores (the op-for-op port in ``oracle/clr_torch_port.py`` -- the reference tree is Python and does
             not travel to the GPU box; ``kind: "port"``).

Rank 0 prints ONE JSON line.  Under torchrun each rank runs its own shard of the global batch (weak scaling:
per-GPU batch fixed) and the only exchange is the all-reduce of the packed class-wise sums.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "clr_fwd_bwd_mpixels_per_s"
UNIT = "Mpixel/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="clr3", choices=["clr3", "align", "dropin"])
    ap.add_argument("--B", type=int, default=8, help="per-GPU batch of each domain")
    ap.add_argument("--C", type=int, default=256)
    ap.add_argument("--H", type=int, default=128)
    ap.add_argument("--K", type=int, default=2)
    ap.add_argument("--T", type=int, default=8)
    ap.add_argument("--up", type=int, default=4, help="image resolution / feature resolution")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block (first two steps vs the eager port on the same GPU)")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip timing the eager port on the same GPU (second baseline)")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    ap.add_argument("--tunable", action="append", default=[], help="library knob name=value (clr_set_tunable)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: in-kernel exchange over peer-mapped memory (default) or NCCL all-reduce between phases")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------- byte model
def algorithmic_bytes(a) -> dict:
    """the ALGORITHMIC traffic of each stage per step and per GPU -- what `roofline` is computed from.
    F = one fp32 pass over one feature map, Lb = one pass over a [B,K,H,W] plane set, Li = the same at image resolution.
      pooling fwd (hard labels)  F + Lb per domain          pooling bwd   F + Lb per domain (write grad + read labels)
      retrify extras             T*Li (read preds) + Li (write std_map) + masks (Lb)
      discriminative fwd         F + Lb (second read of xs); its backward rides in the gradient write (0 bytes)
      consistency fwd            2*Li + Lb
    clr3 at config 1: 4(F+Lb) + (T*Li + Li + Lb) + (F+Lb) + (2*Li+Lb) = 863 MB."""
    B, C, HW, K, T, up = a.B, a.C, a.H * a.H, a.K, a.T, a.up
    F = 4 * B * C * HW
    Lb = 4 * B * K * HW
    Li = 4 * B * K * HW * up * up
    d = {"pool_fwd": 2 * (F + Lb), "pool_bwd": 2 * (F + Lb)}
    if a.workload == "clr3":
        d["mc_stats"] = T * Li + Li + Lb
        d["disc_fwd"] = F + Lb
        d["cons_fwd"] = 2 * Li + Lb
    d["total"] = sum(d.values())
    return d


def implementation_bytes(a) -> dict:
    """What the kernels of THIS implementation move on top of the algorithmic minimum (reported separately, never used
    for a roofline fraction): the full-resolution mean map that clr_mc_stats materialises for the bilinear taps, the
    explicit 2K target weight planes, the coefficient planes of the discriminative term."""
    B, C, HW, K, T, up = a.B, a.C, a.H * a.H, a.K, a.T, a.up
    F = 4 * B * C * HW
    Lb = 4 * B * K * HW
    Li = 4 * B * K * HW * up * up
    d = {"pool_fwd": 2 * F + Lb + (2 * Lb if a.workload == "clr3" else Lb),
         "pool_bwd": 2 * F + Lb + (2 * Lb + Lb if a.workload == "clr3" else Lb)}
    if a.workload == "clr3":
        d["mc_stats"] = T * Li + 2 * Li
        d["retrify_weights"] = Lb + 2 * Li // 4 + 3 * Lb
        d["cons_fwd"] = 2 * Li + Lb
        d["disc_fwd"] = F + Lb + Lb
    d["total"] = sum(d.values())
    return d


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """``nvidia-smi`` clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for (ts, r) in self.rows if self.t_begin is not None and self.t_begin <= ts <= self.t_end + 0.1]
        rows = inside if inside else [r for (_, r) in self.rows]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_inside_timed_region": len(inside), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------- reference arm
def run_reference_cpu(a, steps: int, warmup: int, budget_s: float):
    """The reference's eager CPU path on the host cores, on a bounded sample of the same workload."""
    import torch
    from oracle import clr_torch_port as TP
    from uda_clr_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    use3 = a.workload == "clr3"

    def make(Bs):
        b = synth.make_batch(B=Bs, C=a.C, H=a.H, W=a.H, K=a.K, T=a.T, up=a.up, seed=1234, image_res=use3)
        feats = torch.zeros(a.T * Bs, a.C, a.H, a.H) if use3 else None   # the reference's dead staging buffer
        return b, feats

    first = {}      # results of the FIRST step of a fresh state on the full-size batch (seed 1234 = rank 0's first input set)

    def one(port, b, feats, keep=False):
        xs = b.xs.clone().requires_grad_(True)
        xt = b.xt.clone().requires_grad_(True)
        t0 = time.perf_counter()
        if use3:
            oTa = b.oT_aug.clone().requires_grad_(True)
            res = port.step(xs, b.ys, xt, b.oT_before, preds=b.preds, features=feats, T=a.T, oT=b.oT, oT_aug=oTa, epoch=0.0)
        else:
            res = port.step(xs, b.ys, xt, b.oT_before)
        dt = time.perf_counter() - t0
        if keep and not first:
            first.update({k: float(res[k]) for k in ("intra", "inter", "disc", "aug", "total") if k in res})
            first["Ps"] = torch.cat([p.reshape(1, -1) for p in res["Ps"]])
            first["Pt"] = torch.cat([p.reshape(1, -1) for p in res["Pt"]])
        return dt

    def new_port():
        return TP.ClrStepPort(retrify=use3, use_disc=use3, use_cons=use3, backprop_aug=False)

    # size the sample: one probe step at B=1, then the largest batch <= B whose run fits the budget
    b1, f1 = make(1)
    port = new_port()
    one(port, b1, f1)
    t1 = one(port, b1, f1)
    n_total = steps + warmup
    Bs = int(max(1, min(a.B, budget_s / max(t1 * n_total, 1e-9))))
    if Bs * t1 * n_total > 4 * budget_s:      # even B=1 is too slow for K+W steps: cut the step count
        n_total = max(2, int(4 * budget_s / t1))
        warmup = min(warmup, n_total - 1)
        steps = n_total - warmup
    b, feats = (b1, f1) if Bs == 1 else make(Bs)
    port = new_port()
    seq = [one(port, b, feats, keep=(i == 0 and Bs == a.B)) for i in range(warmup + steps)]
    ts = seq[warmup:]
    med = statistics.median(ts)
    mpix = 2 * Bs * a.H * a.H / med / 1e6
    sample = "%d steps x (B=%d of %d per domain, C=%d, %dx%d, K=%d%s), median of per-step wall times" % (
        steps, Bs, a.B * max(1, a.gpus), a.C, a.H, a.H, a.K, (", T=%d, %dx%d preds" % (a.T, a.H * a.up, a.H * a.up)) if use3 else "")
    if a.gpus > 1:
        sample += "; ONE CPU job on this host timing a %d-sample shard of the %d-sample global batch (throughput in pixels/s does not " \
                  "depend on which shard): compare with the 1-GPU line, not with the %d-GPU aggregate" % (Bs, a.B * a.gpus, a.gpus)
    return dict(value=mpix, unit=UNIT, cores=cores, kind="port", sample=sample, ms_per_step=med * 1e3,
                torch_threads=torch.get_num_threads(), steps=steps, warmup=warmup, first_step=first or None)


# ------------------------------------------------------------------------------------------------- parity / second baseline
TOL = {"loss": 1e-4, "proto": 1e-5, "grad": 1e-4}      # BASELINE.json north_star / SURVEY.md 8(d)


def _rel(x, y):
    return float((x - y).abs().max() / y.abs().max().clamp_min(1e-30))


def run_parity(a, clr, synth, make_plans, host_batch, dev, rank, world, dist_on, use3, names):
    """First two steps of a fresh EMA state through the benched call (``CLRPlan.run``), sharded like the timed steps,
    against ``oracle/clr_torch_port.ClrStepPort`` -- the op-for-op eager restatement of the reference that
    tests/test_oracle_vs_reference.py pins bit for bit on the imported reference -- run on the SAME GPU over the whole
    (concatenated) batch.  The oracle is the checker here, never the thing measured."""
    import torch
    if dist_on:
        import torch.distributed as dist
    K = a.K
    pstep = clr.CLRStep(K=K, retrify=use3, use_disc=use3, use_cons=use3, backprop_aug=False, global_batch=a.B * world)
    pplans = make_plans(pstep)
    snaps, cross = [], {"max_loss_diff": 0.0, "timeout_flag": 0.0}
    for s_ in range(2):
        pl = pplans[s_]
        pl.run()
        torch.cuda.synchronize()
        o = pl.outputs()
        lt = pl.losses.detach().clone()
        if dist_on:
            got = [torch.empty_like(lt) for _ in range(world)]
            dist.all_gather(got, lt)
            cross["max_loss_diff"] = max(cross["max_loss_diff"], max(float((g[:5] - got[0][:5]).abs().max()) for g in got))
            cross["timeout_flag"] = max(cross["timeout_flag"], max(float(g[7]) for g in got))
        else:
            cross["timeout_flag"] = max(cross["timeout_flag"], float(lt[7]))
        snaps.append(dict(losses=lt, Ps=torch.cat([p.reshape(1, -1) for p in o.source_prototypes]).clone(),
                          Pt=torch.cat([p.reshape(1, -1) for p in o.target_prototypes]).clone(),
                          gxs=pl.gxs.clone(), gxt=pl.gxt.clone(),
                          masks=None if o.masks is None else torch.cat(o.masks, 1).clone()))
    del pplans
    if rank != 0:
        return None
    from oracle import clr_torch_port as TP
    port = TP.ClrStepPort(retrify=use3, use_disc=use3, use_cons=use3, backprop_aug=False)
    steps, ok = [], cross["timeout_flag"] == 0.0 and cross["max_loss_diff"] == 0.0
    B = a.B
    for s_ in range(2):
        hb = [host_batch(r, s_) for r in range(world)]
        g = {k: torch.cat([getattr(h, k) for h in hb], 0).to(dev) for k in names if k != "preds"}
        if use3:
            Hi = a.H * a.up
            g["preds"] = torch.cat([h.preds.view(a.T, B, K, Hi, Hi) for h in hb], 1).reshape(a.T * B * world, K, Hi, Hi).to(dev)
        xs, xt = g["xs"].clone().requires_grad_(True), g["xt"].clone().requires_grad_(True)
        if use3:
            res = port.step(xs, g["ys"], xt, g["oT_before"], preds=g["preds"], features=None, T=a.T, oT=g["oT"],
                            oT_aug=g["oT_aug"], epoch=0.0)
        else:
            res = port.step(xs, g["ys"], xt, g["oT_before"])
        sn = snaps[s_]
        keys = ["intra", "inter"] + (["disc", "aug"] if use3 else [])
        idx = {"intra": 0, "inter": 1, "disc": 2, "aug": 3}
        loss_rel = {k: abs(float(sn["losses"][idx[k]]) - float(res[k])) / max(abs(float(res[k])), 1e-30) for k in keys}
        if use3 or True:
            tot_ref = float(res["total"])
            loss_rel["total"] = abs(float(sn["losses"][4]) - tot_ref) / max(abs(tot_ref), 1e-30)
        proto_rel = {"Ps": _rel(sn["Ps"], torch.cat([p.reshape(1, -1) for p in res["Ps"]])),
                     "Pt": _rel(sn["Pt"], torch.cat([p.reshape(1, -1) for p in res["Pt"]]))}
        # this rank's shard of the port's gradients; the sharded step scales its gradients by the world size (DDP averages)
        gs_ref, gt_ref = xs.grad[:B] * world, xt.grad[:B] * world
        bad = (sn["gxs"] - gs_ref).abs() > TOL["grad"] * gs_ref.abs().max()
        grad = {"gxt_rel": _rel(sn["gxt"], gt_ref), "gxs_rel": _rel(sn["gxs"], gs_ref),
                "gxs_pixels_above_tol_frac": float(bad.any(dim=1).float().mean())}
        row = {"loss_rel": loss_rel, "proto_rel": proto_rel, "grad": grad}
        if use3:
            ref_masks = torch.cat(res["masks"], 1)[:B]
            row["mask_mismatch_px"] = int((sn["masks"] != ref_masks).sum())
            row["mask_px"] = int(ref_masks.numel())
        steps.append(row)
        # hinge-kink pixels (A9's gradient flips with the active set) are the only place gxs may exceed the tolerance:
        # tests/test_gpu_step.py proves every such pixel sits on the kink; here their fraction must be vanishing
        ok = ok and all(v < TOL["loss"] for v in loss_rel.values()) and all(v < TOL["proto"] for v in proto_rel.values()) \
            and grad["gxt_rel"] < TOL["grad"] and (grad["gxs_rel"] < TOL["grad"] or grad["gxs_pixels_above_tol_frac"] < 1e-4) \
            and row.get("mask_mismatch_px", 0) == 0
        del xs, xt, g, res
    torch.cuda.empty_cache()
    g0 = snaps[0]
    gpu_first = {"intra": float(g0["losses"][0]), "inter": float(g0["losses"][1]), "disc": float(g0["losses"][2]),
                 "aug": float(g0["losses"][3]), "total": float(g0["losses"][4]), "Ps": g0["Ps"].cpu(), "Pt": g0["Pt"].cpu()}
    return {"ok": bool(ok), "_gpu_first": gpu_first,
            "reference": "oracle/clr_torch_port.ClrStepPort (op-for-op eager restatement of the reference) on the same GPU, "
                                         "global batch %d per domain; this rank's shard of masks / gradients" % (B * world),
            "tolerances": dict(TOL, masks="bit-exact"), "steps": steps, "cross_rank": cross}


def run_eager_gpu(a, devb, use3, n=5):
    """BASELINE.md 4, second baseline: the reference's eager ATen op sequence for the same step on the same GPU."""
    import torch
    from oracle import clr_torch_port as TP
    port = TP.ClrStepPort(retrify=use3, use_disc=use3, use_cons=use3, backprop_aug=False)
    ts = []
    for i in range(n + 2):
        d = devb[i % len(devb)]
        xs, xt = d["xs"].clone().requires_grad_(True), d["xt"].clone().requires_grad_(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if use3:
            port.step(xs, d["ys"], xt, d["oT_before"], preds=d["preds"], features=None, T=a.T, oT=d["oT"], oT_aug=d["oT_aug"], epoch=0.0)
        else:
            port.step(xs, d["ys"], xt, d["oT_before"])
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts[2:])
    torch.cuda.empty_cache()
    return {"value": 2 * a.B * a.H * a.H / med / 1e6, "unit": UNIT, "ms_per_step": med * 1e3, "steps": n,
            "what": "oracle/clr_torch_port.ClrStepPort (the reference's eager op sequence, without its dead 1.28 GB mean(features)) "
                    "on this GPU, wall clock around synchronised steps"}


# ------------------------------------------------------------------------------------------------- drop-in workload
def trainer_protocol_step(ops, st, d, T, decay=0.9, pro_weight=0.1):
    """The CLR block exactly as the shipped trainer drives it (Trainer_prototype_full.py:330-449, ``retrify_pesudo``
    branch): two calls into the prototype ops, the inline EMA (:335-355, :378-398), the inline MSE alignment /
    separation losses (:428-444) and ``backward()`` of ``pro_weight * intra`` (:463-468).  ``ops`` is whatever module
    provides ``gen_prototype`` / ``gen_prototype_retrify`` -- this package (the patched trainer) or the eager port."""
    import torch
    mse = torch.nn.MSELoss()
    xs = d["xs"].detach().requires_grad_(True)
    xt = d["xt"].detach().requires_grad_(True)
    K = d["ys"].shape[1]
    cur_s = ops.gen_prototype(d["ys"], xs)                                                   # :332-334
    out = ops.gen_prototype_retrify(d["oT_before"], xt, d["preds"], None, T, xt.shape[0])    # :370-373
    cur_t = out[:2 * K]
    P = {}
    for dom, cur in (("s", cur_s), ("t", cur_t)):
        if st.get(dom) is None:                                                             # First_src / First (:32-33)
            P[dom] = list(cur)
        else:
            P[dom] = [(1 - decay) * old + decay * c for old, c in zip(st[dom], cur)]
        st[dom] = [p.detach() for p in P[dom]]                                              # :341-344, :384-387
    intra = sum(mse(ps, pt) for ps, pt in zip(P["s"], P["t"]))                              # :428-441
    inter = sum(mse(P["s"][k], P["s"][K + k]) for k in reversed(range(K)))                  # :443-444 (logged only)
    loss = pro_weight * intra                                                               # :463-468
    loss.backward()
    return loss.detach(), inter.detach(), xs.grad, xt.grad


def bench_dropin(a, dev, lib):
    """``--workload dropin``: what a reference user gets from ``patch_reference()`` with ZERO trainer changes -- the
    trainer's inline torch code around the drop-in ops -- next to the same protocol on the eager ops, same GPU."""
    import torch
    import uda_clr_b200 as clr
    from oracle import clr_torch_port as TP      # second arm of the comparison (the eager reference ops), not the product
    from uda_clr_b200 import synth
    names = ["xs", "ys", "xt", "oT_before", "preds"]
    hb = [synth.make_batch(B=a.B, C=a.C, H=a.H, W=a.H, K=a.K, T=a.T, up=a.up, seed=1234 + s_) for s_ in range(2)]
    devb = [{k: getattr(h, k).to(dev) for k in names} for h in hb]

    def run(ops, steps, warm):
        st = {}
        for i in range(warm):
            trainer_protocol_step(ops, st, devb[i % 2], a.T)
        torch.cuda.synchronize()
        l0 = lib.clr_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            res = trainer_protocol_step(ops, st, devb[i % 2], a.T)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, res, lib.clr_launch_count() - l0

    class _NullOps:
        """The protocol's floor: ops that cost nothing on the device (one tiny autograd node each), so what remains is
        the trainer's own inline torch code -- 8 EMA expressions, 6 MSELoss, their autograd graph and backward."""
        _cache = {}

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, protos):
                ctx.shape = x.shape
                return tuple(p.view_as(p) for p in protos)

            @staticmethod
            def backward(ctx, *g):
                return None, None

        @classmethod
        def _protos(cls, x, K):
            key = (x.shape[1], K)
            if key not in cls._cache:
                cls._cache[key] = [torch.randn(1, x.shape[1], 1, 1, device=x.device) for _ in range(2 * K)]
            return cls._Fn.apply(x, cls._cache[key])

        @classmethod
        def gen_prototype(cls, pred, feat):
            return cls._protos(feat, pred.shape[1])

        @classmethod
        def gen_prototype_retrify(cls, o, x, p, f, T, s):
            return cls._protos(x, o.shape[1]) + (None, None, None)

    W_ = max(a.warmup, 3)
    ms, res, launches = run(clr, a.steps, W_)
    ms_ref, res_ref, _ = run(TP, max(3, min(a.steps, 20)), 3)
    ms_null, _, _ = run(_NullOps, a.steps, W_)
    # parity of the two protocols after the same number of EMA steps is covered by tests/test_gpu_integration.py; here
    # the first steps of fresh states are compared
    st1, st2 = {}, {}
    r1 = trainer_protocol_step(clr, st1, devb[0], a.T)
    r2 = trainer_protocol_step(TP, st2, devb[0], a.T)
    torch.cuda.synchronize()
    parity = {"loss_rel": abs(float(r1[0]) - float(r2[0])) / abs(float(r2[0])), "gxs_rel": _rel(r1[2], r2[2]), "gxt_rel": _rel(r1[3], r2[3])}
    parity["ok"] = parity["loss_rel"] < TOL["loss"] and parity["gxs_rel"] < TOL["grad"] and parity["gxt_rel"] < TOL["grad"]
    px = 2 * a.B * a.H * a.H
    line = {"metric": METRIC, "value": px / (ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": W_,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(a, 1), "gpu_launches": int(launches),
            "gpu_eager_baseline": {"value": px / (ms_ref * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_ref,
                                   "what": "the same protocol on the eager port ops, same GPU"},
            "inline_torch_floor": {"ms_per_step": ms_null,
                                   "what": "the same protocol with no-op prototype ops: the trainer's inline EMA / MSE expressions, their "
                                           "autograd graph and backward alone (host-bound eager torch; not replaceable without touching the trainer)"},
            "ops_share_ms": ms - ms_null,
            "parity": parity, "losses": {"total": float(res[0]), "inter": float(res[1])}}
    print(json.dumps(line))
    return 0


def config_dict(a, n_gpus):
    wl = {"clr3": "clr3: A1 hard source + A2 retrify target (MC stats T=%d) + A4 EMA + A5 align + A9 hinge + A10 "
                  "consistency, fwd+bwd" % a.T,
          "align": "align: A1 hard source + A1 soft target + A4 + A5, fwd+bwd (Trainer_prototype_full.py:330-449)",
          "dropin": "dropin: the trainer's own protocol (Trainer_prototype_full.py:330-449: gen_prototype + gen_prototype_retrify "
                    "drop-in ops, inline EMA / MSE in torch, loss.backward()) with the patched ops"}[a.workload]
    return {"workload": wl,
            "per_gpu_batch": a.B, "global_batch": a.B * n_gpus, "channels": a.C, "feature_hw": [a.H, a.H],
            "image_hw": [a.H * a.up, a.H * a.up], "classes": a.K, "mc_passes": a.T,
            "parallelism": "dp%d (batch-sharded; packed class sums exchanged %s)" % (
                n_gpus, "inside the step's kernels over NVLink peer memory" if (a.exchange == "peer" and n_gpus > 1) else
                ("by NCCL all-reduce" if n_gpus > 1 else "-")),
            "l2_policy": "inputs larger than L2: 2 rotating input sets, each step streams > 700 MB vs 126 MB L2",
            "baseline_config": "BASELINE.json configs[0] shape (B=8, 256ch, 128x128, K=2) = the per-GPU CLR workload of configs[1]"}


# ------------------------------------------------------------------------------------------------- main
def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        if rank != 0:
            return 0
        r = run_reference_cpu(a, a.steps, a.warmup, budget_s=max(a.cpu_budget_s, 60.0))
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus,
                "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(a, a.gpus),
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                 "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import uda_clr_b200 as clr
    from uda_clr_b200 import _lib, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
        if a.exchange == "peer":
            clr.dist.enable_peer()
        else:
            clr.dist.enable()
    lib = _lib.load()
    for kv in a.tunable:
        name, val = kv.split("=")
        _lib.check(lib.clr_set_tunable(name.encode(), int(val)), "clr_set_tunable(%s)" % kv)

    if a.workload == "dropin":
        return bench_dropin(a, dev, lib)

    use3 = a.workload == "clr3"
    NSET = 2
    names = ["xs", "ys", "xt", "oT_before"] + (["preds", "oT", "oT_aug"] if use3 else [])

    def host_batch(r, s_):
        return synth.make_batch(B=a.B, C=a.C, H=a.H, W=a.H, K=a.K, T=a.T, up=a.up, seed=1234 + 17 * r + s_, image_res=use3)

    host = [host_batch(rank, s_) for s_ in range(NSET)]
    devb = [{k: getattr(h, k).to(dev) for k in names} for h in host]

    def make_plans(step_obj):
        out = []
        for d in devb:
            if use3:
                out.append(step_obj.plan(d["xs"], d["ys"], d["xt"], oT_before=d["oT_before"], preds=d["preds"], T=a.T,
                                         oT=d["oT"], oT_aug=d["oT_aug"], epoch=0.0))
            else:
                out.append(step_obj.plan(d["xs"], d["ys"], d["xt"], wt=torch.sigmoid(d["oT_before"])))
        return out

    def barrier():
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity (outside the timed region): the first two steps of a FRESH state, sharded exactly like the timed
    #      steps, against the op-for-op eager port of the reference on the SAME GPU (rank 0; at N > 1 on the concatenated
    #      global batch), plus the cross-rank agreement of the losses and the device-side time-out flag ---------------
    parity = None
    if not a.no_parity:
        parity = run_parity(a, clr, synth, make_plans, host_batch, dev, rank, world, dist_on, use3, names)
        barrier()

    step = clr.CLRStep(K=a.K, retrify=use3, use_disc=use3, use_cons=use3, backprop_aug=False, global_batch=a.B * world)
    plans = make_plans(step)

    # ---- warm-up -------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    W_ = max(a.warmup, 3)
    for i in range(W_):
        plans[i % NSET].run()
    barrier()

    # ---- timed region: exactly K steps between two CUDA events on the launching stream; NOTHING else is enqueued
    #      inside it unless K >= 64, where the dominant kernels are bracketed on ~16 evenly spaced steps (an event record
    #      between two kernels defeats programmatic dependent launch, ~2.5 us per bracket: < 0.03 % of such a region) --
    K = a.steps
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    inside = K >= 64
    ev_stride = max(1, K // 16) if inside else 1
    n_samples = len(range(0, K, ev_stride)) if inside else 16
    pool_ev = [(_lib.Event(), _lib.Event()) for _ in range(n_samples)]
    bwd_ev = [(_lib.Event(), _lib.Event()) for _ in range(n_samples)]

    def bracketed(p, j):
        p.set_events(pool_ev[j][0], pool_ev[j][1], bwd_ev[j][0], bwd_ev[j][1])
        p.run()
        p.set_events()

    launches0 = lib.clr_launch_count()
    barrier()
    sampler.mark_begin()
    ev0.record()
    if inside:
        for i in range(K):
            if i % ev_stride == 0:
                bracketed(plans[i % NSET], i // ev_stride)
            else:
                plans[i % NSET].run()
    else:
        for i in range(K):
            plans[i % NSET].run()
    ev1.record()
    launches = lib.clr_launch_count() - launches0
    if not inside:
        # same loop, same rotating inputs, continuing straight on: the kernel-time samples of a short run
        for j in range(n_samples):
            bracketed(plans[(K + j) % NSET], j)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist_on:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / K
    pool_us = statistics.mean(b.elapsed_us(e) for b, e in pool_ev)
    bwd_us = statistics.mean(b.elapsed_us(e) for b, e in bwd_ev)
    value = 2 * a.B * a.H * a.H * world / (ms_per_step * 1e-3) / 1e6
    lt = plans[(K - 1) % NSET].losses.detach().clone()
    flags = [lt]
    if dist_on:
        flags = [torch.empty_like(lt) for _ in range(world)]
        dist.all_gather(flags, lt)
    losses = lt.cpu().tolist()
    timed_flag = max(float(f[7]) for f in flags)
    timed_rank_diff = max(float((f[:5] - flags[0][:5]).abs().max()) for f in flags)

    # ---- supplementary (outside the timed region): per-kernel busy time of a live step from the library's own
    #      device-side %globaltimer stamps -- no event records, programmatic dependent launch intact ----------------
    device_trace = None
    N_TRACE = 6                      # every rank runs the same number of steps (the in-kernel exchange is collective)
    if rank == 0:
        try:
            import ctypes
            n_slots = lib.clr_trace_slots()
            tbuf = (ctypes.c_ulonglong * (4 * n_slots))()
            _lib.check(lib.clr_trace_enable(1), "clr_trace_enable")
            acc = {}
            for i in range(N_TRACE):
                plans[i % NSET].run()
                _lib.check(lib.clr_trace_read(tbuf), "clr_trace_read")
                if i == 0:
                    continue         # the first traced step installs the trace pointers (one memcpy per kernel type)
                rows = [(lib.clr_trace_name(j).decode(), tbuf[4 * j], tbuf[4 * j + 1], tbuf[4 * j + 2]) for j in range(n_slots)
                        if tbuf[4 * j + 3] and tbuf[4 * j]]
                t0 = min(r[1] for r in rows)
                acc.setdefault("step_span", []).append((max(r[3] for r in rows) - t0) / 1e3)
                for name, tf, tr, tl in rows:
                    acc.setdefault(name, []).append((tl - tr) / 1e3)
            _lib.check(lib.clr_trace_enable(0), "clr_trace_enable")
            device_trace = {k: round(statistics.median(v), 2) for k, v in acc.items()}
        except Exception as e:       # profiling aid only
            device_trace = {"error": str(e)[:120]}
    else:
        for i in range(N_TRACE):
            plans[i % NSET].run()
            torch.cuda.synchronize()
    barrier()

    # ---- e2e: pinned host inputs -> device -> step -> losses back to the host, every step ----------
    e2e = None
    if not a.no_e2e:
        pinned = [{k: getattr(h, k).pin_memory() for k in names} for h in host]
        h2d = sum(v.numel() * 4 for v in pinned[0].values())
        out_host = torch.empty(8, dtype=torch.float32).pin_memory()
        n_e2e = max(3, min(K, 10))

        # Two device input sets, so the copy of step i+1 (copy stream) runs under the compute of step i (main stream): every
        # step still pays its own H2D copy of all inputs and the D2H read of its losses inside the timed region; the step
        # is PCIe-bound either way (438 MB per step), the overlap hides the 0.18 ms of compute.
        main_st, copy_st = torch.cuda.current_stream(), torch.cuda.Stream()
        copied = [torch.cuda.Event() for _ in range(NSET)]
        consumed = [torch.cuda.Event() for _ in range(NSET)]

        def e2e_copy(i):
            src, dst = pinned[i % NSET], devb[i % NSET]
            with torch.cuda.stream(copy_st):
                copy_st.wait_event(consumed[i % NSET])          # the step that last used this set is done with it
                for k in names:
                    dst[k].copy_(src[k], non_blocking=True)
                copied[i % NSET].record(copy_st)

        def e2e_compute(i):
            main_st.wait_event(copied[i % NSET])
            p = plans[i % NSET]
            p.run()
            out_host.copy_(p.losses, non_blocking=True)
            consumed[i % NSET].record(main_st)

        def e2e_loop(n):
            e2e_copy(0)
            for i in range(n):
                if i + 1 < n:
                    e2e_copy(i + 1)
                e2e_compute(i)

        for ev in consumed:
            ev.record(main_st)
        e2e_loop(2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        copy_st.wait_event(e0)                                   # no copy of the timed steps starts before the region does
        e2e_loop(n_e2e)
        e1.record()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if dist_on:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ms_e2e = float(te.item()) / n_e2e
        e2e = {"value": 2 * a.B * a.H * a.H * world / (ms_e2e * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32, "ms_per_step": ms_e2e, "steps": n_e2e}

    if rank != 0:
        if dist_on:
            clr.dist.close_peer()
            dist.destroy_process_group()
        return 0

    # ---- second baseline: the reference's eager op sequence on THIS GPU (op-for-op port), outside
    #      the timed region; N = 1 only ------------------------------------------------------------------------
    gpu_eager = None
    if world == 1 and not a.no_gpu_eager:
        gpu_eager = run_eager_gpu(a, devb, use3)

    # ---- roofline of the dominant kernel (the two-domain pooling launch), measured live above ------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    ab = algorithmic_bytes(a)
    ib = implementation_bytes(a)
    # schedule 2 (clr3) pools the two maps in two launches (source first): the bracketed launch is the SOURCE pooling, F + Lb
    sched = int(lib.clr_step_schedule(plans[0]._ref))
    pool_launch_bytes = ab["pool_fwd"] // 2 if sched == 2 else ab["pool_fwd"]
    achieved = pool_launch_bytes / (pool_us * 1e-6) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.isfile(tpath):
        try:
            traffic = json.load(open(tpath)).get("pool_fwd_bytes_per_launch")
        except Exception:
            traffic = None
    if traffic is not None and sched == 2:
        traffic = traffic / 2
    roofline = {"bound": "hbm", "kernel": "pool_fwd_ldg_kernel<%d,4> (%s)" % (2 * a.K, "source pooling launch; the target map is pooled by a second launch "
                                                                                 "of the same kernel" if sched == 2 else "source + target pooling in one launch"),
                "schedule": sched,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": pool_launch_bytes, "kernel_us": pool_us,
                "kernel_us_device_trace": (device_trace or {}).get("pool_fwd"), "device_trace_us": device_trace,
                "kernel_samples": n_samples,
                "kernel_sampling": ("every %d-th step inside the timed region" % ev_stride) if inside else
                                   "%d extra steps right after the timed region (K < 64: nothing but the K steps is enqueued inside it)" % n_samples,
                "bwd_kernel": {"kernel": "pool_bwd_kernel (both gradient maps, one launch)", "kernel_us": bwd_us,
                               "algorithmic_bytes_per_launch": ab["pool_bwd"],
                               "achieved": ab["pool_bwd"] / (bwd_us * 1e-6) / 1e9,
                               "frac": ab["pool_bwd"] / (bwd_us * 1e-6) / 1e9 / peak},
                "step": {"algorithmic_bytes": ab["total"], "algorithmic_bytes_by_stage": ab,
                         "achieved": ab["total"] / (ms_per_step * 1e-3) / 1e9,
                         "frac": ab["total"] / (ms_per_step * 1e-3) / 1e9 / peak,
                         "frac_of_nominal_8TBs": ab["total"] / (ms_per_step * 1e-3) / 1e9 / 8000.0,
                         "implementation_bytes": ib["total"], "implementation_bytes_by_stage": ib}}

    cpu_baseline = None
    gpu_first = parity.pop("_gpu_first", None) if parity else None
    if not a.no_cpu_baseline and world == 1:
        r = run_reference_cpu(a, steps=3, warmup=1, budget_s=a.cpu_budget_s)
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                        "ms_per_step": r["ms_per_step"]}
        # the same first step of a fresh state on the CPU (the reference's own device): cross-device, so the uncertainty masks may
        # differ at knife-edge pixels (the reference's mask depends on the device it runs on) -- losses / prototypes at tolerance
        cf = r.get("first_step")
        if cf and gpu_first and parity is not None:
            keys = [k for k in ("intra", "inter", "disc", "aug", "total") if k in cf]
            rel = {k: abs(gpu_first[k] - cf[k]) / max(abs(cf[k]), 1e-30) for k in keys}
            prot = {k: _rel(gpu_first[k], cf[k]) for k in ("Ps", "Pt")}
            parity["vs_cpu_port_step1"] = {"loss_rel": rel, "proto_rel": prot,
                                           "ok": bool(all(v < TOL["loss"] for v in rel.values()) and all(v < TOL["proto"] for v in prot.values()))}
            parity["ok"] = bool(parity["ok"] and parity["vs_cpu_port_step1"]["ok"])

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(a, world), "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "gpu_eager_baseline": gpu_eager, "parity": parity,
            "timed_region_check": {"exchange_timeout_flag": timed_flag, "max_cross_rank_loss_diff": timed_rank_diff},
            "losses": {"intra": losses[0], "inter": losses[1], "disc": losses[2], "aug": losses[3], "total": losses[4]}}
    print(json.dumps(line))
    if dist_on:
        clr.dist.close_peer()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
