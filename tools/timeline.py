#!/usr/bin/env python3
"""Device-side timeline of ONE live CLR step (GPU box only): which kernel ran when, where the gaps are.

    python tools/timeline.py [--steps 3] [--B 8 --C 256 --H 128 --K 2] [--tunable name=value ...]

Every kernel of the library stamps ``%globaltimer`` into its trace slot (``clr_trace_enable``): earliest CTA start
(before ``griddepcontrol.wait``), earliest return from the wait (= predecessor grid complete), latest CTA exit.
Unlike ncu (serialised, cold caches) or event records (they break programmatic dependent launch) this shows the
step as the benchmark runs it.  Output: one table per traced step, times in microseconds from the step's first stamp.
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--C", type=int, default=256)
    ap.add_argument("--H", type=int, default=128)
    ap.add_argument("--K", type=int, default=2)
    ap.add_argument("--T", type=int, default=8)
    ap.add_argument("--up", type=int, default=4)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--workload", default="clr3", choices=["clr3", "align"])
    ap.add_argument("--tunable", action="append", default=[])
    ap.add_argument("--json", default=None)
    ap.add_argument("--all-ranks", action="store_true", help="under torchrun: every rank writes its own <json>.rank<r>")
    ap.add_argument("--pipelined", type=int, default=0,
                    help="also enqueue N steps back to back under ONE trace window: (last exit - first start) / N is the steady-state step time on the device")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:       # under torchrun: sharded step with the in-kernel exchange; rank 0 prints its own timeline
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        clr.dist.enable_peer()
    lib = _lib.load()
    for kv in a.tunable:
        name, val = kv.split("=")
        _lib.check(lib.clr_set_tunable(name.encode(), int(val)), "clr_set_tunable(%s)" % kv)
    use3 = a.workload == "clr3"
    NSET = 2
    host = [synth.make_batch(B=a.B, C=a.C, H=a.H, W=a.H, K=a.K, T=a.T, up=a.up, seed=1234 + s + 17 * rank, image_res=use3)
            for s in range(NSET)]
    names = ["xs", "ys", "xt", "oT_before"] + (["preds", "oT", "oT_aug"] if use3 else [])
    devb = [{k: getattr(h, k).to(dev) for k in names} for h in host]
    step = clr.CLRStep(K=a.K, retrify=use3, use_disc=use3, use_cons=use3, backprop_aug=False, global_batch=a.B * world)
    plans = []
    for d in devb:
        if use3:
            p = step.plan(d["xs"], d["ys"], d["xt"], oT_before=d["oT_before"], preds=d["preds"], T=a.T,
                          oT=d["oT"], oT_aug=d["oT_aug"], epoch=0.0)
        else:
            p = step.plan(d["xs"], d["ys"], d["xt"], wt=torch.sigmoid(d["oT_before"]))
        plans.append(p)
    for i in range(10):
        plans[i % NSET].run()
    torch.cuda.synchronize()

    n = lib.clr_trace_slots()
    slot_names = [lib.clr_trace_name(i).decode() for i in range(n)]
    buf = (ctypes.c_ulonglong * (4 * n))()
    _lib.check(lib.clr_trace_enable(1), "clr_trace_enable")
    plans[0].run()                       # installs the trace pointer in every translation unit (first launch pays a memcpy)
    _lib.check(lib.clr_trace_read(buf), "clr_trace_read")
    out = []
    for s in range(a.steps):
        # several steps back to back, only the LAST one's stamps survive the min/max (t_first: min -> first step!)
        # so trace exactly one step per read, preceded by an untraced-equivalent warm step in flight
        plans[(s + 1) % NSET].run()
        _lib.check(lib.clr_trace_read(buf), "clr_trace_read")
        rows = []
        for i in range(n):
            t_first, t_ready, t_last, ncta = buf[4 * i], buf[4 * i + 1], buf[4 * i + 2], buf[4 * i + 3]
            if ncta == 0:
                continue
            if t_first == 0:      # per-phase cycle counters (builds with CLR_NVCC_EXTRA=-DCLR_PHASE_PROFILE)
                if rank == 0:
                    print("  %-18s %d cycles" % (slot_names[i], t_last))
                continue
            rows.append((slot_names[i], t_first, t_ready, t_last, ncta))
        t0 = min(r[1] for r in rows)
        rows.sort(key=lambda r: r[2])
        total = (max(r[3] for r in rows) - t0) / 1e3
        out.append({"step": s, "span_us": total,
                    "kernels": [{"name": r[0], "start_us": (r[1] - t0) / 1e3, "ready_us": (r[2] - t0) / 1e3,
                                 "exit_us": (r[3] - t0) / 1e3, "ctas": int(r[4])} for r in rows]})
        if rank != 0:
            continue
        print("step %d (%s, %d GPU%s): kernel, first CTA start, predecessor done, last CTA exit, busy, CTAs [us]"
              % (s, a.workload, world, "s" if world > 1 else ""))
        prev_end = None
        for name, tf, tr, tl, ncta in rows:
            gap = "" if prev_end is None else "  gap_after_prev %+6.2f" % ((tr - prev_end) / 1e3)
            print("  %-18s start %8.2f  ready %8.2f  exit %8.2f  busy %7.2f  ctas %5d%s"
                  % (name, (tf - t0) / 1e3, (tr - t0) / 1e3, (tl - t0) / 1e3, (tl - tr) / 1e3, ncta, gap))
            prev_end = tl
        print("  step span %.2f us" % total)
    if a.pipelined > 0:
        for i in range(a.pipelined):
            plans[i % NSET].run()
        _lib.check(lib.clr_trace_read(buf), "clr_trace_read")
        firsts = [buf[4 * i] for i in range(n) if buf[4 * i + 3] and buf[4 * i]]
        lasts = [buf[4 * i + 2] for i in range(n) if buf[4 * i + 3] and buf[4 * i]]
        if rank == 0:
            per_kernel = {slot_names[i]: (buf[4 * i + 2] - buf[4 * i]) / 1e3 for i in range(n) if buf[4 * i + 3] and buf[4 * i]}
            print("pipelined: %d steps back to back: %.2f us per step (device stamps, first start -> last exit)"
                  % (a.pipelined, (max(lasts) - min(firsts)) / 1e3 / a.pipelined))
    _lib.check(lib.clr_trace_enable(0), "clr_trace_enable")
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if a.json and (rank == 0 or a.all_ranks):
        path = a.json if world == 1 or not a.all_ranks else "%s.rank%d" % (a.json, rank)
        with open(path, "w") as fh:
            json.dump({"config": vars(a), "rank": rank, "world": world, "steps": out}, fh, indent=1)


if __name__ == "__main__":
    main()
