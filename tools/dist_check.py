#!/usr/bin/env python3
"""Multi-GPU parity of the sharded CLR step (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

Every rank holds a shard of the global batch; the sharded step (all-reduce of the packed sums) must reproduce the
single-GPU step on the concatenated batch: identical prototypes / losses on every rank, and feature gradients equal
to world_size x the corresponding slice of the single-GPU gradient (DDP then averages them back).
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uda_clr_b200 as clr  # noqa: E402
from uda_clr_b200 import synth  # noqa: E402


def relerr(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for K, C, H, up, T in ((2, 64, 64, 4, 8), (8, 24, 32, 4, 4)):      # K = 8: 2R*8 + 2R = 288 exchange items > 256 threads
        ok = run_case(rank, world, dev, K, C, H, up, T) and ok
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    clr.dist.close_peer()
    dist.destroy_process_group()
    if int(flag.item()) != 0:
        sys.exit(1)
    if rank == 0:
        print("DIST_CHECK_OK world=%d" % world)


def run_case(rank, world, dev, K, C, H, up, T):
    Bg = 4 * world
    b = synth.make_batch(B=Bg, C=C, H=H, W=H, K=K, T=T, up=up, seed=4321)
    full = {k: getattr(b, k).to(dev) for k in ("xs", "ys", "xt", "oT_before", "oT", "oT_aug")}
    # preds is [T*B,...] with t-major order: shard the batch axis inside every MC pass
    preds = b.preds.view(T, Bg, K, H * up, H * up).to(dev)
    lo, hi = clr.dist.shard_bounds(Bg, rank, world)
    mine = {k: v[lo:hi].contiguous() for k, v in full.items()}
    preds_mine = preds[:, lo:hi].reshape(T * (hi - lo), K, H * up, H * up).contiguous()
    ok = True
    from uda_clr_b200 import _lib
    lib = _lib.load()
    # peer_plan: the default ("sched" = 0: schedule 1 below 8 ranks, schedule 2 + split disc finish from 8 on); peer_plan_s1 /
    # peer_plan_s2 force one schedule so that both are checked at every world size
    for variant, mode in (("align", "nccl"), ("clr3", "nccl"), ("align", "peer"), ("clr3", "peer"), ("clr3", "peer_plan"),
                          ("clr3", "peer_plan_s1"), ("clr3", "peer_plan_s2")):
        lib.clr_set_tunable(b"sched", {"peer_plan_s1": 1, "peer_plan_s2": 2}.get(mode, 0))
        lib.clr_set_tunable(b"dfin_split", {"peer_plan_s1": 2, "peer_plan_s2": 1}.get(mode, 0))
        use3 = variant == "clr3"
        # single-GPU reference on the whole batch (every rank computes it redundantly)
        ref = clr.CLRStep(K=K, retrify=use3, use_disc=use3, use_cons=use3, backprop_aug=use3)
        sh = clr.CLRStep(K=K, retrify=use3, use_disc=use3, use_cons=use3, backprop_aug=use3, global_batch=Bg)
        for it in range(2):
            xs_f, xt_f, a_f = (full[k].clone().requires_grad_(True) for k in ("xs", "xt", "oT_aug"))
            kw_f = dict(oT_before=full["oT_before"], preds=preds.reshape(T * Bg, K, H * up, H * up), T=T,
                        oT=full["oT"], oT_aug=a_f) if use3 else dict(wt=torch.sigmoid(full["oT_before"]))
            out_f = ref(xs_f, full["ys"], xt_f, **kw_f)
            out_f.total.backward()
            # nccl: all-reduce of the packed sums between the step's phases; peer: the exchange runs inside the step's
            # own kernels over peer-mapped memory (autograd call path / prebound plan)
            if mode == "nccl":
                clr.dist.enable()
            else:
                clr.dist.enable_peer()
            xs_s, xt_s, a_s = (mine[k].clone().requires_grad_(True) for k in ("xs", "xt", "oT_aug"))
            kw_s = dict(oT_before=mine["oT_before"], preds=preds_mine, T=T, oT=mine["oT"], oT_aug=a_s) if use3 \
                else dict(wt=torch.sigmoid(mine["oT_before"]))
            if mode.startswith("peer_plan"):
                plan = sh.plan(xs_s.detach(), mine["ys"], xt_s.detach(), **{k: (v.detach() if torch.is_tensor(v) else v) for k, v in kw_s.items()})
                plan.run()
                out_s = plan.outputs()
                xs_s.grad, xt_s.grad, a_s.grad = plan.gxs, plan.gxt, plan.g_oT_aug
                timeout_flag = float(plan.losses[7])
            else:
                out_s = sh(xs_s, mine["ys"], xt_s, **kw_s)
                out_s.total.backward()
                timeout_flag = 0.0
            clr.dist.disable()
            if timeout_flag != 0.0:
                print("rank %d: exchange timeout flag set" % rank, flush=True)
                ok = False
            errs = dict(
                total=abs(float(out_s.total) - float(out_f.total)) / abs(float(out_f.total)),
                Ps=relerr(torch.cat(out_s.source_prototypes), torch.cat(out_f.source_prototypes)),
                Pt=relerr(torch.cat(out_s.target_prototypes), torch.cat(out_f.target_prototypes)),
                gxs=relerr(xs_s.grad / world, xs_f.grad[lo:hi]),
                gxt=relerr(xt_s.grad / world, xt_f.grad[lo:hi]))
            if use3:
                errs["g_aug"] = relerr(a_s.grad / world, a_f.grad[lo:hi])
                errs["disc"] = abs(float(out_s.disc) - float(out_f.disc)) / abs(float(out_f.disc))
                errs["aug"] = abs(float(out_s.aug) - float(out_f.aug)) / abs(float(out_f.aug))
            bad = {k: v for k, v in errs.items() if not (v < (1e-5 if k in ("Ps", "Pt") else 1e-4))}
            print("rank %d K=%d %s/%s step %d: %s %s" % (rank, K, variant, mode, it, {k: "%.1e" % v for k, v in errs.items()},
                                                   "FAIL " + str(bad) if bad else "ok"), flush=True)
            ok = ok and not bad
    lib.clr_set_tunable(b"sched", 0)
    lib.clr_set_tunable(b"dfin_split", 0)
    return ok


if __name__ == "__main__":
    main()
