#!/usr/bin/env python3
"""DDP contract of the sharded CUDA step (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/tools/ddp_check.py

A 1x1 convolution wrapped in ``DistributedDataParallel`` produces the decoder features of this rank's shard of both
domains; the fused CLR step (clr3: retrify target, EMA, alignment, discriminative hinge, consistency) runs on them with
the packed sums exchanged across ranks (in-kernel exchange over peer memory, and NCCL all-reduce); ``loss.backward()``
lets DDP average the convolution's gradients.  They must equal the gradients of a SINGLE process that runs the eager
port of the reference (oracle/clr_torch_port.ClrStepPort) on the concatenated batch -- SURVEY.md 8(e)'s
``(1/G) * sum_ranks (G * local) = full`` under real DDP.  Two steps (first-step copy, then one EMA step).
"""
import os
import sys

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import uda_clr_b200 as clr  # noqa: E402
from oracle import clr_torch_port as TP  # noqa: E402  (checker)
from uda_clr_b200 import synth  # noqa: E402


def relerr(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def make_conv(cin, c, dev):
    torch.manual_seed(11)
    return torch.nn.Conv2d(cin, c, 1).to(dev)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    K, Cin, C, H, up, T = 2, 48, 64, 64, 4, 8
    Bg = 2 * world
    ok = True
    for mode in ("peer", "nccl"):
        conv = DDP(make_conv(Cin, C, dev), device_ids=[local])
        ref_conv = make_conv(Cin, C, dev)
        step = clr.CLRStep(K=K, retrify=True, use_disc=True, use_cons=True, backprop_aug=False, global_batch=Bg)
        port = TP.ClrStepPort(retrify=True, use_disc=True, use_cons=True, backprop_aug=False)
        for it in range(2):
            b = synth.make_batch(B=Bg, C=Cin, H=H, W=H, K=K, T=T, up=up, seed=900 + it)     # xs / xt here are the conv INPUTS
            full = {k: getattr(b, k).to(dev) for k in ("xs", "ys", "xt", "oT_before", "oT", "oT_aug")}
            preds = b.preds.view(T, Bg, K, H * up, H * up).to(dev)
            lo, hi = clr.dist.shard_bounds(Bg, rank, world)
            mine = {k: v[lo:hi].contiguous() for k, v in full.items()}
            preds_mine = preds[:, lo:hi].reshape(T * (hi - lo), K, H * up, H * up).contiguous()
            # ---- sharded: DDP conv -> fused CUDA step -> backward (DDP averages)
            conv.zero_grad(set_to_none=True)
            if mode == "peer":
                clr.dist.enable_peer()
            else:
                clr.dist.enable()
            fs, ft = conv(mine["xs"]), conv(mine["xt"])
            out = step(fs, mine["ys"], ft, oT_before=mine["oT_before"], preds=preds_mine, T=T, oT=mine["oT"],
                       oT_aug=mine["oT_aug"], epoch=0.0)
            out.total.backward()
            torch.cuda.synchronize()
            flag = float(out.error)
            clr.dist.disable()
            # ---- single process, whole batch, eager port (every rank computes it redundantly)
            ref_conv.zero_grad(set_to_none=True)
            rs, rt = ref_conv(full["xs"]), ref_conv(full["xt"])
            res = port.step(rs, full["ys"], rt, full["oT_before"], preds=preds.reshape(T * Bg, K, H * up, H * up), features=None,
                            T=T, oT=full["oT"], oT_aug=full["oT_aug"], epoch=0.0)
            errs = dict(total=abs(float(out.total) - float(res["total"])) / abs(float(res["total"])),
                        gW=relerr(conv.module.weight.grad, ref_conv.weight.grad),
                        # (on the first step the loss is invariant to a common shift of all features -- the bias gradient is
                        #  pure rounding noise around 0 -- so its error is measured on the scale of the weight gradient)
                        gb=float((conv.module.bias.grad - ref_conv.bias.grad).abs().max()
                                 / torch.maximum(ref_conv.bias.grad.abs().max(), ref_conv.weight.grad.abs().max())))
            bad = {k: v for k, v in errs.items() if not v < 1e-4}
            if flag != 0.0:
                bad["timeout_flag"] = flag
            print("rank %d %s step %d: %s %s" % (rank, mode, it, {k: "%.1e" % v for k, v in errs.items()},
                                                 "FAIL " + str(bad) if bad else "ok"), flush=True)
            ok = ok and not bad
    f = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(f)
    clr.dist.close_peer()
    dist.destroy_process_group()
    if int(f.item()) != 0:
        sys.exit(1)
    if rank == 0:
        print("DDP_CHECK_OK world=%d" % world)


if __name__ == "__main__":
    main()
