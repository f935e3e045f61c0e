"""The data-parallel contract of the sharded CLR step under REAL ``DistributedDataParallel``:

  forward : every rank pools its own shard, ONE exchange sums the packed ``[2K][C+1]`` buffers, every rank finalises the
            same prototypes / losses;
  backward: no collective -- each rank writes the gradient of ITS pixels scaled by the world size G, and DDP's gradient
            AVERAGING over ranks, ``(1/G) * sum_ranks (G * local)``, reproduces the single-process gradient of the model
            parameters that produced the features.

Here the contract runs on 2 gloo ranks on the CPU: the per-rank arithmetic is the eager port (the CUDA ops refuse CPU
tensors), wrapped in an autograd Function with exactly the sharded op's semantics (all-reduce in forward, scale-by-G and
no collective in backward).  ``tools/ddp_check.py`` / ``tests/test_gpu_multi.py`` run the same check with the CUDA step.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import clr_torch_port as TP
from uda_clr_b200 import synth
import uda_clr_b200 as clr


class _ShardedPool(torch.autograd.Function):
    """Prototypes from this rank's shard + all-reduced sums; backward as the CUDA path does it: ``dL/dmu`` is identical
    on all ranks, the local adjoint is scaled by G, nothing is exchanged."""

    @staticmethod
    def forward(ctx, feat, w):                      # feat [b,C,H,W], w [b,R,H,W] explicit rows
        import torch.distributed as dist
        S = torch.einsum("bchw,brhw->rc", feat, w)
        N = w.sum(dim=(0, 2, 3))
        packed = torch.cat([S, N[:, None]], 1)
        dist.all_reduce(packed)                     # THE exchange
        ctx.save_for_backward(w, packed)
        ctx.G = dist.get_world_size()
        return packed[:, :-1] / packed[:, -1:]

    @staticmethod
    def backward(ctx, g):                           # g [R,C]
        w, packed = ctx.saved_tensors
        gx = torch.einsum("rc,brhw->bchw", g / packed[:, -1:], w) * ctx.G
        return gx, None


def _step_loss(conv, x_s, ys, x_t, wt, pool):
    """source hard labels + soft target weights -> prototypes -> intra alignment loss (first step: no EMA history)."""
    fs, ft = conv(x_s), conv(x_t)
    K = ys.shape[1]
    ws = torch.cat([ys, 1.0 - ys], 1)
    wtt = torch.cat([wt, 1.0 - wt], 1)
    Ps, Pt = pool(fs, ws), pool(ft, wtt)
    mse = torch.nn.MSELoss()
    return sum(mse(Ps[r], Pt[r]) for r in range(2 * K))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    b = synth.make_batch(B=4, C=6, H=8, W=8, K=2, image_res=False, seed=5)
    return b.xs.double(), b.ys.double(), b.xt.double(), torch.sigmoid(b.oT_before).double()


def _make_conv():
    torch.manual_seed(3)
    return torch.nn.Conv2d(6, 10, 1).double()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        xs, ys, xt, wt = (clr.dist.shard_batch([t], rank, world)[0] for t in _data())
        conv = DDP(_make_conv())
        loss = _step_loss(conv, xs, ys, xt, wt, _ShardedPool.apply)
        loss.backward()                              # DDP averages the parameter gradients over the ranks
        q.put((rank, float(loss), conv.module.weight.grad.numpy().copy(), conv.module.bias.grad.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_ddp_averaging_of_world_scaled_shard_gradients_equals_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process, whole batch, plain autograd through the eager port of the reference's pooling
    xs, ys, xt, wt = _data()
    conv = _make_conv()

    def pool(feat, w):
        return torch.cat([p.reshape(1, -1) for p in TP._pool(feat, [w[:, r:r + 1] for r in range(w.shape[1])])], 0)

    loss = _step_loss(conv, xs, ys, xt, wt, pool)
    loss.backward()
    for rank, l, gw, gb in res:
        assert abs(l - float(loss)) < 1e-12 * abs(float(loss))
        assert np.allclose(gw, conv.weight.grad.numpy(), rtol=1e-9, atol=1e-18)
        assert np.allclose(gb, conv.bias.grad.numpy(), rtol=1e-9, atol=1e-16)
