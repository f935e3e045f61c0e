"""Trainer-protocol integration of the zero-line drop-in (VERDICT r01, row "Trainer-level drop-in"):

The shipped trainer drives the CLR block as: drop-in op -> inline EMA -> inline MSE -> ``loss.backward()``
(Trainer_prototype_full.py:330-355, 378-398, 428-449, 463-468).  ``bench.trainer_protocol_step`` restates that sequence
generically over the module that provides ``gen_prototype`` / ``gen_prototype_retrify``; here it runs for three steps
with this package's ops and with the eager port of the reference's ops on the same GPU: same ``loss_all`` contribution,
same gradients and gradient norms for both feature maps, EMA history included.

With ``UDA_CLR_REFERENCE=/path/to/UDA_CLR`` set on a box that has BOTH the reference tree and a GPU, the second test
runs the real, unmodified ``Trainer.train_epoch`` stock and patched (``patch_reference``) and compares the logged
losses of the first steps.  (The GPU box of this project has no reference tree, the dev container no GPU: it is
skipped in both, the seam itself is covered by tests/test_patch_and_dist.py.)
"""
import os

import numpy as np
import pytest
import torch

import uda_clr_b200 as clr
from bench import trainer_protocol_step
from oracle import clr_torch_port as TP
from uda_clr_b200 import synth
from _util import TOL_GRAD, TOL_LOSS, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("B,C,H", [(4, 305, 128), (2, 64, 32)])
def test_trainer_inline_protocol_three_steps_dropin_vs_eager_ops(B, C, H):
    K, up, T = 2, 4, 8
    st_ours, st_ref = {}, {}
    for it in range(3):
        b = synth.make_batch(B=B, C=C, H=H, W=H, K=K, T=T, up=up, seed=300 + it)
        d = {k: getattr(b, k).to(DEV) for k in ("xs", "ys", "xt", "oT_before", "preds")}
        l1, i1, gxs1, gxt1 = trainer_protocol_step(clr, st_ours, d, T)
        l2, i2, gxs2, gxt2 = trainer_protocol_step(TP, st_ref, d, T)
        assert abs(float(l1) - float(l2)) < TOL_LOSS * abs(float(l2)), it
        assert abs(float(i1) - float(i2)) < TOL_LOSS * abs(float(i2)), it
        assert relerr(gxs1.cpu().numpy(), gxs2.cpu().numpy()) < TOL_GRAD
        assert relerr(gxt1.cpu().numpy(), gxt2.cpu().numpy()) < TOL_GRAD
        for g1, g2 in ((gxs1, gxs2), (gxt1, gxt2)):
            assert abs(float(g1.norm()) - float(g2.norm())) < TOL_GRAD * float(g2.norm())
        # the EMA history the trainer keeps (detached prototypes) agrees too
        for dom in ("s", "t"):
            a = torch.cat([p.reshape(1, -1) for p in st_ours[dom]]).cpu().numpy()
            r = torch.cat([p.reshape(1, -1) for p in st_ref[dom]]).cpu().numpy()
            assert relerr(a, r) < 1e-5


def test_real_trainer_train_epoch_stock_vs_patched():
    ref = os.environ.get("UDA_CLR_REFERENCE")
    if not ref or not os.path.isdir(ref):
        pytest.skip("needs the reference tree (UDA_CLR_REFERENCE) on a GPU box")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
    import trainer_harness as TH
    stock = TH.run_train_epoch(ref, patched=False, steps=2, batch_size=2, seed=1337)
    patched = TH.run_train_epoch(ref, patched=True, steps=2, batch_size=2, seed=1337)
    assert len(stock["calls"]) == len(patched["calls"]) > 0
    for a, b in zip(stock["calls"], patched["calls"]):
        assert a["name"] == b["name"]
        for x, y in zip(a["protos"], b["protos"]):
            assert np.abs(x - y).max() <= 1e-4 * max(np.abs(y).max(), 1e-30)


@pytest.mark.parametrize("running_mean", [False, True])
def test_offline_prototype_extraction_matches_cal_prototype_loop(running_mean):
    """``cal_prototype.py:139-195`` per batch: thresholds 0.5 / 0.1 / 0.5, three bmm-style poolings, the "running mean"
    (which in the reference is the identity: the LAST batch's prototypes are saved), the saved dict.  Checked against a
    direct eager restatement of those lines on the same GPU (``TP.bmm_pool`` is the reference's bmm sequence)."""
    import tempfile
    g = torch.Generator().manual_seed(12)
    acc = clr.OfflinePrototypes(running_mean=running_mean)
    ref = {"bu": None, "cup": None, "disc": None}
    n = 0
    for it in range(3):
        B, h = 2, 64
        o = (2.0 * torch.randn(B, 2, h, h, generator=g)).to(DEV)
        bd = (2.0 * torch.randn(B, 1, h, h, generator=g)).to(DEV)
        x_bu = torch.randn(B, 304, h, h, generator=g).to(DEV)
        x = torch.randn(B, 305, h, h, generator=g).to(DEV)
        out = acc.update(o, bd, x_bu, x)
        pred = torch.sigmoid(o).clone()
        q_disc, q_cup = (pred[:, 1] > 0.5), (pred[:, 0] > 0.1)                       # :145-146
        bu_t = torch.sigmoid(bd).clone()
        bu_t[bu_t > 0.5] = 1
        bu_t[bu_t <= 0.5] = 0                                                        # :149-151
        cur = {"bu": TP.bmm_pool(bu_t, x_bu).squeeze(), "cup": TP.bmm_pool(q_cup.float().unsqueeze(1), x).squeeze(),
               "disc": TP.bmm_pool(q_disc.float().unsqueeze(1), x).squeeze()}
        for k in cur:
            if running_mean and ref[k] is not None:
                ref[k] = (ref[k] * n + cur[k]) / (n + 1)
            else:
                ref[k] = (cur[k] * n + cur[k]) / (n + 1) if not running_mean else cur[k]    # :177-190: the identity
        n = min(n + 1, 3000)
        for k in ("bu", "cup", "disc"):
            assert out[k].shape == ref[k].shape
            assert relerr(out[k].cpu().numpy(), ref[k].cpu().numpy()) < 1e-5, (it, k)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "prototypes_on_Drishti-GS_from_prototype")
        acc.save(path)
        back = torch.load(path)
        assert set(back) == {"bu", "cup", "disc"} and back["bu"].shape == (304,) and back["cup"].shape == (305,)
