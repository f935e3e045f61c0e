"""Host-side logic of bench.py that needs no GPU: the SURVEY 8(d) byte model and the trainer-protocol restatement."""
import types

import torch

import bench
from oracle import clr_torch_port as TP
from uda_clr_b200 import synth


def _args(workload="clr3", **kw):
    d = dict(B=8, C=256, H=128, K=2, T=8, up=4, workload=workload)
    d.update(kw)
    return types.SimpleNamespace(**d)


def test_algorithmic_bytes_are_survey_8d():
    """clr3 at config 1 = 4(F+Lb) + (T*Li + Li + Lb) + (F+Lb) + (2Li+Lb) = 863 MB; align = 4F + 4Lb = 541 MB (SURVEY.md 8(d),
    BASELINE.md 3); the implementation's extra traffic is reported separately and is larger."""
    ab = bench.algorithmic_bytes(_args())
    F, Lb, Li = 4 * 8 * 256 * 128 * 128, 4 * 8 * 2 * 128 * 128, 4 * 8 * 2 * 512 * 512
    assert ab["total"] == 4 * (F + Lb) + (8 * Li + Li + Lb) + (F + Lb) + (2 * Li + Lb)
    assert round(ab["total"] / 1e6) == 863
    assert ab["pool_fwd"] == 2 * (F + Lb) and ab["pool_bwd"] == 2 * (F + Lb)
    al = bench.algorithmic_bytes(_args("align"))
    assert al["total"] == 4 * F + 4 * Lb and round(al["total"] / 1e6) == 541
    assert bench.implementation_bytes(_args())["total"] > ab["total"]


def test_trainer_protocol_restatement_runs_the_reference_sequence_on_cpu():
    """``bench.trainer_protocol_step`` (Trainer_prototype_full.py:330-449) with the eager port's ops on the CPU: first-step copy,
    then EMA; the loss is pro_weight * intra and gradients reach both feature maps."""
    st = {}
    for it in range(2):
        b = synth.make_batch(B=2, C=6, H=8, W=8, K=2, T=4, up=2, seed=it)
        d = dict(xs=b.xs, ys=b.ys, xt=b.xt, oT_before=b.oT_before, preds=b.preds)
        loss, inter, gxs, gxt = bench.trainer_protocol_step(TP, st, d, 4)
        assert torch.isfinite(loss) and torch.isfinite(inter)
        assert gxs.shape == b.xs.shape and gxt.shape == b.xt.shape and float(gxs.abs().max()) > 0 and float(gxt.abs().max()) > 0
        assert len(st["s"]) == 4 and not st["s"][0].requires_grad
    # second step used the stored prototypes: P = 0.1 * stored + 0.9 * current
    cur = TP.gen_prototype(b.ys, b.xs)
    assert not torch.equal(st["s"][0], cur[0].detach())
