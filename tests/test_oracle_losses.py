"""Closed forms of the bytecode-only losses (A9 discriminative hinge, A10 consistency BCE) and of the
variant-A pieces (A6-A8) vs the op-for-op torch transcription under autograd.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import clr_oracle as O
from oracle import clr_torch_port as TP
from uda_clr_b200 import synth


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("K", [2, 3])
def test_disc_loss_and_grads(K):
    b = synth.make_batch(B=2, C=11, H=9, W=8, K=K, image_res=False, seed=31)
    g = torch.Generator().manual_seed(1)
    xs = b.xs.double().requires_grad_(True)
    P = [(0.3 * torch.randn(1, 11, 1, 1, generator=g)).double().requires_grad_(True) for _ in range(2 * K)]
    loss = TP.disc_loss(xs, b.ys.double(), P, 0.01)
    loss.backward()
    Pn = np.concatenate([p.detach().reshape(1, -1).numpy() for p in P], 0)
    l, aux = O.disc_loss(b.xs.numpy(), b.ys.numpy(), Pn, 0.01)
    assert abs(l - float(loss.detach())) < 1e-12
    gx, gP = O.disc_grads(b.xs.numpy(), b.ys.numpy(), Pn, 0.01)
    assert relerr(gx, xs.grad.numpy()) < 1e-10
    assert relerr(gP, np.concatenate([p.grad.reshape(1, -1).numpy() for p in P], 0)) < 1e-10
    # both hinge branches must be exercised by the synthetic case
    assert (aux["coef"] > 0).any() and (aux["coef"] < 0).any() and (aux["coef"] == 0).any()


def test_cons_loss_and_grad():
    g = torch.Generator().manual_seed(3)
    B, K, H, W, up = 2, 2, 6, 5, 4
    oT = 3.0 * torch.randn(B, K, H * up, W * up, generator=g)
    oT_aug = (oT + torch.randn(oT.shape, generator=g)).double().requires_grad_(True)
    masks = [2.0 * (torch.rand(B, 1, H, W, generator=g) > 0.4).float() for _ in range(K)]
    for epoch in (0.0, 37.0, 250.0):
        oT_aug.grad = None
        loss = TP.cons_loss(oT.double(), oT_aug, [m.double() for m in masks], epoch, aug_weight=0.7)
        loss.backward()
        thr = O.consistency_threshold(epoch)
        l, gz, aux = O.cons_loss(oT.numpy(), oT_aug.detach().float().numpy(),
                                 torch.cat(masks, 1).numpy(), thr, 0.7)
        # the oracle takes sigmoid in fp32 (as the fp32 reference does); the fp64 transcription differs ~1e-7
        assert abs(l - float(loss.detach())) < 2e-6 * abs(l)
        assert relerr(gz, oT_aug.grad.numpy()) < 2e-5
        assert aux["y"].min() == 0 and aux["y"].max() == 1


def test_cons_loss_fp32_path_and_saturation():
    """fp32 transcription incl. saturated logits (log clamp at -100, grad denominator clamp 1e-12)."""
    g = torch.Generator().manual_seed(5)
    oT = 3.0 * torch.randn(1, 2, 8, 8, generator=g)
    oT_aug = 3.0 * torch.randn(1, 2, 8, 8, generator=g)
    oT_aug[0, 0, 0, :4] = torch.tensor([40.0, -40.0, 120.0, -120.0])
    oT[0, 0, 0, :4] = torch.tensor([-5.0, 5.0, -5.0, 5.0])
    za = oT_aug.clone().requires_grad_(True)
    masks = [2.0 * torch.ones(1, 1, 2, 2), 2.0 * torch.ones(1, 1, 2, 2)]
    loss = TP.cons_loss(oT, za, masks, 10.0)
    loss.backward()
    l, gz, _ = O.cons_loss(oT.numpy(), oT_aug.numpy(), torch.cat(masks, 1).numpy(), O.consistency_threshold(10.0))
    assert np.isfinite(l) and abs(l - float(loss.detach())) < 1e-5 * abs(l)
    assert relerr(gz, za.grad.numpy()) < 1e-5


def test_variant_a_pieces():
    g = torch.Generator().manual_seed(8)
    x = torch.randn(3, 10, 6, 7, generator=g)
    m = (torch.rand(3, 1, 6, 7, generator=g) > 0.5).float()
    assert relerr(O.bmm_pool(m.numpy(), x.numpy()), TP.bmm_pool(m, x).numpy().reshape(-1)) < 1e-6
    p = torch.randn(10, generator=g)
    assert relerr(O.feat_prototype_distance(x.numpy(), p.numpy()),
                  TP.feat_prototype_distance(x, p)[:, 0].numpy()) < 1e-6
    assert relerr(O.distance_weight(x.numpy(), p.numpy()), TP.distance_weight(x, p)[:, 0].numpy()) < 1e-5
    assert relerr(O.cosine_weight(x.numpy(), p.numpy()), TP.cosine_weight(x, 1, p.view(1, -1, 1, 1)).numpy()) < 1e-6
    obj = np.arange(10, dtype=np.float64)
    assert np.array_equal(O.ema_single_vector(obj, np.zeros(10)), obj)
    v = np.ones(10)
    assert np.allclose(O.ema_single_vector(obj, v), obj * 0.999 + 0.001)
    assert O.adaptation_factor(2.0) == TP.adaptation_factor(2.0)


def test_full_step_matches_port_autograd():
    """The fused closed-form step (A1 + A2 weights + A4 + A5 + A9 + A10) vs autograd through the port, 2 steps."""
    K, C, H, W, B, T, up = 2, 7, 8, 8, 2, 3, 2
    port = TP.ClrStepPort(decay=0.9, pro_weight=0.1, src_reg_weight=0.5, aug_weight=0.8, retrify=False,
                          use_disc=True, use_cons=False)
    stored_s = stored_t = None
    for step in range(2):
        b = synth.make_batch(B=B, C=C, H=H, W=W, K=K, T=T, up=up, seed=40 + step)
        xs = b.xs.double().requires_grad_(True)
        xt = b.xt.double().requires_grad_(True)
        oTb = b.oT_before.double().requires_grad_(True)
        res = port.step(xs, b.ys.double(), xt, oTb)
        pt = torch.sigmoid(b.oT_before.double()).numpy()
        o = O.clr_step(b.xs.numpy(), b.ys.numpy(), b.xt.numpy(), np.concatenate([pt, 1 - pt], 1),
                       stored_s=stored_s, stored_t=stored_t, decay=0.9, w_intra=0.1, w_disc=0.5)
        stored_s, stored_t = o["Ps"], o["Pt"]
        assert abs(o["total"] - float(res["total"])) < 1e-11
        assert abs(o["loss_disc"] - float(res["disc"])) < 1e-11
        assert relerr(o["gxs"], xs.grad.numpy()) < 1e-9
        assert relerr(o["gxt"], xt.grad.numpy()) < 1e-9
        # soft target: grad reaches the logits through sigmoid'
        gw = o["gwt"][:, :K] - o["gwt"][:, K:]
        assert relerr(gw * pt * (1 - pt), oTb.grad.numpy()) < 1e-9


def test_seg_loss_and_entropy_closed_forms_vs_aten():
    """8(f) glue: the fp64 closed forms of loss_seg (Trainer_prototype_full.py:292-294) and of the uncertainty map
    (:452) with their gradients vs the ATen op sequence the trainer runs, under autograd in fp64."""
    g = torch.Generator().manual_seed(5)
    oS = (3.0 * torch.randn(2, 2, 12, 10, generator=g)).double().requires_grad_(True)
    bS = (2.0 * torch.randn(2, 1, 12, 10, generator=g)).double().requires_grad_(True)
    tmap = (torch.rand(2, 2, 12, 10, generator=g) > 0.5).double()
    tbd = torch.rand(2, 1, 12, 10, generator=g).double()
    loss = TP.seg_loss(oS, bS, tmap, tbd)
    loss.backward()
    l, aux = O.seg_loss(oS.detach().numpy(), bS.detach().numpy(), tmap.numpy(), tbd.numpy())
    assert abs(l - float(loss.detach())) < 1e-12
    assert relerr(aux["g_oS"], oS.grad.numpy()) < 1e-10
    assert relerr(aux["g_boundaryS"], bS.grad.numpy()) < 1e-10
    o = (4.0 * torch.randn(3, 2, 9, 7, generator=g)).double().requires_grad_(True)
    w = torch.randn(3, 2, 9, 7, generator=g).double()
    u = TP.uncertainty_map(o)
    (u * w).sum().backward()
    un, du = O.uncertainty_map(o.detach().numpy())
    assert relerr(un, u.detach().numpy()) < 1e-12
    assert relerr(du * w.numpy(), o.grad.numpy()) < 1e-10
