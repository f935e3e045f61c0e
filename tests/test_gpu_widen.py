"""GPU parity of the rows widened in round 2: A7 in the library (clr_ema_rows), the nearest label down-sample
(8(f) rank 2, third op), MC accumulation without the staging buffer (8(f) rank 1)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import uda_clr_b200 as clr
from oracle import clr_torch_port as TP
from uda_clr_b200 import synth
from _util import TOL_GRAD, TOL_LOSS, TOL_PROTO, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_ema_rows_equals_aten_expression_bit_for_bit_and_skips_zero_vectors():
    """Trainer_prototype.py:117-123: ``obj * (1 - 0.001) + 0.001 * vector`` unless ``vector.sum() == 0``."""
    g = torch.Generator().manual_seed(4)
    for C in (304, 305, 7):
        obj = torch.randn(C, generator=g).to(DEV)
        v = torch.randn(1, C, generator=g).to(DEV)
        ref = obj * (1 - 0.001) + 0.001 * v.squeeze()
        assert torch.equal(clr.update_objective_single_vector(obj, v), ref)
        assert torch.equal(clr.update_objective_single_vector(obj, torch.zeros_like(v)), obj)
    # a stack of vectors, one of them all-zero (an empty mask's prototype): per-row decision in ONE launch
    obj = torch.randn(3, 305, generator=g).to(DEV)
    v = torch.randn(3, 305, generator=g).to(DEV)
    v[1] = 0.0
    out = clr.update_objective_single_vector(obj, v)
    ref = obj * (1 - 0.001) + 0.001 * v
    assert torch.equal(out[0], ref[0]) and torch.equal(out[2], ref[2]) and torch.equal(out[1], obj[1])


@pytest.mark.parametrize("B,K,Hi,Wi,H,W", [(8, 2, 512, 512, 128, 128), (2, 2, 100, 75, 33, 20), (1, 3, 64, 64, 64, 64), (2, 1, 37, 41, 50, 60)])
def test_nearest_label_downsample_equals_aten(B, K, Hi, Wi, H, W):
    """Trainer_prototype_full.py:329-330 ``F.interpolate(target_map.clone(), size=oS_before.size()[2:], mode='nearest')``."""
    g = torch.Generator().manual_seed(Hi)
    m = (torch.rand(B, K, Hi, Wi, generator=g) > 0.5).float().to(DEV)
    assert torch.equal(clr.nearest_labels(m, H, W), F.interpolate(m, size=(H, W), mode="nearest"))


def test_fused_step_takes_image_resolution_source_labels():
    """CLRStep with ``pred_oS`` = the trainer's full-resolution target_map: same result as down-sampling first."""
    K, C, H, up, T, B = 2, 40, 32, 4, 8, 2
    b = synth.make_batch(B=B, C=C, H=H, W=H, K=K, T=T, up=up, seed=17)
    t = {k: getattr(b, k).to(DEV) for k in ("xs", "ys", "xt", "oT_before", "preds", "oT", "oT_aug")}
    g = torch.Generator().manual_seed(2)
    full = (torch.rand(B, K, H * up, H * up, generator=g) > 0.6).float().to(DEV)
    small = F.interpolate(full.clone(), size=(H, H), mode="nearest")
    res = []
    for ys in (small, full):
        step = clr.CLRStep(K=K, retrify=True, use_disc=True, use_cons=True)
        plan = step.plan(t["xs"], ys, t["xt"], oT_before=t["oT_before"], preds=t["preds"], T=T, oT=t["oT"], oT_aug=t["oT_aug"])
        plan.run(); plan.run()
        torch.cuda.synchronize()
        res.append((plan.losses.clone(), plan.gxs.clone(), plan.gxt.clone()))
        xs = t["xs"].clone().requires_grad_(True)
        out = clr.CLRStep(K=K, retrify=True, use_disc=True, use_cons=True)(
            xs, ys, t["xt"].clone().requires_grad_(True), oT_before=t["oT_before"], preds=t["preds"], T=T, oT=t["oT"], oT_aug=t["oT_aug"])
        out.total.backward()
        res.append((out.total.detach().clone(), xs.grad.clone()))
    assert all(torch.equal(x, y) for x, y in zip(res[0], res[2]))
    assert all(torch.equal(x, y) for x, y in zip(res[1], res[3]))


@pytest.mark.parametrize("T,passes,B,K,Hi", [(8, 2, 2, 2, 128), (8, 1, 1, 2, 64), (6, 3, 2, 3, 40), (5, 1, 1, 1, 33)])
def test_mc_accumulator_equals_staged_statistics(T, passes, B, K, Hi):
    """8(f) rank 1: running statistics over the MC forwards (no [T*B,...] staging buffer) vs torch.std / torch.mean of the
    staged stack (utils/Utils.py:164-168) and vs clr_mc_stats."""
    g = torch.Generator(device=DEV).manual_seed(T * 10 + Hi)
    base = 2.0 * torch.randn(B, K, Hi, Hi, generator=g, device=DEV)
    preds = base.repeat(T, 1, 1, 1) + 0.35 * torch.randn(T * B, K, Hi, Hi, generator=g, device=DEV)
    acc = clr.MCAccumulator()
    for rep in range(2):                       # the accumulator is reusable across steps
        for i in range(T // passes):
            acc.add(preds[i * passes * B:(i + 1) * passes * B], passes=passes)
        std_a, mean_a = acc.finalize()
    p5 = preds.reshape(T, B, K, Hi, Hi)
    ref_std = torch.std(torch.sigmoid(p5 / 2.0), dim=0)
    ref_mean = torch.mean(torch.sigmoid(p5), dim=0)
    assert float((std_a - ref_std).abs().max()) < 1e-6
    assert float((mean_a - ref_mean).abs().max()) < 1e-6
    std_s, mean_s = clr.mc_statistics(preds, T, B)
    assert float((std_a - std_s).abs().max()) < 1e-6
    # downstream: same weights away from the knife edge
    oT = base[:, :, ::1, ::1][:, :, :Hi // 4 * 4:4, :Hi // 4 * 4:4].contiguous() if Hi % 4 == 0 else None
    if oT is not None:
        h = Hi // 4
        w_a, m_a = clr.retrify_weights(oT, mean_a, std_a, h, h)
        w_s, m_s = clr.retrify_weights(oT, mean_s, std_s, h, h, preds=preds, T=T)
        small = F.interpolate(ref_std, size=(h, h), mode="bilinear", align_corners=True)
        away = (small - 0.04).abs() > 2e-6
        assert torch.equal(m_a[away], m_s[away])
