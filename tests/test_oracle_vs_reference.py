"""Pin the oracle against the UNMODIFIED reference, imported in place (dev container only).

Skipped wherever ``/root/reference`` is absent (the GPU box); the committed fixtures under
``tests/golden`` carry the same pin there (tests/test_golden.py).
"""
import numpy as np
import pytest
import torch

from oracle import clr_oracle as O
from oracle import clr_torch_port as TP
from oracle import ref_import
from uda_clr_b200 import synth

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")]


def _stack(protos):
    return torch.cat([p.reshape(1, -1) for p in protos], 0).detach().numpy()


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("soft", [False, True])
@pytest.mark.parametrize("shape", [(2, 5, 6, 7), (3, 17, 16, 16)])
def test_gen_prototype_fwd_bwd_fp64(shape, soft):
    """A1 forward and the closed-form adjoint vs the reference under autograd, both in float64."""
    B, C, H, W = shape
    g = torch.Generator().manual_seed(7)
    y = synth.nested_ellipse_labels(B, 2, H, W, g)
    pred = torch.sigmoid(synth.confident_logits(y, g)) if soft else y
    x = synth.class_shifted_features(y, C, g)
    pred64 = pred.double().requires_grad_(True)
    x64 = x.double().requires_grad_(True)
    ref = ref_import.ref_gen_prototype(pred64, x64)
    seeds = [torch.randn(p.shape, generator=g, dtype=torch.float64) for p in ref]
    sum((p * s).sum() for p, s in zip(ref, seeds)).backward()

    # the oracle takes (1 - pred) in fp32 like the fp32 reference; feed it the fp64 pair directly
    w = np.concatenate([pred.double().numpy(), 1.0 - pred.double().numpy()], axis=1)
    S, N = O.pool_sums(x.numpy(), w)
    mu = O.prototypes_from_sums(S, N)
    assert relerr(mu, _stack(ref)) < 1e-12
    gseed = np.concatenate([s.reshape(1, -1).numpy() for s in seeds], 0)
    gx, gw = O.pool_backward(x.numpy(), w, gseed)
    assert relerr(gx, x64.grad.numpy()) < 1e-11
    assert relerr(gw[:, :2] - gw[:, 2:], pred64.grad.numpy()) < 1e-10


def test_gen_prototype_fp32_and_port_bitexact():
    """fp32 reference vs oracle (1e-6) and vs the torch port (bit-for-bit, same op sequence)."""
    b = synth.make_batch(B=2, C=19, H=16, W=12, K=2, image_res=False, seed=3)
    xs = b.xs.clone().requires_grad_(True)
    ref = ref_import.ref_gen_prototype(b.ys, xs)
    port = TP.gen_prototype(b.ys, b.xs)
    for r, p in zip(ref, port):
        assert torch.equal(r.detach(), p)
    assert relerr(O.gen_prototype(b.ys.numpy(), b.xs.numpy()), _stack(ref)) < 2e-6
    # hard-label counts are exact integers
    _, N = O.pool_sums(b.xs.numpy(), O.weights_complement(b.ys.numpy()))
    assert np.array_equal(N, np.round(N))
    assert N[0] == float(b.ys[:, 0].sum()) and N[3] == float((1 - b.ys[:, 1]).sum())


def test_empty_class_is_nan_like_reference():
    """utils/Utils.py:127-130 has no guard: an empty class gives 0/0 = NaN."""
    y = torch.zeros(1, 2, 4, 4)
    y[:, 1, 1:3, 1:3] = 1.0
    x = torch.randn(1, 3, 4, 4)
    ref = _stack(ref_import.ref_gen_prototype(y, x))
    ours = O.gen_prototype(y.numpy(), x.numpy())
    assert np.isnan(ref[0]).all() and np.isnan(ours[0]).all()
    assert relerr(ours[1:], ref[1:]) < 1e-6


def test_src_trg_joint_prototypes():
    b = synth.make_batch(B=2, C=9, H=8, W=8, K=2, image_res=False, seed=5)
    pt = torch.sigmoid(b.oT_before)
    ref = _stack(ref_import.ref_gen_prototype_src_trg(b.ys, b.xs, pt, b.xt))
    ours = O.gen_prototype_src_trg(b.ys.numpy(), b.xs.numpy(), pt.numpy(), b.xt.numpy())
    assert relerr(ours, ref) < 2e-6


def _retrify_case(seed=11, B=1, C=3, T=4, Hi=160):
    g = torch.Generator().manual_seed(seed)
    H = W = 128
    yt = synth.nested_ellipse_labels(B, 2, H, W, g)
    xt = synth.class_shifted_features(yt, C, g)
    oT = synth.confident_logits(yt, g)
    base = torch.nn.functional.interpolate(oT, size=(Hi, Hi), mode="nearest")
    preds = base.repeat(T, 1, 1, 1) + 0.35 * torch.randn(T * B, 2, Hi, Hi, generator=g)
    return xt, oT, preds, T, B


def test_retrify_vs_reference():
    """A2: MC statistics, bilinear(align_corners) downsample, thresholds, weighted pooling."""
    xt, oT, preds, T, B = _retrify_case()
    xt_r = xt.clone().requires_grad_(True)
    oT_r = oT.clone().requires_grad_(True)
    ref = ref_import.ref_gen_prototype_retrify(oT_r, xt_r, preds, T, B)
    ours = O.gen_prototype_retrify(oT.numpy(), xt.numpy(), preds.numpy(), T, B)
    assert relerr(ours["std_map"], ref[4].detach().numpy()) < 5e-6
    m0, m1 = ref[5].numpy(), ref[6].numpy()
    # masks: identical except possibly at |std_small - 0.04| knife edges
    diff0 = ours["mask_0"] != m0
    diff1 = ours["mask_1"] != m1
    for d, k in ((diff0, 0), (diff1, 1)):
        if d.any():
            assert np.abs(ours["std_small"][:, k:k + 1][d] - 0.04).max() < 1e-6
    assert diff0.mean() < 1e-3 and diff1.mean() < 1e-3
    assert 0.02 < (m0 > 0).mean() < 0.98, "synthetic case must exercise both mask states"
    if not (diff0.any() or diff1.any()):
        assert relerr(ours["protos"], _stack(ref[:4])) < 1e-5
    # gradient: only xt_feature receives one; oT_before's is exactly zero
    seeds = [torch.randn(p.shape, generator=torch.Generator().manual_seed(1)) for p in ref[:4]]
    sum((p * s).sum() for p, s in zip(ref[:4], seeds)).backward()
    assert float(oT_r.grad.abs().max()) == 0.0
    gseed = np.concatenate([s.reshape(1, -1).numpy() for s in seeds], 0)
    gx, _ = O.pool_backward(xt.numpy(), ours["w"], gseed)
    if not (diff0.any() or diff1.any()):
        assert relerr(gx, xt_r.grad.numpy()) < 1e-5


def test_retrify_port_matches_reference_bitexact():
    xt, oT, preds, T, B = _retrify_case(seed=12, C=2, T=3, Hi=144)
    ref = ref_import.ref_gen_prototype_retrify(oT, xt, preds, T, B)
    port = TP.gen_prototype_retrify(oT, xt, preds, None, T, B)
    assert len(ref) == len(port) == 7
    for r, p in zip(ref, port):
        assert torch.equal(r, p)


def test_bilinear_matches_aten():
    g = torch.Generator().manual_seed(2)
    x = torch.rand(2, 2, 37, 53, generator=g)
    ref = torch.nn.functional.interpolate(x, size=(9, 14), mode="bilinear", align_corners=True)
    assert relerr(O.bilinear_align_corners(x.numpy(), 9, 14), ref.numpy()) < 1e-6


def test_nearest_matches_aten():
    x = torch.arange(2 * 1 * 5 * 7, dtype=torch.float32).reshape(2, 1, 5, 7)
    for size in ((20, 28), (13, 9), (5, 7)):
        ref = torch.nn.functional.interpolate(x, size=size, mode="nearest")
        assert np.array_equal(O.nearest_upsample(x.numpy(), *size), ref.numpy())


def test_cosine_weight_and_adaptation_factor():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 11, 5, 6, generator=g)
    p = torch.randn(1, 11, 1, 1, generator=g)
    ref = ref_import.ref_get_prototype_weight(x, 1, p)
    assert ref.shape == (2, 1, 5, 6)
    assert relerr(O.cosine_weight(x.numpy(), p.numpy()), ref.numpy()) < 1e-6
    for m in (-1, 0, 3, 25.5):
        assert O.adaptation_factor(m) == ref_import.ref_adaptation_factor(m)


def test_trainer_ema_and_alignment_block():
    """A4/A5: run the reference trainer's inline block semantics through the port (which transcribes
    Trainer_prototype_full.py:335-355, 378-398, 428-444) and compare with the closed forms."""
    b = synth.make_batch(B=2, C=13, H=8, W=8, K=2, image_res=False, seed=9)
    ema_s, ema_t = TP.PrototypeEMA(0.9), TP.PrototypeEMA(0.9)
    stored_s = stored_t = None
    for step in range(3):
        bb = synth.make_batch(B=2, C=13, H=8, W=8, K=2, image_res=False, seed=20 + step)
        xs = bb.xs.double().requires_grad_(True)
        xt = bb.xt.double().requires_grad_(True)
        pt = torch.sigmoid(bb.oT_before).double()
        Ps = ema_s.update(ref_import.ref_gen_prototype(bb.ys.double(), xs))
        Pt = ema_t.update(ref_import.ref_gen_prototype(pt, xt))
        intra, inter = TP.align_losses(Ps, Pt)
        (0.1 * intra).backward()
        o = O.clr_step(bb.xs.numpy(), bb.ys.numpy(), bb.xt.numpy(),
                       np.concatenate([pt.numpy(), 1 - pt.numpy()], 1),
                       stored_s=stored_s, stored_t=stored_t, decay=0.9, w_intra=0.1)
        stored_s, stored_t = o["Ps"], o["Pt"]
        assert abs(o["intra"] - float(intra)) < 1e-12 * max(1.0, abs(float(intra)))
        assert abs(o["inter"] - float(inter)) < 1e-12 * max(1.0, abs(float(inter)))
        assert relerr(o["Ps"], _stack(Ps)) < 1e-12
        assert relerr(o["gxs"], xs.grad.numpy()) < 1e-10
        assert relerr(o["gxt"], xt.grad.numpy()) < 1e-10
    del b


# ---- 8(f) rank 4: TransNorm (networks/sync_batchnorm/batchnorm.py:439-521)
@pytest.mark.parametrize("shape", [(4, 6, 5, 7), (6, 10, 8, 8), (5, 3, 4, 4)])
def test_transnorm_oracle_and_port_vs_reference(shape):
    """fp64 reference module vs the oracle (forward, adjoint, running estimates, eval) and fp32 reference vs the torch
    port (bit-for-bit: same ATen sequence)."""
    B, C, H, W = shape
    g = torch.Generator().manual_seed(21)
    x = torch.randn(B, C, H, W, generator=g) * 1.7 + torch.randn(1, C, 1, 1, generator=g)
    w, b = 0.5 + torch.rand(C, generator=g), torch.randn(C, generator=g) * 0.2
    gy = torch.randn(B, C, H, W, generator=g)
    cls = ref_import.ref_transnorm_class()

    m = cls(C).double()
    with torch.no_grad():
        m.weight.copy_(w)
        m.bias.copy_(b)
    x64 = x.double().requires_grad_(True)
    y = m(x64)
    (y * gy.double()).sum().backward()
    fw = O.transnorm_train(x.numpy(), w.numpy(), b.numpy())
    assert relerr(fw["y"], y.detach().numpy()) < 1e-12
    gx, gw, gb = O.transnorm_train_backward(x.numpy(), w.numpy(), gy.numpy())
    assert relerr(gx, x64.grad.numpy()) < 1e-10
    assert relerr(gw, m.weight.grad.numpy()) < 1e-11
    assert relerr(gb, m.bias.grad.numpy()) < 1e-11
    assert relerr(O.transnorm_running(np.zeros(C), fw["mean"][0], 0.1), m.running_mean_source.numpy()) < 1e-12
    assert relerr(O.transnorm_running(np.ones(C), fw["var_u"][1], 0.1), m.running_var_target.numpy()) < 1e-12
    m.eval()
    with torch.no_grad():
        ye = m(x.double())
    yo = O.transnorm_eval(x.numpy(), w.numpy(), b.numpy(), m.running_mean_source.numpy(), m.running_var_source.numpy(),
                          m.running_mean_target.numpy(), m.running_var_target.numpy())
    assert relerr(yo, ye.numpy()) < 1e-12

    m32 = cls(C)
    with torch.no_grad():
        m32.weight.copy_(w)
        m32.bias.copy_(b)
    bufs = [torch.zeros(C), torch.ones(C), torch.zeros(C), torch.ones(C)]
    xr, xp = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    wp, bp = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = m32(xr)
    yp = TP.trans_norm(xp, wp, bp, *bufs, True, 0.1, 1e-5)
    assert torch.equal(yr, yp)
    (yr * gy).sum().backward()
    (yp * gy).sum().backward()
    assert torch.equal(xr.grad, xp.grad) and torch.equal(m32.weight.grad, wp.grad) and torch.equal(m32.bias.grad, bp.grad)
    assert torch.equal(m32.running_var_source, bufs[1]) and torch.equal(m32.running_mean_target, bufs[2])
    m32.eval()
    with torch.no_grad():
        assert torch.equal(m32(x), TP.trans_norm(x, w, b, *bufs, False, 0.1, 1e-5))


# ---- K > 2: the reference is hard-wired to two classes; class k's prototypes depend on channel k only, so the K-class
# oracle / port must equal the reference applied to channel PAIRS
def test_k4_generalisation_equals_reference_on_channel_pairs():
    K = 4
    g = torch.Generator().manual_seed(31)
    B, C, H, W = 2, 7, 12, 10
    y = synth.nested_ellipse_labels(B, K, H, W, g)
    x = synth.class_shifted_features(y, C, g)
    soft = torch.sigmoid(synth.confident_logits(y, g))
    for pred in (y, soft):
        ours = O.gen_prototype(pred.numpy(), x.numpy())                       # rows obj_0..obj_3, bck_0..bck_3
        port = _stack(TP.gen_prototype(pred, x))
        for pair in ((0, 1), (2, 3), (3, 0)):
            ref = _stack(ref_import.ref_gen_prototype(pred[:, list(pair)].contiguous(), x))   # (c0_obj, c1_obj, c0_bck, c1_bck)
            rows = [pair[0], pair[1], K + pair[0], K + pair[1]]
            assert relerr(ours[rows], ref) < 2e-6
            assert np.array_equal(port[rows], ref)                            # same ATen sequence: bit for bit


def test_k4_retrify_equals_reference_on_channel_pairs():
    """A2 at K = 4 (B = 1, 128 x 128 features as the reference hard-codes, T = 3): std_map, masks and prototypes of classes
    (2, 3) from the K-class oracle / port against the reference run on that channel pair."""
    K, T, B, C, Hi = 4, 3, 1, 3, 144
    g = torch.Generator().manual_seed(41)
    yt = synth.nested_ellipse_labels(B, K, 128, 128, g)
    xt = synth.class_shifted_features(yt, C, g)
    oT = synth.confident_logits(yt, g)
    base = torch.nn.functional.interpolate(oT, size=(Hi, Hi), mode="nearest")
    preds = base.repeat(T, 1, 1, 1) + 0.35 * torch.randn(T * B, K, Hi, Hi, generator=g)
    ours = O.gen_prototype_retrify(oT.numpy(), xt.numpy(), preds.numpy(), T, B)
    port = TP.gen_prototype_retrify(oT, xt, preds, None, T, B)
    for pair in ((2, 3), (0, 1)):
        sel = list(pair)
        ref = ref_import.ref_gen_prototype_retrify(oT[:, sel].contiguous(), xt, preds[:, sel].contiguous(), T, B)
        assert relerr(ours["std_map"][:, sel], ref[4].numpy()) < 5e-6
        rows = [pair[0], pair[1], K + pair[0], K + pair[1]]
        masks_ref = (ref[5].numpy(), ref[6].numpy())
        same = all(np.array_equal(ours["masks"][:, k:k + 1], m) for k, m in zip(pair, masks_ref))
        assert same, "seeded case is expected to have no |std_small - 0.04| knife edge"
        assert 0.02 < float((masks_ref[0] > 0).mean()) < 0.98      # both mask states exercised
        assert relerr(ours["protos"][rows], _stack(ref[:4])) < 1e-5
        # port: same ATen sequence generalised over K -> bit for bit
        assert np.array_equal(_stack(port[:2 * K])[rows], _stack(ref[:4]))
        assert torch.equal(port[2 * K][:, sel], ref[4])
        assert torch.equal(port[2 * K + 1 + pair[0]], ref[5]) and torch.equal(port[2 * K + 1 + pair[1]], ref[6])


# ---- 8(f) rank 3: validation metrics (utils/metrics.py:81-168) from exact integer confusion counts
def test_metrics_from_counts_equal_reference_metrics():
    """``dice_from_counts`` / ``pixel_acc_from_counts`` (host-side formulas over the [K, 4] counts that ``clr_seg_counts``
    produces on the device) against the reference's ``dice_coeff_2label`` and ``pixel_acc`` on the same logits / labels."""
    import sys
    from uda_clr_b200 import ops
    ref_import.load_utils()
    M = sys.modules["utils.metrics"]
    had = hasattr(np, "bool")
    if not had:
        np.bool = bool            # the reference predates numpy 1.24 (utils/metrics.py:85-86 use np.bool)
    try:
        g = torch.Generator().manual_seed(77)
        B, K, H, W = 3, 2, 40, 36
        target = synth.nested_ellipse_labels(B, K, H, W, g)
        logits = 2.5 * torch.randn(B, K, H, W, generator=g) + 2.0 * (2 * target - 1)
        d_cup, d_disc = M.dice_coeff_2label(logits.clone(), target.clone())
        pa_cup, pa_disc, iou_cup, iou_disc = M.pixel_acc(logits.clone(), target.clone())
    finally:
        if not had:
            del np.bool
    pred = (torch.sigmoid(logits) > 0.75)
    gt = target != 0
    counts = torch.zeros(K, 4, dtype=torch.int64)
    for k in range(K):
        idx = (2 * gt[:, k].long() + pred[:, k].long()).flatten()
        counts[k] = torch.bincount(idx, minlength=4)
    dice = ops.dice_from_counts(counts)
    pa, miou = ops.pixel_acc_from_counts(counts)
    assert abs(float(dice[0]) - d_cup) < 1e-15 and abs(float(dice[1]) - d_disc) < 1e-15
    assert abs(float(pa[0]) - pa_cup) < 1e-15 and abs(float(pa[1]) - pa_disc) < 1e-15
    assert abs(float(miou[0]) - iou_cup) < 1e-15 and abs(float(miou[1]) - iou_disc) < 1e-15


# ---- A7 / A8: the Trainer_prototype.py helper methods (98-123), called unbound on a stand-in ``self``
def test_distance_weight_and_single_vector_ema_vs_trainer_prototype_methods():
    import sys
    import types
    ref_import.load_utils()
    sys.modules.setdefault("pytz", types.ModuleType("pytz"))
    import train_process.Trainer_prototype as TPR
    g = torch.Generator().manual_seed(5)
    N, C, H, W = 2, 11, 6, 7
    feat = torch.randn(N, C, H, W, generator=g)
    proto = torch.randn(C, generator=g)
    me = types.SimpleNamespace(objective_vectors={"cup": proto.clone()})
    me.feat_prototype_distance = types.MethodType(TPR.Trainer.feat_prototype_distance, me)
    d_ref = TPR.Trainer.feat_prototype_distance(me, feat, proto, 1)
    w_ref = TPR.Trainer.get_prototype_weight(me, feat, 1, "cup")
    assert relerr(O.feat_prototype_distance(feat.numpy(), proto.numpy()), d_ref[:, 0].numpy()) < 1e-6
    assert relerr(O.distance_weight(feat.numpy(), proto.numpy()), w_ref[:, 0].numpy()) < 2e-6
    assert torch.equal(TP.feat_prototype_distance(feat, proto, 1), d_ref) and torch.equal(TP.distance_weight(feat, proto, 1), w_ref)
    # 0.001-EMA of the stored vectors, skipped when the new vector sums to zero (:117-123)
    v = torch.randn(1, C, generator=g)
    TPR.Trainer.update_objective_SingleVector(me, "cup", v)
    assert relerr(O.ema_single_vector(proto.numpy(), v.numpy().reshape(-1)), me.objective_vectors["cup"].numpy()) < 1e-7
    before = me.objective_vectors["cup"].clone()
    TPR.Trainer.update_objective_SingleVector(me, "cup", torch.zeros(1, C))
    assert torch.equal(me.objective_vectors["cup"], before)
    # the product op (clr_ema_rows) is CUDA-only -- it is compared with this same ATen expression bit for bit on the GPU
    # (tests/test_gpu_widen.py) and refuses CPU tensors here
    from uda_clr_b200 import ops
    with pytest.raises(RuntimeError):
        ops.update_objective_single_vector(proto, v)
