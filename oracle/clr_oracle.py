"""CPU oracle for the CLR hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A numpy float64 closed-form restatement of every function SURVEY.md §8(a) puts
on the hot path of fengweie/UDA_CLR, forward *and* backward.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package; the product (``uda_clr_b200``) never
does and has no CPU fallback.

Parity pin: the reference ships no golden vectors or tests, so
this oracle is pinned against the *reference itself*, imported unmodified from
``/root/reference`` (``oracle/ref_import.py``): ``tests/test_oracle_vs_reference.py``
runs both side by side where the reference tree is present, and
``tests/golden/*.npz`` (written by ``tests/golden/make_golden.py`` from the
reference's own outputs) travel to boxes where it is not.  The two
bytecode-only losses (A9, A10) have no importable source: they are restated
from the disassembly (``tools/pyc38_dis.py``) and pinned against a line-by-line
torch transcription of that bytecode in ``oracle/clr_torch_port.py``.

Integer-valued decisions (thresholded pseudo-labels, uncertainty masks) are
taken in float32 exactly as ATen takes them (the Python scalar is cast to the
tensor dtype before the compare); all sums, means and gradients are float64.

Row order of every ``[2K, C]`` prototype matrix: ``obj_0 .. obj_{K-1}, bck_0 ..
bck_{K-1}`` -- for K = 2 that is the reference's return order
``(c0_obj, c1_obj, c0_bck, c1_bck)`` (utils/Utils.py:131).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np

F32 = np.float32
F64 = np.float64


# --------------------------------------------------------------------------- helpers
def sigmoid_f32(x) -> np.ndarray:
    """fp32 logistic ``1 / (1 + exp(-x))`` -- the form ATen uses (UnarySpecialOpsKernel / sigmoid)."""
    x = np.asarray(x, dtype=F32)
    with np.errstate(over="ignore"):
        return (F32(1.0) / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


def adaptation_factor(m: float) -> float:
    """utils/Utils.py:104-107 and Trainer_prototype.py:240-243."""
    return 1.0 / (1.0 + math.exp(-0.8 * (m + 1))) - 0.3


def sigmoid_rampup(current: float, rampup_length: float) -> float:
    """Trainer_prototype_mt bytecode L24-31 (same as utils/Utils.py sigmoid_rampup)."""
    if rampup_length == 0:
        return 1.0
    current = float(np.clip(current, 0.0, rampup_length))
    phase = 1.0 - current / rampup_length
    return float(np.exp(-5.0 * phase * phase))


def consistency_threshold(epoch: float) -> float:
    """Trainer_prototype_mt bytecode L512: ``(0.85 + 0.25*sigmoid_rampup(epoch, 200)) * ln 2``."""
    return (0.85 + 0.25 * sigmoid_rampup(epoch, 200)) * float(np.log(2))


# --------------------------------------------------------------------------- A1: masked pooling
def weights_complement(pred: np.ndarray) -> np.ndarray:
    """``[B,K,H,W]`` -> ``[B,2K,H,W]``: obj = pred, bck = 1 - pred (fp32 subtract, utils/Utils.py:109-112)."""
    pred = np.asarray(pred, dtype=F32)
    return np.concatenate([pred, (F32(1.0) - pred).astype(F32)], axis=1)


def pool_sums(feat: np.ndarray, w: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """``S[r,c] = sum_{b,p} feat[b,c,p] * w[b,r,p]``, ``N[r] = sum w[b,r,p]`` (utils/Utils.py:114-126)."""
    B, C = feat.shape[:2]
    R = w.shape[1]
    x = np.asarray(feat, dtype=F64).reshape(B, C, -1)
    ww = np.asarray(w, dtype=F64).reshape(B, R, -1)
    S = np.einsum("bcp,brp->rc", x, ww)
    N = ww.sum(axis=(0, 2))
    return S, N


def prototypes_from_sums(S: np.ndarray, N: np.ndarray) -> np.ndarray:
    """``mu_r = S_r / N_r`` with the reference's 0/0 -> NaN behaviour (utils/Utils.py:127-130)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return S / N[:, None]


def gen_prototype(pred: np.ndarray, feat: np.ndarray) -> np.ndarray:
    """A1, utils/Utils.py:108-131.  Returns ``[2K, C]`` float64."""
    S, N = pool_sums(feat, weights_complement(pred))
    return prototypes_from_sums(S, N)


def pool_backward(feat: np.ndarray, w: np.ndarray, g: np.ndarray,
                  S: Optional[np.ndarray] = None, N: Optional[np.ndarray] = None
                  ) -> Tuple[np.ndarray, np.ndarray]:
    """Closed-form adjoint of ``mu = S/N``.

    ``g[r,c] = dL/dmu_r[c]``.  Returns ``(dL/dfeat [B,C,H,W], dL/dw [B,R,H,W])``.
    ``S``/``N`` may be *global* sums (multi-GPU: the local shard divides by the global N).
    """
    B, C = feat.shape[:2]
    R = w.shape[1]
    if S is None or N is None:
        S, N = pool_sums(feat, w)
    mu = S / N[:, None]
    G = np.asarray(g, dtype=F64) / N[:, None]                      # [R,C]
    x = np.asarray(feat, dtype=F64).reshape(B, C, -1)
    ww = np.asarray(w, dtype=F64).reshape(B, R, -1)
    gx = np.einsum("rc,brp->bcp", G, ww).reshape(feat.shape)
    gw = (np.einsum("rc,bcp->brp", G, x) - (G * mu).sum(axis=1)[None, :, None]).reshape(w.shape)
    return gx, gw


def gen_prototype_backward(pred: np.ndarray, feat: np.ndarray, g: np.ndarray
                           ) -> Tuple[np.ndarray, np.ndarray]:
    """Adjoint of :func:`gen_prototype`: ``(dL/dfeat, dL/dpred)``; ``dpred_k = dw_obj,k - dw_bck,k``."""
    K = pred.shape[1]
    gx, gw = pool_backward(feat, weights_complement(pred), g)
    return gx, gw[:, :K] - gw[:, K:]


def gen_prototype_src_trg(pred_s, feat_s, pred_t, feat_t) -> np.ndarray:
    """A3, utils/Utils.py:132-158: A1 on the concatenated source+target batch."""
    Ss, Ns = pool_sums(feat_s, weights_complement(pred_s))
    St, Nt = pool_sums(feat_t, weights_complement(pred_t))
    return prototypes_from_sums(Ss + St, Ns + Nt)


# --------------------------------------------------------------------------- A2: MC statistics + retrify
def mc_statistics(preds: np.ndarray, T: int, stride: int) -> Tuple[np.ndarray, np.ndarray]:
    """utils/Utils.py:161-168.  ``preds [T*stride,K,Hi,Wi]`` logits ->
    ``(std_map, prediction)``, both ``[stride,K,Hi,Wi]`` float64:
    unbiased std over T of ``sigmoid(p/2)`` and mean over T of ``sigmoid(p)``; sigmoids in fp32."""
    p = np.asarray(preds, dtype=F32)
    p = p.reshape((T, stride) + p.shape[1:])
    s_half = sigmoid_f32(p / F32(2.0)).astype(F64)
    s_full = sigmoid_f32(p).astype(F64)
    std_map = s_half.std(axis=0, ddof=1)
    prediction = s_full.mean(axis=0)
    return std_map, prediction


def bilinear_align_corners(x: np.ndarray, H: int, W: int) -> np.ndarray:
    """``F.interpolate(x, size=(H,W), mode='bilinear', align_corners=True)`` (utils/Utils.py:170-171).

    Source indices and lambdas are derived in float32 exactly as ATen does
    (``area_pixel_compute_scale``: ``(in-1)/(out-1)`` as float; source index = ``scale * dst``),
    so the *neighbour selection* matches; the blend itself is float64.
    """
    x = np.asarray(x)
    Hi, Wi = x.shape[-2:]

    def axis(n_in, n_out):
        scale = F32(n_in - 1) / F32(n_out - 1) if n_out > 1 else F32(0.0)
        src = (scale * np.arange(n_out, dtype=F32)).astype(F32)
        i0 = src.astype(np.int64)
        i1 = i0 + (i0 < n_in - 1)
        l1 = (src - i0.astype(F32)).astype(F32)
        l0 = (F32(1.0) - l1).astype(F32)
        return i0, i1, l0.astype(F64), l1.astype(F64)

    h0, h1, hl0, hl1 = axis(Hi, H)
    w0, w1, wl0, wl1 = axis(Wi, W)
    xd = x.astype(F64)
    top = xd[..., h0, :][..., :, w0] * wl0 + xd[..., h0, :][..., :, w1] * wl1
    bot = xd[..., h1, :][..., :, w0] * wl0 + xd[..., h1, :][..., :, w1] * wl1
    return top * hl0[:, None] + bot * hl1[:, None]


def retrify_weights(oT_before: np.ndarray, pred_small: np.ndarray, std_small: np.ndarray,
                    pseudo_thr: float = 0.75, std_thr: float = 0.04
                    ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """utils/Utils.py:173-223.  Returns ``(w [B,2K,H,W] float32, pseudo [B,K,H,W] uint8, mask [B,K,H,W] uint8)``.

    ``pseudo_k = sigmoid(oT)_k > 0.75``; ``mask_k = std_small_k < 0.04`` (both fp32 compares);
    ``w_obj,k = pseudo_k*mask_k*pred_small_k``; ``w_bck,k = (1-pseudo_k)*mask_k*(1-pred_small_k)``.
    """
    pseudo = sigmoid_f32(oT_before) > F32(pseudo_thr)
    mask = np.asarray(std_small, dtype=F32) < F32(std_thr)
    ps = np.asarray(pred_small, dtype=F32)
    w_obj = np.where(pseudo & mask, ps, F32(0.0)).astype(F32)
    w_bck = np.where((~pseudo) & mask, (F32(1.0) - ps).astype(F32), F32(0.0)).astype(F32)
    return np.concatenate([w_obj, w_bck], axis=1), pseudo.astype(np.uint8), mask.astype(np.uint8)


def gen_prototype_retrify(oT_before, xt_feature, preds, T: int, stride: int) -> Dict[str, np.ndarray]:
    """A2, utils/Utils.py:159-225 (the ``features`` argument is dead there: only its ``.size()[2:]`` is
    read at :169-171, which equals ``xt_feature``'s spatial size).

    Returns dict: ``protos [2K,C]``, ``std_map [B,K,Hi,Wi]``, ``mask_0``/``mask_1`` ``[B,1,H,W]`` in {0,2}
    (K = 2 naming; ``masks [B,K,H,W]`` in general), plus the intermediates the tests inspect.
    """
    H, W = xt_feature.shape[-2:]
    std_map, prediction = mc_statistics(preds, T, stride)
    pred_small = bilinear_align_corners(prediction.astype(F32), H, W).astype(F32)
    std_small = bilinear_align_corners(std_map.astype(F32), H, W).astype(F32)
    w, pseudo, mask = retrify_weights(oT_before, pred_small, std_small)
    S, N = pool_sums(xt_feature, w)
    out = dict(protos=prototypes_from_sums(S, N), S=S, N=N, std_map=std_map, pred_small=pred_small,
               std_small=std_small, w=w, pseudo=pseudo, masks=(2.0 * mask).astype(F32))
    out["mask_0"] = out["masks"][:, 0:1]
    if mask.shape[1] > 1:
        out["mask_1"] = out["masks"][:, 1:2]
    return out


# --------------------------------------------------------------------------- A4/A5: EMA + alignment
def ema_update(stored: Optional[np.ndarray], cur: np.ndarray, decay: float) -> Tuple[np.ndarray, float]:
    """Trainer_prototype_full.py:335-355 / :378-398.  First call copies; later
    ``P = (1-decay)*stored + decay*cur``.  Returns ``(P, dP/dcur)``."""
    if stored is None:
        return np.array(cur, dtype=F64), 1.0
    return (1.0 - decay) * np.asarray(stored, F64) + decay * np.asarray(cur, F64), float(decay)


def align_losses(Ps: np.ndarray, Pt: np.ndarray) -> Tuple[float, float]:
    """Trainer_prototype_full.py:428-444: ``intra = sum_r mean_c (Ps_r-Pt_r)^2``;
    ``inter = sum_k mean_c (Ps_obj,k - Ps_bck,k)^2``."""
    K = Ps.shape[0] // 2
    intra = float(((Ps - Pt) ** 2).mean(axis=1).sum())
    inter = float(((Ps[:K] - Ps[K:]) ** 2).mean(axis=1).sum())
    return intra, inter


def align_grads(Ps: np.ndarray, Pt: np.ndarray, w_intra: float, w_inter: float
                ) -> Tuple[np.ndarray, np.ndarray]:
    """d(w_intra*intra + w_inter*inter)/d(Ps, Pt)."""
    K, C = Ps.shape[0] // 2, Ps.shape[1]
    gs = w_intra * 2.0 * (Ps - Pt) / C
    gt = -gs.copy()
    d = w_inter * 2.0 * (Ps[:K] - Ps[K:]) / C
    gs[:K] += d
    gs[K:] -= d
    return gs, gt


# --------------------------------------------------------------------------- A9: discriminative hinge
def disc_loss(xs: np.ndarray, y: np.ndarray, P: np.ndarray, margin: float = 0.01,
              npx: Optional[int] = None) -> Tuple[float, Dict[str, np.ndarray]]:
    """Trainer_prototype_mt bytecode L454-474:
    ``loss = sum_k mean_{b,p}( y_k relu(d_obj,k - d_bck,k + m) ) + mean( (1-y_k) relu(d_bck,k - d_obj,k + m) )``
    with ``d_r(b,p) = mean_c (x[b,c,p] - P_r[c])^2``.  ``npx`` overrides the mean's pixel count
    (multi-GPU shard: the global ``G*B*H*W``)."""
    B, C = xs.shape[:2]
    K = y.shape[1]
    x = np.asarray(xs, F64).reshape(B, C, -1)
    yy = np.asarray(y, F64).reshape(B, K, -1)
    P = np.asarray(P, F64)
    d = ((x[:, None] - P[None, :, :, None]) ** 2).mean(axis=2)          # [B,2K,P]
    delta = d[:, :K] - d[:, K:]                                          # [B,K,P]
    h_obj = np.maximum(delta + margin, 0.0)
    h_bck = np.maximum(-delta + margin, 0.0)
    npx = B * x.shape[2] if npx is None else npx
    loss = float(((yy * h_obj).sum() + ((1.0 - yy) * h_bck).sum()) / npx)
    coef = yy * (delta + margin > 0) - (1.0 - yy) * (-delta + margin > 0)  # dL*npx / d delta
    return loss, dict(delta=delta.reshape(y.shape), coef=coef.reshape(y.shape))


def disc_grads(xs: np.ndarray, y: np.ndarray, P: np.ndarray, margin: float = 0.01,
               npx: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Closed-form ``(dL/dxs [B,C,H,W], dL/dP [2K,C])`` of :func:`disc_loss` (SURVEY.md §8(a) A9 identity:
    ``d(d_obj-d_bck)/dx = 2 (P_bck-P_obj)/C`` independent of x)."""
    B, C = xs.shape[:2]
    K = y.shape[1]
    _, aux = disc_loss(xs, y, P, margin)
    coef = aux["coef"].reshape(B, K, -1)
    x = np.asarray(xs, F64).reshape(B, C, -1)
    P = np.asarray(P, F64)
    npx = B * x.shape[2] if npx is None else npx
    D = P[:K] - P[K:]                                                    # [K,C]
    gx = np.einsum("bkp,kc->bcp", coef, -2.0 * D / C) / npx
    n_k = coef.sum(axis=(0, 2))                                          # [K]
    A = np.einsum("bkp,bcp->kc", coef, x)                                # [K,C]
    g_obj = (2.0 / (C * npx)) * (n_k[:, None] * P[:K] - A)
    g_bck = -(2.0 / (C * npx)) * (n_k[:, None] * P[K:] - A)
    return gx.reshape(xs.shape), np.concatenate([g_obj, g_bck], axis=0)


# --------------------------------------------------------------------------- A10: augmented consistency
def nearest_upsample(x: np.ndarray, Ho: int, Wo: int) -> np.ndarray:
    """``F.interpolate(mode='nearest')``: ``src = min(floor(dst * (in/out as fp32)), in-1)``."""
    Hi, Wi = x.shape[-2:]

    def idx(n_in, n_out):
        scale = F32(n_in) / F32(n_out)
        return np.minimum((np.arange(n_out, dtype=F32) * scale).astype(np.int64), n_in - 1)

    return x[..., idx(Hi, Ho), :][..., :, idx(Wi, Wo)]


def cons_loss(oT: np.ndarray, oT_aug: np.ndarray, masks: np.ndarray, threshold: float,
              aug_weight: float = 1.0) -> Tuple[float, np.ndarray, Dict[str, np.ndarray]]:
    """Trainer_prototype_mt bytecode L502-561.

    ``y = sigmoid(oT) > threshold`` (fp32 compare); ``l = BCELoss(none)(sigmoid(oT_aug), y)`` with the
    log clamp at -100; ``m = nearest_up(masks)``; ``loss = sum(m*l)/sum(m) * aug_weight``.
    Returns ``(loss, dloss/doT_aug, aux)``; the backward mirrors ATen's
    ``binary_cross_entropy_backward`` (denominator clamped at 1e-12) chained with sigmoid'.
    """
    Hi, Wi = oT.shape[-2:]
    y = (sigmoid_f32(oT) > F32(threshold)).astype(F64)
    q32 = sigmoid_f32(oT_aug)
    q = q32.astype(F64)
    one_minus_q = (F32(1.0) - q32).astype(F64)
    with np.errstate(divide="ignore"):
        l = -(y * np.maximum(np.log(q), -100.0) + (1.0 - y) * np.maximum(np.log(one_minus_q), -100.0))
    m = nearest_upsample(np.asarray(masks, F64), Hi, Wi)
    msum = m.sum()
    with np.errstate(divide="ignore", invalid="ignore"):
        loss = float((m * l).sum() / msum * aug_weight)
        gl = m / msum * aug_weight
        gq = gl * (q - y) / np.maximum(one_minus_q * q, 1e-12)
    gz = gq * q * one_minus_q
    return loss, gz, dict(y=y.astype(np.uint8), l=l, m=m)


# --------------------------------------------------------------------------- A6/A7/A8: variant-A trainer pieces
def bmm_pool(mask: np.ndarray, feat: np.ndarray) -> np.ndarray:
    """A6, Trainer_prototype.py:364-383 / cal_prototype.py:156-175:
    ``proto = mean_b( (m_b . X_b) / (sum_p m_b + 1) )``.  ``mask [B,1,H,W]`` or ``[B,H,W]`` -> ``[C]``."""
    B, C = feat.shape[:2]
    x = np.asarray(feat, F64).reshape(B, C, -1)
    m = np.asarray(mask, F64).reshape(B, -1)
    s = np.einsum("bp,bcp->bc", m, x)
    return (s / (m.sum(axis=1, keepdims=True) + 1.0)).mean(axis=0)


def ema_single_vector(obj: np.ndarray, vec: np.ndarray, rate: float = 0.001) -> np.ndarray:
    """A7, Trainer_prototype.py:117-123: skipped when ``vec.sum() == 0``."""
    if float(np.asarray(vec, F32).sum()) == 0.0:
        return np.asarray(obj, F64)
    return np.asarray(obj, F64) * (1.0 - rate) + rate * np.asarray(vec, F64).reshape(-1)


def feat_prototype_distance(feat: np.ndarray, proto: np.ndarray) -> np.ndarray:
    """A8, Trainer_prototype.py:98-104: ``D[n,h,w] = || proto - feat[n,:,h,w] ||_2``."""
    x = np.asarray(feat, F64)
    p = np.asarray(proto, F64).reshape(1, -1, 1, 1)
    return np.sqrt(((p - x) ** 2).sum(axis=1))


def distance_weight(feat: np.ndarray, proto: np.ndarray) -> np.ndarray:
    """A8, Trainer_prototype.py:106-116: global min/max normalisation of the distance map."""
    d = feat_prototype_distance(feat, proto)
    return (d - d.min()) / (d.max() - d.min())


def cosine_weight(feat: np.ndarray, proto: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    """A8, utils/Utils.py:86-88: ``cosine_similarity(prototype, feat, dim=1).unsqueeze(1)``
    (ATen: ``x.y / (max(|x|,eps) * max(|y|,eps))``)."""
    x = np.asarray(feat, F64)
    p = np.asarray(proto, F64).reshape(1, -1, 1, 1)
    num = (x * p).sum(axis=1)
    den = np.maximum(np.sqrt((x * x).sum(axis=1)), eps) * max(float(np.sqrt((p * p).sum())), eps)
    return (num / den)[:, None]


# --------------------------------------------------------------------------- 8(f) rank 2: step glue at image resolution
def seg_loss(oS: np.ndarray, boundaryS: Optional[np.ndarray], target_map: np.ndarray,
             target_boundary: Optional[np.ndarray]):
    """Trainer_prototype_full.py:292-294: ``BCELoss(sigmoid(oS), map) + MSELoss(sigmoid(bS), boundary)`` (means) and
    the gradients w.r.t. the two logit maps.  fp64, log terms clamped at -100 like ATen's BCELoss."""
    o = np.asarray(oS, F64)
    y = np.asarray(target_map, F64)
    q = 1.0 / (1.0 + np.exp(-o))
    with np.errstate(divide="ignore"):
        lq = np.maximum(np.log(q), -100.0)
        l1q = np.maximum(np.log1p(-q), -100.0)
    bce = float(np.mean(-(y * lq + (1.0 - y) * l1q)))
    g_o = (q - y) / o.size
    mse, g_b = 0.0, None
    if boundaryS is not None:
        b = np.asarray(boundaryS, F64)
        t = np.asarray(target_boundary, F64)
        qb = 1.0 / (1.0 + np.exp(-b))
        mse = float(np.mean((qb - t) ** 2))
        g_b = 2.0 * (qb - t) * qb * (1.0 - qb) / b.size
    return bce + mse, dict(bce=bce, mse=mse, g_oS=g_o, g_boundaryS=g_b)


def uncertainty_map(o: np.ndarray, smooth: float = 1e-7):
    """Trainer_prototype_full.py:452: ``-sigmoid(o) * log(sigmoid(o) + smooth)`` and d/do (per unit upstream gradient)."""
    x = np.asarray(o, F64)
    q = 1.0 / (1.0 + np.exp(-x))
    u = -q * np.log(q + smooth)
    du = -q * (1.0 - q) * (np.log(q + smooth) + q / (q + smooth))
    return u, du


# --------------------------------------------------------------------------- the fused step (A1/A2 + A4 + A5 + A9 + A10)
# --------------------------------------------------------------------------- 8(f) rank 4: TransNorm
def _tn_flat(x: np.ndarray) -> np.ndarray:
    """[B, C, ...] -> [B, C, HW] float64."""
    x = np.asarray(x, F64)
    return x.reshape(x.shape[0], x.shape[1], -1)


def transnorm_alpha(mean: np.ndarray, var: np.ndarray, eps: float) -> np.ndarray:
    """``alpha = C * prob / sum(prob)``, ``prob = 1 / (1 + |mu_s/sqrt(var_s+eps) - mu_t/sqrt(var_t+eps)|)``
    (networks/sync_batchnorm/batchnorm.py:481-487); ``mean``/``var`` are ``[2, C]`` (source, target)."""
    dis = np.abs(mean[0] / np.sqrt(var[0] + eps) - mean[1] / np.sqrt(var[1] + eps))
    prob = 1.0 / (1.0 + dis)
    return mean.shape[1] * prob / prob.sum()


def transnorm_train(x: np.ndarray, weight: Optional[np.ndarray], bias: Optional[np.ndarray], eps: float = 1e-5
                    ) -> Dict[str, np.ndarray]:
    """Training forward of the reference's TransNorm (batchnorm.py:451-493): each half of the batch normalised with its
    own statistics (biased variance, F.batch_norm), shared affine, times ``1 + alpha`` (unbiased variances, :474-476).
    Returns ``y`` (shape of ``x``), ``alpha [C]``, ``mean``, ``var_b``, ``var_u`` (``[2, C]``)."""
    xf = _tn_flat(x)
    B, C, _ = xf.shape
    h = B // 2
    gamma = np.ones(C) if weight is None else np.asarray(weight, F64)
    beta = np.zeros(C) if bias is None else np.asarray(bias, F64)
    mean, var_b, var_u = np.zeros((2, C)), np.zeros((2, C)), np.zeros((2, C))
    z = np.empty_like(xf)
    for d, sl in enumerate((slice(0, h), slice(h, B))):
        part = xf[sl]
        n = part.shape[0] * part.shape[2]
        mean[d] = part.mean(axis=(0, 2))
        ss = ((part - mean[d][None, :, None]) ** 2).sum(axis=(0, 2))
        var_b[d], var_u[d] = ss / n, ss / (n - 1)
        z[sl] = (part - mean[d][None, :, None]) / np.sqrt(var_b[d] + eps)[None, :, None] * gamma[None, :, None] \
            + beta[None, :, None]
    alpha = transnorm_alpha(mean, var_u, eps)
    y = z * (1.0 + alpha)[None, :, None]
    return {"y": y.reshape(np.shape(x)), "alpha": alpha, "mean": mean, "var_b": var_b, "var_u": var_u}


def transnorm_running(running: np.ndarray, stat: np.ndarray, factor: float) -> np.ndarray:
    """``running = (1 - factor) running + factor stat`` (F.batch_norm's update; the variance fed in is the unbiased one)."""
    return (1.0 - factor) * np.asarray(running, F64) + factor * np.asarray(stat, F64)


def transnorm_train_backward(x: np.ndarray, weight: Optional[np.ndarray], gy: np.ndarray, eps: float = 1e-5):
    """Adjoint of :func:`transnorm_train` with ``alpha`` held constant (``alpha.detach()``, batchnorm.py:493):
    ``(gx, gweight, gbias)``."""
    xf, gf = _tn_flat(x), _tn_flat(gy)
    B, C, _ = xf.shape
    h = B // 2
    gamma = np.ones(C) if weight is None else np.asarray(weight, F64)
    fw = transnorm_train(x, weight, None, eps)
    q = 1.0 + fw["alpha"]
    gx = np.empty_like(xf)
    gw, gb = np.zeros(C), np.zeros(C)
    for d, sl in enumerate((slice(0, h), slice(h, B))):
        part, dz = xf[sl], gf[sl] * q[None, :, None]
        n = part.shape[0] * part.shape[2]
        rstd = 1.0 / np.sqrt(fw["var_b"][d] + eps)
        xhat = (part - fw["mean"][d][None, :, None]) * rstd[None, :, None]
        s1, s2 = dz.sum(axis=(0, 2)), (dz * xhat).sum(axis=(0, 2))
        gx[sl] = (gamma * rstd)[None, :, None] * (dz - (s1 / n)[None, :, None] - xhat * (s2 / n)[None, :, None])
        gw += s2
        gb += s1
    return gx.reshape(np.shape(x)), gw, gb


def transnorm_eval(x: np.ndarray, weight, bias, rm_s, rv_s, rm_t, rv_t, eps: float = 1e-5) -> np.ndarray:
    """Eval forward (batchnorm.py:494-521): target running estimates for every sample, ``alpha`` from both sets."""
    xf = _tn_flat(x)
    C = xf.shape[1]
    gamma = np.ones(C) if weight is None else np.asarray(weight, F64)
    beta = np.zeros(C) if bias is None else np.asarray(bias, F64)
    mean = np.stack([np.asarray(rm_s, F64), np.asarray(rm_t, F64)])
    var = np.stack([np.asarray(rv_s, F64), np.asarray(rv_t, F64)])
    alpha = transnorm_alpha(mean, var, eps)
    z = (xf - mean[1][None, :, None]) / np.sqrt(var[1] + eps)[None, :, None] * gamma[None, :, None] + beta[None, :, None]
    return (z * (1.0 + alpha)[None, :, None]).reshape(np.shape(x))


def clr_step(xs, ys, xt, wt, *, stored_s=None, stored_t=None, decay: float = 0.9,
             w_intra: float = 0.1, w_inter: float = 0.0, w_disc: float = 0.0, margin: float = 0.01,
             cons: Optional[dict] = None, w_aug: float = 0.0,
             global_sums: Optional[dict] = None, grad_scale: float = 1.0) -> Dict[str, object]:
    """One CLR step, forward and backward, in closed form.

    ``ys [B,K,H,W]`` are the source weights in complement format (hard labels).  ``wt`` is either
    ``[B,K,H,W]`` (complement: ``gen_prototype(sigmoid(oT_before), xt)``, Trainer_prototype_full.py:375-377)
    or ``[B,2K,H,W]`` (explicit retrify weights, :369-373).  ``cons`` = dict(oT, oT_aug, masks, threshold,
    aug_weight) enables A10.  ``global_sums`` = dict(Ss, Ns, St, Nt) substitutes all-reduced sums
    (multi-GPU shard); ``grad_scale`` is the DDP factor G.  (The sharded discriminative
    term needs a second exchange and is composed by the caller from :func:`disc_loss`/:func:`disc_grads`
    with ``npx`` = the global pixel count.)

    total = w_intra*intra + w_inter*inter + w_disc*loss_disc + w_aug*loss_aug.
    """
    K = ys.shape[1]
    ws = weights_complement(ys)
    wt_full = weights_complement(wt) if wt.shape[1] == K else np.asarray(wt)
    Ss, Ns = pool_sums(xs, ws)
    St, Nt = pool_sums(xt, wt_full)
    if global_sums is not None:
        Ss, Ns, St, Nt = (global_sums[k] for k in ("Ss", "Ns", "St", "Nt"))
    cur_s, cur_t = prototypes_from_sums(Ss, Ns), prototypes_from_sums(St, Nt)
    Ps, ds = ema_update(stored_s, cur_s, decay)
    Pt, dt = ema_update(stored_t, cur_t, decay)
    intra, inter = align_losses(Ps, Pt)
    gPs, gPt = align_grads(Ps, Pt, w_intra, w_inter)
    out: Dict[str, object] = dict(cur_s=cur_s, cur_t=cur_t, Ps=Ps, Pt=Pt, intra=intra, inter=inter,
                                  Ss=Ss, Ns=Ns, St=St, Nt=Nt)
    gxs_direct = 0.0
    loss_disc = 0.0
    if w_disc != 0.0:
        loss_disc, _ = disc_loss(xs, ys, Ps, margin)
        gx_d, gP_d = disc_grads(xs, ys, Ps, margin)
        gxs_direct = w_disc * gx_d
        gPs = gPs + w_disc * gP_d
    out["loss_disc"] = loss_disc
    gxs, _ = pool_backward(xs, ws, ds * gPs, Ss, Ns)
    gxt, gwt = pool_backward(xt, wt_full, dt * gPt, St, Nt)
    out["gxs"] = grad_scale * (gxs + gxs_direct)
    out["gxt"] = grad_scale * gxt
    out["gwt"] = grad_scale * (gwt[:, :K] - gwt[:, K:] if wt.shape[1] == K else gwt)
    loss_aug = 0.0
    if cons is not None:
        loss_aug, gz, _ = cons_loss(cons["oT"], cons["oT_aug"], cons["masks"], cons["threshold"],
                                    cons.get("aug_weight", 1.0))
        out["g_oT_aug"] = grad_scale * w_aug * gz
    out["loss_aug"] = loss_aug
    out["total"] = w_intra * intra + w_inter * inter + w_disc * loss_disc + w_aug * loss_aug
    return out
