"""Eager-PyTorch restatement of the reference's CLR path -- TEST INFRASTRUCTURE / CPU BASELINE.

The reference is pure Python over ATen and cannot travel to the GPU box (``/root/reference`` is
absent there), so this module restates it op for op -- same ATen call sequence, same
materialised intermediates, same dtype (fp32), generalised from the hard-coded K = 2 slices to
any K -- and is what ``bench.py --impl reference`` / ``cpu_baseline`` time on the host cores
(``kind: "port"``).  Backward is whatever autograd derives from that sequence, exactly as in the
reference's ``loss_all.backward()``.  ``tests/test_oracle_vs_reference.py`` checks it bit-for-bit
against the imported reference where the reference tree exists.

Never imported by the product package.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- A1
def _pool(feat: torch.Tensor, weights: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """One materialised product, one ``sum(dim=[0,2,3])``, one count and one in-place divide per
    row, in the reference's order (utils/Utils.py:114-130)."""
    prods = [feat * w for w in weights]
    sums = [torch.sum(p, dim=[0, 2, 3], keepdim=True) for p in prods]
    cnts = [torch.sum(w, dim=[0, 2, 3], keepdim=True) for w in weights]
    for s, n in zip(sums, cnts):
        s /= n
    return sums


def gen_prototype(pred: torch.Tensor, feat: torch.Tensor) -> Tuple[torch.Tensor, ...]:
    """A1 (utils/Utils.py:108-131) for any K: returns ``(obj_0..obj_{K-1}, bck_0..bck_{K-1})``,
    each ``[1,C,1,1]``."""
    K = pred.shape[1]
    obj = [pred[:, k:k + 1] for k in range(K)]
    bck = [1.0 - o for o in obj]
    return tuple(_pool(feat, obj + bck))


def gen_prototype_src_trg(pred_s, feat_s, pred_t, feat_t):
    """A3 (utils/Utils.py:132-158)."""
    return gen_prototype(torch.cat((pred_s, pred_t), 0), torch.cat((feat_s, feat_t), 0))


# ----------------------------------------------------------------------------- A2
def gen_prototype_retrify(oT_before, xt_feature, preds, features, T: int, stride: int,
                          read_dead_features: bool = True):
    """A2 (utils/Utils.py:159-225) for any K / C / H / W.

    ``features`` is the reference's ``[T*stride, C, H, W]`` staging buffer; the reference averages it
    (:169) and then uses only the result's size.  ``read_dead_features=True`` keeps that dead pass so
    the CPU baseline pays what the reference pays; pass ``features=None`` to skip it.
    """
    K = preds.shape[1]
    H, W = xt_feature.shape[2:]
    preds = preds.reshape(T, stride, K, preds.shape[2], preds.shape[3])
    preds1 = torch.sigmoid(preds)
    preds = torch.sigmoid(preds / 2.0)
    std_map = torch.std(preds, dim=0)
    prediction = torch.mean(preds1, dim=0)
    if features is not None and read_dead_features:
        torch.mean(features.reshape((T, stride) + tuple(features.shape[1:])), dim=0)
    prediction_small = F.interpolate(prediction, size=(H, W), mode='bilinear', align_corners=True)
    std_map_small = F.interpolate(std_map, size=(H, W), mode='bilinear', align_corners=True)

    pseudo = torch.sigmoid(oT_before).clone()
    pseudo[pseudo > 0.75] = 1.0
    pseudo[pseudo <= 0.75] = 0.0
    B = xt_feature.shape[0]
    weights_obj, weights_bck, masks = [], [], []
    for k in range(K):
        t_obj = pseudo[:, k:k + 1]
        t_bck = 1.0 - t_obj
        m_obj = torch.zeros([B, 1, H, W], dtype=xt_feature.dtype, device=xt_feature.device)
        m_bck = torch.zeros([B, 1, H, W], dtype=xt_feature.dtype, device=xt_feature.device)
        sel = std_map_small[:, k:k + 1] < 0.04
        m_obj[sel] = 1.0
        m_bck[sel] = 1.0
        masks.append(m_obj + m_bck)
        weights_obj.append((t_obj, m_obj, prediction_small[:, k:k + 1]))
        weights_bck.append((t_bck, m_bck, 1 - prediction_small[:, k:k + 1]))
    cents, cnts = [], []
    for (t, m, p) in weights_obj + weights_bck:
        cents.append(torch.sum(xt_feature * t * m * p, dim=[0, 2, 3], keepdim=True))
        cnts.append(torch.sum(m * t * p, dim=[0, 2, 3], keepdim=True))
    for c, n in zip(cents, cnts):
        c /= n
    return tuple(cents) + (std_map,) + tuple(masks)


# ----------------------------------------------------------------------------- A4 / A5
class PrototypeEMA:
    """The inline EMA of Trainer_prototype_full.py:335-355 (source) / :378-398 (target)."""

    def __init__(self, decay: float = 0.9):
        self.decay = decay
        self.first = True
        self.stored: Optional[List[torch.Tensor]] = None

    def update(self, current: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        if self.first:
            out = list(current)
            self.first = False
        else:
            d = self.decay
            out = [(1 - d) * s + d * c for s, c in zip(self.stored, current)]
        self.stored = [o.detach() for o in out]
        return out


def align_losses(src: Sequence[torch.Tensor], tgt: Sequence[torch.Tensor]):
    """``intra`` (Trainer_prototype_full.py:428-441) and ``inter`` (:443-444)."""
    mse = torch.nn.MSELoss()
    K = len(src) // 2
    intra = mse(src[0], tgt[0])
    for r in range(1, 2 * K):
        intra = intra + mse(src[r], tgt[r])
    # reference order: class 1 first, then class 0
    inter = None
    for k in reversed(range(K)):
        t = mse(src[k], src[K + k])
        inter = t if inter is None else inter + t
    return intra, inter


# ----------------------------------------------------------------------------- A9 / A10 (bytecode-only in the reference)
def disc_loss(xs: torch.Tensor, pred_oS: torch.Tensor, src: Sequence[torch.Tensor], margin: float = 0.01):
    """Trainer_prototype_mt.cpython-38.pyc L454-474, transcribed from the disassembly."""
    K = pred_oS.shape[1]
    total = None
    for k in range(K):
        d_obj = torch.mean(torch.pow(xs - src[k], 2), dim=1)
        d_bck = torch.mean(torch.pow(xs - src[K + k], 2), dim=1)
        l_obj = torch.mean(pred_oS[:, k] * F.relu(d_obj - d_bck + margin))
        d_bck2 = torch.mean(torch.pow(xs - src[K + k], 2), dim=1)
        d_obj2 = torch.mean(torch.pow(xs - src[k], 2), dim=1)
        l_bck = torch.mean((1 - pred_oS[:, k]) * F.relu(d_bck2 - d_obj2 + margin))
        total = l_obj + l_bck if total is None else total + l_obj + l_bck
    return total


def sigmoid_rampup(current, rampup_length):
    if rampup_length == 0:
        return 1.0
    import numpy as np
    current = np.clip(current, 0.0, rampup_length)
    phase = 1.0 - current / rampup_length
    return float(np.exp(-5.0 * phase * phase))


def cons_loss(oT: torch.Tensor, oT_aug: torch.Tensor, masks: Sequence[torch.Tensor], epoch: float,
              aug_weight: float = 1.0):
    """Trainer_prototype_mt.cpython-38.pyc L502-561 (the photometric augmentation between L546-555 is a
    CPU image op outside the hot path; ``oT_aug`` is the model output on the augmented view)."""
    import numpy as np
    prediction_ot = torch.sigmoid(oT)
    prediction_copy = prediction_ot.clone()
    threshold = (0.85 + 0.25 * sigmoid_rampup(epoch, 200)) * np.log(2)
    for k in range(oT.shape[1]):
        prediction_copy[:, k] = (prediction_ot[:, k] > threshold).long()
    y = prediction_copy.detach()
    mask = torch.cat(tuple(masks), dim=1)
    loss_aug = torch.nn.BCELoss(reduction='none')(torch.sigmoid(oT_aug), y)
    mask = F.interpolate(mask, size=loss_aug.size()[2:], mode='nearest')
    return torch.sum(mask * loss_aug) / torch.sum(mask) * aug_weight


# ----------------------------------------------------------------------------- A6 / A7 / A8
def bmm_pool(mask: torch.Tensor, feat: torch.Tensor) -> torch.Tensor:
    """A6 (Trainer_prototype.py:364-383): ``mean_b( bmm(m[b,1,HW], X[b,HW,C]) / (sum m + 1) )`` -> ``[1,C]``."""
    b, C, h, w = feat.size()
    key = feat.view(b, C, -1).permute(0, 2, 1)
    query = mask.reshape(b, 1, -1).to(feat.dtype)
    proto = torch.bmm(query, key)
    proto = proto / (torch.sum(query, dim=2, keepdim=True) + 1)
    return torch.mean(proto, dim=0)


def feat_prototype_distance(feat: torch.Tensor, prototype: torch.Tensor, class_numbers: int = 1):
    """A8 (Trainer_prototype.py:98-104)."""
    N, C, H, W = feat.shape
    out = -torch.ones((N, class_numbers, H, W), dtype=feat.dtype, device=feat.device)
    for i in range(class_numbers):
        out[:, i, :, :] = torch.norm(prototype.reshape(-1, 1, 1).expand(-1, H, W) - feat, 2, dim=1)
    return out


def distance_weight(feat, prototype, class_num: int = 1):
    """A8 (Trainer_prototype.py:106-116)."""
    d = feat_prototype_distance(feat, prototype, class_num)
    return (d - d.min()) / (d.max() - d.min())


def cosine_weight(feat, class_num, prototype):
    """A8 (utils/Utils.py:86-88)."""
    return torch.cosine_similarity(prototype, feat, dim=1).unsqueeze(1)


def adaptation_factor(m):
    return 1.0 / (1.0 + math.exp(-0.8 * (m + 1))) - 0.3


# ----------------------------------------------------------------------------- step glue (Trainer_prototype_full.py:292-294, 452)
def seg_loss(oS, boundaryS, target_map, target_boundary):
    """``loss_seg1 + loss_seg2`` exactly as the trainer forms them (:18-19, :292-294)."""
    loss = torch.nn.BCELoss()(torch.sigmoid(oS), target_map)
    if boundaryS is not None:
        loss = loss + torch.nn.MSELoss()(torch.sigmoid(boundaryS), target_boundary)
    return loss


def uncertainty_map(o, smooth: float = 1e-7):
    """``-1.0 * torch.sigmoid(oT) * torch.log(torch.sigmoid(oT) + smooth)`` (:452)."""
    return -1.0 * torch.sigmoid(o) * torch.log(torch.sigmoid(o) + smooth)


# ----------------------------------------------------------------------------- one whole CLR step, the way the trainer runs it
# ----------------------------------------------------------------------------- 8(f) rank 4: TransNorm
def trans_norm(x, weight, bias, rm_s, rv_s, rm_t, rv_t, training: bool, factor: float = 0.1, eps: float = 1e-5):
    """The ATen sequence of the reference's TransNorm forward (networks/sync_batchnorm/batchnorm.py:451-521): two
    ``F.batch_norm`` calls + ``cat`` + two transposed copies + four reductions (training), or one ``F.batch_norm`` with
    the target estimates (eval); ``alpha`` detached.  The running estimates are updated in place by ``F.batch_norm``."""
    C = x.shape[1]
    if training:
        h = x.size()[0] // 2
        src, tgt = x[:h], x[h:]
        z = torch.cat((F.batch_norm(src, rm_s, rv_s, weight, bias, True, factor, eps),
                       F.batch_norm(tgt, rm_t, rv_t, weight, bias, True, factor, eps)), dim=0)
        if x.dim() == 4:
            src = src.permute(0, 2, 3, 1).contiguous().view(-1, C)
            tgt = tgt.permute(0, 2, 3, 1).contiguous().view(-1, C)
        m_s, v_s = torch.mean(src, dim=0), torch.var(src, dim=0)
        m_t, v_t = torch.mean(tgt, dim=0), torch.var(tgt, dim=0)
    else:
        z = F.batch_norm(x, rm_t, rv_t, weight, bias, False, factor, eps)
        m_s, v_s, m_t, v_t = rm_s, rv_s, rm_t, rv_t
    dis = torch.abs(m_s / torch.sqrt(v_s + eps) - m_t / torch.sqrt(v_t + eps))
    prob = 1.0 / (1.0 + dis)
    alpha = C * prob / sum(prob)
    alpha = alpha.view(1, C, 1, 1) if x.dim() == 4 else alpha.view(1, C)
    return z * (1 + alpha.detach())


class ClrStepPort:
    """The CLR block of one training step (Trainer_prototype_full.py:328-449 + the two bytecode-only
    losses), fwd + ``backward()``, on whatever device the tensors live on."""

    def __init__(self, decay=0.9, pro_weight=0.1, src_reg_weight=1.0, aug_weight=1.0, margin=0.01,
                 retrify=True, use_disc=True, use_cons=True, backprop_aug=True):
        self.ema_s = PrototypeEMA(decay)
        self.ema_t = PrototypeEMA(decay)
        self.pro_weight, self.src_reg_weight, self.aug_weight = pro_weight, src_reg_weight, aug_weight
        self.margin = margin
        self.retrify, self.use_disc, self.use_cons, self.backprop_aug = retrify, use_disc, use_cons, backprop_aug

    def step(self, xs, ys, xt, oT_before, preds=None, features=None, T=8, oT=None, oT_aug=None, epoch=0.0):
        B = xt.shape[0]
        cur_s = gen_prototype(ys, xs)
        Ps = self.ema_s.update(cur_s)
        K = ys.shape[1]
        masks = None
        if self.retrify:
            out = gen_prototype_retrify(oT_before, xt, preds, features, T, B)
            cur_t, masks = out[:2 * K], out[2 * K + 1:]
            std_map = out[2 * K]
        else:
            cur_t = gen_prototype(torch.sigmoid(oT_before), xt)
        Pt = self.ema_t.update(cur_t)
        intra, inter = align_losses(Ps, Pt)
        total = self.pro_weight * intra
        res = dict(intra=intra.detach(), inter=inter.detach())
        if self.retrify:
            res["masks"] = [m.detach() for m in masks]
            res["std_map"] = std_map.detach()
        if self.use_disc:
            l_disc = disc_loss(xs, ys, Ps, self.margin)
            total = total + self.src_reg_weight * l_disc
            res["disc"] = l_disc.detach()
        if self.use_cons and masks is not None and oT_aug is not None:
            l_aug = cons_loss(oT, oT_aug, masks, epoch, self.aug_weight)
            res["aug"] = l_aug.detach()
            if self.backprop_aug:
                total = total + l_aug
        total.backward()
        res["total"] = total.detach()
        res["Ps"] = [p.detach() for p in Ps]
        res["Pt"] = [p.detach() for p in Pt]
        return res
