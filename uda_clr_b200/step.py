"""The fused CLR step: everything ``Trainer_prototype_full.train_epoch`` does between the model forward
and ``loss_all.backward()`` for the CLR terms (Trainer_prototype_full.py:328-449, plus the two losses
that exist only in ``Trainer_prototype_mt``'s bytecode), as ONE call.

    step = CLRStep(K=2, decay=0.9, pro_weight=0.1, src_reg_weight=1.0, use_disc=True, use_cons=True)
    out = step(xs_feature, pred_oS, xt_feature, oT_before=oT_before, preds=preds_trg, T=8, oT=oT, oT_aug=oT_aug)
    (loss_seg + loss_adv + out.total).backward()

The trainer's inline EMA / MSE block cannot be fused without touching the trainer, so this is an
additional entry point next to the drop-in functions of :mod:`uda_clr_b200.ops`; it owns the EMA state
(the ``self.sourcecentroid_*`` / ``self.targetcentroid_*`` attributes and the ``First*`` flags of the
reference trainer, Trainer_prototype_full.py:32-33, 341-344).
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import _lib
from . import dist as _dist
from ._lib import CLR_W_COMPLEMENT, CLR_W_EXPLICIT, StepArgs, check, ptr
from .ops import PSEUDO_THRESHOLD, STD_THRESHOLD, _require_cuda_f32, _stream


def sigmoid_rampup(current: float, rampup_length: float) -> float:
    """``sigmoid_rampup`` of the reference (Trainer_prototype_mt bytecode L24-31, utils/Utils.py:312+)."""
    if rampup_length == 0:
        return 1.0
    current = min(max(float(current), 0.0), float(rampup_length))
    phase = 1.0 - current / rampup_length
    return math.exp(-5.0 * phase * phase)


def consistency_threshold(epoch: float) -> float:
    """``(0.85 + 0.25*sigmoid_rampup(epoch, 200)) * ln 2`` (Trainer_prototype_mt bytecode L512)."""
    return (0.85 + 0.25 * sigmoid_rampup(epoch, 200)) * math.log(2.0)


@dataclass
class CLRStepOutput:
    total: torch.Tensor              # scalar, differentiable: w_intra*intra + w_inter*inter + w_disc*disc + w_aug*aug
    intra: torch.Tensor              # scalars below are detached views of one device buffer (no host sync)
    inter: torch.Tensor
    disc: torch.Tensor
    aug: torch.Tensor
    source_prototypes: List[torch.Tensor]   # 2K x [1,C,1,1], EMA'd, detached
    target_prototypes: List[torch.Tensor]
    std_map: Optional[torch.Tensor]         # [B,K,Hi,Wi] (retrify)
    masks: Optional[List[torch.Tensor]]     # K x [B,1,H,W] in {0,2} (retrify)
    error: Optional[torch.Tensor] = None    # device scalar, 0 = ok; 1 = a device-side wait of the step timed out (a peer
                                            # of the in-kernel exchange is gone, or a gate was missed): the step's losses
                                            # and gradients are NaN then and the EMA state was left untouched


class CLRStepError(RuntimeError):
    """A device-side wait of the fused step timed out (``losses[7]`` != 0)."""


class _Buffers:
    """All device buffers of one step, carved from a single allocation."""

    def __init__(self, a: "CLRStep", dev, B_s, B_t, C, H, W, K, Hi, Wi):
        R = 2 * K
        spec = [("packed1", 2 * R * (C + 1)), ("packed2", K * (C + 1) + 4), ("P_s", R * C), ("P_t", R * C),
                ("g_s", R * C), ("g_t", R * C), ("losses", 8)]
        if a.use_disc:
            spec += [("disc_vec", K * C), ("disc_beta", K), ("xtab", K * C), ("disc_coef", B_s * K * H * W)]
        if a.retrify:
            spec += [("std_map", B_t * K * Hi * Wi), ("pred_mean", B_t * K * Hi * Wi),
                     ("wt_retrify", B_t * R * H * W), ("masks", B_t * K * H * W)]
        total = sum((n + 63) // 64 * 64 for _, n in spec)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)     # (losses[7]: exchange-timeout flag, starts 0)
        off = 0
        for name, n in spec:
            setattr(self, name, self.flat[off:off + n])
            off += (n + 63) // 64 * 64


def _downsample_labels(lib, holder, st) -> None:
    full = holder["ys_full"]
    if full is not None:
        ys = holder["ys"]
        check(lib.clr_label_downsample(ptr(full), full.shape[0] * full.shape[1], full.shape[2], full.shape[3],
                                       ys.shape[2], ys.shape[3], ptr(ys), st), "clr_label_downsample")


class _ClrStepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, step, args_holder, xs, xt, oT_aug, wt_soft=None):
        lib = _lib.load()
        a: StepArgs = args_holder["args"]
        with torch.cuda.device(xs.device):
            st = _stream()              # the current stream of the tensors' device, not of whatever device is current
            _downsample_labels(lib, args_holder, st)
            if args_holder["peer"] is not None:
                a.seq = _dist.next_seq(args_holder["peer"])
                check(lib.clr_step_fwd(ctypes.byref(a), st), "clr_step_fwd (in-kernel exchange)")
            elif _dist.enabled():
                check(lib.clr_step_fwd_a(ctypes.byref(a), st), "clr_step_fwd_a")
                _dist.all_reduce_sums(args_holder["buf"].packed1)
                check(lib.clr_step_fwd_b(ctypes.byref(a), st), "clr_step_fwd_b")
                _dist.all_reduce_sums(args_holder["buf"].packed2)
                check(lib.clr_step_fwd_c(ctypes.byref(a), st), "clr_step_fwd_c")
            else:
                check(lib.clr_step_fwd(ctypes.byref(a), st), "clr_step_fwd")
        ctx.holder = args_holder
        ctx.xs_shape, ctx.xt_shape = tuple(xs.shape), tuple(xt.shape)
        ctx.aug_shape = None if oT_aug is None else tuple(oT_aug.shape)
        ctx.xt_saved = xt if (wt_soft is not None and ctx.needs_input_grad[5]) else None    # dL/dw needs the features again
        return args_holder["buf"].losses[4].clone()

    @staticmethod
    def backward(ctx, gup):
        lib = _lib.load()
        h = ctx.holder
        a: StepArgs = h["args"]
        dev = h["buf"].flat.device
        gxs = torch.empty(ctx.xs_shape, dtype=torch.float32, device=dev)
        gxt = torch.empty(ctx.xt_shape, dtype=torch.float32, device=dev)
        want_aug = ctx.aug_shape is not None and ctx.needs_input_grad[4] and a.use_cons and a.w_aug != 0.0
        g_aug = torch.empty(ctx.aug_shape, dtype=torch.float32, device=dev) if want_aug else None
        gup = gup.to(torch.float32).contiguous()
        a.gxs, a.gxt, a.g_oT_aug, a.gup = ptr(gxs), ptr(gxt), ptr(g_aug), ptr(gup)
        with torch.cuda.device(dev):
            check(lib.clr_step_bwd(ctypes.byref(a), _stream()), "clr_step_bwd")
        h["keep_bwd"] = (gxs, gxt, g_aug, gup)
        g_wt = None
        if ctx.xt_saved is not None:
            # soft target predictions (gen_prototype(sigmoid(oT_before), xt), Trainer_prototype_full.py:375-377):
            # dL/dw_r[b,p] = sum_c g_r[c]/N_r (x[b,c,p] - mu_r[c]); one more read of xt, chained into sigmoid by autograd
            from .ops import pool_backward_weights
            K, C = a.K, a.C
            R = 2 * K
            buf = h["buf"]
            g_t = buf.g_t.view(R, C)
            sums_t = buf.packed1[R * (C + 1):].view(R, C + 1)
            g_wt = pool_backward_weights(ctx.xt_saved, a.wt_fmt, K, g_t, sums_t, float(a.grad_scale)) * gup
        return None, None, gxs, gxt, g_aug, g_wt


class CLRStep:
    """Stateful fused CLR step (EMA prototypes live here).

    Parameters mirror the reference trainers: ``decay`` = ``global_pro_weight`` (0.9), ``pro_weight`` (0.1),
    ``src_reg_weight`` / ``aug_weight`` (Trainer_prototype_mt ctor, bytecode L34), margin 0.01 (L455).
    ``backprop_aug``: the bytecode computes and logs ``loss_aug`` but never back-propagates it; set True to
    add it to ``total`` (and obtain ``d total / d oT_aug``).
    """

    def __init__(self, K: int = 2, decay: float = 0.9, pro_weight: float = 0.1, inter_weight: float = 0.0,
                 src_reg_weight: float = 1.0, aug_weight: float = 1.0, margin: float = 0.01,
                 retrify: bool = True, use_disc: bool = True, use_cons: bool = True, backprop_aug: bool = False,
                 global_batch: Optional[int] = None):
        self.K, self.decay = K, float(decay)
        self.pro_weight, self.inter_weight = float(pro_weight), float(inter_weight)
        self.src_reg_weight, self.aug_weight, self.margin = float(src_reg_weight), float(aug_weight), float(margin)
        self.retrify, self.use_disc, self.use_cons, self.backprop_aug = retrify, use_disc, use_cons, backprop_aug
        self.global_batch = global_batch
        self.stored_s: Optional[torch.Tensor] = None
        self.stored_t: Optional[torch.Tensor] = None
        self.first_s = True
        self.first_t = True

    # -- state -------------------------------------------------------------------------------------
    def state_dict(self):
        """Snapshot (clones: later steps update the live EMA tensors in place)."""
        return dict(stored_s=None if self.stored_s is None else self.stored_s.clone(),
                    stored_t=None if self.stored_t is None else self.stored_t.clone(),
                    first_s=self.first_s, first_t=self.first_t, K=self.K)

    def load_state_dict(self, sd):
        ss, st_ = sd["stored_s"], sd["stored_t"]
        if (ss is None) != (st_ is None):
            raise ValueError("stored_s / stored_t must both be present or both be None")
        if ss is not None:
            for name, t in (("stored_s", ss), ("stored_t", st_)):
                if not (isinstance(t, torch.Tensor) and t.dtype == torch.float32 and t.dim() == 2 and t.shape[0] == 2 * self.K):
                    raise ValueError("%s must be a float32 [2K=%d, C] tensor" % (name, 2 * self.K))
            if ss.shape != st_.shape or ss.device != st_.device:
                raise ValueError("stored_s / stored_t disagree in shape or device")
            ss, st_ = ss.detach().clone().contiguous(), st_.detach().clone().contiguous()
        self.stored_s, self.stored_t = ss, st_
        self.first_s, self.first_t = bool(sd["first_s"]), bool(sd["first_t"])

    # -- the step ------------------------------------------------------------------------------------
    def _prepare(self, xs_feature, pred_oS, xt_feature, oT_before=None, wt=None, preds=None, T: int = 8,
                 oT=None, oT_aug=None, masks=None, epoch: float = 0.0):
        """Validate inputs, allocate the step's buffers and fill the C argument block."""
        lib = _lib.load()
        xs = _require_cuda_f32(xs_feature, "xs_feature")
        ys = _require_cuda_f32(pred_oS.detach(), "pred_oS")
        xt = _require_cuda_f32(xt_feature, "xt_feature")
        dev = xs.device
        B_s, C, H, W = xs.shape
        # the hard source labels may come at IMAGE resolution (the trainer's target_map): the step then does the
        # trainer's F.interpolate(target_map, size=..., mode='nearest') (Trainer_prototype_full.py:329-330) itself, as one
        # small launch in front of every run (clr_label_downsample)
        ys_full = None
        if ys.shape[:2] == (B_s, self.K) and tuple(ys.shape[2:]) != (H, W):
            ys_full = ys
            ys = torch.empty(B_s, self.K, H, W, dtype=torch.float32, device=dev)
        B_t = xt.shape[0]
        K = self.K
        if ys.shape != (B_s, K, H, W) or xt.shape[1:] != (C, H, W):
            raise ValueError("inconsistent shapes: xs %s ys %s xt %s" % (tuple(xs.shape), tuple(ys.shape), tuple(xt.shape)))
        Hi = Wi = 0
        use_cons = self.use_cons and oT_aug is not None
        if self.retrify:
            if oT_before is None or preds is None:
                raise ValueError("retrify=True needs oT_before and preds")
            oTb = _require_cuda_f32(oT_before.detach(), "oT_before")
            pr = _require_cuda_f32(preds.detach(), "preds")
            if pr.shape[0] != T * B_t or pr.shape[1] != K:
                raise ValueError("preds must be [T*B_t, K, Hi, Wi]")
            Hi, Wi = pr.shape[2:]
            wt_t, wt_fmt = None, CLR_W_EXPLICIT
        else:
            if wt is None:
                if oT_before is None:
                    raise ValueError("retrify=False needs wt (target weights) or oT_before")
                wt = torch.sigmoid(oT_before)     # graph kept: the step returns dL/dwt, autograd chains it into oT_before
            wt_t = _require_cuda_f32(wt.detach(), "wt")
            wt_fmt = CLR_W_COMPLEMENT if wt_t.shape[1] == K else CLR_W_EXPLICIT
            oTb = pr = None
        if use_cons:
            oT_d = _require_cuda_f32(oT.detach(), "oT")
            oTa = _require_cuda_f32(oT_aug, "oT_aug")
            Hi, Wi = oT_d.shape[2:]
            if not self.retrify:
                if masks is None:
                    raise ValueError("consistency without retrify needs explicit masks [B,K,H,W]")
                masks_t = _require_cuda_f32(masks.detach(), "masks")
        else:
            oT_d = oTa = None

        buf = _Buffers(self, dev, B_s, B_t, C, H, W, K, Hi, Wi)
        if self.stored_s is None:
            self.stored_s = torch.zeros(2 * K, C, dtype=torch.float32, device=dev)
            self.stored_t = torch.zeros(2 * K, C, dtype=torch.float32, device=dev)
        elif self.stored_s.shape != (2 * K, C) or self.stored_s.device != dev or not self.stored_s.is_cuda:
            raise ValueError("EMA state is %s on %s, the step needs [%d, %d] on %s (load_state_dict / a different C?)"
                             % (tuple(self.stored_s.shape), self.stored_s.device, 2 * K, C, dev))

        a = StepArgs()
        a.B_s, a.B_t, a.C, a.H, a.W, a.K = B_s, B_t, C, H, W, K
        a.Hi, a.Wi, a.T = Hi, Wi, int(T)
        a.use_retrify, a.use_disc, a.use_cons = int(self.retrify), int(self.use_disc), int(use_cons)
        a.wt_fmt, a.first_s, a.first_t = wt_fmt, int(self.first_s), int(self.first_t)
        world = _dist.world_size()
        gb = self.global_batch if self.global_batch is not None else world * B_s
        a.decay, a.npx_global = self.decay, float(gb * H * W)
        a.w_intra, a.w_inter, a.w_disc = self.pro_weight, self.inter_weight, self.src_reg_weight if self.use_disc else 0.0
        a.w_aug = 1.0 if (use_cons and self.backprop_aug) else 0.0
        a.margin, a.aug_weight = self.margin, self.aug_weight
        a.cons_threshold = consistency_threshold(epoch)
        a.pseudo_thr, a.std_thr = PSEUDO_THRESHOLD, STD_THRESHOLD
        a.grad_scale = _dist.grad_scale() if _dist.enabled() else 1.0
        a.xs, a.ys, a.xt, a.wt = ptr(xs), ptr(ys), ptr(xt), ptr(wt_t)
        a.oT_before, a.preds, a.oT, a.oT_aug = ptr(oTb), ptr(pr), ptr(oT_d), ptr(oTa)
        a.gup = None
        a.stored_s, a.stored_t = ptr(self.stored_s), ptr(self.stored_t)
        for name in ("packed1", "packed2", "P_s", "P_t", "g_s", "g_t", "losses"):
            setattr(a, name, ptr(getattr(buf, name)))
        if self.use_disc:
            for name in ("disc_vec", "disc_beta", "xtab", "disc_coef"):
                setattr(a, name, ptr(getattr(buf, name)))
        if self.retrify:
            for name in ("std_map", "pred_mean", "wt_retrify", "masks"):
                setattr(a, name, ptr(getattr(buf, name)))
        elif use_cons:
            a.masks = ptr(masks_t)
        peer = None
        if _dist.peer_enabled():
            peer = _dist.peer_buffers(K, C, dev)
            a.world, a.rank, a.seq = peer["world"], peer["rank"], 1
            for q, pq in enumerate(peer["ptrs"]):
                a.peer_rx[q] = pq
        a.ws = None
        ws_bytes = lib.clr_step_ws_bytes(ctypes.byref(a))
        # one workspace per call / plan (it holds the per-CTA partials and the completion / gate counters of ONE step in
        # flight): two plans of the same CLRStep never share it.  The EMA state IS shared -- drive one CLRStep from one stream.
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        a.ws, a.ws_bytes = ptr(ws), ws_bytes

        wt_soft = wt if (not self.retrify and isinstance(wt, torch.Tensor) and wt.requires_grad) else None
        holder = dict(args=a, buf=buf, xs=xs, xt=xt, oT_aug=oTa, dims=(B_s, B_t, C, H, W, K, Hi, Wi), peer=peer, wt_soft=wt_soft,
                      ys=ys, ys_full=ys_full,
                      keep=(xs, ys, xt, wt_t, oTb, pr, oT_d, oTa, self.stored_s, self.stored_t, ws,
                            masks_t if (use_cons and not self.retrify) else None))
        return holder

    def _outputs(self, holder, total) -> CLRStepOutput:
        buf = holder["buf"]
        B_s, B_t, C, H, W, K, Hi, Wi = holder["dims"]
        L = buf.losses
        R = 2 * K
        Ps = [buf.P_s[r * C:(r + 1) * C].view(1, C, 1, 1) for r in range(R)]
        Pt = [buf.P_t[r * C:(r + 1) * C].view(1, C, 1, 1) for r in range(R)]
        std_map = mask_list = None
        if self.retrify:
            std_map = buf.std_map.view(B_t, K, Hi, Wi)
            m = buf.masks.view(B_t, K, H, W)
            mask_list = [m[:, k:k + 1] for k in range(K)]
        return CLRStepOutput(total=total, intra=L[0], inter=L[1], disc=L[2], aug=L[3], source_prototypes=Ps,
                             target_prototypes=Pt, std_map=std_map, masks=mask_list, error=L[7])

    def __call__(self, xs_feature, pred_oS, xt_feature, oT_before=None, wt=None, preds=None, T: int = 8,
                 oT=None, oT_aug=None, masks=None, epoch: float = 0.0) -> CLRStepOutput:
        """One differentiable CLR step; ``out.total.backward()`` (or adding it to the trainer's ``loss_all``)
        writes the gradients of ``xs_feature``, ``xt_feature`` (and ``oT_aug`` with ``backprop_aug``).  Without retrify,
        soft target weights that require grad (``wt``, or ``sigmoid(oT_before)`` built here) receive theirs too, as
        in ``gen_prototype(torch.sigmoid(oT_before), xt_feature)`` (Trainer_prototype_full.py:375-377)."""
        holder = self._prepare(xs_feature, pred_oS, xt_feature, oT_before, wt, preds, T, oT, oT_aug, masks, epoch)
        total = _ClrStepFn.apply(self, holder, holder["xs"], holder["xt"], holder["oT_aug"], holder["wt_soft"])
        self.first_s = self.first_t = False
        return self._outputs(holder, total)

    def plan(self, xs_feature, pred_oS, xt_feature, oT_before=None, wt=None, preds=None, T: int = 8,
             oT=None, oT_aug=None, masks=None, epoch: float = 0.0) -> "CLRPlan":
        """Bind the step to fixed input buffers (training loops that reuse their activation buffers, CUDA-graph
        capture, the benchmark): afterwards :meth:`CLRPlan.run` is two C calls -- forward and backward -- with no
        per-step Python work and no autograd tape."""
        holder = self._prepare(xs_feature, pred_oS, xt_feature, oT_before, wt, preds, T, oT, oT_aug, masks, epoch)
        return CLRPlan(self, holder)


class CLRPlan:
    """A prebound fused step: ``run()`` enqueues forward + backward on the current stream.

    Gradients land in ``gxs`` / ``gxt`` (/ ``g_oT_aug``), losses in ``losses`` (device, ``[intra, inter, disc, aug,
    total, ...]``); nothing is synchronised.  The EMA state advances in the owning :class:`CLRStep`.
    """

    def __init__(self, step: CLRStep, holder):
        self.step, self.holder = step, holder
        a: StepArgs = holder["args"]
        buf = holder["buf"]
        dev = buf.flat.device
        self.gxs = torch.empty_like(holder["xs"])
        self.gxt = torch.empty_like(holder["xt"])
        want_aug = holder["oT_aug"] is not None and a.use_cons and a.w_aug != 0.0
        self.g_oT_aug = torch.empty_like(holder["oT_aug"]) if want_aug else None
        a.gxs, a.gxt, a.g_oT_aug, a.gup = ptr(self.gxs), ptr(self.gxt), ptr(self.g_oT_aug), None
        self.losses = buf.losses
        self.device = dev
        self._lib = _lib.load()
        self._ref = ctypes.byref(a)
        # soft target weights (no retrify) that carry a graph: run() also leaves d total / d wt in ``g_wt`` ([B,K,H,W] for
        # complement weights), the quantity autograd chains into oT_before in the __call__ path
        self.g_wt = None
        if holder["wt_soft"] is not None:
            Q = a.K if a.wt_fmt == CLR_W_COMPLEMENT else 2 * a.K
            self.g_wt = torch.empty(a.B_t, Q, a.H, a.W, dtype=torch.float32, device=dev)
            self._gw_ws_bytes = self._lib.clr_pool_bwd_w_ws_bytes(a.C, a.K, a.wt_fmt)
            self._gw_ws = torch.empty(self._gw_ws_bytes, dtype=torch.uint8, device=dev)

    def set_events(self, pool_begin=None, pool_end=None, bwd_begin=None, bwd_end=None) -> None:
        """Have the library record these :class:`uda_clr_b200._lib.Event` objects around the pooling / backward
        launches of the next :meth:`run` (``None`` = do not record)."""
        a: StepArgs = self.holder["args"]
        a.ev_pool_begin = None if pool_begin is None else pool_begin.handle
        a.ev_pool_end = None if pool_end is None else pool_end.handle
        a.ev_bwd_begin = None if bwd_begin is None else bwd_begin.handle
        a.ev_bwd_end = None if bwd_end is None else bwd_end.handle

    @property
    def error(self) -> torch.Tensor:
        """Device scalar (``losses[7]``): 0 = ok, 1 = a device-side wait timed out (see :class:`CLRStepOutput`)."""
        return self.losses[7]

    def check(self) -> None:
        """Synchronise and raise :class:`CLRStepError` if any step run so far on this plan timed out on the device."""
        if float(self.losses[7]) != 0.0:
            raise CLRStepError("fused CLR step: a device-side wait timed out (in-kernel exchange peer missing or gate missed); "
                               "losses / gradients of that step are NaN, the EMA state was not advanced")

    def run(self) -> None:
        with torch.cuda.device(self.device):
            self._run()

    def _run(self) -> None:
        a: StepArgs = self.holder["args"]
        st = _stream()
        a.first_s, a.first_t = int(self.step.first_s), int(self.step.first_t)
        _downsample_labels(self._lib, self.holder, st)
        if self.holder["peer"] is not None:
            a.seq = _dist.next_seq(self.holder["peer"])
            check(self._lib.clr_step_run(self._ref, st), "clr_step_run (in-kernel exchange)")
        elif _dist.enabled():
            check(self._lib.clr_step_fwd_a(self._ref, st), "clr_step_fwd_a")
            _dist.all_reduce_sums(self.holder["buf"].packed1)
            check(self._lib.clr_step_fwd_b(self._ref, st), "clr_step_fwd_b")
            _dist.all_reduce_sums(self.holder["buf"].packed2)
            check(self._lib.clr_step_fwd_c(self._ref, st), "clr_step_fwd_c")
            check(self._lib.clr_step_bwd(self._ref, st), "clr_step_bwd")
        else:
            check(self._lib.clr_step_run(self._ref, st), "clr_step_run")
        if self.g_wt is not None:
            buf = self.holder["buf"]
            R, C = 2 * a.K, a.C
            check(self._lib.clr_pool_bwd_w(ptr(self.holder["xt"]), a.wt_fmt, a.B_t, C, a.H * a.W, a.K, ptr(buf.g_t),
                                           ptr(buf.packed1[R * (C + 1):]), float(a.grad_scale), ptr(self._gw_ws),
                                           self._gw_ws_bytes, ptr(self.g_wt), st), "clr_pool_bwd_w")
        self.step.first_s = self.step.first_t = False

    def outputs(self) -> CLRStepOutput:
        return self.step._outputs(self.holder, self.losses[4])
