"""Host-side mirror of the reference's CLR functions (``utils/Utils.py:86-311``) over the C ABI.

Same names, positional signatures, tuple arity/order, shapes (``[1,C,1,1]``), dtype and device as the
reference; each is a ``torch.autograd.Function`` whose forward/backward enqueue the hand-written
sm_100a kernels of ``libclr_b200.so`` on the current CUDA stream.  PyTorch is plumbing here: it owns
the device memory, the stream and (optionally) the process group.  There is no CPU path -- CPU
tensors, non-fp32 dtypes or a missing library raise.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import CLR_W_COMPLEMENT, CLR_W_EXPLICIT, check, ptr
from . import dist as _dist


# ----------------------------------------------------------------------------------------------- helpers
_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """``cudaStream_t`` of the current stream of the current device, as an integer.  ``torch.cuda.current_stream()`` builds a
    Python ``Stream`` object per call (17 us, seven calls per training step of the drop-in ops -- measured with
    tools/dropin_profile.py); the raw getter behind it costs ~1 us."""
    if _RAW_STREAM is not None:
        return _RAW_STREAM(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _on(device):
    """Device guard for a launch: ``torch.cuda.device(device)`` only when ``device`` is not already current (the context
    manager costs ~10 us of host time per op, and the zero-line drop-in is host-bound)."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def _require_cuda_f32(t: torch.Tensor, name: str, ndim: int = 4) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the CLR ops have no CPU fallback" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32 (got %s)" % (name, t.dtype))
    if t.dim() != ndim:
        raise ValueError("%s must be %d-D (got shape %s)" % (name, ndim, tuple(t.shape)))
    return t.contiguous()


def _check_k(K: int) -> None:
    if not 1 <= K <= _lib.CLR_MAX_K:
        raise ValueError("number of classes K=%d outside [1, %d]" % (K, _lib.CLR_MAX_K))


# Scratch of the drop-in ops (per-CTA partials, consumed inside the same call): cached per (device, stream, size class)
# instead of a torch.empty per call -- the zero-line drop-in is host-bound, every allocator round trip counts.  Keyed by
# stream because two streams may run the same op concurrently; outputs are always fresh tensors (autograd owns them).
_WS_CACHE = {}


def _workspace(nbytes: int, device, tag: str) -> torch.Tensor:
    key = (device.index, _RAW_STREAM(device.index) if (_RAW_STREAM is not None and device.index is not None)
           else torch.cuda.current_stream(device).cuda_stream, tag)
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _WS_CACHE[key] = ws
    return ws


def pool_sums(feat: torch.Tensor, w: torch.Tensor, fmt: int, K: int, with_mu: bool = False):
    """Packed class-wise sums ``[2K, C+1]`` (column C = weight sums) of one domain.  Local to this rank.
    ``with_mu=True`` also returns the prototypes ``[2K, C]`` from the same two launches (``clr_pool_fwd_mu``)."""
    lib = _lib.load()
    B, C, H, W = feat.shape
    HW = H * W
    with _on(feat.device):
        ws_bytes = lib.clr_pool_ws_bytes(B, C, HW, K)
        ws = _workspace(ws_bytes, feat.device, "pool")
        sums = torch.empty(2 * K, C + 1, dtype=torch.float32, device=feat.device)
        if with_mu:
            mu = torch.empty(2 * K, C, dtype=torch.float32, device=feat.device)
            check(lib.clr_pool_fwd_mu(ptr(feat), ptr(w), fmt, B, C, HW, K, ptr(ws), ws_bytes, ptr(sums), ptr(mu), _stream()),
                  "clr_pool_fwd_mu")
            return sums, mu
        check(lib.clr_pool_fwd(ptr(feat), ptr(w), fmt, B, C, HW, K, ptr(ws), ws_bytes, ptr(sums), _stream()),
              "clr_pool_fwd")
    return sums


def protos_from_sums(sums: torch.Tensor) -> torch.Tensor:
    """``mu[r][c] = S_r[c] / N_r`` -> ``[R, C]`` (0/0 -> NaN, like utils/Utils.py:127-130)."""
    lib = _lib.load()
    R, C1 = sums.shape
    mu = torch.empty(R, C1 - 1, dtype=torch.float32, device=sums.device)
    with _on(sums.device):
        check(lib.clr_proto_finalize(ptr(sums), R, C1 - 1, ptr(mu), _stream()), "clr_proto_finalize")
    return mu


def pool_backward_feat(w: torch.Tensor, fmt: int, K: int, feat_shape, g: torch.Tensor, sums: torch.Tensor,
                       scale: float = 1.0, xcoef: Optional[torch.Tensor] = None,
                       xtab: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    B, C, H, W = feat_shape
    grad = torch.empty(feat_shape, dtype=torch.float32, device=w.device)
    Kx = 0 if xcoef is None else xcoef.shape[1]
    with _on(w.device):
        check(lib.clr_pool_bwd(ptr(w), fmt, B, C, H * W, K, ptr(g), ptr(sums), float(scale),
                               ptr(xcoef), ptr(xtab), Kx, ptr(grad), _stream()), "clr_pool_bwd")
    return grad


def pool_backward_weights(feat: torch.Tensor, fmt: int, K: int, g: torch.Tensor, sums: torch.Tensor,
                          scale: float = 1.0) -> torch.Tensor:
    lib = _lib.load()
    B, C, H, W = feat.shape
    Q = K if fmt == CLR_W_COMPLEMENT else 2 * K
    ws_bytes = lib.clr_pool_bwd_w_ws_bytes(C, K, fmt)
    ws = _workspace(ws_bytes, feat.device, "bwd_w")
    out = torch.empty(B, Q, H, W, dtype=torch.float32, device=feat.device)
    with _on(feat.device):
        check(lib.clr_pool_bwd_w(ptr(feat), fmt, B, C, H * W, K, ptr(g), ptr(sums), float(scale),
                                 ptr(ws), ws_bytes, ptr(out), _stream()), "clr_pool_bwd_w")
    return out


def _stack_grads(grads: Sequence[Optional[torch.Tensor]], R: int, C: int, device) -> torch.Tensor:
    if all(gr is not None for gr in grads[:R]):      # the usual case (all 2K prototypes enter the loss): one launch, not 1 + R
        return torch.cat([gr.reshape(1, C) for gr in grads[:R]], 0).to(torch.float32).contiguous()
    g = torch.zeros(R, C, dtype=torch.float32, device=device)
    for r, gr in enumerate(grads[:R]):
        if gr is not None:
            g[r].copy_(gr.reshape(C))
    return g


def _split_protos(mu: torch.Tensor) -> Tuple[torch.Tensor, ...]:
    R, C = mu.shape
    return tuple(mu[r].view(1, C, 1, 1) for r in range(R))


# ----------------------------------------------------------------------------------------------- A1 / A3
class _WeightedPrototypes(torch.autograd.Function):
    """``mu_r = sum x w_r / sum w_r`` over one or two domains whose sums are added before the divide
    (utils/Utils.py:132-158, :227-311).

    inputs: (K, fmts, w_0, feat_0[, w_1, feat_1]); outputs: 2K tensors ``[1,C,1,1]``.
    """

    @staticmethod
    def forward(ctx, K: int, fmts, *tensors):
        n_dom = len(tensors) // 2
        ws, feats = tensors[0::2], tensors[1::2]
        scale = 1.0
        if n_dom == 1 and not _dist.enabled():
            sums, mu = pool_sums(feats[0], ws[0], fmts[0], K, with_mu=True)     # prototypes leave the reduce launch
        else:
            sums = None
            for w, f, fmt in zip(ws, feats, fmts):
                s = pool_sums(f, w, fmt, K)
                sums = s if sums is None else sums + s
            if _dist.enabled():
                _dist.all_reduce_sums(sums)
                scale = _dist.grad_scale()
            mu = protos_from_sums(sums)
        ctx.fmts, ctx.K, ctx.n_dom, ctx.scale = tuple(fmts), K, n_dom, scale
        ctx.shapes = [tuple(f.shape) for f in feats]
        need_w = [ctx.needs_input_grad[2 + 2 * i] for i in range(n_dom)]
        ctx.save_for_backward(sums, *ws, *[f if nw else None for f, nw in zip(feats, need_w)])
        return _split_protos(mu)

    @staticmethod
    def backward(ctx, *grads):
        saved = ctx.saved_tensors
        sums, ws, feats = saved[0], saved[1:1 + ctx.n_dom], saved[1 + ctx.n_dom:]
        K = ctx.K
        C = ctx.shapes[0][1]
        g = _stack_grads(grads, 2 * K, C, sums.device)
        out = [None, None]
        for i in range(ctx.n_dom):
            gw = gf = None
            if ctx.needs_input_grad[2 + 2 * i]:
                gw = pool_backward_weights(feats[i], ctx.fmts[i], K, g, sums, ctx.scale)
            if ctx.needs_input_grad[3 + 2 * i]:
                gf = pool_backward_feat(ws[i], ctx.fmts[i], K, ctx.shapes[i], g, sums, ctx.scale)
            out += [gw, gf]
        return tuple(out)


def gen_prototype(pred_oS: torch.Tensor, xs_feature: torch.Tensor):
    """Drop-in for ``utils.Utils.gen_prototype`` (utils/Utils.py:108-131).

    ``pred_oS [B,K,H,W]`` (hard {0,1} labels or soft sigmoid predictions), ``xs_feature [B,C,H,W]`` ->
    ``(c0_obj, c1_obj, c0_bck, c1_bck)`` for K = 2 (``obj_0..obj_{K-1}, bck_0..bck_{K-1}`` in general),
    each ``[1,C,1,1]``.  Gradients flow to ``xs_feature`` and, if it requires grad, to ``pred_oS``.
    """
    pred = _require_cuda_f32(pred_oS, "pred_oS")
    feat = _require_cuda_f32(xs_feature, "xs_feature")
    K = pred.shape[1]
    _check_k(K)
    if pred.shape[0] != feat.shape[0] or pred.shape[2:] != feat.shape[2:]:
        raise ValueError("pred_oS %s and xs_feature %s disagree" % (tuple(pred.shape), tuple(feat.shape)))
    return _WeightedPrototypes.apply(K, (CLR_W_COMPLEMENT,), pred, feat)


def gen_prototype_src_trg(pred_oS, xs_feature, pred_oT, xt_feature):
    """Drop-in for ``utils.Utils.gen_prototype_src_trg`` (utils/Utils.py:132-158): joint source+target
    prototypes ``(S_s+S_t)/(N_s+N_t)`` without materialising the concatenation."""
    ps, fs = _require_cuda_f32(pred_oS, "pred_oS"), _require_cuda_f32(xs_feature, "xs_feature")
    pt, ft = _require_cuda_f32(pred_oT, "pred_oT"), _require_cuda_f32(xt_feature, "xt_feature")
    K = ps.shape[1]
    _check_k(K)
    if pt.shape[1] != K or fs.shape[1] != ft.shape[1]:
        raise ValueError("source and target disagree on K or C")
    return _WeightedPrototypes.apply(K, (CLR_W_COMPLEMENT, CLR_W_COMPLEMENT), ps, fs, pt, ft)


def weighted_prototypes(weights: torch.Tensor, feat: torch.Tensor):
    """Prototypes from explicit weight planes ``[B,2K,H,W]`` (rows obj_0.., bck_0..)."""
    w = _require_cuda_f32(weights, "weights")
    f = _require_cuda_f32(feat, "feat")
    if w.shape[1] % 2:
        raise ValueError("explicit weights need 2K planes")
    K = w.shape[1] // 2
    _check_k(K)
    return _WeightedPrototypes.apply(K, (CLR_W_EXPLICIT,), w, f)


# ----------------------------------------------------------------------------------------------- A2
PSEUDO_THRESHOLD = 0.75   # utils/Utils.py:176
STD_THRESHOLD = 0.04      # utils/Utils.py:197


def mc_statistics(preds: torch.Tensor, T: int, stride: int):
    """``std_T(sigmoid(p/2))`` (unbiased) and ``mean_T(sigmoid(p))`` of ``preds [T*stride,K,Hi,Wi]``
    (utils/Utils.py:161-168) -> two ``[stride,K,Hi,Wi]`` maps.  One read of ``preds``."""
    lib = _lib.load()
    p = _require_cuda_f32(preds, "preds")
    if p.shape[0] != T * stride:
        raise ValueError("preds has %d maps, expected T*stride = %d" % (p.shape[0], T * stride))
    _, K, Hi, Wi = p.shape
    std_map = torch.empty(stride, K, Hi, Wi, dtype=torch.float32, device=p.device)
    pred_mean = torch.empty_like(std_map)
    with _on(p.device):
        check(lib.clr_mc_stats(ptr(p), T, stride, K, Hi, Wi, ptr(std_map), ptr(pred_mean), _stream()), "clr_mc_stats")
    return std_map, pred_mean


def retrify_weights(oT_before: torch.Tensor, pred_mean: torch.Tensor, std_map: torch.Tensor, H: int, W: int,
                    pseudo_thr: float = PSEUDO_THRESHOLD, std_thr: float = STD_THRESHOLD, debug: bool = False,
                    preds: Optional[torch.Tensor] = None, T: int = 0):
    """Explicit target weights ``[B,2K,H,W]`` and uncertainty masks ``[B,K,H,W]`` in {0,2}
    (utils/Utils.py:170-223).  ``debug=True`` also returns the pseudo-labels and the two down-sampled maps.
    ``preds`` / ``T`` (the MC logits ``mc_statistics`` was computed from): pixels whose std lies within 1e-5 of the
    threshold are re-evaluated from them in ATen's exact order, so the masks equal eager torch bit for bit."""
    lib = _lib.load()
    o = _require_cuda_f32(oT_before, "oT_before")
    B, K = o.shape[:2]
    Hi, Wi = pred_mean.shape[2:]
    weights = torch.empty(B, 2 * K, H, W, dtype=torch.float32, device=o.device)
    masks = torch.empty(B, K, H, W, dtype=torch.float32, device=o.device)
    pseudo = torch.empty(B, K, H, W, dtype=torch.float32, device=o.device) if debug else None
    small = torch.empty(2, B, K, H, W, dtype=torch.float32, device=o.device) if debug else None
    with _on(o.device):
        pr = None if preds is None else _require_cuda_f32(preds, "preds")
        check(lib.clr_retrify_weights(ptr(o), ptr(pred_mean), ptr(std_map), ptr(pr), int(T), B, K, H, W, Hi, Wi,
                                      float(pseudo_thr), float(std_thr), ptr(weights), ptr(masks),
                                      ptr(pseudo), ptr(small), _stream()), "clr_retrify_weights")
    if debug:
        return weights, masks, pseudo, small
    return weights, masks


def mc_retrify(oT_before: torch.Tensor, preds: torch.Tensor, T: int, stride: int, H: int, W: int,
               pseudo_thr: float = PSEUDO_THRESHOLD, std_thr: float = STD_THRESHOLD):
    """``mc_statistics`` + ``retrify_weights`` in one pass over ``preds`` (``clr_mc_retrify``): returns
    ``(std_map [stride,K,Hi,Wi], weights [B,2K,H,W], masks [B,K,H,W])``.  Falls back to the two separate kernels when
    the geometry does not allow the fusion (up-sampling factor below 2, ragged or misaligned maps)."""
    lib = _lib.load()
    o = _require_cuda_f32(oT_before, "oT_before")
    p = _require_cuda_f32(preds, "preds")
    if p.shape[0] != T * stride:
        raise ValueError("preds has %d maps, expected T*stride = %d" % (p.shape[0], T * stride))
    B, K = o.shape[:2]
    Hi, Wi = p.shape[2:]
    std_map = torch.empty(stride, K, Hi, Wi, dtype=torch.float32, device=p.device)
    weights = torch.empty(B, 2 * K, H, W, dtype=torch.float32, device=o.device)
    masks = torch.empty(B, K, H, W, dtype=torch.float32, device=o.device)
    with _on(p.device):
        rc = lib.clr_mc_retrify(ptr(p), ptr(o), int(T), int(stride), K, H, W, Hi, Wi, float(pseudo_thr), float(std_thr),
                                ptr(std_map), None, ptr(weights), ptr(masks), _stream())
    if rc == _lib.CLR_ERR_UNSUPPORTED:
        std_map, pred_mean = mc_statistics(p, T, stride)
        weights, masks = retrify_weights(o, pred_mean, std_map, H, W, pseudo_thr, std_thr, preds=p, T=T)
        return std_map, weights, masks
    check(rc, "clr_mc_retrify")
    return std_map, weights, masks


class _RetrifyPrototypes(torch.autograd.Function):
    """A2 end to end; optional joint source domain (A3 retrify variant, utils/Utils.py:227-311).

    inputs: (T, stride, oT_before, xt_feature, preds[, pred_oS, xs_feature])
    outputs: 2K prototypes, std_map, K masks.
    Gradients: to xt_feature (and xs_feature / pred_oS in the joint form); oT_before receives exact zeros,
    as in the reference where the clone + masked fills cut the graph (utils/Utils.py:175-177).
    """

    @staticmethod
    def forward(ctx, T, stride, oT_before, xt_feature, preds, pred_oS=None, xs_feature=None):
        K = oT_before.shape[1]
        H, W = xt_feature.shape[2:]
        std_map, weights, masks = mc_retrify(oT_before, preds, T, stride, H, W)
        joint = xs_feature is not None
        scale = 1.0
        if not joint and not _dist.enabled():
            sums, mu = pool_sums(xt_feature, weights, CLR_W_EXPLICIT, K, with_mu=True)
        else:
            sums = pool_sums(xt_feature, weights, CLR_W_EXPLICIT, K)
            if joint:
                sums = sums + pool_sums(xs_feature, pred_oS, CLR_W_COMPLEMENT, K)
            if _dist.enabled():
                _dist.all_reduce_sums(sums)
                scale = _dist.grad_scale()
            mu = protos_from_sums(sums)
        ctx.K, ctx.scale, ctx.joint = K, scale, joint
        ctx.n_in = 7 if joint else 5
        ctx.t_shape = tuple(xt_feature.shape)
        ctx.s_shape = tuple(xs_feature.shape) if joint else None
        ctx.o_shape = tuple(oT_before.shape)
        need_ps = joint and ctx.needs_input_grad[5]
        ctx.save_for_backward(sums, weights, pred_oS if joint else None, xs_feature if need_ps else None)
        mask_list = tuple(masks[:, k:k + 1] for k in range(K))
        ctx.mark_non_differentiable(std_map, *mask_list)
        return _split_protos(mu) + (std_map,) + mask_list

    @staticmethod
    def backward(ctx, *grads):
        sums, weights, pred_oS, xs_feature = ctx.saved_tensors
        K = ctx.K
        C = ctx.t_shape[1]
        g = _stack_grads(grads, 2 * K, C, sums.device)
        g_oT = torch.zeros(ctx.o_shape, dtype=torch.float32, device=sums.device) if ctx.needs_input_grad[2] else None
        g_xt = None
        if ctx.needs_input_grad[3]:
            g_xt = pool_backward_feat(weights, CLR_W_EXPLICIT, K, ctx.t_shape, g, sums, ctx.scale)
        g_ps = g_xs = None
        if ctx.joint:
            if ctx.needs_input_grad[5]:
                g_ps = pool_backward_weights(xs_feature, CLR_W_COMPLEMENT, K, g, sums, ctx.scale)
            if ctx.needs_input_grad[6]:
                g_xs = pool_backward_feat(pred_oS, CLR_W_COMPLEMENT, K, ctx.s_shape, g, sums, ctx.scale)
        return (None, None, g_oT, g_xt, None, g_ps, g_xs)[:ctx.n_in]


def _retrify_inputs(oT_before, xt_feature, preds, T, stride):
    o = _require_cuda_f32(oT_before, "oT_before")
    x = _require_cuda_f32(xt_feature, "xt_feature")
    p = _require_cuda_f32(preds, "preds")
    K = o.shape[1]
    _check_k(K)
    if o.shape[0] != x.shape[0] or o.shape[2:] != x.shape[2:] or p.shape[1] != K or stride != x.shape[0]:
        raise ValueError("gen_prototype_retrify: inconsistent shapes oT_before %s, xt_feature %s, preds %s, stride %d"
                         % (tuple(o.shape), tuple(x.shape), tuple(p.shape), stride))
    return o, x, p


def gen_prototype_retrify(oT_before, xt_feature, preds, features, T, stride):
    """Drop-in for ``utils.Utils.gen_prototype_retrify`` (utils/Utils.py:159-225).

    Returns ``(c0_obj, c1_obj, c0_bck, c1_bck, std_map, mask_0, mask_1)`` for K = 2.  ``features`` is
    accepted and ignored: the reference averages it (:169) but only reads the result's spatial size,
    which equals ``xt_feature``'s; the 128x128 / 305-channel hard-codes (:162) are lifted.
    """
    del features
    o, x, p = _retrify_inputs(oT_before, xt_feature, preds, T, stride)
    return _RetrifyPrototypes.apply(int(T), int(stride), o, x, p)


def gen_prototype_src_trg_retrify(pred_oS, xs_feature, oT_before, xt_feature, preds, features, T, stride):
    """Drop-in for ``utils.Utils.gen_prototype_src_trg_retrify`` (utils/Utils.py:227-311): joint prototypes
    ``(S_s + S_t)/(N_s + N_t)`` with retrify weights on the target side.  Returns the 2K prototypes only."""
    del features
    o, x, p = _retrify_inputs(oT_before, xt_feature, preds, T, stride)
    ps, fs = _require_cuda_f32(pred_oS, "pred_oS"), _require_cuda_f32(xs_feature, "xs_feature")
    K = o.shape[1]
    return _RetrifyPrototypes.apply(int(T), int(stride), o, x, p, ps, fs)[:2 * K]


# ----------------------------------------------------------------------------------------------- A6
class _BmmPrototypes(torch.autograd.Function):
    """``proto[r] = mean_b( sum_p m[b,r,p] x[b,:,p] / (sum_p m[b,r,p] + n_add) )`` -- the bmm-style per-sample
    pooling of Trainer_prototype.py:364-383 / cal_prototype.py:156-175.  One read of ``feat`` forward, one write
    backward (the adjoint never re-reads the features)."""

    @staticmethod
    def forward(ctx, masks, feat, n_add):
        lib = _lib.load()
        B, C, H, W = feat.shape
        R, HW = masks.shape[1], H * W
        ws_bytes = 4 * R * (C + 1) + lib.clr_pool_rows_ws_bytes(B, C, HW, R)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=feat.device)
        sums_b = torch.empty(B, R, C + 1, dtype=torch.float32, device=feat.device)
        out = torch.empty(R, C, dtype=torch.float32, device=feat.device)
        with _on(feat.device):
            check(lib.clr_pool_rows_fwd_ps(ptr(feat), ptr(masks), B, C, HW, R, ptr(ws), ws_bytes, ptr(sums_b), _stream()),
                  "clr_pool_rows_fwd_ps")
            check(lib.clr_bmm_finalize(ptr(sums_b), B, R, C, float(n_add), ptr(out), _stream()), "clr_bmm_finalize")
        ctx.shape, ctx.n_add = (B, C, H, W), float(n_add)
        ctx.save_for_backward(masks, sums_b)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        masks, sums_b = ctx.saved_tensors
        B, C, H, W = ctx.shape
        R = masks.shape[1]
        gf = None
        if ctx.needs_input_grad[1]:
            g = g.contiguous()
            gf = torch.empty(ctx.shape, dtype=torch.float32, device=masks.device)
            with _on(masks.device):
                check(lib.clr_pool_bwd_ps(ptr(masks), B, C, H * W, R, ptr(g), ptr(sums_b), ctx.n_add, 1.0 / B,
                                          ptr(gf), _stream()), "clr_pool_bwd_ps")
        return None, gf, None


def bmm_prototypes(masks: torch.Tensor, feat: torch.Tensor, n_add: float = 1.0) -> torch.Tensor:
    """Per-sample-normalised (bmm-style) prototypes of ``Trainer_prototype.py:364-383, 404-451`` and
    ``cal_prototype.py:156-175``: ``masks [B,R,H,W]`` (R mask planes pooled in ONE read of ``feat``: e.g. cup and
    disc together, where the reference runs one cuBLAS bmm per mask) -> ``[R, C]``; row r equals the reference's
    ``torch.mean(bmm(m_r, X) / (sum m_r + 1), dim=0)``.  Gradient flows to ``feat`` (masks are labels or
    thresholded predictions in every reference call site)."""
    m = _require_cuda_f32(masks, "masks")
    f = _require_cuda_f32(feat, "feat")
    if m.shape[0] != f.shape[0] or m.shape[2:] != f.shape[2:]:
        raise ValueError("masks %s and feat %s disagree" % (tuple(m.shape), tuple(f.shape)))
    if not 1 <= m.shape[1] <= 2 * _lib.CLR_MAX_K:
        raise ValueError("number of mask planes R=%d outside [1, %d]" % (m.shape[1], 2 * _lib.CLR_MAX_K))
    return _BmmPrototypes.apply(m, f, float(n_add))


def update_objective_single_vector(obj: torch.Tensor, vector: torch.Tensor, rate: float = 0.001) -> torch.Tensor:
    """``Trainer.update_objective_SingleVector`` (Trainer_prototype.py:117-123): ``obj = obj * (1-rate) + rate * v``
    unless ``v`` sums to zero -- decided on the device inside the kernel (``clr_ema_rows``), without the reference's
    ``.item()`` host sync.  ``obj`` may be one vector ``[C]`` or a stack ``[R, C]`` (e.g. bu / cup / disc with equal C):
    every row gets its own zero test.  Returns a new tensor, like the reference's dict assignment."""
    lib = _lib.load()
    o = obj.detach()
    if not o.is_cuda or o.dtype != torch.float32:
        raise RuntimeError("update_objective_single_vector needs CUDA float32 tensors (no CPU fallback)")
    o = o.contiguous()
    C = o.shape[-1]
    R = o.numel() // C
    v = vector.detach().to(torch.float32).reshape(R, C).contiguous()
    out = torch.empty_like(o)
    with _on(o.device):
        check(lib.clr_ema_rows(ptr(v), ptr(o), R, C, float(rate), ptr(out), _stream()), "clr_ema_rows")
    return out


def nearest_labels(target_map: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """``F.interpolate(target_map, size=(H, W), mode='nearest')`` of the hard source labels
    (Trainer_prototype_full.py:329-330) -> ``[B,K,H,W]``; ATen's nearest source index, bit for bit."""
    lib = _lib.load()
    t = _require_cuda_f32(target_map.detach(), "target_map")
    B, K, Hi, Wi = t.shape
    out = torch.empty(B, K, H, W, dtype=torch.float32, device=t.device)
    with _on(t.device):
        check(lib.clr_label_downsample(ptr(t), B * K, Hi, Wi, int(H), int(W), ptr(out), _stream()), "clr_label_downsample")
    return out


class MCAccumulator:
    """MC-dropout statistics without the trainer's staging buffers (Trainer_prototype_full.py:359-368 fills a
    ``[T*B,2,512,512]`` ``preds_trg`` and a dead 1.28 GB ``features_trg`` every step):

        acc = MCAccumulator()
        for i in range(T // 2):
            with torch.no_grad():
                logits = model(volume_batch_r)[0]            # [2*stride, K, Hi, Wi]: two MC passes per forward
            acc.add(logits, passes=2)
        std_map, pred_mean = acc.finalize()                  # what mc_statistics(preds_trg, T, stride) returns

    Four running maps per position (pivot, shifted sum, shifted sum of squares, sum of sigmoid(p)); more bytes than one
    read of staged logits, no staging memory.  The raw logits are gone afterwards, so ``retrify_weights`` runs without the
    knife-edge guard of the uncertainty mask on this path (``preds=None``)."""

    def __init__(self):
        self.state: Optional[torch.Tensor] = None
        self.shape = None
        self.T = 0

    def reset(self) -> None:
        self.T = 0

    def add(self, logits: torch.Tensor, passes: int = 1) -> None:
        lib = _lib.load()
        x = _require_cuda_f32(logits.detach(), "logits")
        if x.shape[0] % passes:
            raise ValueError("logits has %d maps, not a multiple of passes = %d" % (x.shape[0], passes))
        B, K, Hi, Wi = x.shape[0] // passes, x.shape[1], x.shape[2], x.shape[3]
        if self.T == 0:
            n = lib.clr_mc_state_floats(B, K, Hi, Wi)
            if self.state is None or self.state.numel() != n or self.state.device != x.device:
                self.state = torch.empty(n, dtype=torch.float32, device=x.device)
            self.shape = (B, K, Hi, Wi)
        elif self.shape != (B, K, Hi, Wi):
            raise ValueError("MC passes of one step must agree in shape: %s vs %s" % (self.shape, (B, K, Hi, Wi)))
        with _on(x.device):
            check(lib.clr_mc_accumulate(ptr(x), int(passes), B, K, Hi, Wi, int(self.T == 0), ptr(self.state), _stream()),
                  "clr_mc_accumulate")
        self.T += passes

    def finalize(self):
        if self.T == 0:
            raise RuntimeError("MCAccumulator.finalize() before any add()")
        lib = _lib.load()
        B, K, Hi, Wi = self.shape
        std_map = torch.empty(B, K, Hi, Wi, dtype=torch.float32, device=self.state.device)
        pred_mean = torch.empty_like(std_map)
        with _on(self.state.device):
            check(lib.clr_mc_finalize(ptr(self.state), self.T, B, K, Hi, Wi, ptr(std_map), ptr(pred_mean), _stream()),
                  "clr_mc_finalize")
        self.T = 0
        return std_map, pred_mean


# ----------------------------------------------------------------------------------------------- A8
def feat_prototype_distance(feat: torch.Tensor, prototype: torch.Tensor, class_numbers: int = 1) -> torch.Tensor:
    """``Trainer.feat_prototype_distance`` (Trainer_prototype.py:98-104): ``[N, class_numbers, H, W]`` with
    ``|| prototype - feat[n,:,h,w] ||_2`` in every class slot (the reference broadcasts one prototype).
    ``prototype`` may also be ``[Q, C]`` with ``Q == class_numbers`` for one distance map per prototype.
    Forward only (no autograd), as used by the reference (pseudo-label rectification)."""
    lib = _lib.load()
    f = _require_cuda_f32(feat.detach(), "feat")
    N, C, H, W = f.shape
    P = prototype.detach().to(device=f.device, dtype=torch.float32).reshape(-1, C).contiguous()
    Q = P.shape[0]
    out = torch.empty(N, Q, H, W, dtype=torch.float32, device=f.device)
    with _on(f.device):
        check(lib.clr_proto_distance(ptr(f), N, C, H * W, ptr(P), Q, ptr(out), _stream()), "clr_proto_distance")
    if Q == 1 and class_numbers > 1:
        out = out.expand(N, class_numbers, H, W).contiguous()
    return out


def distance_weight(feat: torch.Tensor, prototype: torch.Tensor, class_num: int = 1) -> torch.Tensor:
    """``Trainer.get_prototype_weight`` (Trainer_prototype.py:106-116): distance map normalised by its
    global min / max."""
    lib = _lib.load()
    d = feat_prototype_distance(feat, prototype, class_num)
    ws = torch.empty(512, dtype=torch.float32, device=d.device)
    with _on(d.device):
        check(lib.clr_minmax_normalize(ptr(d), d.numel(), ptr(ws), _stream()), "clr_minmax_normalize")
    return d


def get_prototype_weight(feat, class_num, prototype):
    """Drop-in for ``utils.Utils.get_prototype_weight`` (utils/Utils.py:86-88):
    ``cosine_similarity(prototype, feat, dim=1).unsqueeze(1)`` -> ``[N,1,H,W]``.  Forward only."""
    del class_num
    lib = _lib.load()
    f = _require_cuda_f32(feat.detach(), "feat")
    N, C, H, W = f.shape
    P = prototype.detach().to(device=f.device, dtype=torch.float32).reshape(C).contiguous()
    out = torch.empty(N, 1, H, W, dtype=torch.float32, device=f.device)
    ws = torch.empty(4, dtype=torch.float32, device=f.device)
    with _on(f.device):
        check(lib.clr_proto_cosine(ptr(f), N, C, H * W, ptr(P), ptr(ws), ptr(out), _stream()), "clr_proto_cosine")
    return out


def adaptation_factor(m):
    """Drop-in for ``utils.Utils.adaptation_factor`` (utils/Utils.py:104-107); host scalar math."""
    den = 1.0 + math.exp(-0.8 * (m + 1))
    return 1.0 / den - 0.3


# ----------------------------------------------------------------------------------------------- 8(f): step glue
class _SegLoss(torch.autograd.Function):
    """``BCELoss(sigmoid(oS), map) + MSELoss(sigmoid(boundaryS), boundary)`` (Trainer_prototype_full.py:292-294):
    one streaming launch forward, one backward; the upstream gradient stays on the device."""

    @staticmethod
    def forward(ctx, oS, boundaryS, target_map, target_boundary):
        lib = _lib.load()
        ws_bytes = lib.clr_seg_loss_ws_bytes()
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=oS.device)
        out = torch.empty(4, dtype=torch.float32, device=oS.device)
        n2 = 0 if boundaryS is None else boundaryS.numel()
        with _on(oS.device):
            check(lib.clr_seg_loss_fwd(ptr(oS), ptr(target_map), oS.numel(), ptr(boundaryS), ptr(target_boundary), n2,
                                       ptr(ws), ws_bytes, ptr(out), _stream()), "clr_seg_loss_fwd")
        ctx.save_for_backward(oS, boundaryS, target_map, target_boundary)
        ctx.parts = out
        return out[2].clone()

    @staticmethod
    def backward(ctx, gup):
        lib = _lib.load()
        oS, boundaryS, target_map, target_boundary = ctx.saved_tensors
        g1 = torch.empty_like(oS)
        g2 = None if boundaryS is None else torch.empty_like(boundaryS)
        gup = gup.to(torch.float32).contiguous()
        n2 = 0 if boundaryS is None else boundaryS.numel()
        with _on(oS.device):
            check(lib.clr_seg_loss_bwd(ptr(oS), ptr(target_map), oS.numel(), ptr(boundaryS), ptr(target_boundary), n2,
                                       ptr(gup), 1.0, ptr(g1), ptr(g2), _stream()), "clr_seg_loss_bwd")
        return g1, g2, None, None


def seg_loss(oS: torch.Tensor, boundaryS: Optional[torch.Tensor], target_map: torch.Tensor,
             target_boundary: Optional[torch.Tensor]) -> torch.Tensor:
    """``bceloss(sigmoid(oS), target_map) + mseloss(sigmoid(boundaryS), target_boundary)`` of the reference's step
    (Trainer_prototype_full.py:292-294) as one differentiable scalar.  ``boundaryS=None`` -> the BCE term alone."""
    o = _require_cuda_f32(oS, "oS")
    y = _require_cuda_f32(target_map.detach(), "target_map")
    if o.shape != y.shape:
        raise ValueError("oS %s and target_map %s disagree" % (tuple(o.shape), tuple(y.shape)))
    b = t = None
    if boundaryS is not None:
        b = _require_cuda_f32(boundaryS, "boundaryS")
        t = _require_cuda_f32(target_boundary.detach(), "target_boundary")
        if b.shape != t.shape:
            raise ValueError("boundaryS %s and target_boundary %s disagree" % (tuple(b.shape), tuple(t.shape)))
    return _SegLoss.apply(o, b, y, t)


class _EntropyMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, o, smooth):
        lib = _lib.load()
        out = torch.empty_like(o)
        with _on(o.device):
            check(lib.clr_entropy_fwd(ptr(o), o.numel(), float(smooth), ptr(out), _stream()), "clr_entropy_fwd")
        ctx.save_for_backward(o)
        ctx.smooth = float(smooth)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        (o,) = ctx.saved_tensors
        gout = gout.contiguous()
        gin = torch.empty_like(o)
        with _on(o.device):
            check(lib.clr_entropy_bwd(ptr(o), ptr(gout), o.numel(), ctx.smooth, ptr(gin), _stream()), "clr_entropy_bwd")
        return gin, None


def uncertainty_map(o: torch.Tensor, smooth: float = 1e-7) -> torch.Tensor:
    """``-1.0 * torch.sigmoid(o) * torch.log(torch.sigmoid(o) + smooth)`` (Trainer_prototype_full.py:452, 481, 500):
    the entropy map fed to the uncertainty discriminator, one launch each way instead of five."""
    return _EntropyMap.apply(_require_cuda_f32(o, "o", o.dim()), smooth)


# ----------------------------------------------------------------------------------------------- 8(f): validation
def validation_counts(pred_logits: torch.Tensor, target: torch.Tensor, thr: float = 0.75) -> torch.Tensor:
    """Per-class 2x2 confusion counts ``[K, 4]`` (int64, index ``2*gt + pred``) of ``sigmoid(pred) > thr`` against the
    binary ``target`` -- the integers behind ``dice_coeff_2label`` and ``pixel_acc`` (utils/metrics.py:118-168),
    computed on the device in one pass instead of a full-resolution ``.cpu()`` copy plus NumPy."""
    lib = _lib.load()
    z = _require_cuda_f32(pred_logits.detach(), "pred_logits")
    t = _require_cuda_f32(target.detach(), "target")
    if z.shape != t.shape:
        raise ValueError("pred_logits %s and target %s disagree" % (tuple(z.shape), tuple(t.shape)))
    B, K, H, W = z.shape
    counts = torch.empty(K, 4, dtype=torch.int64, device=z.device)
    with _on(z.device):
        check(lib.clr_seg_counts(ptr(z), ptr(t), B, K, H * W, float(thr), ptr(counts), _stream()), "clr_seg_counts")
    return counts


def dice_from_counts(counts: torch.Tensor) -> torch.Tensor:
    """``dice_coefficient_numpy`` (utils/metrics.py:81-100) per class: ``(2 |P&G| + 1) / (1 + |P| + |G|)`` -> ``[K]`` float64."""
    c = counts.to(torch.float64)
    inter, seg, gt = c[:, 3], c[:, 1] + c[:, 3], c[:, 2] + c[:, 3]
    return (2.0 * inter + 1.0) / (1.0 + seg + gt)


def pixel_acc_from_counts(counts: torch.Tensor):
    """``pixelAccuracy`` and ``meanIntersectionOverUnion`` of ``SegmentationMetric(2)`` (utils/metrics.py:10-33) per
    class -> ``(PA [K], mIoU [K])`` float64 (``nanmean`` over the two labels, as the reference)."""
    c = counts.to(torch.float64)
    n00, n01, n10, n11 = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
    pa = (n00 + n11) / c.sum(dim=1)
    iou = torch.stack([n00 / (n00 + n01 + n10), n11 / (n11 + n10 + n01)], dim=1)
    return pa, torch.nanmean(iou, dim=1)
