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
def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


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


def pool_sums(feat: torch.Tensor, w: torch.Tensor, fmt: int, K: int) -> torch.Tensor:
    """Packed class-wise sums ``[2K, C+1]`` (column C = weight sums) of one domain.  Local to this rank."""
    lib = _lib.load()
    B, C, H, W = feat.shape
    HW = H * W
    ws_bytes = lib.clr_pool_ws_bytes(B, C, HW, K)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=feat.device)
    sums = torch.empty(2 * K, C + 1, dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        check(lib.clr_pool_fwd(ptr(feat), ptr(w), fmt, B, C, HW, K, ptr(ws), ws_bytes, ptr(sums), _stream()),
              "clr_pool_fwd")
    return sums


def protos_from_sums(sums: torch.Tensor) -> torch.Tensor:
    """``mu[r][c] = S_r[c] / N_r`` -> ``[R, C]`` (0/0 -> NaN, like utils/Utils.py:127-130)."""
    lib = _lib.load()
    R, C1 = sums.shape
    mu = torch.empty(R, C1 - 1, dtype=torch.float32, device=sums.device)
    with torch.cuda.device(sums.device):
        check(lib.clr_proto_finalize(ptr(sums), R, C1 - 1, ptr(mu), _stream()), "clr_proto_finalize")
    return mu


def pool_backward_feat(w: torch.Tensor, fmt: int, K: int, feat_shape, g: torch.Tensor, sums: torch.Tensor,
                       scale: float = 1.0, xcoef: Optional[torch.Tensor] = None,
                       xtab: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    B, C, H, W = feat_shape
    grad = torch.empty(feat_shape, dtype=torch.float32, device=w.device)
    Kx = 0 if xcoef is None else xcoef.shape[1]
    with torch.cuda.device(w.device):
        check(lib.clr_pool_bwd(ptr(w), fmt, B, C, H * W, K, ptr(g), ptr(sums), float(scale),
                               ptr(xcoef), ptr(xtab), Kx, ptr(grad), _stream()), "clr_pool_bwd")
    return grad


def pool_backward_weights(feat: torch.Tensor, fmt: int, K: int, g: torch.Tensor, sums: torch.Tensor,
                          scale: float = 1.0) -> torch.Tensor:
    lib = _lib.load()
    B, C, H, W = feat.shape
    Q = K if fmt == CLR_W_COMPLEMENT else 2 * K
    ws_bytes = lib.clr_pool_bwd_w_ws_bytes(C, K, fmt)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=feat.device)
    out = torch.empty(B, Q, H, W, dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        check(lib.clr_pool_bwd_w(ptr(feat), fmt, B, C, H * W, K, ptr(g), ptr(sums), float(scale),
                                 ptr(ws), ws_bytes, ptr(out), _stream()), "clr_pool_bwd_w")
    return out


def _stack_grads(grads: Sequence[Optional[torch.Tensor]], R: int, C: int, device) -> torch.Tensor:
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
    """``mu_r = sum x w_r / sum w_r`` for one or two concatenated domains.

    inputs: (fmt, K, n_dom, w_0, feat_0[, w_1, feat_1]); outputs: 2K tensors ``[1,C,1,1]``.
    With two domains the sums are added before the divide (utils/Utils.py:132-158, :227-311).
    """

    @staticmethod
    def forward(ctx, fmt: int, K: int, *tensors):
        n_dom = len(tensors) // 2
        ws, feats = tensors[0::2], tensors[1::2]
        sums = None
        for w, f in zip(ws, feats):
            s = pool_sums(f, w, fmt, K)
            sums = s if sums is None else sums + s
        scale = 1.0
        if _dist.enabled():
            _dist.all_reduce_sums(sums)
            scale = _dist.grad_scale()
        mu = protos_from_sums(sums)
        ctx.fmt, ctx.K, ctx.n_dom, ctx.scale = fmt, K, n_dom, scale
        ctx.shapes = [tuple(f.shape) for f in feats]
        need_w = [ctx.needs_input_grad[2 + 2 * i] for i in range(n_dom)]
        ctx.save_for_backward(sums, *ws, *[f if nw else None for f, nw in zip(feats, need_w)])
        return _split_protos(mu)

    @staticmethod
    def backward(ctx, *grads):
        saved = ctx.saved_tensors
        sums, ws, feats = saved[0], saved[1:1 + ctx.n_dom], saved[1 + ctx.n_dom:]
        K, fmt = ctx.K, ctx.fmt
        C = ctx.shapes[0][1]
        g = _stack_grads(grads, 2 * K, C, sums.device)
        out = [None, None]
        for i in range(ctx.n_dom):
            gw = gf = None
            if ctx.needs_input_grad[2 + 2 * i]:
                gw = pool_backward_weights(feats[i], fmt, K, g, sums, ctx.scale)
            if ctx.needs_input_grad[3 + 2 * i]:
                gf = pool_backward_feat(ws[i], fmt, K, ctx.shapes[i], g, sums, ctx.scale)
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
    return _WeightedPrototypes.apply(CLR_W_COMPLEMENT, K, pred, feat)


def gen_prototype_src_trg(pred_oS, xs_feature, pred_oT, xt_feature):
    """Drop-in for ``utils.Utils.gen_prototype_src_trg`` (utils/Utils.py:132-158): joint source+target
    prototypes ``(S_s+S_t)/(N_s+N_t)`` without materialising the concatenation."""
    ps, fs = _require_cuda_f32(pred_oS, "pred_oS"), _require_cuda_f32(xs_feature, "xs_feature")
    pt, ft = _require_cuda_f32(pred_oT, "pred_oT"), _require_cuda_f32(xt_feature, "xt_feature")
    K = ps.shape[1]
    _check_k(K)
    if pt.shape[1] != K or fs.shape[1] != ft.shape[1]:
        raise ValueError("source and target disagree on K or C")
    return _WeightedPrototypes.apply(CLR_W_COMPLEMENT, K, ps, fs, pt, ft)


def weighted_prototypes(weights: torch.Tensor, feat: torch.Tensor):
    """Prototypes from explicit weight planes ``[B,2K,H,W]`` (rows obj_0.., bck_0..)."""
    w = _require_cuda_f32(weights, "weights")
    f = _require_cuda_f32(feat, "feat")
    if w.shape[1] % 2:
        raise ValueError("explicit weights need 2K planes")
    K = w.shape[1] // 2
    _check_k(K)
    return _WeightedPrototypes.apply(CLR_W_EXPLICIT, K, w, f)


def adaptation_factor(m):
    """Drop-in for ``utils.Utils.adaptation_factor`` (utils/Utils.py:104-107); host scalar math."""
    den = 1.0 + math.exp(-0.8 * (m + 1))
    return 1.0 / den - 0.3
