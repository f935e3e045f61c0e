"""TransNorm: host-side mirror of the reference's domain-split batch normalisation.

``TransNorm2d`` mirrors ``networks.sync_batchnorm.batchnorm.BatchNorm2d`` (networks/sync_batchnorm/batchnorm.py:264-388
``_NormBase``, :390-521 ``_BatchNorm.forward``, :523 ``BatchNorm2d``) -- the class ``DeepLab`` builds its backbone,
ASPP and decoder with when ``sync_bn=False`` (``--use_TN``; networks/deeplabv3.py:17-27): same constructor arguments,
parameter and buffer names (``weight``, ``bias``, ``running_mean_source``, ``running_var_source``,
``running_mean_target``, ``running_var_target``, ``num_batches_tracked``), so state dicts are interchangeable.

The arithmetic runs in ``libclr_b200.so`` (``clr_tn_fwd`` / ``clr_tn_bwd`` / ``clr_tn_eval``: three launches each,
one pass for the statistics and one for the normalisation) instead of two cuDNN batch norms, two transposed copies,
four reductions and a dozen small launches.  There is no CPU path.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _lib
from ._lib import check, ptr


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _geom(x: torch.Tensor):
    if x.dim() == 4:
        B, C, H, W = x.shape
        return B, C, H * W
    if x.dim() == 2:
        return x.shape[0], x.shape[1], 1
    raise ValueError("expected 4D input (got {}D input)".format(x.dim()))


def _require(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("TransNorm input must be a CUDA tensor: there is no CPU fallback")
    if x.dtype != torch.float32:
        raise TypeError("TransNorm input must be float32 (got %s)" % x.dtype)
    return x.contiguous()


class _TransNormFn(torch.autograd.Function):
    """inputs: (x, weight, bias, rm_s, rv_s, rm_t, rv_t, factor, eps, training); the running estimates are updated in
    place by the forward kernel (training) or read (eval)."""

    @staticmethod
    def forward(ctx, x, weight, bias, rm_s, rv_s, rm_t, rv_t, factor, eps, training):
        lib = _lib.load()
        B, C, HW = _geom(x)
        ws_bytes = lib.clr_tn_ws_bytes(C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        y = torch.empty_like(x)
        save = torch.empty(5, C, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            if training:
                check(lib.clr_tn_fwd(ptr(x), B, C, HW, ptr(weight), ptr(bias), ptr(rm_s), ptr(rv_s), ptr(rm_t), ptr(rv_t),
                                     float(factor), float(eps), ptr(ws), ws_bytes, ptr(y), ptr(save), _stream()), "clr_tn_fwd")
            else:
                check(lib.clr_tn_eval(ptr(x), B, C, HW, ptr(weight), ptr(bias), ptr(rm_s), ptr(rv_s), ptr(rm_t), ptr(rv_t),
                                      float(eps), ptr(ws), ws_bytes, ptr(y), ptr(save), _stream()), "clr_tn_eval")
        ctx.geom, ctx.training = (B, C, HW), bool(training)
        ctx.save_for_backward(x, weight, save)
        return y

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x, weight, save = ctx.saved_tensors
        B, C, HW = ctx.geom
        gy = gy.contiguous()
        ws_bytes = lib.clr_tn_ws_bytes(C)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        gx = torch.empty_like(x)
        need_w = weight is not None and ctx.needs_input_grad[1]
        need_b = ctx.needs_input_grad[2]
        gw = torch.empty(C, dtype=torch.float32, device=x.device) if need_w else None
        gb = torch.empty(C, dtype=torch.float32, device=x.device) if need_b else None
        with torch.cuda.device(x.device):
            check(lib.clr_tn_bwd(ptr(x), ptr(gy), B, C, HW, ptr(weight), ptr(save), 0 if ctx.training else 1,
                                 ptr(ws), ws_bytes, ptr(gx), ptr(gw), ptr(gb), _stream()), "clr_tn_bwd")
        return gx, gw, gb, None, None, None, None, None, None, None


def trans_norm(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor],
               running_mean_source: Optional[torch.Tensor], running_var_source: Optional[torch.Tensor],
               running_mean_target: Optional[torch.Tensor], running_var_target: Optional[torch.Tensor],
               training: bool, momentum: float = 0.1, eps: float = 1e-5) -> torch.Tensor:
    """Functional form of ``_BatchNorm.forward`` (networks/sync_batchnorm/batchnorm.py:439-521).

    training: ``x[:B//2]`` (source) and ``x[B//2:]`` (target) are normalised with their own batch statistics, the running
    estimates (if given) are updated in place with ``momentum``, and every channel is scaled by ``1 + alpha``.
    eval: normalisation with the target running estimates, ``alpha`` from both sets."""
    x = _require(x)
    B, C, HW = _geom(x)
    for name, t in (("weight", weight), ("bias", bias), ("running_mean_source", running_mean_source),
                    ("running_var_source", running_var_source), ("running_mean_target", running_mean_target),
                    ("running_var_target", running_var_target)):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32 or t.numel() != C or not t.is_contiguous()):
            raise ValueError("%s must be a contiguous float32 CUDA tensor with %d elements" % (name, C))
    if training:
        if B < 2:
            raise ValueError("TransNorm in training mode needs a batch of at least 2 (source half, target half)")
        if (B // 2) * HW < 2:
            # F.batch_norm (batchnorm.py:455) raises the same for a half with a single value per channel
            raise ValueError("Expected more than 1 value per channel when training, got input size {}".format(list(x.shape)))
    elif None in (running_mean_source, running_var_source, running_mean_target, running_var_target):
        raise ValueError("TransNorm in eval mode needs the four running estimates (batchnorm.py:495)")
    return _TransNormFn.apply(x, weight, bias, running_mean_source, running_var_source, running_mean_target,
                              running_var_target, momentum, eps, training)


class _TransNormBase(nn.Module):
    """``_NormBase`` + ``_BatchNorm`` of the reference (batchnorm.py:264-521)."""
    _version = 2
    __constants__ = ["track_running_stats", "momentum", "eps", "num_features", "affine"]

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: Optional[float] = 0.1, affine: bool = True,
                 track_running_stats: bool = True, device=None, dtype=None) -> None:
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        self.num_features = num_features
        self.eps = eps
        self.momentum = momentum
        self.affine = affine
        self.track_running_stats = track_running_stats
        if affine:
            self.weight = nn.Parameter(torch.empty(num_features, **factory_kwargs))
            self.bias = nn.Parameter(torch.empty(num_features, **factory_kwargs))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        for dom in ("source", "target"):
            if track_running_stats:
                self.register_buffer("running_mean_" + dom, torch.zeros(num_features, **factory_kwargs))
                self.register_buffer("running_var_" + dom, torch.ones(num_features, **factory_kwargs))
            else:
                self.register_buffer("running_mean_" + dom, None)
                self.register_buffer("running_var_" + dom, None)
        if track_running_stats:
            self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long, device=device))
        else:
            self.register_buffer("num_batches_tracked", None)
        self.reset_parameters()

    def reset_running_stats(self) -> None:
        if self.track_running_stats:
            self.running_mean_source.zero_()
            self.running_var_source.fill_(1)
            self.running_mean_target.zero_()
            self.running_var_target.fill_(1)
            self.num_batches_tracked.zero_()

    def reset_parameters(self) -> None:
        self.reset_running_stats()
        if self.affine:
            nn.init.ones_(self.weight)
            nn.init.zeros_(self.bias)

    def _check_input_dim(self, input):
        raise NotImplementedError

    def extra_repr(self):
        return ("{num_features}, eps={eps}, momentum={momentum}, affine={affine}, "
                "track_running_stats={track_running_stats}".format(**self.__dict__))

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        self._check_input_dim(input)
        factor = 0.0 if self.momentum is None else self.momentum
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
            if self.momentum is None:   # cumulative moving average (host sync, like the reference :425)
                factor = 1.0 / float(self.num_batches_tracked)
        if not self.training and self.running_mean_source is None:
            raise AssertionError("eval mode needs tracked running statistics (batchnorm.py:495)")
        return trans_norm(input, self.weight, self.bias, self.running_mean_source, self.running_var_source,
                          self.running_mean_target, self.running_var_target, self.training, factor, self.eps)


class TransNorm2d(_TransNormBase):
    """Drop-in for ``networks.sync_batchnorm.batchnorm.BatchNorm2d`` (batchnorm.py:523)."""

    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError("expected 4D input (got {}D input)".format(input.dim()))


class TransNorm1d(_TransNormBase):
    """The ``input.dim() == 2`` branch of the same forward (batchnorm.py:489-490): ``[N, C]`` inputs."""

    def _check_input_dim(self, input):
        if input.dim() != 2:
            raise ValueError("expected 2D input (got {}D input)".format(input.dim()))


BatchNorm2d = TransNorm2d   # the reference's name for it
