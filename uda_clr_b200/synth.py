"""Seeded synthetic inputs for the CLR hot path.

The reference ships no data and no fixtures; every parity test and the bench
draw their inputs here so the oracle, the CUDA path and the CPU baseline see
the same tensors.  Everything is generated on the CPU with an explicit
``torch.Generator`` (bit-stable for a given torch build) and moved by the
caller.

Label semantics mirror the reference's ``to_multilabel`` nesting
(dataloaders/custom_transforms.py:15-19): channel 0 (cup) is a subset of
channel 1 (disc); for K > 2 the classes stay nested, innermost first.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

LN3 = math.log(3.0)  # sigmoid(x) > 0.75  <=>  x > ln 3


def nested_ellipse_labels(B: int, K: int, H: int, W: int, gen: torch.Generator) -> torch.Tensor:
    """Hard labels ``[B,K,H,W]`` in {0,1}; class k is nested inside class k+1; every class non-empty."""
    yy = torch.arange(H, dtype=torch.float32).view(1, H, 1)
    xx = torch.arange(W, dtype=torch.float32).view(1, 1, W)
    cy = (0.4 + 0.2 * torch.rand(B, generator=gen)).view(B, 1, 1) * H
    cx = (0.4 + 0.2 * torch.rand(B, generator=gen)).view(B, 1, 1) * W
    r_out = (0.25 + 0.10 * torch.rand(B, generator=gen)).view(B, 1, 1) * min(H, W)
    ratio = (0.4 + 0.3 * torch.rand(B, generator=gen)).view(B, 1, 1)
    asp = (0.85 + 0.3 * torch.rand(B, generator=gen)).view(B, 1, 1)
    d2 = ((yy - cy) / asp) ** 2 + ((xx - cx) * asp) ** 2
    out = torch.zeros(B, K, H, W, dtype=torch.float32)
    for k in range(K):
        # radius shrinks geometrically from the outermost class (k = K-1) to the innermost (k = 0)
        r = r_out * ratio ** (K - 1 - k)
        out[:, k] = (d2 <= r * r).float()
    # guarantee non-empty classes even for tiny planes
    iy = cy.long().clamp(0, H - 1).view(B)
    ix = cx.long().clamp(0, W - 1).view(B)
    out[torch.arange(B), :, iy, ix] = 1.0
    return out


def class_shifted_features(labels: torch.Tensor, C: int, gen: torch.Generator) -> torch.Tensor:
    """``randn(B,C,H,W) + m[c] * y_outer`` with ``m = linspace(-1,1,C)`` so class means differ."""
    B, K, H, W = labels.shape
    m = torch.linspace(-1.0, 1.0, C).view(1, C, 1, 1)
    x = torch.randn(B, C, H, W, generator=gen)
    x += m * labels[:, K - 1:K]
    if K > 1:
        x += 0.5 * m.flip(1) * labels[:, 0:1]
    return x


def confident_logits(labels: torch.Tensor, gen: torch.Generator, guard: float = 1e-4) -> torch.Tensor:
    """Logits ``2*randn + 3*(2y-1)``; no value within ``guard`` of ln 3 so the 0.75 threshold is unambiguous."""
    z = 2.0 * torch.randn(labels.shape, generator=gen) + 3.0 * (2.0 * labels - 1.0)
    near = (z - LN3).abs() < guard
    z[near] += 4.0 * guard
    return z


def upsample_nearest_int(x: torch.Tensor, f: int) -> torch.Tensor:
    return x.repeat_interleave(f, dim=2).repeat_interleave(f, dim=3)


@dataclass
class ClrBatch:
    """One synthetic CLR step worth of inputs (all CPU fp32)."""
    ys: torch.Tensor          # [B,K,H,W] hard source labels (pred_oS)
    xs: torch.Tensor          # [B,C,H,W] source decoder features
    yt: torch.Tensor          # [B,K,H,W] hidden target labels (only used to draw logits)
    xt: torch.Tensor          # [B,C,H,W] target decoder features
    oT_before: torch.Tensor   # [B,K,H,W] target logits at feature resolution
    preds: Optional[torch.Tensor]   # [T*B,K,Hi,Wi] MC-dropout logits at image resolution
    oT: Optional[torch.Tensor]      # [B,K,Hi,Wi] target logits at image resolution
    oT_aug: Optional[torch.Tensor]  # [B,K,Hi,Wi] logits of the photometrically augmented view
    T: int
    up: int


def make_batch(B: int = 8, C: int = 256, H: int = 128, W: int = 128, K: int = 2, T: int = 8,
               up: int = 4, seed: int = 1234, image_res: bool = True) -> ClrBatch:
    g = torch.Generator().manual_seed(seed)
    ys = nested_ellipse_labels(B, K, H, W, g)
    yt = nested_ellipse_labels(B, K, H, W, g)
    xs = class_shifted_features(ys, C, g)
    xt = class_shifted_features(yt, C, g)
    oT_before = confident_logits(yt, g)
    preds = oT = oT_aug = None
    if image_res:
        base = upsample_nearest_int(oT_before, up)
        preds = base.repeat(T, 1, 1, 1) + 0.3 * torch.randn(T * B, K, H * up, W * up, generator=g)
        oT = base + 0.1 * torch.randn(B, K, H * up, W * up, generator=g)
        oT_aug = base + 0.5 * torch.randn(B, K, H * up, W * up, generator=g)
    return ClrBatch(ys=ys, xs=xs, yt=yt, xt=xt, oT_before=oT_before, preds=preds, oT=oT, oT_aug=oT_aug,
                    T=T, up=up)
