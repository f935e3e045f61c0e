#!/usr/bin/env python3
"""Generate ``tests/golden/transnorm_golden.npz`` by running the UNMODIFIED reference TransNorm module
(``networks.sync_batchnorm.batchnorm.BatchNorm2d``, imported in place from ``/root/reference``; dev container only)
under torch CPU fp32: two training steps (outputs, input / weight / bias gradients, running estimates after each step)
and one eval forward per case.  ``tests/test_golden.py`` (oracle) and ``tests/test_gpu_transnorm.py`` (CUDA path) compare
against it on the GPU box, where the reference tree does not exist.

    python tests/golden/make_transnorm_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "transnorm_golden.npz")

CASES = {           # name: (B, C, H, W, seed)
    "tn_small": (4, 6, 5, 7, 11),         # ragged planes (35 pixels): scalar path
    "tn_vec": (6, 12, 8, 16, 12),          # 128-bit path (HW = 128), halves of 3
    "tn_odd_batch": (5, 9, 4, 8, 13),     # B odd: source 2 samples, target 3 (batchnorm.py:452-454)
}


def inputs(B, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g) * (0.5 + torch.rand(1, C, 1, 1, generator=g) * 2.0) \
        + torch.randn(1, C, 1, 1, generator=g) * 1.5
    x[B // 2:] += torch.randn(1, C, 1, 1, generator=g) * 0.7          # domain shift
    x2 = x + 0.3 * torch.randn(B, C, H, W, generator=g)
    w = 0.5 + torch.rand(C, generator=g)
    b = 0.3 * torch.randn(C, generator=g)
    gy = torch.randn(B, C, H, W, generator=g)
    return x, x2, w, b, gy


def main():
    if not ref_import.available():
        raise SystemExit("reference tree not present")
    cls = ref_import.ref_transnorm_class()
    store = {}
    for name, (B, C, H, W, seed) in CASES.items():
        x, x2, w, b, gy = inputs(B, C, H, W, seed)
        m = cls(C)
        with torch.no_grad():
            m.weight.copy_(w)
            m.bias.copy_(b)
        m.train()
        store[name + "/in_x"], store[name + "/in_x2"] = x.numpy(), x2.numpy()
        store[name + "/in_weight"], store[name + "/in_bias"], store[name + "/seed_gy"] = w.numpy(), b.numpy(), gy.numpy()
        for step, xin in enumerate((x, x2)):
            xr = xin.clone().requires_grad_(True)
            m.zero_grad()
            y = m(xr)
            (y * gy).sum().backward()
            store["%s/out_y%d" % (name, step)] = y.detach().numpy()
            store["%s/grad_x%d" % (name, step)] = xr.grad.numpy()
            store["%s/grad_weight%d" % (name, step)] = m.weight.grad.numpy().copy()
            store["%s/grad_bias%d" % (name, step)] = m.bias.grad.numpy().copy()
            for buf in ("running_mean_source", "running_var_source", "running_mean_target", "running_var_target"):
                store["%s/out_%s%d" % (name, buf, step)] = getattr(m, buf).numpy().copy()
        m.eval()
        with torch.no_grad():
            store[name + "/out_eval"] = m(x).numpy()
    np.savez_compressed(OUT, **store)
    print("wrote %s: %d arrays, %.1f KB" % (OUT, len(store), os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
