#!/usr/bin/env python3
"""Generate ``tests/golden/ref_golden.npz`` by running the UNMODIFIED reference (dev container only).

Every array named ``<case>/out_*`` or ``<case>/grad_*`` below is produced by code imported in place from
``/root/reference`` (``utils/Utils.py``) under torch CPU fp32; ``<case>/in_*`` are the seeded inputs and
``<case>/seed_*`` the upstream gradients fed to ``backward``.  The file travels to the GPU box, where
``tests/test_golden.py`` (oracle) and ``tests/test_gpu_parity.py`` (CUDA path) compare against it.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from uda_clr_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.npz")


def stack(protos):
    return torch.cat([p.reshape(1, -1) for p in protos], 0).detach().numpy()


def case_gen_prototype(store, name, B, C, H, W, soft, seed):
    g = torch.Generator().manual_seed(seed)
    y = synth.nested_ellipse_labels(B, 2, H, W, g)
    x = synth.class_shifted_features(y, C, g)
    pred = torch.sigmoid(synth.confident_logits(y, g)) if soft else y
    xr = x.clone().requires_grad_(True)
    pr = pred.clone().requires_grad_(soft)
    out = ref_import.ref_gen_prototype(pr, xr)
    seeds = torch.randn(4, C, generator=g)
    sum((o.reshape(-1) * s).sum() for o, s in zip(out, seeds)).backward()
    store[name + "/in_pred"] = pred.numpy()
    store[name + "/in_feat"] = x.numpy()
    store[name + "/out_protos"] = stack(out)
    store[name + "/seed_g"] = seeds.numpy()
    store[name + "/grad_feat"] = xr.grad.numpy()
    if soft:
        store[name + "/grad_pred"] = pr.grad.numpy()


def case_src_trg(store, name, seed):
    b = synth.make_batch(B=2, C=6, H=8, W=8, K=2, image_res=False, seed=seed)
    pt = torch.sigmoid(b.oT_before)
    out = ref_import.ref_gen_prototype_src_trg(b.ys, b.xs, pt, b.xt)
    store[name + "/in_pred_s"] = b.ys.numpy()
    store[name + "/in_feat_s"] = b.xs.numpy()
    store[name + "/in_pred_t"] = pt.numpy()
    store[name + "/in_feat_t"] = b.xt.numpy()
    store[name + "/out_protos"] = stack(out)


def case_retrify(store, name, seed, B=1, C=3, T=4, Hi=160):
    g = torch.Generator().manual_seed(seed)
    H = W = 128  # hard-coded in the reference (utils/Utils.py:162)
    yt = synth.nested_ellipse_labels(B, 2, H, W, g)
    xt = synth.class_shifted_features(yt, C, g)
    oT = synth.confident_logits(yt, g)
    base = torch.nn.functional.interpolate(oT, size=(Hi, Hi), mode="nearest")
    # half precision keeps the fixture small; the values are exactly representable in fp32
    preds = (base.repeat(T, 1, 1, 1) + 0.35 * torch.randn(T * B, 2, Hi, Hi, generator=g)).half().float()
    xr = xt.clone().requires_grad_(True)
    out = ref_import.ref_gen_prototype_retrify(oT, xr, preds, T, B)
    seeds = torch.randn(4, C, generator=g)
    sum((o.reshape(-1) * s).sum() for o, s in zip(out[:4], seeds)).backward()
    store[name + "/in_oT_before"] = oT.numpy()
    store[name + "/in_xt"] = xt.numpy()
    store[name + "/in_preds_f16"] = preds.half().numpy()
    store[name + "/in_T"] = np.array(T)
    store[name + "/out_protos"] = stack(out[:4])
    store[name + "/out_std_map"] = out[4].numpy()
    store[name + "/out_mask_0"] = out[5].numpy().astype(np.uint8)
    store[name + "/out_mask_1"] = out[6].numpy().astype(np.uint8)
    store[name + "/seed_g"] = seeds.numpy()
    store[name + "/grad_xt"] = xr.grad.numpy()


def case_cosine(store, name, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(2, 9, 5, 6, generator=g)
    p = torch.randn(1, 9, 1, 1, generator=g)
    store[name + "/in_feat"] = x.numpy()
    store[name + "/in_proto"] = p.numpy()
    store[name + "/out_weight"] = ref_import.ref_get_prototype_weight(x, 1, p).numpy()
    ms = np.array([-1.0, 0.0, 3.0, 25.0, 100.5])
    store[name + "/in_m"] = ms
    store[name + "/out_adaptation_factor"] = np.array([ref_import.ref_adaptation_factor(float(m)) for m in ms])


def main():
    assert ref_import.available(), "needs /root/reference"
    torch.set_num_threads(1)  # fixed reduction order for the recorded fp32 outputs
    store = {}
    case_gen_prototype(store, "hard_ragged", B=2, C=5, H=6, W=7, soft=False, seed=101)
    case_gen_prototype(store, "hard_c305", B=1, C=305, H=16, W=16, soft=False, seed=102)
    case_gen_prototype(store, "soft", B=2, C=8, H=8, W=8, soft=True, seed=103)
    case_gen_prototype(store, "hard_big", B=3, C=33, H=32, W=32, soft=False, seed=104)
    case_src_trg(store, "src_trg", seed=105)
    case_retrify(store, "retrify", seed=106)
    case_cosine(store, "cosine", seed=107)
    store["meta/torch_version"] = np.array(torch.__version__)
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(store), "arrays")


if __name__ == "__main__":
    main()
