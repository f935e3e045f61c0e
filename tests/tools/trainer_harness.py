#!/usr/bin/env python3
"""Run the UNMODIFIED reference trainer (``train_process/Trainer_prototype_full.Trainer.train_epoch``) for a few steps on
synthetic fundus-shaped data, stock or with the CLR ops patched in, and record what the prototype ops returned.

TEST INFRASTRUCTURE ("whole-trainer oracle").  Needs the reference tree (``UDA_CLR_REFERENCE`` or
/root/reference); nothing is copied from it, its modules are imported in place with the import-time stubs of
``oracle/ref_import.py``.  On a CPU-only host ``.cuda()`` is shimmed to the identity (then only ``patched="port"`` -- the
eager port standing in for the CUDA ops -- can be compared with stock: it validates this harness and the seam).

    python tests/tools/trainer_harness.py [--steps 2] [--batch 1] [--patched none|clr|port]
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import sys
import tempfile
import types
from unittest.mock import MagicMock

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _import_trainer(ref_root):
    from oracle import ref_import
    for name in ref_import._STUBS:
        if name != "pytz":
            sys.modules.setdefault(name, MagicMock())
    # a MagicMock time zone breaks datetime.now(tz) (Trainer_prototype_full.py:57-58): a real module with a real tzinfo
    pytz = types.ModuleType("pytz")
    pytz.timezone = lambda name: datetime.timezone.utc
    sys.modules["pytz"] = pytz
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import train_process.Trainer_prototype_full as mod
    import networks.deeplabv3 as deeplab
    import networks.GAN as gan
    return mod, deeplab, gan


def _synthetic_loader(steps, B, seed, size=512):
    """dict(image [B,3,S,S] in [-1,1], map [B,2,S,S] nested ellipses, boundary [B,1,S,S] in [0,1], img_name)."""
    import torch
    from uda_clr_b200 import synth
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(steps):
        m = synth.nested_ellipse_labels(B, 2, size, size, g)
        img = (torch.rand(B, 3, size, size, generator=g) * 2 - 1) * 0.5 + 0.4 * (m[:, 1:2] - 0.5) + 0.3 * m[:, 0:1]
        ring = (m[:, 1:2] - torch.nn.functional.avg_pool2d(m[:, 1:2], 9, 1, 4)).abs()
        out.append(dict(image=img.clamp(-1, 1), map=m, boundary=(ring / ring.amax().clamp_min(1e-6)), img_name=["synthetic"] * B))
    return out


def run_train_epoch(ref_root, patched=False, steps=2, batch_size=1, seed=1337, backbone="mobilenet"):
    """``patched``: False / "none" (stock), "clr" (``uda_clr_b200.patch_reference()``: needs a GPU), "port" (the eager port
    bound at the same seam).  Returns dict(calls=[{name, protos:[np arrays]}...], running_intra, running_inter)."""
    import numpy as np
    import torch
    mod, deeplab, gan = _import_trainer(ref_root)
    use_gpu = torch.cuda.is_available()
    saved = {}
    if not use_gpu:
        saved["t"], saved["m"] = torch.Tensor.cuda, torch.nn.Module.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    import networks.backbone.mobilenet as mb
    mb.MobileNetV2._load_pretrained_model = lambda self: None
    orig = {n: getattr(mod, n) for n in ("gen_prototype", "gen_prototype_retrify")}
    calls = []
    try:
        torch.manual_seed(seed)
        np.random.seed(seed)
        model = deeplab.DeepLab(num_classes=2, backbone=backbone, output_stride=16, sync_bn=True, freeze_bn=False)
        dis, dis2 = gan.BoundaryDiscriminator(), gan.UncertaintyDiscriminator()
        if use_gpu:
            model, dis, dis2 = model.cuda(), dis.cuda(), dis2.cuda()
        og = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.99))
        od = torch.optim.SGD(dis.parameters(), lr=2.5e-5, momentum=0.99, weight_decay=5e-4)
        od2 = torch.optim.SGD(dis2.parameters(), lr=2.5e-5, momentum=0.99, weight_decay=5e-4)
        if patched == "clr":
            import uda_clr_b200 as clr
            clr.patch_reference()
        elif patched == "port":
            from oracle import clr_torch_port as TP
            mod.gen_prototype = TP.gen_prototype
            mod.gen_prototype_retrify = lambda o, x, p, f, T, s: TP.gen_prototype_retrify(o, x, p, None, T, s)

        def record(name):
            fn = getattr(mod, name)

            def wrapped(*a, **k):
                out = fn(*a, **k)
                calls.append(dict(name=name, protos=[t.detach().float().cpu().numpy().copy() for t in out[:4]]))
                return out
            return wrapped
        for n in ("gen_prototype", "gen_prototype_retrify"):
            setattr(mod, n, record(n))
        loaderS = _synthetic_loader(steps, batch_size, seed + 1)
        loaderT = _synthetic_loader(steps, batch_size, seed + 2)
        tr = mod.Trainer(cuda=use_gpu, model_gen=model, model_dis=dis, model_uncertainty_dis=dis2, optimizer_gen=og,
                         optimizer_dis=od, optimizer_uncertainty_dis=od2, val_loader=[], domain_loaderS=loaderS,
                         domain_loaderT=loaderT, out=tempfile.mkdtemp(prefix="clr_trainer_"), max_epoch=1, use_global=True,
                         use_pid=True, retrify_pesudo=True, global_pro_weight=0.9, pro_weight=0.1, batch_size=batch_size,
                         warmup_epoch=-1)
        tr.epoch, tr.iteration = 0, 0
        tr.writer = MagicMock()
        torch.manual_seed(seed + 7)          # the MC-dropout passes draw from the global generator
        tr.train_epoch()
        return dict(calls=calls, running_intra=float(tr.running_intra), running_inter=float(tr.running_inter))
    finally:
        for n, f in orig.items():
            setattr(mod, n, f)
        if patched == "clr":
            import uda_clr_b200 as clr
            if hasattr(clr, "unpatch_reference"):
                clr.unpatch_reference()
        if not use_gpu:
            torch.Tensor.cuda, torch.nn.Module.cuda = saved["t"], saved["m"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("UDA_CLR_REFERENCE", "/root/reference"))
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--patched", default="port", choices=["none", "clr", "port"])
    a = ap.parse_args()
    import numpy as np
    stock = run_train_epoch(a.ref, False, a.steps, a.batch)
    other = run_train_epoch(a.ref, a.patched if a.patched != "none" else False, a.steps, a.batch)
    rows = []
    for x, y in zip(stock["calls"], other["calls"]):
        err = max(float(np.abs(p - q).max() / max(np.abs(p).max(), 1e-30)) for p, q in zip(x["protos"], y["protos"]))
        rows.append(dict(name=x["name"], max_rel_err=err))
    print(json.dumps(dict(patched=a.patched, steps=a.steps, batch=a.batch, calls=rows,
                          stock=dict(intra=stock["running_intra"], inter=stock["running_inter"]),
                          other=dict(intra=other["running_intra"], inter=other["running_inter"])), indent=1))


if __name__ == "__main__":
    main()
