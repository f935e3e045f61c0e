"""Rebind the reference's CLR functions to the B200 ops inside an imported reference tree.

The reference pulls its CLR functions in with ``from utils.Utils import *``
(train_process/Trainer_prototype_full.py:16, Trainer_prototype.py:18, cal_prototype.py:27), which COPIES the
names into each trainer module at import time.  Swapping the implementation therefore means rebinding the
name in ``utils.Utils`` *and* in every already-imported module that star-imported it -- the three call sites
(Trainer_prototype_full.py:332-334, 370-373, 375-377) resolve the name in the trainer module's globals.

    import utils.Utils, train_process.Trainer_prototype_full      # the reference, unmodified
    import uda_clr_b200
    uda_clr_b200.patch_reference()                                  # trainers now run the sm_100a kernels
    ...
    uda_clr_b200.unpatch_reference()
"""
from __future__ import annotations

import sys
from typing import Dict, List, Tuple

from . import ops

#: reference name -> replacement (identical positional signature and return tuple)
REPLACEMENTS = {
    "gen_prototype": ops.gen_prototype,
    "gen_prototype_retrify": ops.gen_prototype_retrify,
    "gen_prototype_src_trg": ops.gen_prototype_src_trg,
    "gen_prototype_src_trg_retrify": ops.gen_prototype_src_trg_retrify,
    "get_prototype_weight": ops.get_prototype_weight,
    "adaptation_factor": ops.adaptation_factor,
}

_saved: List[Tuple[object, str, object]] = []


def _reference_modules():
    """``utils.Utils`` of the reference and every loaded module that holds a star-import copy of its names."""
    utils_mod = sys.modules.get("utils.Utils")
    if utils_mod is None or not hasattr(utils_mod, "gen_prototype_retrify"):
        raise RuntimeError("the reference's utils.Utils is not imported (put the reference tree on sys.path "
                           "and `import utils.Utils` first)")
    mods = [utils_mod]
    for name, mod in list(sys.modules.items()):
        if mod is None or mod is utils_mod or name.startswith("uda_clr_b200"):
            continue
        d = getattr(mod, "__dict__", None)
        if not d:
            continue
        if any(d.get(n) is getattr(utils_mod, n, object()) for n in REPLACEMENTS):
            mods.append(mod)
    return mods


def patch_reference() -> Dict[str, List[str]]:
    """Rebind the six CLR names everywhere they were copied.  Returns ``{name: [module names patched]}``."""
    if any(n in REPLACEMENTS for _, n, _o in _saved):
        raise RuntimeError("reference already patched; call unpatch_reference() first")
    utils_mod = sys.modules.get("utils.Utils")
    report: Dict[str, List[str]] = {n: [] for n in REPLACEMENTS}
    mods = _reference_modules()
    originals = {n: getattr(utils_mod, n) for n in REPLACEMENTS if hasattr(utils_mod, n)}
    for mod in mods:
        for n, orig in originals.items():
            if mod.__dict__.get(n) is orig:
                _saved.append((mod, n, orig))
                setattr(mod, n, REPLACEMENTS[n])
                report[n].append(mod.__name__)
    return report


def patch_transnorm() -> List[str]:
    """Rebind the reference's TransNorm class (``networks.sync_batchnorm.batchnorm.BatchNorm2d``, the ``BatchNorm`` that
    ``DeepLab`` picks with ``sync_bn=False`` / ``--use_TN``, networks/deeplabv3.py:4, 17-23) to :class:`TransNorm2d` in
    every loaded module that imported it by name.  Call before the model is constructed; undone by
    :func:`unpatch_reference`.  Returns the names of the modules patched."""
    from .transnorm import TransNorm2d
    ref_bn = sys.modules.get("networks.sync_batchnorm.batchnorm")
    if ref_bn is None or not hasattr(ref_bn, "BatchNorm2d"):
        raise RuntimeError("the reference's networks.sync_batchnorm.batchnorm is not imported")
    orig = ref_bn.BatchNorm2d
    if orig is TransNorm2d:
        return []
    patched = []
    for name, mod in list(sys.modules.items()):
        if mod is None or name.startswith("uda_clr_b200"):
            continue
        d = getattr(mod, "__dict__", None)
        if d and d.get("BatchNorm2d") is orig:
            _saved.append((mod, "BatchNorm2d", orig))
            setattr(mod, "BatchNorm2d", TransNorm2d)
            patched.append(name)
    return patched


def unpatch_reference() -> None:
    while _saved:
        mod, n, orig = _saved.pop()
        setattr(mod, n, orig)
