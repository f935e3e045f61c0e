"""Import the UNMODIFIED reference (fengweie/UDA_CLR) where its tree is present -- TEST INFRASTRUCTURE.

``/root/reference`` exists only in the development container; it does not travel to the GPU box.
Everything here is therefore optional at run time: :func:`available` says whether the tree is
there, and callers (``tests/test_oracle_vs_reference.py``, ``tests/golden/make_golden.py``) skip
when it is not.  Nothing is copied out of the reference: its modules are imported in place.

Recipe: ten import-time-only dependencies of ``utils/Utils.py`` and the trainers
are absent from this image (skimage, matplotlib, albumentations, tensorboardX, pytz, mypath); none is
touched by the hot path, so they are pre-seeded in ``sys.modules`` as mocks.
``gen_prototype_retrify`` hard-codes ``.cuda()`` (utils/Utils.py:188-195) and
``features.reshape(T, stride, 305, 128, 128)`` (:162); on a CPU-only host ``Tensor.cuda`` is
patched to the identity for the duration of the call and a dummy ``features`` is passed.
"""
from __future__ import annotations

import contextlib
import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("UDA_CLR_REFERENCE", "/root/reference")

_STUBS = ["skimage", "skimage.morphology", "skimage.measure", "skimage.transform", "matplotlib",
          "matplotlib.pyplot", "albumentations", "tensorboardX", "pytz", "mypath"]

_utils = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "Utils.py"))


def load_utils():
    """Return the reference's ``utils.Utils`` module (imported once)."""
    global _utils
    if _utils is not None:
        return _utils
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in _STUBS:
        sys.modules.setdefault(name, MagicMock())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import utils.Utils as ref_utils  # noqa: E402  (the reference's own module, in place)
    _utils = ref_utils
    return _utils


@contextlib.contextmanager
def cpu_cuda_shim():
    """Make ``Tensor.cuda()`` the identity while the reference runs on a CPU-only host."""
    import torch
    if torch.cuda.is_available():
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def ref_gen_prototype(pred, feat):
    return load_utils().gen_prototype(pred, feat)


def ref_gen_prototype_src_trg(pred_s, feat_s, pred_t, feat_t):
    return load_utils().gen_prototype_src_trg(pred_s, feat_s, pred_t, feat_t)


def ref_gen_prototype_retrify(oT_before, xt_feature, preds, T, stride):
    """Calls the reference's ``gen_prototype_retrify``; ``xt_feature`` must be ``[stride,C,128,128]``
    because the reference hard-codes the 128x128 feature size (utils/Utils.py:162)."""
    import torch
    assert tuple(xt_feature.shape[-2:]) == (128, 128), "reference hard-codes 128x128 (utils/Utils.py:162)"
    features = torch.zeros(T * stride, 305, 128, 128, dtype=xt_feature.dtype, device=xt_feature.device)
    with cpu_cuda_shim():
        return load_utils().gen_prototype_retrify(oT_before, xt_feature, preds, features, T, stride)


def ref_get_prototype_weight(feat, class_num, prototype):
    return load_utils().get_prototype_weight(feat, class_num, prototype)


def ref_adaptation_factor(m):
    return load_utils().adaptation_factor(m)


def ref_transnorm_class():
    """The reference's TransNorm module class: ``networks.sync_batchnorm.batchnorm.BatchNorm2d`` (imports cleanly)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import networks.sync_batchnorm.batchnorm as ref_bn  # noqa: E402
    return ref_bn.BatchNorm2d
